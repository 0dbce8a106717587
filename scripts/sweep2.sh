PC_SCL_THREADS=64 PC_SCL_LSM=4 python -m pytest tests/test_gpu_scl.py -x -q 2>&1 | tail -2
FRAMES=32768 bash scripts/sweep_scl.sh "7 256 9" "7 128 9" "6 128 9" "5 128 9" "6 64 9" "5 64 9" "4 64 9" "3 64 9" "4 128 9" "5 64 7" "4 64 12"
