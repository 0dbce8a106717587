# ncu --set full of the block encoder (N = 2^20): bash scripts/prof_encode.sh  -> gpurun_out/prof_enc_<TAG>.ncu-rep
python scripts/bench_encode.py > gpurun_out/plain_enc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:encode_block_kernel -s 30 -c 1 -o gpurun_out/prof_enc_${TAG:-a} -f python scripts/bench_encode.py > gpurun_out/ncu_enc.log 2>&1
tail -5 gpurun_out/plain_enc.log
