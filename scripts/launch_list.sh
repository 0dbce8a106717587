# usage: bash scripts/launch_list.sh <tag> <bench args...>  -> gpurun_out/launches_<tag>.csv (+ plain log)
TAG=$1; shift
CMD="python bench.py $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_ll_$TAG.log 2>&1
tail -c 400 gpurun_out/plain_$TAG.log
