#!/bin/bash
# ncu evidence for the bench command: launch list (gpu__time_duration) + one --set full capture of the top kernels.
# usage: gpu_ncu.sh <tag> [kernel regex] [skip] [count]
set +e
mkdir -p gpurun_out
O=gpurun_out
TAG=${1:-r01}
KRE=${2:-gemm_tc_kernel}
SKIP=${3:-0}
CNT=${4:-7}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > $O/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-381} -c 300 --csv --log-file $O/${TAG}_launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > $O/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT -o $O/${TAG}_prof $CMD > $O/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 $O/ncu_full.log
