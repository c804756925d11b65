#!/bin/bash
# Regression run on the GPU box: all GPU tests (one process per file) + smoke + a short bench.
set +e
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1; echo "build rc=$?"
run() { name=$1; shift; timeout 900 "$@" > $O/$name.log 2>&1; rc=$?; echo "== $name rc=$rc"; tail -n ${TAILN:-6} $O/$name.log; }
TAILN=25 run t_kernels python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider
TAILN=25 run t_path python -m pytest tests/test_gpu_path.py -q -p no:cacheprovider
run smoke python __graft_entry__.py --smoke
if [ "$1" != "nobench" ]; then
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 ${BENCH_ARGS} > $O/bench.log 2> $O/bench.err; echo "== bench rc=$?"; tail -3 $O/bench.err
python tools/show_bench.py $O/bench.log
fi
