#!/bin/bash
# Bring-up run on the GPU box: each stage in its own process (a trap in one kernel must not poison the rest).
set +e
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1; echo "build rc=$?"
run() { name=$1; shift; timeout 600 "$@" > $O/$name.log 2>&1; rc=$?; echo "== $name rc=$rc"; tail -n ${TAILN:-15} $O/$name.log; }
run t_simt python -m pytest tests/test_gpu_kernels.py -q -k "simt" -p no:cacheprovider
run t_tc python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05 or bitwise" -p no:cacheprovider
run t_misc python -m pytest tests/test_gpu_kernels.py -q -k "layernorm or attention" -p no:cacheprovider
TAILN=40 run t_path python -m pytest tests/test_gpu_path.py -q -p no:cacheprovider
run smoke python __graft_entry__.py --smoke
TAILN=5 run bench python bench.py --steps 5 --warmup 3
