#!/bin/bash
# Round-2 multi-GPU measurements on an 8-GPU box (run through gpurun --gpus 8): the one-process pool with and without graph
# replay of the C2 shape, C4 strong scaling through the pool and through torchrun bench.py --config c4, pool tests.
O=gpurun_out
python tools/bench_pool.py 20 > $O/r02f_pool_n8_eager.json 2> $O/r02f_pool_n8_eager.err; tail -c 600 $O/r02f_pool_n8_eager.json
QASR_GRAPH=all python tools/bench_pool.py 20 > $O/r02f_pool_n8_graph.json 2> $O/r02f_pool_n8_graph.err; tail -c 600 $O/r02f_pool_n8_graph.json
for n in 2 4 8; do
  CUDA_VISIBLE_DEVICES=$(seq -s, 0 $((n-1))) python tools/bench_c4.py 5 > $O/r02f_c4_pool_n$n.json 2> $O/r02f_c4_pool_n$n.err; tail -c 500 $O/r02f_c4_pool_n$n.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --config c4 --no-cpu-baseline > $O/r02f_bench_c4_n$n.json 2> $O/r02f_bench_c4_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r02f_bench_c4_n$n.json").read().strip().splitlines()[-1])
    print("bench c4 n=$n", d["value"], d["ms_per_step"], d["e2e"]["value"], d["scaling"])
except Exception as e: print("ERR", e)
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --no-cpu-baseline > $O/r02f_bench_c2_n8.json 2> $O/r02f_bench_c2_n8.err
tail -c 400 $O/r02f_bench_c2_n8.json | head -c 400; echo
python -m pytest tests/test_gpu_pool.py tests/test_gpu_configs.py -q 2>&1 | tail -4
