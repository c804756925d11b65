#!/bin/bash
# Round-2 end-of-round measurements on one B200: every config through bench.py, a sustained 400-step run, the FP8 variant, the
# reference arm, smoke, and the ncu launch list of the default command.
O=gpurun_out
T=r02m
python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_c2.json 2> $O/${T}_bench_c2.err; python tools/show_bench.py $O/${T}_bench_c2.json | head -8
python bench.py --steps 400 --warmup 5 --no-cpu-baseline > $O/${T}_bench_c2_sustained400.json 2>/dev/null; python tools/show_bench.py $O/${T}_bench_c2_sustained400.json | head -4
for c in c1 c3 c4 c5; do python bench.py --config $c --no-cpu-baseline > $O/${T}_bench_$c.json 2>/dev/null; echo "== $c"; python tools/show_bench.py $O/${T}_bench_$c.json | head -1; done
python bench.py --quantize fp8_per_row --no-cpu-baseline > $O/${T}_bench_fp8_per_row.json 2>/dev/null; echo "== fp8_per_row"; python tools/show_bench.py $O/${T}_bench_fp8_per_row.json | head -1
python bench.py --impl reference --steps 1 --warmup 0 > $O/${T}_bench_reference.json 2>/dev/null; tail -c 400 $O/${T}_bench_reference.json; echo
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file $O/${T}_launches_all.csv $CMD > /dev/null 2>&1; echo "launch list rc=$?"
