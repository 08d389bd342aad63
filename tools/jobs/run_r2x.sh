#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2x_pytest_gpu.log
run() { env $1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-reference-cuda $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1 $2]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline'].get('sum_in_graph_us'), d['roofline'].get('frac'))"; }
for a in "" "--edits 8" "--ratio 0.05" "--ratio 0.15" "--ratio 0.30" "--edits 8 --ratio 0.30"; do run "A=0" "$a"; done 2>&1 | tee gpurun_out/r2x_sweep.log
timeout 600 python bench.py --workload sd --steps 50 --warmup 5 > gpurun_out/r2x_bench_sd.json 2> gpurun_out/r2x_bench_sd.log; tail -c 700 gpurun_out/r2x_bench_sd.json
timeout 600 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r2x_bench_gaugan.json 2> gpurun_out/r2x_bench_gaugan.log; tail -c 700 gpurun_out/r2x_bench_gaugan.json
