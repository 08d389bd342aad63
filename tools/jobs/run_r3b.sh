#!/bin/bash
run() { env $1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-reference-cuda 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1]', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['ms_per_step'],4), d['clocks'])"; }
for rep in 1 2 3; do for arm in "SIGE_BENCH_SAMPLE_S=0.01" "SIGE_BENCH_SAMPLE_S=0.1" "SIGE_BENCH_SAMPLE_S=1000"; do run "$arm"; done; done 2>&1 | tee gpurun_out/r3b_e2e_sampler.log
