#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "stride2" 2>&1 | tail -5 | tee gpurun_out/r2w_pytest_s2.log
timeout 300 python tools/microbench_s2.py 2>&1 | tee gpurun_out/r2w_microbench_s2.log
run() { env $1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-reference-cuda $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1 $2]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline'].get('sum_in_graph_us'))"; }
for rep in 1 2; do for arm in "SIGE_TC5_S2=0" "SIGE_TC5_S2=1"; do run "$arm" ""; done; done 2>&1 | tee gpurun_out/r2w_ab_s2.log
for arm in "SIGE_TC5_S2=0" "SIGE_TC5_S2=1"; do run "$arm" "--edits 8"; run "$arm" "--ratio 0.15"; done 2>&1 | tee -a gpurun_out/r2w_ab_s2.log
timeout 400 python tools/trace_graph.py --detail > gpurun_out/r2w_timeline_s2.txt 2>&1
