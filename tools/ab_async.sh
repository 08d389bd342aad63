(timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_fused.py -q -x > gpurun_out/r2j_pytest.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/r2j_pytest.log)
run() { env $1 timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-reference-cuda $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$1 $2]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4), d['roofline'].get('sum_in_graph_us'))"; }
for rep in 1 2; do for arm in "SIGE_TC5_ASYNC_GATHER=0" "SIGE_TC5_ASYNC_GATHER=1"; do run "$arm" ""; run "$arm" "--edits 8"; done; done
run "SIGE_TC5_ASYNC_GATHER=0" "--ratio 0.30"; run "SIGE_TC5_ASYNC_GATHER=1" "--ratio 0.30"
