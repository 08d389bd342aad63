#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -x -q -m gpu -s -k "without_a_host_sync or next_edit or headroom" 2>&1 | grep -v Warning | tail -12 | tee gpurun_out/r2s_pytest_async.log
timeout 600 python -m pytest tests/test_consumers.py -x -q -m gpu -s 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/r2s_pytest_consumers.log
timeout 600 python bench.py --workload sd --steps 50 --warmup 5 > gpurun_out/r2s_bench_sd.json 2> gpurun_out/r2s_bench_sd.log
tail -c 1200 gpurun_out/r2s_bench_sd.json
