#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_glue.py tests/test_consumers.py -x -q -m gpu -k "spade or consumer" -s 2>&1 | grep -v Warn | tail -6 | tee gpurun_out/r3g_pytest.log
timeout 300 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r3g_bench_gaugan.json 2> gpurun_out/r3g_bench_gaugan.log; tail -c 700 gpurun_out/r3g_bench_gaugan.json
