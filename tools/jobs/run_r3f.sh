#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_consumers.py -x -q -m gpu -s 2>&1 | grep -v Warn | tail -5 | tee gpurun_out/r3f_pytest.log
timeout 600 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r3f_bench_gaugan.json 2> gpurun_out/r3f_bench_gaugan.log; tail -c 700 gpurun_out/r3f_bench_gaugan.json
