#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_consumers.py -x -q -m gpu -s 2>&1 | grep -v Warn | tail -5 | tee gpurun_out/r3a_pytest.log
timeout 600 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r3a_bench_gaugan.json 2> gpurun_out/r3a_bench_gaugan.log; tail -c 600 gpurun_out/r3a_bench_gaugan.json
timeout 600 python bench.py --workload sd --steps 50 --warmup 5 > gpurun_out/r3a_bench_sd.json 2> gpurun_out/r3a_bench_sd.log; tail -c 600 gpurun_out/r3a_bench_sd.json
timeout 900 python bench.py > gpurun_out/r3a_bench_default.json 2> gpurun_out/r3a_bench_default.log; cat gpurun_out/r3a_bench_default.json
