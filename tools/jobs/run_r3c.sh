#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r3c_pytest_gpu.log
timeout 600 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r3c_bench_gaugan.json 2> gpurun_out/r3c_bench_gaugan.log; tail -c 650 gpurun_out/r3c_bench_gaugan.json
timeout 300 python tools/profile_consumer.py gaugan > gpurun_out/r3c_prof_gaugan.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
