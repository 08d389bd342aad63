#!/bin/bash
# one gpurun call: sparse-attention parity tests, microbench vs SDPA, SD consumer test, SD full-size bench with ours / SDPA
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py -x -q -m gpu -k "sparse_attention" 2>&1 | tail -15 | tee gpurun_out/r2r_pytest_sattn.log
timeout 300 python tools/microbench_sparse_attention.py 2>&1 | tee gpurun_out/r2r_microbench_sattn.log
timeout 600 python -m pytest tests/test_consumers.py -x -q -m gpu -s 2>&1 | grep -v Warning | tail -8 | tee gpurun_out/r2r_pytest_consumers.log
for arm in 1 0; do
  SIGE_SPARSE_ATTENTION=$arm timeout 600 python bench.py --workload sd --steps 50 --warmup 5 > gpurun_out/r2r_bench_sd_sattn$arm.json 2> gpurun_out/r2r_bench_sd_sattn$arm.log
  tail -c 1500 gpurun_out/r2r_bench_sd_sattn$arm.json
done
