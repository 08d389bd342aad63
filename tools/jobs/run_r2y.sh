#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -m gpu -k "generic" 2>&1 | tail -4 | tee gpurun_out/r2y_pytest_generic.log
timeout 300 python tools/profile_consumer.py gaugan > gpurun_out/r2y_prof_gaugan.txt 2>&1
timeout 300 python tools/profile_consumer.py sd > gpurun_out/r2y_prof_sd.txt 2>&1
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_conv.py tests/test_gpu_glue.py -x -q -m gpu -k "stride2 or sparse_attention_strided or generic_conv_half" 2>&1 | tail -6 | tee gpurun_out/r2y_memcheck.log
