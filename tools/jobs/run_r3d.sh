#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r3d_smoke.log
timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r3d_pytest_gpu.log
