#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_consumers.py tests/test_gpu_fused.py -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r2z_pytest.log
timeout 600 python bench.py --workload sd --steps 50 --warmup 5 > gpurun_out/r2z_bench_sd.json 2> gpurun_out/r2z_bench_sd.log; tail -c 900 gpurun_out/r2z_bench_sd.json
timeout 600 python bench.py --workload gaugan --steps 50 --warmup 5 > gpurun_out/r2z_bench_gaugan.json 2> gpurun_out/r2z_bench_gaugan.log; tail -c 900 gpurun_out/r2z_bench_gaugan.json
bash tools/profile_step.sh r2z
cat > /tmp/sattn_once.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from sige_b200 import ops
q = torch.randn(16, 1008, 40, device="cuda").half(); k = torch.randn(16, 4096, 40, device="cuda").half(); v = torch.randn(16, 4096, 40, device="cuda").half()
for _ in range(3):
    ops.sparse_attention(q, k, v, 40 ** -0.5)
torch.cuda.synchronize()
PY
timeout 280 ncu --set full --import-source on --clock-control none -k regex:sparse_attention -s 2 -c 1 -o gpurun_out/r2z_sattn_full -f python /tmp/sattn_once.py > gpurun_out/r2z_ncu_sattn.log 2>&1
ls -la gpurun_out/r2z_*
