mkdir -p gpurun_out
KSPLIT=1 SIGE_TC5_WIDE=0 timeout 200 python tools/trace_engine.py > gpurun_out/te_n64.txt 2>&1
KSPLIT=1 SIGE_TC5_WIDE=1 timeout 200 python tools/trace_engine.py > gpurun_out/te_n128.txt 2>&1
grep "down.4.block.1.conv1\|mid.block_1.conv1\|down.0.block.1.conv1" gpurun_out/te_n64.txt gpurun_out/te_n128.txt
timeout 600 python -m pytest tests/test_gpu_conv.py tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$* $F]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
for rep in 1 2; do
(cd _base && F="" run BASE=1)
F="" run SIGE_TC5_PF=1 SIGE_TC5_DEEP=0
done
