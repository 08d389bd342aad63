timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_engine.py -m gpu -q -x 2>&1 | tail -3
run() { env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$* $F]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
for rep in 1 2; do
(cd _base && F="" run BASE=1)
F="" run NEW=1
done
