timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$* $F]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
for rep in 1 2; do
(cd _base && F="" run BASE=1)
F="" run NEW=1
done
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.log; tail -1 gpurun_out/bench_final.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.log; tail -1 gpurun_out/bench_ref.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
