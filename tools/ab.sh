timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
run() { env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$* $F]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
for rep in 1 2; do
(cd _base && F="" run BASE=1)
F="" run NEW=1
done
mkdir -p gpurun_out
timeout 200 python tools/trace_graph.py --detail > gpurun_out/tg_new.txt 2>&1
timeout 300 python bench.py --ncu > /dev/null 2>&1; ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01i.csv python bench.py --ncu --steps 1 --warmup 1 > gpurun_out/ncu_b.log 2>&1; tail -3 gpurun_out/ncu_b.log
