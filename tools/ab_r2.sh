run() { (cd $1 && timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-reference-cuda --model intree $3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$2 $3]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"); }
for rep in 1 2 3; do run _base BASE ""; run . NEW ""; done
run . NEW8 "--edits 8"
run . NEWref "--model reference"
