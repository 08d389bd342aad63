run() { env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline $F 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('[$* $F]', round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))"; }
F="" run SIGE_TC5_MIN_TAPS=3
F="" run SIGE_TC5_MIN_TAPS=5
F="" run SIGE_TC5_MIN_TAPS=9
F="" run SIGE_TC5_MIN_TAPS=18
F="" run SIGE_TC5_MIN_TAPS=3
F="--no-fuse-shortcut" run SIGE_TC5_MIN_TAPS=3
