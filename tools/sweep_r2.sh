#!/bin/bash
# BASELINE.json configs[4] on ONE GPU: edit-ratio sweep x batch of edits per step.  Prints one line per run.
for E in 1 8; do for R in 0.01 0.05 0.15 0.30; do
  timeout 400 python bench.py --no-cpu-baseline --no-reference-cuda --steps 50 --warmup 5 --edits $E --ratio $R 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline'] or {}
print('edits=%d ratio=%s  %.3f ms/step  %.0f edit-steps/s  e2e %.3f ms  in-graph tc5: %.0f GB/s (%.1f%% HBM) %.1f TFLOP/s' % (d['config']['edits_per_gpu'], '$R', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], r.get('achieved') or 0, 100*(r.get('frac') or 0), (r.get('tensor') or {}).get('achieved_tflops') or 0))"
done; done
