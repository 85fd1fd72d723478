#!/bin/bash
# A/B of environment knobs: tools/ab_env.sh "VAR=a" "VAR=b" ...   (run on the GPU box); prints per-kernel ms
for kv in "$@"; do
  env $kv python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']
print('$kv', 'value=%.0f ms=%.3f' % (d['value'], d['ms_per_step']), ' '.join('%s=%.3f' % (n, v) for n, v in k.items()))"
done
