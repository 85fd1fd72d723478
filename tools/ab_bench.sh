#!/bin/bash
# A/B timing of library variants: tools/ab_bench.sh lib1.so lib2.so ...   (run on the GPU box)
for lib in "$@"; do
  AVZ_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra-configs 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']
print('$lib', 'value=%.0f ms=%.3f' % (d['value'], d['ms_per_step']), ' '.join('%s=%.3f' % (n.split(' ')[0], v) for n, v in k.items()))"
done
