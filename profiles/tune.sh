#!/bin/bash
# usage: profiles/tune.sh variants/lib_*.so   -- prints per-stage ms per 64-frame launch for each variant
for lib in "$@"; do
  RUF_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 64 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_launch']
print('$lib', 'fps=%.0f' % d['value'], ' '.join('%s=%.1fus' % (k, v*1e3) for k,v in s.items()))"
done
