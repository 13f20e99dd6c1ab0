#!/bin/bash
# usage: profiles/ab.sh [lib.so ...]   -- A/B of library variants (built with RUF_LIB_PATH / RUF_EXTRA_NVCC, see
# realtime_urdf_filter_b200/build.py) on the bench workload: frames/s and per-stage us per 256-frame launch
for lib in "$@"; do
  RUF_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-frames 64 --batch 1024 --ring 1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_launch']
print('$lib', 'fps=%.0f' % d['value'], ' '.join('%s=%.1fus' % (k, v*1e3) for k,v in s.items()), 'e2e=%.0f' % d['e2e']['value'])"
done
