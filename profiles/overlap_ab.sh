#!/bin/bash
# A/B of the sliced two-stream launch (RUF_SLICE_FRAMES) at several batch sizes
for cfg in "512 2 0" "512 2 256" "1024 1 0" "1024 1 256" "2048 1 0"; do set -- $cfg
RUF_SLICE_FRAMES=$3 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-frames 64 --batch $1 --ring $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_launch']
print('batch=$1 ring=$2 slice=$3', 'fps=%.0f' % d['value'], ' '.join('%s=%.1fus' % (k, v*1e3) for k,v in s.items()), 'launches', d['gpu_launches'])"; done
