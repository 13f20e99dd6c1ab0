#!/bin/bash
# e2e A/B: frames per ruf_filter_batch_host call and pipeline chunk size (RUF_HOST_CHUNK)
for cfg in "256 32" "256 16" "256 64" "1024 32" "1024 64" "64 16" "16 4"; do set -- $cfg
RUF_HOST_CHUNK=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames $1 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('e2e_frames=$1 chunk=$2', 'e2e=%.0f' % d['e2e']['value'], 'fps=%.0f' % d['value'])"; done
