#!/bin/bash
# usage: profiles/ab2.sh lib.so ...  -- like ab.sh, plus the whole GPU suite against every variant first (a variant that is
# not bit-exact is not timed) and the floor probe (empty model / C1 / C2 per-kernel us per frame)
for lib in "$@"; do
  echo "== $lib"
  case "$lib" in *lib_r1*) ;; *)
  if ! RUF_LIB_PATH=$PWD/$lib timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -1 | tee /dev/stderr | grep -q " passed"; then echo "   (suite failed: not timed)"; continue; fi ;;
  esac
  RUF_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --e2e-frames 64 --batch 1024 --ring 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms_per_launch']
print('   bench', 'fps=%.0f' % d['value'], ' '.join('%s=%.1fus' % (k, v*1e3) for k,v in s.items()))"
  RUF_LIB_PATH=$PWD/$lib timeout 200 python profiles/floor_probe.py 2>&1 | grep -E "^(empty|C1|C2)" | python -c "
import sys,ast
for l in sys.stdin:
    n,d=l.split(' ',1); d=ast.literal_eval(d)
    print('   %-5s total=%.3f setup=%.3f raster=%.3f us/frame' % (n, d['us_per_frame'], d['setup_bin_us_per_frame'], d['raster_filter_us_per_frame']))"
done
