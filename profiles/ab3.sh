#!/bin/bash
# usage: profiles/ab3.sh lib.so ...  -- floor probe only (empty / C1 / C2 per-kernel us per frame), no suite: quick bisects
for lib in "$@"; do
  RUF_LIB_PATH=$PWD/$lib timeout 200 python profiles/floor_probe.py 2>&1 | grep -E "^(empty|C1|C2)" | python -c "
import sys,ast
out=[]
for l in sys.stdin:
    n,d=l.split(' ',1); d=ast.literal_eval(d)
    out.append('%s raster=%.3f' % (n, d['raster_filter_us_per_frame']))
print('%-40s' % '$lib', '  '.join(out))"
done
