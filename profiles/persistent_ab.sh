#!/bin/bash
# Round-2 opener (prepared at the end of round 1, NOT yet run on a GPU): the raster kernel as persistent CTAs.
#   here (no GPU):   bash profiles/persistent_ab.sh build      -> variants/lib_persist{1,2}.so
#   under gpurun:    bash profiles/persistent_ab.sh            -> A/B on C2, the floor probe and the whole GPU suite per variant
# RUF_PERSISTENT=1: gridDim = SMs x 5 CTAs (made coprime with the tile count) walk the (frame, tile) items;
# RUF_PERSISTENT=2: additionally the record counts of the next item are requested one item ahead.
set -u
if [ "${1:-}" = build ]; then
  mkdir -p variants
  for v in 1 2; do
    RUF_LIB_PATH=$PWD/variants/lib_persist$v.so RUF_EXTRA_NVCC="-DRUF_PERSISTENT=$v" python -m realtime_urdf_filter_b200.build --force
  done
  touch variants/*.so
  exit 0
fi
mkdir -p gpurun_out
bash profiles/ab.sh realtime_urdf_filter_b200/libruf_b200.so variants/lib_persist*.so realtime_urdf_filter_b200/libruf_b200.so 2>&1 | tee gpurun_out/persist_ab.txt
for lib in variants/lib_persist*.so; do
  echo "== $lib" | tee -a gpurun_out/persist_probe.txt
  RUF_LIB_PATH=$PWD/$lib timeout 100 python profiles/floor_probe.py 2>&1 | tail -3 | tee -a gpurun_out/persist_probe.txt
  RUF_LIB_PATH=$PWD/$lib timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee -a gpurun_out/persist_probe.txt
done
