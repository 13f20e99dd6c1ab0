#!/bin/bash
# A/B of library variants (variants/lib_*.so, built with RUF_EXTRA_NVCC) against the default library on one box:
# bench stage times (C2), the floor probe (empty / C1 / C2), then the whole GPU suite against every variant.
mkdir -p gpurun_out
bash profiles/ab.sh realtime_urdf_filter_b200/libruf_b200.so variants/lib_*.so realtime_urdf_filter_b200/libruf_b200.so > gpurun_out/reentry_ab.txt 2>&1
cat gpurun_out/reentry_ab.txt
for lib in realtime_urdf_filter_b200/libruf_b200.so variants/lib_*.so; do
  echo "== $lib" >> gpurun_out/reentry_probe.txt
  RUF_LIB_PATH=$PWD/$lib timeout 100 python profiles/floor_probe.py 2>&1 | tail -3 >> gpurun_out/reentry_probe.txt
done
cat gpurun_out/reentry_probe.txt
for lib in variants/lib_*.so; do
  echo "== $lib" >> gpurun_out/reentry_parity.txt
  RUF_LIB_PATH=$PWD/$lib timeout 120 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 >> gpurun_out/reentry_parity.txt
done
cat gpurun_out/reentry_parity.txt
