#!/bin/bash
# usage: profiles/build_variant.sh <name> [<commit>] [nvcc flags...]
# Builds variants/lib_<name>.so from the CURRENT host/API sources and, when <commit> is given (not "-"), the kernel files
# (csrc/ruf_kernels.cu, csrc/ruf_device.cuh) of that commit: an A/B partner that still speaks today's C ABI.
set -eu
name=$1; commit=${2:--}; shift; shift || true
root=$(cd "$(dirname "$0")/.." && pwd)
tmp=$(mktemp -d)
cp -r "$root/realtime_urdf_filter_b200" "$root/include" "$tmp/"
rm -f "$tmp"/realtime_urdf_filter_b200/*.so
if [ "$commit" != "-" ]; then
  files="ruf_kernels.cu ruf_device.cuh"
  [ "${ALLSRC:-0}" = 1 ] && files="ruf_kernels.cu ruf_device.cuh ruf_api.cu"      # when the C ABI did not change since <commit>
  for f in $files; do git -C "$root" show "$commit:realtime_urdf_filter_b200/csrc/$f" > "$tmp/realtime_urdf_filter_b200/csrc/$f"; done
fi
mkdir -p "$root/variants"
(cd "$tmp" && RUF_LIB_PATH="$root/variants/lib_$name.so" RUF_EXTRA_NVCC="$*" python -m realtime_urdf_filter_b200.build --force > /dev/null)
rm -rf "$tmp"; touch "$root/variants/lib_$name.so"; echo "variants/lib_$name.so"
