#!/usr/bin/env python
"""Instruction / stall-sample share of the phases of the setup and raster kernels (phase = source range between
marker comments of ruf_kernels.cu), from an .ncu-rep with source counters.

    python profiles/ncu_phases.py gpurun_out/prof.ncu-rep raster|setup [lib.so]
"""
import bisect
import collections
import csv
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines

ROOT = ncu_lines.ROOT


def main(rep, which, lib=None):
    lib = lib or os.path.join(ROOT, "realtime_urdf_filter_b200", "libruf_b200.so")
    src = open(os.path.join(ROOT, "realtime_urdf_filter_b200", "csrc", "ruf_kernels.cu")).read().splitlines()

    def find(t, start=0):
        for i, l in enumerate(src):
            if i >= start and t in l:
                return i + 1
        raise KeyError(t)

    if which == "raster":
        k = find("ruf_raster_filter_kernel(Dims d")
        marks = [("inlined helpers", 1), ("prologue / ring start / z clear", k), ("big list: classify", find("per-frame big list, part 1", k)),
                 ("pass control, block maxima", find("const uint32_t nbatches", k)), ("claim + wait for chunk", find("for (;;) {", k)),
                 ("phase 1: record -> tile bbox", find("---- phase 1: one lane per record", k)), ("depth cull", find("if (pass) {", find("---- phase 1", k))),
                 ("phase 1: edge set-up", find("if (kind == 1) {", k)), ("arrive / refill", find("if (lane == 0) {", find("if (kind == 1) {", k))),
                 ("wide records", find("Wide records (more than kMaxUnits units", k)), ("unit table", find("---- phase 2: the units", k)),
                 ("unit loop", find("for (int base = 0; base < items; base += 32) {", k)), ("last barrier, z reload", find("every record of the tile has been rasterised", k) - 4),
                 ("big list: apply", find("per-frame big list, part 2", k)), ("fragment stage", find("---- fused fragment stage", k)),
                 ("(after)", find("Forward kinematics on the device", k))]
        kernel, mangled = "ruf_raster_filter", "ruf_raster_filter_kernelILi1ELb0ELi1"     # the 16UC1 throughput instantiation
    else:
        k = find("ruf_setup_bin_kernel(Model m")
        marks = [("inlined helpers", 1), ("xform", find("__device__ __forceinline__ V4 xform")), ("clip helpers", find("__device__ __forceinline__ float plane_dist")),
                 ("to_window", find("__device__ __forceinline__ bool to_window")), ("setup_window_tri", find("__device__ __forceinline__ bool setup_window_tri")),
                 ("push_big / clip_and_emit", find("__device__ __forceinline__ void push_big")), ("(pose kernel)", find("ruf_pose_kernel(const double")),
                 ("vertex_stage", find("__device__ __forceinline__ uint4 vertex_stage")), ("tri_may_touch", find("__device__ __forceinline__ bool tri_may_touch")),
                 ("CTA prologue", k), ("frame top (prefetch, cull bits)", find("for (int f = f0; f < f1; ++f) {", k)), ("P1 vertices", find("---- P1: vertex stage", k)),
                 ("barrier, matrix staging", find("cp_async_wait_all();", k)), ("P2 classify + compact", find("---- P2: per triangle", k)),
                 ("S3 clip loop", find("---- S3 for the rare triangles", k)), ("P3 set-up", find("---- P3 + P4: one lane per SURVIVOR", k)),
                 ("P4 tile reservation + writes", find("// P4.", k)), ("(after)", find("K4: tile rasteriser"))]
        kernel, mangled = "ruf_setup_bin", "ruf_setup_bin"
    lt = ncu_lines.line_table(lib, mangled)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[0].startswith("0x")][:len(lt)]
    if len(body) != len(lt):
        print(f"warning: {len(body)} profiled instructions vs {len(lt)} in {lib}", file=sys.stderr)
    col = {n: i for i, n in enumerate(hdr)}
    starts = [m[1] for m in marks]
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
    for i, r in enumerate(body):
        name = marks[max(bisect.bisect_right(starts, lt[i]) - 1, 0)][0]
        agg[name][0] += float(r[col["# Samples"]] or 0)
        agg[name][1] += float(r[col["Instructions Executed"]] or 0)
        agg[name][2] += float(r[col["stall_barrier"]] or 0) if "stall_barrier" in col else 0
    ts = sum(v[0] for v in agg.values()) or 1
    ti = sum(v[1] for v in agg.values()) or 1
    print(f"== {kernel}: {ti / 1e6:.1f} M warp instructions, {ts:.0f} stall samples")
    for n, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:34s} {v[1] / ti * 100:5.1f}% instructions {v[0] / ts * 100:5.1f}% samples (barrier {v[2] / ts * 100:4.1f}%)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
