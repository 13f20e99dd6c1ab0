#!/usr/bin/env python
"""Per-CTA phase timestamps of the raster kernel for ONE single-frame call (debug variant built with -DRUF_X_TIMELINE:
profiles/build_variant.sh tl - -DRUF_X_TIMELINE; RUF_LIB_PATH=variants/lib_tl.so python profiles/timeline_probe.py)."""
import ctypes, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

sc = synth.pr2_like_scene()
proj, _, _ = sc.proj()
views, pms = sc.frames(list(range(8)))
lib = ruf.load()
lib.ruf_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
rng = np.random.default_rng(3)
names = ["start", "prologue+sync", "classify+front pass", "block maxima", "back pass+wide", "cluster wait", "z reload+big list", "fragment"]
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    h_in = torch.from_numpy(rng.integers(400, 4000, (sc.height, sc.width)).astype(np.int16)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    h_mask = torch.empty((sc.height, sc.width), dtype=torch.uint8).pin_memory()
    tl = np.zeros(8192 * 8, dtype=np.uint64)
    bl = np.zeros(1024 * 8 * 16 * 4, dtype=np.uint64)
    lib.ruf_debug_batchlog.argtypes = [ctypes.c_void_p]
    lib.ruf_debug_timeline1.argtypes = [ctypes.c_void_p]
    t1 = np.zeros(1024 * 64, dtype=np.uint64)
    for k in range(12):
        lib.ruf_debug_timeline(tl.ctypes.data, tl.size)          # clears
        lib.ruf_debug_batchlog(bl.ctypes.data)
        lib.ruf_debug_timeline1(t1.ctypes.data)
        rc = lib.ruf_filter(ctx._h, h_in.data_ptr(), ruf.ENC_U16_MM, proj.ctypes.data, views[k % 8].ctypes.data, pms[k % 8].ctypes.data,
                            sc.max_diff, sc.replace_value, h_out.data_ptr(), h_mask.data_ptr())
        assert rc == 0
    lib.ruf_debug_timeline(tl.ctypes.data, tl.size)
    lib.ruf_debug_batchlog(bl.ctypes.data)
    lib.ruf_debug_timeline1(t1.ctypes.data)
# setup kernel: per warp marks
w = t1.reshape(-1, 8).astype(np.int64)
w = w[w[:, 0] > 0]
s0 = w[:, 0].min()
live = w[w[:, 6] > 0]
print("setup kernel: warps", len(w), "live", len(live), "span us", (w.max() - s0) / 1e3, "last start", (w[:, 0].max() - s0) / 1e3)
print("slowest warps: start | cull bytes | loads+sync | P1+B1 | P2 | S3 clip | P3+P4 | end")
for i in np.argsort(-live[:, 6])[:10]:
    r = live[i]
    print(" ".join("%7.2f" % x for x in [(r[0] - s0) / 1e3] + [(r[j] - r[j - 1]) / 1e3 for j in range(1, 7)]), " end %.2f" % ((r[6] - s0) / 1e3))
t = tl.reshape(-1, 8).astype(np.int64)
t = t[t[:, 0] > 0]
t0 = t[:, 0].min()
full = t[t[:, 7] > 0]
print("CTAs", len(t), "non-flat", len(full), "kernel span us", (full[:, 7].max() - t0) / 1e3, "last start us", (t[:, 0].max() - t0) / 1e3)
order = np.argsort(-full[:, 7])[:12]
print("slowest CTAs: start, then the duration of each phase (us):", names)
for i in order:
    r = full[i]
    ph = [(r[0] - t0) / 1e3] + [((r[j] - r[j - 1]) / 1e3 if r[j] and r[j - 1] else 0.0) for j in range(1, 8)]
    print(" ".join("%7.2f" % x for x in ph), " end %.2f" % ((r[7] - t0) / 1e3))

# the batch log of the slowest CTA: per warp, per batch (claim -> ready -> phase 1 -> done), units dealt
allt = tl.reshape(-1, 8).astype(np.int64)
slow = int(np.argmax(allt[:, 7]))
if slow < 1024:
    b = bl.reshape(1024, 8, 16, 4)[slow]
    print("batch log of CTA", slow, "(us since kernel start: claim, wait for chunk, phase 1, units; pass, units)")
    for w in range(8):
        for k in range(16):
            q = b[w, k]
            if not q[0]:
                continue
            t3 = int(q[3]) & ((1 << 48) - 1); meta = int(q[3]) >> 48
            print("  warp %d batch %2d  claim %6.2f  wait %5.2f  phase1 %5.2f  units %5.2f   pass %d items %d" % (
                w, k, (int(q[0]) - t0) / 1e3, (int(q[1]) - int(q[0])) / 1e3, (int(q[2]) - int(q[1])) / 1e3, t3 / 1e3,
                meta >> 15, meta & 0x7fff))
