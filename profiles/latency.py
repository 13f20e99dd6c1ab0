#!/usr/bin/env python
"""Single-frame latency of the reference-facing call (ruf_filter: host buffers in, host buffers out, synchronous),
the way RealtimeURDFFilter::filter_callback uses the path at 30 Hz.  C2 scene, 16UC1 and 32FC1, pinned and pageable."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

# optional argument: c1 | c2 | c3 | c5 (bench.py's configurations; default c2)
cfg_name = sys.argv[1] if len(sys.argv) > 1 else "c2"
import bench
sc = bench.make_scene(bench.CONFIGS[cfg_name])
proj, _, _ = sc.proj()
views, pms = sc.frames(list(range(32)))
lib = ruf.load()
out = {}
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    for enc, dt, code in (("16UC1", torch.int16, ruf.ENC_U16_MM), ("32FC1", torch.float32, ruf.ENC_F32_M)):
        for pinned in (True, False):
            h_in = torch.full((sc.height, sc.width), 1500 if enc == "16UC1" else 1.5, dtype=dt)
            h_out = torch.empty_like(h_in)
            h_mask = torch.empty((sc.height, sc.width), dtype=torch.uint8)
            if pinned:
                h_in, h_out, h_mask = h_in.pin_memory(), h_out.pin_memory(), h_mask.pin_memory()
            ts = []
            for k in range(232):
                v, pm = views[k % 32], pms[k % 32]
                t0 = time.perf_counter()
                rc = lib.ruf_filter(ctx._h, h_in.data_ptr(), code, proj.ctypes.data, v.ctypes.data, pm.ctypes.data,
                                    sc.max_diff, sc.replace_value, h_out.data_ptr(), h_mask.data_ptr())
                ts.append(time.perf_counter() - t0)
                assert rc == 0
            ts = np.array(ts[32:]) * 1e6
            out[f"{enc} {'pinned' if pinned else 'pageable'}"] = dict(median_us=round(float(np.median(ts)), 1), p99_us=round(float(np.percentile(ts, 99)), 1))
            print(enc, "pinned" if pinned else "pageable", out[f"{enc} {'pinned' if pinned else 'pageable'}"], flush=True)
    # a caller that alternates between two sets of buffers (double buffering): the raster kernel's arguments are retargeted
    # in the instantiated graph on every call
    sets = []
    for _ in range(2):
        a_in = torch.full((sc.height, sc.width), 1500, dtype=torch.int16).pin_memory()
        sets.append((a_in, torch.empty_like(a_in).pin_memory(), torch.empty((sc.height, sc.width), dtype=torch.uint8).pin_memory()))
    ts = []
    for k in range(232):
        v, pm = views[k % 32], pms[k % 32]
        b_in, b_out, b_mask = sets[k & 1]
        t0 = time.perf_counter()
        rc = lib.ruf_filter(ctx._h, b_in.data_ptr(), ruf.ENC_U16_MM, proj.ctypes.data, v.ctypes.data, pm.ctypes.data,
                            sc.max_diff, sc.replace_value, b_out.data_ptr(), b_mask.data_ptr())
        ts.append(time.perf_counter() - t0)
        assert rc == 0
    ts = np.array(ts[32:]) * 1e6
    out["16UC1 pinned, alternating buffers"] = dict(median_us=round(float(np.median(ts)), 1), p99_us=round(float(np.percentile(ts, 99)), 1))
    print("16UC1 pinned, alternating buffers", out["16UC1 pinned, alternating buffers"], flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency_%s.json" % cfg_name), "w"), indent=1)
