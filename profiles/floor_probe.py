#!/usr/bin/env python
"""Where is the floor?  Per-kernel time per frame (events between kernels, batch 1024) for an EMPTY model (only the
background quad: every raster CTA takes the register-only path), C1 (example.urdf: big-list records only) and C2.
The difference C2 - empty is what the binned records cost; `empty` is the fixed per-tile cost of the path."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

dev = torch.device("cuda:0")
B = int(os.environ.get("PROBE_BATCH", "1024"))
ex = synth.example_scene()
empty = synth.Scene("empty", 640, 480, ex.P, [], [], np.zeros((0, 9), np.float32), np.zeros(0, np.uint32),
                    -1, (0, 0, 0), ex.cam_R)
out = {}
for name, sc in (("empty", empty), ("C1", ex), ("C2", synth.pr2_like_scene())):
    proj, _, _ = sc.proj()
    views, pms = sc.frames(list(range(B)))
    depth = np.random.default_rng(1).integers(300, 6000, (B, sc.height, sc.width)).astype(np.uint16)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in, d_proj, d_view = t(depth.view(np.int16)), t(proj), t(views)
    d_pm = t(pms) if sc.n_parts else torch.zeros(16, dtype=torch.float64, device=dev)
    d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx, torch.cuda.stream(stream):
        ctx.set_stream(stream.cuda_stream)
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.reserve(B)
        args = (B, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        for _ in range(3):
            ctx.filter_batch_device(*args); ctx.sync()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            ctx.filter_batch_device(*args)
        e1.record(stream)
        ctx.sync()
        total_us = e0.elapsed_time(e1) * 1e3 / (n * B)
        ctx.set_profiling(True); ctx.stage_times(reset=True)
        for _ in range(n):
            ctx.filter_batch_device(*args)
        ctx.sync()
        ms, calls = ctx.stage_times(reset=True)
        st = ctx.stats()
    out[name] = dict(us_per_frame=round(total_us, 3), **{k + "_us_per_frame": round(v * 1e3 / (calls * B), 3) for k, v in ms.items()},
                     big_per_frame=st["big_tris"] / B, refs_per_frame=st["binned_refs"] / B)
    print(name, out[name], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "floor_probe.json"), "w"), indent=1)
