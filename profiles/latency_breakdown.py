#!/usr/bin/env python
"""Where does the single-frame latency go?  Per-kernel device time of a ONE-frame launch sequence (events between the
kernels) and the bare copy times of one frame's buffers, next to the end-to-end ruf_filter call (profiles/latency.py)."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

dev = torch.device("cuda:0")
sc = synth.pr2_like_scene()
proj, _, _ = sc.proj()
views, pms = sc.frames(list(range(8)))
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_in = torch.full((1, sc.height, sc.width), 1500, dtype=torch.int16, device=dev)
d_out, d_mask = torch.empty_like(d_in), torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
d_proj, d_view, d_pm = t(proj), t(views), t(pms)
out = {}
stream = torch.cuda.Stream(device=dev)
with ruf.Context(sc.width, sc.height) as ctx, torch.cuda.stream(stream):
    ctx.set_stream(stream.cuda_stream)
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    args = lambda k: (1, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view[k:k + 1].data_ptr(), d_pm[k:k + 1].data_ptr(),
                      sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
    for k in range(8):
        ctx.filter_batch_device(*args(k)); ctx.sync()
    ctx.set_profiling(True); ctx.stage_times(reset=True)
    for k in range(64):
        ctx.filter_batch_device(*args(k % 8)); ctx.sync()
    ms, calls = ctx.stage_times(reset=True)
    out["kernels_us"] = {k: round(v * 1e3 / calls, 2) for k, v in ms.items()}
    ctx.set_profiling(False)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for k in range(64):
        e0.record(stream); ctx.filter_batch_device(*args(k % 8)); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    out["sequence_us_back_to_back"] = round(float(np.median(ts)), 2)
h_in = torch.empty((sc.height, sc.width), dtype=torch.int16).pin_memory()
h_out = torch.empty_like(h_in).pin_memory(); h_mask = torch.empty((sc.height, sc.width), dtype=torch.uint8).pin_memory()
for name, fn in (("h2d_depth_614KB", lambda: d_in[0].copy_(h_in, non_blocking=True)),
                 ("d2h_depth_614KB", lambda: h_out.copy_(d_out[0], non_blocking=True)),
                 ("d2h_mask_307KB", lambda: h_mask.copy_(d_mask[0], non_blocking=True))):
    ts = []
    for _ in range(50):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e6)
    out[name + "_us"] = round(float(np.median(ts[10:])), 1)
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "latency_breakdown.json"), "w"), indent=1)
