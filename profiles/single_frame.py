#!/usr/bin/env python
"""A few single-frame ruf_filter calls (C2 scene, 16UC1, pinned buffers) for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 80 --csv python profiles/single_frame.py"""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
sc = synth.pr2_like_scene()
proj, _, _ = sc.proj()
views, pms = sc.frames(list(range(8)))
lib = ruf.load()
rng = np.random.default_rng(3)
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    h_in = torch.from_numpy(rng.integers(400, 4000, (sc.height, sc.width)).astype(np.int16)).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    h_mask = torch.empty((sc.height, sc.width), dtype=torch.uint8).pin_memory()
    for k in range(n):
        rc = lib.ruf_filter(ctx._h, h_in.data_ptr(), ruf.ENC_U16_MM, proj.ctypes.data, views[k % 8].ctypes.data, pms[k % 8].ctypes.data,
                            sc.max_diff, sc.replace_value, h_out.data_ptr(), h_mask.data_ptr())
        assert rc == 0
print("ok", int(h_mask.sum()))
