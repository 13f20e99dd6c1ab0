import os, sys
os.environ["RUF_SETUP_FRAMES_FORCE"] = "2"      # exercise the frame loop of the setup kernel
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import helpers, realtime_urdf_filter_b200 as ruf
from realtime_urdf_filter_b200 import synth
for name, kw in (("pr2_small", {}), ("example", {})):
    sc = helpers.scene(name)
    proj, _, _ = sc.proj()
    fr = helpers.make_frame(sc, 2, "u16")
    want_d, want_m, _ = helpers.oracle_filter(sc, fr)
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        d, m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
    assert np.array_equal(d, want_d) and np.array_equal(m, want_m)
    print(name, "ok")
# 32FC1 at full size: the 256-bit loads / stores of the float path, through ruf_filter's page-locked staging and the graph
sc = helpers.scene("pr2_small"); proj, _, _ = sc.proj()
fr = helpers.make_frame(sc, 4, "f32")
want_d, want_m, _ = helpers.oracle_filter(sc, fr)
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    d, m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
assert np.array_equal(d.view(np.uint32), want_d.view(np.uint32)) and np.array_equal(m, want_m)
print("f32 ok")
sc = synth.pr2_like_scene(100, 75, n_tris=3000, name="odd")
proj, _, _ = sc.proj(); fr = helpers.make_frame(sc, 1, "f32")
with ruf.Context(100, 75) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    ctx.reserve(1, 2, 64)
    d, m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
print("odd ok")
# a 5-frame device batch through the frame loop (runs of 2, 2, 1) and the two-pass raster (depth cull)
import torch
sc = helpers.scene("pr2_small"); proj, _, _ = sc.proj()
frames = [helpers.make_frame(sc, k, "u16") for k in (0, 5, 9, 14, 21)]
dev = torch.device("cuda:0"); t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
d_in = t(np.stack([f["depth"] for f in frames]).view(np.int16)); d_out = torch.empty_like(d_in)
d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
d_proj, d_view, d_pm = t(proj), t(np.stack([f["view"] for f in frames])), t(np.stack([f["pm"] for f in frames]))
torch.cuda.synchronize()
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    ctx.filter_batch_device(5, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                            sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
    ctx.sync()
for i, fr in enumerate(frames):
    want_d, want_m, _ = helpers.oracle_filter(sc, fr)
    assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d) and np.array_equal(d_mask[i].cpu().numpy(), want_m)
print("batch ok")
# round 2: the single-frame CUDA-graph path (pinned host buffers), the packed mask, and whichever raster kernel variant
# RUF_MULTIPASS selects (pr2_small: 80 % of its records are above 24 units -> parked wide records / multi-pass units)
import ctypes
lib = ruf.load()
sc = helpers.scene("pr2_small"); proj, _, _ = sc.proj()
n = sc.width * sc.height
ptrs = [ruf.host_alloc(n * 2), ruf.host_alloc(n * 2), ruf.host_alloc(n)]
view = lambda p, dt, cnt: np.frombuffer((ctypes.c_uint8 * (cnt * np.dtype(dt).itemsize)).from_address(p), dtype=dt)
h_in, h_out, h_mask = view(ptrs[0], np.uint16, n), view(ptrs[1], np.uint16, n), view(ptrs[2], np.uint8, n)
with ruf.Context(sc.width, sc.height) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    for k in (3, 8):
        fr = helpers.make_frame(sc, k, "u16")
        h_in[:] = fr["depth"].reshape(-1)
        v, pm, pr = [np.ascontiguousarray(a, np.float64) for a in (fr["view"], fr["pm"], proj)]
        rc = lib.ruf_filter(ctx._h, ptrs[0], ruf.ENC_U16_MM, pr.ctypes.data, v.ctypes.data, pm.ctypes.data, sc.max_diff,
                            sc.replace_value, ptrs[1], ptrs[2])
        assert rc == 0
        want_d, want_m, _ = helpers.oracle_filter(sc, fr)
        assert np.array_equal(h_out.reshape(want_d.shape), want_d) and np.array_equal(h_mask.reshape(want_m.shape), want_m)
    ctx.set_mask_format(ruf.MASK_BITS)
    d, bits = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
    assert np.array_equal(np.unpackbits(bits, axis=-1, bitorder="little") * np.uint8(255), want_m)
for p in ptrs:
    ruf.host_free(p)
print("graph + packed mask ok, RUF_MULTIPASS =", os.environ.get("RUF_MULTIPASS", "auto"), "RUF_CLUSTER =", os.environ.get("RUF_CLUSTER", "auto"),
      "RUF_DIRECT =", os.environ.get("RUF_DIRECT", "7"))
