import sys
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
sc = synth.pr2_like_scene(100, 75, n_tris=3000, name="odd")
proj, _, _ = sc.proj(); fr = helpers.make_frame(sc, 1, "f32")
with ruf.Context(100, 75) as ctx:
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    ctx.reserve(1, 2, 64)
    d, m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
print("odd ok")
