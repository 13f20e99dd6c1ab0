"""The multi-GPU group of the C ABI (ruf_group_*): one process, one context + host thread per device,
NCCL broadcast of the model at set-up, frame k -> GPU k mod N, results in sequence order.  Runs with as many devices
as the box has (1 on the default GPU box: the sharding and ordering logic is the same, NCCL is skipped)."""
import numpy as np
import pytest

import helpers
import realtime_urdf_filter_b200 as ruf

pytestmark = pytest.mark.gpu


def _frames(sc, ks, enc="u16"):
    frs = [helpers.make_frame(sc, k, enc) for k in ks]
    return (np.stack([f["depth"] for f in frs]), np.stack([f["view"] for f in frs]), np.stack([f["pm"] for f in frs]), frs)


@pytest.mark.parametrize("frames_per_chunk", [1, 3, 0])
def test_group_matches_single_context_and_oracle(frames_per_chunk):
    import torch
    n_dev = min(torch.cuda.device_count(), 4)
    sc = helpers.scene("pr2_small")
    proj, _, _ = sc.proj()
    ks = list(range(0, 22))
    depth, views, pms, frs = _frames(sc, ks)
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        want_d, want_m = ctx.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value)
    with ruf.Group(sc.width, sc.height, n_devices=n_dev) as grp:
        sent = grp.set_model(sc.tri, sc.tri_part, sc.n_parts)
        assert (sent > 0) == (n_dev > 1)               # one broadcast per static buffer, only when there are peers
        got_d, got_m = grp.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value,
                                             frames_per_chunk=frames_per_chunk)
        # a second call reuses the staging of every member
        got_d2, got_m2 = grp.filter_batch_host(depth, proj, views, pms, sc.max_diff, sc.replace_value,
                                               frames_per_chunk=frames_per_chunk)
    assert np.array_equal(got_d, want_d) and np.array_equal(got_m, want_m)      # sequence order, bit for bit
    assert np.array_equal(got_d2, want_d) and np.array_equal(got_m2, want_m)
    for i in (0, 7, 21):
        od, om, _ = helpers.oracle_filter(sc, frs[i])
        assert np.array_equal(got_d[i], od) and np.array_equal(got_m[i], om)


def test_group_member_contexts_are_independent_streams():
    """BASELINE configs[3]: one camera stream per GPU -- every member is a full ruf_context."""
    import ctypes as C
    import torch
    n_dev = min(torch.cuda.device_count(), 4)
    sc = helpers.scene("example")
    proj, _, _ = sc.proj()
    fr = helpers.make_frame(sc, 0, "u16")
    od, om, _ = helpers.oracle_filter(sc, fr)
    lib = ruf.load()
    with ruf.Group(sc.width, sc.height, n_devices=n_dev) as grp:
        grp.set_model(sc.tri, sc.tri_part, sc.n_parts)
        assert lib.ruf_group_size(grp._h) == n_dev
        for i in range(n_dev):
            h = lib.ruf_group_context(grp._h, i)
            out = np.empty_like(fr["depth"])
            mask = np.empty(fr["depth"].shape, np.uint8)
            pm = np.ascontiguousarray(fr["pm"], np.float64)
            view = np.ascontiguousarray(fr["view"], np.float64)
            pr = np.ascontiguousarray(proj, np.float64)
            rc = lib.ruf_filter(C.c_void_p(h), fr["depth"].ctypes.data, ruf.ENC_U16_MM, pr.ctypes.data, view.ctypes.data,
                                pm.ctypes.data, sc.max_diff, sc.replace_value, out.ctypes.data, mask.ctypes.data)
            assert rc == 0
            assert np.array_equal(out, od) and np.array_equal(mask, om)
        assert lib.ruf_group_context(grp._h, n_dev) is None
