"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit-exact."""
import numpy as np
import pytest

import helpers
import realtime_urdf_filter_b200 as ruf

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["throughput", "small-launch"])
def raster_variant(request, monkeypatch):
    """Every test of this file runs twice: with the throughput kernels (one CTA per tile, coarse meshlet cut) and with what
    small launches get by default (cluster-split raster variant, fine meshlet cut).  Same bits either way."""
    v = "0" if request.param == "throughput" else "1"
    monkeypatch.setenv("RUF_CLUSTER", v)
    monkeypatch.setenv("RUF_FINE_MESHLETS", v)


def _ctx(sc):
    ctx = ruf.Context(sc.width, sc.height)
    ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
    return ctx


@pytest.mark.parametrize("name", ["example", "pr2_small", "pr2"])
@pytest.mark.parametrize("enc", ["u16", "f32"])
def test_single_frame_host_api(name, enc):
    sc = helpers.scene(name)
    proj, _, _ = sc.proj()
    with _ctx(sc) as ctx:
        for k in (0, 11):
            fr = helpers.make_frame(sc, k, enc)
            want_d, want_m, _ = helpers.oracle_filter(sc, fr)
            got_d, got_m = ctx.filter(fr["depth"], proj, fr["view"], fr["pm"], sc.max_diff, sc.replace_value)
            assert np.array_equal(got_m, want_m), f"mask differs at {np.count_nonzero(got_m != want_m)} px"
            if enc == "u16":
                assert np.array_equal(got_d, want_d)
            else:
                assert np.array_equal(got_d.view(np.uint32), want_d.view(np.uint32))   # NaNs included
            # all outcomes present
            assert 0 < np.count_nonzero(want_m) < want_m.size


def test_zbuf_bit_exact_device_batch():
    import torch
    sc = helpers.scene("pr2")
    proj, _, _ = sc.proj()
    ks = [0, 3, 17, 40]
    frames = [helpers.make_frame(sc, k, "u16") for k in ks]
    dev = torch.device("cuda:0")
    d_in = torch.from_numpy(np.stack([f["depth"] for f in frames]).view(np.int16)).to(dev)
    d_out = torch.empty_like(d_in)
    d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
    d_z = torch.empty(d_in.shape, dtype=torch.float32, device=dev)
    d_proj = torch.from_numpy(proj).to(dev)
    d_view = torch.from_numpy(np.stack([f["view"] for f in frames])).to(dev)
    d_pm = torch.from_numpy(np.stack([f["pm"] for f in frames])).to(dev)
    stream = torch.cuda.Stream(device=dev)
    stream.wait_stream(torch.cuda.current_stream())
    with _ctx(sc) as ctx, torch.cuda.stream(stream):
        ctx.set_stream(stream.cuda_stream)
        ctx.filter_batch_device(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(),
                                d_pm.data_ptr(), sc.max_diff, sc.replace_value, d_out.data_ptr(),
                                d_mask.data_ptr(), d_z.data_ptr())
        ctx.sync()
        st = ctx.stats()
    assert st["kernel_launches"] in (3, 4) and st["visible_tris"] > 0     # 3: small launch, cluster-split raster variant (no tile-info kernel)
    z = d_z.cpu().numpy()
    for i, fr in enumerate(frames):
        want_d, want_m, want_z = helpers.oracle_filter(sc, fr)
        assert np.array_equal(z[i].view(np.uint32), want_z.view(np.uint32))
        assert np.array_equal(d_mask[i].cpu().numpy(), want_m)
        assert np.array_equal(d_out[i].cpu().numpy().view(np.uint16), want_d)
