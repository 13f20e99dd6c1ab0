#!/usr/bin/env python
"""us per launch of small device-resident batches (C2 scene) with the small-launch variants forced on / off:
where should the automatic choice (kClusterMaxTiles, kFineMaxCtas in ruf_device.cuh) switch?"""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import realtime_urdf_filter_b200 as ruf
    from realtime_urdf_filter_b200 import synth
    sc = synth.pr2_like_scene()
    proj, _, _ = sc.proj()
    out = {}
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream(device=dev)
    with ruf.Context(sc.width, sc.height) as ctx, torch.cuda.stream(stream):
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.set_stream(stream.cuda_stream)          # the events below are recorded on the stream the kernels run on
        for n in (1, 2, 4, 8, 12, 16, 24, 32):
            views, pms = sc.frames(list(range(n)))
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            d_in = torch.randint(400, 4000, (n, sc.height, sc.width), dtype=torch.int16, device=dev)
            d_out = torch.empty_like(d_in); d_mask = torch.empty(d_in.shape, dtype=torch.uint8, device=dev)
            d_proj, d_view, d_pm = t(proj), t(views), t(pms)
            args = (n, d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(),
                    sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
            for _ in range(20):
                ctx.filter_batch_device(*args)
            ctx.sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(200):
                ctx.filter_batch_device(*args)
            e1.record(); torch.cuda.synchronize()
            out[n] = round(e0.elapsed_time(e1) / 200 * 1e3, 1)
    print(json.dumps(out))
else:
    for cl, fine in (("0", "0"), ("1", "0"), ("0", "1"), ("1", "1")):
        env = dict(os.environ, RUF_CLUSTER=cl, RUF_FINE_MESHLETS=fine)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print("cluster", cl, "fine", fine, "us per launch by frames:", r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:])
