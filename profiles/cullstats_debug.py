import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
import helpers, realtime_urdf_filter_b200 as ruf
sc = helpers.scene("pr2"); proj,_,_ = sc.proj()
dev = torch.device("cuda:0")
for ks in ([0],[19],[40]):
    frames=[helpers.make_frame(sc,k,"u16") for k in ks]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_in=t(np.stack([f["depth"] for f in frames]).view(np.int16)); d_out=torch.empty_like(d_in); d_mask=torch.empty(d_in.shape,dtype=torch.uint8,device=dev)
    d_proj=t(proj); d_view=t(np.stack([f["view"] for f in frames])); d_pm=t(np.stack([f["pm"] for f in frames]))
    torch.cuda.synchronize()
    with ruf.Context(sc.width, sc.height) as ctx:
        ctx.set_model(sc.tri, sc.tri_part, sc.n_parts)
        ctx.filter_batch_device(len(ks), d_in.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(), d_view.data_ptr(), d_pm.data_ptr(), sc.max_diff, sc.replace_value, d_out.data_ptr(), d_mask.data_ptr(), 0)
        ctx.sync(); st=ctx.stats()
        print(ks, 'refs', st['binned_refs'], 'visible', st['visible_tris'], 'back tested', st['h2d_bytes'], 'culled', st['d2h_bytes'])
