"""Section \"exotic sensor values\" of profiles/r02_gl_crosscheck.txt: NaN, +-inf, +-0.0, negatives, 3e38 and denormals in the 32FC1
sensor image, per class: what mix() in the reference's shader makes of them in the GL driver vs the CPU oracle.
TEST INFRASTRUCTURE ONLY."""
import os, sys, subprocess, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT+'/oracle'); sys.path.insert(0, ROOT+'/oracle/gl_ref'); sys.path.insert(0, ROOT+'/tests')
import gl_case, oracle_py as orc, helpers
mesa = subprocess.run(["make","-s","-C",ROOT+"/oracle/gl_ref","mesa_dir"],capture_output=True,text=True).stdout.strip()
env = dict(os.environ, LD_LIBRARY_PATH=ROOT+"/oracle/_ref/fakex:"+mesa)
for seed in (1, 6, 9):
    fc = helpers.fuzz_case(seed)
    rng = np.random.default_rng(seed)
    d = fc["depth"]
    sel = rng.random(d.shape)
    d[sel < 0.03] = np.nan; d[(sel >= 0.03) & (sel < 0.05)] = np.inf; d[(sel >= 0.05) & (sel < 0.07)] = -np.inf
    d[(sel >= 0.07) & (sel < 0.09)] = 0.0; d[(sel >= 0.09) & (sel < 0.11)] = -1.5; d[(sel >= 0.11) & (sel < 0.12)] = 3.0e38
    d[(sel >= 0.12) & (sel < 0.13)] = 1e-42   # denormal
    d[(sel >= 0.13) & (sel < 0.15)] = -0.0
    want_d, want_m = helpers.fuzz_oracle(fc)
    la = np.asarray(orc.lookat()).reshape(4, 4).T
    cam = (np.linalg.inv(la) @ np.asarray(fc["view"]).reshape(4, 4).T).T.reshape(-1)
    with tempfile.TemporaryDirectory() as td:
        gl_case.write_case_raw(td+"/c.bin", fc["W"], fc["H"], fc["proj"], np.eye(4).reshape(-1), cam, fc["pm"], fc["tri"], fc["part"], d, 0.1, 8.0, 0.05, 5.0)
        r = subprocess.run([ROOT+"/oracle/_ref/gl_crosscheck_glx", td+"/c.bin", "-", td+"/o.bin"], capture_output=True, text=True, env=env)
        gd, gm = gl_case.read_dump(td+"/o.bin")
    dm = gm != want_m
    special = ~np.isfinite(d) | (d <= 0) | (d > 1e30) | (np.abs(d) < 1e-38)
    dd = (gd.view(np.uint32) != want_d.view(np.uint32)) & ~dm
    print("seed", seed, "mask diff", int(dm.sum()), "of which on special sensor values", int((dm & special).sum()), "| depth-bit diff (mask equal)", int(dd.sum()), "of which special", int((dd & special).sum()))
    if dd.sum():
        idx = np.argwhere(dd)[:6]
        for y, x in idx: print("   sensor", d[y, x], "gl", gd[y, x], hex(gd.view(np.uint32)[y, x]), "oracle", want_d[y, x], hex(want_d.view(np.uint32)[y, x]))
print("---- per class (last seed)")
classes = {"nan": np.isnan(d), "+inf": d == np.inf, "-inf": d == -np.inf, "zero": (d == 0) & ~np.signbit(d), "neg": d == -1.5, "huge": d == np.float32(3.0e38), "denorm": d == np.float32(1e-42), "negzero": (d == 0) & np.signbit(d)}
for n, sel in classes.items():
    k = sel & ~dm
    nd = int((gd.view(np.uint32)[k] != want_d.view(np.uint32)[k]).sum())
    flt = gm[k] == 255
    ex = np.argwhere(k & (gd.view(np.uint32) != want_d.view(np.uint32)))[:2]
    print(n, "pixels", int(k.sum()), "filtered", int(flt.sum()), "depth-bit diffs", nd, [(hex(gd.view(np.uint32)[y, x]), hex(want_d.view(np.uint32)[y, x]), int(gm[y, x])) for y, x in ex])
# negative zero
