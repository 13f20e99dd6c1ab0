#!/usr/bin/env python
"""Benchmark of the URDF depth self-filter hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c1|c2|c3|c5]

One "step" = one pass of the hot path over RING x BATCH distinct-address synthetic depth frames.
`--config` picks the BASELINE.json configuration (default c2 = configs[1], the one the metric is quoted on):

    c1  640x480   urdf/example.urdf.xml (two boxes + their doubled cubes), static pose        configs[0]
    c2  640x480   PR2-like synthetic model (~90k triangles)                                  configs[1]; with --gpus N: configs[3]
    c3  1280x960  PR2-like model + two static wall boxes                                     configs[2]
    c5  1920x1080 four articulated PR2-like URDFs (~500k triangles), animated joint sweep    configs[4]

Multi-GPU (torchrun, one rank per GPU, NCCL): c1-c3 give every rank its own camera stream (weak scaling,
"8 concurrent streams" of configs[3]); c5 is ONE stream whose frame k goes to GPU k mod N (strong scaling:
the frames per step are fixed), with an ordered gather of the results over NCCL checked after the timed region.
There is no per-frame collective; the mesh, poses and frames are generated on rank 0 and broadcast once.

Prints ONE JSON line (rank 0).  `value` = frames/s with inputs resident in HBM; `e2e` = frames/s through
ruf_filter_batch_host with pinned HOST buffers (H2D + D2H inside the timed region) next to the copy-only
ceiling of the same pipeline on this box; `roofline` = dominant kernel vs the measured HBM peak;
`cpu_baseline` = the CPU oracle timed on this box's host cores on a bounded sample.

`--impl reference` times the CPU restatement of the reference's algorithm (oracle/, OpenMP on all host
threads this process may run on) -- the reference's own GL path cannot run in this image (no
libGL/Mesa/X11, DESIGN.md "Oracle").  That arm builds its scene with the oracle's own host math, so it never
loads the product library.  This script, tests/ and __graft_entry__.smoke() are the only places allowed to
execute oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "depth frames/sec at 640×480, PR2 URDF; achieved HBM GB/s vs peak"
UNIT = "frames/s"
FALLBACK_HBM_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback
MAX_DIFF, REPLACE_VALUE = 0.05, 5.0          # launch/filter_parameters.yaml:14,16

CONFIGS = {
    "c1": dict(scene="example", size=(640, 480), batch=1024, cpu_frames=32, baseline="configs[0]",
               what="urdf/example.urdf.xml (two 4x0.5x2 boxes + their doubled cubes, SURVEY F4), static pose"),
    "c2": dict(scene="pr2", size=(640, 480), batch=1024, cpu_frames=16, baseline="configs[1]",
               what="PR2-like synthetic model"),
    "c3": dict(scene="walls", size=(1280, 960), batch=256, cpu_frames=8, baseline="configs[2]",
               what="PR2-like synthetic model + two static wall boxes"),
    "c5": dict(scene="multi", size=(1920, 1080), batch=128, cpu_frames=2, baseline="configs[4]",
               what="four articulated PR2-like URDFs, animated joint sweep"),
}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(W, H, n_tris, n_parts, elem=2, mask=True):
    """B = W*H*(b_in + b_out + b_mask) + T*36 + L*64 per frame."""
    image = W * H * (elem + elem + (1 if mask else 0))
    geometry = n_tris * 36 + n_parts * 64
    return image, geometry


def make_scene(cfg):
    from realtime_urdf_filter_b200 import synth
    W, H = cfg["size"]
    return {"example": synth.example_scene, "pr2": synth.pr2_like_scene, "walls": synth.walls_scene,
            "multi": synth.multi_robot_scene}[cfg["scene"]](W, H)


def config_dict(name, cfg, n_tris, n_parts, world):
    """The workload-defining keys: identical in the B200 arm and in the reference arm."""
    W, H = cfg["size"]
    streams = "one stream, frame k -> GPU k mod N" if name == "c5" else "one camera stream per GPU"
    return {
        "workload": f"{W}x{H} 16UC1 stream, {cfg['what']} ({n_parts} parts, {n_tris} triangles), mask on "
                    f"(BASELINE.json {cfg['baseline']})",
        "config": name, "width": W, "height": H, "encoding": "16UC1", "mask": True, "triangles": int(n_tris),
        "parts": int(n_parts), "threshold_m": MAX_DIFF, "replace_value_m": REPLACE_VALUE, "near_far_m": [0.1, 8.0],
        "sharding": streams if world > 1 else "single GPU",
    }


# --------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread during the timed region
# --------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.sm_max = None
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.thread:
            self.thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference) on the host cores
# --------------------------------------------------------------------------------------------
def host_threads() -> int:
    """Every core this process may run on.  NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its
    workers, which silently turned the CPU arm into a one-thread baseline."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def load_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py as orc
    return orc


class OracleMath:
    """synth.use_math provider backed by the CPU oracle (bit-identical with the library's host functions:
    tests/test_host_math.py), so the reference arm builds its scene without loading libruf_b200.so."""

    def __init__(self, orc):
        self.o = orc
        for n in ("projection_matrix", "view_matrix", "box_triangles", "cube_triangles", "sphere_triangles",
                  "cylinder_triangles"):
            setattr(self, n, getattr(orc, n))

    def part_model(self, link_q, link_t, off_q=(0, 0, 0, 1), off_t=(0, 0, 0), suffix=None):
        return self.o.link_model(link_q, link_t, off_q, off_t, suffix)


class OracleWorkload:
    """Sample frames of a config for the CPU arm: poses -> MVPs -> synthetic sensor frames."""

    def __init__(self, orc, sc, n_frames, threads, depth=None, frames=None):
        from realtime_urdf_filter_b200 import synth
        self.orc, self.sc, self.threads = orc, sc, threads
        proj, _, _ = sc.proj()
        views, pms = frames if frames is not None else sc.frames(list(range(n_frames)))
        self.mvps = [orc.compose_mvp(proj, views[i], pms[i], sc.n_parts) for i in range(n_frames)]
        self.zn, self.zf = np.float32(synth.Z_NEAR), np.float32(synth.Z_FAR)
        if depth is None:
            depth = []
            for i in range(n_frames):
                z = orc.render(sc.tri, sc.tri_part, self.mvps[i], sc.width, sc.height, np.float32(8.0 * 0.99),
                               nthreads=threads, native=True)
                depth.append(synth.synth_depth(synth.linear_depth(z), i, "u16"))
        self.depth = depth

    def one(self, i):
        j = i % len(self.mvps)
        self.orc.filter_frame(self.depth[j], self.sc.tri, self.sc.tri_part, self.mvps[j], self.zn, self.zf,
                              np.float32(MAX_DIFF), np.float32(REPLACE_VALUE), want_mask=True, nthreads=self.threads,
                              native=True)


def gl_reference_arm(args, cfg, sc):
    """The reference's OWN implementation of the path where it can run: its two shader files (compiled into
    oracle/_ref/gl_crosscheck_glx when the checkout was present at build time) in Mesa llvmpipe -- the libGL inside Nsight
    Compute on oracle/gl_ref/fakex11 --, driven through the GL call sequence of RealtimeURDFFilter::render: sensor upload,
    render, both read-backs per frame.  One harness run per step (another frame of the stream each), timed inside the
    harness in steady state.  -> (frames/s, ms per step, threads, description) or None when any piece is missing."""
    import glob
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "oracle", "_ref", "gl_crosscheck_glx")
    mesa = sorted(glob.glob("/opt/nvidia/nsight-compute/*/host/linux-desktop-glibc_2_11_3-x64/Mesa"))
    if not os.path.exists(exe) or not mesa or os.environ.get("RUF_NO_GL_REFERENCE"):
        return None
    sys.path.insert(0, os.path.join(ROOT, "oracle", "gl_ref"))
    import gl_case
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "oracle", "_ref", "fakex") + ":" + mesa[0])
    per_step = max(8, 2 * cfg["cpu_frames"])          # ~1 s of llvmpipe per step: past the shader JIT and the thread ramp
    ms = []
    try:
        with tempfile.TemporaryDirectory() as td:
            for s_ in range(args.warmup + args.steps):
                case, dump = os.path.join(td, "case.bin"), os.path.join(td, "dump.bin")
                gl_case.write_case(case, sc, s_, gl_case._frame_depth(sc, s_))
                r = subprocess.run([exe, case, "-", dump, str(per_step + 1)], capture_output=True, text=True, env=env, timeout=600)
                t = [l for l in r.stderr.splitlines() if l.startswith("TIMING")]
                if r.returncode != 0 or not t:
                    return None
                if s_ >= args.warmup:
                    ms.append(float(t[0].split()[4]))
            gl_info = r.stderr.splitlines()[0]
    except Exception:
        return None
    mean_ms = sum(ms) / len(ms)
    threads = min(16, os.cpu_count() or 1)            # llvmpipe: one rasteriser thread per online core, at most 16
    return 1e3 / mean_ms, mean_ms * per_step, threads, per_step, gl_info


def run_reference(args, rank, world):
    """--impl reference: the reference's implementation of the path on the host cores (rank 0 only): its GLSL path on Mesa
    llvmpipe where the harness and the driver exist (kind "reference"), else the CPU oracle on all host threads (kind
    "port"); K steps of a bounded sample each."""
    if rank != 0:
        return
    orc = load_oracle()
    from realtime_urdf_filter_b200 import synth
    synth.use_math(OracleMath(orc))
    cfg = CONFIGS[args.config]
    sc = make_scene(cfg)
    gl = gl_reference_arm(args, cfg, sc)
    if gl is not None:
        fps, ms_step, threads, per_step, gl_info = gl
        line = {
            "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if (args.config == "c5" and args.gpus > 1) else "weak",
            "vs_baseline": None, "dtype": "f32 (GLSL)", "data": "synthetic",
            "config": config_dict(args.config, cfg, sc.n_tris, sc.n_parts, args.gpus),
            "run": {"frames_per_step": per_step, "distinct_frames": args.steps},
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "reference",
                             "sample": f"{args.steps} steps x {per_step} frames: the reference's own shader files (unmodified) through the GL "
                                       f"call sequence of RealtimeURDFFilter::render -- sensor upload, render, both read-backs per frame -- in "
                                       f"{gl_info} (oracle/_ref/gl_crosscheck_glx on oracle/gl_ref/fakex11), timed in steady state"},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "native_libraries": sorted({os.path.basename(l.split()[-1]) for l in open("/proc/self/maps")
                                        if "libruf" in l or "ruf_oracle" in l}),
        }
        print(json.dumps(line), flush=True)
        return
    threads = host_threads()
    nf = min(8, cfg["cpu_frames"])
    work = OracleWorkload(orc, sc, nf, threads)
    per_step = cfg["cpu_frames"]     # frames per step: bounded so that K steps finish within minutes
    for s in range(args.warmup):
        for i in range(per_step):
            work.one(s * per_step + i)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for i in range(per_step):
            work.one(s * per_step + i)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong" if (args.config == "c5" and args.gpus > 1) else "weak",
        "vs_baseline": None, "dtype": "f32+i64 (u16 I/O)", "data": "synthetic",
        "config": config_dict(args.config, cfg, sc.n_tris, sc.n_parts, args.gpus),
        "run": {"frames_per_step": per_step, "distinct_frames": nf},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps x {per_step} frames of the same workload, oracle/ruf_oracle.c "
                                   f"-O3 -march=native, OpenMP over {threads} threads (sched_getaffinity; the "
                                   "reference's GLSL path on llvmpipe is the `--impl reference` arm where oracle/_ref/gl_crosscheck_glx exists)"},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "native_libraries": sorted({os.path.basename(l.split()[-1]) for l in open("/proc/self/maps")
                                    if "libruf" in l or "ruf_oracle" in l}),
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def pin_rank_to_cores(local_rank, local_world):
    """Before any pinned allocation: give every rank of the box its own slice of the cores (first-touch places the
    pinned staging buffers next to them; copy threads of different ranks stop migrating over each other)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        if local_world > 1 and len(cores) >= local_world:
            per = len(cores) // local_world
            os.sched_setaffinity(0, cores[local_rank * per:(local_rank + 1) * per])
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def run_b200(args, rank, world, local_rank):
    cores_of_rank = pin_rank_to_cores(local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    import torch
    import torch.distributed as dist
    import realtime_urdf_filter_b200 as ruf
    from realtime_urdf_filter_b200 import sharding, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    cfg = CONFIGS[args.config]
    W_IMG, H_IMG = cfg["size"]
    strong = args.config == "c5" and world > 1
    B_total = args.batch or cfg["batch"]          # frames per launch sequence of the WHOLE job (strong) / per rank (weak)
    R = args.ring
    if strong:
        B_total -= B_total % world
    B = B_total // world if strong else B_total   # frames this rank launches per sequence
    n_distinct = min(B_total, 64)                 # distinct poses / depth frames; ring slots hold rolled repeats of them
    ctx = ruf.Context(W_IMG, H_IMG, device=local_rank)
    # a real (non-default) stream: torch events must sit on the stream the kernels run on, and the
    # legacy default stream (handle 0) cannot be handed to ruf_set_stream (NULL = internal stream)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0

    # ---- set-up on rank 0, then one NCCL broadcast per static buffer (mesh, poses) ----
    sc, arrays = None, None
    if rank == 0:
        sc = make_scene(cfg)
        views, pms = sc.frames(list(range(n_distinct)))
        proj, _, _ = sc.proj()
        arrays = {"tri": sc.tri, "tri_part": sc.tri_part, "proj": proj, "views": views, "pms": pms}
    got = sharding.broadcast_arrays(arrays, 0, dev, rank)
    d_tri, d_part, d_proj, d_views, d_pms = got["tri"], got["tri_part"], got["proj"], got["views"], got["pms"]
    T, P = int(d_tri.shape[0]), int(d_pms.shape[1])
    d_depth0 = torch.empty((n_distinct, H_IMG, W_IMG), dtype=torch.int16, device=dev)
    torch.cuda.synchronize()
    ctx.set_model_device(d_tri.data_ptr(), d_part.data_ptr(), T, P)
    ctx.reserve(B)

    def run_checked(fn):
        """An internal list that turns out too small is grown by the library; the call is repeated once."""
        for attempt in range(4):
            fn()
            try:
                ctx.sync()
                return
            except ruf.RufError as e:
                if e.code != ruf.RUF_ERR_OVERFLOW:
                    raise
        raise RuntimeError("internal lists still overflow after 4 attempts")

    # synthetic sensor frames from the virtual depth of the same poses (rank 0), then broadcast
    frames_np = None
    if rank == 0:
        d_z = torch.empty((n_distinct, H_IMG, W_IMG), dtype=torch.float32, device=dev)
        d_tmp = torch.zeros((n_distinct, H_IMG, W_IMG), dtype=torch.int16, device=dev)
        run_checked(lambda: ctx.filter_batch_device(n_distinct, d_tmp.data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(),
                                                    d_views.data_ptr(), d_pms.data_ptr(), MAX_DIFF, REPLACE_VALUE,
                                                    d_depth0.data_ptr(), 0, d_z.data_ptr()))
        z = d_z.cpu().numpy()
        frames_np = np.stack([synth.synth_depth(synth.linear_depth(z[i]), i, "u16") for i in range(n_distinct)])
        d_depth0.copy_(torch.from_numpy(frames_np.view(np.int16)))
        del d_z, d_tmp
    if world > 1:
        dist.broadcast(d_depth0.view(torch.uint8), 0)

    # ---- this rank's frames.  Weak: its own stream (the distinct frames rolled by a rank-dependent shift).
    # Strong (c5): ONE stream of B_total frames per ring slot, frame k -> GPU k mod N (sharding.frames_for_rank).
    reps_b = (B_total + n_distinct - 1) // n_distinct
    rep = lambda t: t.repeat(reps_b, *([1] * (t.dim() - 1)))[:B_total]
    mine = torch.tensor(sharding.frames_for_rank(B_total, rank, world), device=dev) if strong else None
    ring_in = torch.empty((R, B, H_IMG, W_IMG), dtype=torch.int16, device=dev)
    ring_out = torch.empty_like(ring_in)
    ring_mask = torch.empty((R, B, H_IMG, W_IMG), dtype=torch.uint8, device=dev)
    ring_views = torch.empty((R, B, 16), dtype=torch.float64, device=dev)
    ring_pms = torch.empty((R, B, P, 16), dtype=torch.float64, device=dev)
    for r in range(R):
        sh = (r * 7 + (0 if strong else rank * 3)) % B_total
        pick = (lambda t: torch.roll(rep(t), sh, 0)[mine]) if strong else (lambda t: torch.roll(rep(t), sh, 0))
        ring_in[r], ring_views[r], ring_pms[r] = pick(d_depth0), pick(d_views), pick(d_pms)
    torch.cuda.synchronize()

    def step():
        for r in range(R):
            ctx.filter_batch_device(B, ring_in[r].data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(),
                                    ring_views[r].data_ptr(), ring_pms[r].data_ptr(), MAX_DIFF, REPLACE_VALUE,
                                    ring_out[r].data_ptr(), ring_mask[r].data_ptr(), 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_checked(step)
    for _ in range(args.warmup):
        step()
    ctx.sync()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    ctx.sync()        # raises on deferred overflow
    stats = ctx.stats()
    # per-kernel durations: a separate pass over the same steps with CUDA events recorded between the kernels on
    # the launching stream (ruf_set_profiling).  Events between kernels serialise them, so this pass is kept
    # out of the timed region above.
    ctx.set_profiling(True)
    ctx.stage_times(reset=True)
    for _ in range(max(1, min(args.steps, 5))):
        step()
    ctx.sync()
    stage_ms, seqs = ctx.stage_times(reset=True)
    ctx.set_profiling(False)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms = float(t_ms.item())
    frames_per_step = R * B_total * (1 if strong else world)
    value = frames_per_step * args.steps / (ms * 1e-3)

    # ---- strong scaling: ordered gather over NCCL, checked against rank 0 running the whole stream itself ----
    gather_ok = None
    if strong:
        # raw bytes on the wire: NCCL has no 16-bit integer type
        mine_bytes = ring_out[0].contiguous().view(torch.uint8)
        parts = [torch.empty_like(mine_bytes) for _ in range(world)] if rank == 0 else None
        dist.gather(mine_bytes, parts, dst=0)
        if rank == 0:
            whole = torch.empty((B_total, H_IMG, W_IMG), dtype=torch.int16, device=dev)
            for r_, p_ in enumerate(parts):
                whole[torch.tensor(sharding.frames_for_rank(B_total, r_, world), device=dev)] = p_.view(torch.int16)   # sequence order
            ref_out = torch.empty_like(whole)
            chunk = max(1, B)
            for f0 in range(0, B_total, chunk):
                n = min(chunk, B_total - f0)
                ctx.filter_batch_device(n, rep(d_depth0)[f0:f0 + n].contiguous().data_ptr(), ruf.ENC_U16_MM, d_proj.data_ptr(),
                                        rep(d_views)[f0:f0 + n].contiguous().data_ptr(), rep(d_pms)[f0:f0 + n].contiguous().data_ptr(),
                                        MAX_DIFF, REPLACE_VALUE, ref_out[f0:f0 + n].data_ptr(), 0, 0)
                ctx.sync()
            gather_ok = bool(torch.equal(whole, ref_out))
            del whole, ref_out, parts

    # ---- e2e: pinned host buffers through ruf_filter_batch_host (H2D + D2H inside the region) ----
    n_e2e = max(1, min(args.e2e_frames, 4 * B))
    h_in = torch.empty((n_e2e, H_IMG, W_IMG), dtype=torch.int16).pin_memory()
    h_out = torch.empty_like(h_in).pin_memory()
    h_mask = torch.empty((n_e2e, H_IMG, W_IMG), dtype=torch.uint8).pin_memory()
    reps = (n_e2e + n_distinct - 1) // n_distinct
    h_in.copy_(d_depth0.cpu().repeat(reps, 1, 1)[:n_e2e])
    e_views = d_views.cpu().repeat(reps, 1)[:n_e2e].contiguous().numpy()
    e_pms = d_pms.cpu().repeat(reps, 1, 1)[:n_e2e].contiguous().numpy()
    e_proj = d_proj.cpu().numpy()
    lib = ruf.load()

    def e2e_step():
        rc = lib.ruf_filter_batch_host(ctx._h, n_e2e, h_in.data_ptr(), ruf.ENC_U16_MM, e_proj.ctypes.data,
                                       e_views.ctypes.data, e_pms.ctypes.data, MAX_DIFF, REPLACE_VALUE,
                                       h_out.data_ptr(), h_mask.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.ruf_last_error(ctx._h).decode())

    def copy_step():
        rc = lib.ruf_host_copy_ceiling(ctx._h, n_e2e, h_in.data_ptr(), ruf.ENC_U16_MM, e_proj.ctypes.data,
                                       e_views.ctypes.data, e_pms.ctypes.data, h_out.data_ptr(), h_mask.data_ptr())
        if rc != 0:
            raise RuntimeError(lib.ruf_last_error(ctx._h).decode())

    def timed(fn, n):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()          # synchronous: returns after the last D2H landed
        barrier()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n_e2e * n / float(t.item())

    e2e_steps = max(3, min(args.steps, 10))
    # the copy-only pipeline first (same chunks, slots and streams, no kernels): the PCIe / host-memory ceiling
    copy_value = timed(copy_step, e2e_steps)
    e2e_value = timed(e2e_step, e2e_steps)
    e2e_stats = ctx.stats()
    # opt-in output format RUF_MASK_BITS (1 bit per pixel on the wire instead of one byte): same calls, same buffers
    packed = None
    if W_IMG % 8 == 0:
        ctx.set_mask_format(ruf.MASK_BITS)
        p_ceiling, p_value = timed(copy_step, e2e_steps), timed(e2e_step, e2e_steps)
        p_stats = ctx.stats()
        bits_ok = bool(np.array_equal(
            np.unpackbits(h_mask.view(-1)[:H_IMG * W_IMG // 8].numpy(), bitorder="little") * np.uint8(255),
            ring_mask[0][0].cpu().numpy().reshape(-1))) if not strong and rank * 3 % B == 0 else None
        ctx.set_mask_format(ruf.MASK_BYTES)
        packed = {"value": p_value, "unit": UNIT, "copy_ceiling": p_ceiling, "frac_of_copy_ceiling": p_value / p_ceiling,
                  "d2h_bytes_per_step": int(p_stats["d2h_bytes"]), "h2d_bytes_per_step": int(p_stats["h2d_bytes"]),
                  "matches_byte_mask": bits_ok,
                  "note": "ruf_set_mask_format(RUF_MASK_BITS): opt-in, not the reference's MONO8 wire format"}
    # the reference's own use of the path: ONE frame per call at 30 Hz (ruf_filter, host buffers in and out, synchronous).
    # Latency, not throughput: reported beside the e2e figure, N = 1 only.
    single = None
    if world == 1:
        torch.cuda.synchronize()
        ctx.set_stream(None)       # the context's own stream, as a ROS node uses it
        lat = []
        nf = min(n_e2e, 8)         # a driver with a small ring of frame buffers
        ptrs = [(h_in[f].data_ptr(), e_views[f].ctypes.data, e_pms[f].ctypes.data, h_out[f].data_ptr(), h_mask[f].data_ptr())
                for f in range(nf)]
        for k in range(132):
            p_in, p_view, p_pm, p_out, p_mask = ptrs[k % nf]
            t0 = time.perf_counter()
            rc = lib.ruf_filter(ctx._h, p_in, ruf.ENC_U16_MM, e_proj.ctypes.data, p_view, p_pm, MAX_DIFF, REPLACE_VALUE, p_out, p_mask)
            lat.append(time.perf_counter() - t0)
            if rc != 0:
                raise RuntimeError(lib.ruf_last_error(ctx._h).decode())
        lat = np.array(lat[32:]) * 1e6
        ctx.sync()
        ctx.set_stream(stream.cuda_stream)
        single = {"median_us": round(float(np.median(lat)), 1), "p99_us": round(float(np.percentile(lat, 99)), 1), "calls": int(lat.size),
                  "api": "ruf_filter (pinned host buffers in and out, synchronous): one CUDA graph of three kernels that read / write the host buffers themselves"}
    # sanity: the e2e output equals the device-path output for the same frames (weak arm: ring slot 0 is frame order)
    same = None
    if not strong:
        same = bool(torch.equal(h_out[:n_distinct].to(dev), torch.roll(ring_out[0], -((rank * 3) % B), 0)[:n_distinct]))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peak, peak_src = measured_peak()
    img_b, geo_b = algorithmic_bytes(W_IMG, H_IMG, T, P)
    dom = max(stage_ms, key=stage_ms.get)
    share = {k: v / max(sum(stage_ms.values()), 1e-12) for k, v in stage_ms.items()}
    kernel_bytes = {"setup_bin": geo_b, "raster_filter": img_b}.get(dom, img_b + geo_b)
    dom_avg_ms = stage_ms[dom] / max(seqs, 1)
    achieved = kernel_bytes * B / (dom_avg_ms * 1e-3) / 1e9
    path_gbs = (img_b + geo_b) * frames_per_step * args.steps / (ms * 1e-3) / 1e9 / world     # per GPU
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            tj = json.load(f)
            per_frame = tj.get(args.config, tj).get("bytes_per_frame", {}).get(dom)
            traffic = None if per_frame is None else per_frame * B      # per launch, like `achieved`
    except Exception:
        pass

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        orc = load_oracle()
        threads = host_threads()
        work = OracleWorkload(orc, sc, n_distinct, threads, depth=frames_np, frames=(views, pms))
        for i in range(2):
            work.one(i)
        n, t0 = 0, time.perf_counter()
        while True:
            work.one(n)
            n += 1
            dt = time.perf_counter() - t0
            if dt >= args.cpu_seconds:
                break
        cpu = {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{n} frames of the same workload in {dt:.1f} s, oracle/ruf_oracle.c -O3 -march=native, "
                         f"OpenMP over {threads} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32+i64 (u16 I/O)", "data": "synthetic",
        "config": config_dict(args.config, cfg, T, P, world),
        "run": {
            "frames_per_step": frames_per_step, "frames_per_launch_per_gpu": B, "launch_sequences_per_step": R,
            "l2_policy": f"inputs larger than L2: ring of {R} x {B} distinct-address frames per GPU = "
                         f"{R * B * img_b / 1e6:.0f} MB of image traffic per step vs 126 MB L2",
            "cpu_cores_of_this_rank": cores_of_rank,
        },
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(e2e_stats["h2d_bytes"]), "d2h_bytes_per_step": int(e2e_stats["d2h_bytes"]),
                "frames_per_step": n_e2e, "steps": e2e_steps, "api": "ruf_filter_batch_host (pinned host buffers)",
                "matches_device_path": same,
                "copy_ceiling": copy_value, "frac_of_copy_ceiling": e2e_value / copy_value,
                "copy_ceiling_api": "ruf_host_copy_ceiling: the same chunked H2D/D2H pipeline without the kernels",
                "pcie_gbs_at_ceiling": copy_value * (e2e_stats["h2d_bytes"] + e2e_stats["d2h_bytes"]) / n_e2e / 1e9 / world,
                "packed_mask": packed, "single_frame": single},
        "gpu_launches": int(args.steps * R * stats["kernel_launches"]) * world,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": f"ruf_{dom}_kernel", "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": kernel_bytes * B, "avg_launch_ms": dom_avg_ms,
                     "kernel_share_of_step": share[dom],
                     "stage_ms_per_launch": {k: v / max(seqs, 1) for k, v in stage_ms.items()},
                     "path_achieved_gbs": path_gbs, "path_frac": path_gbs / peak,
                     "path_bytes_per_frame": img_b + geo_b},
        "cpu_baseline": cpu,
        "clocks": clocks.summary(),
        "stats_last_launch": {k: stats[k] for k in ("visible_tris", "binned_refs", "big_tris")},
    }
    if strong:
        line["ordered_gather_matches_single_gpu"] = gather_ok
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="frames per launch sequence (0 = the config's default)")
    ap.add_argument("--ring", type=int, default=2, help="launch sequences (distinct buffers) per step")
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per ruf_filter_batch_host call")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 100 if args.config in ("c1", "c2") else 30
    # stdout carries exactly ONE JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION in this image) off it
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), __file__,
               *sys.argv[1:]]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
