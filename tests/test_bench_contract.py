"""bench.py's reference arm runs without a GPU (the CPU oracle timed on the host's cores): one JSON line with the keys the
driver reads, the same `config` dict the B200 arm prints for that configuration, and no product library loaded."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["c1", "c2"])
def test_reference_arm_prints_one_contract_line(config):
    env = dict(os.environ, OMP_NUM_THREADS="1")        # what torchrun injects: the arm must ignore it and use the affinity mask
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config, "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["value"] == d["value"] and cb["sample"]
    # "reference" = the reference's GLSL path on Mesa llvmpipe (where oracle/_ref/gl_crosscheck_glx and the driver exist: one
    # rasteriser thread per online core, at most 16); "port" = the CPU oracle on every core of the affinity mask
    assert cb["cores"] == (min(16, os.cpu_count()) if cb["kind"] == "reference" else len(os.sched_getaffinity(0)))
    sys.path.insert(0, ROOT)
    import bench
    sc = bench.make_scene(bench.CONFIGS[config])
    assert d["config"] == bench.config_dict(config, bench.CONFIGS[config], len(sc.tri), sc.n_parts, 1)
    assert "libruf_b200" not in json.dumps(d.get("native_so_loaded", []))
