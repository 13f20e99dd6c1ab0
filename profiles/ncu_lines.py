#!/usr/bin/env python
"""Per-source-line hot spots of one kernel: joins the SASS sampling page of an .ncu-rep (ncu --set full
--import-source on) with nvdisasm's line table of the library that was profiled (-lineinfo build).

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep ruf_setup_bin [top] [lib.so|-] [mangled-name substring]
"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER_LINES = {}      # (header file, line) -> id; the line table holds -id for instructions inlined from CUDA headers


def line_table(lib, kernel):
    """-> list of (source line) per instruction of the first function whose name contains `kernel`"""
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
        for cubin in sorted(glob.glob(os.path.join(td, "*.cubin"))):
            dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
            lines, cur, inside = [], 0, False
            for l in dis.splitlines():
                if l.startswith("//---") and ".text." in l:
                    if inside and lines:
                        return lines
                    inside = kernel in l
                    continue
                if not inside:
                    continue
                m = re.search(r'//## File "(.*?)", line (\d+)', l)
                if m:
                    # lines of inlined CUDA headers (shuffles, min/max, ...) must not alias lines of the kernel file:
                    # they are reported as negative numbers keyed by HEADER_LINES
                    if m.group(1).endswith("ruf_kernels.cu"):
                        cur = int(m.group(2))
                    else:
                        key = (os.path.basename(m.group(1)), int(m.group(2)))
                        cur = -HEADER_LINES.setdefault(key, len(HEADER_LINES) + 1)
                elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
                    lines.append(cur)
            if inside and lines:
                return lines
    return []


def main(rep, kernel, top=45, lib=None, mangled=None):
    """kernel: regex for ncu's (demangled) kernel name; mangled: substring of the mangled name in the cubin"""
    mangled = mangled or kernel
    lib = lib or os.path.join(ROOT, "realtime_urdf_filter_b200", "libruf_b200.so")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
    first = rows.index(hdr)
    # several captured launches of the same kernel follow each other: keep the first
    n_inst = 0
    a0 = int(body[0][0], 16)
    for r in body:
        if int(r[0], 16) == a0 and n_inst:
            break
        n_inst += 1
    body = body[:n_inst]
    lt = line_table(lib, mangled)
    if len(lt) != len(body):
        print(f"warning: {len(body)} profiled instructions vs {len(lt)} in {lib}: line numbers may be off", file=sys.stderr)
    col = {n: i for i, n in enumerate(hdr)}
    smp, inst = col["# Samples"], col["Instructions Executed"]
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    agg = collections.defaultdict(lambda: collections.Counter())
    for k, r in enumerate(body):
        ln = lt[k] if k < len(lt) else 0
        a = agg[ln]
        a["smp"] += float(r[smp] or 0)
        a["inst"] += float(r[inst] or 0)
        for n in stall_cols:
            a[n] += float(r[col[n]] or 0)
    tot_s = sum(a["smp"] for a in agg.values()) or 1.0
    tot_i = sum(a["inst"] for a in agg.values()) or 1.0
    src = open(os.path.join(ROOT, "realtime_urdf_filter_b200", "csrc", "ruf_kernels.cu")).read().splitlines()
    print(f"== {kernel}: {tot_s:.0f} samples, {tot_i:.0f} warp instructions executed, {len(body)} SASS instructions")
    for ln, a in sorted(agg.items(), key=lambda kv: -kv[1]["smp"])[:top]:
        s = a["smp"]
        if s == 0:
            break
        stalls = sorted(((a[n], n[6:]) for n in stall_cols), reverse=True)[:3]
        st = " ".join(f"{n}:{v / s * 100:.0f}%" for v, n in stalls if v > 0)
        text = src[ln - 1].strip()[:100] if 0 < ln <= len(src) else ""
        if ln < 0:
            text = "[%s:%d]" % next(k for k, v in HEADER_LINES.items() if v == -ln)
        print(f"{ln:5d} {s / tot_s * 100:5.1f}% smp {a['inst'] / tot_i * 100:5.1f}% inst  [{st}]  {text}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45, sys.argv[4] if len(sys.argv) > 4 and sys.argv[4] != "-" else None,
         sys.argv[5] if len(sys.argv) > 5 else None)
