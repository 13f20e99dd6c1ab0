#!/bin/bash
# usage: profiles/scale_run.sh N [configs...]  -- bench.py on N GPUs of one box for the given configs (default c2 c5), one
# summary line per run; full JSON lines in gpurun_out/scale_N<N>_<config>.json
N=$1; shift; CFGS=${*:-c2 c5}
touch realtime_urdf_filter_b200/*.so
P=$((29800 + N * 10))
for c in $CFGS; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --config $c --steps 10 --warmup 3 --e2e-frames 512 2>gpurun_out/scale_N${N}_$c.err > gpurun_out/scale_N${N}_$c.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_N${N}_$c.json")); r=d["roofline"]; e=d["e2e"]; pm=e.get("packed_mask") or {}
    print("$c N=$N %s fps=%.0f path_frac=%.3f e2e=%.0f ceil=%.0f (%.2f) packed=%.0f/%.0f gather=%s clocks=%s" % (d["scaling"], d["value"], r["path_frac"], e["value"], e["copy_ceiling"], e["frac_of_copy_ceiling"], pm.get("value",0), pm.get("copy_ceiling",0), d.get("ordered_gather_matches_single_gpu"), d["clocks"]["sm_mhz"]))
except Exception as ex:
    print("$c N=$N failed:", ex); print(open("gpurun_out/scale_N${N}_$c.err").read()[-800:])
PY
  P=$((P+1))
done
