#!/bin/bash
# usage: knob_ab.sh <tag> "<ENV=.. ENV=..>" ...  -- one 400-epoch sweep (value + live phase split) per knob set, one line each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; tag=$1; shift; i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e --no-extras > $O/${tag}_$i.json 2> $O/${tag}_$i.err
  python - "$envs" $O/${tag}_$i.json <<'PY' | tee -a $O/${tag}.log || tail -5 $O/${tag}_$i.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1], "| fit-epochs/s", round(d["value"]), "| ms/epoch %.4f" % (d["ms_per_step"] / d["config"]["epochs"]),
      "| phases", {k.replace('_ms_per_epoch', ''): round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")},
      "| cos", round(d["quality"]["cos_keys_mean"], 6), "| clk", d["clocks"]["sm_mhz"])
PY
done
