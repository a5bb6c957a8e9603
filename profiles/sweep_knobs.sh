#!/bin/bash
# usage: sweep_knobs.sh <tag> "<ENV=.. ENV=..>" ...   -- one short bench run per knob set, one summary line each
tag=$1; shift
i=0
for envs in "$@"; do
  i=$((i+1))
  env $envs python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$envs" gpurun_out/${tag}_$i.json <<'PY' || tail -5 gpurun_out/${tag}_$i.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1] or "default", "| fit-epochs/s", round(d["value"]), "| ms/epoch %.4f" % (d["ms_per_step"] / d["config"]["epochs"]),
      "| phases", {k: round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")}, "| cos", round(d["quality"]["cos_keys_mean"], 6),
      "| clk", d["clocks"]["sm_mhz"])
PY
done
