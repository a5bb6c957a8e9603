#!/bin/bash
# round 2, last session: GPU tests + smoke of the final tree, then one short sweep (400 epochs, live phase split) per
# scheduling knob -- shape groups split into units of at most U fits on their own graph branches (NERFATTN_UNIT_FITS)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r2c_gpu_tests.log 2>&1; tail -3 $O/r2c_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c_smoke.log 2>&1; tail -2 $O/r2c_smoke.log
i=0
for envs in "A=0" "NERFATTN_UNIT_FITS=40" "NERFATTN_UNIT_FITS=20" "NERFATTN_UNIT_FITS=10" "NERFATTN_UNIT_FITS=60" "A=1"; do
  i=$((i+1))
  env $envs python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e --no-extras > $O/r2c_knob_$i.json 2> $O/r2c_knob_$i.err
  python - "$envs" $O/r2c_knob_$i.json <<'PY' || tail -5 $O/r2c_knob_$i.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1], "| fit-epochs/s", round(d["value"]), "| ms/epoch %.4f" % (d["ms_per_step"] / d["config"]["epochs"]),
      "| phases", {k.replace('_ms_per_epoch', ''): round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")},
      "| cos", round(d["quality"]["cos_keys_mean"], 6), "| clk", d["clocks"]["sm_mhz"])
PY
done
