#!/bin/bash
# branch order (longest tile first) and branch priorities of the epoch graph: shards of the strong-scaled sweep alone on
# one GPU (35 / 70 / 140 fits) and the whole 280-fit sweep, old behaviour (ORDER=0 PRIO=0) against the knobs
# (the NERFATTN_ORDER / NERFATTN_PRIO hooks existed only in the build this script measured: neutral, patch not kept --
#  DESIGN.md section 8; profiles/order_prio_ab_r02.log)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
for world in 8 4 2; do
  for envs in "NERFATTN_ORDER=0 NERFATTN_PRIO=0" "NERFATTN_ORDER=1 NERFATTN_PRIO=0" "NERFATTN_ORDER=1 NERFATTN_PRIO=1" "NERFATTN_ORDER=0 NERFATTN_PRIO=1"; do
    env $envs python profiles/shard_ab.py $world 0 400 single 2>&1 | grep variant | tail -1 | tee -a $O/r2d_shards.log
  done
done
i=0
for envs in "NERFATTN_ORDER=0 NERFATTN_PRIO=0" "NERFATTN_ORDER=1 NERFATTN_PRIO=0" "NERFATTN_ORDER=1 NERFATTN_PRIO=1" "NERFATTN_ORDER=0 NERFATTN_PRIO=0"; do
  i=$((i+1))
  env $envs python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e --no-extras > $O/r2d_knob_$i.json 2> $O/r2d_knob_$i.err
  python - "$envs" $O/r2d_knob_$i.json <<'PY' || tail -5 $O/r2d_knob_$i.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1], "| fit-epochs/s", round(d["value"]), "| ms/epoch %.4f" % (d["ms_per_step"] / d["config"]["epochs"]),
      "| phases", {k.replace('_ms_per_epoch', ''): round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")},
      "| cos", round(d["quality"]["cos_keys_mean"], 6), "| clk", d["clocks"]["sm_mhz"])
PY
done
