#!/bin/bash
# usage: ab.sh <prefix> tag=lib.so[,ENV=..,ENV=..] ...   -- one short sweep (400 epochs, phases) per variant, one line each
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out; C=nerf-attention_b200/csrc; pre=$1; shift
for spec in "$@"; do
  tag=${spec%%=*}; rest=${spec#*=}; lib=${rest%%,*}; envs=""
  if [[ "$rest" == *,* ]]; then envs=$(echo "${rest#*,}" | tr ',' ' '); fi
  env NERFATTN_LIB=$PWD/$C/$lib NERFATTN_PROF_LIB=$PWD/$C/$lib $envs python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e --no-extras > $O/${pre}_$tag.json 2> $O/${pre}_$tag.err
  python - $tag $O/${pre}_$tag.json <<'PY' || tail -5 $O/${pre}_$tag.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1], "| fit-epochs/s", round(d["value"]), "| phases", {k.replace('_ms_per_epoch',''): round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")},
      "| cos", round(d["quality"]["cos_keys_mean"], 5), "| clk", d["clocks"]["sm_mhz"])
PY
done
