#!/bin/bash
# round 2, session 2: diagnostic pass -- GPU tests, sensitivity builds, per-step cycle counters, launch list, ncu captures
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
C=nerf-attention_b200/csrc
python -m pytest tests -m gpu -x -q > $O/r2b_gpu_tests.log 2>&1; tail -3 $O/r2b_gpu_tests.log
run() {  # tag lib [env...]
  tag=$1; lib=$2; shift 2
  env NERFATTN_LIB=$PWD/$C/$lib NERFATTN_PROF_LIB=$PWD/$C/$lib "$@" python bench.py --steps 1 --warmup 1 --epochs 400 --no-e2e --no-extras > $O/r2b_$tag.json 2> $O/r2b_$tag.err
  python - $tag $O/r2b_$tag.json <<'PY' || tail -5 $O/r2b_$tag.err
import json, sys
d = json.load(open(sys.argv[2]))
ph = d["roofline"].get("phases_ms_per_epoch") or {}
print(sys.argv[1], "| fit-epochs/s", round(d["value"]), "| phases", {k: round(v, 4) for k, v in ph.items() if k.endswith("per_epoch")},
      "| cos", round(d["quality"]["cos_keys_mean"], 5), "| clk", d["clocks"]["sm_mhz"])
PY
}
run base libnerfattn_prof.so
run nol0load exp_NOL0LOAD.so
run pairtiles exp_PAIRTILES.so
run nostore libnerfattn_prof.so NERFATTN_CHAIN_DBG=1
run base2 libnerfattn_prof.so
for a in "medium 120" "large 40" "deep 40"; do
  set -- $a
  NERFATTN_LIB=$PWD/$C/exp_TIMING.so NERFATTN_PHASE=1 python profiles/prof_fit.py $1 $2 3 > $O/r2b_timing_$1.log 2>&1
  grep "chain timing" $O/r2b_timing_$1.log | tail -4
done
# launch list of the current step (serialised, cold cache: shares only)
NERFATTN_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
  --log-file $O/r2b_launches.csv python bench.py --steps 1 --warmup 0 --epochs 3 --no-e2e --no-extras > $O/r2b_launches.log 2>&1
tail -2 $O/r2b_launches.log
cap() {  # tag regex arch nfits
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o $O/r2b_$1 -f python profiles/prof_fit.py $3 $4 3 > $O/r2b_ncu_$1.log 2>&1
  tail -2 $O/r2b_ncu_$1.log
  python profiles/ncu_summary.py $O/r2b_$1.ncu-rep "$1: $3 x $4 fits (prof_fit.py), ncu --set full --clock-control none" > $O/r2b_ncu_$1.txt 2>&1
  python profiles/ncu_source_lines.py $O/r2b_$1.ncu-rep 60 > $O/r2b_ncu_$1_lines.txt 2>&1
}
cap chain256 chain_kernel medium 120
cap chain512 chain_kernel large 40
cap dwadam256 dw_adam_kernel medium 120
cap dwadam512 dw_adam_kernel large 40
rm -f $O/r2b_dwadam256.ncu-rep $O/r2b_dwadam512.ncu-rep
ls -la $O/r2b_*.ncu-rep
