#!/bin/bash
# usage: timing.sh arch nfits  -- per-step cycle counters of the chain kernel (exp_TIMING.so), last launch
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
NERFATTN_LIB=$PWD/nerf-attention_b200/csrc/exp_TIMING.so NERFATTN_PHASE=1 python profiles/prof_fit.py $1 $2 3 2>&1 | grep "chain timing" | tail -6
