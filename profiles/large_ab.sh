#!/bin/bash
# usage: large_ab.sh lib.so ...  -- chain-only time of 40 `large` fits (cycle counters of the timing build where present)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
C=nerf-attention_b200/csrc
for lib in "$@"; do
  echo "== $lib"
  NERFATTN_LIB=$PWD/$C/$lib NERFATTN_PHASE=1 python profiles/prof_fit.py large 40 3 2>&1 | grep "chain timing" | tail -2
  NERFATTN_LIB=$PWD/$C/$lib python - <<'PY'
import os, sys, time
sys.path.insert(0, 'nerf-attention_b200'); sys.path.insert(0, 'profiles')
import torch, nerf_attention as na
from nerf_attention.extract import synthetic_head
from nerf_attention import batched
cfg = next(c for c in na.CONFIGS_FULL if c.name == 'large')
tensors = [synthetic_head(16, h % 8, 2048, 32, 8, 128)[h // 8 % 2] for h in range(16)]
torch.manual_seed(0)
jobs = [na.FitJob(tensors[i % 16], cfg) for i in range(40)]
for mask in ('1', '7'):
    os.environ['NERFATTN_PHASE'] = mask
    b = batched.FitBatch(jobs, epochs=100, device='cuda', precision='bf16', keep_initial=True)
    best = None
    for _ in range(3):
        b.reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.launch(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    print('phase mask', mask, 'ms per epoch %.4f' % (best / 100))
PY
done
