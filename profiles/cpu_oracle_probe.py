"""Where does the torch-CPU fp32 evaluation of a SIREN occasionally lose accuracy?  (test infrastructure probe)
Prints, per sine layer, the fp32-vs-fp64 error of the oracle forward for the case that once failed a parity test."""
import sys; sys.path.insert(0, 'tests'); sys.path.insert(0, 'nerf-attention_b200'); sys.path.insert(0, '.')
import torch
import torch.nn.functional as F
import nerf_attention as na
from gpu_util import seeded_state
from oracle import siren_oracle as orc
cfg = next(c for c in na.CONFIGS_FULL if c.name == 'medium')
n = 512
state = seeded_state(cfg, 128, 70)
sine, (wf, bf) = orc._layers(state)
x32 = orc.positions_for(n); x64 = x32.double()
h32, h64 = x32, x64
msg = [f'threads {torch.get_num_threads()}']
for li, (w, b) in enumerate(sine):
    z32 = F.linear(h32, w, b); z64 = F.linear(h64, w.double(), b.double())
    ez = (z32.double() - z64).abs().max().item()
    h32 = torch.sin(cfg.omega_0 * z32); h64 = torch.sin(cfg.omega_0 * z64)
    d = (h32.double() - h64).abs()
    msg.append(f'L{li}: z {ez:.1e} h {d.max().item():.1e} (#>1e-5: {int((d > 1e-5).sum())})')
y32 = F.linear(h32, wf, bf); y64 = F.linear(h64, wf.double(), bf.double())
msg.append(f'out {((y32.double() - y64).abs().max() / y64.abs().max()).item():.1e}')
print('; '.join(msg))
