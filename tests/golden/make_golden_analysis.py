"""Golden values for the structure analysis and the SVD baseline, from the REAL reference.

Run in the build container only:  python tests/golden/make_golden_analysis.py
Imports /root/reference/nerf_attention/{analyze,experiments/svd}.py with a stub matplotlib (absent in this image;
only the figure code touches it) and runs them on a small synthetic KV cache written by the reference's own
generator (seq_len 192, 4 layers x 2 heads x 16 dims).  Output: tests/golden/analysis.json.
"""

import contextlib
import io
import json
import sys
import tempfile
import types
from pathlib import Path
from unittest import mock

import torch

REF = Path('/root/reference')
OUT = Path(__file__).resolve().parent
SHAPE = dict(seq_len=192, num_layers=4, num_kv_heads=2, head_dim=16)


def import_reference():
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.gridspec'):
        sys.modules[name] = mock.MagicMock()
    pkg = types.ModuleType('nerf_attention')
    pkg.__path__ = [str(REF / 'nerf_attention')]
    sys.modules['nerf_attention'] = pkg
    import nerf_attention.analyze as ranalyze             # noqa: E402
    import nerf_attention.extract as rextract             # noqa: E402
    import nerf_attention.experiments.svd as rsvd         # noqa: E402
    return ranalyze, rextract, rsvd


def main() -> None:
    torch.set_num_threads(1)
    ranalyze, rextract, rsvd = import_reference()
    ranalyze._plot_analysis = lambda *a, **k: None
    out = {'shape': SHAPE, 'torch': torch.__version__}
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        kv_dir = Path(tmp) / 'kv'
        rextract.extract_kv_cache_synthetic(output_dir=kv_dir, **SHAPE)
        ranalyze.analyze_kv_cache(kv_dir, Path(tmp) / 'analysis')
        out['analysis_results'] = json.loads((Path(tmp) / 'analysis' / 'analysis_results.json').read_text())
        blob = torch.load(kv_dir / 'layer_02.pt', weights_only=True)
        out['tensor_L2_H1_K'] = ranalyze._analyze_tensor(blob['keys'][1], 'L2_H1_K')
        out['tensor_L2_H1_V'] = ranalyze._analyze_tensor(blob['values'][1], 'L2_H1_V')
        svd_fn = next(getattr(rsvd, n) for n in dir(rsvd) if n.startswith('run_svd') or n == 'svd_baseline')
        out['svd_function'] = svd_fn.__name__
        res = svd_fn(kv_dir, Path(tmp) / 'svd')
        out['svd_results'] = res if isinstance(res, list) else json.loads(next((Path(tmp) / 'svd').glob('*.json')).read_text())
    (OUT / 'analysis.json').write_text(json.dumps(out, indent=1))
    print('wrote', OUT / 'analysis.json')


if __name__ == '__main__':
    main()
