"""Full pipeline on synthetic data, B200 path (reference quickstart.py:18-68, BASELINE config 1).

Steps 1 and 3 of the reference script -- synthetic KV cache, quick SIREN sweep (layers {0, 2, 3} x
head 0 x key/value x {small, medium}, 2000 epochs at 512 tokens) -- plus the latency profile of the
saved models.  The structure analysis and the figures of the reference (analyze.py, matplotlib) are
outside the hot-path build.  There is no CPU mode: `--cpu` (the reference's flag) fails loudly,
the CPU numbers of this configuration come from `bench.py --impl reference`.
"""

import argparse
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))

from nerf_attention import extract_kv_cache_synthetic, fit_kv_cache, load_results, profile_latency  # noqa: E402


def main() -> None:
    parser = argparse.ArgumentParser()
    parser.add_argument('--cpu', action='store_true', help='Force CPU mode (not available in this build)')
    parser.add_argument('--epochs', type=int, default=2000)
    parser.add_argument('--precision', choices=['fp32', 'bf16'], default=None)
    args = parser.parse_args()
    device = 'cpu' if args.cpu else 'cuda'
    print(f"Device: {device}\n")

    kv_dir = Path('results/kv_cache_quick')
    fits_dir = Path('results/fits_quick')
    figures_dir = Path('results/figures_quick')

    print("=" * 60 + "\nSTEP 1: Generate synthetic KV cache\n" + "=" * 60)
    extract_kv_cache_synthetic(seq_len=512, num_layers=4, num_kv_heads=4, head_dim=128, output_dir=kv_dir)

    print("\n" + "=" * 60 + "\nSTEP 3: Fit SIRENs (quick mode)\n" + "=" * 60)
    fit_kv_cache(kv_dir=kv_dir, output_dir=fits_dir, epochs=args.epochs, device=device, quick=True,
                 precision=args.precision)

    print("\n" + "=" * 60 + "\nSTEP 4: Evaluate\n" + "=" * 60)
    results = load_results(fits_dir)
    print(f"  {len(results)} fits, mean CosSim {sum(r['final_cosine_mean'] for r in results) / len(results):.4f}")
    profile_latency(fits_dir, figures_dir, device=device)
    print("\n" + "=" * 60 + "\nDONE!\n" + "=" * 60)
    print(f"\nResults in: {fits_dir}/ and {figures_dir}/")


if __name__ == '__main__':
    main()
