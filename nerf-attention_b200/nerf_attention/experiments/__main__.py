"""`python -m nerf_attention.experiments {scaling,layer_profile}` (reference experiments/__main__.py:26-109).
multi_prompt needs a real LLM and is outside this build."""

import argparse
from pathlib import Path

from nerf_attention.experiments.scaling import run_full_layer_profile, run_scaling_experiment
from nerf_attention.experiments.svd import run_svd_experiment


def main() -> None:
    parser = argparse.ArgumentParser(description='Run follow-up experiments')
    parser.add_argument('experiment', choices=['scaling', 'multi_prompt', 'svd', 'layer_profile', 'all'])
    parser.add_argument('--model', type=str, default='synthetic')
    parser.add_argument('--device', type=str, default='cuda')
    parser.add_argument('--epochs', type=int, default=2000)
    parser.add_argument('--kv_dir', type=str, default='results/kv_cache')
    parser.add_argument('--siren_dir', type=str, default='results/fits')
    parser.add_argument('--precision', type=str, default=None, choices=['fp32', 'bf16'])
    args = parser.parse_args()
    if args.experiment == 'multi_prompt':
        raise SystemExit('multi_prompt: outside the scope of this build (needs a real LLM)')
    if args.experiment in ('scaling', 'all'):
        print('\n' + '=' * 60 + '\nEXPERIMENT 1: Sequence Length Scaling\n' + '=' * 60)
        run_scaling_experiment(model_name=args.model, seq_lengths=[512, 1024, 2048, 4096, 8192],
                               base_dir=Path('results/scaling'), device=args.device, epochs=args.epochs,
                               precision=args.precision)
    if args.experiment in ('svd', 'all'):
        print('\n' + '=' * 60 + '\nEXPERIMENT 3: SVD Baseline\n' + '=' * 60)
        run_svd_experiment(kv_dir=Path(args.kv_dir), base_dir=Path('results/svd'), device=args.device)
    if args.experiment in ('layer_profile', 'all'):
        print('\n' + '=' * 60 + '\nEXPERIMENT 4: Full Layer Profile\n' + '=' * 60)
        run_full_layer_profile(kv_dir=Path(args.kv_dir), output_dir=Path('results/layer_profile'),
                               device=args.device, epochs=args.epochs, precision=args.precision)


if __name__ == '__main__':
    main()
