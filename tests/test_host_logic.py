"""Host-side logic of the drop-in package, on the CPU: API surface, packing order,
job enumeration, record / checkpoint schemas, shard planner, C-ABI symbol table.
No compute kernel is called here."""

import ctypes
import hashlib
import inspect
import json
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import nerf_attention as na
from nerf_attention import _native, batched, fit as fit_mod, sharding
from nerf_attention.extract import synthetic_layer
from oracle import siren_oracle as orc

ROOT = Path(__file__).resolve().parent.parent


def digest(state):
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


# ------------------------------------------------------------------ API surface
def test_config_tables_match_reference(golden):
    assert [[c.hidden_features, c.hidden_layers, c.omega_0, c.name] for c in na.CONFIGS_FULL] == \
        golden['meta']['configs_full']
    assert [[c.hidden_features, c.hidden_layers, c.omega_0, c.name] for c in na.CONFIGS_QUICK] == \
        golden['meta']['configs_quick']
    assert na.SIRENConfig() == na.SIRENConfig(256, 2, 30.0, 'medium')
    with pytest.raises(Exception):
        na.SIRENConfig().hidden_features = 3           # frozen, like the reference


def test_signatures_match_reference():
    sig = inspect.signature(na.fit_siren)
    names = list(sig.parameters)
    assert names[:7] == ['kv_tensor', 'config', 'epochs', 'lr', 'device', 'log_every', 'verbose']
    assert [sig.parameters[n].default for n in names[2:7]] == [5000, 1e-4, 'cuda', 500, True]
    sig = inspect.signature(na.fit_kv_cache)
    assert list(sig.parameters)[:5] == ['kv_dir', 'output_dir', 'epochs', 'device', 'quick']
    assert sig.parameters['epochs'].default == 5000 and sig.parameters['quick'].default is False
    assert list(inspect.signature(na.profile_latency).parameters) == ['siren_dir', 'output_dir', 'device']
    assert list(inspect.signature(na.SineLayer.__init__).parameters)[1:] == \
        ['in_features', 'out_features', 'omega_0', 'is_first']
    fields = [f.name for f in na.FitResult.__dataclass_fields__.values()]
    assert fields == ['model', 'config', 'target_mean', 'target_std', 'losses', 'final_mse',
                      'final_cosine_mean', 'final_cosine_min', 'final_cosine_std', 'per_pos_mse',
                      'cosine_sims', 'compression_ratio', 'raw_size_bytes', 'siren_size_bytes',
                      'train_time_seconds', 'seq_len', 'd_head', 'num_parameters']


def test_seeded_model_is_bit_identical_to_reference(golden):
    cases = golden['meta']['init_seed1234_d128']
    for cfg in na.CONFIGS_FULL:
        torch.manual_seed(1234)
        model = na.SIREN(cfg, out_features=128)
        sd = model.state_dict()
        assert list(sd.keys()) == cases[cfg.name]['keys']
        assert digest(sd) == cases[cfg.name]['digest']
        assert model.count_parameters() == cases[cfg.name]['num_parameters'] == cfg.param_count(128)
        assert model.size_bytes() == cases[cfg.name]['size_bytes']


def test_model_forward_matches_reference_on_cpu(golden):
    torch.manual_seed(77)
    model = na.SIREN(na.CONFIGS_FULL[2], out_features=128)
    with torch.no_grad():
        out = model(torch.linspace(0, 1, 96).unsqueeze(1)).numpy()
    assert np.abs(out - golden['forward']['fwd_medium']).max() < 1e-6


def test_packing_is_state_dict_order_and_adopt_round_trips():
    torch.manual_seed(3)
    cfg = na.SIRENConfig(64, 2, 30.0, 'x')
    model = na.SIREN(cfg, out_features=16)
    flat = torch.empty(model.count_parameters())
    batched.pack_model(model, flat)
    expect = torch.cat([v.reshape(-1) for v in model.state_dict().values()])
    assert torch.equal(flat, expect)
    lib = _native.lib()
    assert lib.nerfattn_param_count(64, 2, 16) == model.count_parameters()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    batched.adopt_packed(model, flat.clone())
    for k, v in model.state_dict().items():
        assert torch.equal(v, before[k])


def test_lr_schedule_equals_oracle():
    assert np.array_equal(batched.lr_schedule(37, 1e-4), np.asarray(orc.lr_table(37, 1e-4)))


# ------------------------------------------------------------------ synthetic inputs
def test_synthetic_generator_is_bit_identical_to_reference(golden, tmp_path):
    for layer in range(3):
        blob = synthetic_layer(layer, 48, 3, 2, 8)
        assert np.array_equal(blob['keys'].numpy(), golden['synthetic'][f'keys_{layer}'])
        assert np.array_equal(blob['values'].numpy(), golden['synthetic'][f'values_{layer}'])
    meta = na.extract_kv_cache_synthetic(seq_len=48, num_layers=3, num_kv_heads=2, head_dim=8,
                                         output_dir=tmp_path)
    assert meta.to_dict() == golden['meta']['synthetic_metadata']
    assert json.loads((tmp_path / 'metadata.json').read_text()) == golden['meta']['synthetic_metadata']
    blob = torch.load(tmp_path / 'layer_02.pt', weights_only=True)
    assert set(blob) == {'keys', 'values'} and blob['keys'].shape == (2, 48, 8)


# ------------------------------------------------------------------ sweep enumeration
def _fake_layers(layers, heads=8, n=4, d=4):
    return {l: {'keys': torch.zeros(heads, n, d), 'values': torch.ones(heads, n, d)} for l in layers}


def test_full_sweep_is_280_jobs_in_reference_order():
    meta = na.KVMetadata('synthetic', 32, 8, 2048, 128, 2048)
    layers, heads, configs = fit_mod.sweep_selection(meta, quick=False)
    assert layers == [0, 8, 16, 24, 31] and heads == 4 and configs is na.CONFIGS_FULL
    jobs = fit_mod.enumerate_jobs(_fake_layers(layers), layers, heads, configs)
    assert len(jobs) == 280
    assert [j['name'] for j in jobs[:8]] == [
        'L0_H0_key_tiny', 'L0_H0_key_small', 'L0_H0_key_medium', 'L0_H0_key_large', 'L0_H0_key_deep',
        'L0_H0_key_hifreq', 'L0_H0_key_lofreq', 'L0_H0_value_tiny']
    assert jobs[-1]['name'] == 'L31_H3_value_lofreq'
    assert jobs[7]['tensor'].sum() > 0 and jobs[0]['tensor'].sum() == 0     # values vs keys


def test_quick_sweep_and_missing_layer():
    meta = na.KVMetadata('synthetic', 4, 4, 512, 128, 512)
    layers, heads, configs = fit_mod.sweep_selection(meta, quick=True)
    assert layers == [0, 2, 3] and heads == 1 and len(configs) == 2
    jobs = fit_mod.enumerate_jobs(_fake_layers([0, 3], heads=4), layers, heads, configs)   # layer 2 missing
    assert len(jobs) == 8 and {j['layer'] for j in jobs} == {0, 3}
    tiny = na.KVMetadata('synthetic', 2, 2, 64, 8, 64)
    assert fit_mod.sweep_selection(tiny, quick=False)[0] == [0, 1]                        # de-duplicated


def test_record_and_checkpoint_schema(tmp_path):
    cfg = na.CONFIGS_FULL[2]
    torch.manual_seed(0)
    model = na.SIREN(cfg, 8)
    res = na.FitResult(model=model, config=cfg, target_mean=torch.zeros(1, 8), target_std=torch.ones(1, 8),
                       losses=[1.0], final_mse=0.1, final_cosine_mean=0.9, final_cosine_min=0.5,
                       final_cosine_std=0.05, per_pos_mse=np.zeros(4), cosine_sims=np.ones(4),
                       compression_ratio=2.0, raw_size_bytes=64, siren_size_bytes=32, train_time_seconds=1.5,
                       seq_len=4, d_head=8, num_parameters=model.count_parameters())
    rec = fit_mod._result_to_record('L0_H0_key_medium', 0, 0, 'key', res)
    assert list(rec) == ['name', 'layer', 'head', 'kv_type', 'config_name', 'hidden_features', 'hidden_layers',
                         'omega_0', 'final_mse', 'final_cosine_mean', 'final_cosine_min', 'final_cosine_std',
                         'compression_ratio', 'raw_size_bytes', 'siren_size_bytes', 'train_time_seconds',
                         'num_parameters', 'seq_len', 'd_head']
    json.dumps(rec)
    fit_mod._save_model(tmp_path, rec['name'], res, rec)
    ckpt = torch.load(tmp_path / 'L0_H0_key_medium_model.pt', weights_only=True)
    assert set(ckpt) == {'model_state', 'config', 'target_mean', 'target_std', 'metrics'}
    assert ckpt['config'] == {'hidden_features': 256, 'hidden_layers': 2, 'omega_0': 30.0, 'name': 'medium',
                              'out_features': 8}
    from nerf_attention.evaluate import _load_model_from_checkpoint
    again = _load_model_from_checkpoint(ckpt, 'cpu')
    for a, b in zip(again.state_dict().values(), model.state_dict().values()):
        assert torch.equal(a, b)


# ------------------------------------------------------------------ sharding
def test_shard_planner_partitions_and_balances():
    costs = {(l, h, kv): float(1 + (l % 3)) for l in range(5) for h in range(4) for kv in 'kv'}
    for world in (1, 2, 3, 4, 8):
        owned = sharding.plan_shards(costs, world)
        flat = [k for r in owned for k in r]
        assert sorted(flat) == sorted(costs) and len(set(flat)) == len(flat)
        loads = [sum(costs[k] for k in r) for r in owned]
        assert max(loads) - min(loads) <= max(costs.values())          # LPT bound
    assert sharding.plan_shards(costs, 4) == sharding.plan_shards(costs, 4)        # deterministic


def test_shard_jobs_keeps_units_together():
    meta = na.KVMetadata('synthetic', 32, 8, 2048, 128, 2048)
    layers, heads, configs = fit_mod.sweep_selection(meta, False)
    jobs = fit_mod.enumerate_jobs(_fake_layers(layers), layers, heads, configs)
    keys = [(j['layer'], j['head'], j['kv_type']) for j in jobs]
    costs = [j['config'].flops_per_epoch(2048, 128) for j in jobs]
    per_rank = sharding.shard_jobs(keys, costs, 8)
    assert sorted(i for r in per_rank for i in r) == list(range(280))
    assert [len(r) for r in per_rank] == [35] * 8                      # 40 units of 7 fits over 8 ranks
    for r in per_rank:
        assert len({keys[i] for i in r}) * 7 == len(r)
    assert sum(costs) == pytest.approx(1.334e15 / 2000, rel=1e-3)      # 1.334 PFLOP per sweep


# ------------------------------------------------------------------ the C ABI
def test_library_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'nerfattn.h').read_text()
    declared = set(re.findall(r'\b(nerfattn_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_native.EXPORTS)
    handle = ctypes.CDLL(str(_native.library_path()))
    for name in declared:
        assert hasattr(handle, name), name
    assert _native.lib().nerfattn_abi_version() == int(re.search(r'NERFATTN_ABI_VERSION (\d+)', header).group(1))
    assert ctypes.sizeof(_native.NaFit) == 4 * 6 + 8 * 11


def test_argument_validation_without_a_gpu():
    lib = _native.lib()
    need = ctypes.c_size_t(0)
    fits = (_native.NaFit * 1)()
    assert lib.nerfattn_fit_workspace_bytes(fits, 0, 0, ctypes.byref(need)) == -1          # NA_ERR_INVALID
    f = fits[0]
    f.N, f.D, f.H, f.L, f.omega0 = 2048, 128, 256, 2, 30.0
    f.positions = f.targets = f.params = 1 << 20                                            # never dereferenced
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 1, ctypes.byref(need)) == -2          # precision code 1 is unassigned
    assert b'precision' in lib.nerfattn_last_error()
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 0, ctypes.byref(need)) == 0
    fp32_bytes = need.value
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 2, ctypes.byref(need)) == 0
    assert 0 < need.value and fp32_bytes > 8 * 2048 * 256 * 4
    f.N = 2000                                                                              # any N in both modes
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 2, ctypes.byref(need)) == 0
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 0, ctypes.byref(need)) == 0
    f.D = 48                                                                                # bf16: D in {64, 128, 256}
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 2, ctypes.byref(need)) == -2
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 0, ctypes.byref(need)) == 0
    f.D = 128
    f.H = 30
    assert lib.nerfattn_fit_workspace_bytes(fits, 1, 0, ctypes.byref(need)) == -2


def test_no_cpu_fallback():
    with pytest.raises(_native.NativeError, match='no CPU fallback'):
        na.fit_siren(torch.randn(16, 4), na.SIRENConfig(8, 1, 30.0, 't'), epochs=1, device='cpu', verbose=False)
    assert _native.precision_code('bf16') == 2 and _native.precision_code(None) in (0, 2)
    with pytest.raises(ValueError):
        _native.precision_code('fp64')
    pkg = Path(na.__file__).parent
    for src in pkg.glob('*.py'):
        assert 'oracle' not in src.read_text().replace('siren_oracle', 'oracle') or src.name == '__init__.py', \
            f'{src.name} must not reference the oracle'


def test_crossover_fit_reproduces_reference_published_numbers():
    """experiments.scaling.crossover_data against the reference's own committed results: its
    scaling_results.json latency columns in, its crossover_data.json out (scaling.py:279-291)."""
    from nerf_attention.experiments.scaling import crossover_data
    gold = json.loads((Path(__file__).parent / 'golden' / 'reference_scaling.json').read_text())
    rows = {int(k): v for k, v in gold['scaling_results'].items()}
    got, want = crossover_data(rows), gold['crossover_data']
    assert got['siren_fit_log_slope'] == pytest.approx(want['siren_fit_log_slope'], rel=1e-9)
    assert got['siren_fit_log_intercept'] == pytest.approx(want['siren_fit_log_intercept'], rel=1e-9)
    assert got['latency_ratio_range'] == pytest.approx(want['latency_ratio_range'], rel=1e-9)
    assert got['crossover_4060_tokens'] == pytest.approx(want['crossover_4060_tokens'], rel=1e-6)
    assert got['crossover_h100_tokens'] == pytest.approx(want['crossover_h100_tokens'], rel=1e-6)
    assert got['siren_scaling'] == want['siren_scaling']


def test_async_layer_loader_and_job_specs(tmp_path):
    """layer_XX.pt files (reference format) come back unchanged from the background loader, and the tensor-free
    job list is the reference's loop order (fit.py:54-65)."""
    from nerf_attention.fit import enumerate_job_specs, enumerate_jobs, load_layers_async
    blobs = {l: {'keys': torch.randn(2, 24, 8), 'values': torch.randn(2, 24, 8)} for l in (0, 2)}
    for l, blob in blobs.items():
        torch.save(blob, tmp_path / f'layer_{l:02d}.pt')
    loaded = {l: f.result() for l, f in load_layers_async(tmp_path, [0, 2], pin=False).items()}
    for l in blobs:
        assert torch.equal(loaded[l]['keys'], blobs[l]['keys']) and torch.equal(loaded[l]['values'], blobs[l]['values'])
    assert load_layers_async(tmp_path, [], pin=False) == {}
    specs = enumerate_job_specs([0, 2], 2, na.CONFIGS_QUICK)
    jobs = enumerate_jobs(loaded, [0, 1, 2], 2, na.CONFIGS_QUICK)           # layer 1 has no file: skipped
    assert [j['name'] for j in specs] == [j['name'] for j in jobs]
    assert specs[0]['name'] == 'L0_H0_key_small' and specs[1]['name'] == 'L0_H0_key_medium'
    assert specs[2]['name'] == 'L0_H0_value_small' and specs[-1]['name'] == 'L2_H1_value_medium'
    assert all(s['tensor'] is None for s in specs)
    assert torch.equal(jobs[2]['tensor'], blobs[0]['values'][0]) and torch.equal(jobs[-1]['tensor'], blobs[2]['values'][1])


def test_checkpoint_holds_only_its_own_weights(tmp_path):
    """After a batched fit every model's parameters are views into one flat buffer with all jobs' weights; a
    checkpoint must still be ~4 * num_parameters bytes (the reference's size), not the whole buffer."""
    from nerf_attention.types import FitResult
    cfg = na.SIRENConfig(64, 1, 30.0, 'medium')
    model = na.SIREN(cfg, out_features=16)
    p = model.count_parameters()
    big = torch.zeros(4_000_000)                               # stands in for FitBatch.params.buf (16 MB)
    batched.pack_model(model, big[1000:1000 + p])
    batched.adopt_packed(model, big[1000:1000 + p])
    assert model.network[0].linear.weight.untyped_storage().nbytes() == big.numel() * 4     # really a view
    result = FitResult(model=model, config=cfg, target_mean=torch.zeros(1, 16), target_std=torch.ones(1, 16), losses=[0.1],
                       final_mse=0.0, final_cosine_mean=1.0, final_cosine_min=1.0, final_cosine_std=0.0,
                       per_pos_mse=np.zeros(8, np.float32), cosine_sims=np.ones(8, np.float32), compression_ratio=1.0,
                       raw_size_bytes=256, siren_size_bytes=4 * p, train_time_seconds=0.0, seq_len=8, d_head=16,
                       num_parameters=p)
    record = fit_mod._result_to_record('L0_H0_key_medium', 0, 0, 'key', result)
    fit_mod._save_model(tmp_path, 'L0_H0_key_medium', result, record)
    size = (tmp_path / 'L0_H0_key_medium_model.pt').stat().st_size
    assert 4 * p <= size <= 4 * p + 16384, (size, 4 * p)
    ckpt = torch.load(tmp_path / 'L0_H0_key_medium_model.pt', weights_only=True)
    assert all(torch.equal(ckpt['model_state'][k], v) for k, v in model.state_dict().items())
    from nerf_attention.experiments import scaling
    scaling._save_scaling_checkpoint(tmp_path / 's.pt', 'L0_H0_K', result, 8)
    assert (tmp_path / 's.pt').stat().st_size <= 4 * p + 16384


def test_reference_export_names_exist_and_say_what_they_are():
    """Every name the reference package exports (reference nerf_attention/__init__.py) imports; the figure functions and
    the real-LLM extraction are outside the hot path and raise a clear error instead of an ImportError."""
    reference_exports = ['CONFIGS_FULL', 'CONFIGS_QUICK', 'AnalysisResult', 'FitResult', 'KVMetadata', 'LayerSummary',
                         'SIRENConfig', 'SIREN', 'SineLayer', 'fit_siren', 'extract_kv_cache', 'extract_kv_cache_synthetic',
                         'analyze_kv_cache', 'fit_kv_cache', 'load_results', 'plot_pareto_frontier', 'plot_keys_vs_values',
                         'plot_per_position_error', 'profile_latency', 'generate_summary_figure']
    for name in reference_exports:
        assert hasattr(na, name), name
    for name in ('plot_pareto_frontier', 'plot_keys_vs_values', 'generate_summary_figure'):
        with pytest.raises(NotImplementedError, match='outside the scope'):
            getattr(na, name)([], Path('.'))


def test_bench_strong_scaling_shards_partition_the_one_sweep():
    """bench.py --gpus N shards ONE 280-fit sweep (BASELINE config 3) by (layer, head, key|value) unit: every fit exactly
    once over the ranks, 280 / N per rank, units kept whole; --scaling weak gives every rank a whole sweep."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', ROOT / 'bench.py')
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    full = bench.sweep_specs(0, 1, 'strong', 2048)
    assert len(full) == 280 and full[0] == (0, 0, 0, 0) and full[-1] == (31, 3, 1, 6)
    for world in (2, 4, 8):
        shards = [bench.sweep_specs(r, world, 'strong', 2048) for r in range(world)]
        assert sorted(s for sh in shards for s in sh) == sorted(full)
        assert all(len(sh) == 280 // world for sh in shards)
        for sh in shards:                                         # a unit's 7 architectures stay on one rank
            units = {}
            for layer, head, is_value, ci in sh:
                units.setdefault((layer, head, is_value), []).append(ci)
            assert all(sorted(v) == list(range(7)) for v in units.values())
        weak = bench.sweep_specs(1, world, 'weak', 2048)
        assert len(weak) == 280 and weak[0][0] == 1
