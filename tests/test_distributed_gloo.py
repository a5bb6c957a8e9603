"""world_size-2 gloo run of the sweep's multi-rank host logic (CPU): shard planning is
consistent across ranks, the one all-gather reassembles every record exactly once."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_attention import fit as fit_mod, sharding
import nerf_attention as na


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    assert sharding.ensure_process_group('gloo')
    meta = na.KVMetadata('synthetic', 8, 4, 256, 16, 256)
    layers, heads, configs = fit_mod.sweep_selection(meta, quick=False)
    tensors = {l: {'keys': torch.zeros(4, 256, 16), 'values': torch.zeros(4, 256, 16)} for l in layers}
    jobs = fit_mod.enumerate_jobs(tensors, layers, heads, configs)
    keys = [(j['layer'], j['head'], j['kv_type']) for j in jobs]
    costs = [j['config'].flops_per_epoch(256, 16) for j in jobs]
    mine = sharding.shard_jobs(keys, costs, world)[rank]
    # stand-in for the per-fit results of this rank: the value encodes the job index
    local = []
    for i in mine:
        cfg = jobs[i]['config']
        rec = {k: 0 for k in fit_mod.RECORD_KEYS}
        rec.update(name=jobs[i]['name'], layer=jobs[i]['layer'], head=jobs[i]['head'], kv_type=jobs[i]['kv_type'],
                   config_name=cfg.name, hidden_features=cfg.hidden_features, hidden_layers=cfg.hidden_layers,
                   omega_0=cfg.omega_0, final_cosine_mean=i / 1000.0, final_mse=float(rank), seq_len=256, d_head=16)
        local.append((i, rec))
    records = fit_mod._gather_records(jobs, local, 'cpu')
    if rank == 0:
        np.save(os.path.join(out_dir, 'cos.npy'), np.array([r['final_cosine_mean'] for r in records]))
        np.save(os.path.join(out_dir, 'owner.npy'), np.array([r['final_mse'] for r in records]))
        with open(os.path.join(out_dir, 'names.txt'), 'w') as f:
            f.write('\n'.join(r['name'] for r in records))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cos = np.load(tmp_path / 'cos.npy')
    owner = np.load(tmp_path / 'owner.npy')
    names = (tmp_path / 'names.txt').read_text().split('\n')
    assert len(cos) == 5 * 4 * 2 * 7 == len(names)
    assert np.allclose(cos, np.arange(280) / 1000.0)            # every job exactly once, in job order
    assert names[0] == 'L0_H0_key_tiny' and names[-1] == 'L7_H3_value_lofreq'
    assert set(owner) == {0.0, 1.0} and abs((owner == 0).sum() - 140) <= 7
