"""CPU suite, part 3: the N>1 path (graph sharding, loss weighting, flat gradient bucket, one all-reduce) on
world_size-2 gloo, driven with the oracle model as the compute stand-in (the CUDA product cannot run here)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(seed=0):
    from oracle import gsat_oracle as O
    cfg = {'model_name': 'GIN', 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.0, 'use_edge_attr': False}
    torch.manual_seed(seed)
    clf = O.get_model(10, 0, 2, False, cfg)
    ext = O.ExtractorMLP(16, {'learn_edge_att': True, 'extractor_dropout_p': 0.0})
    g = O.GSAT(clf, ext, O.Criterion(2, False), learn_edge_att=True, final_r=0.5)
    g.train()
    for m in clf.modules():                       # BatchNorm batch statistics are shard-local by design (DDP
        if isinstance(m, torch.nn.BatchNorm1d):   # semantics); use running stats so 1-rank == 2-rank exactly
            m.eval()
    return g


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from dp_gsat_b200.data import ba2motifs_batch, shard_batch
    from dp_gsat_b200.parallel import TrainStep, broadcast_parameters
    torch.set_num_threads(1)
    full = ba2motifs_batch(12, seed=5)
    full.x = torch.rand(full.x.shape, generator=torch.Generator().manual_seed(3))
    u_full = torch.rand(full.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-6, 1 - 1e-6)
    g = _make(seed=rank)                          # different init per rank -> broadcast must fix it
    broadcast_parameters(g.clf)
    broadcast_parameters(g.extractor)
    shard = shard_batch(full, rank, world)
    e0 = sum(shard_batch(full, r, world).num_edges for r in range(rank))
    step = TrainStep(g, lr=1e-2, fused_adam=False)
    step(shard, 0, noise_u=u_full[e0:e0 + shard.num_edges])
    if rank == 0:
        torch.save({'flat': step.bucket.flat.clone(), 'w': g.clf.convs[0].nn[0].weight.detach().clone()}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_rank(tmp_path):
    from dp_gsat_b200.data import ba2motifs_batch
    from dp_gsat_b200.parallel import TrainStep
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    full = ba2motifs_batch(12, seed=5)
    full.x = torch.rand(full.x.shape, generator=torch.Generator().manual_seed(3))
    u_full = torch.rand(full.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-6, 1 - 1e-6)
    g = _make(seed=0)
    step = TrainStep(g, lr=1e-2, fused_adam=False)
    step(full, 0, noise_u=u_full)
    assert torch.allclose(got['flat'], step.bucket.flat, rtol=1e-4, atol=1e-6)
    assert torch.allclose(got['w'], g.clf.convs[0].nn[0].weight.detach(), rtol=1e-4, atol=1e-6)


def test_flat_bucket_views():
    from dp_gsat_b200.parallel import FlatGradBucket
    lin = torch.nn.Linear(3, 2)
    b = FlatGradBucket(lin.parameters())
    lin(torch.ones(1, 3)).sum().backward()
    assert b.flat.numel() == 8 and torch.equal(b.flat[:6].view(2, 3), lin.weight.grad)
    assert lin.weight.grad.data_ptr() == b.flat.data_ptr()


def _preload_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import importlib.util
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('_bench', os.path.join(root, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    calls = []
    buf = torch.zeros(1)

    def step():                      # a "training step": different speed per rank, one collective inside
        time.sleep(0.002 * (1 + 3 * rank))
        dist.all_reduce(buf)
        calls.append(1)

    n = bench.preload(step, dist.barrier, world, torch.device('cpu'), seconds=0.08)
    dist.barrier()
    out[rank] = (n, len(calls))
    dist.destroy_process_group()


def test_bench_preload_runs_the_same_number_of_steps_on_every_rank():
    """bench.preload keeps the GPU loaded before the timed region; every step holds collectives, so the number of extra
    steps must be agreed across ranks (a per-rank wall-clock loop dead-locked the 2-GPU bench once)."""
    world = 2
    out = mp.Manager().dict()
    mp.spawn(_preload_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0] == out[1]
    assert out[0][0] >= 1 and out[0][1] == out[0][0] + 2
