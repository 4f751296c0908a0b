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


# ---------------------------------------------------------------------------------------------------------------
# the PRODUCT model on two gloo ranks (kernels on the host SIMT emulator, tests/simt): loss weighting, flat bucket,
# all-reduce and -- with sync BatchNorm -- exact equality of the sharded and the single-device step
# ---------------------------------------------------------------------------------------------------------------
def _product(seed, model_name='GIN', precision='fp32'):
    import dp_gsat_b200 as G
    cfg = {'model_name': model_name, 'hidden_size': 16, 'n_layers': 2, 'dropout_p': 0.0, 'use_edge_attr': False,
           'aggregators': ['mean', 'min', 'max', 'std'], 'scalers': False, 'deg': torch.ones(10)}
    torch.manual_seed(seed)
    clf = G.get_model(10, 0, 2, False, cfg, 'cpu')
    ext = G.ExtractorMLP(16, {'learn_edge_att': True, 'extractor_dropout_p': 0.0})
    clf.precision = ext.precision = precision   # 'bf16': the tcgen05 kernels (on their host models under emulation)
    g = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5)
    g.train()                                    # BatchNorm in TRAINING mode: batch statistics matter
    return g


def _full_batch():
    from dp_gsat_b200.data import ba2motifs_batch
    full = ba2motifs_batch(12, seed=5)
    full.x = torch.rand(full.x.shape, generator=torch.Generator().manual_seed(3))
    u = torch.rand(full.num_edges, 1, generator=torch.Generator().manual_seed(1)).clamp(1e-6, 1 - 1e-6)
    return full, u


def _product_worker(rank, world, port, out, model_name, sync_bn, precision='fp32'):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from tests.simt import emulate
    emulate.patch_product(setattr)
    from dp_gsat_b200.data import shard_batch
    from dp_gsat_b200.parallel import TrainStep, broadcast_parameters
    torch.set_num_threads(1)
    full, u_full = _full_batch()
    g = _product(seed=rank, model_name=model_name, precision=precision)
    broadcast_parameters(g.clf)
    broadcast_parameters(g.extractor)
    shard = shard_batch(full, rank, world)
    e0 = sum(shard_batch(full, r, world).num_edges for r in range(rank))
    step = TrainStep(g, lr=1e-2, fused_adam=False, sync_bn=sync_bn)
    step(shard, 0, noise_u=u_full[e0:e0 + shard.num_edges])
    local = None
    if sync_bn:                                   # the same sharded step with shard-local statistics (DDP semantics)
        g2 = _product(seed=0, model_name=model_name, precision=precision)
        broadcast_parameters(g2.clf)
        broadcast_parameters(g2.extractor)
        step2 = TrainStep(g2, lr=1e-2, fused_adam=False)
        step2(shard, 0, noise_u=u_full[e0:e0 + shard.num_edges])
        local = step2.bucket.flat.clone()
    if rank == 0:
        bn = next(m for m in g.clf.modules() if isinstance(m, torch.nn.BatchNorm1d))
        torch.save({'flat': step.bucket.flat.clone(), 'running_var': bn.running_var.clone(),
                    'running_mean': bn.running_mean.clone(), 'flat_local_bn': local}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('model_name,precision', [('PNA', 'fp32'), ('GIN', 'bf16')])
def test_product_two_ranks_with_sync_batchnorm_equal_single_rank(tmp_path, monkeypatch, model_name, precision):
    """Graph-sharded step of the product model (emulated kernels) on 2 ranks with enable_sync_batchnorm == the
    single-rank step on the whole batch: every parameter gradient and the BatchNorm running statistics -- on the strict
    fp32 path and on the tensor-core path (statistics from the GEMM epilogue all-reduced on the device; every row goes
    through the same bf16 roundings in both runs, only the statistics' summation order differs)."""
    from tests.simt import emulate
    from dp_gsat_b200.parallel import TrainStep
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_product_worker, args=(2, _free_port(), out, model_name, True, precision), nprocs=2, join=True)
    got = torch.load(out)
    emulate.patch_product(monkeypatch.setattr)
    full, u_full = _full_batch()
    g = _product(seed=0, model_name=model_name, precision=precision)
    step = TrainStep(g, lr=1e-2, fused_adam=False)
    step(full, 0, noise_u=u_full)
    scale = float(step.bucket.flat.abs().max())
    rtol, atol = (2e-4, 2e-6) if precision == 'fp32' else (1e-4, 1e-5)      # measured: 1.5e-7 relative L2 (0.2 without sync)
    assert torch.allclose(got['flat'], step.bucket.flat, rtol=rtol, atol=atol * max(1.0, scale))
    bn = next(m for m in g.clf.modules() if isinstance(m, torch.nn.BatchNorm1d))
    assert torch.allclose(got['running_mean'], bn.running_mean, rtol=1e-5 if precision == 'fp32' else 1e-3, atol=1e-6)
    assert torch.allclose(got['running_var'], bn.running_var, rtol=1e-5 if precision == 'fp32' else 1e-3, atol=1e-6)
    # the default (shard-local statistics) is NOT the single-device step: the option actually changes the math
    assert not torch.allclose(got['flat_local_bn'], step.bucket.flat, rtol=2e-4, atol=1e-6)


def test_bench_control_flow_on_two_ranks_without_a_gpu():
    """bench.py's own b200 arm (warm-up, agreed pre-load, timed region, e2e loop, max over ranks, JSON line) on two gloo
    ranks with the kernels on the emulator (tools/bench_dry_run.py): every rank issues the same collectives and the line
    carries every contract key.  Timings are emulation and are not looked at."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'bench_dry_run.py'), '--ranks', '2', '--graphs', '48',
                        '--steps', '1', '--warmup', '1'], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=900)
    assert r.returncode == 0 and 'control flow ok on 2 rank(s)' in r.stdout, r.stdout[-3000:]
