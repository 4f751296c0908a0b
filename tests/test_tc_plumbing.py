"""Autograd plumbing of the tensor-core GIN layer (dp_gsat_b200/tc.py) with the tcgen05 kernels STUBBED OUT: every C-ABI
call is a no-op that leaves its outputs zero-filled, so nothing numeric is checked here -- only what can break without
a GPU in sight: the number and order of gradients each autograd.Function returns, tensor shapes and dtypes, the
optional data-parallel group argument (sync BatchNorm) and the running-statistics updates.  Numerics of these kernels
are the job of tests/test_gpu_tc.py on a B200."""
import types

import pytest
import torch


class _StubLib:
    launches = 0

    class cdll:
        @staticmethod
        def gsatb_tc_stat_partials_elems(h):
            return 4 * h

        @staticmethod
        def gsatb_tc_dw_workspace(rows, m, n):
            return 16

        @staticmethod
        def gsatb_gin_rows_supported(k, h1, h):          # the row-owner kernels take hidden 64 / 128 (csrc/gin_rows.cu)
            return int(k == h1 == h and h in (64, 128))

        @staticmethod
        def gsatb_gin_rows_stat_partials_elems(h):
            return 8 * h

    def call(self, name, *args):
        self.launches += 1
        if name == 'gsatb_bn_fold_fwd' and args[10] and args[9] is not None and args[9].value:
            import ctypes                                  # the one side effect the plumbing test looks at
            ctypes.c_int64.from_address(args[9].value).value += 1


@pytest.fixture
def tc(monkeypatch):
    import dp_gsat_b200.tc as tc
    stub = _StubLib()
    monkeypatch.setattr(tc, 'lib', lambda: stub)
    monkeypatch.setattr(tc, 'stream', lambda: None)
    monkeypatch.setattr(torch, 'empty', lambda *a, **k: torch.zeros(*a, **k))        # "kernel outputs" are zeros
    return tc


def _layer(H=16):
    import dp_gsat_b200 as G
    torch.manual_seed(0)
    conv = G.GINConv(G.GIN.MLP(H, H))
    conv.train()
    return conv


def _index(N, E):
    return types.SimpleNamespace(N=N, E=E, rowptr_dst=torch.zeros(N + 1, dtype=torch.int32),
                                 eid_by_dst=torch.zeros(E, dtype=torch.int32), src_by_dst=torch.zeros(E, dtype=torch.int32),
                                 rowptr_src=torch.zeros(N + 1, dtype=torch.int32),
                                 eid_by_src=torch.zeros(E, dtype=torch.int32), dst_by_src=torch.zeros(E, dtype=torch.int32))


@pytest.fixture
def one_rank_group():
    import os
    import socket
    import torch.distributed as dist
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=0, world_size=1)
    yield dist.group.WORLD
    dist.destroy_process_group()


@pytest.mark.parametrize('H', [16, 64])          # channel-owner skeleton kernels / row-owner kernels (hidden 64, 128)
@pytest.mark.parametrize('sync', [False, True])
def test_gin_layer_function_returns_one_gradient_per_input(tc, one_rank_group, sync, H):
    N, E = 12, 30
    conv = _layer(H)
    bn = conv.nn[1]
    if sync:
        bn.sync_group = one_rank_group
    x = torch.randn(N, H, requires_grad=True)
    att = torch.rand(E, 1, requires_grad=True)
    out = tc.gin_layer(x, att, _index(N, E), conv, training=True, pdrop=0.3, drop_seed=1)
    assert out.shape == (N, H) and out.dtype == torch.float32
    assert int(bn.num_batches_tracked) == 1
    out.sum().backward()                                   # raises if a Function returns the wrong number of gradients
    assert x.grad.shape == x.shape and att.grad.shape == att.shape
    for name, p in conv.nn.named_parameters():
        assert p.grad is not None and p.grad.shape == p.shape and p.grad.dtype == torch.float32, name
    out2 = tc.gin_mlp_relu(torch.randn(N, H, requires_grad=True), conv.nn, training=True, pdrop=0.0)
    out2.sum().backward()
    # eval mode: running statistics, no group traffic
    conv.eval()
    tc.gin_layer(x.detach(), None, _index(N, E), conv, training=False)
    assert int(bn.num_batches_tracked) == 2
