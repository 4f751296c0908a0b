#!/usr/bin/env python
"""Headline benchmark: GSAT-GIN training-step throughput in directed edges / second on the large synthetic
BA-motif batch (BASELINE.json configs[3]: ~10 M edges, hidden 128), graph-sharded data-parallel at N GPUs.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

A step = forward_pass(training=True) (both GNN passes, extractor, sampler, reverse-average, losses) + backward +
gradient all-reduce (N>1) + Adam step.  `value` is measured with the shard resident in HBM; `e2e` is measured through
the public API from pinned HOST buffers (H2D of the batch + index build + step + D2H of the loss inside the timed
region).  One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'gsat_gin_train_step_directed_edges_per_sec'
UNIT = 'edges/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--graphs', type=int, default=196000, help='BA-2Motifs-shaped graphs in the global batch (cfg4)')
    ap.add_argument('--hidden', type=int, default=128)
    ap.add_argument('--layers', type=int, default=2)
    ap.add_argument('--e2e-steps', type=int, default=5)
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32'],
                    help='bf16: tcgen05 MLPs (bf16 operands, fp32 accumulate); fp32: strict mode, the same tcgen05 GEMM '
                         'kernels on split-bf16 x3 operands (fp32-accurate products, the rtol-1e-5 parity mode)')
    ap.add_argument('--cuda-graph', default='auto', choices=['auto', 'on', 'off'],
                    help='replay the whole step (fwd + bwd + all-reduce + Adam) as ONE CUDA graph in the timed region; '
                         'auto = on at every N (the same launch mode for the whole scaling run), eager if capture fails')
    ap.add_argument('--sync-bn', action='store_true',
                    help='exact multi-GPU mode: BatchNorm batch statistics span all ranks (the N-GPU step then computes '
                         'the single-GPU step); default = shard-local statistics (DDP semantics)')
    ap.add_argument('--strict-steps', type=int, default=2,
                    help='N = 1, bf16 run only: also time this many steps of the strict fp32 mode on the same workload '
                         '(reported under "strict_mode"; 0 = skip)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--cpu-sample-graphs', type=int, default=0, help='0 = size automatically (~10-30 s of CPU work)')
    return ap.parse_args()


def model_cfg(a):
    return ({'model_name': 'GIN', 'hidden_size': a.hidden, 'n_layers': a.layers, 'dropout_p': 0.3,
             'use_edge_attr': False}, {'learn_edge_att': True, 'extractor_dropout_p': 0.5})


def workload_name(a):
    return f'cfg4 large synthetic BA-2Motifs batch: {a.graphs} graphs, GSAT-GIN hidden {a.hidden}, {a.layers} layers'


# ---------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.idx), '-lms', '50'], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm),
                'window': 'identical untimed steps (>= 0.7 s, sampler already running) + the timed region'}


def preload(run_one, sync, world, dev, seconds=0.7):
    """Keep the GPU under the SAME load for about `seconds` before the timed region, so that the nvidia-smi sampler
    (which needs ~100 ms to produce its first row) sees the clocks the timed steps run at even when K steps last
    < 0.1 s.  The number of extra steps is AGREED across ranks (every step holds collectives: a per-rank, wall-clock
    bounded loop would let ranks run different counts and dead-lock): two steps are timed, the maximum over ranks is
    all-reduced, and every rank derives the same count from it."""
    import math
    sync()
    t0 = time.perf_counter()
    run_one()
    run_one()
    sync()
    t = torch.tensor([(time.perf_counter() - t0) / 2.0], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    n = int(max(1, min(500, math.ceil(seconds / max(float(t.item()), 1e-5)))))
    for _ in range(n):
        run_one()
    sync()
    return n


# ---------------------------------------------------------------------------------------------------------------
# CPU path (the oracle port of the reference's algorithm) -- cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------------------------------
def _data_module():
    """The synthetic generators (dp_gsat_b200/data.py: numpy + torch only) loaded BY FILE, not through the package:
    importing the package maps libgsat_b200.so, and the reference arm must not touch the product's library."""
    import importlib.util
    name = '_gsatb_data_standalone'
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, 'dp_gsat_b200', 'data.py'))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def cpu_oracle_rate(a, n_graphs, steps, warmup, threads):
    """edges/s of the oracle's training step on `n_graphs` BA-2Motifs graphs, host cores."""
    from oracle import gsat_oracle as O
    ba2motifs_batch = _data_module().ba2motifs_batch
    torch.set_num_threads(threads)
    cfg, shared = model_cfg(a)
    torch.manual_seed(0)
    b = ba2motifs_batch(n_graphs, seed=0)
    clf = O.get_model(b.x.shape[1], 0, 2, False, cfg)
    ext = O.ExtractorMLP(a.hidden, shared)
    g = O.GSAT(clf, ext, O.Criterion(2, False), learn_edge_att=True, final_r=0.5)
    g.train()
    opt = torch.optim.Adam(list(ext.parameters()) + list(clf.parameters()), lr=1e-3)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, loss, _, _ = g.forward_pass(b, 0, True)
        opt.zero_grad()
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return b.num_edges * len(times) / tot, b.num_edges, tot / len(times)


def auto_cpu_sample(a, threads, budget_s):
    """Pick a sample size whose (warmup + steps) oracle run costs about budget_s, from a small probe."""
    rate, _, _ = cpu_oracle_rate(a, 200, 1, 1, threads)
    return rate


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    probe_rate = auto_cpu_sample(a, threads, 0)
    total_steps = a.steps + a.warmup
    budget = 150.0
    n_graphs = a.cpu_sample_graphs or int(max(200, min(20000, probe_rate * budget / total_steps / 51.0)))
    rate, n_edges, sec = cpu_oracle_rate(a, n_graphs, a.steps, a.warmup, threads)
    cfg = {'workload': workload_name(a), 'hidden': a.hidden, 'layers': a.layers, 'global_graphs': a.graphs,
           'sample': f'{n_graphs} graphs / {n_edges} edges per step (bounded sample of the same generator)'}
    line = {'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT, 'n_gpus': a.gpus, 'steps': a.steps,
            'warmup': a.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': cfg,
            'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                             'sample': cfg['sample'] + '; oracle/gsat_oracle.py (pure-PyTorch restatement; the '
                             'reference cannot be imported: torch_geometric/torch_scatter/torch_sparse absent)'},
            'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# roofline: algorithmic bytes / flops per launch of every kernel of ours (DESIGN.md "Kernels"), live CUDA-event times
# ---------------------------------------------------------------------------------------------------------------
def kernel_models(N, E, G, H):
    """{C-ABI entry[:variant]: (algorithmic bytes per launch, flops per launch)} for the GSAT-GIN step at hidden H
    (extractor widths 2H -> 4H -> H -> 1).  Bytes = every distinct input element read once + every output element
    written once; the bf16 intermediates a kernel reads / writes are part of ITS contract."""
    C1 = 4 * H
    return {
        # the fused extractor, SURVEY section 8d's contract figures: bytes = emb + indices + logits (fwd), 2x emb + 12E
        # (bwd); flops = GEMM1 (2H x 4H) + GEMM2 (4H x H) forward, recomputed GEMM1 + dh1 + d f12 backward
        'gsatb_ext_fused_fwd': (4.0 * N * H + 12.0 * E, E * (24.0 * H * H + 2.0 * H)),
        'gsatb_ext_fused_bwd': (8.0 * N * H + 12.0 * E, E * 40.0 * H * H),
        'gsatb_gin_aggregate_fwd:att': (8.0 * N * H + 8.0 * E + 4.0 * N, 2.0 * E * H),
        'gsatb_gin_aggregate_fwd:noatt': (8.0 * N * H + 4.0 * E + 4.0 * N, 1.0 * E * H),
        'gsatb_gin_aggregate_bwd:att': (12.0 * N * H + 16.0 * E, 4.0 * E * H),
        'gsatb_gin_aggregate_bwd:noatt': (8.0 * N * H + 8.0 * E, 1.0 * E * H),
        'gsatb_gin_aggregate_fwd_bf16:att': (6.0 * N * H + 8.0 * E + 4.0 * N, 2.0 * E * H),
        'gsatb_gin_aggregate_fwd_bf16:noatt': (6.0 * N * H + 4.0 * E + 4.0 * N, 1.0 * E * H),
        # (gsatb_tc_linear_bf16_fwd is not modelled: at H = 64 / 128 the node MLP runs on the row-owner kernels and the
        # remaining calls are the small encoder / classifier Linears, whose shapes differ from call to call)
        # row-owner node-MLP kernels (csrc/gin_rows.cu): lin1 = bf16 agg in, bf16 z1 out; lin2 = bf16 z1 in, bf16 a1 +
        # fp32 h + sign bits out; bwd1 = bf16 g, z1 in, bf16 dz1 + fp32 dx out; bwd2 = fp32 dh, sign bits, bf16 z1 in,
        # bf16 d2, g out
        'gsatb_gin_rows_lin1': (4.0 * N * H, 2.0 * N * H * H),
        'gsatb_gin_rows_lin2': (N * (2.0 * H + 2 * H + 4 * H + H / 8.0), 2.0 * N * H * H),
        'gsatb_gin_rows_bwd1': (N * (2.0 * H + 2 * H + 2 * H + 4 * H), 2.0 * N * H * H),
        'gsatb_gin_rows_bwd2': (N * (4.0 * H + H / 8.0 + 2 * H + 2 * H + 2 * H), 2.0 * N * H * H),
        'gsatb_bn_relu_bf16': (4.0 * N * H, 0.0),
        'gsatb_tc_linear_fwd': (8.0 * N * H, 2.0 * N * H * H),
        'gsatb_tc_gin_bwd2': (N * (4.0 * H + H / 8.0 + 2 * H + 2 * H + 2 * H), 2.0 * N * H * H),   # dh, sign bits, z1 in; d2, g out
        'gsatb_tc_gin_bwd1': (N * (2.0 * H + 2 * H + 2 * H + 4 * H), 2.0 * N * H * H),
        'gsatb_tc_ext_fwd1': (2.0 * E * 2 * H + 2.0 * E * C1, 2.0 * E * 2 * H * C1),
        'gsatb_tc_ext_fwd2': (2.0 * E * C1 + 2.0 * E * H + 4.0 * E, 2.0 * E * C1 * H),
        'gsatb_tc_ext_bwd_head': (4.0 * E + 2.0 * E * H + 2.0 * E * H, 8.0 * E * H),
        'gsatb_tc_ext_bwd1': (2.0 * E * H + 2.0 * E * C1 + 2.0 * E * C1, 2.0 * E * H * C1),
        'gsatb_tc_linear_bf16in_fwd': (2.0 * E * C1 + 4.0 * E * 2 * H, 2.0 * E * C1 * 2 * H),
        'gsatb_tc_ext_make_f12': (4.0 * N * H + 8.0 * E + 2.0 * E * 2 * H, 0.0),
        'gsatb_tc_ext_make_h1': (4.0 * E * C1, 0.0),
        'gsatb_linear_small_dw': (4.0 * N * H + 40.0 * N, 2.0 * N * H * 11),
        'gsatb_gather_concat_bwd': (4.0 * E * 2 * H + 8.0 * E + 4.0 * N * H, 2.0 * E * H),
        'gsatb_gather_concat_bwd_bf16': (2.0 * E * 2 * H + 8.0 * E + 4.0 * N * H, 2.0 * E * H),
        'gsatb_sample_avg_info_fwd': (20.0 * E, 0.0),
        'gsatb_sample_avg_info_bwd': (20.0 * E, 0.0),
        'gsatb_pool_fwd': (4.0 * N * H + 4.0 * G * H, 1.0 * N * H),
        'gsatb_pool_bwd': (4.0 * N * H + 4.0 * G * H + 4.0 * N, 0.0),
    }


def build_roofline(timer, N, E, G, H, steps, ms_step, unmodelled=()):
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    hbm = float(peaks.get('hbm_gbs', 6650.0))
    tc = float(peaks.get('bf16_tflops_sustained', peaks.get('bf16_tflops', 1400.0)))    # timed inside a long step
    models = kernel_models(N, E, G, H)
    rows = []
    for name, evs in timer.items():
        by_tag = {}
        for e0, e1, tag in evs:
            by_tag.setdefault(tag or '', []).append(e0.elapsed_time(e1))
        for tag, ms in by_tag.items():
            key = name + (':' + tag if tag else '')
            tot = sum(ms)
            row = {'kernel': key, 'launches_per_step': len(ms) / steps, 'avg_launch_ms': tot / len(ms),
                   'share_of_step': tot / steps / ms_step}
            if name in unmodelled:
                models.pop(key, None)
            elif name == 'gsatb_tc_dw' and tag:          # tag = 'rows x M x N' of the product dW[M, N] = A[rows, M]^T B[rows, N]
                r_, m_, n_ = (float(v) for v in tag.split('x'))
                models[key] = (2.0 * r_ * (m_ + n_) + 4.0 * m_ * n_, 2.0 * r_ * m_ * n_)
            if key in models:
                nbytes, flops = models[key]
                t_hbm, t_tc = nbytes / (hbm * 1e9), flops / (tc * 1e12)
                sec = row['avg_launch_ms'] * 1e-3
                if t_tc > t_hbm:
                    row.update(bound='tensor', achieved=flops / sec / 1e12, peak=tc, unit='TFLOP/s')
                else:
                    row.update(bound='hbm', achieved=nbytes / sec / 1e9, peak=hbm, unit='GB/s')
                row['frac'] = row['achieved'] / row['peak']
                if row['unit'] == 'GB/s':                 # both the measured and the nominal 8 TB/s fraction
                    row['frac_of_nominal_8TBs'] = row['achieved'] / 8000.0
                row['algorithmic_bytes_per_launch'], row['flops_per_launch'] = nbytes, flops
            rows.append(row)
    rows.sort(key=lambda r: -r['share_of_step'])
    modelled = [r for r in rows if 'frac' in r]
    top = dict(modelled[0]) if modelled else {'kernel': None, 'bound': 'hbm', 'achieved': 0.0, 'peak': hbm,
                                              'unit': 'GB/s', 'frac': 0.0}
    # dram__bytes_read + write per launch from the committed ncu --set full captures (same shapes), where one exists
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'r2_ncu_traffic.json')))
    except Exception:
        pass
    for r in rows:
        r['traffic'] = traffic.get(r['kernel']) if (N, E, H) == (4900000, 9996000, 128) else None
    top['traffic'] = traffic.get(top.get('kernel')) if (N, E, H) == (4900000, 9996000, 128) else None
    top['peak_source'] = ('measured (MEASURED_PEAKS.json: hbm_gbs, bf16_tflops_sustained)' if peaks
                          else 'fallback 6650 GB/s / 1400 TFLOP/s')
    top['frac_of_nominal_8TBs'] = top['achieved'] / 8000.0 if top.get('unit') == 'GB/s' else None
    top['dominant_by'] = 'largest share of the timed step among our kernels (CUDA events around every C-ABI call)'
    top['kernels'] = [{k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items()} for r in rows[:16]]
    top['ours_share_of_step'] = sum(r['share_of_step'] for r in rows)
    return top


# ---------------------------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    import dp_gsat_b200 as G
    from dp_gsat_b200._lib import lib
    from dp_gsat_b200.data import ba2motifs_batch, shard_batch
    from dp_gsat_b200.parallel import TrainStep, broadcast_parameters

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg, shared = model_cfg(a)
    full = ba2motifs_batch(a.graphs, seed=0)
    E_global = full.num_edges
    shard_host = shard_batch(full, rank, world).pin_memory()
    del full
    torch.manual_seed(0)
    clf = G.get_model(shard_host.x.shape[1], 0, 2, False, cfg, dev)
    ext = G.ExtractorMLP(a.hidden, shared).to(dev)
    broadcast_parameters(clf)
    broadcast_parameters(ext)
    clf.precision = ext.precision = a.precision
    gsat = G.GSAT(clf, ext, G.Criterion(2, False), learn_edge_att=True, final_r=0.5, lazy_metrics=True)
    gsat.train()
    step = TrainStep(gsat, lr=1e-3, sync_bn=a.sync_bn)
    data = shard_host.to(dev)
    N_loc, E_loc, H = data.num_nodes, data.num_edges, a.hidden

    for _ in range(a.warmup):
        step(data, 0)
    barrier()

    L = lib()
    use_graph = a.cuda_graph in ('on', 'auto')
    graphed = step.enable_cuda_graph(data, 0, warmup=2) if use_graph else False
    if world > 1:                                            # every rank must take the same path
        flag = torch.tensor([1 if graphed else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if use_graph and int(flag.item()) == 0:
            step.disable_cuda_graph()
            graphed = False
    if graphed:
        # headline timing: graph replays.  Per-kernel roofline times cannot be taken inside a replay, so they come from
        # the same number of instrumented EAGER steps right after (identical kernels, same data).
        for _ in range(2):
            step(data, 0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks = Clocks(local)
        if rank == 0:
            clocks.start()
        preload(lambda: step(data, 0), barrier, world, dev)
        barrier()
        ev0.record()
        for _ in range(a.steps):
            _, loss, _, _ = step(data, 0)
        ev1.record()
        barrier()
        ms_graph = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_graph, op=dist.ReduceOp.MAX)
        clk_graph = clocks.stop() if rank == 0 else None
        launches_graph = step.launches_per_step * a.steps
        step.disable_cuda_graph()
        torch.cuda.empty_cache()
    L.timer, L.timer_all = {}, True          # CUDA events around every C-ABI call of the timed region (rank-local)
    def tag(name, args):           # variants of one entry point that have different algorithmic bytes
        if name in ('gsatb_gin_aggregate_fwd', 'gsatb_gin_aggregate_fwd_bf16'):
            return 'att' if args[1] is not None else 'noatt'
        if name == 'gsatb_gin_aggregate_bwd':
            return 'att' if args[2] is not None else 'noatt'
        if name == 'gsatb_tc_linear_bf16_fwd':
            return 'bf16' if args[5] else 'fp32'
        if name == 'gsatb_tc_dw':
            return f'{int(args[6])}x{int(args[7])}x{int(args[8])}'
        return ''
    L.timer_tag = tag
    clocks = Clocks(local)
    if rank == 0:
        clocks.start()
    L.timer_all = False                       # (the preload steps are not part of the per-kernel statistics)
    preload(lambda: step(data, 0), barrier, world, dev)
    L.timer, L.timer_all = {}, True
    launches0 = L.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(a.steps):
        _, loss, _, _ = step(data, 0)
    ev1.record()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if rank == 0 else None
    launches = L.launches - launches0
    timer, L.timer, L.timer_all, L.timer_tag = L.timer, None, False, None
    ms_step = float(ms.item()) / a.steps
    # (strict mode runs the GEMM entry points on 6x wider split operands: their bf16-mode byte models do not apply)
    roofline = build_roofline(timer, N_loc, E_loc, data.num_graphs, H, a.steps, ms_step,
                              unmodelled=('gsatb_tc_linear_bf16_fwd', 'gsatb_tc_dw') if a.precision == 'fp32' else ())
    if graphed:
        roofline['timed_in'] = (f'{a.steps} instrumented eager steps ({ms_step:.3f} ms/step) run right after the timed '
                                'CUDA-graph replays: per-kernel events cannot be taken inside a graph replay')
        ms_step, clk, launches = float(ms_graph.item()) / a.steps, clk_graph, launches_graph
    value = E_global / (ms_step * 1e-3)

    # end-to-end through the public API from pinned host buffers.  Every step's inputs are copied host -> device inside
    # the timed region; as a data loader would, the copy of step i+1 is issued on a copy stream while step i computes
    # (double buffering), the K0 index build of the fresh tensors and the D2H read of the loss are in the step.
    G.clear_index_cache()
    e2e_steps = max(1, a.e2e_steps)
    copy_stream = torch.cuda.Stream(device=dev)
    main_stream = torch.cuda.current_stream()

    from dp_gsat_b200 import tc as _tc
    ext_plans = [('edge', _tc.ext_tile_slots(H, True))] if a.precision == 'bf16' else []

    def fetch():
        with torch.cuda.stream(copy_stream):
            d = shard_host.to(dev, non_blocking=True)                  # H2D of one step's inputs
        for t in (d.x, d.edge_index, d.batch, d.y, d.edge_attr, d.edge_label):
            if t is not None:
                t.record_stream(main_stream)                           # allocated on the copy stream, used on the main one
        # the loader's prefetch stage also builds the batch's index (K0 + tile plan) behind the copy, on the copy stream:
        # ~40 short dependent launches that hide under the step the main stream is running
        G.prefetch_graph_index(d.edge_index, d.batch, getattr(d, 'num_graphs', None), on_stream=copy_stream,
                               for_stream=main_stream, ext_plans=ext_plans)
        ev = torch.cuda.Event()
        ev.record(copy_stream)
        return d, ev

    # With the step captured as a CUDA graph on the resident shard (`data`), a fresh batch of the same shape is copied --
    # with the index K0 built for it on the copy stream -- into the static buffers the graph reads (TrainStep.load_batch:
    # device-to-device, ~1 GB) and the step is ONE replay; a batch that does not fit takes an eager step.
    e2e_graph = bool(graphed) and step.enable_cuda_graph(data, 0, warmup=1)
    if world > 1:
        flag = torch.tensor([1 if e2e_graph else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if e2e_graph and int(flag.item()) == 0:
            step.disable_cuda_graph()
        e2e_graph = bool(int(flag.item()))
    replayed = [0, 0]

    def e2e_loop(n):
        nxt = fetch()
        last = None
        for i in range(n):
            d, ev = nxt
            main_stream.wait_event(ev)
            gi_d = G.get_graph_index(d.edge_index, d.batch, getattr(d, 'num_graphs', None))      # the prefetched entry
            if e2e_graph and step.load_batch(d, gi_d):
                _, loss, _, _ = step(data, 0)               # one graph replay on the refreshed resident batch ...
                replayed[0] += 1
            else:
                _, loss, _, _ = step(d, 0)                  # (eager: the index is already cached)
                replayed[1] += 1
            if i + 1 < n:
                nxt = fetch()                               # ... which runs while batch i + 1 is copied and indexed
            last = float(loss.item())                       # D2H of the step's result
            G.evict_graph_index(d.edge_index, d.batch)      # this batch is done: its index blocks go back to the allocator
        return last

    # untimed warm-up of THIS loop (W >= 3 applies to it as well): the first passes grow the copy stream's allocator pool
    # with cudaMalloc calls, which synchronise the device and would be charged to the timed steps
    e2e_loop(3)
    barrier()
    replayed[0] = replayed[1] = 0
    t0 = time.perf_counter()
    loss_host = e2e_loop(e2e_steps)
    barrier()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = {'value': E_global / float(t_e2e.item()), 'unit': UNIT, 'h2d_bytes_per_step': shard_host.nbytes() * world,
           'd2h_bytes_per_step': 4 * world, 'ms_per_step': float(t_e2e.item()) * 1e3, 'steps': e2e_steps,
           'last_loss': loss_host, 'warmup': 3,
           'how': 'pinned host batch -> H2D + K0 index build on the copy stream, prefetched one step ahead -> D2D refresh of the '
                  "captured step's static batch + index -> one CUDA-graph replay -> loss.item()" if replayed[0] else
                  'pinned host batch -> H2D + K0 index build on the copy stream, prefetched one step ahead -> eager step -> loss.item()',
           'graph_replays': replayed[0], 'eager_steps': replayed[1]}
    if e2e_graph:
        step.disable_cuda_graph()
        torch.cuda.empty_cache()

    # strict-mode figure beside the bf16 one (N = 1): the same workload, model and step with precision='fp32' -- the
    # mode the rtol-1e-5 parity tests hold to the fp32 oracle (same tcgen05 GEMM kernels, split-bf16 x3 operands)
    strict = None
    if world == 1 and a.precision == 'bf16' and a.strict_steps > 0:
        try:
            step.disable_cuda_graph()
            G.clear_index_cache()
            torch.cuda.empty_cache()
            clf.precision = ext.precision = 'fp32'
            step(data, 0)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for _ in range(a.strict_steps):
                step(data, 0)
            ev1.record()
            torch.cuda.synchronize()
            ms_strict = ev0.elapsed_time(ev1) / a.strict_steps
            strict = {'value': E_global / (ms_strict * 1e-3), 'unit': UNIT, 'ms_per_step': ms_strict, 'steps': a.strict_steps,
                      'warmup': 1, 'dtype': 'f32 (split-bf16 x3 tensor-core products)', 'step_launch': 'eager launches'}
        except Exception as exc:                          # an extra figure must never cost the bench line
            strict = {'error': repr(exc)}
        finally:
            clf.precision = ext.precision = a.precision

    cpu_baseline = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        probe = auto_cpu_sample(a, threads, 0)
        n_graphs = a.cpu_sample_graphs or int(max(200, min(8000, probe * 20.0 / 3 / 51.0)))
        rate, n_edges, sec = cpu_oracle_rate(a, n_graphs, 2, 1, threads)
        cpu_baseline = {'value': rate, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                        'sample': f'{n_graphs} graphs / {n_edges} edges per step of the same generator, 1 warm-up + 2 '
                                  f'timed steps, {sec:.2f} s/step; oracle/gsat_oracle.py on torch CPU'}
        try:     # the reference's own thread setting (torch.set_num_threads(5), src/run_gsat.py:1049), half the sample
            r5, e5, s5 = cpu_oracle_rate(a, max(200, n_graphs // 2), 1, 1, min(5, threads))
            cpu_baseline['reference_thread_setting'] = {'value': r5, 'unit': UNIT, 'cores': min(5, threads),
                                                        'sample': f'{e5} edges per step, 1 warm-up + 1 timed step, '
                                                                  f'{s5:.2f} s/step'}
        except Exception as exc:                          # an optional extra figure must never cost the bench line
            cpu_baseline['reference_thread_setting'] = {'error': repr(exc)}
        finally:
            torch.set_num_threads(threads)
    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': a.steps, 'warmup': a.warmup,
                'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'bf16' if a.precision == 'bf16' else 'f32', 'data': 'synthetic',
                'config': {'workload': workload_name(a), 'precision': a.precision + (' tensor-core MLP operands, fp32 accumulate / gather / scatter / sampler' if a.precision == 'bf16' else ''), 'global_edges': E_global, 'global_nodes': a.graphs * 25,
                           'hidden': a.hidden, 'layers': a.layers, 'parallelism': f'graph-sharded dp{world}',
                           'step_launch': 'one CUDA graph replay per step' if graphed else 'eager launches',
                           'batchnorm': 'statistics over all ranks (exact: equals the single-GPU step)' if a.sync_bn
                           else ('shard-local statistics (DDP semantics)' if world > 1 else 'single device'),
                           'l2_policy': 'inputs larger than L2 (per-rank activations >> 126 MB)'},
                'roofline': roofline, 'cpu_baseline': cpu_baseline, 'e2e': e2e, 'strict_mode': strict, 'gpu_launches': launches,
                'clocks': clk}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)
