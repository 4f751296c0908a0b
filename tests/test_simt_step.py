"""The GPU suites' own test bodies, executed on the CPU against the host-SIMT build of the kernel sources.

`tests/simt/emulate.py` points the product's Python stack at tests/simt/build/libgsat_sim.so (the unmodified .cu files
compiled by g++ against the emulator) and redirects `.cuda()` / `device='cuda'` to host tensors for the duration of one
test; the functions called below are the `-m gpu` tests of tests/test_gpu_parity.py and tests/test_gpu_z_next_rows.py,
unchanged, with the oracle as the checker.  What this tier proves and what it does not is stated in DESIGN.md section 2a:
indexing, masks, warp-level reductions, argument order and autograd routing of every kernel, the tensor-core ones through
functional models of mbarrier / TMA / tcgen05.mma / TMEM -- not inter-warp races, not the hardware's accumulation order,
not speed.  It is not a GPU result.

Left out on purpose: the CUDA-graph test, the BASELINE-size property test, and two whole-step cases whose ill-conditioned BatchNorm statistics (constant BA-2Motifs features; see the docstring of
test_gsat_step_parity.check) put the CPU library ops that replace torch's CUDA ops under emulation outside the bound.
"""
import pytest

import tests.test_gpu_parity as P
import tests.test_gpu_tc as T
import tests.test_gpu_z_next_rows as Z
import tests.test_gpu_dense as D


@pytest.fixture
def G(monkeypatch):
    from tests.simt import emulate
    emulate.redirect_torch_to_cpu(monkeypatch.setattr)
    emulate.patch_product(monkeypatch.setattr)
    import dp_gsat_b200 as g
    yield g
    g.clear_index_cache()


def _case(fn, **kw):
    ident = fn.__name__[5:] + ''.join(f'-{v}' for v in kw.values())
    return pytest.param(fn, kw, id=ident)


def case_gin_rows_kernels(G, H, N, grid):
    mp = pytest.MonkeyPatch()
    try:
        T.test_gin_rows_kernels_match_the_channel_owner_kernels(G, H, N, grid, mp)
    finally:
        mp.undo()


CASES = [
    # K0: index / CSR / CSC / reverse map / flags, bit-exact
    *[_case(P.test_index_build_bit_exact, name=n) for n in ('ba2motifs', 'molhiv', 'mutag', 'shuffled', 'directed',
                                                             'duplicates', 'empty')],
    _case(P.test_mutag_reverse_is_xor1_on_gpu),
    # K3 / K5 / segment norm / sampler / lift, forward and backward
    _case(P.test_gin_aggregate_fwd_bwd, H=64, with_att=True),
    _case(P.test_gin_aggregate_fwd_bwd, H=300, with_att=False),
    _case(P.test_gin_aggregate_fwd_bwd, H=16, with_att=True),
    _case(P.test_gin_aggregate_high_degree_and_isolated, deg=500, H=64),
    _case(P.test_pool, mean=False),
    _case(P.test_pool, mean=True),
    _case(P.test_instance_norm, C=64),
    _case(P.test_instance_norm, C=300),
    _case(P.test_sample_avg_info, training=True, info_on='att', tensor_r=False),
    _case(P.test_sample_avg_info, training=True, info_on='edge_att', tensor_r=True),
    _case(P.test_sample_avg_info, training=False, info_on='att', tensor_r=False),
    _case(P.test_sampler_philox_is_regenerable_and_uniform),
    _case(P.test_lift),
    # whole steps
    _case(P.test_gsat_step_parity, cfgname='cfg1_L3'),
    _case(P.test_gsat_step_parity, cfgname='mutag_dual_avg'),
    _case(P.test_gsat_step_parity, cfgname='lift_path'),
    _case(P.test_gsat_step_parity, cfgname='fork_info_on_edge_att'),
    _case(P.test_gsat_step_parity, cfgname='eval_mode'),
    _case(P.test_fork_glue_composes_with_autograd),
    # K4 PNA
    _case(P.test_pna_aggregate_fwd_bwd, H=80, with_ea=True, with_att=True),
    _case(P.test_pna_aggregate_fwd_bwd, H=16, with_ea=False, with_att=True),
    _case(P.test_pna_empty_rows_and_constant_segments),
    _case(P.test_gsat_pna_step_parity, use_edge_attr=False, learn_edge_att=False),
    _case(P.test_gsat_pna_step_parity, use_edge_attr=True, learn_edge_att=False),
    # node-encoder weight gradient, line-graph builder, GINE, metrics
    _case(P.test_small_linear_weight_gradient, N=1000, F_=10, H=64),
    _case(P.test_line_graph_dual_bit_exact, case='kat4', halve=False),
    _case(P.test_line_graph_dual_bit_exact, case='mutag', halve=True),
    _case(P.test_line_graph_dual_bit_exact, case='directed', halve=False),
    _case(P.test_gine_aggregate_fwd_bwd, H=64, with_att=True),
    _case(P.test_gsat_gin_with_edge_features_step_parity, atom_encoder=True),
    _case(P.test_on_device_metrics, k=5),
    # the tensor-core path: tcgen05.mma / TMEM / TMA / mbarrier kernels on their host models (tests/simt/tc_sim.h)
    _case(T.test_tc_linear_matches_bf16_reference, rows=1000, K=128, OUT=128),
    _case(T.test_tc_linear_matches_bf16_reference, rows=3333, K=256, OUT=512),
    _case(T.test_tc_linear_matches_bf16_reference, rows=777, K=80, OUT=80),
    _case(T.test_fused_extractor_fwd_bwd, case='ba_edge_H64'),
    _case(T.test_fused_extractor_fwd_bwd, case='ba_edge_H128'),
    _case(T.test_fused_extractor_fwd_bwd, case='mol_node_H64'),
    _case(T.test_fused_extractor_fwd_bwd, case='mol_edge_H80_p03'),
    _case(T.test_fused_extractor_fwd_bwd, case='eval_mode'),
    _case(T.test_gin_mlp_fused_matches_torch, H=64),
    _case(T.test_gin_mlp_fused_matches_torch, H=128),
    # the row-owner node-MLP kernels (csrc/gin_rows.cu) against the channel-owner ones, many tiles per CTA included
    _case(case_gin_rows_kernels, H=64, N=777, grid=148),
    _case(case_gin_rows_kernels, H=128, N=777, grid=148),
    _case(case_gin_rows_kernels, H=128, N=4096 + 40, grid=3),
    _case(T.test_weight_grad_pairs_of_a_blocks, M=256, N=192, rows=5000, layouts=(0, 0)),
    _case(T.test_weight_grad_pairs_of_a_blocks, M=512, N=256, rows=4096, layouts=(1, 0)),
    _case(T.test_gather_concat_bwd_bf16_equals_the_fp32_reduction_of_the_same_values, H=64),
    _case(T.test_gsat_step_bf16_mode_tracks_oracle),
    _case(T.test_word_dropout_rate_and_scale, p=0.3),
    _case(T.test_bf16_mode_layer_by_layer_path, case='mutag_dual_big_graphs'),
    _case(T.test_gsat_pna_step_bf16_mode, use_edge_attr=True),
    _case(D.test_bf16x2_linear, rows=513, K=640, OUT=80),
    _case(D.test_strict_linear_matches_fp32, rows=513, K=10, OUT=64),
    _case(D.test_strict_linear_matches_fp32, rows=129, K=3, OUT=3),
    _case(D.test_batch_norm_matches_torch, rows=1000, C=80, relu=False),
    _case(D.test_batch_norm_matches_torch, rows=2, C=16, relu=True),
    # SURVEY 8f rows built after the GPU budget was spent: emulator runs are all they have had so far
    _case(Z.test_le_aggregate_fwd_bwd, H=32, with_w=True, with_att=True),
    _case(Z.test_le_aggregate_fwd_bwd, H=300, with_w=True, with_att=True),
    _case(Z.test_leconv_layer_and_state_dict),
    _case(Z.test_gsat_spmotifnet_step_parity, learn_edge_att=True),
    _case(Z.test_gsat_spmotifnet_step_parity, learn_edge_att=False),
    _case(Z.test_embedding_sum_fwd_bwd, M=5000, H=80, dims=[119, 4, 12, 12, 10, 6, 6, 2, 2]),
    _case(Z.test_embedding_sum_out_of_range_is_clamped_and_flagged),
    _case(Z.test_fused_encoders_inside_the_step, model_name='PNA'),
    _case(Z.test_fused_encoders_inside_the_step, model_name='GIN'),
    _case(Z.test_device_collate_bit_exact, ids=[3, 1, 4, 1, 5, 9, 2, 6]),
    _case(Z.test_device_collate_edge_cases_and_loader),
    _case(Z.test_step_on_a_device_collated_batch_equals_the_host_batch),
    # the fork's two-model step
    _case(Z.test_dual_forward_pass_parity, primal_learn_edge_att=True, epoch=3),
    _case(Z.test_dual_forward_pass_parity, primal_learn_edge_att=False, epoch=3),
    _case(Z.test_dual_forward_pass_parity, primal_learn_edge_att=True, epoch=60),
    _case(Z.test_dual_forward_pass_parity, primal_learn_edge_att=False, epoch=60),
    _case(Z.test_dual_train_and_eval_one_batch),
    # product directly against the outputs of the reference's own class bodies (tests/golden/ref_fork.pt)
    _case(Z.test_product_layers_reproduce_the_reference_layers),
    _case(Z.test_product_backbones_reproduce_the_reference_classes, tag='pna'),
    _case(Z.test_product_backbones_reproduce_the_reference_classes, tag='gin'),
    _case(Z.test_product_backbones_reproduce_the_reference_classes, tag='spmotif'),
    _case(Z.test_product_dual_forward_pass_reproduces_the_reference_body, epoch=3),
    _case(Z.test_product_dual_forward_pass_reproduces_the_reference_body, epoch=57),
    _case(Z.test_product_metrics_reproduce_the_reference_bodies),
    _case(Z.test_product_line_graph_reproduces_the_reference_loops, tag='ba2motifs'),
    _case(Z.test_product_line_graph_reproduces_the_reference_loops, tag='mol'),
    _case(Z.test_dense_dual_reproduces_the_reference_loops),
]


@pytest.mark.parametrize('fn,kw', CASES)
def test_gpu_test_body_on_the_emulator(G, fn, kw):
    fn(G, **kw)


def test_sync_batchnorm_single_rank_body_on_the_emulator(G):
    gen = Z.single_rank_group.__wrapped__()            # the fixture's generator: a one-rank gloo group here
    group = next(gen)
    try:
        Z.test_sync_batchnorm_on_a_single_rank_group_equals_the_default(G, group, 'fp32')
        Z.test_sync_batchnorm_on_a_single_rank_group_equals_the_default(G, group, 'bf16')
    finally:
        next(gen, None)
