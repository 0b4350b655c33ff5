"""Round-2 parity tests of the CONFIGURATIONS THAT ARE MEASURED (VERDICT r01 "next round" item 1), all through the
C ABI on a B200 (-m gpu):

  * the multi-GPU mode the scaling run times -- calibration rows split over ranks, Gram matrices summed over the
    ranks, recurrence per neuron slice -- emulated serially on one GPU for G in {2, 4, 8} and compared with the ORACLE;
  * VGG-16 with the L1 (soft-threshold) alphabet, teacher-forced, including fc6 (d = 25088);
  * ResNet-50 at 3 bits (K = 4), teacher-forced;
  * ResNet-18 free-running: final logits of the quantized network against the oracle's quantized network;
  * K = 128 ("8-bit") and K = 2^15, which the int8 `levels` output used to forbid;
  * tensors on a device that is not the current one.
"""
import numpy as np
import pytest
import torch
import torchvision

import golden_cases as gc
from oracle import gpfq_oracle as orc

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _serial_row_shards(sa, W, X, Xq, G, step, K, reg, lam):
    """What G ranks of the sharded calibration forward compute for one layer, run one after the other on one GPU:
    rank r holds calibration rows [r*m/G, (r+1)*m/G), forms the Gram matrices of ITS rows (gpfq_gram_f32), the
    matrices are summed in rank order (the all-reduce), and rank r runs the recurrence for ITS neuron slice
    (gpfq_gram_path_f32).  Returns (Q, err, rel)."""
    from quantized_neural_nets_b200.sharding import neuron_slice
    m, d = X.shape
    assert m % G == 0
    ml = m // G
    local = []
    for r in range(G):
        Xfm, ld = sa.feature_major(X[r * ml:(r + 1) * ml].contiguous())
        Xqfm, _ = sa.feature_major(Xq[r * ml:(r + 1) * ml].contiguous())
        local.append(sa.local_gram_matrices(Xfm, Xqfm, ld, d, ml))
    total = local[0].clone()
    for g in local[1:]:
        total += g
    N = W.shape[0]
    Q = torch.zeros_like(W)
    e2 = torch.zeros(N, dtype=torch.float64, device=W.device)
    r2 = torch.zeros(N, dtype=torch.float64, device=W.device)
    for r in range(G):
        n0, n1 = neuron_slice(N, 1, world=G, rank=r)
        calls = []

        def reducer(mine, r=r, calls=calls):      # stands in for dist.all_reduce: local matrices in, the sum out
            assert torch.equal(mine, local[r])
            calls.append(1)
            return total

        Xl, Xql = X[r * ml:(r + 1) * ml].contiguous(), Xq[r * ml:(r + 1) * ml].contiguous()
        Qr, er, rr = sa.quantize_layer_impl(W, Xl, Xql, ml, step, K, 1, reg, lam, 1, False, DEV, neuron_range=(n0, n1),
                                            return_partials=True, rows_split_over=reducer)
        assert len(calls) == 1
        Q[n0:n1], e2[n0:n1], r2[n0:n1] = Qr[n0:n1], er[n0:n1], rr[n0:n1]
    err, rel, _, _ = sa.reduce_errors(e2, r2, 1)
    return Q, err, rel


@pytest.mark.parametrize("G", [2, 4, 8])
@pytest.mark.parametrize("reg,lam", [(None, 0.0), ("L1", 0.003)])
def test_row_sharded_gram_reduce_matches_oracle(G, reg, lam):
    from quantized_neural_nets_b200 import step_algorithm as sa
    W, X, Xq = gc._problem(seed=90 + G, N=136, d=72, m=4096, relu=True, xq_noise=0.02, zero_xq=(11,))
    K, step = 8, 1.16 / 8
    Qo, erro, relo, _, _ = orc.quantize_layer(W, X, Xq, X.shape[0], step, K, 1, reg, lam, 1, False)
    Q, err, rel = _serial_row_shards(sa, W.to(DEV), X.to(DEV), Xq.to(DEV), G, step, K, reg, lam)
    delta = orc.layer_step_size(W, step, K, 1, reg, lam)
    lv, lvo = orc.level_index(Q.cpu(), delta, reg, lam), orc.level_index(Qo, delta, reg, lam)
    agree = (lv == lvo).float().mean().item()
    assert agree >= 0.999, (G, agree)
    if agree < 1.0:
        margin = orc.exact_decision_margin(W, X, Xq, Qo, delta, K, reg, lam)
        diff = lv != lvo
        for n in diff.any(dim=1).nonzero().flatten().tolist():
            assert margin[n, int(diff[n].nonzero()[0])] < 1e-4
    assert abs(rel.item() - relo.item()) <= 1e-3 * relo.item()
    assert abs(err.item() - erro.item()) <= 1e-3 * erro.item()


def test_resnet50_row_sharded_layers_teacher_forced():
    """Every ResNet-50 layer the sharded forward would solve from all-reduced Gram matrices (gram_reduce_eligible),
    teacher-forced at batch 16 with G = 4 emulated ranks, against the oracle: the parity evidence of the mode the
    multi-GPU bench measures."""
    import quantized_neural_nets_b200 as qb
    from quantized_neural_nets_b200 import step_algorithm as sa
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).eval().to(DEV)
    layers = []
    qb.extract_layers(model, layers)
    batch, G, K = 16, 4, 8
    g = torch.Generator().manual_seed(1)
    loader = [(torch.randn(batch, 3, 224, 224, generator=g), None) for _ in layers]
    np.random.seed(0)
    qnn = qb.QuantizeNeuralNet(model, "resnet50", batch, loader, 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25, False, DEV)
    checked, weights, same = 0, 0.0, 0.0
    for i, layer in enumerate(qnn.analog_network_layers):
        X, Xq = qnn._populate_linear_layer_input(i)
        W = layer.weight.data.view(layer.weight.shape[0], -1)
        N, d = W.shape
        m = X.shape[0]
        if m % G == 0 and sa.gram_reduce_eligible(N, d, m):
            Q, err, rel = _serial_row_shards(sa, W, X.contiguous(), Xq.contiguous(), G, 1.16 / K, K, None, 0.1)
            Wc, Xc, Xqc = W.cpu(), X.cpu().contiguous(), Xq.cpu().contiguous()
            Qo, erro, relo, _, _ = orc.quantize_layer(Wc, Xc, Xqc, m, 1.16 / K, K, 1, None, 0.1, 1, False)
            delta = orc.layer_step_size(Wc, 1.16 / K, K, 1, None, 0.1)
            agree = (orc.level_index(Q.cpu(), delta) == orc.level_index(Qo, delta)).float().mean().item()
            assert agree >= 0.99, (i, tuple(W.shape), m, agree)
            assert abs(rel.item() - float(relo)) <= 1e-3 * float(relo), (i, rel.item(), float(relo))
            checked += 1
            weights += N * d
            same += agree * N * d
        else:
            Q = qb.StepAlgorithm._quantize_layer(W, X, Xq, m, 1.16 / K, K, 1, None, 0.1, layer.groups
                                                 if hasattr(layer, "groups") else 1, False, DEV)[0]
        qnn.quantized_network_layers[i].weight.data = Q.reshape(layer.weight.shape).float()
    assert checked >= 20, checked
    assert same / weights >= 0.999, same / weights


def test_vgg16_soft_threshold_teacher_forced_including_fc6():
    """BASELINE.json configs[4]: VGG-16, reg='L1'; every layer teacher-forced against the oracle, fc6 = 4096 x 25088."""
    from test_gpu_networks import run_teacher_forced, check
    report = run_teacher_forced("vgg16", batch=4, reg="L1", lam=1e-3)
    assert len(report) == 16
    assert report[13][1] == (4096, 25088)
    assert check(report) >= 0.999


def test_resnet50_three_bit_teacher_forced():
    """BASELINE.json configs[3] names 3 and 4 bits: K = 2^(3-1) = 4, alphabet delta*{-4..4} (quantize_neural_net.py:87-93)."""
    from test_gpu_networks import run_teacher_forced, check
    report = run_teacher_forced("resnet50", batch=8, bits=3, solver=1)
    assert len(report) == 54
    assert check(report) >= 0.999


def test_resnet18_free_running_logits_vs_oracle():
    """The reference's only end-to-end check is the quantized model's outputs (main.py:153-155).  Free-running
    quantize_network() of a random-init ResNet-18 (batch 8, 4 bits) on the GPU against the oracle's quantize_network on
    the CPU, same seeds and batches, compared on a held-out batch.

    The tolerance is MEASURED ON THE REFERENCE ITSELF (tests/golden/resnet18_sensitivity.json, written by
    tests/golden/make_sensitivity.py from the unmodified reference): GPFQ is chaotic -- a decision at a rounding tie
    flips under any last-bit change of its argument and every later layer then sees a different quantized input -- so
    the reference run twice with its images perturbed by ONE ulp keeps only 44.6 % of its weights and moves its own
    logits by 6.9e-2 (relative L2; the quantization itself moves them by 5.7e-2).  Two correct fp32 implementations
    with different convolution kernels cannot be closer than that, so the gate is: logits within 1.5 x the
    reference's own one-ulp sensitivity, the same quantization quality (distance from the fp32 network's logits within
    15 %), the leading layers (before the first tie flips) identical, and overall weight agreement no worse than 0.8 x
    the reference's own.  Layer-by-layer parity proper is the teacher-forced tests."""
    import copy
    import json
    import os
    import quantized_neural_nets_b200 as qb
    sens = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "resnet18_sensitivity.json")))
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet18(weights=None).eval()
    batch = 8
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(batch, 3, 224, 224, generator=g), None) for _ in range(21)]
    probe = torch.randn(8, 3, 224, 224, generator=g)
    np.random.seed(0)
    log = []
    q_cpu = orc.quantize_network(copy.deepcopy(model), batches, mlp_bits=4, cnn_bits=4, log=log)
    np.random.seed(0)
    qnn = qb.QuantizeNeuralNet(copy.deepcopy(model).to(DEV), "resnet18", batch, batches, 4, 4, [], 1.16, 1.16, 1, 1, None,
                               0.1, 0.25, False, DEV)
    q_gpu = qnn.quantize_network()
    with torch.no_grad():
        want, got, fp = q_cpu(probe), q_gpu(probe.to(DEV)).cpu(), model(probe)
    # the oracle run is the reference's run A (same seeds) up to this host's CPU kernels: its distance from the fp32
    # network's logits must be the fixture's, to the accuracy the chaos allows
    effect_o = ((want - fp).norm() / fp.norm()).item()
    assert abs(effect_o - sens["logits_rel_l2_A_vs_fp32"]) <= 0.15 * sens["logits_rel_l2_A_vs_fp32"]
    rel = ((got - want).norm() / want.norm()).item()
    effect_g = ((got - fp).norm() / fp.norm()).item()
    layers_o, layers_g = [], []
    orc.extract_layers(q_cpu, layers_o)
    qb.extract_layers(q_gpu, layers_g)
    per_layer = [torch.isclose(a.weight.data, b.weight.data.cpu(), rtol=1e-6, atol=0).float().mean().item()
                 for a, b in zip(layers_o, layers_g)]
    tot = sum(l.weight.numel() for l in layers_o)
    same = sum(f * l.weight.numel() for f, l in zip(per_layer, layers_o)) / tot
    rels = [abs(float(r) - ro) / ro for (_, _, r), (_, _, ro) in zip(qnn.layer_log, log)]
    print(f"resnet18 free-running vs oracle: logits rel-L2 {rel:.3e} (reference's one-ulp sensitivity "
          f"{sens['logits_rel_l2_A_vs_B']:.3e}); distance from fp32 logits {effect_g:.3e} (oracle {effect_o:.3e}); identical "
          f"weights {same:.4f} (reference vs itself {sens['identical_weights']:.4f}); worst per-layer rel-err deviation "
          f"{max(rels):.2e} (reference {sens['worst_layer_rel_err_deviation']:.2e})")
    assert rel <= 1.5 * sens["logits_rel_l2_A_vs_B"], rel
    assert abs(effect_g - effect_o) <= 0.15 * effect_o, (effect_g, effect_o)
    assert min(per_layer[:4]) >= 0.999, per_layer[:4]
    assert same >= 0.8 * sens["identical_weights"], same
    assert max(rels) <= 2.0 * sens["worst_layer_rel_err_deviation"], max(rels)
    assert (got.argmax(1) == want.argmax(1)).float().mean().item() >= 0.75


@pytest.mark.parametrize("K,reg,lam", [(128, None, 0.0), (128, "L0", 0.002), (128, "L1", 0.002)])
def test_eight_bit_alphabet(K, reg, lam):
    """--bits 8 (K = 128): the reference has no bound on K; only the optional int8 `levels` output has one."""
    from quantized_neural_nets_b200 import _lib, step_algorithm as sa
    from quantized_neural_nets_b200.export import pack_layer, unpack_layer
    W, X, Xq = gc._problem(seed=95, N=70, d=60, m=300, relu=True, xq_noise=0.02)
    step = 1.16 / K
    Qo, erro, relo, _, _ = orc.quantize_layer(W, X, Xq, X.shape[0], step, K, 1, reg, lam, 1, False)
    delta = orc.layer_step_size(W, step, K, 1, reg, lam)
    lvo = orc.level_index(Qo, delta, reg, lam)
    assert int(lvo.abs().max()) > 16          # the run really uses the wide alphabet
    for solver in (_lib.SOLVER_DIRECT, _lib.SOLVER_GRAM_F64):
        Q, e2, r2 = sa.quantize_layer_impl(W.to(DEV), X.to(DEV), Xq.to(DEV), X.shape[0], step, K, 1, reg, lam, 1, False,
                                           DEV, solver=solver, return_partials=True)
        lv = orc.level_index(Q.cpu(), delta, reg, lam)
        diff = lv != lvo
        # a 257-level alphabet has 16 x more rounding boundaries per unit than the 4-bit one: a neuron may leave the
        # oracle's path at a genuine tie (and then differs for the rest of its row), nowhere else
        assert (~diff).float().mean().item() >= 0.97, (K, reg, solver)
        if diff.any():
            margin = orc.exact_decision_margin(W, X, Xq, Qo, delta, K, reg, lam)
            for n in diff.any(dim=1).nonzero().flatten().tolist():
                assert margin[n, int(diff[n].nonzero()[0])] < 2e-5 * K, (n, float(margin[n, int(diff[n].nonzero()[0])]))
        else:
            rel = (e2.sum().sqrt() / (X.to(DEV) @ W.to(DEV).T).norm()).item()
            assert abs(rel - relo.item()) <= 1e-3 * relo.item()
    # int8 level indices cannot hold this alphabet in L0 mode (levels -129..129): it must fail loudly, not wrap around
    if reg == "L0":
        with pytest.raises(RuntimeError):
            _solve_with_levels(sa, W, X, Xq, delta, K, reg, lam)
    # packed export: 9-bit codes round-trip losslessly
    packed = pack_layer(Q, delta, K, reg, lam)
    assert packed.bits == 9
    assert torch.equal(unpack_layer(packed).view_as(Q), Q)


def test_sixteen_bit_codes_round_trip():
    """Packed export up to 16-bit codes (K = 2^15 - 2): values on the alphabet survive pack -> unpack bit for bit."""
    from quantized_neural_nets_b200.export import pack_layer, unpack_layer
    K = 2 ** 15 - 2
    delta = torch.tensor(3.0517578125e-05)
    lv = torch.randint(-K, K + 1, (37, 53), generator=torch.Generator().manual_seed(5)).float()
    Q = (torch.sign(lv) * delta * lv.abs()).to(DEV)          # sign * delta * k, the alphabet map's operation order
    packed = pack_layer(Q, delta, K)
    assert packed.bits == 16
    assert torch.equal(unpack_layer(packed), Q)


def _solve_with_levels(sa, W, X, Xq, delta, K, reg, lam):
    Wd = W.to(DEV)
    Xfm, ld = sa.feature_major(X.to(DEV))
    Xqfm, _ = sa.feature_major(Xq.to(DEV))
    lv = torch.zeros(W.shape, dtype=torch.int8, device=DEV)
    sa.solve_rows(Wd, Xfm, Xqfm, ld, X.shape[0], sa._delta_tensor(delta, DEV), K, sa.mode_of(reg, False), lam,
                  torch.zeros_like(Wd), 0, W.shape[0], levels=lv)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device():
    """The reference takes a `device` argument and works on any device without set_device; so must the drop-in
    (the C side launches on the CURRENT device)."""
    import quantized_neural_nets_b200 as qb
    assert torch.cuda.current_device() == 0
    dev1 = torch.device("cuda:1")
    W, X, Xq = gc._problem(seed=96, N=40, d=50, m=120, relu=True, xq_noise=0.02)
    a = qb.StepAlgorithm._quantize_layer(W.to(DEV), X.to(DEV), Xq.to(DEV), 120, 1.16 / 8, 8, 1, None, 0.1, 1, False, DEV)
    b = qb.StepAlgorithm._quantize_layer(W.to(dev1), X.to(dev1), Xq.to(dev1), 120, 1.16 / 8, 8, 1, None, 0.1, 1, False, dev1)
    assert torch.cuda.current_device() == 0
    assert b[0].device == dev1 and torch.equal(a[0].cpu(), b[0].cpu()) and float(a[2]) == float(b[2])
    hook = qb.SaveInputConv2d(3, 1, 1, 1, 1, 0.5)
    np.random.seed(1)
    with pytest.raises(qb.InterruptException):
        hook(None, (torch.randn(2, 4, 9, 9, device=dev1),), None)
    assert hook.inputs[0].device == dev1


def test_mixed_devices_are_rejected():
    from quantized_neural_nets_b200._lib import lib, launch
    x = torch.zeros(8, device=DEV)
    with pytest.raises((TypeError, ValueError)):
        launch(lib.gpfq_quantize_f32, x, torch.zeros(8), 8, torch.ones(1, device=DEV), 8, 0, 0.0, 0)


# ------------------------------------------------------------------------------------------------------------------
# calibration forward: tcgen05 1x1 convolution fused with BatchNorm / residual / ReLU (gpfq_conv1x1_bn_act_f32)
@pytest.mark.parametrize("B,C,N,H,W,with_res,lo,hi", [
    (3, 64, 64, 8, 8, False, 0.0, float("inf")),            # one k-block pair, one tile
    (2, 256, 64, 56, 56, False, 0.0, float("inf")),         # ResNet-50 layer1 conv1: N < tile, 25 pixel tiles, tail 64 px
    (2, 64, 256, 56, 56, True, 0.0, float("inf")),          # conv3 + residual
    (5, 72, 200, 14, 14, True, 0.0, 6.0),                   # ragged C (zero-filled k tail), ragged N, HW = 196
    (2, 40, 24, 6, 6, False, -float("inf"), float("inf")),  # HW = 36 < one tile, no clamp
    (1, 1024, 300, 14, 14, False, 0.0, float("inf")),       # 32 k-blocks: the ring and the four accumulators wrap
    (160, 32, 130, 4, 4, True, 0.0, float("inf")),          # many tiny tiles per CTA (persistent loop)
])
def test_conv1x1_tensor_core_kernel(B, C, N, H, W, with_res, lo, hi):
    """fp32-level accuracy against float64 and exactness of the fused epilogue.

    Gates: relative L2 error of the convolution <= 2e-7 (measured 1.2-1.7e-7; cuDNN's fp32 kernels ~0.6e-7) and
    worst single output <= 1e-6 of its sum of |terms| -- the split keeps 22 of the 24 significand bits of the leading
    products (|x - hi - lo'| <= 2^-21 |x| after the tensor core reads lo at TF32 width) and the tensor core adds the 12
    partial products of a 32-channel k-block with truncation, so a short reduction (C = 64) shows up to 5e-7 on its
    worst element where fp32 FMA chains show 2e-7; a single TF32 pass would show 5e-4.  The fused epilogue must be
    bit-identical to the separate elementwise pass applied to the same kernel's plain convolution output."""
    from quantized_neural_nets_b200._lib import lib, launch
    g = torch.Generator().manual_seed(B * 1000 + C)
    x = torch.relu(torch.randn(B, C, H, W, generator=g)).to(DEV)
    w = (torch.randn(N, C, generator=g) * 0.1).to(DEV)
    alpha = (torch.rand(N, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(N, generator=g) * 0.1).to(DEV)
    res = torch.randn(B, N, H, W, generator=g).to(DEV) if with_res else None
    assert lib.gpfq_conv1x1_fused_supported(C, N, H * W, H * W) == 1
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(N, C), dtype=torch.uint8, device=DEV)
    plain = torch.full((B, N, H, W), float("nan"), device=DEV)
    launch(lib.gpfq_conv1x1_bn_act_f32, x, H * W, w, None, None, None, plain, B, C, N, H * W, -float("inf"), float("inf"), ws,
           ws.numel())
    ref = torch.einsum("nc,bchw->bnhw", w.double(), x.double())
    mag = torch.einsum("nc,bchw->bnhw", w.double().abs(), x.double().abs())
    err = ((plain.double() - ref).abs() / (mag + 1e-30)).max().item()
    l2 = ((plain.double() - ref).norm() / ref.norm()).item()
    assert err <= 1e-6, err
    assert l2 <= 2e-7, l2
    fused = torch.full((B, N, H, W), float("nan"), device=DEV)
    launch(lib.gpfq_conv1x1_bn_act_f32, x, H * W, w, res, alpha, beta, fused, B, C, N, H * W, lo, hi, ws, ws.numel())
    two_pass = torch.empty_like(plain)
    launch(lib.gpfq_bn_act_f32, plain, res, alpha, beta, two_pass, B * N, N, H * W, lo, hi)
    assert torch.equal(fused, two_pass)
    # and against cuDNN's fp32 convolution
    torch.backends.cudnn.allow_tf32 = False
    cud = torch.nn.functional.conv2d(x, w.view(N, C, 1, 1))
    assert (plain - cud).norm() <= 2e-6 * cud.norm()


def test_conv1x1_two_call_route_and_plane_cache():
    """gpfq_conv1x1_split_weight_f32 + gpfq_conv1x1_bn_act_planes_f32 give the bits of the one-call route, tiles that
    straddle images included (B * ceil(HW / 32) chunks not a multiple of 4), and FusedConvBNAct re-splits a weight whose
    value changed (in place or by assignment) and only then."""
    from quantized_neural_nets_b200 import _lib
    from quantized_neural_nets_b200._lib import lib, launch
    from quantized_neural_nets_b200.forward_fusion import FusedConvBNAct
    g = torch.Generator().manual_seed(7)
    B, C, N, H, W = 7, 96, 160, 14, 14          # 7 chunks per image, 49 chunks: tiles straddle images, last tile 1 chunk
    x = torch.relu(torch.randn(B, C, H, W, generator=g)).to(DEV)
    w = (torch.randn(N, C, generator=g) * 0.1).to(DEV)
    alpha = (torch.rand(N, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(N, generator=g) * 0.1).to(DEV)
    res = torch.randn(B, N, H, W, generator=g).to(DEV)
    ws = torch.empty(lib.gpfq_conv1x1_workspace_bytes(N, C), dtype=torch.uint8, device=DEV)
    one = torch.full((B, N, H, W), float("nan"), device=DEV)
    launch(lib.gpfq_conv1x1_bn_act_f32, x, H * W, w, res, alpha, beta, one, B, C, N, H * W, 0.0, float("inf"), ws, ws.numel())
    planes = torch.empty_like(ws)
    launch(lib.gpfq_conv1x1_split_weight_f32, w, N, C, planes, planes.numel())
    two = torch.full((B, N, H, W), float("nan"), device=DEV)
    launch(lib.gpfq_conv1x1_bn_act_planes_f32, x, H * W, res, alpha, beta, two, B, C, N, H * W, 0.0, float("inf"), planes,
           planes.numel())
    assert torch.equal(one, two)
    ref = torch.relu(torch.einsum("nc,bchw->bnhw", w.double(), x.double()) * alpha.double()[None, :, None, None]
                     + beta.double()[None, :, None, None] + res.double())
    assert (one.double() - ref).norm() <= 2e-7 * ref.norm()
    # the module's cache
    conv = torch.nn.Conv2d(C, N, 1, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(N).eval().to(DEV)
    mod = FusedConvBNAct(conv, bn, 0.0, float("inf"))
    with torch.no_grad():
        y0 = mod(x)
        n0 = _lib.launch_count()
        y1 = mod(x)
        assert _lib.launch_count() - n0 == 1 and torch.equal(y0, y1)        # planes reused: one launch
        conv.weight.mul_(2.0)                                                # in-place change
        y2 = mod(x)
        assert torch.allclose(y2, torch.relu(bn(conv(x))), rtol=1e-5, atol=1e-5)
        conv.weight.data = torch.randn(N, C, 1, 1, device=DEV) * 0.1        # assignment (what quantize_network does)
        n0 = _lib.launch_count()
        y3 = mod(x)
        assert _lib.launch_count() - n0 == 2
        assert torch.allclose(y3, torch.relu(bn(conv(x))), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("C,N,H,W,k,stride,pad,bias", [(3, 64, 40, 40, 7, 2, 3, False),      # ResNet stem
                                                      (32, 48, 28, 28, 3, 2, 1, False),    # stride-2 3x3
                                                      (64, 96, 14, 14, 1, 2, 0, False),    # stride-2 shortcut, 7 x 7 output
                                                      (96, 40, 7, 7, 1, 1, 0, True),       # 7 x 7 planes: padded pitch; bias
                                                      (16, 24, 9, 11, 3, 2, 0, True)])
def test_fused_conv_bn_act_module_any_convolution(C, N, H, W, k, stride, pad, bias):
    """FusedConvBNAct (patch matrix + tensor-core GEMM + fused epilogue) against PyTorch's conv2d -> BatchNorm2d -> + r -> ReLU."""
    from quantized_neural_nets_b200.forward_fusion import FusedConvBNAct
    g = torch.Generator().manual_seed(C * 100 + N)
    conv = torch.nn.Conv2d(C, N, k, stride=stride, padding=pad, bias=bias).to(DEV)
    bn = torch.nn.BatchNorm2d(N).eval()
    bn.running_mean = torch.randn(N, generator=g) * 0.1
    bn.running_var = torch.rand(N, generator=g) + 0.5
    bn.weight.data = torch.rand(N, generator=g) + 0.5
    bn.bias.data = torch.randn(N, generator=g) * 0.1
    bn = bn.to(DEV)
    x = torch.randn(5, C, H, W, generator=g).to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        y = bn(conv(x))
        res = torch.randn(y.shape, generator=g).to(DEV)
        want = torch.relu(y + res)
        mod = FusedConvBNAct(conv, bn, 0.0, float("inf"))
        assert mod.route is not None
        from quantized_neural_nets_b200 import _lib
        before = _lib.launch_count()
        got = mod(x, res)
        assert _lib.launch_count() - before >= 2        # the tensor-core path ran (weight split + GEMM (+ patches))
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


def test_fused_resnet50_forward_with_tensor_core_convolutions():
    """The traced ResNet-50 with conv1x1 + BN (+ add) + ReLU sites on the tensor-core kernel: same logits as the plain
    network to fp32 accuracy, hooks on a fused convolution still fire (the site falls back to the modules)."""
    from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None).eval().to(DEV)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    fused, sites = fuse_inference_forward(model)
    assert sites == 53 and fused.fused_conv_sites == 40      # 33 stride-1 1x1 + 3 stride-2 1x1 + 3 stride-2 3x3 + the stem
    x = torch.randn(8, 3, 224, 224, device=DEV)
    with torch.no_grad():
        want, got = model(x), fused(x)
    assert (got - want).norm() <= 2e-5 * want.norm(), ((got - want).norm() / want.norm()).item()
    seen = []
    handle = model.layer2[1].conv3.register_forward_hook(lambda m, i, o: seen.append((tuple(i[0].shape), tuple(o.shape))))
    with torch.no_grad():
        got2 = fused(x)
    handle.remove()
    assert seen == [((8, 128, 28, 28), (8, 512, 28, 28))]
    assert (got2 - want).norm() <= 2e-5 * want.norm()


@pytest.mark.parametrize("reg,lam,groups", [(None, 0.0, 1), ("L0", 0.004, 1), ("L1", 0.002, 1), (None, 0.0, 4)])
def test_slice_exchange_kernels_round_trip(reg, lam, groups):
    """The layer's all-gather payload (int8 levels + two fp64 norms per neuron, gpfq_pack_slice_f32) unpacked from the
    concatenation of 1 / 2 / 3 / 8 ranks' buffers reproduces the solver's Q bit for bit and the norms exactly."""
    from quantized_neural_nets_b200 import _lib, step_algorithm as sa
    from quantized_neural_nets_b200.sharding import neuron_slice, slice_rows
    lib, launch = _lib.lib, _lib.launch
    N, d, m, K = 52, 45, 200, 8
    W, X, Xq = gc._problem(seed=97, N=N, d=d * groups, m=m, relu=True, xq_noise=0.02)
    W = W[:, :d].contiguous()
    step = 1.16 / K
    Wd, Xd, Xqd = W.to(DEV), X.to(DEV), Xq.to(DEV)
    Q, e2, r2 = sa.quantize_layer_impl(Wd, Xd, Xqd, m, step, K, 1, reg, lam, groups, False, DEV, return_partials=True)
    delta = sa._delta_tensor(sa.layer_delta(Wd, step, K, 1, reg, lam), DEV)
    mode = sa.mode_of(reg, False)
    rb = int(lib.gpfq_slice_row_bytes(d))
    assert rb == (d + 7) // 8 * 8 + 16
    for world in (1, 2, 3, 8):
        per = slice_rows(N, groups, world)
        bufs = []
        for rank in range(world):
            n0, n1 = neuron_slice(N, groups, world=world, rank=rank)
            Qr = torch.zeros_like(Q)
            Qr[n0:n1] = Q[n0:n1]                   # what a rank holds: its own rows, zeros elsewhere
            buf = torch.full((per * rb,), 0xAB, dtype=torch.uint8, device=DEV)
            bad = torch.ones(1, dtype=torch.int32, device=DEV)
            launch(lib.gpfq_pack_slice_f32, Qr, Qr.stride(0), d, n0, n1, per, delta, K, mode, float(lam), e2, r2, buf, bad)
            assert int(bad.item()) == 0
            bufs.append(buf)
        full = torch.cat(bufs)
        Qf = torch.full((N, d), float("nan"), device=DEV)
        ef = torch.zeros(N, dtype=torch.float64, device=DEV)
        rf = torch.zeros(N, dtype=torch.float64, device=DEV)
        launch(lib.gpfq_unpack_slices_f32, full, N, d, delta, K, mode, float(lam), Qf, d, ef, rf)
        assert torch.equal(Qf, Q) and torch.equal(ef, e2) and torch.equal(rf, r2), world
    # a value off the alphabet is reported, not rounded away
    Qbad = Q.clone()
    Qbad[3, 7] += 1e-3
    buf = torch.zeros(N * rb, dtype=torch.uint8, device=DEV)
    bad = torch.zeros(1, dtype=torch.int32, device=DEV)
    launch(lib.gpfq_pack_slice_f32, Qbad, d, d, 0, N, N, delta, K, mode, float(lam), e2, r2, buf, bad)
    assert int(bad.item()) == 1


@pytest.mark.parametrize("shape,k,s,p", [((3, 5, 17, 23), 3, 2, 1), ((2, 4, 8, 8), 2, 2, 0), ((1, 3, 9, 9), 3, 1, 1),
                                         ((2, 64, 112, 112), 3, 2, 1), ((2, 3, 30, 31), 5, 3, 2),
                                         ((1, 2, 3, 3), 3, 2, 1), ((2, 3, 15, 24), 3, 2, 1), ((3, 2, 4, 4), 3, 2, 1)])
def test_maxpool_kernel_matches_pytorch(shape, k, s, p):
    from quantized_neural_nets_b200.forward_fusion import FastMaxPool
    x = torch.randn(shape, generator=torch.Generator().manual_seed(11)).to(DEV)
    x[0, 0, 1, 1] = float("nan")
    x[-1, -1, -1, -1] = float("nan")
    pool = torch.nn.MaxPool2d(k, s, p)
    got, want = FastMaxPool(pool)(x), pool(x)
    assert got.shape == want.shape
    assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))
    assert torch.isnan(got).sum() == torch.isnan(want).sum() > 0


@pytest.mark.parametrize("name,min_conv_sites", [("mobilenet_v2", 30), ("resnet18", 4), ("googlenet", 20)])
def test_fused_forward_of_other_model_families(name, min_conv_sites):
    """The fused calibration forward on the other families the reference's CLI whitelists: MobileNetV2 (1x1 layers with
    C = 16 ... 960 -- channel tails of a k-block --, ReLU6 clamps, depthwise layers left to cuDNN), ResNet-18 (only the
    stem and the shortcuts are 1x1 / strided), GoogLeNet (BasicConv2d: functional ReLU, eps = 1e-3)."""
    from quantized_neural_nets_b200.forward_fusion import fuse_inference_forward
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    kw = dict(weights=None)
    if name == "googlenet":
        kw.update(aux_logits=False, init_weights=True)
    model = getattr(torchvision.models, name)(**kw).eval().to(DEV)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.1)
            mod.running_var.uniform_(0.5, 1.5)
    fused, sites = fuse_inference_forward(model)
    assert fused.fused_conv_sites >= min_conv_sites, fused.fused_conv_sites
    x = torch.randn(4, 3, 224, 224, device=DEV)
    with torch.no_grad():
        want, got = model(x), fused(x)
    assert torch.isfinite(got).all()
    assert (got - want).norm() <= 5e-5 * want.norm(), ((got - want).norm() / want.norm()).item()


def test_stem_patch_matrix_is_shared_between_the_two_passes():
    """The analog and the quantized pass of a layer read the same image batch: the designated batch's patch matrix is
    built once (forward_fusion.share_patches_of) and the second pass gives the same bits as an unshared one."""
    from quantized_neural_nets_b200 import forward_fusion as ff
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(3, 64, 7, 2, 3, bias=False).to(DEV)
    bn = torch.nn.BatchNorm2d(64).eval().to(DEV)
    bn.running_mean.normal_(0, 0.1), bn.running_var.uniform_(0.5, 1.5)
    site = ff.FusedConvBNAct(conv, bn, 0.0, float("inf"))
    x = torch.randn(4, 3, 64, 64, device=DEV)
    with torch.no_grad():
        plain = site(x)
        ff.share_patches_of(x)
        try:
            first = site(x)
            assert len(ff._SharedPatches.store) == 1
            cached = next(iter(ff._SharedPatches.store.values()))
            second = site(x)
            assert len(ff._SharedPatches.store) == 1 and next(iter(ff._SharedPatches.store.values())) is cached
            x.add_(1.0)                                    # an in-place change of the batch must not hit the cache
            third = site(x)
            assert len(ff._SharedPatches.store) == 2
        finally:
            ff.release_shared_patches()
        assert ff._SharedPatches.source is None and not ff._SharedPatches.store
        assert torch.equal(plain, first) and torch.equal(plain, second)
        assert torch.equal(third, site(x)) and not torch.equal(third, plain)
