"""Parity of the CUDA path (through the C ABI of libgpfq_b200.so) against the golden outputs of
the reference and against the CPU oracle on the same seeded inputs.  Needs a B200: -m gpu.

Tolerances (BASELINE.json north_star): >= 99.9 % identical quantization levels per layer,
differences only at rounding ties; per-layer relative error within 1e-3 (relative) of the
reference's.  The small golden cases are expected to match bit for bit."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from oracle import gpfq_oracle as orc

pytestmark = pytest.mark.gpu

DEV = torch.device("cuda:0")


@pytest.fixture(scope="module")
def qb():
    import quantized_neural_nets_b200 as qb
    return qb


def cuda_quantizer(qb, mode):
    SA = qb.StepAlgorithm
    return {"msq": SA._msq, "soft": SA._soft_thresholding_msq, "hard": SA._hard_thresholding_msq}[mode]


def test_quantizer_tables_bit_exact(qb, golden):
    g = golden("quantizers.npz")
    for tag, (x, delta, K, lam) in gc.quantizer_inputs().items():
        for mode in ("msq", "soft", "hard"):
            got = cuda_quantizer(qb, mode)(delta.to(DEV), x.to(DEV), K, lam).cpu().numpy()
            want = g[f"{tag}_{mode}"]
            np.testing.assert_array_equal(got, want, err_msg=f"{tag}/{mode}")
            np.testing.assert_array_equal(np.signbit(got), np.signbit(want), err_msg=f"{tag}/{mode} signed zero")


@pytest.mark.parametrize("tag", list(gc.greedy_inputs().keys()))
def test_greedy_path_matches_reference(qb, golden, tag):
    g = golden("greedy_path.npz")
    c = gc.greedy_inputs()[tag]
    W, X, Xq = c["W"].to(DEV), c["X"].to(DEV), c["Xq"].to(DEV)
    Q = torch.zeros_like(W)
    U = torch.zeros(W.shape[0], X.shape[0], device=DEV)
    qb.StepAlgorithm._quantization(W, Q, U, X, Xq, cuda_quantizer(qb, c["mode"]), c["delta"].to(DEV), c["K"], c["lam"])
    np.testing.assert_array_equal(Q.cpu().numpy(), g[f"{tag}_Q"])
    # identical decisions => the residual is reproduced bit for bit (same update order and roundings)
    np.testing.assert_array_equal(U.cpu().numpy(), g[f"{tag}_U"])


@pytest.mark.parametrize("tag", list(gc.layer_inputs().keys()))
def test_quantize_layer_matches_reference(qb, golden, tag):
    g = golden("quantize_layer.npz")
    c = gc.layer_inputs()[tag]
    W, X, Xq = c["W"].to(DEV), c["X"].to(DEV), c["Xq"].to(DEV)
    Q, err, rel, adder, rel_adder = qb.StepAlgorithm._quantize_layer(
        W, X, Xq, X.shape[0], c["step"], c["K"], c["pct"], c["reg"], c["lam"], c["groups"], False, DEV)
    assert isinstance(err, torch.Tensor) and err.dim() == 0 and isinstance(rel, torch.Tensor)
    # delta is averaged on the host exactly as the reference does, so Q matches bit for bit
    np.testing.assert_array_equal(Q.cpu().numpy(), g[f"{tag}_Q"])
    delta = orc.layer_step_size(c["W"], c["step"], c["K"], c["pct"], c["reg"], c["lam"])
    lv_got = orc.level_index(Q.cpu(), delta, c["reg"], c["lam"])
    lv_want = orc.level_index(torch.from_numpy(g[f"{tag}_Q"]), delta, c["reg"], c["lam"])
    assert torch.equal(lv_got, lv_want)
    np.testing.assert_allclose(err.item(), g[f"{tag}_err"], rtol=1e-5)
    np.testing.assert_allclose(rel.item(), g[f"{tag}_rel"], rtol=1e-5)
    if c["groups"] == 1:
        np.testing.assert_allclose(adder.cpu().numpy(), g[f"{tag}_adder"], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(rel_adder.cpu().numpy(), g[f"{tag}_rel_adder"], rtol=1e-4)
    else:
        assert adder is None and rel_adder is None


@pytest.mark.parametrize("tag", list(gc.conv_inputs().keys()))
def test_conv_capture_bit_exact(qb, golden, tag):
    g = golden("conv_capture.npz")
    c = gc.conv_inputs()[tag]
    hook = qb.SaveInputConv2d(c["kernel"], c["dilation"], c["padding"], c["stride"], c["groups"], c["p"])
    np.random.seed(c["np_seed"])
    for which in ("a", "q"):
        with pytest.raises(qb.InterruptException):
            hook(None, (c["inp_" + which].to(DEV),), None)
    np.testing.assert_array_equal(hook.rand_indices, g[f"{tag}_idx"])
    np.testing.assert_array_equal(hook.inputs[0].cpu().numpy(), g[f"{tag}_rows_a"])
    np.testing.assert_array_equal(hook.inputs[1].cpu().numpy(), g[f"{tag}_rows_q"])


@pytest.mark.parametrize("tag", list(gc.network_inputs().keys()))
def test_tiny_network_matches_reference(qb, golden, tag):
    g = golden("tiny_network.npz")
    c = gc.network_inputs()[tag]
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = c["model"].to(DEV)
    np.random.seed(c["np_seed"])
    qnn = qb.QuantizeNeuralNet(model, "tiny", c["batch"], c["loader"](), c["bits"], c["bits"], c["ignore"],
                               c["scalar"], c["scalar"], 1, 1, c["reg"], c["lam"], c["p"], False, DEV)
    qmodel = qnn.quantize_network()
    for i, layer in enumerate(qnn.quantized_network_layers):
        want = g[f"{tag}_layer{i}"]
        got = layer.weight.data.cpu().numpy()
        # free-running: later layers see cuDNN-vs-CPU forward differences; allow rare tie flips
        same = np.isclose(got, want, rtol=1e-6, atol=0).mean()
        assert same >= 0.99, (tag, i, same)
    with torch.no_grad():
        logits = qmodel(c["probe"].to(DEV)).cpu().numpy()
    assert np.linalg.norm(logits - g[f"{tag}_logits"]) <= 2e-2 * np.linalg.norm(g[f"{tag}_logits"])


@pytest.mark.parametrize("reg,lam,relu", [(None, 0.0, True), ("L1", 0.003, True), ("L0", 0.003, False)])
def test_medium_layer_vs_oracle(qb, reg, lam, relu):
    """Teacher-forced parity on a layer big enough to exercise every tail path
    (N, d, m not multiples of the tile sizes; several feature blocks; zero-norm column)."""
    W, X, Xq = gc._problem(seed=77, N=203, d=150, m=1111, relu=relu, xq_noise=0.02, zero_xq=(5,))
    K, step = 8, 1.16 / 8
    Qo, erro, relo, addero, _ = orc.quantize_layer(W, X, Xq, X.shape[0], step, K, 1, reg, lam, 1, False)
    Q, err, rel, adder, _ = qb.StepAlgorithm._quantize_layer(W.to(DEV), X.to(DEV), Xq.to(DEV), X.shape[0], step, K, 1,
                                                            reg, lam, 1, False, DEV)
    delta = orc.layer_step_size(W, step, K, 1, reg, lam)
    lv, lvo = orc.level_index(Q.cpu(), delta, reg, lam), orc.level_index(Qo, delta, reg, lam)
    agree = (lv == lvo).float().mean().item()
    assert agree >= 0.999, agree
    if agree < 1.0:   # every first divergence of a neuron must sit at a rounding tie
        margin = orc.exact_decision_margin(W, X, Xq, Qo, delta, K, reg, lam)
        diff = (lv != lvo)
        for n in diff.any(dim=1).nonzero().flatten().tolist():
            t = int(diff[n].nonzero()[0])
            assert margin[n, t] < 1e-4, (n, t, float(margin[n, t]))
    assert abs(rel.item() - relo.item()) <= 1e-3 * relo.item()
    assert abs(err.item() - erro.item()) <= 1e-3 * erro.item()


def test_shard_count_invariance(qb):
    """Q solved as 1, 2, 4, 8 neuron slices is bit-identical (per-neuron arithmetic does not depend
    on the slice) -- SURVEY.md section 8e."""
    from quantized_neural_nets_b200.step_algorithm import quantize_layer_impl
    from quantized_neural_nets_b200.sharding import neuron_slice
    W, X, Xq = gc._problem(seed=78, N=300, d=96, m=700, relu=True, xq_noise=0.02)
    W, X, Xq = W.to(DEV), X.to(DEV), Xq.to(DEV)
    full = None
    for world in (1, 2, 4, 8):
        Q = torch.zeros_like(W)
        e2 = torch.zeros(W.shape[0], dtype=torch.float64, device=DEV)
        for rank in range(world):
            n0, n1 = neuron_slice(W.shape[0], 1, world=world, rank=rank)
            Qr, er, _ = quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / 8, 8, 1, None, 0.1, 1, False, DEV,
                                            neuron_range=(n0, n1), return_partials=True)
            Q[n0:n1] = Qr[n0:n1]
            e2[n0:n1] = er[n0:n1]
            assert Qr[:n0].abs().sum() == 0 and Qr[n1:].abs().sum() == 0
        if full is None:
            full = (Q.clone(), e2.clone())
        else:
            assert torch.equal(Q, full[0])
            assert torch.equal(e2, full[1])


def test_config1_full_size(qb, golden):
    """BASELINE.json configs[0] at full size against the reference's own levels."""
    g = golden("config1.npz")
    c = gc.config1_inputs()
    if float(g["inp"]) != gc.checksum(c["W"], c["X"]):
        pytest.skip("seeded inputs regenerate differently on this machine")
    X = c["X"].to(DEV)
    Q, err, rel, _, _ = qb.StepAlgorithm._quantize_layer(c["W"].to(DEV), X, X, 2048, c["step"], c["K"], 1, None, 0.1, 1,
                                                        False, DEV)
    lv = orc.level_index(Q.cpu(), torch.tensor(float(g["delta"]))).numpy()
    agree = (lv == g["levels"]).mean()
    assert agree >= 0.999, agree
    assert abs(rel.item() - float(g["rel"])) <= 1e-3 * float(g["rel"])


def test_errors_are_loud(qb):
    with pytest.raises(RuntimeError):
        qb.StepAlgorithm._quantize_layer(torch.zeros(4, 4), torch.zeros(8, 4), torch.zeros(8, 4), 8, 0.1, 8, 1, None, 0.1,
                                         1, False, torch.device("cpu"))
    with pytest.raises(NotImplementedError):
        z = torch.zeros(4, 4, device=DEV)
        qb.StepAlgorithm._quantization(z, z.clone(), torch.zeros(4, 8, device=DEV), torch.zeros(8, 4, device=DEV),
                                       torch.zeros(8, 4, device=DEV), (lambda *a: None), 0.1, 8, 0.0)


@pytest.mark.parametrize("solver", [2, 1], ids=["gram_f64", "gram_tcgen05"])
@pytest.mark.parametrize("reg,lam", [(None, 0.0), ("L1", 0.003), ("L0", 0.003)])
def test_gram_solvers_vs_oracle(qb, reg, lam, solver):
    """Gram-form solvers (fp64 SIMT and tcgen05 split-TF32 Gram matrices) against the oracle on a layer with
    m >> d and N >= d, the regime they are selected for; levels, both error norms and the denominators must match."""
    from quantized_neural_nets_b200 import _lib
    from quantized_neural_nets_b200.step_algorithm import quantize_layer_impl, reduce_errors
    W, X, Xq = gc._problem(seed=79, N=150, d=70, m=2500, relu=True, xq_noise=0.02, zero_xq=(9,))
    K, step = 8, 1.16 / 8
    Qo, erro, relo, _, rel_addo = orc.quantize_layer(W, X, Xq, X.shape[0], step, K, 1, reg, lam, 1, False)
    Q, e2, r2 = quantize_layer_impl(W.to(DEV), X.to(DEV), Xq.to(DEV), X.shape[0], step, K, 1, reg, lam, 1, False, DEV,
                                    solver=solver, return_partials=True)
    err, rel, _, rel_add = reduce_errors(e2, r2, 1)
    delta = orc.layer_step_size(W, step, K, 1, reg, lam)
    lv, lvo = orc.level_index(Q.cpu(), delta, reg, lam), orc.level_index(Qo, delta, reg, lam)
    assert (lv == lvo).float().mean().item() >= 0.999
    assert abs(rel.item() - relo.item()) <= 1e-3 * relo.item()
    assert abs(err.item() - erro.item()) <= 1e-3 * erro.item()
    if torch.equal(lv, lvo):
        np.testing.assert_allclose(rel_add.cpu().numpy(), rel_addo.numpy(), rtol=2e-3)


def test_auto_solver_picks_from_measured_time(qb):
    from quantized_neural_nets_b200 import _lib, step_algorithm as sa
    W, X, Xq = gc._problem(seed=80, N=256, d=64, m=20000, relu=True, xq_noise=0.02)
    W, X, Xq = W.to(DEV), X.to(DEV), Xq.to(DEV)
    sa._AUTO_CHOICE.clear()
    Qa, e2a, r2a = sa.quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / 8, 8, 1, None, 0.1, 1, False, DEV, solver=sa.AUTO,
                                          return_partials=True)
    assert len(sa._AUTO_CHOICE) == 1 and sa.AUTO_LOG, "the eligible shape must have been timed"
    key, times, agree, chosen = sa.AUTO_LOG[-1]
    assert set(times) >= {_lib.SOLVER_DIRECT, _lib.SOLVER_GRAM, _lib.SOLVER_GRAM_F64} and chosen in times
    Qd, e2d, r2d = sa.quantize_layer_impl(W, X, Xq, X.shape[0], 1.16 / 8, 8, 1, None, 0.1, 1, False, DEV,
                                          solver=_lib.SOLVER_DIRECT, return_partials=True)
    assert (Qa == Qd).float().mean().item() >= 0.999
    assert abs(e2a.sum().item() - e2d.sum().item()) <= 2e-3 * e2d.sum().item()
    assert abs(r2a.sum().item() - r2d.sum().item()) <= 1e-4 * r2d.sum().item()


def test_gram_matrices_tcgen05_accuracy(qb):
    """The tensor-core Gram kernel (TMA + tcgen05.mma kind::tf32 x3 + TMEM ping-pong) against float64."""
    from quantized_neural_nets_b200._lib import lib, launch
    d, m = 200, 5000      # not multiples of the 128 x 128 x 32 tile
    g = torch.Generator(device=DEV).manual_seed(0)
    ld = (m + 3) // 4 * 4
    X = torch.relu(torch.randn(d, ld, device=DEV, generator=g)); X[:, m:] = 0
    Xq = torch.relu(X + 0.02 * torch.randn(d, ld, device=DEV, generator=g)); Xq[:, m:] = 0
    ldg = (d + 63) // 64 * 64
    for solver, tol in ((1, 2e-6), (2, 1e-13)):
        nbytes = lib.gpfq_workspace_bytes(solver, 1, d, m)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=DEV)
        out = [torch.zeros((ldg, ldg), dtype=torch.float64, device=DEV) for _ in range(3)]
        launch(lib.gpfq_gram_f32, solver, X, Xq, ld, d, m, out[0], out[1], out[2], ws, nbytes)   # straight through the C ABI
        Xd, Xqd = X[:, :m].double(), Xq[:, :m].double()
        for got, want in zip(out, (Xd @ Xqd.T, Xqd @ Xqd.T, Xd @ Xd.T)):
            rel = ((got[:d, :d] - want).norm() / want.norm()).item()
            assert rel < tol, (solver, rel)


@pytest.mark.parametrize("env", [{"GPFQ_RESIDENT": "1"},
                                 {"GPFQ_RESIDENT": "0"},
                                 {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": "1", "GPFQ_RESIDENT_TN": "32"},
                                 {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": "2", "GPFQ_RESIDENT_TN": "32"},
                                 {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": "4", "GPFQ_RESIDENT_TN": "16"},
                                 {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": "8", "GPFQ_RESIDENT_TN": "32"},
                                 {"GPFQ_RESIDENT": "1", "GPFQ_RESIDENT_CLUSTER": "16", "GPFQ_RESIDENT_TN": "16"}],
                         ids=["resident", "multi_launch", "resident_c1_t32_lane_neuron",
                              "resident_c2_t32", "resident_c4_t16", "resident_c8_t32", "resident_c16_t16"])
def test_direct_solver_variants_in_subprocess(env):
    """The launch structures of the direct solver (multi-launch sweep+recur, single-launch resident with 1-16 CTA
    clusters) are selected per layer by a heuristic; force each one over the same small
    problems (oracle-checked inside tools/sanitize_smoke.py).  The switches are read once per process."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_smoke.py")], env={**os.environ, **env},
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "FAIL" not in out.stdout


def test_stochastic_map_distribution(qb):
    """SGPFQ map (reference step_algorithm.py:7-35): every output is one of the two neighbouring grid points,
    the mean over draws is the input (unbiased) inside the alphabet, and clipping holds."""
    delta, K = torch.tensor(0.25), 4
    x0 = torch.tensor([0.10, -0.10, 0.30, 0.70, -0.62, 0.0, 0.25, 3.0, -3.0], device=DEV)
    torch.manual_seed(0)
    draws = torch.stack([qb.StepAlgorithm._stochastic_msq(delta.to(DEV), x0.clone(), K, 0.0) for _ in range(4000)])
    lo = torch.floor(x0 / 0.25) * 0.25
    inside = x0.abs() < 1.0
    for j in range(x0.numel()):
        vals = set(draws[:, j].tolist())
        if inside[j]:
            assert vals <= {float(lo[j]), float(lo[j]) + 0.25}, (j, vals)
            sigma = 0.25 * 0.5 / (4000 ** 0.5)
            assert abs(draws[:, j].mean().item() - x0[j].item()) < 5 * sigma
        else:
            assert vals == {float(torch.sign(x0[j])) * 1.0}, (j, vals)      # clipped to +-delta*K
    # in place, like the reference
    y = x0.clone()
    out = qb.StepAlgorithm._stochastic_msq(delta.to(DEV), y, K, 0.0)
    assert out.data_ptr() == y.data_ptr()


def test_stochastic_solver_statistics_and_invariance(qb):
    """SGPFQ through the solver: reproducible for a fixed torch seed, independent of neuron slicing and of the
    solver structure, and statistically equivalent to the oracle's stochastic run (its RNG differs)."""
    from quantized_neural_nets_b200 import _lib
    from quantized_neural_nets_b200.step_algorithm import quantize_layer_impl
    W, X, Xq = gc._problem(seed=81, N=120, d=90, m=1500, relu=True, xq_noise=0.02)
    Wd, Xd, Xqd = W.to(DEV), X.to(DEV), Xq.to(DEV)

    def run(solver, rng=(0, 120), seed=1234):
        Q, e2, r2 = quantize_layer_impl(Wd, Xd, Xqd, 1500, 1.16 / 8, 8, 1, None, 0.1, 1, True, DEV, solver=solver,
                                        neuron_range=rng, return_partials=True, seed=seed)
        return Q, float((e2.sum() / r2.sum()).sqrt())

    Q0, rel0 = run(_lib.SOLVER_DIRECT)
    Q1, _ = run(_lib.SOLVER_DIRECT)
    assert torch.equal(Q0, Q1)
    Qa, _ = run(_lib.SOLVER_DIRECT, (0, 50))
    Qb, _ = run(_lib.SOLVER_DIRECT, (50, 120))
    assert torch.equal(Qa[:50], Q0[:50]) and torch.equal(Qb[50:], Q0[50:])
    Qg, relg = run(_lib.SOLVER_GRAM_F64)
    assert (Qg == Q0).float().mean().item() >= 0.999
    Q2, _ = run(_lib.SOLVER_DIRECT, seed=99)
    assert not torch.equal(Q2, Q0)
    torch.manual_seed(3)
    _, _, relo, _, _ = orc.quantize_layer(W, X, Xq, 1500, 1.16 / 8, 8, 1, None, 0.1, 1, True)
    assert abs(rel0 - float(relo)) <= 0.15 * float(relo), (rel0, float(relo))


@pytest.mark.parametrize("reg,lam,K", [(None, 0.0, 8), ("L1", 0.004, 8), ("L0", 0.004, 4), (None, 0.0, 2)])
def test_packed_export_is_lossless_and_matches_format(qb, reg, lam, K):
    """gpfq_pack_levels_f32 / gpfq_unpack_levels_f32: the codes equal the numpy restatement of the container format,
    unpacking restores the solver's fp32 weights exactly, off-alphabet input is refused."""
    N, d, m = 37, 45, 96      # 1665 weights: not a multiple of 8
    W, X, Xq = (t.to(DEV) for t in gc._problem(11, N, d, m, relu=True, xq_noise=0.02))
    Q, _, _, _, _ = qb.StepAlgorithm._quantize_layer(W, X, Xq, m, 1.16 / K, K, 1, reg, lam, 1, False, DEV)
    delta = orc.layer_step_size(W.cpu(), 1.16 / K, K, 1, reg, lam)
    packed = qb.pack_layer(Q, delta, K, reg, lam)
    assert packed.bits == orc.packed_bits(K, reg) and packed.nbytes == (N * d + 7) // 8 * packed.bits
    levels = orc.level_index(Q.cpu(), delta, reg, lam).numpy()
    np.testing.assert_array_equal(packed.codes.cpu().numpy(), orc.pack_levels(levels, K, reg))
    back, lv = qb.unpack_layer(packed, want_levels=True)
    assert torch.equal(back, Q)
    np.testing.assert_array_equal(lv.cpu().numpy(), levels)
    bad = Q.clone()
    bad[3, 5] += 0.37 * float(delta)
    with pytest.raises(ValueError, match="not on the alphabet"):
        qb.pack_layer(bad, delta, K, reg, lam)


def test_export_packed_network_round_trip(qb):
    """Quantize the BN network after the reference's conv/BN fusion pre-pass, export it packed, and rebuild a
    bit-identical network from the codes."""
    import copy
    torch.backends.cudnn.allow_tf32 = False
    model = gc.bn_cnn(0).to(DEV)
    qb.fusion_layers_inplace(model, DEV)
    np.random.seed(5)
    qnn = qb.QuantizeNeuralNet(model, "bn", 6, gc.image_batches(4, 6, 8, 51), 4, 3, [], 1.16, 1.16, 1, 1, None, 0.1,
                               0.5, False, DEV)
    qmodel = qnn.quantize_network()
    packed = qb.export_packed(qnn)
    assert sorted(packed) == [0, 1, 2, 3]
    assert [p.bits for p in packed.values()] == [4, 4, 4, 5]        # 3-bit convs: 9 levels; 4-bit Linear: 17 levels
    rebuilt = copy.deepcopy(model)
    layers = []
    qb.extract_layers(rebuilt, layers)
    qb.load_packed(layers, packed)
    probe = gc.image_batches(1, 5, 8, 52)[0][0].to(DEV)
    with torch.no_grad():
        assert torch.equal(rebuilt(probe), qmodel(probe))
    assert sum(p.nbytes for p in packed.values()) * 6 < sum(l.weight.numel() * 4 for l in layers)


def _grouped_problem(seed, groups, per, dg, m):
    """W (groups*per x dg) and (m x groups*dg) layer inputs with correlated, post-ReLU columns inside each group."""
    g = torch.Generator().manual_seed(seed)
    W = torch.randn(groups * per, dg, generator=g) * 0.1
    X = torch.relu(torch.randn(m, groups, dg, generator=g) + 0.5 * torch.randn(m, groups, 1, generator=g)).reshape(m, -1)
    Xq = torch.relu(X + 0.03 * torch.randn(m, groups * dg, generator=g))
    return W, X.contiguous(), Xq.contiguous()


@pytest.mark.parametrize("groups,per,dg,m,reg,lam", [(24, 1, 9, 777, None, 0.0), (24, 1, 9, 777, "L1", 0.004),
                                                      (24, 1, 25, 301, "L0", 0.004), (4, 3, 18, 1030, None, 0.0),
                                                      (130, 2, 32, 150, None, 0.0)])
def test_grouped_batched_solver_vs_oracle(qb, groups, per, dg, m, reg, lam):
    """gpfq_solve_grouped_f32 (all groups of a depthwise / grouped conv in one batched solve) against the oracle's
    loop over groups (step_algorithm.py:221-247): >= 99.9 % identical levels, group-averaged errors within 1e-3."""
    from quantized_neural_nets_b200 import step_algorithm as sa
    K = 8
    W, X, Xq = _grouped_problem(21, groups, per, dg, m)
    Qo, erro, relo, _, _ = orc.quantize_layer(W, X, Xq, m, 1.16 / K, K, 1, reg, lam, groups, False)
    Q, err, rel, adder, rel_adder = sa.quantize_layer_impl(W.to(DEV), X.to(DEV), Xq.to(DEV), m, 1.16 / K, K, 1, reg, lam,
                                                           groups, False, DEV, solver=sa.GROUPED)
    assert adder is None and rel_adder is None
    delta = orc.layer_step_size(W, 1.16 / K, K, 1, reg, lam)
    agree = (orc.level_index(Q.cpu(), delta, reg, lam) == orc.level_index(Qo, delta, reg, lam)).float().mean().item()
    assert agree >= 0.999, agree
    assert abs(float(err) - float(erro)) <= 1e-3 * float(erro)
    assert abs(float(rel) - float(relo)) <= 1e-3 * float(relo)
    # the loop over groups with the direct solver (the default) stays available and agrees as well
    Q2, err2, rel2, _, _ = sa.quantize_layer_impl(W.to(DEV), X.to(DEV), Xq.to(DEV), m, 1.16 / K, K, 1, reg, lam, groups,
                                                  False, DEV)
    assert (Q2 == Q).float().mean().item() >= 0.999 and abs(float(rel2) - float(rel)) <= 1e-3 * float(rel)


def test_grouped_batched_solver_slices_and_auto(qb):
    from quantized_neural_nets_b200 import step_algorithm as sa
    K, groups, per, dg, m = 8, 96, 2, 9, 2000
    W, X, Xq = (t.to(DEV) for t in _grouped_problem(22, groups, per, dg, m))
    full, e_full, r_full = sa.quantize_layer_impl(W, X, Xq, m, 1.16 / K, K, 1, None, 0.0, groups, False, DEV,
                                                  solver=sa.GROUPED, return_partials=True)
    Q = torch.zeros_like(full)
    for n0, n1 in ((0, 64), (64, 66), (66, 192)):      # whole groups per slice: bit-identical rows
        part, e2, r2 = sa.quantize_layer_impl(W, X, Xq, m, 1.16 / K, K, 1, None, 0.0, groups, False, DEV,
                                              solver=sa.GROUPED, return_partials=True, neuron_range=(n0, n1))
        assert torch.equal(part[n0:n1], full[n0:n1]) and torch.equal(e2[n0:n1], e_full[n0:n1])
        assert torch.equal(r2[n0:n1], r_full[n0:n1])
        assert not part[:n0].any() and not part[n1:].any()
    with pytest.raises(ValueError, match="whole groups"):
        sa.quantize_layer_impl(W, X, Xq, m, 1.16 / K, K, 1, None, 0.0, groups, False, DEV, solver=sa.GROUPED,
                               neuron_range=(1, 64))
    # 'auto' times the batched solve against the loop over 96 groups and keeps the faster one behind the level gate
    sa.AUTO_LOG.clear()
    Qa, _, _, _, _ = sa.quantize_layer_impl(W, X, Xq, m, 1.16 / K, K, 1, None, 0.0, groups, False, DEV, solver=sa.AUTO)
    (key, times, agree, chosen), = [rec for rec in sa.AUTO_LOG if rec[0][0] == "96g"]
    assert chosen == sa.GROUPED and agree >= 0.999 and times[sa.GROUPED] < times["groups_loop"], (times, agree)
    assert torch.equal(Qa, full)
