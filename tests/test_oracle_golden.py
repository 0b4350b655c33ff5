"""The oracle (oracle/gpfq_oracle.py) replayed against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from oracle import gpfq_oracle as orc

ORACLE_QUANTIZER = {"msq": orc.msq, "soft": orc.soft_msq, "hard": orc.hard_msq}


def same_inputs(stored, *tensors):
    if float(stored) != gc.checksum(*tensors):
        pytest.skip("seeded inputs regenerate differently on this machine; golden outputs do not apply")


def test_quantizer_tables_bit_exact(golden):
    g = golden("quantizers.npz")
    for tag, (x, delta, K, lam) in gc.quantizer_inputs().items():
        for mode, fn in ORACLE_QUANTIZER.items():
            got = fn(x.clone(), delta, K, lam).numpy()
            np.testing.assert_array_equal(got, g[f"{tag}_{mode}"], err_msg=f"{tag}/{mode}")


def test_quantizer_known_answers():
    # SURVEY.md section 8c probe of the reference: delta=0.25, K=8, lam=0.1
    d = torch.tensor(0.25)
    x = d * torch.tensor([-2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 7.5, 8.5, 100.0, -100.0, 0.0, -0.0])
    lv = (orc.msq(x, d, 8) / d).tolist()
    assert lv == [-2, -1, -0.0, 1, 2, 3, 8, 8, 8, -8, 0, 0]
    np.testing.assert_allclose(orc.soft_msq(x, d, 8, 0.1).numpy(),
                               [-.5, -.25, -0., 0, .25, .5, 1.75, 2, 2, -2, 0, 0])
    np.testing.assert_allclose(orc.hard_msq(x, d, 8, 0.1).numpy(),
                               [-.6, -.35, -.1, .1, .35, .6, 1.85, 2.1, 2.1, -2.1, 0, 0], rtol=1e-6)


@pytest.mark.parametrize("tag", list(gc.greedy_inputs().keys()))
def test_greedy_path_bit_exact(golden, tag):
    g = golden("greedy_path.npz")
    c = gc.greedy_inputs()[tag]
    same_inputs(g[f"{tag}_in"], c["W"], c["X"], c["Xq"])
    Q = torch.zeros_like(c["W"])
    U = torch.zeros(c["W"].shape[0], c["X"].shape[0])
    orc.greedy_path(c["W"], Q, U, c["X"], c["Xq"], ORACLE_QUANTIZER[c["mode"]], c["delta"], c["K"], c["lam"])
    np.testing.assert_array_equal(Q.numpy(), g[f"{tag}_Q"])
    np.testing.assert_array_equal(U.numpy(), g[f"{tag}_U"])


@pytest.mark.parametrize("tag", list(gc.layer_inputs().keys()))
def test_quantize_layer_bit_exact(golden, tag):
    g = golden("quantize_layer.npz")
    c = gc.layer_inputs()[tag]
    same_inputs(g[f"{tag}_in"], c["W"], c["X"], c["Xq"])
    Q, err, rel, adder, rel_adder = orc.quantize_layer(c["W"], c["X"], c["Xq"], c["X"].shape[0], c["step"], c["K"],
                                                       c["pct"], c["reg"], c["lam"], c["groups"], False)
    np.testing.assert_array_equal(Q.numpy(), g[f"{tag}_Q"])
    np.testing.assert_array_equal(np.asarray(err), g[f"{tag}_err"])
    np.testing.assert_array_equal(np.asarray(rel), g[f"{tag}_rel"])
    if c["groups"] == 1:
        np.testing.assert_array_equal(adder.numpy(), g[f"{tag}_adder"])
        np.testing.assert_array_equal(rel_adder.numpy(), g[f"{tag}_rel_adder"])
    else:
        assert adder is None and rel_adder is None


@pytest.mark.parametrize("tag", list(gc.conv_inputs().keys()))
def test_conv_capture_bit_exact(golden, tag):
    g = golden("conv_capture.npz")
    c = gc.conv_inputs()[tag]
    same_inputs(g[f"{tag}_in"], c["inp_a"], c["inp_q"])
    B, _, H, W = c["inp_a"].shape
    lh, lw = orc.patch_count(H, W, c["kernel"], c["dilation"], c["padding"])
    np.random.seed(c["np_seed"])
    idx = orc.draw_patch_indices(B, lh * lw, c["p"])
    np.testing.assert_array_equal(idx, g[f"{tag}_idx"])
    for which in ("a", "q"):
        rows = orc.conv_patches(c["inp_" + which], c["kernel"], c["dilation"], c["padding"], idx)
        np.testing.assert_array_equal(rows.numpy(), g[f"{tag}_rows_{which}"])


@pytest.mark.parametrize("tag", list(gc.network_inputs().keys()))
def test_tiny_network_bit_exact(golden, tag):
    g = golden("tiny_network.npz")
    c = gc.network_inputs()[tag]
    model = c["model"]
    same_inputs(g[f"{tag}_in"], *[p.data for p in model.parameters()], *[b[0] for b in c["loader"]()])
    np.random.seed(c["np_seed"])
    q = orc.quantize_network(model, c["loader"](), mlp_bits=c["bits"], cnn_bits=c["bits"], ignore_layers=c["ignore"],
                             mlp_scalar=c["scalar"], cnn_scalar=c["scalar"], reg=c["reg"], lam=c["lam"],
                             retain_rate=c["p"])
    layers = []
    orc.extract_layers(q, layers)
    assert len(layers) == 4
    for i, layer in enumerate(layers):
        np.testing.assert_array_equal(layer.weight.data.numpy(), g[f"{tag}_layer{i}"])
    with torch.no_grad():
        np.testing.assert_array_equal(q(c["probe"]).numpy(), g[f"{tag}_logits"])


def test_extract_layers_order_resnet18():
    import torchvision
    m = torchvision.models.resnet18(weights=None)
    layers = []
    orc.extract_layers(m, layers)
    assert len(layers) == 21
    # downsample conv follows conv2 of its block (definition order)
    assert layers[7] is m.layer2[0].downsample[0]
    assert layers[-1] is m.fc


def test_config1_matches_reference(golden):
    """BASELINE.json configs[0] (about 20 s of CPU).  Levels are compared with a tie allowance
    because MKL's sgemv reduction order may differ between the build box and this one."""
    g = golden("config1.npz")
    c = gc.config1_inputs()
    same_inputs(g["inp"], c["W"], c["X"])
    Q, err, rel, _, _ = orc.quantize_layer(c["W"], c["X"], c["X"], 2048, c["step"], c["K"], 1, None, 0.1, 1, False)
    delta = orc.layer_step_size(c["W"], c["step"], c["K"], 1, None, 0.1)
    assert float(delta) == float(g["delta"])
    lv = orc.level_index(Q, delta).numpy()
    agree = (lv == g["levels"]).mean()
    assert agree >= 0.9999, agree
    assert abs(float(rel) - float(g["rel"])) <= 1e-3 * float(g["rel"])
    assert abs(float(rel) - 0.06468) < 5e-5       # value probed in SURVEY.md section 6
    assert lv.min() >= -8 and lv.max() <= 8


def test_stochastic_paths_bit_exact_under_a_fixed_torch_seed(golden):
    g = golden("stochastic.npz")
    x, delta, K, lam = gc.quantizer_inputs()["rand"]
    torch.manual_seed(7)
    np.testing.assert_array_equal(orc.stochastic_msq(x.clone(), delta, K, lam).numpy(), g["map"])
    c = gc.layer_inputs()["l_msq"]
    torch.manual_seed(8)
    Q, err, rel, _, _ = orc.quantize_layer(c["W"], c["X"], c["Xq"], c["X"].shape[0], c["step"], c["K"], c["pct"], None,
                                           c["lam"], 1, True)
    np.testing.assert_array_equal(Q.numpy(), g["layer_Q"])
    np.testing.assert_array_equal(np.asarray(rel), g["layer_rel"])
