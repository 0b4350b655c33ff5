"""Generate the golden fixtures in this directory by EXECUTING THE REFERENCE ITSELF.

Run only in the build container (needs /root/reference, read-only, imported unmodified):

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference has no tests or golden vectors of its own, so these
files are what pins the oracle (oracle/gpfq_oracle.py) and, through it, the CUDA path.
Every array stored here is an output of the reference's own functions:
  StepAlgorithm._msq/_soft_thresholding_msq/_hard_thresholding_msq   src/step_algorithm.py:38-104
  StepAlgorithm._quantization                                        src/step_algorithm.py:107-148
  StepAlgorithm._quantize_layer                                      src/step_algorithm.py:151-249
  SaveInputConv2d.__call__                                           src/quantize_neural_net.py:325-350
  QuantizeNeuralNet.quantize_network                                 src/quantize_neural_net.py:117-214
Nothing in tests/, bench.py or smoke() reads /root/reference at run time.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(HERE))

from step_algorithm import StepAlgorithm as SA                      # noqa: E402  (the reference)
from quantize_neural_net import QuantizeNeuralNet, SaveInputConv2d  # noqa: E402
from utils import InterruptException                                # noqa: E402

import golden_cases as gc                                           # noqa: E402  (shared seeded input builders)


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        yield


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(f"wrote {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in out.items()))


def quantizer_tables():
    out = {}
    for tag, (x, delta, K, lam) in gc.quantizer_inputs().items():
        out[f"{tag}_msq"] = SA._msq(delta, x.clone(), K, lam)
        out[f"{tag}_soft"] = SA._soft_thresholding_msq(delta, x.clone(), K, lam)
        out[f"{tag}_hard"] = SA._hard_thresholding_msq(delta, x.clone(), K, lam)
    save("quantizers.npz", **out)


REF_QUANTIZER = {"msq": SA._msq, "soft": SA._soft_thresholding_msq, "hard": SA._hard_thresholding_msq}


def greedy_cases():
    out = {}
    for tag, c in gc.greedy_inputs().items():
        Q = torch.zeros_like(c["W"])
        U = torch.zeros(c["W"].shape[0], c["X"].shape[0])
        with quiet():
            SA._quantization(c["W"], Q, U, c["X"], c["Xq"], REF_QUANTIZER[c["mode"]], c["delta"], c["K"], c["lam"])
        out[f"{tag}_Q"] = Q
        out[f"{tag}_U"] = U
        out[f"{tag}_in"] = gc.checksum(c["W"], c["X"], c["Xq"])
    save("greedy_path.npz", **out)


def layer_cases():
    out = {}
    for tag, c in gc.layer_inputs().items():
        with quiet():
            Q, err, rel, adder, rel_adder = SA._quantize_layer(
                c["W"], c["X"], c["Xq"], c["X"].shape[0], c["step"], c["K"], c["pct"], c["reg"], c["lam"],
                c["groups"], False, torch.device("cpu"))
        out[f"{tag}_Q"] = Q
        out[f"{tag}_err"] = err
        out[f"{tag}_in"] = gc.checksum(c["W"], c["X"], c["Xq"])
        out[f"{tag}_rel"] = rel
        if adder is not None:
            out[f"{tag}_adder"] = adder
            out[f"{tag}_rel_adder"] = rel_adder
    save("quantize_layer.npz", **out)


def stochastic_cases():
    """SGPFQ (step_algorithm.py:7-35, 198-208): consumes torch's global generator, so the fixtures fix its seed."""
    out = {}
    x, delta, K, lam = gc.quantizer_inputs()["rand"]
    torch.manual_seed(7)
    out["map"] = SA._stochastic_msq(delta, x.clone(), K, lam)
    c = gc.layer_inputs()["l_msq"]
    torch.manual_seed(8)
    with quiet():
        Q, err, rel, _, _ = SA._quantize_layer(c["W"], c["X"], c["Xq"], c["X"].shape[0], c["step"], c["K"], c["pct"], None,
                                               c["lam"], 1, True, torch.device("cpu"))
    out["layer_Q"], out["layer_err"], out["layer_rel"] = Q, err, rel
    save("stochastic.npz", **out)


def conv_capture_cases():
    out = {}
    for tag, c in gc.conv_inputs().items():
        hook = SaveInputConv2d(c["kernel"], c["dilation"], c["padding"], c["stride"], c["groups"], c["p"])
        np.random.seed(c["np_seed"])
        for which in ("a", "q"):
            try:
                hook(None, (c["inp_" + which],), None)
            except InterruptException:
                pass
        out[f"{tag}_idx"] = hook.rand_indices
        out[f"{tag}_in"] = gc.checksum(c["inp_a"], c["inp_q"])
        out[f"{tag}_rows_a"] = hook.inputs[0]
        out[f"{tag}_rows_q"] = hook.inputs[1]
    save("conv_capture.npz", **out)


def tiny_network():
    out = {}
    for tag, c in gc.network_inputs().items():
        model = c["model"]
        out[f"{tag}_in"] = gc.checksum(*[p.data for p in model.parameters()], *[b[0] for b in c["loader"]()])
        np.random.seed(c["np_seed"])
        with quiet():
            qnn = QuantizeNeuralNet(model, "tiny", c["batch"], c["loader"](), c["bits"], c["bits"], c["ignore"],
                                    c["scalar"], c["scalar"], 1, 1, c["reg"], c["lam"], c["p"], False,
                                    torch.device("cpu"))
            qmodel = qnn.quantize_network()
        for i, layer in enumerate(qnn.quantized_network_layers):
            out[f"{tag}_layer{i}"] = layer.weight.data
        with torch.no_grad():
            out[f"{tag}_logits"] = qmodel(c["probe"])
    save("tiny_network.npz", **out)


def model_utils():
    """utils.fusion_layers_inplace (utils.py:96-130), eval_sparsity (:133-159), test_accuracy (:54-73)."""
    import utils as ref_utils
    out = {}
    net = gc.bn_cnn(0)
    probe = gc.image_batches(1, 4, 8, 62)[0][0]
    with torch.no_grad():
        out["logits_before"] = net(probe)
    out["sparsity"] = ref_utils.eval_sparsity(net)
    with quiet():
        out["topk"] = ref_utils.test_accuracy(net, gc.labelled_loader(), torch.device("cpu"), topk=(1, 3))
    ref_utils.fusion_layers_inplace(net, torch.device("cpu"))
    for name, t in net.state_dict().items():
        out["fused_" + name.replace(".", "_")] = t
    out["fused_eps"] = np.array([m.eps for m in net if isinstance(m, nn.BatchNorm2d)])
    # the reference leaves eps = 0 (utils.py:121), which torch >= 2.x rejects in F.batch_norm; the smallest normal
    # fp32 is arithmetically the same (1 + tiny == 1), so the fused network is evaluated with it
    for mod in net:
        if isinstance(mod, nn.BatchNorm2d):
            mod.eps = float(torch.finfo(torch.float32).tiny)
    with torch.no_grad():
        out["logits_after"] = net(probe)
    save("model_utils.npz", **out)


def config1():
    c = gc.config1_inputs()
    with quiet():
        Q, err, rel, _, _ = SA._quantize_layer(c["W"], c["X"], c["X"], c["X"].shape[0], c["step"], c["K"], 1, None, 0.1,
                                               1, False, torch.device("cpu"))
    delta = c["step"] * torch.quantile(c["W"].abs(), 1, axis=1).mean()
    levels = torch.round(Q / delta).to(torch.int8)
    assert torch.equal(levels.float() * delta, Q), "cfg1 Q is not exactly on the delta grid"
    save("config1.npz", levels=levels, err=err, rel=rel, delta=delta, inp=gc.checksum(c["W"], c["X"]))


if __name__ == "__main__":
    torch.set_num_threads(8)
    quantizer_tables()
    greedy_cases()
    layer_cases()
    stochastic_cases()
    conv_capture_cases()
    tiny_network()
    model_utils()
    if "--skip-cfg1" not in sys.argv:
        config1()
