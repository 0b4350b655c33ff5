"""How far do the REFERENCE's own results move when its inputs move by one ulp?

GPFQ is a chaotic recurrence: a decision that sits at a rounding tie flips under any change of the last bit of its
argument, the flipped level changes the layer's output by a whole alphabet step, and every later layer of a
free-running quantize_network() then sees a different quantized input.  Two correct fp32 implementations -- the
reference on a CPU and the reference on a GPU, say -- therefore agree layer by layer only when they are fed the same
(W, X, X~) (the teacher-forced tests), not when they run free.  This script measures the size of that effect on the
unmodified reference itself, so that the free-running logits tolerance of tests/test_gpu_round2.py is a measured
number and not a guess:

  run A: the reference's QuantizeNeuralNet.quantize_network() on a random-init ResNet-18, batch 8, 4 bits, CPU;
  run B: the same, with every calibration image multiplied by (1 + 2^-23) -- a relative perturbation of one fp32 ulp,
         far below the difference between two convolution algorithms.

Stored in resnet18_sensitivity.json: relative L2 distance of the two quantized networks' logits on a held-out batch,
fraction of bit-identical quantized weights, worst per-layer relative-error deviation, and the distance of either
network's logits from the fp32 network's (the quantization effect itself).

    python tests/golden/make_sensitivity.py          # build container only (imports /root/reference/src unmodified)
"""
import contextlib
import copy
import io
import json
import os
import sys

import numpy as np
import torch
import torchvision

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)

from quantize_neural_net import QuantizeNeuralNet      # noqa: E402  (the reference)
from utils import extract_layers                       # noqa: E402


def run(model, batches, capture):
    np.random.seed(0)
    log = io.StringIO()
    with contextlib.redirect_stdout(log), contextlib.redirect_stderr(io.StringIO()):
        q = QuantizeNeuralNet(copy.deepcopy(model), "resnet18", 8, batches, 4, 4, [], 1.16, 1.16, 1, 1, None, 0.1, 0.25,
                              False, torch.device("cpu")).quantize_network()
    rel = [float(line.split(" is ")[1].rstrip(".")) for line in log.getvalue().splitlines()
           if line.startswith("The relative quantization error")]
    capture.append(rel)
    return q


def main():
    torch.manual_seed(0)
    model = torchvision.models.resnet18(weights=None).eval()
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(8, 3, 224, 224, generator=g), None) for _ in range(21)]
    probe = torch.randn(8, 3, 224, 224, generator=g)
    rels = []
    qa = run(model, batches, rels)
    qb = run(model, [(x * (1.0 + 2.0 ** -23), y) for x, y in batches], rels)
    with torch.no_grad():
        fa, fb, fp = qa(probe), qb(probe), model(probe)
    la, lb = [], []
    extract_layers(qa, la)
    extract_layers(qb, lb)
    tot = sum(l.weight.numel() for l in la)
    same = sum((a.weight.data == b.weight.data).sum().item() for a, b in zip(la, lb)) / tot
    per_layer = [float((a.weight.data == b.weight.data).float().mean()) for a, b in zip(la, lb)]
    out = {
        "what": "unmodified reference, ResNet-18 random init, batch 8, 4 bits, CPU: run A vs run B = images * (1 + 2^-23)",
        "logits_rel_l2_A_vs_B": float((fa - fb).norm() / fa.norm()),
        "logits_rel_l2_A_vs_fp32": float((fa - fp).norm() / fp.norm()),
        "logits_rel_l2_B_vs_fp32": float((fb - fp).norm() / fp.norm()),
        "identical_weights": same,
        "identical_weights_per_layer": per_layer,
        "worst_layer_rel_err_deviation": max(abs(a - b) / a for a, b in zip(*rels)),
        "top1_agreement_A_vs_B": float((fa.argmax(1) == fb.argmax(1)).float().mean()),
        "torch": torch.__version__,
    }
    with open(os.path.join(HERE, "resnet18_sensitivity.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
