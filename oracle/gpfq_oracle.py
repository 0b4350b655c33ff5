"""CPU oracle for the GPFQ per-layer quantization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``quantized_neural_nets_b200/`` may import this
module; it is used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs as the checker / CPU baseline, never as the
thing shipped.

What it is: a plain-PyTorch (CPU, fp32) restatement of the reference algorithm of
YixuanSeanZhou/Quantized_Neural_Nets for the hot path

    QuantizeNeuralNet.quantize_network      src/quantize_neural_net.py:117-214
      -> _populate_linear_layer_input       src/quantize_neural_net.py:217-274
           SaveInputMLP / SaveInputConv2d   src/quantize_neural_net.py:277-350
      -> StepAlgorithm._quantize_layer      src/step_algorithm.py:151-249
           StepAlgorithm._quantization      src/step_algorithm.py:107-148
           quantizers                       src/step_algorithm.py:7-104

The reference's arithmetic *is* ATen; the oracle therefore issues the same ATen ops in the
same order (elementwise mul / add_ / sub_, ``mv``, ``linalg.vector_norm`` then square, true
division, floor), so on one machine it is bit-identical to the reference.

Parity pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against *outputs of the reference itself* executed in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference/src`` unmodified, runs it on seeded
inputs and stores the results in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays the oracle against those files bit-exactly.
"""
from __future__ import annotations

import copy
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

# --------------------------------------------------------------------------------------
# alphabet maps (reference: src/step_algorithm.py:38-104; stochastic :7-35)
# --------------------------------------------------------------------------------------


def _levels(x: torch.Tensor, delta, K: int) -> torch.Tensor:
    """|floor(x/delta + 1/2)| clipped at K -- the level count used by all three maps
    (step_algorithm.py:56,80,104).  True division; ties go toward +inf."""
    cap = torch.ones_like(x) * K
    return torch.minimum(torch.abs(torch.floor(x / delta + 0.5)), cap)


def msq(x: torch.Tensor, delta, K: int, lam: float = 0.0) -> torch.Tensor:
    """Nearest point of delta*{-K..K} (step_algorithm.py:38-56)."""
    return torch.sign(x) * delta * _levels(x, delta, K)


def _shrink(x: torch.Tensor, lam: float) -> torch.Tensor:
    """sign(x) * max(|x| - lam, 0)  (step_algorithm.py:79,103)."""
    return torch.sign(x) * torch.maximum(torch.abs(x) - lam, torch.zeros_like(x))


def soft_msq(x: torch.Tensor, delta, K: int, lam: float) -> torch.Tensor:
    """reg='L1': soft-threshold by lam, then msq (step_algorithm.py:84-104)."""
    return msq(_shrink(x, lam), delta, K)


def hard_msq(x: torch.Tensor, delta, K: int, lam: float) -> torch.Tensor:
    """reg='L0': keep |x| > lam, emit sign * (lam + delta*k) with k counted on the shrunk
    value (step_algorithm.py:59-81).  Off-grid alphabet {0, +-(lam + k*delta)}."""
    kept = torch.nn.functional.threshold(torch.abs(x), lam, 0) * torch.sign(x)
    k = _levels(_shrink(kept, lam), delta, K)
    return torch.sign(kept) * (lam + delta * k) * (torch.abs(kept) > lam).float()


def stochastic_msq(x: torch.Tensor, delta, K: int, lam: float = 0.0) -> torch.Tensor:
    """Stochastic rounding between the two neighbouring grid points, then clipping.
    Works IN PLACE on ``x`` like the reference (step_algorithm.py:7-35); consumes the global
    torch RNG through ``torch.bernoulli``."""
    lo = torch.floor(x / delta)
    go_down = torch.bernoulli(1 - x / delta + lo).bool()
    x[go_down] = delta * torch.floor(x[go_down] / delta)
    x[~go_down] = delta * (torch.floor(x[~go_down] / delta) + 1)
    too_big = torch.abs(x) > delta * K
    x[too_big] = torch.sign(x[too_big]) * delta * K
    return x


def pick_quantizer(reg: Optional[str], stochastic: bool) -> Callable:
    """Dispatch of step_algorithm.py:198-208."""
    if reg == "L1":
        return soft_msq
    if reg == "L0":
        return hard_msq
    return stochastic_msq if stochastic else msq


# --------------------------------------------------------------------------------------
# greedy path following (reference: src/step_algorithm.py:107-148)
# --------------------------------------------------------------------------------------


def greedy_path(W, Q, U, X, Xq, quantizer, delta, K, lam, steps: Optional[int] = None) -> None:
    """In place on Q (N x d) and U (N x m).  For feature t:
        U += w_t (x) x_t ; n = ||xq_t||_2 ** 2 ; a = U xq_t / n  (0 when n == 0)
        q_t = quantizer(a) ; U -= q_t (x) xq_t
    ``steps`` limits the loop to the first ``steps`` features (bench sampling only)."""
    d = W.shape[1] if steps is None else min(steps, W.shape[1])
    for t in range(d):
        xt, xqt = X[:, t], Xq[:, t]
        U += torch.outer(W[:, t], xt)                      # mul then add_, both rounded (:141)
        n = torch.linalg.vector_norm(xqt, 2) ** 2          # sqrt, then square (:142)
        if n > 0:
            a = torch.mv(U, xqt) / n                       # :144
        else:
            a = torch.zeros_like(U[:, 0])                  # :146
        Q[:, t] = quantizer(a, delta, K, lam)              # :147
        U -= torch.outer(Q[:, t], xqt)                     # :148


def layer_step_size(W: torch.Tensor, step: float, K: int, pct: float, reg: Optional[str], lam: float):
    """delta = step * mean_i quantile_pct(|W_i|)  (minus lam/K for L0)  (step_algorithm.py:191-192)."""
    rad = torch.quantile(torch.abs(W), pct, axis=1).mean()
    return step * rad - lam / K if reg == "L0" else step * rad


def quantize_layer(W, X, Xq, m, step, K, pct, reg, lam, groups, stochastic, steps=None):
    """Restatement of StepAlgorithm._quantize_layer (step_algorithm.py:151-249).
    Returns (Q, err, rel_err, adder, rel_adder) with the reference's types."""
    delta = layer_step_size(W, step, K, pct, reg, lam)
    N, d = W.shape
    Q = torch.zeros_like(W)
    U = torch.zeros(N, m)
    quantizer = pick_quantizer(reg, stochastic)
    if groups == 1:
        greedy_path(W, Q, U, X, Xq, quantizer, delta, K, lam, steps)
        adder = U.T
        ref = X @ W.T
        rel_adder = torch.linalg.norm(adder, axis=0) / (torch.linalg.norm(ref, axis=0) + 1e-5)
        err = torch.linalg.norm(adder, ord="fro")
        rel_err = err / torch.linalg.norm(ref, ord="fro")
        return Q, err, rel_err, adder, rel_adder
    # grouped convolution: each group is an independent problem; errors are averaged
    # over groups and the adders are dropped (step_algorithm.py:221-247)
    Wg, Qg, Ug = W.view(groups, -1, d), Q.view(groups, -1, d), U.view(groups, -1, m)
    Xg = X.view(X.shape[0], groups, -1)
    Xqg = Xq.view(Xq.shape[0], groups, -1)
    err = 0
    rel_err = 0
    for g in range(groups):
        greedy_path(Wg[g], Qg[g], Ug[g], Xg[:, g, :], Xqg[:, g, :], quantizer, delta, K, lam, steps)
        e = torch.linalg.norm(Ug[g].T, ord="fro")
        err = err + e
        rel_err = rel_err + e / torch.linalg.norm(Xg[:, g, :] @ Wg[g].T, ord="fro")
    return Qg.view(-1, d), err / groups, rel_err / groups, None, None


# --------------------------------------------------------------------------------------
# layer inputs (reference: src/quantize_neural_net.py:277-350)
# --------------------------------------------------------------------------------------


def patch_count(H: int, W: int, kernel, dilation, padding) -> Tuple[int, int]:
    """Patch grid of nn.Unfold with stride == kernel_size (quantize_neural_net.py:320)."""
    kh, kw = kernel
    lh = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // kh + 1
    lw = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // kw + 1
    return lh, lw


def kept_per_image(L: int, p: float) -> int:
    """int(p*L + 1), or p*L (= L) when p == 1  (quantize_neural_net.py:343)."""
    return int(p * L + 1 if p != 1 else p * L)


def draw_patch_indices(B: int, L: int, p: float) -> np.ndarray:
    """Per image, ``kept_per_image`` draws WITH replacement from that image's L patches, via
    the GLOBAL numpy RNG, concatenated over images (quantize_neural_net.py:340-345)."""
    keep = kept_per_image(L, p)
    return np.concatenate([np.random.choice(np.arange(L * i, L * (i + 1)), size=keep) for i in range(B)])


def conv_patches(inp: torch.Tensor, kernel, dilation, padding, idx: np.ndarray) -> torch.Tensor:
    """(B,C,H,W) -> rows ``idx`` of the (B*L, C*kh*kw) non-overlapping patch matrix
    (quantize_neural_net.py:334-347; the conv's own stride is ignored, :320)."""
    cols = nn.Unfold(kernel, dilation, padding, kernel)(inp)      # (B, C*kh*kw, L)
    rows = torch.transpose(cols, 1, 2).reshape(-1, cols.size(1))  # (B*L, C*kh*kw)
    return rows[idx]


class _Stop(Exception):
    """Early exit from a forward pass once the hooked layer was reached (utils.py:24)."""


BLOCK_TYPES: Tuple[type, ...] = ()


def _block_types():
    global BLOCK_TYPES
    if not BLOCK_TYPES:
        from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet
        from torchvision.models.googlenet import BasicConv2d, Inception, InceptionAux
        from torchvision.models.efficientnet import Conv2dNormActivation, SqueezeExcitation, MBConv
        from torchvision.models.mobilenetv2 import InvertedResidual
        BLOCK_TYPES = (nn.Sequential, Bottleneck, BasicBlock, ResNet, BasicConv2d, Inception,
                       InceptionAux, Conv2dNormActivation, SqueezeExcitation, MBConv, InvertedResidual)
    return BLOCK_TYPES


def extract_layers(model: nn.Module, out: List[nn.Module]) -> None:
    """Definition-order walk collecting leaf Linear / Conv2d, recursing only into whitelisted
    container types (exact-type match) -- the layer-index contract of utils.py:76-93."""
    blocks = _block_types()
    for child in model.children():
        if type(child) in blocks:
            extract_layers(child, out)
        if not list(child.children()) and type(child) in (nn.Linear, nn.Conv2d):
            out.append(child)


def capture_layer_inputs(analog: nn.Module, quant: nn.Module, a_layer: nn.Module, q_layer: nn.Module,
                         images: torch.Tensor, retain_rate: float):
    """Inputs of one layer in both networks for one batch (quantize_neural_net.py:217-274).
    Conv: patch rows; the index list is drawn once and reused for the quantized network."""
    grabbed: List[torch.Tensor] = []
    state = {"idx": None}

    def hook(module, args, output):
        if len(args) != 1:
            raise TypeError("The number of input layer is not equal to one!")
        x = args[0]
        if isinstance(module, nn.Conv2d):
            lh, lw = patch_count(x.shape[2], x.shape[3], module.kernel_size, module.dilation, module.padding)
            if state["idx"] is None:
                state["idx"] = draw_patch_indices(x.shape[0], lh * lw, retain_rate)
            x = conv_patches(x, module.kernel_size, module.dilation, module.padding, state["idx"])
        grabbed.append(x)
        raise _Stop

    if type(a_layer) not in (nn.Linear, nn.Conv2d):
        raise TypeError(f"The layer type {type(a_layer)} is not currently supported")
    with torch.no_grad():
        for net, layer in ((analog, a_layer), (quant, q_layer)):
            handle = layer.register_forward_hook(hook)
            try:
                net(images)
            except _Stop:
                pass
            handle.remove()
    return grabbed[0], grabbed[1]


def quantize_network(model: nn.Module, loader: Iterable, *, mlp_bits: int, cnn_bits: int,
                     ignore_layers: Sequence[int] = (), mlp_scalar: float = 1.16, cnn_scalar: float = 1.16,
                     mlp_pct: float = 1.0, cnn_pct: float = 1.0, reg: Optional[str] = None, lam: float = 0.1,
                     retain_rate: float = 0.25, stochastic: bool = False, log: Optional[list] = None) -> nn.Module:
    """Restatement of QuantizeNeuralNet.__init__ + quantize_network
    (quantize_neural_net.py:32-114,117-214): one fresh batch per layer, both networks re-run
    from the image, Q written back into the deep copy."""
    it = iter(loader)
    quant = copy.deepcopy(model)
    a_layers: List[nn.Module] = []
    q_layers: List[nn.Module] = []
    extract_layers(model, a_layers)
    extract_layers(quant, q_layers)
    for i in range(len(q_layers)):
        if i in ignore_layers:
            continue
        images, _ = next(it)
        X, Xq = capture_layer_inputs(model, quant, a_layers[i], q_layers[i], images, retain_rate)
        is_fc = type(a_layers[i]) is nn.Linear
        bits, scalar, pct = (mlp_bits, mlp_scalar, mlp_pct) if is_fc else (cnn_bits, cnn_scalar, cnn_pct)
        K = 2 ** (bits - 1)
        W = a_layers[i].weight.data
        shape = W.shape
        W2 = W if is_fc else W.view(W.size(0), -1)
        groups = 1 if is_fc else a_layers[i].groups
        Q, err, rel, _, _ = quantize_layer(W2, X, Xq, X.shape[0], scalar / K, K, pct, reg, lam, groups, stochastic)
        q_layers[i].weight.data = Q.reshape(shape).float()
        if log is not None:
            log.append((i, float(err), float(rel)))
    return quant


# --------------------------------------------------------------------------------------
# helpers used by tests (not part of the reference)
# --------------------------------------------------------------------------------------


def level_index(Q: torch.Tensor, delta, reg: Optional[str] = None, lam: float = 0.0) -> torch.Tensor:
    """Signed level index of alphabet values: Q/delta for msq/soft, sign*(|Q|-lam)/delta
    (+1 so that the smallest kept magnitude lam is level 1) for the L0 alphabet."""
    Q = Q.double()
    delta = float(delta)
    if reg == "L0":
        mag = torch.where(Q != 0, torch.round((Q.abs() - lam) / delta) + 1, torch.zeros_like(Q))
        return (torch.sign(Q) * mag).to(torch.int32)
    return torch.round(Q / delta).to(torch.int32)


def exact_decision_margin(W, X, Xq, Q, delta, K, reg=None, lam=0.0) -> torch.Tensor:
    """For a GIVEN path Q, the float64 distance of every decision argument a_t/delta + 1/2
    from the nearest integer (N x d).  A level that differs between two correct fp32
    implementations must sit at a rounding tie, i.e. have a margin of a few fp32 ulps."""
    W, X, Xq, Q = W.double(), X.double(), Xq.double(), Q.double()
    N, d = W.shape
    u = torch.zeros(N, X.shape[0], dtype=torch.float64)
    margin = torch.zeros(N, d, dtype=torch.float64)
    for t in range(d):
        u += torch.outer(W[:, t], X[:, t])
        n = (Xq[:, t] ** 2).sum()
        a = u @ Xq[:, t] / n if n > 0 else torch.zeros(N, dtype=torch.float64)
        if reg in ("L1", "L0"):
            a = torch.sign(a) * torch.clamp(a.abs() - lam, min=0)
        z = a / float(delta) + 0.5
        margin[:, t] = (z - torch.round(z)).abs()
        u -= torch.outer(Q[:, t], Xq[:, t])
    return margin


# ----------------------------------------------------------------------------------------------
# Packed low-bit export (no counterpart in the reference, which saves fp32 tensors, main.py:127-131): the
# checker restates the container format documented in include/gpfq_b200.h in numpy.
def packed_bits(K: int, reg: Optional[str] = None) -> int:
    top = K + 1 if reg == "L0" else K
    bits = 1
    while (1 << bits) < 2 * top + 1:
        bits += 1
    return bits


def pack_levels(levels, K: int, reg: Optional[str] = None):
    """Signed level indices (any shape) -> uint8 array: codes level+top, 8 codes little-endian in `bits` bytes."""
    import numpy as np
    top = K + 1 if reg == "L0" else K
    bits = packed_bits(K, reg)
    codes = (np.asarray(levels).astype(np.int64).ravel() + top)
    assert codes.min(initial=0) >= 0 and codes.max(initial=0) < (1 << bits)
    pad = (-len(codes)) % 8
    codes = np.concatenate([codes, np.full(pad, top, dtype=np.int64)]).reshape(-1, 8)
    words = np.zeros(len(codes), dtype=np.uint64)
    for i in range(8):
        words |= codes[:, i].astype(np.uint64) << np.uint64(i * bits)
    out = np.zeros((len(codes), bits), dtype=np.uint8)
    for b in range(bits):
        out[:, b] = ((words >> np.uint64(8 * b)) & np.uint64(0xFF)).astype(np.uint8)
    return out.ravel()
