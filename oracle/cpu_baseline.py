"""CPU arm of bench.py: the reference's own implementation of the hot path timed on the host cores.

TEST / BASELINE INFRASTRUCTURE ONLY -- imported by bench.py's ``--impl reference`` arm and ``cpu_baseline`` leg, never
by the product (quantized_neural_nets_b200/ must not import anything under oracle/).

What is timed (reference src/main.py:120-122 times ``QuantizeNeuralNet(...).quantize_network()``):
  * kind "reference": the UNMODIFIED reference imported from oracle/_ref (oracle/make_ref.py vendors the three source
    files there, untracked) -- ``StepAlgorithm._quantization`` for the greedy loop, torchvision's forward for the
    calibration passes, and ``QuantizeNeuralNet.quantize_network`` in full for the validation run;
  * kind "port": the oracle restatement (gpfq_oracle.py), only when oracle/_ref is absent.

A full ResNet-50 bs=256 run takes several minutes per step on host cores, and the driver asks for K steps, so one
"step" here is a BOUNDED SAMPLE of that workload:
  1. greedy loop: one layer shape per class of residual-matrix size N*m (the loop's cost per feature is five passes over
     the N x m matrix U, whatever d is, step_algorithm.py:141-148), the first k features of it timed on
     column-strided views of full-width W / X (so memory access is that of the full problem), extrapolated to every
     layer by N*d*m through the nearest class in log(N*m);
  2. calibration forward: ONE full fp32 forward at the bench's batch size, scaled by the reference's schedule (both
     networks re-run from the image up to layer i for every layer i, quantize_neural_net.py:256-269), i.e. by
     2 * sum_i flops(layers before i) / flops(network).
``validate()`` checks this extrapolation against ONE complete, real ``quantize_network()`` run of a model that finishes
in about a minute (AlexNet at the bench's batch size), in the same process.
"""
import contextlib
import os
import time

import numpy as np
import torch

from . import gpfq_oracle as orc
from . import make_ref

UNIT = "weights*samples/s"


def host_threads():
    """Use every host core this process may run on (torchrun exports OMP_NUM_THREADS=1 for N > 1)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def reference_modules():
    """(StepAlgorithm class, QuantizeNeuralNet class, extract_layers fn, kind)"""
    mods = make_ref.import_reference()
    if mods is not None:
        sa, qnn, ut = mods
        return sa.StepAlgorithm, qnn.QuantizeNeuralNet, ut.extract_layers, "reference"
    return None, None, orc.extract_layers, "port"


def build_model(name):
    import torchvision
    torch.manual_seed(0)
    return getattr(torchvision.models, name)(weights=None).eval()


def layer_shapes(model, batch, retain, extract_layers, image=224):
    """(N, d, m, groups) of every quantizable layer, in the reference's layer order (utils.py:76-93)."""
    layers = []
    extract_layers(model, layers)
    spatial = {}

    def hook(mod, args, out):
        spatial[mod] = tuple(args[0].shape)

    handles = [l.register_forward_hook(hook) for l in layers]
    with torch.no_grad():
        model(torch.zeros(1, 3, image, image, device=next(model.parameters()).device))
    for h in handles:
        h.remove()
    shapes = []
    for l in layers:
        if isinstance(l, torch.nn.Linear):
            shapes.append((l.out_features, l.in_features, batch, 1))
        else:
            _, C, H, W = spatial[l]
            kh, kw = l.kernel_size
            lh, lw = orc.patch_count(H, W, l.kernel_size, l.dilation, l.padding)
            keep = orc.kept_per_image(lh * lw, retain)
            shapes.append((l.out_channels, C // l.groups * kh * kw, batch * keep, l.groups))
    return shapes


def pick_classes(shapes, max_classes=9):
    """Representative (N, d, m) shapes: the layer list is cut into ``max_classes`` bins of log(N*m) and the shape
    carrying the most units in each bin represents it."""
    keyed = {}
    for (N, d, m, g) in shapes:
        keyed[(N, d, m)] = keyed.get((N, d, m), 0.0) + float(N) * d * m
    logs = sorted(np.log(N * m) for (N, d, m) in keyed)
    lo, hi = logs[0], logs[-1] + 1e-9
    bins = {}
    for (N, d, m), units in keyed.items():
        b = min(max_classes - 1, int((np.log(N * m) - lo) / (hi - lo) * max_classes))
        if b not in bins or units > bins[b][1]:
            bins[b] = ((N, d, m), units)
    return [bins[b][0] for b in sorted(bins)]


@contextlib.contextmanager
def _quiet():
    """The reference prints per-layer errors to stdout and wraps its feature loop in a tqdm bar on stderr."""
    with open(os.devnull, "w") as null, contextlib.redirect_stdout(null), contextlib.redirect_stderr(null):
        yield


def time_greedy_features(StepAlgorithm, N, d, m, k, gen, bits=4, keep=None):
    """Seconds the reference's loop takes for the first k features of an (N, d, m) layer: W, X are full width and the
    loop runs on their first-k-column views, so columns are read with stride d as in the complete layer.  With a list
    ``keep`` the problem and the reference's decisions are appended to it (W_k, X_k, Q_k, delta, K) so that the caller can
    hand the same inputs to the CUDA solver (GPFQ decides feature t from features <= t only, so the first k columns of Q
    are a complete sub-problem)."""
    K = 2 ** (bits - 1)
    W = torch.randn(N, d, generator=gen) * 0.05
    X = torch.relu(torch.randn(m, d, generator=gen))
    delta = orc.layer_step_size(W, 1.16 / K, K, 1, None, 0.1)
    Q = torch.zeros_like(W)
    U = torch.zeros(N, m)
    Wk, Qk, Xk = W[:, :k], Q[:, :k], X[:, :k]
    with _quiet():
        t0 = time.perf_counter()
        if StepAlgorithm is not None:      # the unmodified reference (step_algorithm.py:107-148)
            StepAlgorithm._quantization(Wk, Qk, U, Xk, Xk, StepAlgorithm._msq, delta, K, 0.1)
        else:
            orc.greedy_path(Wk, Qk, U, Xk, Xk, orc.msq, delta, K, 0.0)
        dt = time.perf_counter() - t0
    if keep is not None:
        keep.append({"shape": (N, d, m), "k": k, "W": Wk.contiguous(), "X": Xk.contiguous(), "Q": Qk.contiguous().clone(),
                     "delta": delta.clone(), "K": K})
    return dt


def sample_greedy(shapes, StepAlgorithm, seconds_per_class, bits=4, max_classes=9, keep=None):
    """-> (extrapolated seconds for all layers, seconds spent, {class: units/s})."""
    gen = torch.Generator().manual_seed(3)
    rates, spent = {}, 0.0
    for (N, d, m) in pick_classes(shapes, max_classes):
        k0 = min(2, d)
        dt0 = time_greedy_features(StepAlgorithm, N, d, m, k0, gen, bits)       # also touches pages / warms caches
        k = int(min(d, max(k0, seconds_per_class / (dt0 / k0))))
        dt = time_greedy_features(StepAlgorithm, N, d, m, k, gen, bits, keep)
        spent += dt0 + dt
        rates[(N, m)] = float(N) * m * k / dt
    total = 0.0
    for (N, d, m, groups) in shapes:
        near = min(rates, key=lambda c: abs(np.log(c[0] * c[1]) - np.log(N * m)))
        total += float(N) * d * m / rates[near]
    return total, spent, rates


def sample_forward(model_name, batch, extract_layers, image=224, sample_batch=None):
    """-> (extrapolated seconds of all calibration forward passes of one quantize_network(), seconds spent,
    seconds of one full forward at ``batch``, full-forward equivalents per network).  With ``sample_batch`` the
    forward is timed on fewer images and scaled linearly (warm-up steps only)."""
    model = build_model(model_name)
    layers = []
    extract_layers(model, layers)
    flops = {}

    def hook(mod, args, out):
        flops[mod] = 2.0 * out[0].numel() * mod.weight[0].numel()

    handles = [l.register_forward_hook(hook) for l in layers]
    nb = min(batch, sample_batch or batch)
    x = torch.randn(nb, 3, image, image, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        model(x[:2])
        for h in handles:
            h.remove()
        t0 = time.perf_counter()
        model(x)
        spent = time.perf_counter() - t0
    full = spent * batch / nb
    total = sum(flops[l] for l in layers)
    before, equiv = 0.0, 0.0
    for l in layers:
        equiv += before / total
        before += flops[l]
    return 2.0 * equiv * full, spent, full, equiv


class _Loader:
    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        i = 0
        while True:
            yield self.batches[i % len(self.batches)], None
            i += 1


def run_full_reference(QuantizeNeuralNet, model_name, batch, retain, bits=4, reg=None, lamb=0.1, pool=2):
    """One COMPLETE quantize_network() of the unmodified reference on the host cores (main.py:105-122 with a synthetic
    loader); returns seconds."""
    model = build_model(model_name)
    g = torch.Generator().manual_seed(1)
    batches = [torch.randn(batch, 3, 224, 224, generator=g) for _ in range(pool)]
    np.random.seed(0)
    with _quiet():
        q = QuantizeNeuralNet(model, model_name, batch, _Loader(batches), bits, bits, [], 1.16, 1.16, 1, 1, reg, lamb,
                              retain, False, torch.device("cpu"))
        t0 = time.perf_counter()
        q.quantize_network()
        return time.perf_counter() - t0


def validate(model_name="alexnet", batch=256, retain=0.25, seconds_per_class=0.3):
    """Measured-vs-extrapolated on a model whose complete reference run fits in about a minute."""
    StepAlgorithm, QuantizeNeuralNet, extract_layers, kind = reference_modules()
    if kind != "reference":
        return {"skipped": "oracle/_ref is absent: no real reference to run in full"}
    shapes = layer_shapes(build_model(model_name), batch, retain, extract_layers)
    solver_s, spent, _ = sample_greedy(shapes, StepAlgorithm, seconds_per_class)
    fwd_s, fwd_spent, full, equiv = sample_forward(model_name, batch, extract_layers)
    measured = run_full_reference(QuantizeNeuralNet, model_name, batch, retain)
    return {"model": model_name, "batch": batch, "what": "one complete quantize_network() of the unmodified reference "
            "(oracle/_ref) on the host cores against the sampled-and-extrapolated figure of the same procedure",
            "measured_s": round(measured, 2), "extrapolated_s": round(solver_s + fwd_s, 2),
            "extrapolated_over_measured": round((solver_s + fwd_s) / measured, 3),
            "extrapolated_solver_s": round(solver_s, 2), "extrapolated_forward_s": round(fwd_s, 2),
            "sampling_s": round(spent + fwd_spent, 2)}


def sample_step(model_name, batch, retain, bits, shapes, seconds_per_class=0.3, forward_batch=64, keep=None):
    """One bounded-sample "step" -> dict(seconds (extrapolated quantize_network() time), spent, solver_s, forward_s,
    full_forward_s, equiv, kind)."""
    StepAlgorithm, _, extract_layers, kind = reference_modules()
    solver_s, spent, rates = sample_greedy(shapes, StepAlgorithm, seconds_per_class, bits, keep=keep)
    fwd_s, fwd_spent, full, equiv = sample_forward(model_name, batch, extract_layers, sample_batch=forward_batch)
    return {"seconds": solver_s + fwd_s, "spent": spent + fwd_spent, "solver_s": solver_s, "forward_s": fwd_s,
            "full_forward_s": full, "equiv": equiv, "kind": kind, "classes": len(rates),
            "forward_batch": min(batch, forward_batch or batch)}


def describe(step, cores, n_layers, batch):
    impl = ("the UNMODIFIED reference (oracle/_ref: StepAlgorithm._quantization + torchvision forward)"
            if step["kind"] == "reference" else "oracle port of the reference (torch CPU, the reference's own ATen ops)")
    return (f"{impl} on {cores} host threads; bounded sample per step: greedy loop timed on the first k features of "
            f"{step['classes']} layer shapes (one per class of N*m) and extrapolated by N*d*m over the {n_layers} layers "
            f"({step['solver_s']:.0f} s), plus the calibration forward passes: one full fp32 forward, timed on "
            f"{step['forward_batch']} images and scaled to bs={batch} ({step['full_forward_s']:.1f} s), x 2 networks x {step['equiv']:.1f} full-forward equivalents of prefix "
            f"passes ({step['forward_s']:.0f} s); {step['spent']:.1f} s of CPU work per step")
