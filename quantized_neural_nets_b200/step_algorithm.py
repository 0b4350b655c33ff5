"""Host-side mirror of the reference's ``StepAlgorithm`` (src/step_algorithm.py) on top of
libgpfq_b200.so.  Same names, argument order, in-place behaviour and return types as the
reference, so ``from step_algorithm import StepAlgorithm`` can be swapped for
``from quantized_neural_nets_b200.step_algorithm import StepAlgorithm``.

Everything numerically relevant runs in hand-written sm_100a CUDA (csrc/); this file only
shapes arguments.  PyTorch is used for device memory, the layer radius (``torch.quantile``,
once per layer, step_algorithm.py:191) and the plain ``W @ X^T`` library GEMM behind the
relative-error denominators (step_algorithm.py:217,219)."""
import torch

from . import _lib
from ._lib import lib, launch, require_cuda

# solver used by _quantization/_quantize_layer: _lib.SOLVER_DIRECT or _lib.SOLVER_GRAM
DEFAULT_SOLVER = _lib.SOLVER_DIRECT


def _delta_tensor(step_size, device):
    """The reference passes the step size as a 0-dim tensor (step * radius, :192); the unit-test
    entry points may also be handed a Python float."""
    if isinstance(step_size, torch.Tensor):
        return step_size.detach().to(device=device, dtype=torch.float32).reshape(1).contiguous()
    return torch.tensor([float(step_size)], dtype=torch.float32, device=device)


def draw_seed():
    """Seed of one stochastic solve, taken from torch's global CPU generator (so ``torch.manual_seed`` makes SGPFQ
    runs reproducible, and ranks seeded alike draw alike)."""
    return int(torch.randint(0, 2 ** 62, (1,)).item())


def _elementwise(mode, step_size, x, boundary_idx, lamb, seed=0):
    require_cuda(x)
    xc = x.contiguous()
    out = torch.empty_like(xc)
    delta = _delta_tensor(step_size, x.device)
    launch(lib.gpfq_quantize_f32, xc, out, xc.numel(), delta, int(boundary_idx), mode, float(lamb), int(seed))
    return out


def feature_major(X):
    """(m x d) layer input -> (feature-major tensor whose row t is column t of X, leading dim ld).
    Zero-copy when ``X`` already is a transposed view of a 16-byte aligned (d x ld) buffer with
    ld % 4 == 0 (what SaveInputConv2d / SaveInputMLP of this package hand out); otherwise one
    transpose kernel."""
    require_cuda(X)
    m, d = X.shape
    ld = X.stride(1)
    if X.stride(0) == 1 and ld >= m and ld % 4 == 0 and X.data_ptr() % 16 == 0:
        return X.t(), ld
    ld = (m + 3) // 4 * 4
    out = torch.empty((d, ld), dtype=torch.float32, device=X.device)
    Xc = X if X.stride(1) == 1 else X.contiguous()
    launch(lib.gpfq_transpose_f32, Xc, m, d, Xc.stride(0), out, ld)
    return out, ld


def solve_rows(W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, Q, n0, n1, want_err=True, want_residual=False,
               levels=None, solver=None, want_ref=False, seed=0):
    """Run the greedy path for neurons [n0, n1) of W (N x d) against feature-major inputs.
    Writes rows n0..n1-1 of Q.  Returns (row_err2 | None, U (n1-n0, m) | None, row_ref2 | None);
    row_ref2 (||X w_n||^2) is produced by the Gram solvers only."""
    solver = DEFAULT_SOLVER if solver is None else solver
    N, d = W.shape
    rows = n1 - n0
    dev = W.device
    gram = solver in (_lib.SOLVER_GRAM, _lib.SOLVER_GRAM_F64)
    row_err2 = torch.empty(rows, dtype=torch.float64, device=dev) if want_err else None
    row_ref2 = torch.empty(rows, dtype=torch.float64, device=dev) if (want_ref and gram) else None
    U = torch.empty((rows, m), dtype=torch.float32, device=dev) if (want_residual and not gram) else None
    if rows == 0 or d == 0:
        for t in (row_err2, row_ref2):
            if t is not None:
                t.zero_()
        return row_err2, U, row_ref2
    nbytes = lib.gpfq_workspace_bytes(solver, rows, d, m)
    if nbytes == 0:
        raise RuntimeError(f"libgpfq_b200: solver {solver} does not support a (d={d}, m={m}) layer")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    launch(lib.gpfq_solve_f32, solver, W, W.stride(0), Xfm, Xqfm, ldx, N, d, m, n0, n1, delta, int(K), mode, float(lamb),
           int(seed), Q, Q.stride(0), levels, row_err2, row_ref2, U, m, ws, nbytes)
    if want_residual and gram:      # the Gram solvers never form U; rebuild it only when somebody asks for it
        U = torch.matmul(W[n0:n1], Xfm[:, :m]) - torch.matmul(Q[n0:n1], Xqfm[:, :m])
    return row_err2, U, row_ref2


# ---------------------------------------------------------------------------------------------
# per-layer solver choice "from measured time" (BASELINE.json north_star): the first time a LAYER is seen
# with solver='auto' (key: the caller's ``layer_key`` -- the orchestrator passes the identity of the analog layer
# module -- plus (rows, d, m, mode); without a layer_key the key is the shape alone), every eligible solver is
# run once on the layer's real data and timed with CUDA events; a Gram candidate is kept only if it is faster
# AND reproduces at least 99.9 % of the direct solver's levels ON THAT LAYER.  The choice is cached for the
# process, so the gate is measured once per layer (during warm-up), not once per shape.
AUTO = "auto"
GATE = 0.999
_AUTO_CHOICE = {}
AUTO_LOG = []          # (key, {solver: ms}, agreement, chosen) for reports


def min_gated_agreement():
    """Smallest level agreement (vs the direct solver, on the layer's own data) among the layers for which a Gram
    variant was CHOSEN; 1.0 when none was."""
    picked = [ag for (k, tm, ag, ch) in AUTO_LOG if ch not in (_lib.SOLVER_DIRECT, "groups_loop", "gather_inputs")]
    return min(picked) if picked else 1.0


def gram_eligible(rows, d, m):
    """Shapes worth TIMING the Gram form on: many more calibration rows than features and Gram matrices of
    moderate size.  Whether it is used is decided by the measurement (3 d^2 m tensor-core flops + O(N d^2) fp64
    against 5 N d m fp32 instructions plus the direct path's per-block latency), not by this rule."""
    return d <= 2048 and m >= 2 * d


def _auto_pick(W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, n0, n1, candidates, seed=0):
    N, d = W.shape
    times, results = {}, {}
    for sv in candidates:
        Qs = torch.zeros((N, d), dtype=torch.float32, device=W.device)
        for rep in range(2):                                   # first repetition warms caches / attributes
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            solve_rows(W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, Qs, n0, n1, want_err=True, solver=sv, want_ref=True,
                       seed=seed)
            b.record()
        b.synchronize()
        times[sv] = a.elapsed_time(b)
        results[sv] = Qs[n0:n1]
    base = results[_lib.SOLVER_DIRECT]
    best, agree_best = _lib.SOLVER_DIRECT, 1.0
    for sv in candidates:
        if sv == _lib.SOLVER_DIRECT:
            continue
        agree = float((results[sv] == base).float().mean())
        if agree >= GATE and times[sv] < times[best]:
            best, agree_best = sv, agree
    return best, times, agree_best


def _time_call(fn):
    """(milliseconds of the second of two calls, its result) -- the first call warms caches and attributes."""
    for rep in range(2):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
    b.synchronize()
    return a.elapsed_time(b), out


def resolve_solver(solver, W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, n0, n1, seed=0, layer_key=None):
    if solver != AUTO:
        return DEFAULT_SOLVER if solver is None else solver
    rows, d = n1 - n0, W.shape[1]
    key = (rows, d, m, mode) if layer_key is None else (rows, d, m, mode, layer_key)
    if key not in _AUTO_CHOICE:
        candidates = [_lib.SOLVER_DIRECT]
        if gram_eligible(rows, d, m):
            for sv in AUTO_GRAM_CANDIDATES:
                if sv == _lib.SOLVER_GRAM_F64 and float(d) * d * m > 2e10:
                    continue            # fp64 SIMT Gram matrices of this size cannot win; do not spend warm-up on them
                if lib.gpfq_workspace_bytes(sv, rows, d, m) > 0:
                    candidates.append(sv)
        if len(candidates) == 1:
            _AUTO_CHOICE[key] = _lib.SOLVER_DIRECT
        else:
            best, times, agree = _auto_pick(W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, n0, n1, candidates, seed)
            _AUTO_CHOICE[key] = best
            AUTO_LOG.append((key, times, agree, best))
    return _AUTO_CHOICE[key]


AUTO_GRAM_CANDIDATES = [_lib.SOLVER_GRAM, _lib.SOLVER_GRAM_F64]

# grouped / depthwise convolutions: all groups of the layer in one batched solve (gpfq_solve_grouped_f32) instead
# of the reference's loop over groups (step_algorithm.py:221-247)
GROUPED = "grouped"


def solve_grouped(W, Xfm, Xqfm, ldx, m, delta, K, mode, lamb, Q, n0, n1, groups, levels=None, seed=0):
    """Neurons [n0, n1) (whole groups) of a grouped layer: W (N x d_group), X / Xq feature-major
    ((groups * d_group) x ldx).  Writes rows n0..n1-1 of Q; returns (row_err2, row_ref2)."""
    N, dg = W.shape
    rows = n1 - n0
    per = N // groups
    if n0 % per or n1 % per:
        raise ValueError(f"grouped solve: the neuron range [{n0}, {n1}) must cover whole groups of {per} neurons")
    dev = W.device
    row_err2 = torch.zeros(rows, dtype=torch.float64, device=dev)
    row_ref2 = torch.zeros(rows, dtype=torch.float64, device=dev)
    if rows == 0:
        return row_err2, row_ref2
    nbytes = lib.gpfq_grouped_workspace_bytes(rows // per, dg, m)
    if nbytes == 0:
        raise RuntimeError(f"libgpfq_b200: the grouped solver does not support d_group={dg}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    launch(lib.gpfq_solve_grouped_f32, W, W.stride(0), Xfm, Xqfm, ldx, N, dg, m, groups, n0, n1, delta, int(K), mode,
           float(lamb), int(seed), Q, Q.stride(0), levels, row_err2, row_ref2, ws, nbytes)
    return row_err2, row_ref2


def grouped_eligible(groups, dg, m):
    return groups > 1 and lib.gpfq_grouped_workspace_bytes(groups, dg, m) > 0


def gram_reduce_eligible(N, d, m_total):
    """Layers for which, with the calibration rows split over ranks, summing d x d Gram matrices over the ranks
    beats all-gathering the (m x d) inputs: the Gram solver is the faster one on a single GPU for these shapes
    (profiles/r01_bench_auto_solver_n1.json) and the exchanged volume drops from 8*m*d to 24*d*d bytes."""
    return m_total >= 2 * d and (d <= 512 or (d <= 1024 and N >= 2 * d))


def local_gram_matrices(Xfm_local, Xqfm_local, ldx, d, m_local):
    """(3, ldg, ldg) fp64: GT = X Xq^T, H = Xq Xq^T, A = X X^T of THIS rank's calibration rows, formed on the tensor
    cores (gpfq_gram_f32, split-TF32)."""
    ldg = (d + 63) // 64 * 64
    grams = torch.empty((3, ldg, ldg), dtype=torch.float64, device=Xfm_local.device)
    nbytes = lib.gpfq_gram_workspace_bytes(_lib.SOLVER_GRAM, d, m_local)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=Xfm_local.device)
    launch(lib.gpfq_gram_f32, _lib.SOLVER_GRAM, Xfm_local, Xqfm_local, ldx, d, m_local, grams[0], grams[1], grams[2], ws,
           nbytes)
    return grams


def solve_rows_from_grams(W, grams, delta, K, mode, lamb, Q, n0, n1, seed=0, levels=None):
    """The Gram-form recurrence for neurons [n0, n1) from GIVEN (3, ldg, ldg) Gram matrices (gpfq_gram_path_f32).
    Returns (row_err2, row_ref2)."""
    N, d = W.shape
    rows = n1 - n0
    e2 = torch.zeros(rows, dtype=torch.float64, device=W.device)
    r2 = torch.zeros(rows, dtype=torch.float64, device=W.device)
    if rows > 0:
        launch(lib.gpfq_gram_path_f32, W, W.stride(0), N, d, n0, n1, grams[0], grams[1], grams[2], grams.shape[2], delta,
               int(K), mode, float(lamb), int(seed), Q, Q.stride(0), levels, e2, r2)
    return e2, r2


def all_reduce_sum(group):
    """The reducer of a real multi-GPU run: one in-place fp64 all-reduce over ``group``."""
    def reduce(grams):
        import torch.distributed as dist
        dist.all_reduce(grams, group=group)
        return grams
    return reduce


def solve_rows_gram_reduced(W, Xfm_local, Xqfm_local, ldx, m_local, delta, K, mode, lamb, Q, n0, n1, reduce, seed=0):
    """Gram solver over calibration rows that are SPLIT over ranks: every rank forms the Gram matrices of its own
    rows on the tensor cores, ``reduce`` (a callable: local (3, ldg, ldg) fp64 tensor -> the sum over all ranks; in a
    real run ``all_reduce_sum(group)``, in the single-GPU parity test a serial sum over emulated shards) adds them
    up, then the rank runs the recurrence for its neuron slice.  Returns (row_err2, row_ref2) for neurons [n0, n1)."""
    grams = reduce(local_gram_matrices(Xfm_local, Xqfm_local, ldx, W.shape[1], m_local))
    return solve_rows_from_grams(W, grams, delta, K, mode, lamb, Q, n0, n1, seed)


class StepAlgorithm:
    # ------------------------------------------------------------------ alphabet maps
    def _msq(step_size, x, boundary_idx, lamb):
        """Nearest-alphabet map (reference step_algorithm.py:38-56), CUDA elementwise kernel."""
        return _elementwise(_lib.MODE_MSQ, step_size, x, boundary_idx, lamb)

    def _soft_thresholding_msq(step_size, x, boundary_idx, lamb):
        """reg='L1' map (reference step_algorithm.py:84-104)."""
        return _elementwise(_lib.MODE_SOFT, step_size, x, boundary_idx, lamb)

    def _hard_thresholding_msq(step_size, x, boundary_idx, lamb):
        """reg='L0' map (reference step_algorithm.py:59-81)."""
        return _elementwise(_lib.MODE_HARD, step_size, x, boundary_idx, lamb)

    def _stochastic_msq(step_size, x, boundary_idx, lamb):
        """SGPFQ map (reference step_algorithm.py:7-35): stochastic rounding to the two neighbouring grid
        points, then clipping; works IN PLACE on ``x`` like the reference.  The reference's Bernoulli draws
        come from torch's global generator, ours from a Philox stream seeded from it, so results agree in
        distribution (E[q] = x inside the alphabet), not bit for bit."""
        out = _elementwise(_lib.MODE_STOCHASTIC, step_size, x, boundary_idx, lamb, seed=draw_seed())
        x.copy_(out.view_as(x))
        return x

    _MODE_OF = {}

    # ------------------------------------------------------------------ greedy path
    def _quantization(W, Q, U, analog_layer_input, quantized_layer_input, quantizer,
                      step_size, boundary_idx, lamb):
        """In place on Q (N x d) and U (N x m), as the reference (step_algorithm.py:107-148)."""
        mode = StepAlgorithm._MODE_OF.get(quantizer)
        if mode is None:
            raise NotImplementedError(f"quantizer {quantizer} has no CUDA implementation")
        require_cuda(W, Q, U, analog_layer_input, quantized_layer_input)
        N, d = W.shape
        m = analog_layer_input.shape[0]
        if U.shape != (N, m) or Q.shape != (N, d):
            raise ValueError("Q / U shapes do not match W and the layer inputs")
        Wc = W if W.stride(1) == 1 else W.contiguous()
        Xfm, ldx = feature_major(analog_layer_input)
        Xqfm, ldq = feature_major(quantized_layer_input)
        if ldq != ldx:
            Xqfm, ldq = _repack(Xqfm, m, ldx)
        delta = _delta_tensor(step_size, W.device)
        Qc = Q if (Q.stride(1) == 1) else torch.empty((N, d), dtype=torch.float32, device=W.device)
        seed = draw_seed() if mode == _lib.MODE_STOCHASTIC else 0
        _, Ures, _ = solve_rows(Wc, Xfm, Xqfm, ldx, m, delta, boundary_idx, mode, lamb, Qc, 0, N,
                                want_err=False, want_residual=True, seed=seed)
        if Qc is not Q:
            Q.copy_(Qc)
        # the reference accumulates into the U it is handed, which is always zeros (step_algorithm.py:196); the CUDA
        # solver starts from a zero residual, so any other U is rejected instead of being silently overwritten
        if bool(U.any()):
            raise NotImplementedError("_quantization expects the zero-initialised U of the reference's call sites")
        U.copy_(Ures)

    # ------------------------------------------------------------------ one layer
    def _quantize_layer(W, analog_layer_input, quantized_layer_input, m,
                        step_size, boundary_idx, percentile,
                        reg, lamb, groups, stochastic_quantization, device):
        """Drop-in for the reference's per-layer entry (step_algorithm.py:151-249): returns
        (Q, quantize_error, relative_quantize_error, quantize_adder, relative_adder) with errors
        as 0-dim tensors on the device."""
        return quantize_layer_impl(W, analog_layer_input, quantized_layer_input, m, step_size, boundary_idx,
                                   percentile, reg, lamb, groups, stochastic_quantization, device,
                                   want_adder=True)


StepAlgorithm._MODE_OF = {
    StepAlgorithm._msq: _lib.MODE_MSQ,
    StepAlgorithm._soft_thresholding_msq: _lib.MODE_SOFT,
    StepAlgorithm._hard_thresholding_msq: _lib.MODE_HARD,
    StepAlgorithm._stochastic_msq: _lib.MODE_STOCHASTIC,
}


def _repack(Xfm, m, ld):
    out = torch.zeros((Xfm.shape[0], ld), dtype=torch.float32, device=Xfm.device)
    out[:, :m] = Xfm[:, :m]
    return out, ld


def row_radius(W, percentile):
    """quantile_pct(|W_i|) per neuron, on the device (percentile == 1 is the row maximum)."""
    absW = torch.abs(W)
    return absW.amax(dim=1) if percentile == 1 else torch.quantile(absW, percentile, dim=1)


def delta_from_radii(radii_cpu, step_size, boundary_idx, reg, lamb):
    """delta = step * mean_i(radius_i), minus lamb/K for L0 (step_algorithm.py:191-192).

    The mean over the N per-neuron radii is taken by torch ON THE HOST: an fp32 mean is not
    reduction-order independent, a device reduction can land one ulp away from the reference's CPU
    value, and in ill-conditioned layers (m << d) a one-ulp change of the alphabet is amplified by the
    greedy recurrence into visibly different paths.  N floats per layer; computed before the solve."""
    rad = radii_cpu.mean()
    return step_size * rad - lamb / boundary_idx if reg == 'L0' else step_size * rad


def layer_delta(W, step_size, boundary_idx, percentile, reg, lamb):
    """The layer's alphabet step as a 0-dim CPU tensor (synchronises the stream: N floats D2H)."""
    return delta_from_radii(row_radius(W, percentile).cpu(), step_size, boundary_idx, reg, lamb)


def mode_of(reg, stochastic_quantization):
    if reg == 'L1':
        return _lib.MODE_SOFT
    if reg == 'L0':
        return _lib.MODE_HARD
    return _lib.MODE_STOCHASTIC if stochastic_quantization else _lib.MODE_MSQ


def quantize_layer_impl(W, X, Xq, m, step_size, boundary_idx, percentile, reg, lamb, groups,
                        stochastic_quantization, device, want_adder=False, neuron_range=None, levels=None,
                        solver=None, return_partials=False, delta=None, seed=None, rows_split_over=None,
                        layer_key=None):
    """Shared body of ``StepAlgorithm._quantize_layer`` and of the sharded orchestrator.

    neuron_range=(n0, n1) restricts the solve to a contiguous slice of output neurons (rows
    outside it are left zero in Q); with return_partials=True the per-neuron squared norms
    (||u_n||^2, ||X w_n||^2 as float64, full length N, zero outside the slice) are returned instead
    of the reduced errors so that the caller can all-gather them.  ``delta`` (0-dim tensor) skips the
    alphabet computation when the caller has already done it (the orchestrator does it for all layers
    up front, so the per-layer path never synchronises the host)."""
    if torch.device(device).type != 'cuda':
        raise RuntimeError("libgpfq_b200 runs on CUDA devices only; there is no CPU fallback")
    require_cuda(W, X, Xq)
    mode = mode_of(reg, stochastic_quantization)
    if seed is None:    # one seed per layer; every neuron slice / rank must be handed the same one
        seed = draw_seed() if mode == _lib.MODE_STOCHASTIC else 0
    N, d = W.shape
    dev = W.device
    n0, n1 = (0, N) if neuron_range is None else neuron_range
    Wc = W if W.stride(1) == 1 else W.contiguous()
    if delta is None:
        delta = layer_delta(Wc, step_size, boundary_idx, percentile, reg, lamb)
    delta = _delta_tensor(delta, dev)
    Q = torch.zeros((N, d), dtype=torch.float32, device=dev)
    Xfm, ldx = feature_major(X)
    Xqfm, ldq = feature_major(Xq)
    if ldq != ldx:
        Xqfm, ldq = _repack(Xqfm, m, ldx)
    err2 = torch.zeros(N, dtype=torch.float64, device=dev)
    ref2 = torch.zeros(N, dtype=torch.float64, device=dev)
    adder = None
    if rows_split_over is not None:
        # X / Xq hold only this rank's calibration rows (m of them); the ranks of the group exchange Gram matrices
        if groups != 1 or want_adder:
            raise ValueError("rows_split_over supports ungrouped layers without the adder output")
        reducer = rows_split_over if callable(rows_split_over) else all_reduce_sum(rows_split_over)
        e2, r2 = solve_rows_gram_reduced(Wc, Xfm, Xqfm, ldx, m, delta, boundary_idx, mode, lamb, Q, n0, n1, reducer, seed)
        err2[n0:n1], ref2[n0:n1] = e2, r2
        return (Q, err2, ref2) if return_partials else (Q,) + reduce_errors(err2, ref2, groups, None)
    n_per_group = N // groups
    if groups > 1 and solver in (GROUPED, AUTO) and grouped_eligible(groups, d, m) and not want_adder:
        use = solver == GROUPED
        key = ("grouped", groups, N, d, m, mode, n0, n1)
        if solver == AUTO:
            if key not in _AUTO_CHOICE:          # time the batched solve against the loop over groups, once
                t_loop, Q_loop = _time_call(lambda: quantize_layer_impl(
                    W, X, Xq, m, step_size, boundary_idx, percentile, reg, lamb, groups, stochastic_quantization, device,
                    neuron_range=neuron_range, solver=None, return_partials=True, delta=delta, seed=seed)[0])
                Qb = torch.zeros((N, d), dtype=torch.float32, device=dev)
                t_batched, _ = _time_call(lambda: solve_grouped(Wc, Xfm, Xqfm, ldx, m, delta, boundary_idx, mode, lamb,
                                                                 Qb, n0, n1, groups, None, seed))
                agree = float((Qb[n0:n1] == Q_loop[n0:n1]).float().mean())
                _AUTO_CHOICE[key] = agree >= 0.999 and t_batched < t_loop
                AUTO_LOG.append(((f"{groups}g", N, d, m), {"groups_loop": t_loop, GROUPED: t_batched}, agree,
                                 GROUPED if _AUTO_CHOICE[key] else "groups_loop"))
            use = _AUTO_CHOICE[key]
        if use:
            e2, r2 = solve_grouped(Wc, Xfm, Xqfm, ldx, m, delta, boundary_idx, mode, lamb, Q, n0, n1, groups, levels, seed)
            err2[n0:n1], ref2[n0:n1] = e2, r2
            return (Q, err2, ref2) if return_partials else (Q,) + reduce_errors(err2, ref2, groups, None)
    loop_solver = None if solver == GROUPED else solver
    for g in range(groups):
        g0, g1 = max(n0, g * n_per_group), min(n1, (g + 1) * n_per_group)
        if g0 >= g1:
            continue
        Xg = Xfm[g * d:(g + 1) * d]
        Xqg = Xqfm[g * d:(g + 1) * d]
        sv = resolve_solver(loop_solver, Wc, Xg, Xqg, ldx, m, delta, boundary_idx, mode, lamb, g0, g1, seed,
                            layer_key if groups == 1 else None)
        e2, Ures, r2 = solve_rows(Wc, Xg, Xqg, ldx, m, delta, boundary_idx, mode, lamb, Q, g0, g1,
                                  want_err=True, want_residual=(want_adder and groups == 1), levels=levels,
                                  solver=sv, want_ref=True, seed=seed)
        err2[g0:g1] = e2
        if r2 is not None:
            ref2[g0:g1] = r2
        else:
            # ||X w_n||^2 : plain library GEMM (cuBLAS fp32) -- step_algorithm.py:217,219
            Y = torch.matmul(Wc[g0:g1], Xg[:, :m])
            ref2[g0:g1] = torch.linalg.vector_norm(Y, dim=1).double() ** 2
        if Ures is not None:
            adder = Ures.t()
    if return_partials:
        return Q, err2, ref2
    return (Q,) + reduce_errors(err2, ref2, groups, adder)


def reduce_errors(err2, ref2, groups, adder=None):
    """(quantize_error, relative_quantize_error, quantize_adder, relative_adder) from per-neuron
    squared norms, with the reference's conventions (step_algorithm.py:215-219, 239-245)."""
    if groups == 1:
        err = err2.sum().sqrt().float()
        rel = (err2.sum().sqrt() / ref2.sum().sqrt()).float()
        rel_adder = (err2.sqrt().float() / (ref2.sqrt().float() + 1e-5))
        return err, rel, adder, rel_adder
    e = err2.view(groups, -1).sum(dim=1).sqrt()
    r = ref2.view(groups, -1).sum(dim=1).sqrt()
    return (e.sum() / groups).float(), ((e / r).sum() / groups).float(), None, None
