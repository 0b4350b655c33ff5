"""Host-side mirror of the reference's orchestrator (src/quantize_neural_net.py):
``QuantizeNeuralNet`` with the same 16-argument constructor and ``quantize_network()``, and the
two forward-hook classes ``SaveInputMLP`` / ``SaveInputConv2d`` with the same constructor
arguments and ``.inputs`` protocol.

Differences from the reference are confined to where the work runs:
  * layer inputs are produced directly in the solver's feature-major layout by the fused
    im2col + patch-gather CUDA kernel (the tensors handed out are (m x d) transposed views);
  * the per-layer solve is libgpfq_b200 (StepAlgorithm mirror in step_algorithm.py);
  * with torch.distributed initialised (one process per GPU), every rank runs the calibration
    forward on the same batches, solves a contiguous slice of each layer's output neurons and
    the slices are exchanged with ONE all-gather per layer.
"""
import copy

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import lib, launch, require_cuda
from .sharding import neuron_slice, gather_layer, gather_inputs, world_and_rank
from . import step_algorithm as _sa
from .step_algorithm import (quantize_layer_impl, reduce_errors, row_radius, delta_from_radii,
                             gram_reduce_eligible, feature_major)
from .utils import InterruptException, extract_layers
from .forward_fusion import share_patches_of, release_shared_patches

LINEAR_MODULE_TYPE = nn.Linear
CONV2D_MODULE_TYPE = nn.Conv2d

# {(id(analog layer), world size, total calibration rows): bool} -- verdict of the per-layer parity gate of the
# Gram-reduce mode of the sharded calibration forward (QuantizeNeuralNet._gate_gram_reduce)
_GRAM_REDUCE_GATE = {}


def _default_group():
    import torch.distributed as dist
    return dist.group.WORLD


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


class SaveInputMLP:
    """Stores the input of a Linear layer (reference quantize_neural_net.py:277-292)."""

    def __init__(self):
        self.inputs = []

    def __call__(self, module, module_in, module_out):
        if len(module_in) != 1:
            raise TypeError('The number of input layer is not equal to one!')
        self.inputs.append(self.capture(module_in[0]))
        raise InterruptException

    def capture(self, x):
        return x


class SaveInputConv2d:
    """Stores the sub-sampled patch matrix of a Conv2d layer's input (reference
    quantize_neural_net.py:295-350).

    Semantics kept from the reference: patches are nn.Unfold patches with stride == kernel_size
    (the layer's own ``stride`` is accepted and ignored, :320); per image ``int(p*L + 1)`` patch
    rows (L when p == 1) are drawn WITH replacement from numpy's global RNG on the first call and
    reused on the second call (:340-347).  The gather itself is one CUDA kernel that writes the
    feature-major matrix; ``inputs[k]`` is its (m x C*kh*kw) transposed view."""

    def __init__(self, kernel_size, dilation, padding, stride, groups, retain_rate, image_range=None,
                 full_batch=None):
        self.p = retain_rate
        # B200 addition: when the calibration forward is sharded over ranks, this rank's forward sees
        # images [i0, i1) of a ``full_batch``-image batch; indices are still drawn for ALL images so
        # that numpy's RNG stream is consumed exactly as in the reference.
        self.image_range = image_range
        self.full_batch = full_batch
        self.kernel_size = _pair(kernel_size)
        self.dilation = _pair(dilation)
        self.padding = _pair(padding)
        self.stride = stride          # unused, as in the reference
        self.groups = groups
        self.inputs = []
        self.call_count = 0
        self.rand_indices = None
        self._idx_dev = None

    def _draw(self, batch_size, num_blocks):
        """Patch rows to keep: per image i, ``keep`` draws with replacement from [L*i, L*(i+1)) off numpy's GLOBAL
        generator (reference :340-345: one ``np.random.choice(np.arange(L*i, L*(i+1)), size=keep)`` per image).
        choice() without weights is ``randint(0, L, size)`` on the same stream, and consecutive randint calls
        with one bound consume the stream exactly like a single larger call, so ONE (batch x keep) draw returns
        the same indices and leaves the generator in the same state (tests/test_host_cpu.py) -- 256 numpy calls
        and 256 aranges per layer less on the host."""
        keep = int(self.p * num_blocks + 1 if self.p != 1 else self.p * num_blocks)
        offsets = np.arange(batch_size, dtype=np.int64) * num_blocks
        return (np.random.randint(0, num_blocks, size=(batch_size, keep)) + offsets[:, None]).reshape(-1)

    def __call__(self, module, module_in, module_out):
        if len(module_in) != 1:
            raise TypeError('The number of input layer is not equal to one!')
        self.inputs.append(self.capture(module_in[0]))
        raise InterruptException

    def capture(self, x):
        """The sub-sampled patch matrix of one (B, C, H, W) layer input as an (m x C*kh*kw) view of a
        feature-major buffer.  The first call draws the patch indices, later calls reuse them."""
        require_cuda(x)
        x = x.contiguous()
        B, C, H, W = x.shape
        (kh, kw), (dh, dw), (ph, pw) = self.kernel_size, self.dilation, self.padding
        Lh = (H + 2 * ph - dh * (kh - 1) - 1) // kh + 1
        Lw = (W + 2 * pw - dw * (kw - 1) - 1) // kw + 1
        if self.call_count == 0:
            L = Lh * Lw
            if self.image_range is None:
                self.rand_indices = self._draw(B, L)
                local = self.rand_indices
            else:
                i0, i1 = self.image_range
                self.rand_indices = self._draw(self.full_batch, L)
                keep = len(self.rand_indices) // self.full_batch
                local = self.rand_indices[i0 * keep:i1 * keep] - i0 * L
            self._idx_dev = torch.from_numpy(np.ascontiguousarray(local, dtype=np.int64)).to(x.device)
        self.call_count += 1
        m = int(self._idx_dev.numel())
        ld = (m + 3) // 4 * 4
        feats = C * kh * kw
        out = torch.empty((feats, ld), dtype=torch.float32, device=x.device)
        launch(lib.gpfq_im2col_gather_f32, x, B, C, H, W, kh, kw, dh, dw, ph, pw, 0, C, self._idx_dev, m, out, ld)
        return out[:, :m].t()


class QuantizeNeuralNet:
    """Drop-in for the reference's QuantizeNeuralNet (quantize_neural_net.py:19-274)."""

    def __init__(self,
                 network_to_quantize, network_name, batch_size, data_loader,
                 mlp_bits, cnn_bits,
                 ignore_layers,
                 mlp_alphabet_scalar, cnn_alphabet_scalar,
                 mlp_percentile, cnn_percentile,
                 reg, lamb, retain_rate, stochastic_quantization, device,
                 *, process_group=None, solver=None, verbose=False, profile=False, shard_forward=False,
                 overlap_solve=False, gram_reduce=True, calibration='fresh', fuse_forward=False, pointwise_gemm=False):
        self.network_name = network_name
        self.analog_network = network_to_quantize          # not copied, as in the reference (:82)
        self.batch_size = batch_size
        self.data_loader_iter = iter(data_loader)

        self.mlp_boundary_idx = 2 ** (mlp_bits - 1)         # alphabet delta*{-K..K}  (:87-88)
        self.cnn_boundary_idx = 2 ** (cnn_bits - 1)
        self.mlp_alphabet_scalar = mlp_alphabet_scalar
        self.mlp_alphabet_step_size = mlp_alphabet_scalar / self.mlp_boundary_idx
        self.cnn_alphabet_step_size = cnn_alphabet_scalar / self.cnn_boundary_idx
        self.mlp_bits = mlp_bits
        self.cnn_bits = cnn_bits
        self.mlp_percentile = mlp_percentile
        self.cnn_percentile = cnn_percentile
        self.ignore_layers = ignore_layers
        self.retain_rate = retain_rate
        self.reg = reg
        self.lamb = lamb
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError("quantized_neural_nets_b200 runs on CUDA devices only; there is no CPU fallback")
        self.stochastic_quantization = stochastic_quantization

        self.quantized_network = copy.deepcopy(self.analog_network)
        self.analog_network_layers = []
        extract_layers(self.analog_network, self.analog_network_layers)
        self.quantized_network_layers = []
        extract_layers(self.quantized_network, self.quantized_network_layers)

        # B200 additions (keyword-only, defaults reproduce the single-GPU reference behaviour)
        self.process_group = process_group   # None = default group when torch.distributed is initialised; False = never shard
        self.shard_forward = shard_forward   # split each calibration batch over the ranks and all-gather X / X~ columns
        self.solver = solver
        # Optional: run layer i's solve on a high-priority side stream while the main stream already runs the
        # analog network for layer i+1 (it does not depend on Q_i); the quantized forward of layer i+1 waits
        # for the solve.  Measured on 1 x B200 (r01): no gain -- the fp32 forward saturates the SMs, so the
        # solver's kernels only time-slice with cuDNN's (3.90 s vs 3.80 s per ResNet-50 step); off by default.
        self.overlap_solve = overlap_solve
        # With shard_forward: for the layers the Gram solver is best at, do not all-gather the layer inputs but
        # all-reduce the d x d Gram matrices of each rank's own calibration rows (24 d^2 instead of 8 m d bytes,
        # and the Gram formation is divided by the world size).
        self.gram_reduce = gram_reduce
        self._rows_split = False
        self._gate_pending = None
        # 'fresh' (reference behaviour, :234): a new loader batch and two forward passes from the image for EVERY
        # layer, O(L^2) layer evaluations.  'reuse' (SURVEY.md section 8f rank 1): ONE batch calibrates all layers --
        # one pass of the analog network records every layer's input, then one pass of the quantized network
        # quantizes each layer when the pass reaches it, O(L) layer evaluations.  'reuse' computes what the
        # reference computes when its loader yields the same batch for every layer.
        if calibration not in ('fresh', 'reuse'):
            raise ValueError(f"calibration must be 'fresh' or 'reuse', not {calibration!r}")
        self.calibration = calibration
        # Run the calibration forward passes through a torch.fx copy of each network in which every 1x1 / strided
        # Conv2d -> BatchNorm2d (-> + residual) (-> ReLU) site is ONE tensor-core kernel, every other inference BatchNorm2d
        # (+ add) (+ ReLU) one elementwise launch and every MaxPool2d one pass (forward_fusion.py).
        # The fused callables share all Conv2d / Linear modules with the networks, so hooks and weight updates behave as
        # before.
        self.fuse_forward = fuse_forward
        # Stride-1 1x1 convolutions that are NOT part of a fused conv + BatchNorm site (fuse_forward) through
        # gpfq_conv1x1_f32 -- the tensor-core kernel without an epilogue (one strided-batched SGEMM on 7 x 7 planes);
        # the Conv2d modules stay in place, only their ``forward`` is overridden on the instance while quantize_network()
        # runs (forward_fusion.pointwise_convs_as_gemm).
        self.pointwise_gemm = pointwise_gemm
        self._fused = {}
        if fuse_forward:
            from .forward_fusion import fuse_inference_forward
            try:
                for net in (self.analog_network, self.quantized_network):
                    self._fused[id(net)] = fuse_inference_forward(net)[0]
            except Exception as exc:      # not traceable (data-dependent control flow): keep the plain modules
                import warnings
                warnings.warn(f"fuse_forward: torch.fx could not trace the network ({exc}); using the unfused forward")
                self._fused = {}
        # host -> device copy of the NEXT layer's batch runs on a copy stream while this layer computes
        self._copy_stream = None
        self._stage_bufs = [None, None]  # persistent device buffers for host batches (see _fetch)
        self._stage_next = 0
        self._prefetched = None      # (device images, ready event, sharded?) of the next layer
        self._layers_left = 0
        self.verbose = verbose
        self.layer_deltas = {}
        self.layer_log = []      # (layer_idx, quantize_error tensor, relative_quantize_error tensor)
        self.profile = profile   # record CUDA-event timings of the phases of every layer
        self._marks = []         # (layer_idx, phase, start_event, end_event)

    # ------------------------------------------------------------------
    class _Phase:
        """CUDA-event bracket around one phase of one layer (only when profile=True)."""

        def __init__(self, owner, layer_idx, name):
            self.owner, self.layer_idx, self.name = owner, layer_idx, name

        def __enter__(self):
            if self.owner.profile:
                self.a = torch.cuda.Event(enable_timing=True)
                self.b = torch.cuda.Event(enable_timing=True)
                self.a.record()
            return self

        def __exit__(self, *exc):
            if self.owner.profile:
                self.b.record()
                self.owner._marks.append((self.layer_idx, self.name, self.a, self.b))
            return False

    def phase_times_ms(self):
        """{phase: total ms} and per-layer list, after a profile=True run."""
        torch.cuda.synchronize()
        totals, per_layer = {}, {}
        for idx, name, a, b in self._marks:
            ms = a.elapsed_time(b)
            totals[name] = totals.get(name, 0.0) + ms
            per_layer.setdefault(idx, {})[name] = per_layer.setdefault(idx, {}).get(name, 0.0) + ms
        return totals, per_layer

    # ------------------------------------------------------------------
    def quantize_network(self):
        """Quantize every non-ignored layer in definition order and return the quantized copy
        (reference quantize_neural_net.py:117-214)."""
        layers_to_quantize = [i for i in range(len(self.quantized_network_layers)) if i not in self.ignore_layers]
        if self.verbose:
            print(f'Layer indices to quantize {layers_to_quantize}')
            print(f'Total number of layers to quantize {len(layers_to_quantize)}')
        deltas = self._layer_deltas(layers_to_quantize)
        self.layer_deltas = deltas             # {layer index: alphabet step}; read by export.export_packed
        # cached TF32 planes of the convolution weights are keyed by (data pointer, version, shape); an in-place write
        # through ``weight.data`` between two calls changes none of them, so every run starts from fresh planes
        for gm in self._fused.values():
            for mod in gm.modules():
                if hasattr(mod, '_planes'):
                    mod._planes = None
        if self.pointwise_gemm:                # 1x1 convolutions of the calibration passes as batched SGEMMs
            from .forward_fusion import pointwise_convs_as_gemm
            with pointwise_convs_as_gemm(self.analog_network, self.quantized_network):
                return self._quantize_layers(layers_to_quantize, deltas)
        return self._quantize_layers(layers_to_quantize, deltas)

    def _quantize_layers(self, layers_to_quantize, deltas):
        if self.calibration == 'reuse':
            self._quantize_network_reuse(layers_to_quantize, deltas)
            return self.quantized_network
        side = torch.cuda.Stream(device=self.device, priority=-1) if self.overlap_solve else None   # high priority
        pending = None
        self._layers_left = len(layers_to_quantize)
        for layer_idx in layers_to_quantize:
            self._layers_left -= 1
            capture = self._begin_capture(layer_idx)          # fresh batch, analog forward        (main stream)
            if pending is not None:
                self._finish_layer(pending)                   # wait for the solve, gather Q, write it back
            analog_layer_input, quantized_layer_input = self._end_capture(capture)   # quantized forward
            pending = self._launch_solve(layer_idx, analog_layer_input, quantized_layer_input, deltas[layer_idx], side)
            del analog_layer_input, quantized_layer_input
        if pending is not None:
            self._finish_layer(pending)
        return self.quantized_network

    def _quantize_network_reuse(self, layers_to_quantize, deltas):
        """calibration='reuse': one batch for all layers, two network passes in total.

        Pass 1 runs the analog network once; a forward pre-hook on every layer to be quantized turns the layer's
        input into its (sub-sampled) X matrix on the spot.  Pass 2 runs the quantized copy once; when it reaches a
        layer, the pre-hook builds X~ from the activation that has just been computed by the already-quantized
        layers before it, solves the layer and installs Q as the module's weight BEFORE the module's own forward
        executes, so the rest of the pass sees the quantized layer.  Both passes stop at the last layer they need.
        With a loader that yields the same batch every time this is the reference's computation
        (quantize_neural_net.py:117-214, :217-274) provided the layers execute in the order extract_layers lists
        them (true of the torchvision models the reference supports); conv patch indices are drawn when the analog
        pass reaches the layer."""
        if not layers_to_quantize:
            return
        self._layers_left = 0                      # a single batch: nothing to prefetch
        images, sharded, shard_range, full_batch = self._next_images()
        savers, analog_X = {}, {}

        def run(network, layers, hook_of, name):
            left = set(layers_to_quantize)
            handles = [layers[i].register_forward_pre_hook(hook_of(i, left)) for i in layers_to_quantize]
            with torch.no_grad(), self._Phase(self, -1, name):
                try:
                    self._fused.get(id(network), network)(images)
                except InterruptException:
                    pass
                finally:
                    for h in handles:
                        h.remove()
            if left:
                raise RuntimeError(f"layers {sorted(left)} were not reached by the calibration forward pass")

        def analog_hook(i, left):
            def hook(module, module_in):
                if len(module_in) != 1:
                    raise TypeError('The number of input layer is not equal to one!')
                savers[i] = self._make_saver(self.analog_network_layers[i], sharded, shard_range, full_batch)
                X = savers[i].capture(module_in[0])
                if isinstance(savers[i], SaveInputMLP):
                    X = feature_major(X)[0][:, :X.shape[0]].t()     # a copy: later in-place ops cannot touch it
                analog_X[i] = X
                left.discard(i)
                if not left:
                    raise InterruptException
            return hook

        def quantized_hook(i, left):
            def hook(module, module_in):
                if len(module_in) != 1:
                    raise TypeError('The number of input layer is not equal to one!')
                Xq = savers.pop(i).capture(module_in[0])
                X, Xq = self._exchange_inputs(i, analog_X.pop(i), Xq, sharded)
                self._finish_layer(self._launch_solve(i, X, Xq, deltas[i], None))
                left.discard(i)
                if not left:
                    raise InterruptException
            return hook

        run(self.analog_network, self.analog_network_layers, analog_hook, 'forward_analog')
        run(self.quantized_network, self.quantized_network_layers, quantized_hook, 'quantized_pass')
        self.layer_log.sort(key=lambda rec: rec[0])

    def _launch_solve(self, layer_idx, X, Xq, delta, side):
        """Enqueue the solve of one layer (on the side stream when overlapping) and return the handle
        ``_finish_layer`` completes."""
        layer = self.analog_network_layers[layer_idx]
        if type(layer) == LINEAR_MODULE_TYPE:
            groups = 1
            W = layer.weight.data
            W_shape = W.shape
        elif type(layer) == CONV2D_MODULE_TYPE:
            groups = layer.groups
            W_shape = layer.weight.data.shape
            W = layer.weight.data.view(W_shape[0], -1)
        else:
            raise TypeError(f'The layer type {type(layer)} is not currently supported')
        step, K, pct = self._layer_params(layer)
        m = X.shape[0]
        n0, n1 = neuron_slice(W.shape[0], groups, self.process_group)
        main = torch.cuda.current_stream(self.device)
        stream = side if side is not None else main
        if side is not None:
            side.wait_stream(main)                 # layer inputs are produced on the main stream
            for t in (X, Xq):
                t.record_stream(side)              # keep their memory alive until the side stream is done
        with torch.cuda.stream(stream):
            with self._Phase(self, layer_idx, 'solve'):
                common = dict(neuron_range=(n0, n1), return_partials=True, delta=delta)
                if self._gate_pending is not None:
                    # first sight of a layer the sharded forward wants to solve from all-reduced Gram matrices: X / Xq
                    # are the all-gathered inputs; solve with the DIRECT solver (the result that is kept), solve again
                    # from the all-reduced Gram matrices of the local rows, and keep the Gram-reduce mode for this layer
                    # only if every rank reproduces >= 99.9 % of the direct levels (one MIN all-reduce: the decision
                    # must be identical on all ranks because it selects which collectives later steps issue)
                    Q, err2, ref2 = quantize_layer_impl(W, X, Xq, m, step, K, pct, self.reg, self.lamb, groups,
                                                        self.stochastic_quantization, self.device,
                                                        solver=_lib.SOLVER_DIRECT, **common)
                    self._gate_gram_reduce(layer, W, Q, n0, n1, step, K, pct, groups, delta)
                else:
                    Q, err2, ref2 = quantize_layer_impl(W, X, Xq, m, step, K, pct, self.reg, self.lamb, groups,
                                                        self.stochastic_quantization, self.device, solver=self.solver,
                                                        rows_split_over=((self.process_group or _default_group())
                                                                         if self._rows_split else None),
                                                        layer_key=id(layer), **common)
            done = torch.cuda.Event()
            done.record(stream)
        alphabet = (_sa._delta_tensor(delta, self.device), K, _sa.mode_of(self.reg, self.stochastic_quantization),
                    float(self.lamb) if self.reg in ('L1', 'L0') else 0.0)
        return layer_idx, Q, err2, ref2, n0, n1, groups, W_shape, done, side, alphabet

    def _finish_layer(self, pending):
        layer_idx, Q, err2, ref2, n0, n1, groups, W_shape, done, side, alphabet = pending
        main = torch.cuda.current_stream(self.device)
        if side is not None:
            main.wait_event(done)
            for t in (Q, err2, ref2):
                t.record_stream(main)
        with self._Phase(self, layer_idx, 'gather'):
            Q, err2, ref2 = gather_layer(Q, err2, ref2, n0, n1, groups, self.process_group, alphabet)
        quantize_error, relative_quantize_error, _, _ = reduce_errors(err2, ref2, groups)
        self.quantized_network_layers[layer_idx].weight.data = Q.reshape(W_shape).float()
        self.layer_log.append((layer_idx, quantize_error, relative_quantize_error))
        if self.verbose:
            print(f'The quantization error of layer {layer_idx} is {quantize_error.cpu().numpy()}.')
            print(f'The relative quantization error of layer {layer_idx} is '
                  f'{relative_quantize_error.cpu().numpy()}.\n')

    # ------------------------------------------------------------------
    def _layer_params(self, layer):
        if type(layer) == LINEAR_MODULE_TYPE:
            return self.mlp_alphabet_step_size, self.mlp_boundary_idx, self.mlp_percentile
        return self.cnn_alphabet_step_size, self.cnn_boundary_idx, self.cnn_percentile

    def _layer_deltas(self, layer_indices):
        """Alphabet step of every layer to be quantized (reference step_algorithm.py:191-192).  They
        depend on the analog weights only, so all per-neuron radii are computed on the device, brought
        to the host in ONE copy, and averaged there (see step_algorithm.delta_from_radii); the per-layer
        loop then never synchronises the host."""
        radii, sizes = [], []
        for i in layer_indices:
            layer = self.analog_network_layers[i]
            if type(layer) not in (LINEAR_MODULE_TYPE, CONV2D_MODULE_TYPE):
                raise TypeError(f'The layer type {type(layer)} is not currently supported')
            W = layer.weight.data.view(layer.weight.shape[0], -1)
            radii.append(row_radius(W, self._layer_params(layer)[2]))
            sizes.append(W.shape[0])
        if not radii:
            return {}
        host = torch.cat(radii).cpu()
        deltas, at = {}, 0
        for i, n in zip(layer_indices, sizes):
            step, K, _ = self._layer_params(self.analog_network_layers[i])
            deltas[i] = delta_from_radii(host[at:at + n].clone(), step, K, self.reg, self.lamb).to(self.device)
            at += n
        return deltas

    # ------------------------------------------------------------------
    def _populate_linear_layer_input(self, layer_idx):
        """Inputs of layer ``layer_idx`` in the analog and in the (partially) quantized network for
        one FRESH batch of the loader (reference quantize_neural_net.py:217-274)."""
        return self._end_capture(self._begin_capture(layer_idx))

    def _run_to_hook(self, name, network, layers, layer_idx, save_input, images):
        # The reference registers SaveInput* as a forward hook (quantize_neural_net.py:256-269): the layer's own forward
        # runs, then the hook stores the INPUT and interrupts the pass -- the output is never used.  The same object is
        # registered here as a forward PRE-hook: identical captured input, but the captured layer's convolution (as much
        # as a tenth of a prefix pass) is not computed just to be thrown away.
        handle = layers[layer_idx].register_forward_pre_hook(lambda module, module_in: save_input(module, module_in, None))
        with torch.no_grad(), self._Phase(self, layer_idx, name):
            try:
                self._fused.get(id(network), network)(images)
            except InterruptException:
                pass
            finally:
                handle.remove()

    def _fetch(self, stream):
        """Draw the next batch from the loader and enqueue its host -> device copy on ``stream``."""
        raw_input_data, _ = next(self.data_loader_iter)
        world, rank = world_and_rank(self.process_group)
        sharded = self.shard_forward and world > 1
        shard_range, B = None, raw_input_data.shape[0]
        if sharded:
            if B % world != 0:
                raise ValueError(f"shard_forward needs the batch ({B}) to be a multiple of the world size ({world})")
            shard_range = (rank * (B // world), (rank + 1) * (B // world))
            raw_input_data = raw_input_data[shard_range[0]:shard_range[1]]
        main = torch.cuda.current_stream(self.device)
        if raw_input_data.is_cuda:
            images = raw_input_data.to(self.device)      # already resident (a no-op on the same device)
            ready = torch.cuda.Event()
            ready.record(main)
            return images, ready, sharded, shard_range, B
        # Host batches go through TWO persistent device buffers used alternately (layer i reads one while layer i + 1's
        # batch lands in the other) instead of a fresh 150 MB allocation per layer: allocations made on the copy stream and
        # handed to the main stream are recycled late by the caching allocator, which kept calling cudaMalloc -- a
        # synchronising call -- in the middle of steps (seen as e2e steps of 1250 ... 1430 ms tracking the number of
        # cudaMalloc calls inside them).
        j = self._stage_next
        self._stage_next ^= 1
        buf = self._stage_bufs[j]
        if buf is None or buf.shape != raw_input_data.shape or buf.dtype != raw_input_data.dtype:
            buf = torch.empty(raw_input_data.shape, dtype=raw_input_data.dtype, device=self.device)
            self._stage_bufs[j] = buf
        # the buffer's previous batch (two layers back) was consumed by work already enqueued on the main stream
        free = torch.cuda.Event()
        free.record(main)
        stream.wait_event(free)
        with torch.cuda.stream(stream):
            buf.copy_(raw_input_data, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(stream)
        return buf, ready, sharded, shard_range, B

    def _next_images(self):
        """Device images of the layer about to be captured; the copy of the FOLLOWING layer's batch (the loader is
        consumed one batch per layer, in order, exactly as the reference does) is started on a copy stream so that
        it overlaps this layer's forward passes and solve."""
        main = torch.cuda.current_stream(self.device)
        if self._prefetched is None:
            images, ready, sharded, shard_range, B = self._fetch(main)
        else:
            images, ready, sharded, shard_range, B = self._prefetched
            self._prefetched = None
            main.wait_event(ready)
        if self._layers_left > 0:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            self._prefetched = self._fetch(self._copy_stream)
        return images, sharded, shard_range, B

    def _make_saver(self, analog_layer, sharded, shard_range, full_batch):
        if type(analog_layer) == LINEAR_MODULE_TYPE:
            save_input = SaveInputMLP()
        elif type(analog_layer) == CONV2D_MODULE_TYPE:
            save_input = SaveInputConv2d(kernel_size=analog_layer.kernel_size, dilation=analog_layer.dilation,
                                         padding=analog_layer.padding, stride=analog_layer.stride,
                                         groups=analog_layer.groups, retain_rate=self.retain_rate)
        else:
            raise TypeError(f'The layer type {type(analog_layer)} is not currently supported')
        if sharded and isinstance(save_input, SaveInputConv2d):
            save_input.image_range, save_input.full_batch = shard_range, full_batch
        return save_input

    def _begin_capture(self, layer_idx):
        """Draw the layer's batch, copy it to the device and run the ANALOG network up to the layer."""
        images, sharded, shard_range, full_batch = self._next_images()
        save_input = self._make_saver(self.analog_network_layers[layer_idx], sharded, shard_range, full_batch)
        share_patches_of(images)              # the quantized pass reads the same batch: one stem patch matrix for both
        self._run_to_hook('forward_analog', self.analog_network, self.analog_network_layers, layer_idx, save_input, images)
        return layer_idx, save_input, images, sharded

    def _end_capture(self, capture):
        """Run the (partially) QUANTIZED network up to the layer on the same batch; returns (X, X~)."""
        layer_idx, save_input, images, sharded = capture
        try:
            self._run_to_hook('forward_quantized', self.quantized_network, self.quantized_network_layers, layer_idx,
                              save_input, images)
        finally:
            release_shared_patches()
        return self._exchange_inputs(layer_idx, save_input.inputs[0], save_input.inputs[1], sharded)

    def _exchange_inputs(self, layer_idx, X, Xq, sharded):
        """With a sharded calibration forward X / X~ hold this rank's calibration rows: either all-gather them,
        or leave them split and let the solve all-reduce the layer's Gram matrices instead."""
        self._rows_split = False
        self._gate_pending = None
        if sharded:
            layer = self.analog_network_layers[layer_idx]
            world, _ = world_and_rank(self.process_group)
            N = layer.weight.shape[0]
            d = layer.weight[0].numel()
            m_total = X.shape[0] * world
            if self.gram_reduce and getattr(layer, 'groups', 1) == 1 and gram_reduce_eligible(N, d, m_total) \
                    and not self.stochastic_quantization:
                verdict = _GRAM_REDUCE_GATE.get((id(layer), world, m_total))
                if verdict is True:
                    self._rows_split = True      # the solve exchanges Gram matrices instead (see _launch_solve)
                    return X, Xq
                if verdict is None:              # not yet verified on this layer: see _launch_solve
                    self._gate_pending = (X, Xq, (id(layer), world, m_total))
            with self._Phase(self, layer_idx, 'gather_inputs'):
                return gather_inputs(X, Xq, self.process_group)
        return X, Xq

    def _gate_gram_reduce(self, layer, W, Q_direct, n0, n1, step, K, pct, groups, delta):
        """Per-LAYER parity gate of the Gram-reduce mode (run once per layer, during warm-up)."""
        import torch.distributed as dist
        X_local, Xq_local, key = self._gate_pending
        self._gate_pending = None
        group = self.process_group or _default_group()
        Qg, _, _ = quantize_layer_impl(W, X_local, Xq_local, X_local.shape[0], step, K, pct, self.reg, self.lamb, groups,
                                       self.stochastic_quantization, self.device, neuron_range=(n0, n1),
                                       return_partials=True, delta=delta, rows_split_over=group)
        same = (Qg[n0:n1] == Q_direct[n0:n1]).float().mean() if n1 > n0 else torch.ones((), device=self.device)
        agree = same.reshape(1).clone()
        dist.all_reduce(agree, op=dist.ReduceOp.MIN, group=group)
        agree = float(agree.item())
        _GRAM_REDUCE_GATE[key] = agree >= _sa.GATE
        _sa.AUTO_LOG.append((("gram_reduce", W.shape[0], W.shape[1], key[2]), {}, agree,
                             "gram_reduce" if _GRAM_REDUCE_GATE[key] else "gather_inputs"))
