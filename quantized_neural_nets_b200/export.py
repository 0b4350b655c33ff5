"""Packed low-bit export of a quantized network.

The reference keeps quantized weights as fp32 tensors whose values lie on the layer's alphabet
(quantize_neural_net.py:163,193) and saves the whole module with torch.save (main.py:127-131).  Here every
quantized layer can instead be stored as ceil(log2(#alphabet values))-bit codes plus its alphabet step: 5 bits per
weight for the 17-value "4-bit" alphabet delta*{-8..8} (step_algorithm.py:56).  Packing and unpacking run in
libgpfq_b200 (gpfq_pack_levels_f32 / gpfq_unpack_levels_f32); unpacking reproduces the fp32 weights exactly."""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from ._lib import lib, launch, require_cuda
from .step_algorithm import mode_of


@dataclass
class PackedLayer:
    codes: torch.Tensor           # uint8, ceil(numel / 8) * bits bytes
    shape: Tuple[int, ...]        # the layer's weight shape
    delta: torch.Tensor           # 0-dim fp32 alphabet step
    boundary_idx: int             # K = 2^(bits-1)
    reg: Optional[str]            # None | 'L1' | 'L0'  (which alphabet)
    lamb: float
    bits: int                     # code width

    @property
    def numel(self):
        n = 1
        for s in self.shape:
            n *= s
        return n

    @property
    def nbytes(self):
        return self.codes.numel()


def pack_layer(Q, delta, boundary_idx, reg=None, lamb=0.0, stochastic_quantization=False):
    """fp32 alphabet-valued weights -> PackedLayer.  Raises if any entry of ``Q`` is not exactly on the alphabet
    delta*{-K..K} (or the L0 alphabet), so a successful export is lossless by construction."""
    require_cuda(Q)
    mode = mode_of(reg, stochastic_quantization)
    lam = float(lamb) if reg in ('L1', 'L0') else 0.0
    Qc = Q.contiguous()
    n = Qc.numel()
    bits = int(lib.gpfq_packed_bits(int(boundary_idx), mode))
    if bits <= 0:
        raise ValueError(f"unsupported alphabet: K={boundary_idx}")
    delta_dev = torch.as_tensor(delta, dtype=torch.float32).reshape(()).to(Q.device)
    codes = torch.empty(((n + 7) // 8) * bits, dtype=torch.uint8, device=Q.device)
    bad = torch.zeros(1, dtype=torch.int32, device=Q.device)
    launch(lib.gpfq_pack_levels_f32, Qc, n, delta_dev, int(boundary_idx), mode, lam, codes, bad)
    n_bad = int(bad.item())
    if n_bad:
        raise ValueError(f"{n_bad} of {n} weights are not on the alphabet (delta={float(delta_dev):.6g}, "
                         f"K={boundary_idx}, reg={reg}); the layer is not a GPFQ-quantized layer with these parameters")
    return PackedLayer(codes, tuple(Q.shape), delta_dev, int(boundary_idx), reg, lam, bits)


def unpack_layer(packed, want_levels=False):
    """PackedLayer -> fp32 weights of the original shape (and the int8 signed level indices if asked)."""
    dev = packed.codes.device
    if dev.type != 'cuda':
        raise RuntimeError("libgpfq_b200 runs on CUDA devices only; there is no CPU fallback")
    mode = mode_of(packed.reg, False)
    n = packed.numel
    Q = torch.empty(packed.shape, dtype=torch.float32, device=dev)
    levels = torch.empty(packed.shape, dtype=torch.int8, device=dev) if want_levels else None
    launch(lib.gpfq_unpack_levels_f32, packed.codes, n, packed.delta.to(dev), packed.boundary_idx, mode, packed.lamb, Q,
           levels)
    return (Q, levels) if want_levels else Q


def export_packed(quantizer):
    """{layer index: PackedLayer} for every layer ``quantizer.quantize_network()`` has quantized."""
    if not getattr(quantizer, 'layer_deltas', None):
        raise RuntimeError("export_packed: run quantize_network() first")
    out = {}
    for idx, delta in quantizer.layer_deltas.items():
        layer = quantizer.quantized_network_layers[idx]
        _, K, _ = quantizer._layer_params(layer)
        out[idx] = pack_layer(layer.weight.data, delta, K, quantizer.reg, quantizer.lamb,
                              quantizer.stochastic_quantization)
    return out


def load_packed(network_layers, packed_layers):
    """Install unpacked weights into the layers of a network (the inverse of export_packed)."""
    for idx, packed in packed_layers.items():
        network_layers[idx].weight.data = unpack_layer(packed)
