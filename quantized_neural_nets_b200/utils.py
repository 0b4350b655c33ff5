"""Host utilities of the hot path: layer discovery and the early-exit exception.

Mirrors the two names quantize_neural_net.py imports from the reference's utils.py
(``InterruptException`` utils.py:24, ``extract_layers`` utils.py:76-93)."""
import torch.nn as nn


class InterruptException(Exception):
    """Raised by the input-capture hooks to abort a forward pass at the hooked layer."""


SUPPORTED_LAYER_TYPE = {nn.Linear, nn.Conv2d}
SUPPORTED_BLOCK_TYPE = None


def _default_block_types():
    from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet
    from torchvision.models.googlenet import BasicConv2d, Inception, InceptionAux
    from torchvision.models.efficientnet import Conv2dNormActivation, SqueezeExcitation, MBConv
    from torchvision.models.mobilenetv2 import InvertedResidual
    return {nn.Sequential, Bottleneck, BasicBlock, ResNet, BasicConv2d, Inception, InceptionAux,
            Conv2dNormActivation, SqueezeExcitation, MBConv, InvertedResidual}


def extract_layers(model, layer_list, supported_block_type=None, supported_layer_type=SUPPORTED_LAYER_TYPE):
    """Append the quantizable leaf layers of ``model`` to ``layer_list`` in definition order.

    Same contract as the reference (utils.py:76-93): containers are entered only when their
    exact type is whitelisted, leaves are taken only when their exact type is Linear/Conv2d.
    The resulting positions are the layer indices used by ``ignore_layers``."""
    global SUPPORTED_BLOCK_TYPE
    if supported_block_type is None:
        if SUPPORTED_BLOCK_TYPE is None:
            SUPPORTED_BLOCK_TYPE = _default_block_types()
        supported_block_type = SUPPORTED_BLOCK_TYPE
    for child in model.children():
        if type(child) in supported_block_type:
            extract_layers(child, layer_list, supported_block_type, supported_layer_type)
        if type(child) in supported_layer_type and next(child.children(), None) is None:
            layer_list.append(child)


def fusion_layers_inplace(model, device):
    """Fold every BatchNorm2d that directly follows a Conv2d (in extract_layers order) into that convolution, in
    place, before quantization -- the reference's ``-f`` pre-pass (utils.py:96-130, main.py:85-87).

    The convolution's weights take the per-channel scale ``gamma / sqrt(var + eps)``; the BN module is left as an
    identity scale that only adds the folded shift (or as a full identity when the convolution has a bias, which
    then receives the shift), so the module graph, and hence the layer indices, do not change.  The arithmetic is
    the reference's, operation for operation.  One deliberate difference: the reference sets ``bn.eps = 0``
    (:121), which current PyTorch rejects in F.batch_norm; the smallest normal fp32 is used instead, which leaves
    ``var + eps == 1`` exactly, so the fused network computes the same values."""
    import torch
    chain = []
    extract_layers(model, chain, supported_layer_type=[nn.Conv2d, nn.BatchNorm2d])
    for conv, bn in zip(chain, chain[1:]):
        if not (isinstance(conv, nn.Conv2d) and isinstance(bn, nn.BatchNorm2d)):
            continue
        std = torch.sqrt(bn.running_var + bn.eps)
        scale = bn.weight.data / std
        shift = bn.bias.data - bn.weight.data * bn.running_mean / std
        conv.weight.data = conv.weight.data * scale[:, None, None, None]
        n = bn.num_features
        bn.running_var = torch.ones(n, device=device)
        bn.running_mean = torch.zeros(n, device=device)
        bn.weight.data = torch.ones(n, device=device)
        bn.eps = float(torch.finfo(torch.float32).tiny)
        if conv.bias is None:
            bn.bias.data = shift
        else:
            conv.bias.data = conv.bias.data * scale + shift
            bn.bias.data = torch.zeros(n, device=device)


def eval_sparsity(model):
    """Fraction of exactly-zero parameters (weights and biases) over the Linear / Conv2d layers of ``model``,
    rounded to 4 decimals (reference utils.py:133-159; printed after sparse-mode quantization, main.py:158)."""
    import numpy as np
    layers = []
    extract_layers(model, layers)
    total = zeros = 0
    for layer in layers:
        for p in (layer.weight, layer.bias):
            if p is not None:
                total += p.numel()
                zeros += int(p.eq(0).sum().item())
    return np.around(zeros / total, 4)


def test_accuracy(model, test_dl, device, topk=(1,)):
    """Top-k accuracies of ``model`` over a loader of (images, labels), as a numpy array aligned with ``topk``
    (reference utils.py:54-73; the denominator is ``len(test_dl.dataset)``)."""
    import numpy as np
    import torch
    model.eval()
    maxk = max(topk)
    hits = np.zeros(len(topk))
    with torch.no_grad():
        for images, target in test_dl:
            pred = torch.topk(model(images.to(device)), maxk, dim=1).indices
            correct = pred.eq(target.to(device).view(-1, 1))
            for i, k in enumerate(topk):
                hits[i] += correct[:, :k].sum().item()
    return hits / len(test_dl.dataset)


test_accuracy.__test__ = False      # not a pytest test
