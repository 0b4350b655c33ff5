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
