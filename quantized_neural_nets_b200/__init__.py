"""quantized_neural_nets_b200 -- B200-native GPFQ per-layer quantization hot path.

Drop-in for the hot path of YixuanSeanZhou/Quantized_Neural_Nets:
    QuantizeNeuralNet(...).quantize_network()  ->  StepAlgorithm._quantize_layer(...)
Host code is Python/PyTorch; all numerics run in libgpfq_b200.so (hand-written sm_100a CUDA,
C ABI in include/gpfq_b200.h).  Importing the package loads the shared library and fails if
it has not been built."""
from . import _lib  # noqa: F401  (loads libgpfq_b200.so; raises ImportError when missing)
from .step_algorithm import StepAlgorithm
from .quantize_neural_net import QuantizeNeuralNet, SaveInputMLP, SaveInputConv2d
from .utils import InterruptException, extract_layers, fusion_layers_inplace, eval_sparsity, test_accuracy
from .export import PackedLayer, pack_layer, unpack_layer, export_packed, load_packed

__all__ = ["StepAlgorithm", "QuantizeNeuralNet", "SaveInputMLP", "SaveInputConv2d", "InterruptException",
           "extract_layers", "fusion_layers_inplace", "eval_sparsity", "test_accuracy",
           "PackedLayer", "pack_layer", "unpack_layer", "export_packed", "load_packed"]
