"""ctypes binding of libgpfq_b200.so (C ABI declared in include/gpfq_b200.h).

The CUDA library is the product; there is no CPU or PyTorch fallback.  If the shared object
has not been built (``python -c "import __graft_entry__ as g; g.build()"``) importing this
module raises ImportError, and every compute call raises RuntimeError when no CUDA device is
present."""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpfq_b200.so")

MODE_MSQ, MODE_SOFT, MODE_HARD, MODE_STOCHASTIC = 0, 1, 2, 3
SOLVER_DIRECT, SOLVER_GRAM, SOLVER_GRAM_F64 = 0, 1, 2

c_ptr = ctypes.c_void_p  # device pointers travel as integers
c_i64 = ctypes.c_int64
c_i32 = ctypes.c_int32
c_f32 = ctypes.c_float

# name -> (restype, argtypes); mirrors include/gpfq_b200.h one to one
SIGNATURES = {
    "gpfq_abi_version": (c_i32, []),
    "gpfq_last_error": (ctypes.c_char_p, []),
    "gpfq_launch_count": (c_i64, []),
    "gpfq_profile_begin": (c_i32, []),
    "gpfq_profile_end": (c_i32, [ctypes.POINTER(ctypes.c_double)]),
    "gpfq_profile_kind": (c_i32, [c_i32, ctypes.POINTER(ctypes.c_double)]),
    "gpfq_quantize_f32": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_i32, c_i32, c_f32, ctypes.c_uint64, c_ptr]),
    "gpfq_bn_act_f32": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_f32, c_f32, c_ptr]),
    "gpfq_conv1x1_f32": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr, ctypes.c_size_t, c_ptr]),
    "gpfq_conv1x1_workspace_bytes": (ctypes.c_size_t, [c_i32, c_i32]),
    "gpfq_conv1x1_fused_supported": (c_i32, [c_i32, c_i32, c_i32, c_i64]),
    "gpfq_conv1x1_bn_act_f32": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_f32,
                                        c_f32, c_ptr, ctypes.c_size_t, c_ptr]),
    "gpfq_conv1x1_split_weight_f32": (c_i32, [c_ptr, c_i32, c_i32, c_ptr, ctypes.c_size_t, c_ptr]),
    "gpfq_conv1x1_bn_act_planes_f32": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_f32,
                                               c_f32, c_ptr, ctypes.c_size_t, c_ptr]),
    "gpfq_conv_patches_f32": (c_i32, [c_ptr] + [c_i32] * 12 + [c_ptr, c_i64, c_ptr]),
    "gpfq_maxpool2d_f32": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    "gpfq_packed_bits": (c_i32, [c_i32, c_i32]),
    "gpfq_pack_levels_f32": (c_i32, [c_ptr, c_i64, c_ptr, c_i32, c_i32, c_f32, c_ptr, c_ptr, c_ptr]),
    "gpfq_unpack_levels_f32": (c_i32, [c_ptr, c_i64, c_ptr, c_i32, c_i32, c_f32, c_ptr, c_ptr, c_ptr]),
    "gpfq_slice_row_bytes": (c_i64, [c_i32]),
    "gpfq_pack_slice_f32": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_i32, c_i32, c_ptr, c_i32, c_i32, c_f32, c_ptr, c_ptr, c_ptr,
                                    c_ptr, c_ptr]),
    "gpfq_unpack_slices_f32": (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i32, c_i32, c_f32, c_ptr, c_i64, c_ptr, c_ptr, c_ptr]),
    "gpfq_transpose_f32": (c_i32, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr]),
    "gpfq_im2col_gather_f32": (c_i32, [c_ptr] + [c_i32] * 12 + [c_ptr, c_i64, c_ptr, c_i64, c_ptr]),
    "gpfq_workspace_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32, c_i32]),
    "gpfq_gram_workspace_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    "gpfq_gram_f32": (c_i32, [c_i32, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.c_size_t, c_ptr]),
    "gpfq_gram_path_f32": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i32,
                                   c_i32, c_f32, ctypes.c_uint64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr]),
    "gpfq_grouped_workspace_bytes": (ctypes.c_size_t, [c_i32, c_i32, c_i32]),
    "gpfq_solve_grouped_f32": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr,
                                       c_i32, c_i32, c_f32, ctypes.c_uint64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr,
                                       ctypes.c_size_t, c_ptr]),
    "gpfq_solve_f32": (c_i32, [c_i32, c_ptr, c_i64, c_ptr, c_ptr, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32,
                               c_ptr, c_i32, c_i32, c_f32, ctypes.c_uint64, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_ptr, c_i64,
                               c_ptr, ctypes.c_size_t, c_ptr]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA library first "
            "(python -c \"import __graft_entry__ as g; g.build()\").  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise RuntimeError("libgpfq_b200: " + lib.gpfq_last_error().decode("utf-8", "replace"))


def require_cuda(*tensors):
    if not torch.cuda.is_available():
        raise RuntimeError("libgpfq_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    for t in tensors:
        if t is not None and (not t.is_cuda or t.dtype != torch.float32):
            raise TypeError(f"expected a CUDA float32 tensor, got {t.device} {t.dtype}")


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def launch(fn, *args):
    """Call one stream-ordered entry point of the library.  Tensor arguments are passed as device pointers (None as
    NULL); the device they live on -- all of them must share it -- is made current for the call, and that device's
    current stream is appended as the trailing ``stream`` argument.  The C side (cudaFuncSetAttribute, occupancy
    queries, tensor maps, launches) works on the CURRENT device, so without this a tensor on cuda:1 in a process
    whose current device is cuda:0 would be launched on the wrong GPU; the reference takes a ``device`` argument and
    works on any device without set_device."""
    device = None
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise TypeError(f"libgpfq_b200: expected CUDA tensors, got one on {a.device}")
            if device is None:
                device = a.device
            elif a.device != device:
                raise ValueError(f"libgpfq_b200: tensors on different devices ({device} and {a.device}) in one call")
            conv.append(a.data_ptr())
        elif a is None:
            conv.append(None)
        else:
            conv.append(a)
    if device is None:
        raise ValueError("libgpfq_b200: a launch needs at least one tensor argument")
    # the device guard costs several microseconds of host time per call (the calibration forward makes ~10^4 calls per
    # step and is host-bound at small per-GPU batches): enter it only when the tensors are NOT on the current device
    if device.index == torch.cuda.current_device():
        rc = fn(*conv, torch.cuda.current_stream().cuda_stream)
    else:
        with torch.cuda.device(device):
            rc = fn(*conv, torch.cuda.current_stream(device).cuda_stream)
    if rc != 0:
        check(rc)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def launch_count():
    return int(lib.gpfq_launch_count())


def profile_begin():
    check(lib.gpfq_profile_begin())


def profile_end():
    """-> dict(sweep_launches, sweep_ms, sweep_bytes, sweep_fp32_instr, other_launches, resident_launches,
    resident_ms, resident_fp32_instr, bn_act_launches, bn_act_ms, bn_act_bytes)"""
    out = (ctypes.c_double * 12)()
    check(lib.gpfq_profile_end(out))
    prof = dict(sweep_launches=int(out[0]), sweep_ms=out[1], sweep_bytes=out[2], sweep_fp32_instr=out[3],
                other_launches=int(out[4]), resident_launches=int(out[5]), resident_ms=out[6],
                resident_fp32_instr=out[7], bn_act_launches=int(out[8]), bn_act_ms=out[9], bn_act_bytes=out[10])
    for kind, name in ((0, "sweep"), (3, "conv"), (4, "gram_tc"), (5, "gram_path"), (6, "recur")):
        k = (ctypes.c_double * 5)()
        check(lib.gpfq_profile_kind(kind, k))
        prof[name] = dict(launches=int(k[0]), ms=k[1], bytes=k[2], flops=k[3], aux=k[4])
    return prof
