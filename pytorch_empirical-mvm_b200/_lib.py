"""ctypes binding of libvsw_b200.so (the C ABI declared in include/vsw.h).

There is NO CPU fallback and no alternative backend: if the shared library cannot be loaded (and
cannot be built with nvcc), importing the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvsw_b200.so")

VSW_F32, VSW_BF16, VSW_F16 = 0, 1, 2
EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_GELU_GRAD = 0, 1, 2, 3
GEMM_AUTO, GEMM_SIMT, GEMM_TCGEN05 = 0, 1, 2

_DTYPES = {torch.float32: VSW_F32, torch.bfloat16: VSW_BF16, torch.float16: VSW_F16}

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); kept in the order of include/vsw.h
SIGNATURES = {
    "vsw_version": (_i, []),
    "vsw_last_error": (_i, [C.c_char_p, _sz]),
    "vsw_set_gemm_backend": (_i, [_i]),
    "vsw_get_gemm_backend": (_i, []),
    "vsw_launch_count": (C.c_longlong, []),
    "vsw_window_maps": (_i, [_i] * 9 + [_vp, _vp, _vp]),
    "vsw_rel_pos_index": (_i, [_i, _i, _i, _vp, _vp]),
    "vsw_shift_mask": (_i, [_vp, _i, _i, _vp, _i, _vp]),
    "vsw_merge_map": (_i, [_i, _i, _i, _vp, _vp]),
    "vsw_ln_fwd": (_i, [_vp] * 7 + [_i] * 4 + [_f, _i, _i, _vp]),
    "vsw_ln_bwd_workspace": (_sz, [_i]),
    "vsw_ln_bwd": (_i, [_vp] * 10 + [_i] * 6 + [_vp, _sz, _vp]),
    "vsw_ln_bwd_ex": (_i, [_vp] * 10 + [_i] * 7 + [_vp, _sz, _vp]),
    "vsw_drop_path_scale": (_i, [_vp, _f, _vp, _i, _i, _vp]),
    "vsw_residual_add": (_i, [_vp] * 5 + [_i] * 6 + [_vp]),
    "vsw_merge_ln_fwd": (_i, [_vp] * 7 + [_i] * 4 + [_f, _i, _vp]),
    "vsw_merge_ln_bwd": (_i, [_vp] * 9 + [_i] * 5 + [_vp, _sz, _vp]),
    "vsw_linear_fwd": (_i, [_vp] * 4 + [_i] * 4 + [_vp] * 4 + [_i, _i, _i, _vp]),
    "vsw_linear_dgrad": (_i, [_vp] * 3 + [_i] * 3 + [_vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "vsw_linear_dgrad_mul": (_i, [_vp] * 3 + [_i] * 3 + [_vp, _vp, _i, _i, _vp, _vp, _i, _vp]),
    "vsw_linear_wgrad_workspace": (_sz, [_i, _i, _i]),
    "vsw_linear_wgrad": (_i, [_vp] * 4 + [_i] * 5 + [_vp, _sz, _vp]),
    "vsw_window_attn_fwd": (_i, [_vp] * 8 + [_i] * 6 + [_f, _i, _i, _vp]),
    "vsw_window_attn_bwd_workspace": (_sz, [_i] * 5),
    "vsw_window_attn_bwd": (_i, [_vp] * 11 + [_i] * 6 + [_f, _i, _i, _vp, _sz, _vp]),
    "vsw_patch_im2col": (_i, [_vp, _vp] + [_i] * 10 + [_vp]),
    "vsw_patch_col2im": (_i, [_vp, _vp] + [_i] * 10 + [_vp]),
    "vsw_enc_video_tail_fwd": (_i, [_vp] * 13 + [_i] * 6 + [_f, _i, _i, _vp]),
    "vsw_enc_video_tail_bwd_workspace": (_sz, [_i] * 4),
    "vsw_enc_video_tail_bwd": (_i, [_vp] * 17 + [_i] * 8 + [_vp, _sz, _vp]),
    "vsw_block_mask_apply": (_i, [_vp] * 4 + [_i] * 6 + [_vp]),
    "vsw_masked_l1_workspace": (_sz, []),
    "vsw_masked_l1_fwd": (_i, [_vp] * 5 + [C.c_longlong, _i, _f, _i, _i, _vp, _sz, _vp]),
    "vsw_masked_l1_bwd": (_i, [_vp] * 6 + [C.c_longlong, _i, _f, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            # the built .so normally travels with the tree; building needs nvcc only (no GPU).  Under torchrun every rank gets
            # here at once: an inter-process file lock lets one of them build while the others wait, then find the library.
            import fcntl
            import importlib.util
            with open(os.path.join(_HERE, ".build.lock"), "w") as lockf:
                fcntl.flock(lockf, fcntl.LOCK_EX)
                try:
                    if not os.path.exists(LIB_PATH):
                        spec = importlib.util.spec_from_file_location("_vsw_build", os.path.join(_HERE, "build.py"))
                        mod = importlib.util.module_from_spec(spec)
                        spec.loader.exec_module(mod)
                        mod.build()
                finally:
                    fcntl.flock(lockf, fcntl.LOCK_UN)
        lib = C.CDLL(LIB_PATH)  # raises OSError loudly if missing / unloadable
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def lib():
    return _load()


class VswError(RuntimeError):
    pass


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().vsw_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise VswError(f"{what or 'vsw'} failed with status {rc}: {last_error()}")


def dt(t_or_dtype) -> int:
    d = t_or_dtype.dtype if isinstance(t_or_dtype, torch.Tensor) else t_or_dtype
    try:
        return _DTYPES[d]
    except KeyError:
        raise VswError(f"unsupported dtype {d}") from None


def ptr(t):
    """raw device pointer (or NULL); tensors must be contiguous CUDA tensors"""
    if t is None:
        return None
    if not t.is_cuda:
        raise VswError("vsw kernels are CUDA-only: got a CPU tensor (there is no CPU fallback)")
    if not t.is_contiguous():
        raise VswError("vsw kernels need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def set_gemm_backend(b: int) -> int:
    return lib().vsw_set_gemm_backend(int(b))


def get_gemm_backend() -> int:
    return lib().vsw_get_gemm_backend()


def launch_count() -> int:
    return int(lib().vsw_launch_count())
