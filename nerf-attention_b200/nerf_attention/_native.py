"""ctypes binding of libnerfattn.so (include/nerfattn.h).

There is deliberately no fallback: if the CUDA library is missing or no B200 is
visible every entry point raises.  PyTorch is used only for device memory and the
current stream handle.
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path

ABI_VERSION = 8

PREC_FP32, PREC_BF16 = 0, 2          # include/nerfattn.h: code 1 (a TF32 variant) is unassigned
PRECISIONS = {'fp32': PREC_FP32, 'bf16': PREC_BF16}
FIT_TARGETS_PRENORMALISED = 1        # na_fit_t.flags

_LIB_PATH = Path(__file__).resolve().parent.parent / 'csrc' / 'libnerfattn.so'
_PROF_LIB_PATH = _LIB_PATH.with_name('libnerfattn_prof.so')
_lib = None
_prof_lib = None

c_void_p, c_int32, c_float, c_double, c_size_t = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_float,
                                                  ctypes.c_double, ctypes.c_size_t)


class NaFit(ctypes.Structure):
    """struct na_fit (include/nerfattn.h)."""
    _fields_ = [
        ('N', c_int32), ('D', c_int32), ('H', c_int32), ('L', c_int32),
        ('omega0', c_float), ('flags', c_int32),
        ('positions', c_void_p), ('targets', c_void_p), ('mean', c_void_p), ('std', c_void_p),
        ('params', c_void_p), ('adam_m', c_void_p), ('adam_v', c_void_p), ('losses', c_void_p),
        ('cos_sims', c_void_p), ('per_pos_mse', c_void_p), ('scalars', c_void_p),
    ]


class NaSynthStream(ctypes.Structure):
    """struct na_synth_stream (include/nerfattn.h)."""
    _fields_ = [('seed', ctypes.c_uint32), ('n_spikes', c_int32), ('max_width', c_int32),
                ('keys', c_void_p), ('values', c_void_p)]


class NativeError(RuntimeError):
    pass


_SIGNATURES = {
    'nerfattn_abi_version': (c_int32, []),
    'nerfattn_last_error': (ctypes.c_char_p, []),
    'nerfattn_param_count': (c_size_t, [c_int32, c_int32, c_int32]),
    'nerfattn_fit_workspace_bytes': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32,
                                               ctypes.POINTER(c_size_t)]),
    'nerfattn_fit_batched': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32,
                                       ctypes.POINTER(c_double), c_double, c_double, c_double,
                                       c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    'nerfattn_fit_batched_ex': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32,
                                          ctypes.POINTER(c_double), c_double, c_double, c_double,
                                          c_int32, c_int32, c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    'nerfattn_fit_launch_count': (ctypes.c_longlong, [ctypes.POINTER(NaFit), c_int32, c_int32, c_int32]),
    'nerfattn_forward_workspace_bytes': (c_int32, [ctypes.POINTER(NaFit), c_int32,
                                                   ctypes.POINTER(c_size_t)]),
    'nerfattn_siren_forward': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32,
                                         ctypes.POINTER(c_void_p), c_void_p, c_size_t, c_void_p]),
    'nerfattn_decode_workspace_bytes': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32,
                                                  ctypes.POINTER(c_size_t)]),
    'nerfattn_decode_qk': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_void_p,
                                     ctypes.POINTER(c_void_p), c_int32, c_int32, c_void_p, c_size_t,
                                     c_void_p]),
    'nerfattn_kvread_qk': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                     c_void_p]),
    'nerfattn_softmax': (c_int32, [c_void_p, c_int32, c_int32, c_float, c_void_p]),
    'nerfattn_pv_workspace_bytes': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_int32, c_int32, c_int32,
                                              ctypes.POINTER(c_size_t)]),
    'nerfattn_kvread_pv': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_size_t,
                                     c_void_p]),
    'nerfattn_decode_pv': (c_int32, [ctypes.POINTER(NaFit), c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                     c_size_t, c_void_p]),
    'nerfattn_synth_workspace_bytes': (c_int32, [c_int32, c_int32, c_int32, ctypes.POINTER(c_size_t)]),
    'nerfattn_synth_kv': (c_int32, [ctypes.POINTER(NaSynthStream), c_int32, c_int32, c_int32, c_void_p, c_void_p,
                                    c_size_t, c_void_p]),
    'nerfattn_debug_gemm_bf16': (c_int32, [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                           c_int32, c_int32, c_int32, c_void_p]),
    'nerfattn_debug_sincos': (c_int32, [c_void_p, c_void_p, c_void_p, ctypes.c_int64, c_int32, c_void_p]),
}
EXPORTS = tuple(_SIGNATURES)


def library_path() -> Path:
    return Path(os.environ.get('NERFATTN_LIB', _LIB_PATH))


def _load(path: Path) -> ctypes.CDLL:
    if not path.exists():
        raise NativeError(
            f'{path} not found: build it with `python __graft_entry__.py` (or `make -C '
            f'{path.parent}`); nerf_attention has no CPU / PyTorch fallback')
    handle = ctypes.CDLL(str(path))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(handle, name)          # AttributeError if the .so is stale
        fn.restype, fn.argtypes = res, args
    got = handle.nerfattn_abi_version()
    if got != ABI_VERSION:
        raise NativeError(f'{path}: ABI version {got}, binding expects {ABI_VERSION}; rebuild')
    return handle


def lib() -> ctypes.CDLL:
    """Load (once) and type the shared library; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        _lib = _load(library_path())
    return _lib


def prof_lib() -> ctypes.CDLL:
    """The -DNA_PROFILING build of the same sources (libnerfattn_prof.so): it honours NERFATTN_PHASE, which makes
    a fit launch only one class of kernels per epoch.  For bench.py / profiles only -- nothing in the package
    calls this, and results of a phase-masked run are meaningless."""
    global _prof_lib
    if _prof_lib is None:
        _prof_lib = _load(Path(os.environ.get('NERFATTN_PROF_LIB', _PROF_LIB_PATH)))   # experiment builds (profiles/)
    return _prof_lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().nerfattn_last_error().decode(errors='replace')
        raise NativeError(f'{what} failed (code {rc}): {msg}')


def require_cuda(device) -> 'torch.device':
    import torch
    dev = torch.device(device)
    if dev.type != 'cuda':
        raise NativeError(f"device {device!r}: this build runs the SIREN path on a B200 only "
                          "(no CPU fallback); pass device='cuda'")
    if not torch.cuda.is_available():
        raise NativeError('CUDA is not available: nerf_attention needs a B200 (sm_100a)')
    return dev


def stream_handle() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def precision_code(precision: str | int | None) -> int:
    if precision is None:
        precision = os.environ.get('NERFATTN_PRECISION', 'fp32')
    if isinstance(precision, int):
        return precision
    try:
        return PRECISIONS[precision.lower()]
    except KeyError:
        raise ValueError(f'precision must be one of {sorted(PRECISIONS)}, got {precision!r}') from None
