"""ctypes mirror of libtwkb_tools.so (tomahawk_b200/csrc/tools.cu): measurement tooling of bench.py and the
tests -- a device-side synthetic genotype generator and two peak probes. Nothing here computes LD and the
product library never loads it. torch is used for device memory only.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import VARIANT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
TOOLS_PATH = os.path.join(_HERE, "libtwkb_tools.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(TOOLS_PATH):
            raise ImportError(f"{TOOLS_PATH} is missing (make -C tomahawk_b200/csrc)")
        L = ctypes.CDLL(TOOLS_PATH)
        L.twkb_tools_last_error.restype = ctypes.c_char_p
        L.twkb_tools_synth.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                       ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_double, ctypes.c_uint32,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.twkb_tools_popc_rate.argtypes = [ctypes.POINTER(ctypes.c_double)] * 3
        L.twkb_tools_fp4_gemm.argtypes = [ctypes.c_uint32, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
        _lib = L
    return _lib


def words_per_variant(n_samples: int) -> int:
    w = (2 * n_samples + 63) // 64
    return (w + 1) // 2 * 2


def synth_device(n_samples: int, n_variants: int, seed: int = 1, missing_rate: float = 0.0, rare_fraction: float = 0.0,
                 pos_step: int = 100, p_copy: float = 0.7, redraw: float = 0.05, first: int = 0, n_rows: int | None = None,
                 device=None):
    """Rows [first, first + n_rows) of the synthetic matrix (LD blocks, AF spectrum 0.5 U^3) generated on
    the current CUDA device. Returns (data, mask or None, meta): torch int64 tensors [n_rows, stride] in the
    twk_igt_vec row layout and a numpy VARIANT_DTYPE array. The stream is keyed on the global variant index:
    any slice of the same (seed, n_samples, n_variants) is the same data on every GPU."""
    import torch

    n_rows = n_variants - first if n_rows is None else n_rows
    stride = words_per_variant(n_samples)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    data = torch.empty((n_rows, stride), dtype=torch.int64, device=dev)
    mask = torch.empty((n_rows, stride), dtype=torch.int64, device=dev) if missing_rate > 0 else None
    meta = np.zeros(n_rows, dtype=VARIANT_DTYPE)
    torch.cuda.synchronize()
    rc = lib().twkb_tools_synth(seed, n_samples, n_variants, first, n_rows, p_copy, redraw, rare_fraction, missing_rate, pos_step,
                                data.data_ptr(), mask.data_ptr() if mask is not None else None, stride, meta.ctypes.data)
    if rc != 0:
        raise RuntimeError(lib().twkb_tools_last_error().decode())
    return data, mask, meta


def rows_to_alleles(data: np.ndarray, mask: np.ndarray | None, n_samples: int) -> np.ndarray:
    """Packed rows (uint64 [n, stride]) -> allele codes uint8 [n, 2N] (0 ref, 1 alt, 2 missing), for writing .twk files."""
    H = 2 * n_samples
    bits = np.unpackbits(np.ascontiguousarray(data).view(np.uint8), axis=1, bitorder="little")[:, :H]
    out = bits.astype(np.uint8)
    if mask is not None:
        mb = np.unpackbits(np.ascontiguousarray(mask).view(np.uint8), axis=1, bitorder="little")[:, :H]
        out[mb == 1] = 2
    return out


def popc_rate():
    """(POPC/s, POPC per clock per SM at the nominal clock, nominal SM MHz) of AND+POPC+ADD streams on the current device."""
    a, b, c = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    if lib().twkb_tools_popc_rate(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)) != 0:
        raise RuntimeError(lib().twkb_tools_last_error().decode())
    return a.value, b.value, c.value


def fp4_gemm_tflops(n: int = 8192, sustain_s: float = 2.0):
    """(burst, sustained) TFLOP/s of a cuBLASLt block-scaled e2m1 GEMM n^3, or raises if cuBLASLt has none."""
    a, b = ctypes.c_double(), ctypes.c_double()
    if lib().twkb_tools_fp4_gemm(n, sustain_s, ctypes.byref(a), ctypes.byref(b)) != 0:
        raise RuntimeError(lib().twkb_tools_last_error().decode())
    return a.value, b.value
