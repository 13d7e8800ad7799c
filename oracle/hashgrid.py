"""ctypes front-end for oracle/hashgrid_oracle.c -- TEST INFRASTRUCTURE ONLY.

Presents the same call signature as the reference's pybind module
(`encoder/hashencoder/src/bindings.cpp:5-8`, `hashencoder.h:13-14`) so it can be slotted in as
`encoder.hashencoder.backend._backend` when the *reference's own* Python is imported on CPU
(oracle/make_golden.py), and is what oracle/nsr_oracle.py calls for the encoder.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "hashgrid_oracle.c")
_LIB = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement with gcc (a few hundred ms)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-ffp-contract=off", "-shared", "-fPIC",
                               "-o", _LIB, _SRC, "-lm"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_level_scale.restype = ctypes.c_float
        _lib.oracle_level_scale.argtypes = [ctypes.c_uint32, ctypes.c_float, ctypes.c_uint32]
    return _lib


def _p(t):
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        assert t.device.type == "cpu" and t.is_contiguous()
        return ctypes.c_void_p(t.data_ptr())
    assert t.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(t.ctypes.data)


def level_scales(L, S, H):
    return np.array([lib().oracle_level_scale(l, ctypes.c_float(S), H) for l in range(L)], dtype=np.float32)


def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                        corner_ids=None, scales=None):
    """Same argument order as the reference op (hashencoder.h:13) + two optional extras."""
    assert inputs.dtype == torch.float32 and embeddings.dtype == torch.float32 and offsets.dtype == torch.int32
    lib().oracle_hashgrid_forward(_p(inputs), _p(embeddings), _p(offsets), _p(outputs),
                                  ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L),
                                  ctypes.c_float(S), ctypes.c_uint32(H), ctypes.c_int(int(bool(calc_grad_inputs))),
                                  _p(dy_dx if calc_grad_inputs else None), _p(corner_ids), _p(scales))


def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H,
                         calc_grad_inputs, dy_dx, grad_inputs, scales=None):
    """Same argument order as the reference op (hashencoder.h:14)."""
    lib().oracle_hashgrid_backward(_p(grad), _p(inputs), _p(offsets), _p(grad_embeddings),
                                   ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L),
                                   ctypes.c_float(S), ctypes.c_uint32(H), ctypes.c_int(int(bool(calc_grad_inputs))),
                                   _p(dy_dx if calc_grad_inputs else None),
                                   _p(grad_inputs if calc_grad_inputs else None), _p(scales))


def grid_offsets(input_dim=3, num_levels=16, per_level_scale=None, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=2048):
    """Level offset table exactly as the reference sizes it (encoder/hashencoder/hashgrid.py:84-108)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    cap = 2 ** log2_hashmap_size
    offs, total = [], 0
    for i in range(num_levels):
        res = int(np.ceil(base_resolution * per_level_scale ** i))
        offs.append(total)
        total += min(cap, (res + 1) ** input_dim)
    offs.append(total)
    return np.array(offs, dtype=np.int32), float(per_level_scale)


def encode(x01, embeddings, offsets, per_level_scale, base_resolution=16, want_ids=False, scales=None):
    """x01 [B,D] in [0,1] -> features [B, L*C] (the layout HashEncoder.forward returns,
    hashgrid.py:41), optionally the corner ids [L,B,2^D]."""
    x01 = x01.contiguous().float()
    B, D = x01.shape
    L = offsets.shape[0] - 1
    C = embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    out = torch.empty(L, B, C, dtype=torch.float32)
    ids = torch.empty(L, B, 1 << D, dtype=torch.int32) if want_ids else None
    hash_encode_forward(x01, embeddings.contiguous(), offsets.contiguous(), out, B, D, C, L, S,
                        base_resolution, False, None, corner_ids=ids, scales=scales)
    feats = out.permute(1, 0, 2).reshape(B, L * C)
    return (feats, ids) if want_ids else feats
