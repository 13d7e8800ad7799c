"""`_backend` with the reference's operator names and argument order
(encoder/hashencoder/backend.py:6-15 builds it with torch.utils.cpp_extension.load;
bindings: encoder/hashencoder/src/bindings.cpp:5-8).  Here it forwards to the C ABI."""
import torch

from ... import _lib


def _check_f32(*ts):
    for t in ts:
        if t.dtype != torch.float32:
            raise RuntimeError("avatarcraft_b200 hash encoder computes in float32 (parity with the reference's "
                               "fp32 path); got " + str(t.dtype))


class _Backend:
    @staticmethod
    def hash_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, calc_grad_inputs, dy_dx,
                            corner_ids=None):
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")
        _check_f32(inputs, embeddings, outputs)
        rc = _lib.lib().ac_hash_encode_forward(_lib.ptr(inputs), _lib.ptr(embeddings), _lib.ptr(offsets),
                                               _lib.ptr(outputs), B, D, C, L, float(S), H, int(bool(calc_grad_inputs)),
                                               _lib.ptr(dy_dx) if calc_grad_inputs else None, _lib.ptr(corner_ids),
                                               _lib.stream_ptr())
        _lib.check(rc, "hash_encode_forward")

    @staticmethod
    def hash_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, calc_grad_inputs,
                             dy_dx, grad_inputs):
        if offsets.dtype != torch.int32:
            raise RuntimeError("offsets must be an int tensor")
        _check_f32(grad, inputs, grad_embeddings)
        rc = _lib.lib().ac_hash_encode_backward(_lib.ptr(grad), _lib.ptr(inputs), _lib.ptr(embeddings),
                                                _lib.ptr(offsets), _lib.ptr(grad_embeddings), B, D, C, L, float(S), H,
                                                int(bool(calc_grad_inputs)),
                                                _lib.ptr(dy_dx) if calc_grad_inputs else None,
                                                _lib.ptr(grad_inputs) if calc_grad_inputs else None, _lib.stream_ptr())
        _lib.check(rc, "hash_encode_backward")


_backend = _Backend()
