"""`_backend` with the reference's SH operator names (encoder/shencoder/src/bindings.cpp:5-8) on the C ABI."""
import torch

from ... import _lib


class _Backend:
    @staticmethod
    def sh_encode_forward(inputs, outputs, B, D, C, calc_grad_inputs, dy_dx):
        if inputs.dtype != torch.float32:
            raise RuntimeError("avatarcraft_b200 SH encoder computes in float32")
        _lib.check(_lib.lib().ac_sh_encode_forward(_lib.ptr(inputs), _lib.ptr(outputs), B, D, C, int(bool(calc_grad_inputs)),
                                                   _lib.ptr(dy_dx) if calc_grad_inputs else None, _lib.stream_ptr()), "sh_encode_forward")

    @staticmethod
    def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
        _lib.check(_lib.lib().ac_sh_encode_backward(_lib.ptr(grad), _lib.ptr(inputs), B, D, C, _lib.ptr(dy_dx), _lib.ptr(grad_inputs),
                                                    _lib.stream_ptr()), "sh_encode_backward")


_backend = _Backend()
