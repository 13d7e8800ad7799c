"""SHEncoder with the reference's module API (encoder/shencoder/sphere_harmonics.py:58-83)."""
import torch
import torch.nn as nn
from torch.autograd import Function

from .backend import _backend


class _sh_encoder(Function):
    @staticmethod
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous()
        B, input_dim = inputs.shape
        outputs = torch.empty(B, degree ** 2, dtype=inputs.dtype, device=inputs.device)
        dy_dx = torch.empty(B, input_dim * degree ** 2, dtype=inputs.dtype, device=inputs.device) if calc_grad_inputs \
            else torch.empty(1, dtype=inputs.dtype, device=inputs.device)
        _backend.sh_encode_forward(inputs, outputs, B, input_dim, degree, calc_grad_inputs, dy_dx)
        ctx.save_for_backward(inputs, dy_dx)
        ctx.dims, ctx.calc_grad_inputs = (B, input_dim, degree), calc_grad_inputs
        return outputs

    @staticmethod
    def backward(ctx, grad):
        if not ctx.calc_grad_inputs:
            return None, None, None
        inputs, dy_dx = ctx.saved_tensors
        B, input_dim, degree = ctx.dims
        grad_inputs = torch.zeros_like(inputs)
        _backend.sh_encode_backward(grad.contiguous(), inputs, B, input_dim, degree, dy_dx, grad_inputs)
        return grad_inputs, None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert 0 < self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        return sh_encode(flat, self.degree, flat.requires_grad).reshape(prefix + [self.output_dim])
