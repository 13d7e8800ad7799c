"""Real spherical-harmonics direction encoder module (drop-in for `encoder.shencoder.SHEncoder`,
reference: encoder/shencoder/sphere_harmonics.py:58-83; degree 1..8, output degree^2 coefficients)."""
import torch
import torch.nn as nn

from ... import _lib


class _SHBasis(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dirs, degree):
        if dirs.dtype != torch.float32:
            raise RuntimeError("avatarcraft_b200 SH encoder computes in float32")
        dirs = dirs.contiguous()
        n, want_dx = dirs.shape[0], dirs.requires_grad
        out = torch.empty(n, degree * degree, device=dirs.device, dtype=torch.float32)
        jac = torch.empty(n, 3 * degree * degree, device=dirs.device, dtype=torch.float32) if want_dx else None
        _lib.check(_lib.lib().ac_sh_encode_forward(_lib.ptr(dirs), _lib.ptr(out), n, 3, degree, int(want_dx), _lib.ptr(jac),
                                                   _lib.stream_ptr()), "sh_encode_forward")
        ctx.degree, ctx.want_dx = degree, want_dx
        ctx.save_for_backward(dirs, jac if want_dx else dirs.new_empty(0))
        return out

    @staticmethod
    def backward(ctx, g):
        if not ctx.want_dx:
            return None, None
        dirs, jac = ctx.saved_tensors
        g_dirs = torch.zeros_like(dirs)
        _lib.check(_lib.lib().ac_sh_encode_backward(_lib.ptr(g.contiguous()), _lib.ptr(dirs), dirs.shape[0], 3, ctx.degree, _lib.ptr(jac),
                                                    _lib.ptr(g_dirs), _lib.stream_ptr()), "sh_encode_backward")
        return g_dirs, None


def sh_encode(dirs, degree):
    return _SHBasis.apply(dirs, degree)


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        if input_dim != 3:
            raise AssertionError("SH encoder only support input dim == 3")
        if not 0 < degree <= 8:
            raise AssertionError("SH encoder only supports degree in [1, 8]")
        self.input_dim, self.degree, self.output_dim = input_dim, degree, degree * degree

    def extra_repr(self):
        return f"degree={self.degree}"

    def forward(self, inputs, size=1):
        lead = inputs.shape[:-1]
        return sh_encode((inputs / size).reshape(-1, 3), self.degree).reshape(*lead, self.output_dim)
