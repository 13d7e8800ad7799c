"""Multi-resolution hash-grid encoder module (drop-in for `encoder.hashencoder.HashEncoder`,
reference: encoder/hashencoder/hashgrid.py:79-142 -- same constructor arguments, same `embeddings`
parameter / `offsets` buffer, same forward contract).

Design differences from the reference wrapper: the op is point-major end to end -- the kernel writes the
[B, L*C] feature matrix directly and the backward kernel consumes the [B, L*C] gradient directly -- so the two
[L,B,C] <-> [B,L*C] permute copies per call are gone, and the level table is described by one small dataclass
shared by module, autograd function and tests."""
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

from ... import _lib


@dataclass(frozen=True)
class GridSpec:
    dim: int
    levels: int
    channels: int
    base_resolution: int
    per_level_scale: float
    log2_hashmap_size: int

    @property
    def log2_scale(self) -> float:
        return float(np.log2(self.per_level_scale))

    def level_offsets(self) -> np.ndarray:
        """First table entry of every level (+ total): a level is dense while (res+1)^dim fits in
        2^log2_hashmap_size entries, hashed beyond (sizing rule of hashgrid.py:99-108)."""
        cap = 1 << self.log2_hashmap_size
        sizes = [min(cap, (int(np.ceil(self.base_resolution * self.per_level_scale ** l)) + 1) ** self.dim)
                 for l in range(self.levels)]
        return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)


class _GridLookup(torch.autograd.Function):
    """features = lookup(x01, table): fp32, CUDA only; d/d(table) always, d/d(x01) when x01 requires grad."""

    @staticmethod
    def forward(ctx, x01, table, offsets, spec: GridSpec):
        if x01.dtype != torch.float32 or table.dtype != torch.float32:
            raise RuntimeError("avatarcraft_b200 hash encoder computes in float32")
        x01, table = x01.contiguous(), table.contiguous()
        n = x01.shape[0]
        want_dx = x01.requires_grad
        feats = torch.empty(n, spec.levels * spec.channels, device=x01.device, dtype=torch.float32)
        jac = torch.empty(n, spec.levels * spec.dim * spec.channels, device=x01.device, dtype=torch.float32) if want_dx else None
        _lib.check(_lib.lib().ac_hash_encode_forward_pm(_lib.ptr(x01), _lib.ptr(table), _lib.ptr(offsets), _lib.ptr(feats), n,
                                                        spec.dim, spec.channels, spec.levels, spec.log2_scale,
                                                        spec.base_resolution, int(want_dx), _lib.ptr(jac), _lib.stream_ptr()),
                   "hash_encode_forward")
        ctx.save_for_backward(x01, offsets, jac if want_dx else x01.new_empty(0))
        ctx.spec, ctx.table_shape, ctx.want_dx = spec, table.shape, want_dx
        return feats

    @staticmethod
    def backward(ctx, g):
        x01, offsets, jac = ctx.saved_tensors
        spec, n = ctx.spec, x01.shape[0]
        g = g.contiguous()
        g_table = torch.zeros(ctx.table_shape, device=g.device, dtype=torch.float32)
        g_x = torch.empty_like(x01) if ctx.want_dx else None
        _lib.check(_lib.lib().ac_hash_encode_backward_pm(_lib.ptr(g), _lib.ptr(x01), _lib.ptr(offsets), _lib.ptr(g_table), n, spec.dim,
                                                         spec.channels, spec.levels, spec.log2_scale, spec.base_resolution,
                                                         int(ctx.want_dx), _lib.ptr(jac) if ctx.want_dx else None, _lib.ptr(g_x),
                                                         _lib.stream_ptr()), "hash_encode_backward")
        return g_x, g_table, None, None


def hash_encode(x01, table, offsets, spec: GridSpec):
    return _GridLookup.apply(x01, table, offsets, spec)


class HashEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None):
        super().__init__()
        if desired_resolution is not None:        # geometric progression from base to desired resolution (hashgrid.py:84-86)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.spec = GridSpec(input_dim, num_levels, level_dim, base_resolution, float(per_level_scale), log2_hashmap_size)
        offsets = self.spec.level_offsets()
        # attributes the reference module exposes
        self.input_dim, self.num_levels, self.level_dim = input_dim, num_levels, level_dim
        self.per_level_scale, self.base_resolution, self.log2_hashmap_size = per_level_scale, base_resolution, log2_hashmap_size
        self.output_dim = num_levels * level_dim
        self.max_params = 1 << log2_hashmap_size
        self.n_params = int(offsets[-1]) * level_dim
        self.register_buffer("offsets", torch.from_numpy(offsets))
        self.embeddings = nn.Parameter(torch.empty(int(offsets[-1]), level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.uniform_(self.embeddings, -1e-4, 1e-4)       # hashgrid.py:119-121

    def extra_repr(self):
        return (f"dim={self.input_dim} levels={self.num_levels} channels={self.level_dim} base={self.base_resolution} "
                f"scale={self.per_level_scale:.4f} table={tuple(self.embeddings.shape)}")

    def forward(self, inputs, size=1):
        """inputs [..., dim] in [-size, size] -> [..., levels*channels]; points outside the box encode to zeros."""
        lead = inputs.shape[:-1]
        x01 = ((inputs + size) / (2 * size)).reshape(-1, self.input_dim)
        return hash_encode(x01, self.embeddings, self.offsets, self.spec).reshape(*lead, self.output_dim)
