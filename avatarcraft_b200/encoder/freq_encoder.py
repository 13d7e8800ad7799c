"""NeRF sin/cos positional encoder (reference: encoder/freq_encoder.py:10-54).  Plain torch:
it is only reachable through get_encoder("frequency") for the legacy models.neus path, which no
entry point of the reference uses (stylize.py:150-151 raises NotImplementedError for neus)."""
import torch
import torch.nn as nn


class FreqEncoder(nn.Module):
    def __init__(self, input_dim=3, multires=6, include_input=True, log_sampling=True):
        super().__init__()
        self.input_dim, self.include_input = input_dim, include_input
        max_freq = multires - 1
        bands = 2.0 ** torch.linspace(0.0, max_freq, multires) if log_sampling else torch.linspace(1.0, 2.0 ** max_freq, multires)
        self.register_buffer("freq_bands", bands, persistent=False)
        self.output_dim = input_dim * (int(include_input) + 2 * multires)

    def forward(self, x, **kwargs):
        out = [x] if self.include_input else []
        for f in self.freq_bands:
            out += [torch.sin(x * f), torch.cos(x * f)]
        return torch.cat(out, dim=-1)


def get_freq_embedder(multires, input_dims=3):
    enc = FreqEncoder(input_dims, multires)
    return enc, enc.output_dim
