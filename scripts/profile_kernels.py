"""Launch each secondary kernel once on representative shapes (for `ncu --set full -k regex:...`):
sd_gemm (TMA main loop: a 3x3 conv of the 64x64 level, a weight-streaming split-K shape, a small-K projection),
the fused attention of the 64x64 level, the training kernels (tensor-core forward / fused backward), flat Adam."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avatarcraft_b200 import _lib
from avatarcraft_b200.models import sd_native
from avatarcraft_b200.models.instant_nsr import NeRFNetwork, _SdfQuery
from avatarcraft_b200.utils import synthetic as syn
from avatarcraft_b200.utils.optim import FlatAdam

g = torch.Generator().manual_seed(0)
for (M, N, K) in ((8192, 640, 5760), (128, 1280, 11520), (8192, 320, 320)):
    A = torch.randn(M, K, generator=g).cuda().half(); W = torch.randn(N, K, generator=g).cuda().half()
    bias = torch.randn(N, generator=g).cuda(); res = torch.randn(M, N, generator=g).cuda()
    for _ in range(2):
        sd_native.gemm(A, W, M, N, K, bias=bias, residual=res)
B, heads, L, d = 2, 8, 4096, 40
inner = heads * d
q = torch.randn(B * L, inner, generator=g).cuda().half(); k = torch.randn(B * L, inner, generator=g).cuda().half()
vt = torch.randn(B, inner, L, generator=g).cuda().half(); o = torch.empty(B * L, inner, device="cuda", dtype=torch.float16)
for _ in range(2):
    _lib.check(_lib.lib().ac_sd_flash_attention_f16(sd_native._p(q), sd_native._p(k), sd_native._p(vt), sd_native._p(o), B, heads, L, L, d, inner, inner, L,
                                                    inner, d ** -0.5, _lib.stream_ptr()), "flash")
torch.cuda.synchronize()
net = NeRFNetwork(); net.load_state_dict(syn.synthetic_state_dict("trained", 43)); net = net.cuda().train()
x = ((torch.rand(3670016, 3, generator=g) * 2 - 1) * 0.8).cuda()
w = [torch._weight_norm(l.weight_v, l.weight_g, 0) for l in net.sdf_net]
out = _SdfQuery.apply(x, net.encoder.embeddings, w[0], net.sdf_net[0].bias, w[1], net.sdf_net[1].bias, net, 1.6)
R = torch.zeros_like(out); R[:, 0] = 1.0; R[:524288] = 1.0
(out * R).sum().backward()
opt = FlatAdam(net.parameters(), lr=5e-3)
opt.step(); opt.step()
torch.cuda.synchronize()
print("done")
