"""Run under torchrun (NCCL, one rank per GPU) by tests/test_gpu_multi.py:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_grad_check.py OUT.json

1-vs-N rank gradient equality of the stylisation step (SURVEY.md section 4 "multi-GPU"; the reference is single-GPU, so the
single-process step IS the semantics to preserve): every rank runs utils/train_utils.stylize_patch_step on its shard of the
patches (and, in the coarse case, on its slice of ONE split patch, with the masked-eikonal share) followed by the in-place NCCL
all-reduce; rank 0 then repeats the step alone (world = 1) on the same rays / jitter / pixel gradient and compares the flat
gradient and the updated parameters."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def main():
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    from avatarcraft_b200.utils import synthetic as syn
    from avatarcraft_b200.utils.optim import FlatAdam
    from avatarcraft_b200.utils.train_utils import native_patch_step
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime
    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))      # a mismatched collective fails fast
    solo = dist.new_group([0])           # rank 0's single-process rerun must not join the job-wide all-reduce
    sd = syn.synthetic_state_dict("trained", 43)
    o, d = syn.pinhole_rays(syn.orbit_pose(30.0), 64, 64)
    gen = torch.Generator().manual_seed(21)
    report = {"world": world}
    for case, (sel, batch) in {"patches": (torch.arange(64 * 16, 64 * 48), 512),          # 2048 rays = 4 whole patches, round-robin
                               "split_patch": (torch.arange(64 * 20, 64 * 36), 1024)}.items():     # 1024 rays = ONE patch split over the ranks
        oo, dd = o[sel].to(dev), d[sel].to(dev)
        G = torch.randn(oo.shape[0], 3, generator=gen).to(dev)
        jit = torch.rand(oo.shape[0], 64, generator=gen).to(dev)

        def fresh(group=None):
            net = NeRFNetwork(); net.load_state_dict(sd); net = net.to(dev).train()
            gt = NeRFNetwork(); gt.load_state_dict(sd); gt = gt.to(dev).eval()
            with torch.no_grad():
                gt.sdf_net[1].bias[0] += 0.05                        # a frozen copy that differs: non-zero opacity term
            for p in gt.parameters():
                p.requires_grad_(False)
            return net, gt, FlatAdam(net.parameters(), lr=5e-3, group=group)
        net, gt, opt = fresh()
        native_patch_step(net, gt, opt, oo, dd, G, batch_size=batch, rank=rank, world=world, jitter=jit)
        torch.cuda.synchronize()
        grad_n, param_n = opt.flat_grad.clone(), opt.flat_param.clone()
        # every rank must hold the same reduced gradient and the same parameters after the step
        ref_g = grad_n.clone(); dist.broadcast(ref_g, 0)
        same_across_ranks = bool(torch.equal(ref_g, grad_n))
        flags = torch.tensor([int(same_across_ranks)], device=dev); dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        if rank == 0:
            net1, gt1, opt1 = fresh(solo)
            native_patch_step(net1, gt1, opt1, oo, dd, G, batch_size=batch, rank=0, world=1, jitter=jit)
            torch.cuda.synchronize()
            per = {}
            for (k, p), (a, n) in zip(net1.named_parameters(), opt1._spans):
                per[k] = rel_l2(grad_n[a:a + n], opt1.flat_grad[a:a + n])
            report[case] = {"grad_rel_l2": per, "grad_rel_l2_max": max(per.values()), "param_rel_l2": rel_l2(param_n, opt1.flat_param),
                            "identical_on_all_ranks": bool(flags.item())}
        dist.barrier()
    if rank == 0:
        with open(sys.argv[1], "w") as f:
            json.dump(report, f, indent=1)
        print(json.dumps(report))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
