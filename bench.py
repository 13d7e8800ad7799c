#!/usr/bin/env python
"""bench.py -- headline benchmark of the AvatarCraft hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): render_canonical.py at 256x256, 64+64 samples per ray,
hash-encoded NeuS (Instant-NSR), bound 1.6, white background, eval mode.  One STEP = one full
frame = 65 536 rays through the complete `NeRFRenderer.run` semantics (1008 SDF evaluations
and 128 colour evaluations per ray).  Synthetic "trained-like" checkpoint (seed 43) and the
reference's canonical orbit camera; no dataset or checkpoint is available offline.

Reported on ONE JSON line:
  value     rays/s, whole job, rays already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the reference-facing driver `render_instantnsr_naive` with
            HOST (pinned) ray buffers: H2D of the rays and D2H of the rgb inside the timed region
  roofline  dominant kernel (nsr_render_kernel): algorithmic gather bytes (SURVEY.md 8d:
            1 032 192 B/ray + 56 B/ray compulsory + the 48.96 MB table once per launch) / measured
            launch time vs the measured HBM peak; plus the MLP FLOP rate vs the tensor peak
  cpu_baseline  the oracle port (oracle/nsr_oracle.py, all host threads) on a 4096-ray batch of
            the same frame -- a reported baseline, not the target
N > 1 (torchrun): every rank renders its own view of the orbit (weak scaling, no collective on
the data path -- rays are independent); value = all ranks' rays / max-over-ranks time.

`--impl reference` times the reference's CPU implementation of the path (the oracle port: the
reference's Python cannot travel to the GPU box and has no CPU hash kernel of its own) on the
host cores, each step one 4096-ray batch of the same frame.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W_IMG = H_IMG = 256
NUM_STEPS, UPSAMPLE_STEPS, BOUND = 64, 64, 1.6
RAYS_PER_FRAME = W_IMG * H_IMG
GATHER_BYTES_PER_RAY = 1008 * 16 * 8 * 8          # evals x levels x corners x 8 B
COMPULSORY_BYTES_PER_RAY = 56
TABLE_BYTES = 6119857 * 2 * 4
MLP_FLOP_PER_RAY = 1008 * 6528 + 128 * 11264
CPU_SAMPLE_RAYS = 4096
WORKLOAD = "render_canonical 256x256, 64+64 samples/ray, hash-encoded NeuS (Instant-NSR), bound 1.6"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(hbm_gbs=float(j["hbm_gbs"]), bf16_tflops=float(j["bf16_tflops"]), source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, f"/tmp/bench_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [v for v in sm if v >= 0.5 * max(sm)]
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def frame_rays(view_index):
    from avatarcraft_b200.utils import synthetic as syn
    return syn.pinhole_rays(syn.orbit_pose(30.0 + 6.0 * view_index), W_IMG, H_IMG)


def cpu_threads():
    """Threads for the CPU baseline: every core up to 32 -- beyond that the small per-batch GEMMs and the C hash
    restatement (OpenMP) get slower, not faster (measured on the 128-core GPU box: 128 threads 120 rays/s)."""
    n = min(os.cpu_count() or 1, 32)
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


_CPU_MODEL = {}


def cpu_reference_kind():
    """"reference": the reference's own NeRFRenderer.run / NeRFNetwork (unmodified Python from baseline/_ref, torch CPU) with the
    CUDA-only hash kernel replaced by the oracle's C restatement; "port": the oracle restatement of `run` (baseline/_ref absent)."""
    from baseline import ref_loader
    return "reference" if ref_loader.available() else "port"


def cpu_oracle_rate(n_rays, repeats=1):
    """rays/s of the reference's CPU implementation of the path (see cpu_reference_kind), all host threads."""
    import torch
    from avatarcraft_b200.utils import synthetic as syn
    torch.set_num_threads(cpu_threads())
    kind = cpu_reference_kind()
    if "m" not in _CPU_MODEL:
        sd = syn.synthetic_state_dict("trained", 43)
        if kind == "reference":
            from baseline import ref_loader
            net = ref_loader.load_reference(cuda=False).NeRFNetwork()
            net.load_state_dict(sd)
            net.eval()
            _CPU_MODEL["m"] = lambda o, d: net.run(o[None], d[None], NUM_STEPS, BOUND, UPSAMPLE_STEPS, None, cos_anneal_ratio=1.0,
                                                   normal_epsilon_ratio=0.0, render_can=True)
        else:
            from oracle.nsr_oracle import OracleNSR
            m = OracleNSR(sd)
            _CPU_MODEL["m"] = lambda o, d: m.run(o, d, NUM_STEPS, BOUND, UPSAMPLE_STEPS)
    o, d = frame_rays(0)
    sel = slice(RAYS_PER_FRAME // 2 - n_rays // 2, RAYS_PER_FRAME // 2 + n_rays // 2)   # central rows: rays that hit
    best = None
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            _CPU_MODEL["m"](o[sel], d[sel])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
    return n_rays / best, best


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path (the oracle port -- the reference's own hash
    kernel is CUDA-only and its Python cannot travel to the box) on the host cores.  One step = a bounded sample of
    the frame: as many central rays (a multiple of 256, at most one 4096-ray batch) as keep K steps within ~150 s."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_threads()
    for _ in range(max(args.warmup, 1)):
        cpu_oracle_rate(256)
    rate, _ = cpu_oracle_rate(512)
    budget_s = 150.0 / max(args.steps, 1)
    n_rays = int(max(256, min(CPU_SAMPLE_RAYS, (rate * budget_s) // 256 * 256)))
    t = 0.0
    for _ in range(args.steps):
        _, dt = cpu_oracle_rate(n_rays)
        t += dt
    value = n_rays * args.steps / t
    kind = cpu_reference_kind()
    sample = (f"{n_rays} central rays of the 256x256 frame per step, 64+64 samples (" +
              ("the reference's own NeRFRenderer.run, torch CPU, C restatement of its CUDA-only hash kernel)" if kind == "reference"
               else "oracle port, torch CPU + C hash grid)"))
    print(json.dumps({
        "impl": "reference", "metric": "rays_per_sec_volume_render", "value": value, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rays_per_step": n_rays, "device": "host CPU"},
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from avatarcraft_b200 import _lib
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    from avatarcraft_b200.utils import synthetic as syn
    from avatarcraft_b200.utils.render_utils import render_instantnsr_naive

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU baseline")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    lib = _lib.lib()

    net = NeRFNetwork()
    net.load_state_dict(syn.synthetic_state_dict("trained", 43))
    net = net.to(dev).eval()
    o_h, d_h = frame_rays(rank)
    o_pin, d_pin = o_h.pin_memory(), d_h.pin_memory()
    rgb_pin = torch.empty(RAYS_PER_FRAME, 3).pin_memory()
    o, d = o_h.to(dev), d_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    torch.set_grad_enabled(False)          # inference: the fused kernel, no autograd graph (render_canonical.py is no-grad)

    def step_resident():
        return net.run(o[None], d[None], NUM_STEPS, BOUND, UPSAMPLE_STEPS, None, 1.0, 0.0, per_sample_outputs=False)

    def step_e2e():
        ro = o_pin.to(dev, non_blocking=True)
        rd = d_pin.to(dev, non_blocking=True)
        rgb, _ = render_instantnsr_naive(net, ro, rd, rays_per_batch=4096, render_can=True, perturb=False,
                                         num_steps=NUM_STEPS, upsample_steps=UPSAMPLE_STEPS, bound=BOUND)
        rgb_pin.copy_(rgb, non_blocking=True)

    def timed(fn, steps):
        """Sum of per-step CUDA-event times on the current stream; L2 flushed before every step."""
        total = 0.0
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            total += e0.elapsed_time(e1)
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                     # nvidia-smi takes ~0.5 s to start: begin before the warm-up
    for _ in range(max(args.warmup, 3)):
        step_resident()
        step_e2e()
    barrier()
    launches0 = lib.ac_launch_count()
    barrier()
    ms = timed(step_resident, args.steps)
    barrier()
    launches = lib.ac_launch_count() - launches0
    ms_e2e = timed(step_e2e, args.steps)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # second half of BASELINE.json's metric ("... + SDS steps/sec"): the stylisation optimiser step, outside the timed region above
    torch.set_grad_enabled(True)
    try:
        if os.environ.get("AC_BENCH_SKIP_SDS"):          # profiling runs (ncu) only want the render launches
            raise RuntimeError("skipped (AC_BENCH_SKIP_SDS)")
        full = measure_train(10, 3, world, rank, dev, sd_guidance=True)      # BASELINE.json configs[2] with the guidance in the loop
        sds = {k: full[k] for k in ("metric", "value", "unit", "ms_per_step", "steps", "scaling", "config", "gpu_launches", "phases_ms")}
        nerf = measure_train(10, 3, world, rank, dev)                         # the same step with a stand-in pixel gradient
        sds["nerf_side_only"] = {k: nerf[k] for k in ("value", "unit", "ms_per_step", "steps", "gpu_launches", "phases_ms")}
        c = measure_train(20, 3, world, rank, dev, coarse=True)               # coarse stage (one 4096-ray patch), stand-in gradient
        sds["coarse_stage_nerf_side_only"] = {k: c[k] for k in ("value", "unit", "ms_per_step", "steps", "gpu_launches", "phases_ms")}
        if world == 1:
            sds["reference_gpu_path"] = measure_reference_gpu(dev)            # the reference's own model code + CUDA kernel on this GPU
        try:
            c5 = measure_train(4, 4, world, rank, dev, sd_guidance=True, config5=True)    # BASELINE.json configs[4]: one step per camera box
            sds["config5_512_multibbox_sd21"] = {k: c5[k] for k in ("value", "unit", "ms_per_step", "steps", "gpu_launches", "phases_ms", "config")}
        except Exception as e:
            sds["config5_512_multibbox_sd21"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    except Exception as e:           # never lose the headline line to the secondary workload
        sds = {"metric": "sds_style_steps_per_sec", "error": f"{type(e).__name__}: {e}"[:300]}
    torch.set_grad_enabled(False)
    try:
        if os.environ.get("AC_BENCH_SKIP_WARP"):
            raise RuntimeError("skipped (AC_BENCH_SKIP_WARP)")
        warp = measure_warp_frame(max(args.steps // 2, 5), dev, world, rank)   # BASELINE.json configs[3]: animate frame through the warp
    except Exception as e:
        warp = {"metric": "warp_frame_rays_per_sec", "error": f"{type(e).__name__}: {e}"[:300]}
    t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        rays_total = RAYS_PER_FRAME * args.steps * world
        value = rays_total / (ms * 1e-3)
        e2e = rays_total / (ms_e2e * 1e-3)
        launch_s = ms * 1e-3 / args.steps                     # one render launch per step (+ a 1-block reduce)
        algo_bytes = RAYS_PER_FRAME * (GATHER_BYTES_PER_RAY + COMPULSORY_BYTES_PER_RAY) + TABLE_BYTES
        achieved = algo_bytes / launch_s / 1e9
        tflops = RAYS_PER_FRAME * MLP_FLOP_PER_RAY / launch_s / 1e12
        cpu_rate, cpu_dt = cpu_oracle_rate(CPU_SAMPLE_RAYS)
        import glob
        profs = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_render_kernel_traffic.json")))      # newest round's ncu capture
        traffic = json.load(open(profs[-1])).get("dram_bytes_per_launch") if profs else None
        traffic_src = os.path.relpath(profs[-1], ROOT) if profs else None
        line = {
            "metric": "rays_per_sec_volume_render", "value": value, "unit": "rays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "rays_per_step_per_gpu": RAYS_PER_FRAME, "num_steps": NUM_STEPS,
                       "upsample_steps": UPSAMPLE_STEPS, "sdf_evals_per_ray": 1008, "color_evals_per_ray": 128,
                       "checkpoint": "synthetic trained-like seed 43", "parallelism": f"ray-shard x{world} (one view per rank)",
                       "l2": "flushed before every timed step (256 MiB memset); per-step CUDA events summed"},
            "roofline": {"bound": "hbm", "kernel": "nsr_render_tc_kernel", "achieved": achieved, "peak": pk["hbm_gbs"],
                         "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "traffic_source": traffic_src,
                         "bound_ncu": "l1tex data pipe (gathers served by L1/L2; DRAM < 0.1 % of peak)",
                         "peak_source": pk["source"], "algorithmic_bytes_per_launch": algo_bytes,
                         "note": "algorithmic gather bytes (no reuse) per SURVEY.md 8(d); gathers are served from L1/L2, "
                                 "so frac may exceed 1 -- see traffic (ncu dram bytes) and profiles/",
                         "mlp_tflops": tflops, "mlp_frac_of_bf16_peak": tflops / pk["bf16_tflops"],
                         "measured_limiter": "L1 data pipe: l1tex__throughput 76 % of peak, data-stage wavefronts 71 %, issue slots 51 %, "
                                             "tensor pipe 3.7 %, DRAM 79 MB per launch (ncu, profiles/r02p_render_kernel_tc.md)"},
            "cpu_baseline": {"value": cpu_rate, "unit": "rays/s", "cores": cpu_threads(), "kind": cpu_reference_kind(),
                             "sample": f"{CPU_SAMPLE_RAYS}-ray batch (central rows) of the same frame, {cpu_dt:.1f} s"},
            "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": 2 * RAYS_PER_FRAME * 12,
                    "d2h_bytes_per_step": RAYS_PER_FRAME * 12, "ms_per_step": ms_e2e / args.steps,
                    "api": "avatarcraft_b200.utils.render_utils.render_instantnsr_naive (pinned host rays in, pinned host rgb out)"},
            "gpu_launches": int(launches), "clocks": clocks, "sds_step": sds, "warp_frame": warp}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


BBOXES = {"body": dict(center=(0.0, 0.0, 0.0), dist=1.7), "head": dict(center=(0.0, 0.45, 0.0), dist=0.6),
          "upper": dict(center=(0.0, 0.2, 0.0), dist=1.0), "lower": dict(center=(0.0, -0.35, 0.0), dist=1.0)}


def measure_train(steps, warmup, world, rank, dev, sd_guidance=False, coarse=False, config5=False):
    """Secondary workload (BASELINE.json configs[2]): one stylisation optimiser step of stylize.py on a 256x256 view =
    65 536 rays = 16 patches of batch_size 4096 (or, `coarse`, the stride-4 sub-sampled view = one patch):
    pass 1 no-grad render (ray-sharded + one all-gather), the pixel gradient, pass 2 patch re-renders with gradients +
    eikonal + opacity vs a frozen copy (patch-sharded), ONE in-place gradient all-reduce, ONE flat Adam launch.
    `sd_guidance`: the pixel gradient comes from the SDS guidance (SD-1.5-shaped UNet on the native tcgen05 path + VAE
    encoder forward/backward; random weights -- the HF checkpoints are not available offline); otherwise a fixed randn
    tensor stands in and the number is the NeRF side of the step alone.
    Returns the JSON dict (process group must already exist when world > 1)."""
    import torch
    import torch.distributed as dist
    from avatarcraft_b200 import _lib
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    from avatarcraft_b200.utils import synthetic as syn
    from avatarcraft_b200.utils.optim import FlatAdam
    from avatarcraft_b200.utils.render_utils import render_instantnsr_naive
    from avatarcraft_b200.utils.train_utils import stylize_patch_step
    sd = syn.synthetic_state_dict("trained", 43)
    net = NeRFNetwork(); net.load_state_dict(sd); net = net.to(dev).train()
    gt = NeRFNetwork(); gt.load_state_dict(sd); gt = gt.to(dev).eval()
    for p in gt.parameters():
        p.requires_grad_(False)
    opt = FlatAdam(net.parameters(), lr=5e-3)
    from avatarcraft_b200.utils.distributed import render_rays_sharded
    o, d = frame_rays(0)
    side = 64 if coarse else 256
    if coarse:            # coarse stage of stylize.py: the 256x256 view sub-sampled with stride 4 = ONE 4096-ray patch
        o, d = o.reshape(256, 256, 3)[::4, ::4].reshape(-1, 3), d.reshape(256, 256, 3)[::4, ::4].reshape(-1, 3)
    views = None
    if config5:           # BASELINE.json configs[4]: 512x512 views of four camera boxes, cycled step by step; SD-2.1-shaped UNet
        side = 512
        views = [tuple(t.contiguous().to(dev) for t in syn.pinhole_rays(syn.orbit_pose(30.0 + 90.0 * i, **bb), 512, 512))
                 for i, bb in enumerate(BBOXES.values())]
        o, d = views[0]
    o, d = o.contiguous().to(dev), d.contiguous().to(dev)      # fine stage (BASELINE.json configs[2]): 65 536 rays = 16 patches of 4096
    G = torch.randn(o.shape[0], 3, generator=torch.Generator().manual_seed(44)).to(dev)
    torch.manual_seed(1000 + rank)
    sd = emb = None
    if sd_guidance:       # the real guidance: SD-1.5-shaped UNet (native tcgen05 forward) + VAE encoder with gradient, random weights
        from avatarcraft_b200.models.diffusion import StableDiffusion
        sd = StableDiffusion(dev, "2.1" if config5 else "1.5")
        sd.cfg_parallel = world > 1          # every rank seeds the step identically (pixel_gradient(seed=...)): the CFG pair may be split
        emb = sd.get_text_embeds("a 3D rendering of a knight in bronze armour")
    count = [0, 0]
    events = []

    def mark(name):
        if events is not None and timing[0]:
            e = torch.cuda.Event(enable_timing=True); e.record(); events.append((name, e))
    timing = [False]
    from avatarcraft_b200.utils.train_utils import native_patch_step

    def step():
        nonlocal o, d, G
        if views is not None:
            o, d = views[count[1] % len(views)]
            count[1] += 1
            if G.shape[0] != o.shape[0]:
                G = torch.randn(o.shape[0], 3, generator=torch.Generator().manual_seed(44)).to(dev)
        mark("start")
        with torch.no_grad():                                                                            # pass 1, ray-sharded + one all-gather
            rgb = render_rays_sharded(lambda a, b: render_instantnsr_naive(net, a, b, rays_per_batch=4096, render_can=True, perturb=True)[0],
                                      o, d, rank, world)
        mark("pass1")
        g = G
        if sd is not None:
            count[0] += 1
            g = sd.pixel_gradient(emb, rgb, side, side, 100.0, seed=77 + count[0])                        # SDS (models/diffusion.py:92-149)
        mark("guidance")
        native_patch_step(net, gt, opt, o, d, g, batch_size=4096, rank=rank, world=world, mark=mark)      # pass 2 + allreduce + Adam

    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = _lib.lib().ac_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timing[0] = True
    e0.record()
    for _ in range(steps):
        step()
    e1.record(); e1.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier(); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    phases = {}
    for (n0, a), (n1, b) in zip(events[:-1], events[1:]):            # rank 0's per-phase device time, averaged over the timed steps
        if n1 != "start":
            phases[n1] = phases.get(n1, 0.0) + a.elapsed_time(b) / steps
    return {"metric": "sds_style_steps_per_sec", "phases_ms": {k: round(v, 3) for k, v in phases.items()}, "value": steps / (float(ms) * 1e-3), "unit": "steps/s", "n_gpus": world,
            "steps": steps, "warmup": max(warmup, 3), "ms_per_step": float(ms) / steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": ("stylize.py coarse-stage step: 4096 rays (256x256 stride 4)" if coarse else
                                    "multi-bbox stylize step on a 512x512 view (4 camera boxes body/head/upper/lower cycled): 262 144 rays = 64 patches "
                                    "of batch_size 4096, SD-2.1-shaped fp16 UNet, guidance at 512x512 without resize" if config5 else
                                    "stylize.py step on a 256x256 view: 65 536 rays = 16 patches of batch_size 4096") + ", 64+64 samples, pass1 + pass2 "
                                   "(grad, eikonal 0.01, opacity vs frozen copy) + grad all-reduce + Adam; pixel gradient " +
                                   ("from the SDS guidance: SD-" + ("2.1" if config5 else "1.5") + "-shaped UNet (random init, native tcgen05 forward on the "
                                    "(uncond, text) pair, fp16 operands) + native VAE encoder forward/backward at 512x512" if sd_guidance else "randn seed 44 (guidance excluded)"),
                       "parallelism": f"pass 1 ray-sharded + all-gather, pass 2 patch-sharded x{world}, one 49 MB gradient all-reduce, "
                                      "SD guidance: classifier-free pair split over ranks, VAE replicated"},
            "gpu_launches": int(_lib.lib().ac_launch_count() - l0)}


def measure_reference_gpu(dev):
    """B-REF-GPU: the REFERENCE'S OWN model code (baseline/_ref/models/instant_nsr.py, unmodified) with
    its own hash-encoder CUDA kernel compiled for sm_100a (oracle/_ref/_ref_hash_encoder.so) on this GPU: one 256x256 inference
    frame in 4096-ray batches (render_instantnsr_naive's loop, utils/render_utils.py:514-600) and one pass-2 patch of the
    trainer (stylize.py:153-199: render with gradients, pixel gradient + eikonal + opacity vs a frozen copy, backward)."""
    import torch
    from baseline import ref_loader
    if not ref_loader.available():
        return {"unavailable": "baseline/_ref not installed"}
    try:
        ref = ref_loader.load_reference(cuda=True)
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    from avatarcraft_b200.utils import synthetic as syn
    sd = syn.synthetic_state_dict("trained", 43)
    net = ref.NeRFNetwork(); net.load_state_dict(sd); net = net.to(dev).train()
    gt = ref.NeRFNetwork(); gt.load_state_dict(sd); gt = gt.to(dev).eval()
    for p in gt.parameters():
        p.requires_grad_(False)
    optim = torch.optim.Adam(net.parameters(), lr=5e-3)
    o, d = frame_rays(0)
    o, d = o.to(dev), d.to(dev)
    G = torch.randn(o.shape[0], 3, generator=torch.Generator().manual_seed(44)).to(dev)
    kw = dict(num_steps=NUM_STEPS, upsample_steps=UPSAMPLE_STEPS, bound=BOUND, staged=False, bg_color=None, cos_anneal_ratio=1.0,
              normal_epsilon_ratio=0.0, render_can=True)

    def frame():
        with torch.no_grad():
            for s0 in range(0, RAYS_PER_FRAME, 4096):
                net.render(o[None, s0:s0 + 4096], d[None, s0:s0 + 4096], perturb=False, **kw)

    def patch(s0):
        out = net.render(o[None, s0:s0 + 4096], d[None, s0:s0 + 4096], perturb=True, **kw)
        with torch.no_grad():
            og = gt.render(o[None, s0:s0 + 4096], d[None, s0:s0 + 4096], perturb=True, **kw)["weight_sum"]
        loss = (out["rgb"].reshape(-1, 3) * G[s0:s0 + 4096]).sum() + 0.01 * out["gradient_error"] + \
            torch.nn.functional.smooth_l1_loss(out["weight_sum"].clamp(0, 1), og.clamp(0, 1)) * 1e5
        loss.backward()

    def timed(fn, n):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); e1.synchronize()
        return e0.elapsed_time(e1) / n
    ms_frame = timed(frame, 2)
    optim.zero_grad()
    ms_patch = timed(lambda: patch(8 * 4096), 3)
    optim.zero_grad()

    def full_step():
        frame()
        optim.zero_grad()
        for s0 in range(0, RAYS_PER_FRAME, 4096):
            patch(s0)
        optim.step()
    ms_step = timed(full_step, 1)
    return {"what": "reference models/instant_nsr.py (unmodified) + reference hashencoder.cu rebuilt for sm_100a, eager torch, this GPU",
            "inference_frame_ms": ms_frame, "inference_rays_per_sec": RAYS_PER_FRAME / ms_frame * 1e3, "train_patch_ms": ms_patch,
            "train_step_nerf_side_ms": ms_step, "train_steps_per_sec_nerf_side": 1e3 / ms_step}


def measure_warp_frame(steps, dev, world, rank):
    """BASELINE.json configs[3]: one render_warp.py --render_type animate frame at 256x256, 32+32 samples, batch 8192, on a
    synthetic SMPL-shaped body (6890 vertices / 13 776 faces, 24 joints): per-frame posed-mesh preparation + mesh-guided
    near/far + inverse-LBS warp of every sample (96 closest-point queries per ray) + the fused render.  Each rank renders its
    own frame of the sequence (weak scaling, no collective)."""
    import torch
    import torch.distributed as dist
    from avatarcraft_b200 import _lib
    from avatarcraft_b200.models.instant_nsr import NeRFNetwork
    from avatarcraft_b200.utils import synthetic as syn
    from avatarcraft_b200.utils.render_utils import render_instantnsr_naive
    net = NeRFNetwork(); net.load_state_dict(syn.synthetic_state_dict("trained", 43)); net = net.to(dev).eval()
    net.warp_skip_masked = os.environ.get("AC_WARP_EXACT_ALL", "") == ""       # what render_warp.py sets (it keeps only the image)
    body = syn.synthetic_body()
    o, d = syn.pinhole_rays(syn.orbit_pose(10.0 + 3.0 * rank), W_IMG, H_IMG)
    o, d = o.to(dev), d.to(dev)

    def frame():
        return render_instantnsr_naive(net, o, d, 8192, render_can=False, perturb=False, verts=body["world_verts"], faces=body["faces"],
                                       Ts=body["Ts"], num_steps=32, upsample_steps=32, bound=BOUND)
    def timed(n):
        for _ in range(3):
            frame()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            frame()
        e1.record(); e1.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.barrier(); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / n
    # the same frame with every masked-out sample searched to the end and evaluated (the flag off): reported beside the headline
    skip = net.warp_skip_masked
    net.warp_skip_masked = False
    per_all = timed(max(3, steps // 2))
    net.warp_skip_masked = skip
    l0 = _lib.lib().ac_launch_count()
    per = timed(steps)
    l0 += 3 * (_lib.lib().ac_launch_count() - l0) // (steps + 3)          # the warm-up frames' launches are not in the timed region
    return {"metric": "warp_frame_rays_per_sec", "value": world * RAYS_PER_FRAME / per * 1e3, "unit": "rays/s", "ms_per_frame": per, "steps": steps,
            "closest_point_queries_per_sec": world * RAYS_PER_FRAME * 96 / per * 1e3, "gpu_launches": int(_lib.lib().ac_launch_count() - l0),
            "config": {"workload": "render_warp.py animate frame 256x256, 32+32 samples/ray, batch 8192, synthetic SMPL-shaped body "
                                   "(6890 verts / 13 776 faces), mesh prep + near/far + warp (96 closest-point queries/ray) + render",
                       "sdf_evals_per_ray": 496, "color_evals_per_ray": 64,
                       "warp_skip_masked": bool(net.warp_skip_masked)},
            "all_samples_evaluated": {"ms_per_frame": per_all, "value": world * RAYS_PER_FRAME / per_all * 1e3, "unit": "rays/s",
                                      "what": "warp_skip_masked off: masked-out samples (alpha x 0) are searched to the end and evaluated "
                                              "too; image, depth, opacity and weights are bit-identical either way"}}


def run_train(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_train(args.steps, args.warmup, world, rank, torch.device("cuda", local), sd_guidance=args.sd_guidance, coarse=args.coarse)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="render", choices=["render", "train"])
    ap.add_argument("--coarse", action="store_true", help="--workload train: the coarse stage (one 4096-ray patch) instead of the 256x256 view")
    ap.add_argument("--sd_guidance", action="store_true", help="--workload train: put the SD UNet + VAE guidance inside the step")
    args = ap.parse_args()
    if args.workload == "train" and args.impl == "ours":
        return run_train(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
