#!/usr/bin/env python
"""Headline benchmark: DDPM U-Net training steps/s and 1000-step samples/s (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, via the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU
    python bench.py --config celeba64 ...                    # BASELINE.json configs[2]: CelebA-64 U-Net, 32 images per GPU
    python bench.py --config vqvae ...                       # BASELINE.json configs[4]: VQ-VAE 128x128, 32 images per GPU

One JSON line on stdout (rank 0).  Default workload = BASELINE.json configs[1]: DDPM CIFAR-10 32x32 U-Net
(dims 3->64->128->256, T=1000), batch 128 per GPU.  A "step" is one training step: noise + q_sample + U-Net forward +
L1 loss + backward + [gradient all-reduce] + Adam on one synthetic batch.  `value` = rank-steps per second over all GPUs
(weak scaling).  The same line carries
  * `e2e`: the same metric through DDPM.training_step with pinned-host batches (H2D + loss read-back inside the timed region),
  * `sustained`: the same loop run for >= 2.5 s (SM clocks settle below boost on a long dense run),
  * `samples_per_sec_1000step` (M2: a full captured 1000-step chain) with its own e2e (final images copied to the host),
  * `eager_gpu`: the reference algorithm as PyTorch eager on the SAME GPU (cuDNN / cuBLAS fp32, TF32 off and on) --
    the comparator that matters; the CPU baseline is only a reported number,
  * `pixelcnn` (M3: PixelCNN MNIST sampling, batch 64) and `vqvae` (M4: VQ lookup vectors/s and VQ-VAE train steps/s),
    each with its own roofline / eager-GPU / CPU figures (one-GPU run of the default config only),
  * `roofline` of the dominant kernel class and `cpu_baseline` (oracle port on the host cores).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DIM, CH, T = 64, 3, 1000
# per-config constants: BASELINE.md section 3 (FlopCounterMode on the reference modules)
CONFIGS = {
    "cifar10": dict(mults=(1, 2, 4), H=32, W=32, batch=128, train_gflop=5.153, fwd_gflop=1.719,
                    name="DDPM CIFAR-10 32x32 U-Net dims 3-64-128-256 T=1000 batch={B}/GPU (BASELINE.json configs[1])"),
    "celeba64": dict(mults=(1, 2, 4, 8), H=64, W=64, batch=32, train_gflop=26.195, fwd_gflop=8.737,
                     name="DDPM CelebA 64x64 U-Net dims 3-64-128-256-512 T=1000 batch={B}/GPU "
                          "(BASELINE.json configs[2]: global 256 over 8 GPUs)"),
}
METRIC = "ddpm_unet_train_steps_per_sec"


def unit_of(B):
    return f"steps/s ({B} images per GPU-step)"


def workload_name(cfg, B):
    """config.workload, shared verbatim by our arm and the reference arm."""
    return CONFIGS[cfg]["name"].format(B=B) + "; train step = noise+q_sample+fwd+L1+bwd+allreduce+Adam"


def synth_batch(B, seed, H, W):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, CH, H, W, generator=g) * 0.5).clamp(-1, 1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------
# the reference algorithm as plain PyTorch (the oracle port; the unmodified reference modules when the tree is mounted)
# on any device: the CPU baseline AND the PyTorch-eager-on-the-same-GPU comparator
# ---------------------------------------------------------------------------
def torch_train_setup(cfg, B, dev="cpu"):
    from oracle import ddpm_oracle as O
    c = CONFIGS[cfg]
    spec = O.UnetSpec(DIM, CH, c["mults"])
    params = {k: torch.nn.Parameter(v.to(dev)) for k, v in O.init_params(spec, seed=0).items()}
    opt = torch.optim.Adam(params.values(), lr=1e-4, betas=(0.9, 0.999))
    buf = {k: v.to(dev) for k, v in O.diffusion_buffers(T).items()}
    x = synth_batch(B, 0, c["H"], c["W"]).to(dev)
    g = torch.Generator(device=dev).manual_seed(1)

    def step():
        t = torch.randint(0, T, (B,), generator=g, device=dev)
        noise = torch.randn(B, CH, c["H"], c["W"], generator=g, device=dev)
        opt.zero_grad()
        loss = O.p_losses(params, spec, buf, x, t, noise, "l1")
        loss.backward()
        opt.step()
        return loss

    def sample_step(img, i):
        with torch.no_grad():
            return O.p_sample(params, spec, buf, img, torch.full((B,), i, dtype=torch.long, device=dev),
                              torch.randn(B, CH, c["H"], c["W"], generator=g, device=dev))

    return step, sample_step


def cpu_baseline(cfg, B, train_steps=3, sample_steps=4):
    torch.set_num_threads(os.cpu_count() or 1)
    c = CONFIGS[cfg]
    step, sample_step = torch_train_setup(cfg, B)
    step()
    t0 = time.perf_counter()
    for _ in range(train_steps):
        step().item()
    dt = (time.perf_counter() - t0) / train_steps
    img = torch.randn(B, CH, c["H"], c["W"])
    img = sample_step(img, T - 1)
    t0 = time.perf_counter()
    for i in range(sample_steps):
        img = sample_step(img, T - 2 - i)
    ds = (time.perf_counter() - t0) / sample_steps
    return {
        "value": 1.0 / dt, "unit": unit_of(B), "cores": torch.get_num_threads(), "kind": "port",
        "sample": f"{train_steps} train steps (+1 warm-up) at B={B} and {sample_steps} p_sample steps at B={B}, "
                  f"oracle/ddpm_oracle.py (torch CPU fp32, oneDNN)",
        "ms_per_train_step": dt * 1e3, "ms_per_sample_step": ds * 1e3,
        "samples_per_sec_1000step_extrapolated": B / (ds * T),
    }


def eager_gpu(cfg, B, dev, steps=6, warmup=3, sample_steps=10):
    """The reference algorithm as PyTorch eager on this GPU: cuDNN / cuBLAS fp32 with TF32 off (the parity-grade
    comparator) and on.  Same batch, same optimizer, same step contents as our arm; timed with CUDA events."""
    out = {"impl": "oracle/ddpm_oracle.py on cuda: the reference's ATen op sequence (F.conv2d, F.group_norm, softplus/tanh, "
                   "einsum, torch.optim.Adam), pinned bit-exact to the unmodified reference modules on CPU",
           "batch": B, "cudnn_benchmark": False}
    c = CONFIGS[cfg]
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for mode, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            step, sample_step = torch_train_setup(cfg, B, dev)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            img = torch.randn(B, CH, c["H"], c["W"], device=dev)
            for i in range(3):
                img = sample_step(img, T - 1 - i)
            torch.cuda.synchronize()
            e0.record()
            for i in range(sample_steps):
                img = sample_step(img, T - 4 - i)
            e1.record()
            torch.cuda.synchronize()
            ms_s = e0.elapsed_time(e1) / sample_steps
            out[mode] = {"train_steps_per_sec": 1e3 / ms, "ms_per_train_step": ms, "ms_per_denoise_step": ms_s,
                         "samples_per_sec_1000step_extrapolated": B / (ms_s * T / 1e3),
                         "train_steps_timed": steps, "denoise_steps_timed": sample_steps}
            del step, sample_step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    if args.config == "vqvae":
        import bench_secondary
        return bench_secondary.run_reference_vqvae(args)
    B = args.batch
    step, _ = torch_train_setup(args.config, B)
    for _ in range(args.warmup):
        step().item()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step().item()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": unit_of(B), "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, B),
                   "where": "host CPU only (torch fp32, all cores), ONE rank doing one per-GPU batch: at N > 1 the ratio "
                            "to our N-rank value is not like for like",
                   "parallelism": f"dp{args.gpus}", "global_batch": B * args.gpus},
        "cpu_baseline": {"value": v, "unit": unit_of(B), "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} full train steps at B={B} (oracle port of the reference algorithm; "
                                   "the reference itself is torch-on-CPU and needs Lightning, absent here)"},
        "e2e": {"value": v, "unit": unit_of(B), "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist

    from types import SimpleNamespace

    import igm_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line (the JSON): anything libraries print there (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        opts = None
        if os.environ.get("IGM_NCCL_PRIO") == "1":   # experiment: NCCL kernels on a high-priority stream
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    if args.config == "vqvae":
        import bench_secondary
        line = bench_secondary.run_vqvae(args, dev, rank, world)
        if rank == 0:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        if world > 1:
            dist.destroy_process_group()
        return
    cfg = CONFIGS[args.config]
    B, H, W, MULTS = args.batch, cfg["H"], cfg["W"], cfg["mults"]
    UNIT = unit_of(B)

    torch.manual_seed(0)   # identical initial weights on every rank (= the reference's default init, seed 0)
    dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
    model = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T,
                          loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
    unet, gd = model.denoising_model, model.diffusion_model
    opt = model.configure_optimizers()
    torch.manual_seed(1234 + rank)   # per-rank t / noise streams
    n_host = 8
    host = [synth_batch(B, 100 * rank + i, H, W).pin_memory() for i in range(n_host)]
    dev_batches = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def train_step(i):
        opt.zero_grad()
        loss = gd(dev_batches[i % n_host])
        loss.backward()
        opt.step()
        return loss

    def e2e_step(i):
        x = host[i % n_host].to(dev, non_blocking=True)
        opt.zero_grad()
        loss = model.training_step((x, None), i)    # includes loss.item(): the device->host read of the result
        loss.backward()
        opt.step()
        return loss

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = unet.launch_count()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), unet.launch_count() - l0

    with ClockSampler(local) as clk:
        ms, launches = timed(train_step, args.steps, args.warmup)
    ms_step = ms / args.steps
    value = world * args.steps / (ms / 1e3)
    ms_e2e, _ = timed(e2e_step, args.steps, max(3, args.warmup // 2))
    e2e_value = world * args.steps / (ms_e2e / 1e3)

    # ---- the same loop for >= 2.5 s: SM clocks settle below boost on a long dense run -------------------------------
    sustained = None
    if args.sustain_s > 0:
        n_sus = max(args.steps, int(math.ceil(args.sustain_s * 1e3 / ms_step)))
        with ClockSampler(local) as clk_s:
            ms_s, _ = timed(train_step, n_sus, 3)
        sustained = {"value": world * n_sus / (ms_s / 1e3), "unit": UNIT, "steps": n_sus, "seconds": ms_s / 1e3,
                     "ms_per_step": ms_s / n_sus, "clocks": clk_s.summary()}

    # ---- sampler: T-step reverse diffusion, one C call, CUDA-graph replay ------------
    s_steps = T if args.sample_steps <= 0 else args.sample_steps
    img = torch.randn(B, CH, H, W, device=dev)
    gd._run_sampler(img, T - 1, min(20, T), seed=1)          # warm-up + graph capture
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = unet.launch_count()
    with ClockSampler(local) as clk_smp:
        e0.record()
        out = gd._run_sampler(img, T - 1, s_steps, seed=2)
        e1.record()
        barrier()
    s_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(s_ms, op=dist.ReduceOp.MAX)
    s_ms = s_ms.item()
    s_launches = unet.launch_count() - l0
    samples_per_sec = world * B / (s_ms / 1e3 * (T / s_steps))
    finite = bool(torch.isfinite(out).all())
    # end to end through the public call: GaussianDiffusion.sample(B) (draws x_T, runs the chain) + the images on the host
    host_img = torch.empty(B, CH, H, W).pin_memory()
    s_e2e = None
    if s_steps == T:
        barrier()
        t0 = time.perf_counter()
        e0.record()
        fake = gd.sample(B)
        host_img.copy_(fake, non_blocking=True)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        s_e2e = {"value": world * B / (ms2.item() / 1e3), "unit": "samples/s", "ms": ms2.item(), "wall_ms": wall * 1e3,
                 "h2d_bytes": 0, "d2h_bytes": B * CH * H * W * 4,
                 "api": "GaussianDiffusion.sample(B) + copy of the images to pinned host memory",
                 "finite": bool(torch.isfinite(host_img).all())}

    # ---- roofline pass (event pair around every launch; outside the timed region) ------
    roof = None
    prof = None
    if rank == 0:
        P = peaks()
        unet.ddp_sync = False      # rank 0 profiles alone: no collective may be issued here
        unet.profile_start()
        for i in range(3):
            train_step(i)
        prof = unet.profile_stop()
        unet.ddp_sync = True
        tot = sum(v["ms"] for v in prof.values()) or 1.0
        # dominant kernel = conv_tc_kernel: every forward conv and every data-gradient conv is a launch of it (the
        # fprop + dgrad scopes; they also hold the two SIMT stem convs and the SIMT 1x1 head dgrad, < 3 % of their time)
        d = {k: prof["conv_fprop"][k] + prof["conv_dgrad"][k] for k in ("ms", "flops", "bytes", "launches")}
        ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
        traffic, tnote = None, "no ncu DRAM capture committed for this build"
        for tp in ("r2_conv_tc_dram.json", "r1final_conv_tc_dram.json"):
            tp = os.path.join(ROOT, "profiles", tp)
            if os.path.exists(tp) and args.config == "cifar10":
                tj = json.load(open(tp))
                traffic, tnote = tj["avg_dram_bytes_per_launch"], tj["note"]
                break
        # each launch is event-timed alone at boost clocks: the burst peak is the matching denominator
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (fprop + dgrad launches)", "achieved": ach,
                "peak": P["tf_burst"], "unit": "TFLOP/s",
                "frac": ach / P["tf_burst"], "traffic": traffic,
                "peak_source": P["src"] + " bf16 dense burst (kernels timed alone, one event pair per launch)",
                "avg_launch_ms": d["ms"] / max(d["launches"], 1), "share_of_step": d["ms"] / tot,
                "algorithmic_flops_per_launch": d["flops"] / max(d["launches"], 1),
                "algorithmic_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
                "frac_of_bf16x3_ceiling": ach / (P["tf_burst"] / 3.0),
                "traffic_note": tnote,
                "engine": "simt-fp32" if unet._engine.lib.igm_get_conv_engine(unet._engine.ctx) == 0 else "tcgen05-bf16x3",
                "classes": {k: {"ms_per_step": v["ms"] / 3, "launches_per_step": v["launches"] // 3,
                                "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0,
                                "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0}
                            for k, v in prof.items() if v["launches"]},
                "whole_step": {"achieved": B * cfg["train_gflop"] / ms_step, "unit": "TFLOP/s",
                               "frac": B * cfg["train_gflop"] / ms_step / P["tf_sust"], "peak": P["tf_sust"]},
                "sampler": {"achieved": samples_per_sec / world * T * cfg["fwd_gflop"] / 1e3, "unit": "TFLOP/s",
                            "frac": samples_per_sec / world * T * cfg["fwd_gflop"] / 1e3 / P["tf_sust"], "peak": P["tf_sust"]}}

    eager = None
    secondary = {}
    cpu = None
    if rank == 0 and world == 1:
        # free our engine's activations before the eager comparator allocates its own
        if not args.no_eager:
            eager = eager_gpu(args.config, B, dev)
            eager["speedup_train_vs_fp32"] = value / eager["fp32"]["train_steps_per_sec"]
            eager["speedup_train_vs_tf32"] = value / eager["tf32"]["train_steps_per_sec"]
            eager["speedup_sampler_vs_fp32"] = samples_per_sec / eager["fp32"]["samples_per_sec_1000step_extrapolated"]
            eager["speedup_sampler_vs_tf32"] = samples_per_sec / eager["tf32"]["samples_per_sec_1000step_extrapolated"]
        if args.config == "cifar10" and not args.no_secondary:
            import bench_secondary
            secondary = bench_secondary.run_secondary(dev, peaks(), cpu=not args.no_cpu)
        if not args.no_cpu:
            cpu = cpu_baseline(args.config, B)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.config, B),
                       "l2": "per-step working set (GBs of activations) >> 126 MB L2; 8 rotating input batches",
                       "parallelism": f"dp{world}", "global_batch": B * world},
            "images_per_sec": value * B,
            "samples_per_sec_1000step": samples_per_sec,
            "sampler": {"steps_timed": s_steps, "ms": s_ms, "ms_per_denoise_step": s_ms / s_steps,
                        "extrapolated": s_steps != T, "finite": finite, "gpu_launches": s_launches, "batch": B,
                        "clocks": clk_smp.summary(), "e2e": s_e2e},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * CH * H * W * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "H2D of the pinned batch + DDPM.training_step (enqueues fwd+bwd, reads the loss back on a copy stream) "
                           "+ loss.backward() + FusedAdam.step()"},
            "sustained": sustained,
            "gpu_launches": launches,
            "roofline": roof,
            "eager_gpu": eager,
            "cpu_baseline": cpu,
        }
        line.update(secondary)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cifar10", choices=list(CONFIGS) + ["vqvae"])
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (0 = the config's: 128 cifar10, 32 celeba64 / vqvae)")
    ap.add_argument("--sample-steps", type=int, default=0, help="denoise steps to time (0 = the full T=1000 chain)")
    ap.add_argument("--sustain-s", type=float, default=2.5, help="seconds of the sustained train loop (0 = skip)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--no-eager", action="store_true", help="skip the PyTorch-eager-on-this-GPU comparator")
    ap.add_argument("--no-secondary", action="store_true", help="skip the PixelCNN (M3) and VQ / VQ-VAE (M4) legs")
    args = ap.parse_args()
    if args.batch <= 0:
        args.batch = 32 if args.config == "vqvae" else CONFIGS[args.config]["batch"]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
