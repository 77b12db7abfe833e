#!/usr/bin/env python
"""Headline benchmark: DDPM CIFAR-10 32x32 U-Net (dims 3->64->128->256, T=1000), batch 128 per GPU.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, via the C ABI)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host CPU

One JSON line on stdout (rank 0).  A "step" is one training step of BASELINE.json's
configs[1] workload: noise + q_sample + U-Net forward + L1 loss + backward +
[gradient all-reduce] + Adam, on one synthetic batch of 128 images per GPU.
`value` = rank-steps per second over all GPUs (weak scaling: 128 images per GPU per step).
The same line carries the 1000-step sampler throughput (samples/s), the end-to-end
number through DDPM.training_step with pinned-host inputs, the roofline of the
dominant kernel class and the CPU baseline timed on this box.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

DIM, CH, MULTS, H, W, T = 64, 3, (1, 2, 4), 32, 32, 1000
TRAIN_GFLOP_PER_SAMPLE = 5.153   # BASELINE.md section 3 (fwd+bwd, FlopCounterMode on the reference)
FWD_GFLOP_PER_SAMPLE = 1.719
METRIC = "ddpm_unet_train_steps_per_sec"
UNIT = "steps/s (128 images per GPU-step)"


def workload_name(B):
    """config.workload, shared verbatim by our arm and the reference arm (BASELINE.json configs[1])."""
    return (f"DDPM CIFAR-10 32x32 U-Net dims 3-64-128-256 T=1000 batch={B}/GPU "
            "(BASELINE.json configs[1]); train step = noise+q_sample+fwd+L1+bwd+allreduce+Adam")



def synth_batch(B, seed):
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
            self._stop.wait(0.1)

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
# CPU arm: the oracle port of the reference algorithm on the host cores
# ---------------------------------------------------------------------------
def cpu_train_setup(B):
    from oracle import ddpm_oracle as O
    spec = O.UnetSpec(DIM, CH, MULTS)
    params = {k: torch.nn.Parameter(v) for k, v in O.init_params(spec, seed=0).items()}
    opt = torch.optim.Adam(params.values(), lr=1e-4, betas=(0.9, 0.999))
    buf = O.diffusion_buffers(T)
    x = synth_batch(B, 0)
    g = torch.Generator().manual_seed(1)

    def step():
        t = torch.randint(0, T, (B,), generator=g)
        noise = torch.randn(B, CH, H, W, generator=g)
        opt.zero_grad()
        loss = O.p_losses(params, spec, buf, x, t, noise, "l1")
        loss.backward()
        opt.step()
        return loss.item()

    def sample_step(img, i):
        with torch.no_grad():
            return O.p_sample(params, spec, buf, img, torch.full((B,), i, dtype=torch.long),
                              torch.randn(B, CH, H, W, generator=g))

    return step, sample_step


def cpu_baseline(B=128, train_steps=3, sample_steps=4):
    torch.set_num_threads(os.cpu_count() or 1)
    step, sample_step = cpu_train_setup(B)
    step()
    t0 = time.perf_counter()
    for _ in range(train_steps):
        step()
    dt = (time.perf_counter() - t0) / train_steps
    img = torch.randn(B, CH, H, W)
    img = sample_step(img, T - 1)
    t0 = time.perf_counter()
    for i in range(sample_steps):
        img = sample_step(img, T - 2 - i)
    ds = (time.perf_counter() - t0) / sample_steps
    return {
        "value": 1.0 / dt, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
        "sample": f"{train_steps} train steps (+1 warm-up) at B={B} and {sample_steps} p_sample steps at B={B}, "
                  f"oracle/ddpm_oracle.py (torch CPU fp32, oneDNN)",
        "ms_per_train_step": dt * 1e3, "ms_per_sample_step": ds * 1e3,
        "samples_per_sec_1000step_extrapolated": B / (ds * T),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    B = args.batch
    step, _ = cpu_train_setup(B)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    v = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(B), "where": "host CPU (torch fp32, all cores), one rank",
                   "parallelism": f"dp{args.gpus}", "global_batch": B * args.gpus},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} full train steps at B={B} (oracle port of the reference algorithm; "
                                   "the reference itself is torch-on-CPU and needs Lightning, absent here)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch

    torch.manual_seed(0)   # identical initial weights on every rank (= the reference's default init, seed 0)
    dm = SimpleNamespace(width=W, height=H, channels=CH, transforms=SimpleNamespace(normalize=True))
    model = igm_b200.DDPM(dm, hidden_dim=DIM, dim_mults=MULTS, timesteps=T,
                          loss_type="l1", lr=1e-4, b1=0.9, b2=0.999).to(dev)
    unet, gd = model.denoising_model, model.diffusion_model
    opt = model.configure_optimizers()
    torch.manual_seed(1234 + rank)   # per-rank t / noise streams
    n_host = 8
    host = [synth_batch(B, 100 * rank + i).pin_memory() for i in range(n_host)]
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

    # ---- sampler: T-step reverse diffusion, one C call, CUDA-graph replay ------------
    s_steps = T if args.sample_steps <= 0 else args.sample_steps
    img = torch.randn(B, CH, H, W, device=dev)
    gd._run_sampler(img, T - 1, min(20, T), seed=1)          # warm-up + graph capture
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = unet.launch_count()
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
        tp = os.path.join(ROOT, "profiles", "r1final_conv_tc_dram.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic, tnote = tj["avg_dram_bytes_per_launch"], tj["note"]
        roof = {"bound": "tensor", "kernel": "conv_tc_kernel (fprop + dgrad launches)", "achieved": ach,
                "peak": P["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / P["tf_sust"], "traffic": traffic, "peak_source": P["src"] + " bf16 dense sustained",
                "avg_launch_ms": d["ms"] / max(d["launches"], 1), "share_of_step": d["ms"] / tot,
                "algorithmic_flops_per_launch": d["flops"] / max(d["launches"], 1),
                "algorithmic_bytes_per_launch": d["bytes"] / max(d["launches"], 1),
                "frac_of_bf16x3_ceiling": ach / (P["tf_sust"] / 3.0),
                "traffic_note": tnote,
                "engine": "simt-fp32" if unet._engine.lib.igm_get_conv_engine(unet._engine.ctx) == 0 else "tcgen05-bf16x3",
                "classes": {k: {"ms_per_step": v["ms"] / 3, "launches_per_step": v["launches"] // 3,
                                "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] else 0.0,
                                "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] else 0.0}
                            for k, v in prof.items() if v["launches"]},
                "whole_step": {"achieved": B * TRAIN_GFLOP_PER_SAMPLE / ms_step, "unit": "TFLOP/s",
                               "frac": B * TRAIN_GFLOP_PER_SAMPLE / ms_step / P["tf_sust"]},
                "sampler": {"achieved": samples_per_sec / world * T * FWD_GFLOP_PER_SAMPLE / 1e3, "unit": "TFLOP/s",
                            "frac": samples_per_sec / world * T * FWD_GFLOP_PER_SAMPLE / 1e3 / P["tf_sust"]}}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(B)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(B),
                       "l2": "per-step working set ~3.5 GB of activations >> 126 MB L2; 8 rotating input batches",
                       "parallelism": f"dp{world}", "global_batch": B * world},
            "images_per_sec": value * B,
            "samples_per_sec_1000step": samples_per_sec,
            "sampler": {"steps_timed": s_steps, "ms": s_ms, "ms_per_denoise_step": s_ms / s_steps,
                        "extrapolated": s_steps != T, "finite": finite, "gpu_launches": s_launches, "batch": B},
            "clocks": clk.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": B * CH * H * W * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "api": "H2D of the pinned batch + DDPM.training_step (enqueues fwd+bwd, reads the loss back on a copy stream) "
                           "+ loss.backward() + FusedAdam.step()"},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
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
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--sample-steps", type=int, default=0, help="denoise steps to time (0 = the full T=1000 chain)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
