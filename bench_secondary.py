"""Secondary metrics of bench.py: M3 PixelCNN MNIST sampling at batch 64 (BASELINE.json
configs[3]) and M4 VQ lookup vectors/s + VQ-VAE 128x128 training steps/s at 32 images per GPU (configs[4]).

Each figure is reported three ways: this framework (CUDA events), the reference algorithm as PyTorch eager on the same
GPU (the oracle functions on cuda, TF32 off), and a bounded sample of the same algorithm on the host cores.
`python bench.py --config vqvae --gpus N` runs the VQ-VAE step data-parallel (flat-arena all-reduce) as its own line.
"""
import os
import time
from types import SimpleNamespace

import torch

PIX_FWD_GFLOP = 2.077          # one full PixelCNN forward per sample = the incremental sampler's work
PIX_REF_GFLOP = 843.5          # the reference algorithm: 784 cropped forwards per sample
VQVAE_FWD_GFLOP = 2.11         # encoder + decoder forward per 128x128 sample
FP32_FMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12   # nominal: 148 SMs x 128 FMA lanes x 2 FLOP x 1.965 GHz


def _dm(c, h, w, normalize):
    return SimpleNamespace(width=w, height=h, channels=c, transforms=SimpleNamespace(normalize=normalize))


def _events():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def _time_ms(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


# ---------------------------------------------------------------------------
# M3: PixelCNN MNIST 28x28 sampling, batch 64
# ---------------------------------------------------------------------------
def pixelcnn_leg(dev, cpu=True, N=64, H=28, W=28, hidden=64):
    import igm_b200
    from oracle import pixelcnn_oracle as PO

    params = PO.init_params(1, hidden, seed=0)
    model = igm_b200.PixelCNN(_dm(1, H, W, False), hidden_dim=hidden).to(dev)
    model.load_state_dict(params)
    shape = (N, 1, H, W)
    seed = [10]

    def ours():
        seed[0] += 1
        return model.sample(shape, seed=seed[0])

    ms = _time_ms(ours, 5)
    host = torch.empty(shape).pin_memory()
    torch.cuda.synchronize()
    e0, e1 = _events()
    e0.record()
    img = ours()
    host.copy_(img, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    ok = bool(((host >= 0) & (host <= 1)).all()) and float(host.std()) > 0

    out = {
        "metric": "pixelcnn_mnist_samples_per_sec", "value": N / (ms / 1e3), "unit": "samples/s", "batch": N,
        "ms_per_batch": ms, "workload": "PixelCNN MNIST 28x28 hidden 64, full autoregressive sample, batch 64 "
                                        "(BASELINE.json configs[3]); in-kernel Philox draws",
        "e2e": {"value": N / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes": 0, "d2h_bytes": N * H * W * 4,
                "api": "PixelCNN.sample(shape) + copy of the images to pinned host memory", "valid_pixels": ok},
        "roofline": {"bound": "fp32-fma issue (per-pixel GEMVs on CUDA cores; decisions must reproduce fp32 argmax / CDF)",
                     "achieved": N * PIX_FWD_GFLOP / ms, "unit": "TFLOP/s", "peak": FP32_FMA_TFLOPS,
                     "frac": N * PIX_FWD_GFLOP / ms / FP32_FMA_TFLOPS, "peak_source": "nominal 148 SMs x 128 FMA x 2 x 1.965 GHz",
                     "algorithmic_gflop_per_sample": PIX_FWD_GFLOP,
                     "reference_algorithm_gflop_per_sample": PIX_REF_GFLOP},
    }

    # the reference algorithm (784 cropped full forwards, one host sync per pixel) as PyTorch eager on this GPU
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        pd = {k: v.to(dev) for k, v in params.items()}
        u = torch.rand(H * W, N, device=dev)
        PO.sample(pd, (N, 1, 4, W), u[: 4 * W], False, img=torch.zeros(N, 1, 4, W, device=dev) - 1)   # warm-up (4 rows)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        PO.sample(pd, shape, u, False, img=torch.zeros(shape, device=dev) - 1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["eager_gpu"] = {"value": N / dt, "unit": "samples/s", "seconds_per_batch": dt,
                            "impl": "oracle/pixelcnn_oracle.py::sample on cuda (the reference loop: a cropped full forward and a "
                                    "host sync per pixel), TF32 off", "speedup": (N / (ms / 1e3)) / (N / dt)}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved

    if cpu:
        # bounded sample: the per-pixel cropped forward at rows 4 / 14 / 28, linear in the row count, summed over 784 pixels
        torch.set_num_threads(os.cpu_count() or 1)
        rows = (4, 14, 28)
        ts = []
        with torch.no_grad():
            for r in rows:
                x = torch.rand(N, 1, r, W)
                PO.forward(params, x)
                t0 = time.perf_counter()
                for _ in range(2):
                    PO.forward(params, x)
                ts.append((time.perf_counter() - t0) / 2)
        # least-squares line t(r) = a + b r
        n = len(rows)
        mr, mt = sum(rows) / n, sum(ts) / n
        b = sum((r - mr) * (t - mt) for r, t in zip(rows, ts)) / sum((r - mr) ** 2 for r in rows)
        a = mt - b * mr
        total = sum(W * (a + b * (h + 1)) for h in range(H))
        out["cpu_baseline"] = {"value": N / total, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                               "sample": f"cropped full forward (oracle/pixelcnn_oracle.py) at {rows} rows, batch {N}, 2 runs each; "
                                         "extrapolated over the 784 pixels of the reference loop (softmax / draw excluded)",
                               "seconds_per_batch_extrapolated": total}
    return out


# ---------------------------------------------------------------------------
# M4: VQ lookup and the VQ-VAE training step (CelebA 128x128, K = 512, D = 64, 32 images per GPU)
# ---------------------------------------------------------------------------
def _vqvae_torch_step(dev, B, S):
    """One training step of the reference algorithm (oracle functions + torch.optim.Adam) on `dev`."""
    from oracle import vqvae_oracle as VO
    p0 = VO.init_params(3, 64, 512, seed=0)
    leaves, params = {}, {}
    for k, v in p0.items():                      # tied residual layers stay ONE leaf
        if id(v) not in leaves:
            leaves[id(v)] = torch.nn.Parameter(v.to(dev))
        params[k] = leaves[id(v)]
    opt = torch.optim.Adam(list(leaves.values()), lr=2e-4, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev)

    def step():
        opt.zero_grad()
        total = VO.training_losses(params, x, 0.25)[0]
        total.backward()
        opt.step()
        return total

    return step


def build_vqvae(dev, S=128):
    import igm_b200
    from oracle import vqvae_oracle as VO
    model = igm_b200.VQVAE(_dm(3, S, S, True), latent_dim=64, num_embeddings=512, beta=0.25, lr=2e-4, b1=0.5, b2=0.999).to(dev)
    model.load_state_dict(VO.init_params(3, 64, 512, seed=0))
    return model


def vq_leg(dev, P, cpu=True, B=32, S=128, K=512, D=64):
    import igm_b200
    from oracle import vq_oracle

    out = {}
    # ---- lookup: [B, 64, 32, 32] latents against the 512 x 64 codebook -------------------------------------------
    hw = S // 4
    vq = igm_b200.VectorQuantizer(K, D, 0.25).to(dev)
    n_vec = B * hw * hw
    g = torch.Generator().manual_seed(3)
    zs = [torch.randn(B, D, hw, hw, generator=g).mul_(1.0 / K).to(dev) for _ in range(20)]   # 20 x 8.4 MB > L2
    i = [0]

    def look():
        i[0] += 1
        with torch.no_grad():
            return vq(zs[i[0] % len(zs)])

    ms = _time_ms(look, 40, 5)
    # the same lookup straight through the C ABI with preallocated outputs: the public call above is bound by its host
    # side (four torch allocations + the autograd.Function per call), this is the kernel pair itself
    import ctypes as C
    from igm_b200 import _lib
    lib = _lib.load()
    emb_c = vq.embedding.detach().contiguous()
    idx = torch.empty(n_vec, dtype=torch.int64, device=dev)
    quant = torch.empty_like(zs[0])
    losses = torch.empty(2, device=dev)
    wsb = torch.empty(lib.igm_vq_workspace_floats(B, hw * hw), device=dev)
    P_ = lambda t: C.c_void_p(t.data_ptr())
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def raw():
        i[0] += 1
        rc = lib.igm_vq_forward(P_(zs[i[0] % len(zs)]), P_(emb_c), P_(idx), P_(quant), P_(losses), B, D, hw * hw, K, 0.25, P_(wsb), stream)
        assert rc == 0

    ms_k = _time_ms(raw, 200, 10)
    alg_bytes = n_vec * (2 * D * 4 + 8) + K * D * 4
    out["vq_lookup"] = {
        "metric": "vq_lookup_vectors_per_sec", "value": n_vec / (ms / 1e3), "unit": "vectors/s", "ms": ms,
        "workload": f"VectorQuantizer.forward on [{B},{D},{hw},{hw}] latents, K={K} (per-GPU share of BASELINE.json configs[4]); "
                    "20 rotating inputs (168 MB > L2); through the public module call (host-bound: allocations + autograd.Function)",
        "c_abi": {"value": n_vec / (ms_k / 1e3), "unit": "vectors/s", "ms": ms_k,
                  "api": "igm_vq_forward with preallocated outputs, back to back"},
        "roofline": {"bound": "fp32-fma (distance GEMM on CUDA cores: the argmin must reproduce the fp32 reference)",
                     "achieved": n_vec * K * D * 2 / (ms_k * 1e-3) / 1e12, "unit": "TFLOP/s", "peak": FP32_FMA_TFLOPS,
                     "frac": n_vec * K * D * 2 / (ms_k * 1e-3) / 1e12 / FP32_FMA_TFLOPS,
                     "peak_source": "nominal 148 SMs x 128 FMA x 2 x 1.965 GHz", "timed": "c_abi",
                     "algorithmic_bytes": alg_bytes, "hbm_gbs": alg_bytes / (ms_k * 1e-3) / 1e9, "hbm_frac": alg_bytes / (ms_k * 1e-3) / 1e9 / P["hbm"]},
    }
    emb = vq.embedding.detach()

    def look_eager():
        i[0] += 1
        with torch.no_grad():
            return vq_oracle.vq_forward(zs[i[0] % len(zs)], emb, 0.25)

    ms_e = _time_ms(look_eager, 20, 3)
    out["vq_lookup"]["eager_gpu"] = {"value": n_vec / (ms_e / 1e3), "unit": "vectors/s", "ms": ms_e,
                                     "impl": "oracle/vq_oracle.py::vq_forward on cuda (torch.cdist + argmin + gather + 2 mse)",
                                     "speedup": ms_e / ms}
    del zs

    # ---- VQ-VAE training step -----------------------------------------------------------------------------------
    model = build_vqvae(dev, S)
    opt = model.configure_optimizers()
    g = torch.Generator().manual_seed(0)
    x = (torch.rand(B, 3, S, S, generator=g) * 2 - 1).to(dev)

    def step():
        opt.zero_grad()
        loss = model.training_step((x, None), 0)
        loss.backward()
        opt.step()
        return loss

    ms_t = _time_ms(step, 10, 3)
    out["vqvae_train"] = {
        "metric": "vqvae_train_steps_per_sec", "value": 1e3 / ms_t, "unit": f"steps/s ({B} images per GPU-step)", "ms_per_step": ms_t,
        "workload": f"VQVAE.training_step + backward + Adam on [{B},3,{S},{S}] (BASELINE.json configs[4]: 128 images over 4 GPUs)",
        "tflops": 3 * B * VQVAE_FWD_GFLOP / ms_t,
    }
    del model, opt
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        est = _vqvae_torch_step(dev, B, S)
        ms_te = _time_ms(est, 6, 3)
        out["vqvae_train"]["eager_gpu"] = {"value": 1e3 / ms_te, "ms_per_step": ms_te, "speedup": ms_te / ms_t,
                                           "impl": "oracle/vqvae_oracle.py::training_losses + torch.optim.Adam on cuda, TF32 off"}
        del est
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    if cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        cst = _vqvae_torch_step("cpu", B, S)
        cst()
        t0 = time.perf_counter()
        for _ in range(2):
            cst().item()
        dt = (time.perf_counter() - t0) / 2
        out["vqvae_train"]["cpu_baseline"] = {"value": 1.0 / dt, "unit": f"steps/s ({B} images per step)", "cores": torch.get_num_threads(),
                                              "kind": "port", "sample": f"2 train steps (+1 warm-up) at B={B}, oracle/vqvae_oracle.py on CPU"}
    return out


def run_secondary(dev, P, cpu=True):
    out = {}
    try:
        out["pixelcnn"] = pixelcnn_leg(dev, cpu)
    except Exception as e:   # a secondary leg must not cost the headline line
        out["pixelcnn"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    try:
        out["vqvae"] = vq_leg(dev, P, cpu)
    except Exception as e:
        out["vqvae"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


# ---------------------------------------------------------------------------
# `bench.py --config vqvae`: the VQ-VAE training step, data parallel (BASELINE.json configs[4])
# ---------------------------------------------------------------------------
VQ_METRIC = "vqvae_train_steps_per_sec"


def _vq_workload(B):
    return (f"VQ-VAE CelebA 128x128 K=512 D=64 batch={B}/GPU (BASELINE.json configs[4]: 128 images over 4 GPUs); "
            "train step = encoder + VQ + decoder + losses + backward + allreduce + Adam")


def run_vqvae(args, dev, rank, world):
    import torch.distributed as dist
    B, S = args.batch, 128
    torch.manual_seed(0)
    model = build_vqvae(dev, S)
    model.sync_parameters()
    opt = model.configure_optimizers()
    host = [(torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(100 * rank + i)) * 2 - 1).pin_memory() for i in range(8)]
    dev_b = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i, e2e=False):
        x = host[i % 8].to(dev, non_blocking=True) if e2e else dev_b[i % 8]
        opt.zero_grad()
        loss = model.training_step((x, None), i)
        loss.backward()
        model.on_after_backward()           # Lightning's hook: the data-parallel gradient exchange
        opt.step()
        if e2e:
            loss.item()
        return loss

    from igm_b200 import _lib
    lib = _lib.load()
    counted = {}

    def timed(steps, warmup, e2e):
        for i in range(warmup):
            step(i, e2e)
        barrier()
        n0 = int(lib.igm_ops_launch_count())
        e0, e1 = _events()
        e0.record()
        for i in range(steps):
            step(i, e2e)
        e1.record()
        barrier()
        counted["launches"] = int(lib.igm_ops_launch_count()) - n0   # this rank's kernels inside the timed region
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    from bench import ClockSampler, peaks
    clk = ClockSampler(dev.index or 0)
    with clk:
        ms = timed(args.steps, args.warmup, False)
    launches = counted["launches"]
    ms_e = timed(args.steps, 3, True)
    unit = f"steps/s ({B} images per GPU-step)"
    P = peaks()
    tf = 3 * B * VQVAE_FWD_GFLOP / (ms / args.steps)          # fwd + data-gradient + weight-gradient passes, TFLOP/s
    extra = {}
    if world == 1 and not getattr(args, "no_cpu", False):      # bounded CPU sample: 2 steps of the oracle on the host cores
        torch.set_num_threads(os.cpu_count() or 1)
        cst = _vqvae_torch_step("cpu", B, S)
        cst()
        t0 = time.perf_counter()
        for _ in range(2):
            cst().item()
        dt = (time.perf_counter() - t0) / 2
        extra["cpu_baseline"] = {"value": 1.0 / dt, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"2 train steps (+1 warm-up) at B={B}, oracle/vqvae_oracle.py (torch CPU fp32)"}
    return {
        **extra,
        "clocks": clk.summary(),
        "roofline": {"bound": "tensor", "kernel": "whole step (tcgen05 bf16x3 convs via igm_conv2d_*; 3- and 32-channel layers on CUDA cores)",
                     "achieved": tf, "peak": P["tf_burst"], "unit": "TFLOP/s", "frac": tf / P["tf_burst"], "traffic": None,
                     "peak_source": f"{P['src']} bf16 dense burst; algorithmic 3 x {VQVAE_FWD_GFLOP} GFLOP per sample and step"},
        "metric": VQ_METRIC, "value": world * args.steps / (ms / 1e3), "unit": unit, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": _vq_workload(B), "parallelism": f"dp{world}", "global_batch": B * world,
                   "l2": "activations of one step (~2 GB) >> L2; 8 rotating input batches"},
        "e2e": {"value": world * args.steps / (ms_e / 1e3), "unit": unit, "h2d_bytes_per_step": B * 3 * S * S * 4,
                "d2h_bytes_per_step": 4, "api": "H2D of the pinned batch + VQVAE.training_step + backward + on_after_backward + Adam + loss.item()"},
        "gpu_launches": launches * world,
    }


def run_reference_vqvae(args):
    import json
    B = args.batch
    st = _vqvae_torch_step("cpu", B, 128)
    for _ in range(args.warmup):
        st().item()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st().item()
    dt = (time.perf_counter() - t0) / max(args.steps, 1)
    unit = f"steps/s ({B} images per GPU-step)"
    print(json.dumps({
        "impl": "reference", "metric": VQ_METRIC, "value": 1.0 / dt, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": _vq_workload(B), "where": "host CPU only, one rank"},
        "cpu_baseline": {"value": 1.0 / dt, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{args.steps} train steps at B={B}, oracle/vqvae_oracle.py"},
        "e2e": {"value": 1.0 / dt, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
