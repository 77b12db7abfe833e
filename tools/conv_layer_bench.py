"""Per-layer timing of the 3x3 / 1x1 convolutions of the DDPM U-Nets on the two tcgen05 engines (GPU tool, not a test):
engine 1 = per-tap conv_tc.cu, engine 3 = CTA pair conv_tc2.cu (comparison only).  Uses the C-ABI timing entry
igm_debug_conv_bench (operands staged once, CUDA events around `iters` back-to-back launches, warm L2).

    python tools/conv_layer_bench.py [--pair] [--batch 128] [--iters 50]

Prints one line per (layer shape, mode, engine): microseconds per launch and fp32-equivalent TFLOP/s, next to the
bf16x3 ceiling (measured bf16 dense peak / 3)."""
import argparse
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (H, W, Cin, Cout, K, launches per forward pass) of the stride-1 convs of the CIFAR-10 and CelebA-64 topologies
CIFAR = [(32, 32, 64, 64, 3, 4), (32, 32, 64, 384, 1, 1), (32, 32, 128, 64, 1, 2),
         (16, 16, 64, 128, 3, 1), (16, 16, 128, 128, 3, 6), (16, 16, 256, 64, 3, 1), (16, 16, 64, 64, 3, 3),
         (8, 8, 128, 256, 3, 1), (8, 8, 256, 256, 3, 7), (8, 8, 512, 128, 3, 1), (8, 8, 128, 128, 3, 3)]
CELEBA = [(64, 64, 64, 64, 3, 4), (32, 32, 64, 128, 3, 1), (32, 32, 128, 128, 3, 6), (32, 32, 256, 64, 3, 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--pair", action="store_true", help="also time engine 3 where the shape is eligible")
    ap.add_argument("--celeba", action="store_true")
    args = ap.parse_args()
    import torch
    from igm_b200 import _lib
    lib = _lib.load()
    torch.cuda.init()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    except OSError:
        pass
    ceil = peaks.get("bf16_tflops_sustained", 1378.6) / 3.0
    shapes = CELEBA if args.celeba else CIFAR
    B = 32 if args.celeba else args.batch
    ms = C.c_float(0)
    for (H, W, Cin, Cout, K, n) in shapes:
        for mode in (0, 1):
            for engine in [1] + ([3] if args.pair else []):
                rc = lib.igm_debug_conv_bench(engine, mode, B, H, W, Cin, Cout, K, 1 if (mode == 0 and K == 3) else 0, 5,
                                              args.iters, C.byref(ms), None)
                if rc != 0:
                    if engine == 1:
                        print(f"{H}x{W} {Cin}->{Cout} k{K} mode {mode} engine {engine}: {lib.igm_last_error(None).decode()}")
                    continue
                fl = 2.0 * B * H * W * Cin * Cout * K * K
                tf = fl / (ms.value * 1e-3) / 1e12
                print(f"{H:3d}x{W:<3d} {Cin:4d}->{Cout:<4d} k{K} {'fprop' if mode == 0 else 'dgrad'} engine {engine}: "
                      f"{ms.value * 1e3:8.1f} us  {tf:7.1f} TFLOP/s  ({100 * tf / ceil:5.1f} % of the bf16x3 ceiling)  x{n}/pass")


if __name__ == "__main__":
    main()
