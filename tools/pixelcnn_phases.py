"""Where the PixelCNN sampler's time goes (IGM_PCNN_PROF=1): cycles of the row pass, the per-pixel chain and the head + draw,
summed over the CTAs (one per image), per pixel.  Diagnosis, not a benchmark."""
import ctypes as C
import os
import sys

os.environ["IGM_PCNN_PROF"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import igm_b200  # noqa: E402
from igm_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
from bench_secondary import _dm  # noqa: E402

model = igm_b200.PixelCNN(_dm(1, 28, 28, False), hidden_dim=64).to(dev)   # default init: timing does not depend on the weights
out = (C.c_ulonglong * 8)()
model.sample((64, 1, 28, 28), seed=1)
torch.cuda.synchronize()
lib.igm_debug_pixelcnn_prof(out)
model.sample((64, 1, 28, 28), seed=1)
torch.cuda.synchronize()
lib.igm_debug_pixelcnn_prof(out)
row, chain, head, px, f0, f1, f2, f3 = (int(v) for v in out)
print(f"pixels {px}: row pass {row / px:.0f} cycles/pixel, chain {chain / px:.0f}, head + draw {head / px:.0f}  "
      f"(total {(row + chain + head) / px:.0f} cycles = {(row + chain + head) / px / 1965:.1f} us per pixel at 1965 MHz)")
print(f"chain per pixel: input fill {f0 / px:.0f}, horiz_conv GEMV {f1 / px:.0f}, gate {f2 / px:.0f}, conv1x1_2 GEMV + tail {f3 / px:.0f} cycles")
