"""Bring-up probe for the tcgen05 wgrad kernel: tries both MN-major descriptor conventions."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from igm_b200 import _lib
lib = _lib.load()
for shape in [(2, 32, 32, 64, 64, 3), (2, 16, 16, 128, 128, 3), (2, 8, 8, 128, 64, 1)]:
    B, H, W, Cin, Cout, K = shape
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, Cin, generator=g); dy = torch.randn(B, H, W, Cout, generator=g)
    w = torch.zeros(Cout, Cin, K, K, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2).double(), w, padding=(K - 1) // 2)
    ref, = torch.autograd.grad(y, w, dy.permute(0, 3, 1, 2).double())
    for variant in (0, 1):
        gw = torch.zeros(Cout, Cin, K, K, device="cuda")
        p = lambda t: C.c_void_p(t.data_ptr())
        rc = lib.igm_debug_wgrad(1, variant, p(x.cuda()), p(dy.cuda()), p(gw), B, H, W, Cin, Cout, K, None)
        torch.cuda.synchronize()
        e = (gw.cpu().double() - ref).norm() / ref.norm()
        print(shape, "variant", variant, "rc", rc, "rel-L2", float(e), flush=True)
