#!/bin/bash
mkdir -p gpurun_out
echo "== vq + pixelcnn" ; timeout 900 python -m pytest tests/test_vq.py tests/test_pixelcnn.py -q -m gpu --tb=short > gpurun_out/pytest_new.log 2>&1 ; echo "rc=$?" ; tail -40 gpurun_out/pytest_new.log
python - <<'PY' 2>&1 | tail -5
import time, torch, sys
sys.path.insert(0, '.')
import igm_b200
from types import SimpleNamespace
dm = SimpleNamespace(width=28, height=28, channels=1, transforms=SimpleNamespace(normalize=False))
torch.manual_seed(0)
m = igm_b200.PixelCNN(dm, hidden_dim=64).cuda()
m.sample((64, 1, 28, 28), seed=1); torch.cuda.synchronize()
t0 = time.time(); m.sample((64, 1, 28, 28), seed=2); torch.cuda.synchronize(); dt = time.time() - t0
print(f"pixelcnn sample B=64 28x28: {dt*1e3:.1f} ms -> {64/dt:.1f} samples/s")
vq = igm_b200.VectorQuantizer(512, 64, 0.25).cuda()
z = torch.randn(32, 64, 32, 32, device='cuda')
vq(z); torch.cuda.synchronize()
t0 = time.time()
for _ in range(10): vq(z)
torch.cuda.synchronize(); dt = (time.time() - t0) / 10
print(f"vq forward 32768 vectors x 512 codes: {dt*1e3:.3f} ms -> {32768/dt/1e6:.1f} Mvec/s")
PY
