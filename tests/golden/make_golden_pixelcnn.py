"""Golden fixtures of the PixelCNN path from the UNMODIFIED reference (src/models/pixelcnn.py).

    python tests/golden/make_golden_pixelcnn.py

Weights come from oracle.pixelcnn_oracle.init_params(C, hidden, seed) and are loaded into the
reference module; torch.multinomial is patched with the inverse-CDF rule on seeded uniforms so the
reference's own sample() loop produces reproducible pixels.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pixelcnn_oracle as PO  # noqa: E402
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: (channels, hidden, N, H, W, normalize)
    "mnist": (1, 64, 2, 28, 28, False),
    "rgb_small": (3, 32, 2, 9, 11, True),
    "cond_small": (1, 32, 3, 8, 8, False),
}
N_CLASSES = {"cond_small": 4}                            # class_condition=True cases
SAMPLE_HW = {"mnist": (12, 28), "rgb_small": (9, 11), "cond_small": (8, 8)}   # sampled crop (full width, fewer rows)


def labels(case):
    C, Hd, N, H, W, norm = CASES[case]
    if case not in N_CLASSES:
        return None
    return torch.arange(N) % N_CLASSES[case]


def sub(t, n=256):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].numpy().copy()


def inputs(case):
    C, Hd, N, H, W, norm = CASES[case]
    g = torch.Generator().manual_seed(21)
    x = torch.randint(0, 256, (N, C, H, W), generator=g).float() / 255
    if norm:
        x = x * 2 - 1
    sh, sw = SAMPLE_HW[case]
    u = torch.rand(sh * sw, N * C, generator=g)
    return x, u


def main():
    ref = ref_loader.load("pixelcnn")
    for case, (C, Hd, N, H, W, norm) in CASES.items():
        nc = N_CLASSES.get(case)
        p = PO.init_params(C, Hd, seed=1, n_classes=nc)
        m = ref.PixelCNN(ref_loader.datamodule_cfg(C, H, W, normalize=norm), hidden_dim=Hd, class_condition=nc is not None,
                         n_classes=nc)
        m.load_state_dict(p)
        x, u = inputs(case)
        sh, sw = SAMPLE_HW[case]
        lab = labels(case)
        y = None if lab is None else torch.nn.functional.one_hot(lab, nc).float()
        with torch.no_grad():
            logits = m(x, y)
            bpd = m.calc_likelihood(x, y)
        # one training step of the reference: loss + every parameter gradient (masked taps included)
        loss = m.training_step((x, lab), 0)
        loss.backward()
        # (the last layer's vertical gate output is unused, so its cond_proj_vert* never get a gradient)
        grads = {"grad:" + k: sub(q.grad) for k, q in m.named_parameters() if q.grad is not None}
        grads["no_grad"] = np.array([k for k, q in m.named_parameters() if q.grad is None])
        state = {"i": 0}
        real = torch.multinomial

        def fake(probs, num_samples=1):
            k = PO.pick(probs, u[state["i"]])
            state["i"] += 1
            return k[:, None]

        torch.multinomial = fake
        try:
            s_u = m.sample((N, C, sh, sw), cond=y)
        finally:
            torch.multinomial = real
        s_g = PO.sample(p, (N, C, sh, sw), None, input_normalize=norm, y=y)
        np.savez_compressed(os.path.join(HERE, f"pixelcnn_{case}.npz"), logits_sub=logits[:, :, :, ::3, ::3].numpy(),
                            bpd=np.float32(bpd.item()), train_loss=np.float32(loss.item()), sample_u=s_u.numpy(),
                            sample_greedy=s_g.numpy(), **grads)
        print(case, "bpd", bpd.item(), "sampled", tuple(s_u.shape))


if __name__ == "__main__":
    main()
