"""Golden fixtures of the VQ-VAE model path from the UNMODIFIED reference
(src/models/vqvae.py + src/networks/vqvae.py), run here on CPU.

    python tests/golden/make_golden_vqvae.py

Weights come from oracle.vqvae_oracle.init_params and are loaded into the reference module; the fixture keeps
the losses of one training_step, every parameter gradient (full for the small case, sub-sampled otherwise), the
chosen codes and a sub-sampled reconstruction.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import vqvae_oracle as VO  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    # name: dict(channels, size, latent_dim, K, h_dim, res_h_dim, n_res, N, codebook_scale, normalize)
    "small": dict(C=3, S=16, D=16, K=32, h=32, rh=16, n_res=2, N=4, cs=0.3, norm=True),
    "cifar": dict(C=3, S=32, D=64, K=512, h=128, rh=128, n_res=3, N=2, cs=None, norm=True),
    "mnist": dict(C=1, S=28, D=32, K=64, h=64, rh=32, n_res=3, N=3, cs=0.2, norm=False),
}


def build_inputs(c):
    g = torch.Generator().manual_seed(5)
    x = torch.rand(c["N"], c["C"], c["S"], c["S"], generator=g)
    return x * 2 - 1 if c["norm"] else x


def params_of(c, seed=3):
    return VO.init_params(c["C"], c["D"], c["K"], c["h"], c["rh"], c["n_res"], seed=seed, codebook_scale=c["cs"])


def reference_model(c):
    ref = ref_loader.load("vqvae")
    enc = {"_target_": "src.networks.vqvae.Encoder", "n_res_layers": c["n_res"], "res_h_dim": c["rh"]}
    dec = {"_target_": "src.networks.vqvae.Decoder", "h_dim": c["h"], "n_res_layers": c["n_res"], "res_h_dim": c["rh"]}
    m = ref.VQVAE(ref_loader.datamodule_cfg(c["C"], c["S"], c["S"], normalize=c["norm"]), encoder=enc, decoder=dec,
                  latent_dim=c["D"], num_embeddings=c["K"], beta=0.25)
    m.load_state_dict(params_of(c))
    return m


def sub(t, n=4096):
    f = t.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].numpy().copy()


def main():
    for name, c in CASES.items():
        m = reference_model(c)
        x = build_inputs(c)
        total = m.training_step((x, None), 0)
        total.backward()
        with torch.no_grad():
            recon = m(x)
            ez = m.encoder(x.clone())
            idx = torch.argmin(torch.cdist(ez.reshape(c["N"], c["D"], -1).permute(0, 2, 1).reshape(-1, c["D"]),
                                           m.vector_quntizer.embedding), dim=1)
        out = {"total": np.float32(total.item()), "recon_loss": np.float32(m._logged["train_loss/recon_loss"].item()),
               "vq_loss": np.float32(m._logged["train_loss/vq_loss"].item()),
               "commit_loss": np.float32(m._logged["train_loss/commit_loss"].item()),
               "indices": idx.numpy().astype(np.int64), "recon_sub": sub(recon), "encoder_z_sub": sub(ez)}
        seen = set()
        for k, p in m.named_parameters():           # tied layers are reported once (stack.0)
            if id(p) in seen:
                continue
            seen.add(id(p))
            out["grad:" + k] = p.grad.numpy().copy() if name == "small" else sub(p.grad)
        np.savez_compressed(os.path.join(HERE, f"vqvae_{name}.npz"), **out)
        print(name, {k: float(out[k]) for k in ("total", "recon_loss", "vq_loss", "commit_loss")}, "codes used",
              len(set(idx.tolist())))


if __name__ == "__main__":
    main()
