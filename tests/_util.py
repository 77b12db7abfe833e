"""Shared helpers for the parity tests."""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES

REL_TOL = 1e-3   # north_star: outputs within 1e-3 rel fp32 of the reference path


def golden(case):
    return dict(np.load(os.path.join(GOLDEN, f"ddpm_{case}.npz")))


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """(rel-L2, max-abs / max-ref) of a against reference b (SURVEY.md section 8(c) definition)."""
    a = a.detach().double().cpu().reshape(-1)
    b = b.detach().double().cpu().reshape(-1)
    l2 = (a - b).norm().item() / max(b.norm().item(), 1e-30)
    mx = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    return l2, mx


def assert_close(a, b, what, tol=REL_TOL):
    l2, mx = rel_err(torch.as_tensor(a), torch.as_tensor(b))
    assert l2 <= tol and mx <= tol, f"{what}: rel-L2 {l2:.3e}, max-rel {mx:.3e} exceed {tol:g}"
    return l2, mx
