"""Column algebra of csrc/conv_tc2.cu (CTA pair, cta_group::2) as an executable numpy model; not on any product path.

Assumed instruction semantics (PTX tcgen05.mma.cta_group::2, M = 256): A rows 0..127 come from CTA 0's shared memory and
128..255 from CTA 1's; B's N rows are split - CTA r supplies rows r*N/2 .. (r+1)*N/2 - 1 FROM THE SAME DESCRIPTOR OFFSET of
its own shared memory; each CTA's TMEM receives its 128 rows of D with all N columns (this is the operand partitioning
CUTLASS encodes for SM100_MMA_F16BF16_2x1SM_SS: cute/atom/mma_traits_sm100.hpp, ALayout / BLayout / CLayout with a CTA
stride of M/2, N/2, M/2).  Under that assumption the model
checks the kernel's operand staging and epilogue: CTA r stages B'_r = [w_hi[64r:64r+64] ; w_lo[64r:64r+64]],
MMA 1 = a_hi x B' (N = 256), MMA 2 = a_lo x (first 64 rows of each B'_r) (N = 128) accumulated 64 columns to the right,
epilogue channel c -> columns (c, c + 64) for c < 64 and (128 + c - 64, 192 + c - 64) otherwise.
"""
import numpy as np


def pair_tile(a_hi, a_lo, w_hi, w_lo):
    """a_*: [256, K] (rows 0..127 = CTA 0's M tile), w_*: [128, K] (one N tile).  Returns the [256, 128] output the
    epilogues of the two CTAs assemble."""
    Bp = [np.concatenate([w_hi[64 * r:64 * r + 64], w_lo[64 * r:64 * r + 64]]) for r in range(2)]   # per-CTA tile, 128 rows
    D = np.zeros((256, 256))
    # MMA 1: N = 256 -> CTA r supplies rows 0..127 of its tile
    B1 = np.concatenate([Bp[0][:128], Bp[1][:128]])
    D += a_hi @ B1.T
    # MMA 2: N = 128 -> CTA r supplies rows 0..63 of its tile (same descriptor), D address + 64 columns
    B2 = np.concatenate([Bp[0][:64], Bp[1][:64]])
    D[:, 64:192] += a_lo @ B2.T
    out = np.zeros((256, 128))
    for c0 in range(0, 128, 32):
        col = c0 if c0 < 64 else 128 + c0 - 64
        out[:, c0:c0 + 32] = D[:, col:col + 32] + D[:, col + 64:col + 96]
    return out
