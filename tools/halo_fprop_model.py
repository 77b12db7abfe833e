"""Index algebra of the planned halo-reuse forward / data-gradient convolution ("Next"), as an
executable numpy model.  Not a kernel and not on any product path: it pins down the tile -> pixel map the tcgen05 kernel
will use and is checked against torch's conv2d by tests/test_halo_model.py.

Layout of one CTA work item (a band of BH image rows of one image, one 64-channel K chunk):

  * shared-memory tile S: rows y0-1 .. y0+BH and columns -1 .. W of the NHWC activation, i.e. (BH+2) x (W+2) pixels of
    C channels, zero outside the image (the TMA out-of-bounds fill), flattened in padded-linear order
    s = r * (W+2) + c; two guard rows follow (read only by junk outputs);
  * GEMM M index m in [0, BH * (W+2)): output pixel (y0 + m // (W+2), m % (W+2)); columns W and W+1 are junk and are
    dropped in the epilogue (valid fraction W / (W+2));
  * tap (ky, kx) of a 3x3 stride-1 pad-1 conv reads A row  m + ky * (W+2) + kx  of S: a constant row offset, i.e. a
    shared-memory descriptor start address shifted by whole 128-byte rows (tools/desc_probe.cu shows that is legal with
    SWIZZLE_128B), so ONE tile serves all nine taps instead of nine TMA loads;
  * an MMA covers 128 consecutive m; a band needs ceil(BH * (W+2) / 128) of them per (tap, K sub-step).

Shared-memory traffic per 64-channel k-step at C = 64 drops from 9 x 32 KB of activation fills per 128-pixel tile to
(BH+2)/BH x (W+2)/W x 32 KB per 128 valid pixels (1.37x for W = 32, BH = 7).
"""
import numpy as np


def band_plan(H, W, BH):
    """(m_rows, n_mma_tiles, valid_fraction, tile_pixels) of one band."""
    PW = W + 2
    m_rows = BH * PW
    n_tiles = -(-m_rows // 128)
    return m_rows, n_tiles, (BH * W) / (n_tiles * 128.0), (BH + 2) * PW + 2 * PW


def conv3x3_halo(x_nhwc, w_oihw, BH):
    """3x3 / stride 1 / pad 1 convolution computed band by band exactly as the kernel will index it.
    x: [B, H, W, C] float, w: [Cout, C, 3, 3] float -> [B, H, W, Cout]."""
    B, H, W, C = x_nhwc.shape
    Cout = w_oihw.shape[0]
    PW = W + 2
    out = np.zeros((B, H, W, Cout), dtype=np.float64)
    wt = w_oihw.astype(np.float64).transpose(2, 3, 1, 0)          # [ky][kx][ci][co]
    for b in range(B):
        for y0 in range(0, H, BH):
            bh = min(BH, H - y0)
            # zero-filled halo tile + guard rows, padded-linear
            S = np.zeros(((BH + 2 + 2) * PW, C), dtype=np.float64)
            for r in range(BH + 2):
                y = y0 - 1 + r
                if 0 <= y < H:
                    S[r * PW + 1:r * PW + 1 + W] = x_nhwc[b, y]
            m_rows = BH * PW
            n_tiles = -(-m_rows // 128)
            acc = np.zeros((n_tiles * 128, Cout), dtype=np.float64)
            for ky in range(3):
                for kx in range(3):
                    off = ky * PW + kx                               # the descriptor shift of this tap, in rows
                    for t in range(n_tiles):
                        rows = np.arange(t * 128, (t + 1) * 128) + off
                        rows = np.minimum(rows, S.shape[0] - 1)      # rows past the guard belong to junk outputs only
                        acc[t * 128:(t + 1) * 128] += S[rows] @ wt[ky, kx]
            for m in range(m_rows):                                  # epilogue: drop the junk columns
                yy, xx = divmod(m, PW)
                if xx < W and yy < bh:
                    out[b, y0 + yy, xx] = acc[m]
    return out


def kernel_plan(W):
    """BH of csrc/conv_halo.cu: two M = 128 tiles per band."""
    return 256 // (W + 2)


def conv3x3_halo_kernel_model(x_nhwc, w_oihw, bias=None):
    """The work decomposition and the epilogue of csrc/conv_halo.cu, step by step: bands of BH = 256 // (W+2) image rows,
    two 128-row MMA tiles (the second skipped when the band's live rows fit the first), per epilogue warp (32 accumulator
    rows) one clipped row store per image row it touches, one GroupNorm slot per (band, tile, warp).
    Returns (out [B,H,W,Cout], writes [B,H,W] = how often each pixel was stored, slot_sums [B, slots] = sum over the
    output channels of each slot's valid rows)."""
    B, H, W, C = x_nhwc.shape
    Cout = w_oihw.shape[0]
    PW = W + 2
    BH = kernel_plan(W)
    bands = -(-H // BH)
    out = np.full((B, H, W, Cout), np.nan)
    writes = np.zeros((B, H, W), dtype=np.int64)
    slots = np.zeros((B, bands * 8))
    wt = w_oihw.astype(np.float64).transpose(2, 3, 1, 0)
    for b in range(B):
        for band_i in range(bands):
            y0 = band_i * BH
            bh = min(BH, H - y0)
            n_t = 2 if bh * PW > 128 else 1
            # the TMA box: BH+2 rows from y0-1, W+2 columns from -1, zeros outside the image; reads past the box hit
            # whatever follows in shared memory (modelled as NaN: junk rows must never reach the output)
            S = np.full(((256 + 2 * PW + 2), C), np.nan)
            S[:(BH + 2) * PW] = 0.0
            for r in range(BH + 2):
                y = y0 - 1 + r
                if 0 <= y < H:
                    S[r * PW + 1:r * PW + 1 + W] = x_nhwc[b, y]
            acc = np.zeros((256, Cout))
            for t in range(n_t):
                for ky in range(3):
                    for kx in range(3):
                        rows = np.arange(t * 128, (t + 1) * 128) + ky * PW + kx      # descriptor start shifted by whole rows
                        acc[t * 128:(t + 1) * 128] += S[rows] @ wt[ky, kx]
            if bias is not None:
                acc = acc + bias
            for t in range(2):
                for q in range(4):
                    m0w = t * 128 + q * 32
                    m = m0w + np.arange(32)
                    yy, xx = m // PW, m % PW
                    valid = (t < n_t) & (xx < W) & (yy < bh)
                    slots[b, (band_i * 2 + t) * 4 + q] = acc[m][valid].sum() if valid.any() else 0.0
                    if t >= n_t:
                        continue
                    yr0, yr1 = m0w // PW, min((m0w + 31) // PW, bh - 1)
                    for yr in range(yr0, yr1 + 1):
                        xs = m0w - yr * PW
                        if xs >= W:
                            continue
                        for r in range(32):                 # one TMA store: box row r lands at column xs + r, clipped to [0, W)
                            xcol = xs + r
                            if 0 <= xcol < W:
                                out[b, y0 + yr, xcol] = acc[m0w + r]
                                writes[b, y0 + yr, xcol] += 1
    return out, writes, slots


if __name__ == "__main__":
    for (H, W) in ((32, 32), (16, 16), (8, 8), (28, 28), (64, 64)):
        for BH in (3, 7, 15):
            if BH > H:
                continue
            m, n, f, px = band_plan(H, W, BH)
            print(f"{H}x{W} BH={BH}: {m} M rows -> {n} MMA tiles, {100 * f:.0f} % valid lanes, tile {px} pixels "
                  f"({px * 128 * 2 / 1024:.0f} KB hi+lo per 64-channel chunk), activation fill per valid pixel "
                  f"{px / (BH * W):.2f}x (now 9x)")
