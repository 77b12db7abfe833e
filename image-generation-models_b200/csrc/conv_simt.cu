// Generic fp32 implicit-GEMM convolution family on CUDA cores (NHWC activations).
//
// This is the shape-agnostic engine: it serves every convolution of the U-Net
// that the tcgen05 engine (conv_tc.cu) does not take (C_in = 1/3 stem, stride-2
// Downsample, 4x4 ConvTranspose2d Upsample and all their gradients), and is the
// bit-for-bit-fp32 comparator for the tensor-core engine in the tests.
//
// Replaces the ATen calls behind reference src/models/ddpm.py:70 (ConvTranspose2d),
// :79 (Conv2d s2), :116 (Conv2d 3x3), :134/:151/:152 (1x1) and their autograd.
#include "common.cuh"

namespace igm {

namespace {

constexpr int BM = 128;   // output pixels per CTA
constexpr int BN = 64;    // output channels per CTA
constexpr int BK = 16;    // input channels per K step (within one tap)
constexpr int APAD = 4;
constexpr int MAX_TAPS = 32;

struct RowCoord {
  int b, oy, ox;
  bool valid;
};

// decode a K-step (tap) into an input coordinate for one output pixel
__device__ __forceinline__ bool gather_coord(const ConvArgs& a, int oy, int ox, int ky, int kx, int& iy,
                                             int& ix) {
  const int pw = a.pad_w < 0 ? a.pad : a.pad_w;
  if (!a.transposed) {
    iy = oy * a.stride - a.pad + ky * a.dil;
    ix = ox * a.stride - pw + kx * a.dil;
    return iy >= 0 && iy < a.IH && ix >= 0 && ix < a.IW;
  } else {
    int ty = oy + a.pad - ky * a.dil;
    int tx = ox + pw - kx * a.dil;
    if (ty < 0 || tx < 0) return false;
    if (a.stride > 1 && ((ty % a.stride) != 0 || (tx % a.stride) != 0)) return false;
    iy = ty / a.stride;
    ix = tx / a.stride;
    return iy < a.IH && ix < a.IW;
  }
}

// grid: x = M tiles (within a phase), y = N tiles, z = phases (stride^2 for strided transposed mode)
template <bool VEC>
__global__ void __launch_bounds__(256) conv_igemm_kernel(const ConvArgs a) {
  __shared__ __align__(16) float As[2][BK][BM + APAD];
  __shared__ __align__(16) float Bs[2][BK][BN];
  __shared__ int s_taps[MAX_TAPS];
  __shared__ int s_ntaps;

  const int tid = threadIdx.x;
  const int Ctot = a.C0 + a.C1;

  // ---- phase decomposition (only for strided gather-transposed mode) ----
  const int ps = (a.transposed && a.stride > 1) ? a.stride : 1;
  const int py = blockIdx.z / ps, px = blockIdx.z % ps;
  const int OHp = (a.OH - py + ps - 1) / ps;   // rows of this phase
  const int OWp = (a.OW - px + ps - 1) / ps;
  const int Mp = a.B * OHp * OWp;
  const int m0 = blockIdx.x * BM;
  if (m0 >= Mp) return;
  const int n0 = blockIdx.y * BN;

  if (tid == 0) {
    int nt = 0;
    for (int ky = 0; ky < a.KH; ++ky)
      for (int kx = 0; kx < a.KW; ++kx) {
        bool live = true;
        if (ps > 1) {
          int ry = ((py + a.pad - ky * a.dil) % ps + ps) % ps;
          int rx = ((px + (a.pad_w < 0 ? a.pad : a.pad_w) - kx * a.dil) % ps + ps) % ps;
          live = (ry == 0) && (rx == 0);
        }
        if (live) s_taps[nt++] = ky * a.KW + kx;
      }
    s_ntaps = nt;
  }
  __syncthreads();
  const int ntaps = s_ntaps;
  const int nchunks = (Ctot + BK - 1) / BK;
  const int KT = ntaps * nchunks;

  // ---- per-thread A-load assignment: 2 x (row, 4 channels) ----
  RowCoord rc[2];
  int rrow[2], rc4[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int idx = tid + i * 256;
    rrow[i] = idx >> 2;
    rc4[i] = idx & 3;
    int m = m0 + rrow[i];
    rc[i].valid = m < Mp;
    int mm = rc[i].valid ? m : 0;
    int b = mm / (OHp * OWp);
    int r = mm - b * (OHp * OWp);
    int oyp = r / OWp;
    int oxp = r - oyp * OWp;
    rc[i].b = b;
    rc[i].oy = oyp * ps + py;
    rc[i].ox = oxp * ps + px;
  }
  // B-load assignment: row k = tid / 16, 4 columns at (tid % 16) * 4
  const int bk = tid >> 4, bn4 = (tid & 15) * 4;

  float4 ra[2];
  float4 rb;

  auto load_tile = [&](int kt) {
    const int ti = kt / nchunks;
    const int ch = kt - ti * nchunks;
    const int tap = s_taps[ti];
    const int ky = tap / a.KW, kx = tap - ky * a.KW;
    const int c0 = ch * BK;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      int iy, ix;
      if (rc[i].valid && gather_coord(a, rc[i].oy, rc[i].ox, ky, kx, iy, ix)) {
        const int c = c0 + rc4[i] * 4;
        const int64_t pix = ((int64_t)rc[i].b * a.IH + iy) * a.IW + ix;
        if (VEC) {
          if (c < a.C0)
            v = __ldg(reinterpret_cast<const float4*>(a.in0 + pix * a.C0 + c));
          else if (c < Ctot)
            v = __ldg(reinterpret_cast<const float4*>(a.in1 + pix * a.C1 + (c - a.C0)));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int cj = c + j;
            t[j] = 0.f;
            if (cj < a.C0)
              t[j] = __ldg(a.in0 + pix * a.C0 + cj);
            else if (cj < Ctot)
              t[j] = __ldg(a.in1 + pix * a.C1 + (cj - a.C0));
          }
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      ra[i] = v;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int k = c0 + bk;
      const int n = n0 + bn4;
      if (k < Ctot) {
        const float* wp = a.w + ((int64_t)tap * Ctot + k) * a.N + n;
        if (VEC) {
          if (n < a.N) v = __ldg(reinterpret_cast<const float4*>(wp));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (n + j < a.N) ? __ldg(wp + j) : 0.f;
          v = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      rb = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = rc4[i] * 4;
      As[buf][k + 0][rrow[i]] = ra[i].x;
      As[buf][k + 1][rrow[i]] = ra[i].y;
      As[buf][k + 2][rrow[i]] = ra[i].z;
      As[buf][k + 3][rrow[i]] = ra[i].w;
    }
    *reinterpret_cast<float4*>(&Bs[buf][bk][bn4]) = rb;
  };

  const int tn = tid & 15;   // 16 x 4 columns
  const int tm = tid >> 4;   // 16 x 8 rows
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  if (KT > 0) {
    load_tile(0);
    store_tile(0);
  }
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][tm * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tn * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue: + bias, + addend, split store ----
  const int n = n0 + tn * 4;
  if (n >= a.N) return;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (a.bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (n + j < a.N) bias[j] = __ldg(a.bias + n + j);
  }
  const int N1 = a.N - a.N0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    if (m >= Mp) continue;
    // map phase-local m back to the dense output pixel index
    int64_t opix;
    if (ps == 1) {
      opix = m;
    } else {
      int b = m / (OHp * OWp);
      int r = m - b * (OHp * OWp);
      int oyp = r / OWp, oxp = r - oyp * OWp;
      opix = ((int64_t)b * a.OH + (oyp * ps + py)) * a.OW + (oxp * ps + px);
    }
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = acc[i][j] + bias[j];
    if (VEC) {
      // N0 and N are multiples of 4 on this path, so a float4 never straddles the split
      if (n < a.N0) {
        float* o = a.out0 + opix * a.N0 + n;
        if (a.add0) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.add0 + opix * a.N0 + n));
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        float* o = a.out1 + opix * N1 + (n - a.N0);
        if (a.add1) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(a.add1 + opix * N1 + (n - a.N0)));
          v[0] += r4.x; v[1] += r4.y; v[2] += r4.z; v[3] += r4.w;
        }
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nj = n + j;
        if (nj >= a.N) break;
        if (nj < a.N0) {
          float r = a.add0 ? __ldg(a.add0 + opix * a.N0 + nj) : 0.f;
          a.out0[opix * a.N0 + nj] = v[j] + r;
        } else {
          float r = a.add1 ? __ldg(a.add1 + opix * N1 + (nj - a.N0)) : 0.f;
          a.out1[opix * N1 + (nj - a.N0)] = v[j] + r;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// wgrad: G[tap][qc][pc] = sum_pix Q[gather(pix,tap), qc] * P[pix, pc]
// CTA tile: 64 (qc) x 64 (pc), K step = 16 pixels, split-K over pixel ranges.
// grid: x = split, y = qc tiles * pc tiles, z = tap
// ---------------------------------------------------------------------------
constexpr int WB = 64;
constexpr int WK = 16;

template <bool QVEC, bool PVEC>
__global__ void __launch_bounds__(256) wgrad_kernel(const WgradArgs a, int pix_per_split, int p_tiles) {
  __shared__ __align__(16) float Qs[2][WK][WB];
  __shared__ __align__(16) float Ps[2][WK][WB];
  const int tid = threadIdx.x;
  const int tap = blockIdx.z;
  const int ky = tap / a.KW, kx = tap - ky * a.KW;
  const int qt = blockIdx.y / p_tiles, pt = blockIdx.y - qt * p_tiles;
  const int q0 = qt * WB, p0 = pt * WB;
  const int64_t npix = (int64_t)a.B * a.PH * a.PW;
  const int64_t pix_begin = (int64_t)blockIdx.x * pix_per_split;
  int64_t pix_end = pix_begin + pix_per_split;
  if (pix_end > npix) pix_end = npix;
  if (pix_begin >= pix_end) return;
  const int KT = (int)((pix_end - pix_begin + WK - 1) / WK);

  // loader: pixel row lr = tid / 16, 4 channels at lc4 = (tid % 16) * 4
  const int lr = tid >> 4, lc4 = (tid & 15) * 4;
  float4 rq, rp;
  auto load_tile = [&](int kt) {
    const int64_t pix = pix_begin + (int64_t)kt * WK + lr;
    rq = make_float4(0.f, 0.f, 0.f, 0.f);
    rp = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pix < pix_end) {
      const int b = (int)(pix / (a.PH * a.PW));
      const int r = (int)(pix - (int64_t)b * a.PH * a.PW);
      const int y = r / a.PW, x = r - y * a.PW;
      {
        const int c = p0 + lc4;
        const float* src = a.P + pix * a.PC + c;
        if (PVEC) {
          if (c < a.PC) rp = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (c + j < a.PC) ? __ldg(src + j) : 0.f;
          rp = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
      const int iy = y * a.stride - a.pad + ky * a.dil;
      const int ix = x * a.stride - (a.pad_w < 0 ? a.pad : a.pad_w) + kx * a.dil;
      if (iy >= 0 && iy < a.QH && ix >= 0 && ix < a.QW) {
        const int c = q0 + lc4;
        const float* src = a.Q + (((int64_t)b * a.QH + iy) * a.QW + ix) * a.QC + c;
        if (QVEC) {
          if (c < a.QC) rq = __ldg(reinterpret_cast<const float4*>(src));
        } else {
          float t[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) t[j] = (c + j < a.QC) ? __ldg(src + j) : 0.f;
          rq = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    }
  };
  auto store_tile = [&](int buf) {
    *reinterpret_cast<float4*>(&Qs[buf][lr][lc4]) = rq;
    *reinterpret_cast<float4*>(&Ps[buf][lr][lc4]) = rp;
  };

  const int tq = tid >> 4, tp = tid & 15;   // 16 x 16 threads, 4 x 4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < KT; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < KT) load_tile(kt + 1);
#pragma unroll
    for (int k = 0; k < WK; ++k) {
      const float4 q = *reinterpret_cast<const float4*>(&Qs[buf][k][tq * 4]);
      const float4 p = *reinterpret_cast<const float4*>(&Ps[buf][k][tp * 4]);
      const float qv[4] = {q.x, q.y, q.z, q.w};
      const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(qv[i], pv[j], acc[i][j]);
    }
    if (kt + 1 < KT) store_tile(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qc = q0 + tq * 4 + i;
    if (qc >= a.QC) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int pc = p0 + tp * 4 + j;
      if (pc >= a.PC) continue;
      atomicAdd(a.grad + qc * a.sq + pc * a.sp + tap, acc[i][j]);
    }
  }
}

// Forward of the stem convs (C_in = 1..4 image channels; 3x3 pad 1 or 1x1): a thread owns one output
// channel, keeps its KS*KS*C_in weights in registers and walks pixels; the image patch loads are warp
// broadcasts and the stores are 256-byte coalesced.  grid (pixel splits, C_out/64).
template <int CI, int KS>
__global__ void __launch_bounds__(256) conv_stem_kernel(const float* __restrict__ X, const float* __restrict__ Wp,
                                                        const float* __restrict__ bias, float* __restrict__ out, int B,
                                                        int H, int W, int Cout, int pix_per_cta) {
  constexpr int T = KS * KS, PAD = (KS - 1) / 2;
  __shared__ __align__(16) float sw[T * CI][64];   // this CTA's 64-channel slab of the packed [tap][ci][co] weights
  const int co0 = blockIdx.y * 64;
  for (int i = threadIdx.x; i < T * CI * 64; i += 256) {
    const int r = i >> 6, c = i & 63;
    sw[r][c] = (co0 + c < Cout) ? __ldg(Wp + (int64_t)r * Cout + co0 + c) : 0.f;
  }
  __syncthreads();
  const int q4 = (threadIdx.x & 15) * 4;   // 4 output channels per thread
  const int pl = threadIdx.x >> 4;         // 16 pixels in flight
  const int co = co0 + q4;
  if (co >= Cout) return;
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + co));
  const int64_t npix = (int64_t)B * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta;
  int64_t p1 = p0 + pix_per_cta;
  if (p1 > npix) p1 = npix;
  for (int64_t p = p0 + pl; p < p1; p += 16) {
    const int b = (int)(p / (H * W));
    const int r = (int)(p - (int64_t)b * H * W);
    const int y = r / W, x = r - y * W;
    float4 acc = bv;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
      const int iy = y + ky - PAD;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const int ix = x + kx - PAD;
        if (ix < 0 || ix >= W) continue;
        const float* xp = X + (((int64_t)b * H + iy) * W + ix) * CI;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          const float xv = __ldg(xp + ci);
          const float4 w4 = *reinterpret_cast<const float4*>(&sw[(ky * KS + kx) * CI + ci][q4]);
          acc.x = fmaf(xv, w4.x, acc.x); acc.y = fmaf(xv, w4.y, acc.y);
          acc.z = fmaf(xv, w4.z, acc.z); acc.w = fmaf(xv, w4.w, acc.w);
        }
      }
    }
    *reinterpret_cast<float4*>(out + p * Cout + co) = acc;
  }
}

// Same contract, row-band form.  The kernel above is bound by shared-memory weight traffic (one 128-bit weight load per
// four FMAs, nothing reused across pixels).  Here a lane owns TWO output channels and keeps all KS*KS*C_in weight pairs
// in registers; a CTA stages the zero-padded input rows of a band of R image rows in shared memory, and every input
// value is then one broadcast shared-memory load feeding one packed FMA (FFMA2) per lane.  Out-of-image taps multiply
// zeros, in the same tap order as above, so results are bit-identical.  grid (B * ceil(H / R), C_out / 64), 8 warps.
template <int CI, int KS>
__global__ void __launch_bounds__(256) conv_stem_rows_kernel(const float* __restrict__ X, const float* __restrict__ Wp,
                                                             const float* __restrict__ bias, float* __restrict__ out,
                                                             int B, int H, int W, int Cout, int R) {
  constexpr int T = KS * KS, PAD = (KS - 1) / 2;
  // [R + 2 PAD][W + 2 PAD] pixels of FOUR floats (channels CI .. 3 are padding): one aligned 128-bit broadcast load brings
  // all input channels of a tap (the scalar form issued CI loads per tap and was bound by the shared-memory pipe)
  extern __shared__ __align__(16) float sx[];
  const int bands = (H + R - 1) / R;
  const int b = blockIdx.x / bands, y0 = (blockIdx.x - b * bands) * R;
  const int rows = min(R, H - y0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = blockIdx.y * 64 + lane * 2;
  const bool cok = co < Cout;
  float2 w[T * CI];
#pragma unroll
  for (int k = 0; k < T * CI; ++k)
    w[k] = cok ? __ldg(reinterpret_cast<const float2*>(Wp + (int64_t)k * Cout + co)) : make_float2(0.f, 0.f);
  float2 bv = make_float2(0.f, 0.f);
  if (bias && cok) bv = __ldg(reinterpret_cast<const float2*>(bias + co));
  const int PWp = W + 2 * PAD;               // padded pixels per row
  const int nrow = rows + 2 * PAD;
  for (int i = threadIdx.x; i < nrow * PWp; i += 256) {
    const int rr = i / PWp, px = i - rr * PWp;
    const int iy = y0 - PAD + rr, ix = px - PAD;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* xp = X + (((int64_t)b * H + iy) * W + ix) * CI;
      float* ve = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) ve[ci] = __ldg(xp + ci);
    }
    *reinterpret_cast<float4*>(sx + (size_t)i * 4) = v;
  }
  __syncthreads();
  if (!cok) return;
  // two pixels per iteration: two independent FMA chains per lane (each pixel keeps the tap order of the kernel above)
  const int npx = rows * W;
  for (int p = warp; p < npx; p += 16) {
    const int pb = (p + 8 < npx) ? p + 8 : p;
    const int ya = p / W, xa = p - ya * W;
    const int yb = pb / W, xb = pb - yb * W;
    const float* sa = sx + ((size_t)ya * PWp + xa) * 4;
    const float* sb = sx + ((size_t)yb * PWp + xb) * 4;
    float2 acca = bv, accb = bv;
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float4 va4 = *reinterpret_cast<const float4*>(sa + ((size_t)ky * PWp + kx) * 4);
        const float4 vb4 = *reinterpret_cast<const float4*>(sb + ((size_t)ky * PWp + kx) * 4);
        const float va[4] = {va4.x, va4.y, va4.z, va4.w}, vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w};
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) {
          ffma2(acca, make_float2(va[ci], va[ci]), w[(ky * KS + kx) * CI + ci]);
          ffma2(accb, make_float2(vb[ci], vb[ci]), w[(ky * KS + kx) * CI + ci]);
        }
      }
    *reinterpret_cast<float2*>(out + (((int64_t)b * H + y0 + ya) * W + xa) * Cout + co) = acca;
    if (pb != p) *reinterpret_cast<float2*>(out + (((int64_t)b * H + y0 + yb) * W + xb) * Cout + co) = accb;
  }
}

template <int CI>
static void launch_stem(const ConvArgs& a, dim3 grid, int per, cudaStream_t st) {
  static const bool rows_off = [] { const char* e = getenv("IGM_STEM_ROWS"); return e && e[0] == '0'; }();
  if (!rows_off && a.N % 2 == 0) {
    // bands of 8 rows while that still leaves two CTAs per SM (fewer prologues, one wave); else 4
    int R = ((int64_t)a.B * cdiv(a.IH, 8) * cdiv(a.N, 64) >= 296) ? 8 : 4;
    if (a.IH < R) R = a.IH;
    const int PADk = (a.KH - 1) / 2;
    const size_t smem = (size_t)(R + 2 * PADk) * (a.IW + 2 * PADk) * 4 * sizeof(float);
    if (smem <= 40 * 1024) {
      dim3 g2((unsigned)(a.B * cdiv(a.IH, R)), (unsigned)cdiv(a.N, 64));
      if (a.KH == 3)
        conv_stem_rows_kernel<CI, 3><<<g2, 256, smem, st>>>(a.in0, a.w, a.bias, a.out0, a.B, a.IH, a.IW, a.N, R);
      else
        conv_stem_rows_kernel<CI, 1><<<g2, 256, smem, st>>>(a.in0, a.w, a.bias, a.out0, a.B, a.IH, a.IW, a.N, R);
      return;
    }
  }
  if (a.KH == 3)
    conv_stem_kernel<CI, 3><<<grid, 256, 0, st>>>(a.in0, a.w, a.bias, a.out0, a.B, a.IH, a.IW, a.N, per);
  else
    conv_stem_kernel<CI, 1><<<grid, 256, 0, st>>>(a.in0, a.w, a.bias, a.out0, a.B, a.IH, a.IW, a.N, per);
}

// Weight gradient of the stem conv (C_in = 1..4 image channels, 3x3, stride 1, pad 1):
//   dW[co][ci][ky][kx] += sum_pix dY[pix][co] * X[pix + (ky-1, kx-1)][ci]
// The generic 64x64-tile kernel wastes >90 % of its tile on 3 input channels; here a thread owns one
// output channel and keeps its 9*C_in partial sums in registers.  grid (pixel splits, C_out/64).
template <int CI>
__global__ void __launch_bounds__(256) wgrad_stem_kernel(const float* __restrict__ X, const float* __restrict__ dY,
                                                         float* __restrict__ grad, int B, int H, int W, int Cout,
                                                         int pix_per_cta) {
  __shared__ float red[4][64][9 * CI + 1];
  const int co = blockIdx.y * 64 + (threadIdx.x & 63);
  const int lane4 = threadIdx.x >> 6;
  const int64_t npix = (int64_t)B * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta;
  int64_t p1 = p0 + pix_per_cta;
  if (p1 > npix) p1 = npix;
  float acc[9 * CI];
#pragma unroll
  for (int i = 0; i < 9 * CI; ++i) acc[i] = 0.f;
  if (co < Cout) {
    for (int64_t p = p0 + lane4; p < p1; p += 4) {
      const float g = __ldg(dY + p * Cout + co);
      const int b = (int)(p / (H * W));
      const int r = (int)(p - (int64_t)b * H * W);
      const int y = r / W, x = r - y * W;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int ix = x + kx - 1;
          if (ix < 0 || ix >= W) continue;
          const float* xp = X + (((int64_t)b * H + iy) * W + ix) * CI;
#pragma unroll
          for (int ci = 0; ci < CI; ++ci) acc[(ky * 3 + kx) * CI + ci] = fmaf(g, __ldg(xp + ci), acc[(ky * 3 + kx) * CI + ci]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 9 * CI; ++i) red[lane4][threadIdx.x & 63][i] = acc[i];
  __syncthreads();
  if (lane4 == 0 && co < Cout) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) {
        const int i = t * CI + ci;
        const float v = red[0][threadIdx.x][i] + red[1][threadIdx.x][i] + red[2][threadIdx.x][i] + red[3][threadIdx.x][i];
        atomicAdd(grad + ((int64_t)co * CI + ci) * 9 + t, v);   // OIHW
      }
  }
}

// "Thin" weight gradients: one operand has <= 4 channels (the image / the predicted noise), the other is wide.
//   grad[w*s_wide + t*s_thin + tap] += sum_pix Wide[pix][w] * Thin[pix + tap][t]
// (stem 3x3 and 1x1 convs: Wide = dY, Thin = X gathered through the taps; final 1x1 conv: Wide = activations,
// Thin = d_pred, KS = 1).  A CTA owns RB image rows of one image and 64 wide channels; the zero-padded thin tile
// sits in shared memory, so the inner loop is one coalesced load of the wide value and KS*KS*TC broadcast LDS + FMA.
template <int TC, int KS>
__global__ void __launch_bounds__(256) wgrad_thin_kernel(const float* __restrict__ Wide, const float* __restrict__ Thin,
                                                         float* __restrict__ grad, int H, int W, int WC, int RB,
                                                         int64_t s_wide, int64_t s_thin) {
  constexpr int T = KS * KS, PAD = (KS - 1) / 2, NA = T * TC;
  extern __shared__ __align__(16) float smem_thin[];   // tile [(RB + 2 PAD)][(W + 2 PAD)] of float4 (TC padded to 4), then red
  const int TW = W + 2 * PAD, TH = RB + 2 * PAD;
  float4* tile = reinterpret_cast<float4*>(smem_thin);
  float (*red)[64][NA + 1] = reinterpret_cast<float (*)[64][NA + 1]>(smem_thin + TH * TW * 4);
  const int tiles_per_img = (H + RB - 1) / RB;
  const int b = blockIdx.x / tiles_per_img, y0 = (blockIdx.x % tiles_per_img) * RB;
  for (int i = threadIdx.x; i < TH * TW; i += 256) {
    const int tx = i % TW, ty = i / TW;
    const int y = y0 + ty - PAD, x = tx - PAD;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < H && x >= 0 && x < W) {
      const float* src = Thin + (((int64_t)b * H + y) * W + x) * TC;
#pragma unroll
      for (int ci = 0; ci < TC; ++ci) v[ci] = __ldg(src + ci);
    }
    tile[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const int c = blockIdx.y * 64 + (threadIdx.x & 63);
  const int lane4 = threadIdx.x >> 6;
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  if (c < WC) {
    // each of the 4 pixel lanes walks a contiguous run of pixels: along a row the KS x KS window of thin values
    // slides, so only its new column (KS 128-bit LDS) is fetched per pixel
    const int npix = min(RB, H - y0) * W;
    const int per = (npix + 3) >> 2;
    const int i0 = lane4 * per, i1 = min(i0 + per, npix);
    int yy = i0 / W, xx = i0 - yy * W;
    float4 win[KS][KS];
    bool fresh = true;
    const float* wp = Wide + (((int64_t)b * H + y0) * W + i0) * WC + c;
    constexpr int UB = 8;   // wide values requested per batch (memory-level parallelism: the loop is otherwise latency-bound)
    for (int ib = i0; ib < i1; ib += UB, wp += (int64_t)UB * WC) {
      float gb[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) gb[u] = (ib + u < i1) ? __ldg(wp + (int64_t)u * WC) : 0.f;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (ib + u >= i1) break;
        const float g = gb[u];

        if (fresh) {
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) win[ky][kx] = tile[(yy + ky) * TW + xx + kx];
          fresh = false;
        } else {
#pragma unroll
          for (int ky = 0; ky < KS; ++ky) {
#pragma unroll
            for (int kx = 0; kx + 1 < KS; ++kx) win[ky][kx] = win[ky][kx + 1];
            win[ky][KS - 1] = tile[(yy + ky) * TW + xx + KS - 1];
          }
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float tv[4] = {win[ky][kx].x, win[ky][kx].y, win[ky][kx].z, win[ky][kx].w};
#pragma unroll
            for (int ci = 0; ci < TC; ++ci) acc[(ky * KS + kx) * TC + ci] = fmaf(g, tv[ci], acc[(ky * KS + kx) * TC + ci]);
          }
        if (++xx == W) { xx = 0; ++yy; fresh = true; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NA; ++i) red[lane4][threadIdx.x & 63][i] = acc[i];
  __syncthreads();
  if (lane4 == 0 && c < WC) {
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int ci = 0; ci < TC; ++ci) {
        const int i = t * TC + ci;
        const float v = red[0][threadIdx.x][i] + red[1][threadIdx.x][i] + red[2][threadIdx.x][i] + red[3][threadIdx.x][i];
        atomicAdd(grad + (int64_t)c * s_wide + (int64_t)ci * s_thin + t, v);
      }
  }
}

template <int TC, int KS>
static int launch_thin_impl(const LaunchCtx& lc, const float* wide, const float* thin, float* grad, int B, int H, int W, int WC,
                            int64_t s_wide, int64_t s_thin) {
  constexpr int PAD = (KS - 1) / 2, NA = KS * KS * TC;
  int RB = (int)cdiv64((int64_t)B * H, 296);   // ~2 CTAs per SM and 64-channel slab (fewer same-address atomics)
  if (RB < 1) RB = 1;
  if (RB > H) RB = H;
  while (RB > 1 && (RB + 2 * PAD) * (W + 2 * PAD) * 4 > 4096) --RB;
  const int tile_f = (RB + 2 * PAD) * (W + 2 * PAD) * 4;
  const size_t smem = (size_t)(tile_f + 4 * 64 * (NA + 1)) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(wgrad_thin_kernel<TC, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  if (smem > 64 * 1024) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "thin wgrad: image row too wide");
  dim3 grid((unsigned)(B * cdiv(H, RB)), (unsigned)cdiv(WC, 64));
  wgrad_thin_kernel<TC, KS><<<grid, 256, smem, lc.stream>>>(wide, thin, grad, H, W, WC, RB, s_wide, s_thin);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

template <int KS>
static int launch_thin(const LaunchCtx& lc, int TC, const float* wide, const float* thin, float* grad, int B, int H, int W,
                       int WC, int64_t s_wide, int64_t s_thin) {
  switch (TC) {
    case 1: return launch_thin_impl<1, KS>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    case 2: return launch_thin_impl<2, KS>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    case 3: return launch_thin_impl<3, KS>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    default: return launch_thin_impl<4, KS>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
  }
}

// ---- thin 4x4 stride-2 layers (first Conv2d / last ConvTranspose2d of the VQ-VAE: 3 image channels on one side) ----------
// The generic 64 x 64-tile kernels above spend > 90 % of a tile on padding when one side has 3 channels (ncu, round 2: 159 us
// for Conv2d(3, 32, 4, 2, 1) on 32 x 128 x 128 images, 715 us for ConvTranspose2d(32, 3, 4, 2, 1), 557 us for each of their
// weight gradients -- 40 % of the VQ-VAE step for 0.3 % of its FLOPs).  Same ideas as the stem kernels of the U-Net: the thin
// operand lives zero-padded in shared memory as 4-float pixels, a lane owns wide channels, thin values are broadcasts.

// Conv2d(CI <= 4 -> N, 4, 2, 1), also the data gradient of ConvTranspose2d(N -> CI, 4, 2, 1).  Wp packed [16][CI][N].
// grid (B * ceil(OH / R), N / (32 * CPL)); a lane owns CPL output channels and keeps their 16 * CI weights in registers.
template <int CI, int CPL>
__global__ void __launch_bounds__(256) conv_thin_in_s2_kernel(const float* __restrict__ X, const float* __restrict__ Wp,
                                                              const float* __restrict__ bias, float* __restrict__ out,
                                                              int H, int W, int OH, int OW, int Cout, int R) {
  constexpr int KS = 4, T = 16;
  extern __shared__ __align__(16) float sx[];   // [(R - 1) * 2 + 4][2 OW + 2] pixels of four floats
  const int bands = (OH + R - 1) / R;
  const int b = blockIdx.x / bands, y0 = (blockIdx.x - b * bands) * R;
  const int rows = min(R, OH - y0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int co = (blockIdx.y * 32 + lane) * CPL;
  float w[T * CI][CPL];
#pragma unroll
  for (int k = 0; k < T * CI; ++k)
#pragma unroll
    for (int j = 0; j < CPL; ++j) w[k][j] = __ldg(Wp + (int64_t)k * Cout + co + j);
  float bv[CPL];
#pragma unroll
  for (int j = 0; j < CPL; ++j) bv[j] = bias ? __ldg(bias + co + j) : 0.f;
  const int PWp = 2 * OW + 2;
  const int nrow = (rows - 1) * 2 + KS;
  for (int i = threadIdx.x; i < nrow * PWp; i += 256) {
    const int rr = i / PWp, px = i - rr * PWp;
    const int iy = 2 * y0 - 1 + rr, ix = px - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const float* xp = X + (((int64_t)b * H + iy) * W + ix) * CI;
      float* ve = reinterpret_cast<float*>(&v);
#pragma unroll
      for (int ci = 0; ci < CI; ++ci) ve[ci] = __ldg(xp + ci);
    }
    *reinterpret_cast<float4*>(sx + (size_t)i * 4) = v;
  }
  __syncthreads();
  const int npx = rows * OW;
  for (int p = warp; p < npx; p += 16) {   // two pixels per iteration: independent FMA chains
    const int pb = (p + 8 < npx) ? p + 8 : p;
    const int ya = p / OW, xa = p - ya * OW;
    const int yb = pb / OW, xb = pb - yb * OW;
    const float* sa = sx + ((size_t)(2 * ya) * PWp + 2 * xa) * 4;
    const float* sb = sx + ((size_t)(2 * yb) * PWp + 2 * xb) * 4;
    float acca[CPL], accb[CPL];
#pragma unroll
    for (int j = 0; j < CPL; ++j) { acca[j] = bv[j]; accb[j] = bv[j]; }
#pragma unroll
    for (int ky = 0; ky < KS; ++ky)
#pragma unroll
      for (int kx = 0; kx < KS; ++kx) {
        const float4 va4 = *reinterpret_cast<const float4*>(sa + ((size_t)ky * PWp + kx) * 4);
        const float4 vb4 = *reinterpret_cast<const float4*>(sb + ((size_t)ky * PWp + kx) * 4);
        const float va[4] = {va4.x, va4.y, va4.z, va4.w}, vb[4] = {vb4.x, vb4.y, vb4.z, vb4.w};
#pragma unroll
        for (int ci = 0; ci < CI; ++ci)
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            acca[j] = fmaf(va[ci], w[(ky * KS + kx) * CI + ci][j], acca[j]);
            accb[j] = fmaf(vb[ci], w[(ky * KS + kx) * CI + ci][j], accb[j]);
          }
      }
    float* oa = out + (((int64_t)b * OH + y0 + ya) * OW + xa) * Cout + co;
    float* ob = out + (((int64_t)b * OH + y0 + yb) * OW + xb) * Cout + co;
#pragma unroll
    for (int j = 0; j < CPL; ++j) oa[j] = acca[j];
    if (pb != p) {
#pragma unroll
      for (int j = 0; j < CPL; ++j) ob[j] = accb[j];
    }
  }
}

// ConvTranspose2d(Cin -> CO <= 4, 4, 2, 1) forward.  Wp packed [16][Cin][CO] (staged in shared memory, read as broadcasts).
// A thread produces the 2 x 2 output block (2j + a, 2i + b) from the 3 x 3 input pixels around (j, i): every tap is used
// exactly once per block (oy = 2 iy - 1 + ky), 16 * Cin * CO FMAs per thread.  grid: ceil(B * IH * IW / 256).
template <int CO>
__global__ void __launch_bounds__(256) convT_thin_out_s2_kernel(const float* __restrict__ X, const float* __restrict__ Wp,
                                                                const float* __restrict__ bias, float* __restrict__ out,
                                                                int B, int IH, int IW, int Cin) {
  extern __shared__ __align__(16) float swt[];   // [16][Cin][CO]
  for (int i = threadIdx.x; i < 16 * Cin * CO; i += 256) swt[i] = __ldg(Wp + i);
  __syncthreads();
  const int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (t >= (int64_t)B * IH * IW) return;
  const int i = (int)(t % IW);
  const int j = (int)((t / IW) % IH);
  const int b = (int)(t / ((int64_t)IW * IH));
  float acc[2][2][CO];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c)
#pragma unroll
      for (int o = 0; o < CO; ++o) acc[a][c][o] = bias ? __ldg(bias + o) : 0.f;
  // output row parity a uses (ky, window row): a = 0 -> (1, 1), (3, 0);  a = 1 -> (0, 2), (2, 1)   (window row r <-> iy = j - 1 + r)
  constexpr int kk[2][2] = {{1, 3}, {0, 2}};
  constexpr int rr[2][2] = {{1, 0}, {2, 1}};
  for (int c4 = 0; c4 < Cin; c4 += 4) {
    float4 xin[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int iy = j - 1 + r, ix = i - 1 + c;
        xin[r][c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < IH && ix >= 0 && ix < IW)
          xin[r][c] = __ldg(reinterpret_cast<const float4*>(X + (((int64_t)b * IH + iy) * IW + ix) * Cin + c4));
      }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int tx = 0; tx < 2; ++tx) {
            const int tap = kk[a][ty] * 4 + kk[c][tx];
            const float4 xv4 = xin[rr[a][ty]][rr[c][tx]];
            const float xv[4] = {xv4.x, xv4.y, xv4.z, xv4.w};
            const float* wp = swt + ((size_t)tap * Cin + c4) * CO;   // 4 * CO consecutive floats, 16-byte aligned
            float wv[4 * CO];
#pragma unroll
            for (int q = 0; q < CO; ++q) *reinterpret_cast<float4*>(&wv[4 * q]) = *reinterpret_cast<const float4*>(wp + 4 * q);
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
#pragma unroll
              for (int o = 0; o < CO; ++o) acc[a][c][o] = fmaf(xv[ci], wv[ci * CO + o], acc[a][c][o]);
          }
  }
  const int OH = 2 * IH, OW = 2 * IW;
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float* o = out + (((int64_t)b * OH + 2 * j + a) * OW + 2 * i + c) * CO;
#pragma unroll
      for (int q = 0; q < CO; ++q) o[q] = acc[a][c][q];
    }
}

// Weight gradient of both thin 4x4 stride-2 layers:
//   grad[w * s_wide + t * s_thin + tap] += sum_{b, y, x} Wide[b, y, x][w] * Thin[b, 2y - 1 + ky, 2x - 1 + kx][t]
// (Conv2d: Wide = dY, Thin = X; ConvTranspose2d: Wide = X, Thin = dY).  wgrad_thin_kernel with a stride-2 window: a CTA owns RB
// coarse rows of one image and CW wide channels, 256 / CW pixel lanes walk contiguous runs of coarse pixels, the window of
// 4 x 4 thin pixels slides by two columns per step (two new columns = 8 broadcast 128-bit loads per 16 * TC FMAs).
template <int TC, int CW>
__global__ void __launch_bounds__(256) wgrad_thin_s2_kernel(const float* __restrict__ Wide, const float* __restrict__ Thin,
                                                            float* __restrict__ grad, int H, int W, int WC, int RB,
                                                            int64_t s_wide, int64_t s_thin) {
  constexpr int KS = 4, T = 16, NA = T * TC, LG = 256 / CW;
  extern __shared__ __align__(16) float smem_thin[];   // tile [(RB - 1) * 2 + 4][2 W + 2] of float4, then red[LG][CW][NA + 1]
  const int TW = 2 * W + 2, TH = (RB - 1) * 2 + KS;
  const int HF = 2 * H, WF = 2 * W;
  float4* tile = reinterpret_cast<float4*>(smem_thin);
  float (*red)[CW][NA + 1] = reinterpret_cast<float (*)[CW][NA + 1]>(smem_thin + (size_t)TH * TW * 4);
  const int tiles_per_img = (H + RB - 1) / RB;
  const int b = blockIdx.x / tiles_per_img, y0 = (blockIdx.x % tiles_per_img) * RB;
  for (int i = threadIdx.x; i < TH * TW; i += 256) {
    const int tx = i % TW, ty = i / TW;
    const int y = 2 * y0 - 1 + ty, x = tx - 1;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (y >= 0 && y < HF && x >= 0 && x < WF) {
      const float* src = Thin + (((int64_t)b * HF + y) * WF + x) * TC;
#pragma unroll
      for (int ci = 0; ci < TC; ++ci) v[ci] = __ldg(src + ci);
    }
    tile[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const int cl = threadIdx.x % CW, lg = threadIdx.x / CW;
  const int c = blockIdx.y * CW + cl;
  float acc[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) acc[i] = 0.f;
  if (c < WC) {
    const int npix = min(RB, H - y0) * W;
    const int per = (npix + LG - 1) / LG;
    const int i0 = lg * per, i1 = min(i0 + per, npix);
    int yy = i0 / W, xx = i0 - yy * W;
    float4 win[KS][KS];
    bool fresh = true;
    const float* wp = Wide + (((int64_t)b * H + y0) * W + i0) * WC + c;
    constexpr int UB = 8;
    for (int ib = i0; ib < i1; ib += UB, wp += (int64_t)UB * WC) {
      float gb[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) gb[u] = (ib + u < i1) ? __ldg(wp + (int64_t)u * WC) : 0.f;
#pragma unroll
      for (int u = 0; u < UB; ++u) {
        if (ib + u >= i1) break;
        const float g = gb[u];
        const float4* t0 = tile + (size_t)(2 * yy) * TW + 2 * xx;
        if (fresh) {
#pragma unroll
          for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) win[ky][kx] = t0[ky * TW + kx];
          fresh = false;
        } else {
#pragma unroll
          for (int ky = 0; ky < KS; ++ky) {
            win[ky][0] = win[ky][2];
            win[ky][1] = win[ky][3];
            win[ky][2] = t0[ky * TW + 2];
            win[ky][3] = t0[ky * TW + 3];
          }
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
          for (int kx = 0; kx < KS; ++kx) {
            const float tv[4] = {win[ky][kx].x, win[ky][kx].y, win[ky][kx].z, win[ky][kx].w};
#pragma unroll
            for (int ci = 0; ci < TC; ++ci) acc[(ky * KS + kx) * TC + ci] = fmaf(g, tv[ci], acc[(ky * KS + kx) * TC + ci]);
          }
        if (++xx == W) { xx = 0; ++yy; fresh = true; }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NA; ++i) red[lg][cl][i] = acc[i];
  __syncthreads();
  if (lg == 0 && c < WC) {
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
      for (int ci = 0; ci < TC; ++ci) {
        const int i = t * TC + ci;
        float v = 0.f;
#pragma unroll
        for (int l = 0; l < LG; ++l) v += red[l][cl][i];
        atomicAdd(grad + (int64_t)c * s_wide + (int64_t)ci * s_thin + t, v);
      }
  }
}

template <int TC, int CW>
static int launch_thin_s2_impl(const LaunchCtx& lc, const float* wide, const float* thin, float* grad, int B, int H, int W, int WC,
                               int64_t s_wide, int64_t s_thin) {
  constexpr int NA = 16 * TC, LG = 256 / CW;
  const int red_f = LG * CW * (NA + 1);
  int RB = (int)cdiv64((int64_t)B * H, 296);   // ~2 CTAs per SM and channel slab
  if (RB < 1) RB = 1;
  if (RB > H) RB = H;
  auto tile_f = [&](int rb) { return ((rb - 1) * 2 + 4) * (2 * W + 2) * 4; };
  while (RB > 1 && (size_t)(tile_f(RB) + red_f) * sizeof(float) > 96 * 1024) --RB;
  const size_t smem = (size_t)(tile_f(RB) + red_f) * sizeof(float);
  if (smem > 96 * 1024) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "thin stride-2 wgrad: image row too wide");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(wgrad_thin_s2_kernel<TC, CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr = true;
  }
  dim3 grid((unsigned)(B * cdiv(H, RB)), (unsigned)cdiv(WC, CW));
  wgrad_thin_s2_kernel<TC, CW><<<grid, 256, smem, lc.stream>>>(wide, thin, grad, H, W, WC, RB, s_wide, s_thin);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

static int launch_thin_s2(const LaunchCtx& lc, int TC, const float* wide, const float* thin, float* grad, int B, int H, int W,
                          int WC, int64_t s_wide, int64_t s_thin) {
  const bool narrow = WC <= 32;
  switch (TC) {
    case 1: return narrow ? launch_thin_s2_impl<1, 32>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin)
                          : launch_thin_s2_impl<1, 64>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    case 2: return narrow ? launch_thin_s2_impl<2, 32>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin)
                          : launch_thin_s2_impl<2, 64>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    case 3: return narrow ? launch_thin_s2_impl<3, 32>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin)
                          : launch_thin_s2_impl<3, 64>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
    default: return narrow ? launch_thin_s2_impl<4, 32>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin)
                           : launch_thin_s2_impl<4, 64>(lc, wide, thin, grad, B, H, W, WC, s_wide, s_thin);
  }
}

template <int CI>
static int launch_thin_in_s2(const LaunchCtx& lc, const ConvArgs& a) {
  const int cpl = (a.N % 64 == 0) ? 2 : 1;
  int R = ((int64_t)a.B * cdiv(a.OH, 8) * (a.N / (32 * cpl)) >= 296) ? 8 : 4;
  if (a.OH < R) R = a.OH;
  const size_t smem = (size_t)((R - 1) * 2 + 4) * (2 * a.OW + 2) * 4 * sizeof(float);
  if (smem > 96 * 1024) return -1;
  dim3 grid((unsigned)(a.B * cdiv(a.OH, R)), (unsigned)(a.N / (32 * cpl)));
  if (cpl == 2) {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_thin_in_s2_kernel<CI, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    conv_thin_in_s2_kernel<CI, 2><<<grid, 256, smem, lc.stream>>>(a.in0, a.w, a.bias, a.out0, a.IH, a.IW, a.OH, a.OW, a.N, R);
  } else {
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(conv_thin_in_s2_kernel<CI, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024); attr = true; }
    conv_thin_in_s2_kernel<CI, 1><<<grid, 256, smem, lc.stream>>>(a.in0, a.w, a.bias, a.out0, a.IH, a.IW, a.OH, a.OW, a.N, R);
  }
  return 0;
}

template <int CO>
static void launch_thin_out_s2(const LaunchCtx& lc, const ConvArgs& a) {
  const size_t smem = (size_t)16 * a.C0 * CO * sizeof(float);
  const int64_t nthr = (int64_t)a.B * a.IH * a.IW;
  convT_thin_out_s2_kernel<CO><<<(unsigned)cdiv64(nthr, 256), 256, smem, lc.stream>>>(a.in0, a.w, a.bias, a.out0, a.B, a.IH, a.IW, a.C0);
}

// out[n] += sum_m x[m, n]; grid.x = row splits; block (32 x 8): 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, int64_t M, int N,
                                                     float* __restrict__ out, int rows_per_cta) {
  __shared__ float red[8][33];
  const int col = blockIdx.y * 32 + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
  int64_t r1 = r0 + rows_per_cta;
  if (r1 > M) r1 = M;
  float s = 0.f;
  if (col < N)
    for (int64_t r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(x + r * N + col);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && col < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(out + col, t);
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ src, float* __restrict__ dst, int taps, int K,
                                   int N, int64_t sk, int64_t sn) {
  const int64_t total = (int64_t)taps * K * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int64_t r = i / N;
    const int k = (int)(r % K);
    const int tap = (int)(r / K);
    dst[i] = src[k * sk + n * sn + tap];
  }
}

}  // namespace

int launch_conv(const LaunchCtx& lc, const ConvArgs& a) {
  const int Ctot = a.C0 + a.C1;
  if (a.KH * a.KW > MAX_TAPS) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv: too many taps");
  if (a.N0 <= 0 || a.N0 > a.N || (a.N0 < a.N && !a.out1))
    IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv: bad output split");
  if (!a.in1 && a.C1 == 0 && a.C0 >= 1 && a.C0 <= 4 && !a.transposed && a.stride == 1 && a.dil == 1 && a.KH == a.KW &&
      a.pad_w < 0 &&
      ((a.KH == 3 && a.pad == 1) || (a.KH == 1 && a.pad == 0)) && a.IH == a.OH && a.IW == a.OW && !a.add0 && a.N0 == a.N &&
      a.N % 4 == 0) {
    const int64_t npix = (int64_t)a.B * a.OH * a.OW;
    int per = (int)cdiv64(npix, 148 * 8);
    per = (per + 15) & ~15;
    dim3 grid((unsigned)cdiv64(npix, per), (unsigned)cdiv(a.N, 64));
    ProfScope ps_(lc, a.kclass, 2.0 * npix * (double)a.N * a.C0 * a.KH * a.KW, 4.0 * npix * (a.N + a.C0));
    switch (a.C0) {
      case 1: launch_stem<1>(a, grid, per, lc.stream); break;
      case 2: launch_stem<2>(a, grid, per, lc.stream); break;
      case 3: launch_stem<3>(a, grid, per, lc.stream); break;
      default: launch_stem<4>(a, grid, per, lc.stream); break;
    }
    IGM_POST_LAUNCH(lc);
    return IGM_OK;
  }
  static const bool thin_s2 = [] { const char* e = getenv("IGM_THIN_S2"); return !(e && e[0] == '0'); }();
  const bool s2_geo = thin_s2 && !a.in1 && a.C1 == 0 && a.stride == 2 && a.dil == 1 && a.KH == 4 && a.KW == 4 && a.pad == 1 &&
                      (a.pad_w < 0 || a.pad_w == 1) && !a.add0 && a.N0 == a.N;
  if (s2_geo && !a.transposed && a.C0 >= 1 && a.C0 <= 4 && a.N % 32 == 0 && a.IH == 2 * a.OH && a.IW == 2 * a.OW) {
    // Conv2d(<= 4 -> N, 4, 2, 1) (and the data gradient of the matching ConvTranspose2d)
    ProfScope ps_(lc, a.kclass, 2.0 * a.B * a.OH * a.OW * (double)a.N * a.C0 * 16, 4.0 * a.B * ((double)a.OH * a.OW * a.N + (double)a.IH * a.IW * a.C0));
    int rc = -1;
    switch (a.C0) {
      case 1: rc = launch_thin_in_s2<1>(lc, a); break;
      case 2: rc = launch_thin_in_s2<2>(lc, a); break;
      case 3: rc = launch_thin_in_s2<3>(lc, a); break;
      default: rc = launch_thin_in_s2<4>(lc, a); break;
    }
    if (rc == 0) {
      IGM_POST_LAUNCH(lc);
      return IGM_OK;
    }
  }
  if (s2_geo && a.transposed && a.N >= 1 && a.N <= 4 && a.C0 % 4 == 0 && 16 * a.C0 * a.N <= 12288 && a.OH == 2 * a.IH &&
      a.OW == 2 * a.IW) {
    // ConvTranspose2d(C -> <= 4, 4, 2, 1)
    ProfScope ps_(lc, a.kclass, 2.0 * a.B * a.IH * a.IW * (double)a.N * a.C0 * 16, 4.0 * a.B * ((double)a.OH * a.OW * a.N + (double)a.IH * a.IW * a.C0));
    switch (a.N) {
      case 1: launch_thin_out_s2<1>(lc, a); break;
      case 2: launch_thin_out_s2<2>(lc, a); break;
      case 3: launch_thin_out_s2<3>(lc, a); break;
      default: launch_thin_out_s2<4>(lc, a); break;
    }
    IGM_POST_LAUNCH(lc);
    return IGM_OK;
  }
  const int ps = (a.transposed && a.stride > 1) ? a.stride : 1;
  const int OHp = cdiv(a.OH, ps), OWp = cdiv(a.OW, ps);
  const int64_t Mp = (int64_t)a.B * OHp * OWp;
  dim3 grid((unsigned)cdiv64(Mp, BM), (unsigned)cdiv(a.N, BN), (unsigned)(ps * ps));
  const double flops = 2.0 * a.B * a.OH * a.OW * (double)a.N * Ctot * a.KH * a.KW / (a.transposed ? a.stride * a.stride : 1);
  const double bytes = 4.0 * ((double)a.B * a.IH * a.IW * Ctot + (double)a.B * a.OH * a.OW * a.N * (a.add0 ? 2 : 1) +
                              (double)a.KH * a.KW * Ctot * a.N);
  ProfScope ps_(lc, a.kclass, flops, bytes);
  const bool vec = (a.C0 % BK == 0) && (a.C1 % BK == 0) && (a.N % 4 == 0) && (a.N0 % 4 == 0) && Ctot > 0;
  if (vec)
    conv_igemm_kernel<true><<<grid, 256, 0, lc.stream>>>(a);
  else
    conv_igemm_kernel<false><<<grid, 256, 0, lc.stream>>>(a);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_wgrad(const LaunchCtx& lc, const WgradArgs& a) {
  const int64_t npix = (int64_t)a.B * a.PH * a.PW;
  const bool same_grid = a.stride == 1 && a.dil == 1 && a.pad_w < 0 && a.PH == a.QH && a.PW == a.QW && a.KH == a.KW;
  if (same_grid && a.QC <= 4 && a.PC >= 32 && ((a.KH == 3 && a.pad == 1) || (a.KH == 1 && a.pad == 0))) {
    // few input channels gathered (Q = image), C_out enumerated (P = dY): stem 3x3 / 1x1 convs
    ProfScope ps_(lc, K_CONV_WGRAD, 2.0 * npix * (double)a.PC * a.QC * a.KH * a.KW, 4.0 * npix * (a.PC + a.QC));
    if (a.KH == 3) return launch_thin<3>(lc, a.QC, a.P, a.Q, a.grad, a.B, a.PH, a.PW, a.PC, a.sp, a.sq);
    return launch_thin<1>(lc, a.QC, a.P, a.Q, a.grad, a.B, a.PH, a.PW, a.PC, a.sp, a.sq);
  }
  if (same_grid && a.PC <= 4 && a.QC >= 32 && a.KH == 1 && a.pad == 0) {
    // few output channels (P = d_pred), wide input (Q): the final 1x1 conv
    ProfScope ps_(lc, K_CONV_WGRAD, 2.0 * npix * (double)a.PC * a.QC, 4.0 * npix * (a.PC + a.QC));
    return launch_thin<1>(lc, a.PC, a.Q, a.P, a.grad, a.B, a.PH, a.PW, a.QC, a.sq, a.sp);
  }
  static const bool thin_s2 = [] { const char* e = getenv("IGM_THIN_S2"); return !(e && e[0] == '0'); }();
  if (thin_s2 && a.stride == 2 && a.dil == 1 && (a.pad_w < 0 || a.pad_w == 1) && a.KH == 4 && a.KW == 4 && a.pad == 1 && a.QC >= 1 &&
      a.QC <= 4 &&
      a.PC >= 32 && a.QH == 2 * a.PH && a.QW == 2 * a.PW && a.PW <= 320 /* one band of thin rows must fit shared memory */) {
    // thin operand gathered on the fine grid (Q), wide operand enumerated on the coarse grid (P): 4x4 stride-2 layers
    ProfScope ps_(lc, K_CONV_WGRAD, 2.0 * npix * (double)a.PC * a.QC * 16, 4.0 * npix * (a.PC + 4.0 * a.QC));
    return launch_thin_s2(lc, a.QC, a.P, a.Q, a.grad, a.B, a.PH, a.PW, a.PC, a.sp, a.sq);
  }
  const int q_tiles = cdiv(a.QC, WB), p_tiles = cdiv(a.PC, WB);
  const int taps = a.KH * a.KW;
  // aim for ~4 CTAs per SM in total
  int64_t base = (int64_t)q_tiles * p_tiles * taps;
  int split = (int)((148 * 4 + base - 1) / base);
  int64_t max_split = cdiv64(npix, 256);
  if (split > max_split) split = (int)max_split;
  if (split < 1) split = 1;
  int pix_per_split = (int)cdiv64(npix, split);
  pix_per_split = cdiv(pix_per_split, WK) * WK;
  split = (int)cdiv64(npix, pix_per_split);
  dim3 grid((unsigned)split, (unsigned)(q_tiles * p_tiles), (unsigned)taps);
  ProfScope ps_(lc, K_CONV_WGRAD, 2.0 * npix * (double)a.PC * a.QC * taps,
                4.0 * (npix * (double)a.PC + (double)a.B * a.QH * a.QW * a.QC + (double)taps * a.PC * a.QC));
  const bool qv = a.QC % 4 == 0, pv = a.PC % 4 == 0;
  if (qv && pv)
    wgrad_kernel<true, true><<<grid, 256, 0, lc.stream>>>(a, pix_per_split, p_tiles);
  else if (qv)
    wgrad_kernel<true, false><<<grid, 256, 0, lc.stream>>>(a, pix_per_split, p_tiles);
  else if (pv)
    wgrad_kernel<false, true><<<grid, 256, 0, lc.stream>>>(a, pix_per_split, p_tiles);
  else
    wgrad_kernel<false, false><<<grid, 256, 0, lc.stream>>>(a, pix_per_split, p_tiles);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_colsum(const LaunchCtx& lc, const float* x, int64_t M, int N, float* out) {
  int splits = (int)cdiv64(M, 512);
  if (splits > 148 * 2) splits = 148 * 2;
  if (splits < 1) splits = 1;
  const int rows = (int)cdiv64(M, splits);
  dim3 grid((unsigned)cdiv64(M, rows), (unsigned)cdiv(N, 32));
  ProfScope ps_(lc, K_CONV_WGRAD, (double)M * N, 4.0 * M * N);
  colsum_kernel<<<grid, dim3(32, 8), 0, lc.stream>>>(x, M, N, out, rows);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_pack_weight(const LaunchCtx& lc, const float* src, float* dst, int taps, int K, int N,
                       int64_t sk, int64_t sn) {
  const int64_t total = (int64_t)taps * K * N;
  int blocks = (int)cdiv64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps_(lc, K_PACK, 0.0, 8.0 * total);
  pack_weight_kernel<<<blocks, 256, 0, lc.stream>>>(src, dst, taps, K, N, sk, sn);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
