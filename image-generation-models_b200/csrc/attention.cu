// Linear attention of the reference U-Net (src/models/ddpm.py:154-166), forward
// and backward, on the NHWC qkv tensor produced by the to_qkv 1x1 convolution.
//
//   q,k,v : [B, heads=4, d=32, n]     (channel = which*128 + head*32 + d)
//   k     <- softmax over n            (ddpm.py:160)
//   ctx   =  einsum('bhdn,bhen->bhde', k, v)        (:161)
//   out   =  einsum('bhde,bhdn->bhen', ctx, q)      (:162)
//
// One CTA per (batch, head): the 32x32 context lives in shared memory / registers and the
// spatial axis is streamed through in 32-row chunks.  <2 % of the U-Net FLOPs, so CUDA-core FMAs
// (fp32-exact) are used rather than tensor cores; the inner products are register-tiled (4x4 outer
// products, context rows/columns held in registers) so shared-memory traffic is one 128-bit
// broadcast load per 4..16 FMAs.
#include "common.cuh"

#include <algorithm>
#include <stdlib.h>

namespace igm {
namespace {

constexpr int D = kDimHead;          // 32
constexpr int QKV = 3 * kHeads * D;  // 384
constexpr int HD = kHeads * D;       // 128

constexpr int CH = 128;   // pixels per CTA: the spatial axis of one (batch, head) is split over ceil(n / CH) CTAs

// acc[i][j] += sum_{rows nn == grp (mod 4), nn < rows} X[nn][d0+i] * Y[nn][e0+j]   (4x4 register tile)
__device__ __forceinline__ void outer4x4(const float (*Xs)[D], const float (*Ys)[D], int rows, int grp, int d0, int e0,
                                         float (&acc)[4][4], float (&xsum)[4]) {
#pragma unroll 4
  for (int nn = grp; nn < rows; nn += 4) {
    const float4 x = *reinterpret_cast<const float4*>(&Xs[nn][d0]);
    const float4 y = *reinterpret_cast<const float4*>(&Ys[nn][e0]);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float2 y01 = make_float2(y.x, y.y), y23 = make_float2(y.z, y.w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xsum[i] += xv[i];
      const float2 xx = make_float2(xv[i], xv[i]);
      ffma2(*reinterpret_cast<float2*>(&acc[i][0]), xx, y01);   // acc[i][0..1] += x_i * y_{0,1}
      ffma2(*reinterpret_cast<float2*>(&acc[i][2]), xx, y23);
    }
  }
}

// reduce the four row-groups' 4x4 tiles through shared memory -> full[32][32] (+ row sums of X)
__device__ __forceinline__ void reduce_groups(float (&acc)[4][4], float (&xsum)[4], int grp, int t64, int d0, int e0,
                                              float (*part)[64][17], float (*full)[D + 1], float* rowsum) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) part[grp][t64][i * 4 + j] = acc[i][j];
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        full[d0 + i][e0 + j] = part[0][t64][i * 4 + j] + part[1][t64][i * 4 + j] + part[2][t64][i * 4 + j] +
                               part[3][t64][i * 4 + j];
  }
  __syncthreads();
  if (rowsum) {
    // every thread with e0 == 0 holds a partial row sum for rows d0..d0+3 of its row group
    if (e0 == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) part[grp][t64][i] = xsum[i];
    }
    __syncthreads();
    if (grp == 0 && e0 == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) rowsum[d0 + i] = part[0][t64][i] + part[1][t64][i] + part[2][t64][i] + part[3][t64][i];
    }
    __syncthreads();
  }
}

// rows [n0, n0+rows) of one 32-wide column block of an [n, ld] tensor -> smem tile (128-bit loads, zero fill).
// All R / 32 loads of a thread are issued before the first store: with a load -> store pair per iteration the compiler
// kept ONE load in flight per thread (ncu, round 2: 35 % of the ctx kernel's stall samples sat on the two STS.128).
template <int R = CH>
__device__ __forceinline__ void load_tile(float (*dst)[D], const float* __restrict__ src, int ld, int rows, int tid) {
  constexpr int IT = R * (D / 4) / 256;
  float4 v[IT];
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nn < rows) v[k] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)nn * ld + c4));
  }
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    *reinterpret_cast<float4*>(&dst[i >> 3][(i & 7) * 4]) = v[k];
  }
}
// two tiles at once (both tiles' loads in flight together)
template <int R = CH>
__device__ __forceinline__ void load_tile2(float (*d0)[D], const float* __restrict__ s0, int ld0, float (*d1)[D],
                                           const float* __restrict__ s1, int ld1, int rows, int tid) {
  constexpr int IT = R * (D / 4) / 256;
  float4 a[IT], b[IT];
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    a[k] = b[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nn < rows) {
      a[k] = __ldg(reinterpret_cast<const float4*>(s0 + (int64_t)nn * ld0 + c4));
      b[k] = __ldg(reinterpret_cast<const float4*>(s1 + (int64_t)nn * ld1 + c4));
    }
  }
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    *reinterpret_cast<float4*>(&d0[i >> 3][(i & 7) * 4]) = a[k];
    *reinterpret_cast<float4*>(&d1[i >> 3][(i & 7) * 4]) = b[k];
  }
}

// three tiles at once
template <int R = CH>
__device__ __forceinline__ void load_tile3(float (*d0)[D], const float* __restrict__ s0, int ld0, float (*d1)[D],
                                           const float* __restrict__ s1, int ld1, float (*d2)[D], const float* __restrict__ s2,
                                           int ld2, int rows, int tid) {
  constexpr int IT = R * (D / 4) / 256;
  float4 a[IT], b[IT], c[IT];
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    a[k] = b[k] = c[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nn < rows) {
      a[k] = __ldg(reinterpret_cast<const float4*>(s0 + (int64_t)nn * ld0 + c4));
      b[k] = __ldg(reinterpret_cast<const float4*>(s1 + (int64_t)nn * ld1 + c4));
      c[k] = __ldg(reinterpret_cast<const float4*>(s2 + (int64_t)nn * ld2 + c4));
    }
  }
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    *reinterpret_cast<float4*>(&d0[i >> 3][(i & 7) * 4]) = a[k];
    *reinterpret_cast<float4*>(&d1[i >> 3][(i & 7) * 4]) = b[k];
    *reinterpret_cast<float4*>(&d2[i >> 3][(i & 7) * 4]) = c[k];
  }
}

// the same tile from a bf16 hi / lo staging pair (value = hi + lo, 16 significant bits): `ld` in elements
template <int R = CH>
__device__ __forceinline__ void load_tile_hl(float (*dst)[D], const __nv_bfloat16* __restrict__ hi,
                                             const __nv_bfloat16* __restrict__ lo, int ld, int rows, int tid) {
  constexpr int IT = R * (D / 4) / 256;
  uint2 h[IT], l[IT];
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    h[k] = l[k] = make_uint2(0u, 0u);
    if (nn < rows) {
      h[k] = __ldg(reinterpret_cast<const uint2*>(hi + (int64_t)nn * ld + c4));
      l[k] = __ldg(reinterpret_cast<const uint2*>(lo + (int64_t)nn * ld + c4));
    }
  }
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int i = tid + k * 256;
    float4 v;
    v.x = __uint_as_float(h[k].x << 16) + __uint_as_float(l[k].x << 16);
    v.y = __uint_as_float(h[k].x & 0xffff0000u) + __uint_as_float(l[k].x & 0xffff0000u);
    v.z = __uint_as_float(h[k].y << 16) + __uint_as_float(l[k].y << 16);
    v.w = __uint_as_float(h[k].y & 0xffff0000u) + __uint_as_float(l[k].y & 0xffff0000u);
    *reinterpret_cast<float4*>(&dst[i >> 3][(i & 7) * 4]) = v;
  }
}

// The same tiles in two halves -- global loads into registers, registers to shared memory -- so that a kernel walking
// several chunks can have the NEXT chunk's loads in flight while it computes on the current one.
constexpr int kTileIt = CH * (D / 4) / 256;
__device__ __forceinline__ void tile_fetch(float4 (&v)[kTileIt], const float* __restrict__ src, int ld, int rows, int tid) {
#pragma unroll
  for (int k = 0; k < kTileIt; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (nn < rows) v[k] = __ldg(reinterpret_cast<const float4*>(src + (int64_t)nn * ld + c4));
  }
}
__device__ __forceinline__ void tile_fetch_hl(uint2 (&h)[kTileIt], uint2 (&l)[kTileIt], const __nv_bfloat16* __restrict__ hi,
                                              const __nv_bfloat16* __restrict__ lo, int ld, int rows, int tid) {
#pragma unroll
  for (int k = 0; k < kTileIt; ++k) {
    const int i = tid + k * 256;
    const int nn = i >> 3, c4 = (i & 7) * 4;
    h[k] = l[k] = make_uint2(0u, 0u);
    if (nn < rows) {
      h[k] = __ldg(reinterpret_cast<const uint2*>(hi + (int64_t)nn * ld + c4));
      l[k] = __ldg(reinterpret_cast<const uint2*>(lo + (int64_t)nn * ld + c4));
    }
  }
}
__device__ __forceinline__ void tile_store(float (*dst)[D], const float4 (&v)[kTileIt], int tid) {
#pragma unroll
  for (int k = 0; k < kTileIt; ++k) {
    const int i = tid + k * 256;
    *reinterpret_cast<float4*>(&dst[i >> 3][(i & 7) * 4]) = v[k];
  }
}
__device__ __forceinline__ void tile_store_hl(float (*dst)[D], const uint2 (&h)[kTileIt], const uint2 (&l)[kTileIt], int tid) {
#pragma unroll
  for (int k = 0; k < kTileIt; ++k) {
    const int i = tid + k * 256;
    float4 v;
    v.x = __uint_as_float(h[k].x << 16) + __uint_as_float(l[k].x << 16);
    v.y = __uint_as_float(h[k].x & 0xffff0000u) + __uint_as_float(l[k].x & 0xffff0000u);
    v.z = __uint_as_float(h[k].y << 16) + __uint_as_float(l[k].y << 16);
    v.w = __uint_as_float(h[k].y & 0xffff0000u) + __uint_as_float(l[k].y & 0xffff0000u);
    *reinterpret_cast<float4*>(&dst[i >> 3][(i & 7) * 4]) = v;
  }
}

constexpr int PART = 2 * D + D * D;   // per-chunk partial: max[32] | sum[32] | S[32][32]

// "last CTA of the group finishes the job": returns true in exactly one CTA of the gridDim.y CTAs that
// share blockIdx.x, after all the others have published their partials; the counter re-arms itself.
__device__ __forceinline__ bool last_chunk_done(unsigned int* counter, int tid) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(counter, 1u);
    s_last = (prev + 1 == gridDim.y);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// Pass 1, grid (B*heads, nsplit): softmax statistics of k over the pixels and the context
//   m[d] = max_n k[n][d],  l[d] = sum_n exp(k - m),  S[d][e] = sum_n exp(k[n][d] - m[d]) v[n][e],  ctx = S / l.
// A CTA walks `cpc` consecutive 128-pixel chunks: chunk-local statistics as before, merged into running (m, l, S) with the
// usual rescaling (fixed order -> deterministic): the one-chunk-per-CTA form paid a partial write, a gpu-scope fence and a
// share of the last-CTA merge per chunk.  (Prefetching the next chunk into registers was measured as well: 122 registers,
// two CTAs per SM, slower than letting five resident CTAs hide the load latency.)  With more than one CTA per
// (batch, head) the last one merges the CTAs' running statistics into ctx / kstat.
// HL: q / k / v come from the bf16 hi / lo staging pair of the to_qkv output (tensor-core attention path, no fp32 qkv)
template <bool HL>
__global__ void __launch_bounds__(256, 4) linattn_ctx_kernel(const float* __restrict__ qkv, const __nv_bfloat16* __restrict__ q_hi,
                                                             const __nv_bfloat16* __restrict__ q_lo, float* __restrict__ part_ws,
                                                             unsigned int* __restrict__ counters, float* __restrict__ ctx,
                                                             float* __restrict__ kstat, int N, int ld, int koff, int cpc) {
  __shared__ __align__(16) float buf[2 * CH * D];   // k | v tiles, later the reduction scratch
  float (*Xs)[D] = reinterpret_cast<float (*)[D]>(buf);
  float (*Ys)[D] = reinterpret_cast<float (*)[D]>(buf + CH * D);
  float (*part)[64][17] = reinterpret_cast<float (*)[64][17]>(buf);
  __shared__ float ctxs[D][D + 1];
  __shared__ float r_ctx[D][D + 1];                 // running S
  __shared__ float red[8][D];
  __shared__ float s_kmax[D], s_ksum[D], r_max[D], r_sum[D];
  pdl_wait();
  const int bh = blockIdx.x, b = bh / kHeads, h = bh % kHeads;
  const int nsplit = gridDim.y;
  const int nchunks = (N + CH - 1) / CH;
  const int c_begin = blockIdx.y * cpc, c_end = min(nchunks, c_begin + cpc);
  const int tid = threadIdx.x;
  // rows of `ld` floats, k at column koff, v right behind it: [M, 384] q | k | v (koff = 128) or a compact [M, 256] k | v
  const int64_t boff0 = (int64_t)b * N * ld + koff + h * D;
  auto load = [&](int c) {
    const int64_t boff = boff0 + (int64_t)c * CH * ld;
    const int rows = min(CH, N - c * CH);
    if (HL) {
      load_tile_hl(Xs, q_hi + boff, q_lo + boff, ld, rows, tid);
      load_tile_hl(Ys, q_hi + boff + HD, q_lo + boff + HD, ld, rows, tid);
    } else {
      load_tile2(Xs, qkv + boff, ld, Ys, qkv + boff + HD, ld, rows, tid);
    }
  };
  load(c_begin);
  __syncthreads();
  const int grp = tid >> 6, t64 = tid & 63;
  const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
  for (int c = c_begin; c < c_end; ++c) {
    const int rows = min(CH, N - c * CH);
    const bool more = c + 1 < c_end;
    {
      const int d = tid & 31, r = tid >> 5;
      float m = -INFINITY;
      for (int nn = r; nn < rows; nn += 8) m = fmaxf(m, Xs[nn][d]);
      red[r][d] = m;
      __syncthreads();
      if (tid < D) {
        float t = red[0][tid];
#pragma unroll
        for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i][tid]);
        s_kmax[tid] = t;
      }
      __syncthreads();
      for (int nn = r; nn < rows; nn += 8) Xs[nn][d] = expf(Xs[nn][d] - s_kmax[d]);
      __syncthreads();
    }
    float acc[4][4], xsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xsum[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    outer4x4(Xs, Ys, rows, grp, d0, e0, acc, xsum);
    __syncthreads();   // `part` aliases the tiles
    reduce_groups(acc, xsum, grp, t64, d0, e0, part, ctxs, s_ksum);
    // running (m, l, S) <- merge with this chunk's
    if (c == c_begin) {
      for (int i = tid; i < D * D; i += 256) r_ctx[i >> 5][i & 31] = ctxs[i >> 5][i & 31];
      if (tid < D) { r_max[tid] = s_kmax[tid]; r_sum[tid] = s_ksum[tid]; }
    } else {
      for (int i = tid; i < D * D; i += 256) {
        const int d = i >> 5, e = i & 31;
        const float mo = r_max[d], mc = s_kmax[d], mn = fmaxf(mo, mc);
        r_ctx[d][e] = fmaf(r_ctx[d][e], expf(mo - mn), ctxs[d][e] * expf(mc - mn));
      }
      __syncthreads();   // every reader of r_max is done
      if (tid < D) {
        const float mo = r_max[tid], mc = s_kmax[tid], mn = fmaxf(mo, mc);
        r_sum[tid] = fmaf(r_sum[tid], expf(mo - mn), s_ksum[tid] * expf(mc - mn));
        r_max[tid] = mn;
      }
    }
    __syncthreads();
    if (more) {
      load(c + 1);
      __syncthreads();
    }
  }
  float* cdst = ctx + (int64_t)bh * D * D;
  float* kdst = kstat + (int64_t)bh * D * 2;
  if (nsplit == 1) {
    for (int i = tid; i < D * D; i += 256) cdst[i] = r_ctx[i >> 5][i & 31] / r_sum[i >> 5];
    if (tid < D) { kdst[tid * 2] = r_max[tid]; kdst[tid * 2 + 1] = r_sum[tid]; }
    return;
  }
  float* dst = part_ws + ((int64_t)bh * nsplit + blockIdx.y) * PART;
  if (tid < D) { dst[tid] = r_max[tid]; dst[D + tid] = r_sum[tid]; }
  for (int i = tid; i < D * D; i += 256) dst[2 * D + i] = r_ctx[i >> 5][i & 31];
  if (!last_chunk_done(counters + bh, tid)) return;
  // merge: global max / denominator, rescaled sum of the partial contexts (fixed order -> deterministic)
  const float* src = part_ws + (int64_t)bh * nsplit * PART;
  float (*s_scale)[D] = reinterpret_cast<float (*)[D]>(buf);   // [nsplit][32]
  if (tid < D) {
    float m = -INFINITY;
    for (int sp = 0; sp < nsplit; ++sp) m = fmaxf(m, __ldcg(src + (int64_t)sp * PART + tid));
    float l = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
      const float sc = expf(__ldcg(src + (int64_t)sp * PART + tid) - m);
      s_scale[sp][tid] = sc;
      l = fmaf(__ldcg(src + (int64_t)sp * PART + D + tid), sc, l);
    }
    s_kmax[tid] = m;
    s_ksum[tid] = l;
    kdst[tid * 2] = m;
    kdst[tid * 2 + 1] = l;
  }
  __syncthreads();
  for (int i = tid; i < D * D; i += 256) {
    const int d = i >> 5;
    float a = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) a = fmaf(__ldcg(src + (int64_t)sp * PART + 2 * D + i), s_scale[sp][d], a);
    cdst[i] = a / s_ksum[d];
  }
}

constexpr int kMaxSplit = 64;   // n <= 8192 pixels per image (s_scale fits the tile buffer)

// Pass 2, grid (B*heads, nsplit): out[n][e] = sum_d ctx[d][e] * q[n][d] on this CTA's pixel chunk
__global__ void __launch_bounds__(256) linattn_out_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx,
                                                          float* __restrict__ out, int N,
                                                          __nv_bfloat16* __restrict__ out_hi,
                                                          __nv_bfloat16* __restrict__ out_lo) {
  __shared__ __align__(16) float Xs[CH][D];
  pdl_wait();
  const int bh = blockIdx.x, b = bh / kHeads, h = bh % kHeads;
  const int n0 = blockIdx.y * CH, rows = min(CH, N - n0);
  const int tid = threadIdx.x;
  load_tile(Xs, qkv + ((int64_t)b * N + n0) * QKV + h * D, QKV, rows, tid);
  // lane e keeps column e of the context in registers
  const int e = tid & 31, r = tid >> 5;
  float2 col[D / 2];
#pragma unroll
  for (int dd = 0; dd < D; dd += 2)
    col[dd >> 1] = make_float2(__ldg(ctx + (int64_t)bh * D * D + dd * D + e), __ldg(ctx + (int64_t)bh * D * D + (dd + 1) * D + e));
  __syncthreads();
  for (int nn = r; nn < rows; nn += 8) {
    float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int d4 = 0; d4 < D; d4 += 4) {
      const float4 x = *reinterpret_cast<const float4*>(&Xs[nn][d4]);
      ffma2(a2, col[d4 >> 1], make_float2(x.x, x.y));
      ffma2(a2, col[(d4 >> 1) + 1], make_float2(x.z, x.w));
    }
    const float a = a2.x + a2.y;
    const int64_t o = ((int64_t)b * N + n0 + nn) * HD + h * D + e;
    if (out) out[o] = a;
    if (out_hi) {
      const __nv_bfloat16 hv = __float2bfloat16_rn(a);
      out_hi[o] = hv;
      out_lo[o] = __float2bfloat16_rn(a - __bfloat162float(hv));
    }
  }
}

// Backward pass 1, grid (B*heads, nsplit): dctx[d][e] = sum_n q[n][d] * dO[n][e]; the last CTA of each
// (batch, head) sums the chunk partials into dctx_full.
template <bool HL>
__global__ void __launch_bounds__(256) linattn_bwd_dctx_kernel(const float* __restrict__ qkv,
                                                               const float* __restrict__ d_out,
                                                               const __nv_bfloat16* __restrict__ q_hi,
                                                               const __nv_bfloat16* __restrict__ q_lo,
                                                               const __nv_bfloat16* __restrict__ d_hi,
                                                               const __nv_bfloat16* __restrict__ d_lo,
                                                               float* __restrict__ part_ws,
                                                               unsigned int* __restrict__ counters,
                                                               float* __restrict__ dctx_full, int N, int cpc) {
  __shared__ __align__(16) float buf[2 * CH * D];   // q | dO tiles, later the reduction scratch
  float (*Xs)[D] = reinterpret_cast<float (*)[D]>(buf);
  float (*Ys)[D] = reinterpret_cast<float (*)[D]>(buf + CH * D);
  float (*part)[64][17] = reinterpret_cast<float (*)[64][17]>(buf);
  __shared__ float full[D][D + 1];
  pdl_wait();
  const int bh = blockIdx.x, b = bh / kHeads, h = bh % kHeads;
  const int nsplit = gridDim.y;
  const int nchunks = (N + CH - 1) / CH;
  const int c_begin = blockIdx.y * cpc, c_end = min(nchunks, c_begin + cpc);
  const int tid = threadIdx.x;
  const int grp = tid >> 6, t64 = tid & 63;
  const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
  float acc[4][4], xsum[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  // a CTA walks `cpc` consecutive chunks and keeps accumulating its 4x4 register tiles (see linattn_ctx_kernel)
  for (int c = c_begin; c < c_end; ++c) {
    const int n0 = c * CH, rows = min(CH, N - n0);
    if (HL) {
      const int64_t qo = ((int64_t)b * N + n0) * QKV + h * D, dof = ((int64_t)b * N + n0) * HD + h * D;
      load_tile_hl(Xs, q_hi + qo, q_lo + qo, QKV, rows, tid);
      load_tile_hl(Ys, d_hi + dof, d_lo + dof, HD, rows, tid);
    } else {
      load_tile2(Xs, qkv + ((int64_t)b * N + n0) * QKV + h * D, QKV, Ys, d_out + ((int64_t)b * N + n0) * HD + h * D, HD, rows, tid);
    }
    __syncthreads();
    outer4x4(Xs, Ys, rows, grp, d0, e0, acc, xsum);
    __syncthreads();   // the tiles are overwritten by the next chunk / aliased by `part`
  }
  reduce_groups(acc, xsum, grp, t64, d0, e0, part, full, nullptr);
  float* fdst = dctx_full + (int64_t)bh * D * D;
  if (nsplit == 1) {
    for (int i = tid; i < D * D; i += 256) fdst[i] = full[i >> 5][i & 31];
    return;
  }
  float* dst = part_ws + ((int64_t)bh * nsplit + blockIdx.y) * (D * D);
  for (int i = tid; i < D * D; i += 256) dst[i] = full[i >> 5][i & 31];
  if (!last_chunk_done(counters + bh, tid)) return;
  const float* src = part_ws + (int64_t)bh * nsplit * (D * D);
  for (int i = tid; i < D * D; i += 256) {
    float a = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) a += __ldcg(src + (int64_t)sp * (D * D) + i);
    fdst[i] = a;
  }
}

// Backward pass 2, grid (B*heads, nsplit): per-pixel gradients of q, k, v on this CTA's chunk
//   dq[n][d] = sum_e ctx[d][e] dO[n][e];  dv[n][e] = sum_d dctx[d][e] p[n][d]
//   dk[n][d] = p[n][d] * (sum_e dctx[d][e] v[n][e] - sum_e dctx[d][e] ctx[d][e]),  p = softmax_n(k)
// Three sweeps over the staged rows, each with ONE 32-entry row/column of ctx or dctx in registers.
__global__ void __launch_bounds__(256, 3) linattn_bwd_rows_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx,
                                                               const float* __restrict__ kstat,
                                                               const float* __restrict__ d_out,
                                                               const float* __restrict__ dctx_full,
                                                               float* __restrict__ d_qkv, int N,
                                                               __nv_bfloat16* __restrict__ d_hi,
                                                               __nv_bfloat16* __restrict__ d_lo) {
  constexpr int SUB = 64;                      // rows staged at a time (three tiles must fit 48 KB of static smem)
  __shared__ __align__(16) float Xs[SUB][D];   // softmax(k)
  __shared__ __align__(16) float Ys[SUB][D];   // dO
  __shared__ __align__(16) float Vs[SUB][D];   // v
  __shared__ float ctxs[D][D + 1];
  __shared__ float dctxs[D][D + 1];
  pdl_wait();
  const int bh = blockIdx.x, b = bh / kHeads, h = bh % kHeads;
  const int tid = threadIdx.x;
  for (int i = tid; i < D * D; i += 256) {
    ctxs[i >> 5][i & 31] = __ldg(ctx + (int64_t)bh * D * D + i);
    dctxs[i >> 5][i & 31] = __ldg(dctx_full + (int64_t)bh * D * D + i);
  }
  __syncthreads();
  const int c = tid & 31, r = tid >> 5;
  float cdot = 0.f;
#pragma unroll
  for (int e = 0; e < D; ++e) cdot = fmaf(dctxs[c][e], ctxs[c][e], cdot);
  const float kmax = kstat[((int64_t)bh * D + c) * 2 + 0], kinv = 1.f / kstat[((int64_t)bh * D + c) * 2 + 1];
  const int qcol = h * D, kcol = HD + h * D, vcol = 2 * HD + h * D;
  const int n_end = min((int)(blockIdx.y + 1) * CH, N);
  for (int n0 = blockIdx.y * CH; n0 < n_end; n0 += SUB) {
    const int rows = min(SUB, n_end - n0);
    const float* base = qkv + ((int64_t)b * N + n0) * QKV;
    load_tile3<SUB>(Xs, base + kcol, QKV, Vs, base + vcol, QKV, Ys, d_out + ((int64_t)b * N + n0) * HD + h * D, HD, rows, tid);
    __syncthreads();
    for (int nn = r; nn < rows; nn += 8) Xs[nn][c] = expf(Xs[nn][c] - kmax) * kinv;
    __syncthreads();
    const int64_t obase = ((int64_t)b * N + n0) * QKV;
    // gradient element -> fp32 tensor and/or the bf16 hi/lo staging copy the to_qkv backward convs read
    auto emit = [&](int64_t o, float val) {
      if (d_qkv) d_qkv[o] = val;
      if (d_hi) {
        const __nv_bfloat16 hv = __float2bfloat16_rn(val);
        d_hi[o] = hv;
        d_lo[o] = __float2bfloat16_rn(val - __bfloat162float(hv));
      }
    };
    float2 w[D / 2];
    // 32-term dot product of the register-resident row/column w with one staged row (packed FMAs: two terms per issue)
    auto dot = [&](const float* row) {
      float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j4 = 0; j4 < D; j4 += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(row + j4);
        ffma2(a2, w[j4 >> 1], make_float2(g4.x, g4.y));
        ffma2(a2, w[(j4 >> 1) + 1], make_float2(g4.z, g4.w));
      }
      return a2.x + a2.y;
    };
    // dq = ctx[c][:] . dO[n][:]
#pragma unroll
    for (int j = 0; j < D; j += 2) w[j >> 1] = make_float2(ctxs[c][j], ctxs[c][j + 1]);
    for (int nn = r; nn < rows; nn += 8) emit(obase + (int64_t)nn * QKV + qcol + c, dot(&Ys[nn][0]));
    // dk = p * (dctx[c][:] . v[n][:] - cdot)
#pragma unroll
    for (int j = 0; j < D; j += 2) w[j >> 1] = make_float2(dctxs[c][j], dctxs[c][j + 1]);
    for (int nn = r; nn < rows; nn += 8) emit(obase + (int64_t)nn * QKV + kcol + c, Xs[nn][c] * (dot(&Vs[nn][0]) - cdot));
    // dv = dctx[:][c] . p[n][:]
#pragma unroll
    for (int j = 0; j < D; j += 2) w[j >> 1] = make_float2(dctxs[j][c], dctxs[j + 1][c]);
    for (int nn = r; nn < rows; nn += 8) emit(obase + (int64_t)nn * QKV + vcol + c, dot(&Xs[nn][0]));
    __syncthreads();
  }
}

// Inference shortcut for n >= 2C: out = ctx^T q and y = W_out out + b collapse into one per-image C x C matrix
//   M_b = A_b W_q,   A_b[c][(h,d)] = sum_e W_out[c][(h,e)] ctx_b[h][d][e]        (y[:, n] = M_b LN(x)[:, n] + b)
// so the attention output never exists; the to_out conv then runs with per-image weights M_b on LN(x).
// grid (B, C / 32): 32 rows of M_b per CTA; emitted as bf16 hi / lo rows [b * C + c][c'] (K-major weight tile rows).
__global__ void __launch_bounds__(256) linattn_mb_kernel(const float* __restrict__ ctx, const float* __restrict__ w_out,
                                                         const float* __restrict__ w_q, int C,
                                                         __nv_bfloat16* __restrict__ mb_hi, __nv_bfloat16* __restrict__ mb_lo) {
  extern __shared__ __align__(16) float msm[];
  float* s_ctx = msm;               // [4 * 32 rows][33]  (padded: the 8 lanes of a row group read 8 different d rows)
  float* s_wo = msm + 4224;         // [32 rows][128]
  float* s_a = msm + 4224 + 4096;   // [32 rows][128]
  pdl_wait();
  const int b = blockIdx.x, c0 = blockIdx.y * 32, tid = threadIdx.x;
  for (int i = tid; i < 4096; i += 256) {
    s_ctx[(i >> 5) * 33 + (i & 31)] = __ldg(ctx + (int64_t)b * 4096 + i);
    s_wo[i] = __ldg(w_out + (int64_t)(c0 + (i >> 7)) * HD + (i & 127));
  }
  __syncthreads();
  // A[c][(h,d)]: warp -> 4 rows, lane -> d; W_out values are warp-wide broadcasts, ctx rows are 33 floats apart
  // (the lane's ctx row is read once per head and kept in registers for the warp's four rows, W_out arrives as 128-bit
  // broadcasts: 64 shared-memory loads per head and warp instead of 256; same summation order as the scalar form)
  {
    const int w = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int h = 0; h < kHeads; ++h) {
      float cx[D];
      const float* cxp = s_ctx + (h * D + lane) * 33;
#pragma unroll
      for (int e = 0; e < D; ++e) cx[e] = cxp[e];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = w * 4 + rr;
        const float4* wo = reinterpret_cast<const float4*>(s_wo + r * HD + h * D);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int e4 = 0; e4 < D / 4; ++e4) {
          const float4 wv = wo[e4];
          a0 = fmaf(wv.x, cx[4 * e4 + 0], a0);
          a1 = fmaf(wv.y, cx[4 * e4 + 1], a1);
          a0 = fmaf(wv.z, cx[4 * e4 + 2], a0);
          a1 = fmaf(wv.w, cx[4 * e4 + 3], a1);
        }
        s_a[r * HD + h * D + lane] = a0 + a1;
      }
    }
  }
  __syncthreads();
  // M[c][c'] = sum_k A[c][k] Wq[k][c']: thread -> column c' (coalesced Wq rows) x 4 rows, k in steps of 4
  for (int o = tid; o < 8 * C; o += 256) {
    const int cp = o % C, rg = o / C;       // rows rg*4 .. rg*4+3
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < HD; k += 4) {
      float wq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) wq[u] = __ldg(w_q + (int64_t)(k + u) * C + cp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 av = *reinterpret_cast<const float4*>(s_a + (rg * 4 + i) * HD + k);
        acc[i] = fmaf(av.x, wq[0], acc[i]);
        acc[i] = fmaf(av.y, wq[1], acc[i]);
        acc[i] = fmaf(av.z, wq[2], acc[i]);
        acc[i] = fmaf(av.w, wq[3], acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t dst = ((int64_t)b * C + c0 + rg * 4 + i) * C + cp;
      const __nv_bfloat16 hv = __float2bfloat16_rn(acc[i]);
      mb_hi[dst] = hv;
      mb_lo[dst] = __float2bfloat16_rn(acc[i] - __bfloat162float(hv));
    }
  }
}


// ---- tensor-core attention path (training, n >= 1024 pixels): every per-pixel contraction of the block is a 1x1
// convolution whose 128 x 128 weight matrix differs per image and is block-diagonal over the four heads, i.e. exactly what
// conv_tc_kernel runs with a tc_plan_img plan:
//     out = conv(q; ctx)   dq = conv(dO; ctx^T)   dv = conv(p; dctx)   T = conv(v; dctx^T)   dk = p * (T - c)
// (reference ddpm.py:161-162 and their autograd).  The kernels below are the glue that stays on CUDA cores.

// Per-image block-diagonal weight matrix Wt[b][row][col] (row = output channel, col = input channel, K-major rows, what
// the conv engine's weight descriptor reads) from m[b][h][d][e]:
//   transpose = 0:  Wt[(h,e)][(h,d)] = m[d][e]      (out = conv(q; ctx), dv = conv(p; dctx))
//   transpose = 1:  Wt[(h,d)][(h,e)] = m[d][e]      (dq = conv(dO; ctx^T), T = conv(v; dctx^T))
// grid (B), 256 threads; cc (nullable): cc[b][(h,d)] = sum_e m[d][e] * m2[d][e]  (the softmax-normaliser term of dk)
__global__ void __launch_bounds__(256) linattn_wt_kernel(const float* __restrict__ m, int transpose,
                                                         __nv_bfloat16* __restrict__ w_hi, __nv_bfloat16* __restrict__ w_lo,
                                                         const float* __restrict__ m2, float* __restrict__ cc) {
  __shared__ float sm[kHeads][D][D + 1];
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < kHeads * D * D; i += 256) sm[i >> 10][(i >> 5) & 31][i & 31] = __ldg(m + (int64_t)b * kHeads * D * D + i);
  __syncthreads();
  for (int i = tid; i < HD * HD / 2; i += 256) {          // two adjacent columns per thread
    const int row = i / (HD / 2), col = (i - row * (HD / 2)) * 2;
    const int hr = row >> 5, hc = col >> 5;
    float v0 = 0.f, v1 = 0.f;
    if (hr == hc) {
      if (transpose) { v0 = sm[hr][row & 31][col & 31]; v1 = sm[hr][row & 31][(col + 1) & 31]; }
      else { v0 = sm[hr][col & 31][row & 31]; v1 = sm[hr][(col + 1) & 31][row & 31]; }
    }
    uint32_t h, l;
    split_pair(v0, v1, h, l);
    const int64_t o = ((int64_t)b * HD + row) * HD + col;
    *reinterpret_cast<uint32_t*>(w_hi + o) = h;
    *reinterpret_cast<uint32_t*>(w_lo + o) = l;
  }
  if (cc && tid < HD) {
    const int h = tid >> 5, d = tid & 31;
    const float* r2 = m2 + ((int64_t)b * kHeads + h) * D * D + d * D;
    float a = 0.f;
#pragma unroll
    for (int e = 0; e < D; ++e) a = fmaf(sm[h][d][e], __ldg(r2 + e), a);
    cc[(int64_t)b * HD + tid] = a;
  }
}

// p = softmax_n(k) as a bf16 hi / lo pair [M, 128] (operand of dv = conv(p; dctx) and factor of dk), from the k third
// of the to_qkv staging pair and the (max, sum) statistics the forward kept.  One thread per 4 channels of a pixel.
__global__ void __launch_bounds__(256) linattn_p_kernel(const __nv_bfloat16* __restrict__ q_hi, const __nv_bfloat16* __restrict__ q_lo,
                                                        const float* __restrict__ kstat, __nv_bfloat16* __restrict__ p_hi,
                                                        __nv_bfloat16* __restrict__ p_lo, int64_t M, int N) {
  pdl_wait();
  const int64_t total = M * (HD / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i >> 5;
    const int c = (int)(i & 31) * 4;
    const int64_t b = pix / N;
    const int64_t src = pix * QKV + HD + c;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(q_hi + src));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(q_lo + src));
    const float k0 = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
    const float k1 = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
    const float k2 = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
    const float k3 = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
    const float4 s01 = __ldg(reinterpret_cast<const float4*>(kstat + (b * HD + c) * 2));       // (max, sum) x 2 channels
    const float4 s23 = __ldg(reinterpret_cast<const float4*>(kstat + (b * HD + c + 2) * 2));
    float4 o;
    o.x = expf(k0 - s01.x) / s01.y;
    o.y = expf(k1 - s01.z) / s01.w;
    o.z = expf(k2 - s23.x) / s23.y;
    o.w = expf(k3 - s23.z) / s23.w;
    store_split4(p_hi, p_lo, pix * HD + c, o);
  }
}

// dk = p * (T - c) into the k third of the d(qkv) staging pair (pitch 384): T = conv(v; dctx^T) fp32 [M, 128]
__global__ void __launch_bounds__(256) linattn_dk_kernel(const __nv_bfloat16* __restrict__ p_hi, const __nv_bfloat16* __restrict__ p_lo,
                                                         const float* __restrict__ T, const float* __restrict__ cc,
                                                         __nv_bfloat16* __restrict__ d_hi, __nv_bfloat16* __restrict__ d_lo,
                                                         int64_t M, int N) {
  pdl_wait();
  const int64_t total = M * (HD / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i >> 5;
    const int c = (int)(i & 31) * 4;
    const int64_t b = pix / N;
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p_hi + pix * HD + c));
    const uint2 l = __ldg(reinterpret_cast<const uint2*>(p_lo + pix * HD + c));
    const float4 t = __ldg(reinterpret_cast<const float4*>(T + pix * HD + c));
    const float4 cv = __ldg(reinterpret_cast<const float4*>(cc + b * HD + c));
    float4 o;
    o.x = (__uint_as_float(h.x << 16) + __uint_as_float(l.x << 16)) * (t.x - cv.x);
    o.y = (__uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u)) * (t.y - cv.y);
    o.z = (__uint_as_float(h.y << 16) + __uint_as_float(l.y << 16)) * (t.z - cv.z);
    o.w = (__uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u)) * (t.w - cv.w);
    store_split4(d_hi, d_lo, pix * QKV + HD + c, o);
  }
}

}  // namespace

// scratch layout: kCtrCap chunk counters (zero on entry, self re-arming) | [B*heads][32*32] dctx | chunk partials.
// The counter block has a FIXED size: were it sized by the batch of the call, the dctx / partial regions of a small batch
// would land on counters a later, larger batch of the same context expects to find zero.
// chunks per CTA of the statistics / context kernel: up to 4 (a whole image where it has no more: no cross-CTA merge at
// all), halved while the launch would leave fewer than three CTAs per SM.  Measured: 4 chunks per CTA +0.7 % on the CIFAR-10
// sampler; unconditionally it cost CelebA-64 at 32 images 0.4 % (its 32x32 / 16x16 blocks were left with 256 / 128 CTAs).
// IGM_ATTN_CPC overrides (1 = one chunk per CTA as in round 1).
static int ctx_cpc(int nchunks, int B) {
  static const int env = [] { const char* e = getenv("IGM_ATTN_CPC"); return e ? atoi(e) : 0; }();
  int cpc = env > 0 ? env : 4;
  if (cpc > nchunks) cpc = nchunks;
  if (env <= 0)
    while (cpc > 1 && B * kHeads * cdiv(nchunks, cpc) < 148 * 3) cpc >>= 1;
  return cpc;
}
constexpr int kCtrCap = 8192;   // (batch, head) pairs: B <= 2048
int linattn_ws_floats(int B, int n) { return kCtrCap + B * kHeads * (D * D + cdiv(n, CH) * PART); }

int launch_linattn_forward(const LaunchCtx& lc, const float* qkv, float* out, float* ctx, float* kstat,
                           int B, int n, float* ws, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
  const int nsplit = cdiv(n, CH);
  if (nsplit > kMaxSplit) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: more than 8192 pixels per image");
  ProfScope ps_(lc, K_ATTN, 4.0 * B * kHeads * (double)n * D * D, 4.0 * B * (double)n * (QKV + HD));
  if (B * kHeads > kCtrCap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: batch too large");
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* parts = ws + kCtrCap + (int64_t)B * kHeads * (D * D);
  { cudaError_t le_ = launch_pdl(linattn_ctx_kernel<false>, dim3(B * kHeads, cdiv(nsplit, ctx_cpc(nsplit, B))), dim3(256), (size_t)0, lc.stream, qkv, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr, parts, counters, ctx, kstat, n, QKV, HD, ctx_cpc(nsplit, B)); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  { cudaError_t le_ = launch_pdl(linattn_out_kernel, dim3(B * kHeads, nsplit), dim3(256), (size_t)0, lc.stream, qkv, ctx, out, n, out_hi, out_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// statistics + context only (the first half of launch_linattn_forward)
int launch_linattn_ctx(const LaunchCtx& lc, const float* qkv, float* ctx, float* kstat, int B, int n, float* ws, bool kv_only) {
  const int nsplit = cdiv(n, CH);
  if (nsplit > kMaxSplit) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: more than 8192 pixels per image");
  ProfScope ps_(lc, K_ATTN, 2.0 * B * kHeads * (double)n * D * D, 4.0 * B * (double)n * 2 * HD);
  if (B * kHeads > kCtrCap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: batch too large");
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* parts = ws + kCtrCap + (int64_t)B * kHeads * (D * D);
  { cudaError_t le_ = launch_pdl(linattn_ctx_kernel<false>, dim3(B * kHeads, cdiv(nsplit, ctx_cpc(nsplit, B))), dim3(256), (size_t)0, lc.stream, qkv, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr, parts, counters, ctx, kstat, n, kv_only ? 2 * HD : QKV, kv_only ? 0 : HD, ctx_cpc(nsplit, B)); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_linattn_mb(const LaunchCtx& lc, const float* ctx, const float* w_out, const float* w_q, int B, int C,
                      __nv_bfloat16* mb_hi, __nv_bfloat16* mb_lo) {
  if (C % 32 != 0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linattn_mb: C must be a multiple of 32");
  const size_t smem = (4224 + 2 * 4096) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(linattn_mb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr = true;
  }
  ProfScope ps_(lc, K_ATTN, 2.0 * B * C * (double)HD * (D + C), 4.0 * B * (4096.0 + C * C));
  { cudaError_t le_ = launch_pdl(linattn_mb_kernel, dim3(B, C / 32), dim3(256), smem, lc.stream, ctx, w_out, w_q, C, mb_hi, mb_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_linattn_backward(const LaunchCtx& lc, const float* qkv, const float* ctx, const float* kstat,
                            const float* d_out, float* d_qkv, int B, int n, float* ws, __nv_bfloat16* d_hi,
                            __nv_bfloat16* d_lo) {
  const int nsplit = cdiv(n, CH);
  if (nsplit > kMaxSplit) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: more than 8192 pixels per image");
  ProfScope ps_(lc, K_ATTN, 8.0 * B * kHeads * (double)n * D * D, 4.0 * B * (double)n * (2 * QKV + HD));
  if (B * kHeads > kCtrCap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: batch too large");
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* dctx = ws + kCtrCap;
  float* parts = ws + kCtrCap + (int64_t)B * kHeads * (D * D);
  { cudaError_t le_ = launch_pdl(linattn_bwd_dctx_kernel<false>, dim3(B * kHeads, cdiv(nsplit, ctx_cpc(nsplit, B))), dim3(256), (size_t)0, lc.stream, qkv, d_out, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)nullptr, parts, counters, dctx, n, ctx_cpc(nsplit, B)); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  { cudaError_t le_ = launch_pdl(linattn_bwd_rows_kernel, dim3(B * kHeads, nsplit), dim3(256), (size_t)0, lc.stream, qkv, ctx, kstat, d_out, dctx, d_qkv, n, d_hi, d_lo); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// ---- tensor-core attention path: launches of the CUDA-core glue kernels ------------------------------------------------
#define IGM_LAUNCH_PDL(...) do { cudaError_t le_ = launch_pdl(__VA_ARGS__); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); IGM_POST_LAUNCH(lc); } while (0)

// statistics + context from the k / v thirds of the bf16 hi / lo staging pair of the to_qkv output ([M, 384])
int launch_linattn_ctx_hl(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, float* ctx, float* kstat,
                          int B, int n, float* ws) {
  const int nsplit = cdiv(n, CH);
  if (nsplit > kMaxSplit) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: more than 8192 pixels per image");
  if (B * kHeads > kCtrCap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: batch too large");
  ProfScope ps_(lc, K_ATTN, 2.0 * B * kHeads * (double)n * D * D, 2.0 * B * (double)n * 2 * HD * 2);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* parts = ws + kCtrCap + (int64_t)B * kHeads * (D * D);
  IGM_LAUNCH_PDL(linattn_ctx_kernel<true>, dim3(B * kHeads, cdiv(nsplit, ctx_cpc(nsplit, B))), dim3(256), (size_t)0, lc.stream, (const float*)nullptr, q_hi, q_lo,
                 parts, counters, ctx, kstat, n, QKV, HD, ctx_cpc(nsplit, B));
  return IGM_OK;
}

// dctx[b][h][d][e] = sum_n q[n][d] dO[n][e] from the staging pairs; result at linattn_dctx_ptr(ws)
float* linattn_dctx_ptr(float* ws) { return ws + kCtrCap; }
int launch_linattn_dctx_hl(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, const __nv_bfloat16* d_hi,
                           const __nv_bfloat16* d_lo, int B, int n, float* ws) {
  const int nsplit = cdiv(n, CH);
  if (nsplit > kMaxSplit) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: more than 8192 pixels per image");
  if (B * kHeads > kCtrCap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "linear attention: batch too large");
  ProfScope ps_(lc, K_ATTN, 2.0 * B * kHeads * (double)n * D * D, 2.0 * B * (double)n * 2 * HD * 2);
  unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
  float* dctx = ws + kCtrCap;
  float* parts = ws + kCtrCap + (int64_t)B * kHeads * (D * D);
  IGM_LAUNCH_PDL(linattn_bwd_dctx_kernel<true>, dim3(B * kHeads, cdiv(nsplit, ctx_cpc(nsplit, B))), dim3(256), (size_t)0, lc.stream, (const float*)nullptr,
                 (const float*)nullptr, q_hi, q_lo, d_hi, d_lo, parts, counters, dctx, n, ctx_cpc(nsplit, B));
  return IGM_OK;
}

int launch_linattn_wt(const LaunchCtx& lc, const float* m, int transpose, int B, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo,
                      const float* m2, float* cc) {
  ProfScope ps_(lc, K_ATTN, 0.0, 4.0 * B * (4096.0 + HD * HD));
  IGM_LAUNCH_PDL(linattn_wt_kernel, dim3(B), dim3(256), (size_t)0, lc.stream, m, transpose, w_hi, w_lo, m2, cc);
  return IGM_OK;
}

int launch_linattn_p(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, const float* kstat,
                     __nv_bfloat16* p_hi, __nv_bfloat16* p_lo, int B, int n) {
  const int64_t M = (int64_t)B * n;
  int blocks = (int)cdiv64(M * (HD / 4), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps_(lc, K_ATTN, 4.0 * M * HD, 8.0 * M * HD);
  IGM_LAUNCH_PDL(linattn_p_kernel, dim3(blocks), dim3(256), (size_t)0, lc.stream, q_hi, q_lo, kstat, p_hi, p_lo, M, n);
  return IGM_OK;
}

int launch_linattn_dk(const LaunchCtx& lc, const __nv_bfloat16* p_hi, const __nv_bfloat16* p_lo, const float* T, const float* cc,
                      __nv_bfloat16* d_hi, __nv_bfloat16* d_lo, int B, int n) {
  const int64_t M = (int64_t)B * n;
  int blocks = (int)cdiv64(M * (HD / 4), 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps_(lc, K_ATTN, 2.0 * M * HD, 12.0 * M * HD);
  IGM_LAUNCH_PDL(linattn_dk_kernel, dim3(blocks), dim3(256), (size_t)0, lc.stream, p_hi, p_lo, T, cc, d_hi, d_lo, M, n);
  return IGM_OK;
}

}  // namespace igm
