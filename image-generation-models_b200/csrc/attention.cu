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

namespace igm {
namespace {

constexpr int D = kDimHead;          // 32
constexpr int QKV = 3 * kHeads * D;  // 384
constexpr int HD = kHeads * D;       // 128

// acc[i][j] += sum_{rows nn == grp (mod 4)} X[nn][d0+i] * Y[nn][e0+j]   (4x4 register tile)
__device__ __forceinline__ void outer4x4(const float (*Xs)[D], const float (*Ys)[D], int grp, int d0, int e0,
                                         float (&acc)[4][4], float (&xsum)[4]) {
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int nn = grp + r * 4;
    const float4 x = *reinterpret_cast<const float4*>(&Xs[nn][d0]);
    const float4 y = *reinterpret_cast<const float4*>(&Ys[nn][e0]);
    const float xv[4] = {x.x, x.y, x.z, x.w};
    const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xsum[i] += xv[i];
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], yv[j], acc[i][j]);
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

__global__ void __launch_bounds__(256) linattn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                          float* __restrict__ ctx, float* __restrict__ kstat,
                                                          int N, __nv_bfloat16* __restrict__ out_hi,
                                                          __nv_bfloat16* __restrict__ out_lo) {
  __shared__ __align__(16) float Xs[32][D];
  __shared__ __align__(16) float Ys[32][D];
  __shared__ float ctxs[D][D + 1];
  __shared__ float part[4][64][17];
  __shared__ float red[8][D];
  __shared__ float s_kmax[D], s_ksum[D];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int tid = threadIdx.x;
  const float* base = qkv + (int64_t)b * N * QKV;
  const int qcol = h * D, kcol = HD + h * D, vcol = 2 * HD + h * D;

  // 1) max over n of k[:, d]
  {
    const int d = tid & 31, r = tid >> 5;
    float m = -INFINITY;
    for (int n = r; n < N; n += 8) m = fmaxf(m, __ldg(base + (int64_t)n * QKV + kcol + d));
    red[r][d] = m;
    __syncthreads();
    if (tid < D) {
      float t = red[0][tid];
#pragma unroll
      for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i][tid]);
      s_kmax[tid] = t;
    }
    __syncthreads();
  }
  // 2) unnormalised context sum_n exp(k - max)[n][d] * v[n][e] and the softmax denominators
  const int grp = tid >> 6, t64 = tid & 63;
  const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
  {
    float acc[4][4], xsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xsum[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256;
        const int nn = idx >> 5, c = idx & 31;
        const int n = n0 + nn;
        float xv = 0.f, yv = 0.f;
        if (n < N) {
          xv = expf(__ldg(base + (int64_t)n * QKV + kcol + c) - s_kmax[c]);
          yv = __ldg(base + (int64_t)n * QKV + vcol + c);
        }
        Xs[nn][c] = xv;
        Ys[nn][c] = yv;
      }
      __syncthreads();
      outer4x4(Xs, Ys, grp, d0, e0, acc, xsum);
      __syncthreads();
    }
    reduce_groups(acc, xsum, grp, t64, d0, e0, part, ctxs, s_ksum);
  }
  // normalise, publish ctx / statistics
  for (int i = tid; i < D * D; i += 256) {
    const int d = i >> 5, e = i & 31;
    const float v = ctxs[d][e] / s_ksum[d];
    ctxs[d][e] = v;
    ctx[((int64_t)b * kHeads + h) * D * D + i] = v;
  }
  if (tid < D) {
    float* ksd = kstat + (((int64_t)b * kHeads + h) * D + tid) * 2;
    ksd[0] = s_kmax[tid];
    ksd[1] = s_ksum[tid];
  }
  __syncthreads();
  // 3) out[n][e] = sum_d ctx[d][e] * q[n][d]; the context column of this thread sits in registers
  const int e = tid & 31, r = tid >> 5;
  float col[D];
#pragma unroll
  for (int dd = 0; dd < D; ++dd) col[dd] = ctxs[dd][e];
  for (int n0 = 0; n0 < N; n0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int nn = idx >> 5, c = idx & 31;
      const int n = n0 + nn;
      Xs[nn][c] = (n < N) ? __ldg(base + (int64_t)n * QKV + qcol + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nn = r + i * 8;
      const int n = n0 + nn;
      float a = 0.f;
#pragma unroll
      for (int d4 = 0; d4 < D; d4 += 4) {
        const float4 x = *reinterpret_cast<const float4*>(&Xs[nn][d4]);
        a = fmaf(col[d4], x.x, a);
        a = fmaf(col[d4 + 1], x.y, a);
        a = fmaf(col[d4 + 2], x.z, a);
        a = fmaf(col[d4 + 3], x.w, a);
      }
      if (n < N) {
        const int64_t o = ((int64_t)b * N + n) * HD + h * D + e;
        out[o] = a;
        if (out_hi) {
          const __nv_bfloat16 hv = __float2bfloat16_rn(a);
          out_hi[o] = hv;
          out_lo[o] = __float2bfloat16_rn(a - __bfloat162float(hv));
        }
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) linattn_bwd_kernel(const float* __restrict__ qkv,
                                                          const float* __restrict__ ctx,
                                                          const float* __restrict__ kstat,
                                                          const float* __restrict__ d_out,
                                                          float* __restrict__ d_qkv, int N) {
  __shared__ __align__(16) float Xs[32][D];   // q (pass A) / softmax(k) (pass B)
  __shared__ __align__(16) float Ys[32][D];   // dO
  __shared__ __align__(16) float Vs[32][D];   // v
  __shared__ float ctxs[D][D + 1];
  __shared__ float dctxs[D][D + 1];
  __shared__ float part[4][64][17];
  __shared__ float s_kmax[D], s_kinv[D], s_cdot[D];
  const int b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
  const int tid = threadIdx.x;
  const float* base = qkv + (int64_t)b * N * QKV;
  const float* dob = d_out + (int64_t)b * N * HD + h * D;
  float* dqb = d_qkv + (int64_t)b * N * QKV;
  const int qcol = h * D, kcol = HD + h * D, vcol = 2 * HD + h * D;

  for (int i = tid; i < D * D; i += 256) ctxs[i >> 5][i & 31] = __ldg(ctx + ((int64_t)b * kHeads + h) * D * D + i);
  if (tid < D) {
    const float* ksd = kstat + (((int64_t)b * kHeads + h) * D + tid) * 2;
    s_kmax[tid] = ksd[0];
    s_kinv[tid] = 1.f / ksd[1];
  }
  __syncthreads();

  // A) dctx[d][e] = sum_n q[n][d] * dO[n][e]
  {
    const int grp = tid >> 6, t64 = tid & 63;
    const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
    float acc[4][4], xsum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      xsum[i] = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    for (int n0 = 0; n0 < N; n0 += 32) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int idx = tid + i * 256;
        const int nn = idx >> 5, c = idx & 31;
        const int n = n0 + nn;
        float xv = 0.f, yv = 0.f;
        if (n < N) {
          xv = __ldg(base + (int64_t)n * QKV + qcol + c);
          yv = __ldg(dob + (int64_t)n * HD + c);
        }
        Xs[nn][c] = xv;
        Ys[nn][c] = yv;
      }
      __syncthreads();
      outer4x4(Xs, Ys, grp, d0, e0, acc, xsum);
      __syncthreads();
    }
    reduce_groups(acc, xsum, grp, t64, d0, e0, part, dctxs, nullptr);
  }
  if (tid < D) {
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < D; ++e) t = fmaf(dctxs[tid][e], ctxs[tid][e], t);
    s_cdot[tid] = t;
  }
  __syncthreads();

  // B) per-row gradients; thread (c, r) keeps row c of ctx and dctx and column c of dctx in registers
  const int c = tid & 31, r = tid >> 5;
  float crow[D], drow[D], dcol[D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    crow[j] = ctxs[c][j];
    drow[j] = dctxs[c][j];
    dcol[j] = dctxs[j][c];
  }
  const float cdot = s_cdot[c];
  for (int n0 = 0; n0 < N; n0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int nn = idx >> 5, cc = idx & 31;
      const int n = n0 + nn;
      float p = 0.f, v = 0.f, g = 0.f;
      if (n < N) {
        p = expf(__ldg(base + (int64_t)n * QKV + kcol + cc) - s_kmax[cc]) * s_kinv[cc];
        v = __ldg(base + (int64_t)n * QKV + vcol + cc);
        g = __ldg(dob + (int64_t)n * HD + cc);
      }
      Xs[nn][cc] = p;
      Vs[nn][cc] = v;
      Ys[nn][cc] = g;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int nn = r + i * 8;
      const int n = n0 + nn;
      float dq = 0.f, dp = 0.f, dv = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < D; j4 += 4) {
        const float4 g4 = *reinterpret_cast<const float4*>(&Ys[nn][j4]);
        const float4 v4 = *reinterpret_cast<const float4*>(&Vs[nn][j4]);
        const float4 p4 = *reinterpret_cast<const float4*>(&Xs[nn][j4]);
        dq = fmaf(crow[j4], g4.x, dq); dq = fmaf(crow[j4 + 1], g4.y, dq);
        dq = fmaf(crow[j4 + 2], g4.z, dq); dq = fmaf(crow[j4 + 3], g4.w, dq);
        dp = fmaf(drow[j4], v4.x, dp); dp = fmaf(drow[j4 + 1], v4.y, dp);
        dp = fmaf(drow[j4 + 2], v4.z, dp); dp = fmaf(drow[j4 + 3], v4.w, dp);
        dv = fmaf(dcol[j4], p4.x, dv); dv = fmaf(dcol[j4 + 1], p4.y, dv);
        dv = fmaf(dcol[j4 + 2], p4.z, dv); dv = fmaf(dcol[j4 + 3], p4.w, dv);
      }
      if (n < N) {
        float* o = dqb + (int64_t)n * QKV;
        o[qcol + c] = dq;
        o[kcol + c] = Xs[nn][c] * (dp - cdot);
        o[vcol + c] = dv;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_linattn_forward(const LaunchCtx& lc, const float* qkv, float* out, float* ctx, float* kstat,
                           int B, int n, __nv_bfloat16* out_hi, __nv_bfloat16* out_lo) {
  ProfScope ps_(lc, K_ATTN, 4.0 * B * kHeads * (double)n * D * D, 4.0 * B * (double)n * (QKV + HD));
  linattn_fwd_kernel<<<B * kHeads, 256, 0, lc.stream>>>(qkv, out, ctx, kstat, n, out_hi, out_lo);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_linattn_backward(const LaunchCtx& lc, const float* qkv, const float* ctx, const float* kstat,
                            const float* d_out, float* d_qkv, int B, int n) {
  ProfScope ps_(lc, K_ATTN, 8.0 * B * kHeads * (double)n * D * D, 4.0 * B * (double)n * (2 * QKV + HD));
  linattn_bwd_kernel<<<B * kHeads, 256, 0, lc.stream>>>(qkv, ctx, kstat, d_out, d_qkv, n);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
