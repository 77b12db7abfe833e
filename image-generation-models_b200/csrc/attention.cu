// Linear attention of the reference U-Net (src/models/ddpm.py:154-166), forward
// and backward, on the NHWC qkv tensor produced by the to_qkv 1x1 convolution.
//
//   q,k,v : [B, heads=4, d=32, n]     (channel = which*128 + head*32 + d)
//   k     <- softmax over n            (ddpm.py:160)
//   ctx   =  einsum('bhdn,bhen->bhde', k, v)        (:161)
//   out   =  einsum('bhde,bhdn->bhen', ctx, q)      (:162)
//
// One CTA per (batch, head): the 32x32 context lives in shared memory and the
// spatial axis is streamed through in 32-row chunks.  <2 % of the U-Net FLOPs,
// so CUDA-core FMAs (fp32-exact) are used rather than tensor cores.
#include "common.cuh"

namespace igm {
namespace {

constexpr int D = kDimHead;          // 32
constexpr int QKV = 3 * kHeads * D;  // 384
constexpr int HD = kHeads * D;       // 128

// acc[d][e0..e0+3] += sum_{rows} X[row][d] * Y[row][e]; thread owns d = tid>>3, e0 = (tid&7)*4.
// X is transformed by f(x, d) on load.
template <typename F>
__device__ __forceinline__ void outer_accumulate(const float* __restrict__ base, int N, int xcol, int ycol,
                                                 F f, float (*Xs)[D + 1], float (*Ys)[D], float acc[4],
                                                 float* xsum) {
  const int tid = threadIdx.x;
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  for (int n0 = 0; n0 < N; n0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int nn = idx >> 5, c = idx & 31;
      const int n = n0 + nn;
      float xv = 0.f, yv = 0.f;
      if (n < N) {
        xv = f(__ldg(base + (int64_t)n * QKV + xcol + c), c);
        yv = __ldg(base + (int64_t)n * QKV + ycol + c);
      }
      Xs[nn][c] = xv;
      Ys[nn][c] = yv;
    }
    __syncthreads();
#pragma unroll 8
    for (int nn = 0; nn < 32; ++nn) {
      const float x = Xs[nn][d];
      const float4 y = *reinterpret_cast<const float4*>(&Ys[nn][e0]);
      acc[0] = fmaf(x, y.x, acc[0]);
      acc[1] = fmaf(x, y.y, acc[1]);
      acc[2] = fmaf(x, y.z, acc[2]);
      acc[3] = fmaf(x, y.w, acc[3]);
      if (xsum) *xsum += x;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) linattn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                          float* __restrict__ ctx, float* __restrict__ kstat,
                                                          int N, __nv_bfloat16* __restrict__ out_hi,
                                                          __nv_bfloat16* __restrict__ out_lo) {
  __shared__ float Xs[32][D + 1];
  __shared__ __align__(16) float Ys[32][D];
  __shared__ float ctxs[D][D + 1];
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
  // 2) unnormalised context and softmax denominator
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float ks = 0.f;
  outer_accumulate(
      base, N, kcol, vcol, [&](float x, int c) { return expf(x - s_kmax[c]); }, Xs, Ys, acc, &ks);
  const int d = tid >> 3, e0 = (tid & 7) * 4;
  if ((tid & 7) == 0) s_ksum[d] = ks;
  __syncthreads();
  {
    const float inv = 1.f / s_ksum[d];
    float* cg = ctx + (((int64_t)b * kHeads + h) * D + d) * D + e0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = acc[j] * inv;
      ctxs[d][e0 + j] = v;
      cg[j] = v;
    }
    if (tid < D) {
      float* ksd = kstat + (((int64_t)b * kHeads + h) * D + tid) * 2;
      ksd[0] = s_kmax[tid];
      ksd[1] = s_ksum[tid];
    }
  }
  __syncthreads();
  // 3) out[n][e] = sum_d ctx[d][e] * q[n][d]
  const int e = tid & 31, r = tid >> 5;
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
      for (int dd = 0; dd < D; ++dd) a = fmaf(ctxs[dd][e], Xs[nn][dd], a);
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
  __shared__ float Xs[32][D + 1];
  __shared__ __align__(16) float Ys[32][D];
  __shared__ float Vs[32][D + 1];
  __shared__ float ctxs[D][D + 1];
  __shared__ float dctxs[D][D + 1];
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
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  {
    const int d = tid >> 3, e0 = (tid & 7) * 4;
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
#pragma unroll 8
      for (int nn = 0; nn < 32; ++nn) {
        const float x = Xs[nn][d];
        const float4 y = *reinterpret_cast<const float4*>(&Ys[nn][e0]);
        acc[0] = fmaf(x, y.x, acc[0]);
        acc[1] = fmaf(x, y.y, acc[1]);
        acc[2] = fmaf(x, y.z, acc[2]);
        acc[3] = fmaf(x, y.w, acc[3]);
      }
      __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) dctxs[d][e0 + j] = acc[j];
  }
  __syncthreads();
  if (tid < D) {
    float t = 0.f;
#pragma unroll
    for (int e = 0; e < D; ++e) t = fmaf(dctxs[tid][e], ctxs[tid][e], t);
    s_cdot[tid] = t;
  }
  __syncthreads();

  // B) per-row gradients
  const int c = tid & 31, r = tid >> 5;
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
      for (int j = 0; j < D; ++j) {
        dq = fmaf(ctxs[c][j], Ys[nn][j], dq);
        dp = fmaf(dctxs[c][j], Vs[nn][j], dp);
        dv = fmaf(dctxs[j][c], Xs[nn][j], dv);
      }
      if (n < N) {
        float* o = dqb + (int64_t)n * QKV;
        o[qcol + c] = dq;
        o[kcol + c] = Xs[nn][c] * (dp - s_cdot[c]);
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
