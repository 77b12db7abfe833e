// Timestep-embedding path of the reference U-Net, forward and backward:
//   SinusoidalPosEmb(dim) -> Linear(dim, 4dim) -> Mish -> Linear(4dim, dim)   (src/models/ddpm.py:47-59, :188-193)
//   per ResnetBlock: Mish -> Linear(dim, C_out)                                 (:126-130, applied at :139-140)
// All the per-block projections share the input mish(t_emb), so they are
// evaluated by ONE kernel over a device table of (weight, bias, C_out, offset).
// Negligible FLOPs (<0.02 % of a forward); plain CUDA-core kernels.
#include "common.cuh"

namespace igm {
namespace {

// one CTA per sample; every dot product reads its weight row with independent 128-bit loads
__global__ void __launch_bounds__(256) time_mlp_fwd_kernel(const TimeMlpParams p, const int64_t* __restrict__ t,
                                                           float* __restrict__ emb, float* __restrict__ h1,
                                                           float* __restrict__ temb, float* __restrict__ act) {
  extern __shared__ __align__(16) float sm[];   // emb[d] | a1[4d]
  const int d = p.dim, d4 = 4 * p.dim;
  float* s_emb = sm;
  float* s_a1 = sm + d;
  const int b = blockIdx.x;
  const float tf = (float)t[b];
  const int half = d / 2;
  // emb = log(10000)/(half-1); freq_j = exp(j * -emb) in fp32, like torch (ddpm.py:55-57)
  const float neg = (float)(-(9.210340371976184 /* ln 1e4 */ / (double)(half - 1)));
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const int j = i < half ? i : i - half;
    const float arg = tf * expf((float)j * neg);
    const float v = i < half ? sinf(arg) : cosf(arg);
    s_emb[i] = v;
    emb[(int64_t)b * d + i] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d4; j += blockDim.x) {
    const float4* w = reinterpret_cast<const float4*>(p.w1 + (int64_t)j * d);
    float a0 = __ldg(p.b1 + j), a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int i = 0; i < d / 4; ++i) {
      const float4 wv = __ldg(w + i);
      const float4 e = *reinterpret_cast<const float4*>(s_emb + 4 * i);
      a0 = fmaf(wv.x, e.x, a0); a1 = fmaf(wv.y, e.y, a1); a2 = fmaf(wv.z, e.z, a2); a3 = fmaf(wv.w, e.w, a3);
    }
    const float a = (a0 + a1) + (a2 + a3);
    h1[(int64_t)b * d4 + j] = a;
    s_a1[j] = mish_f(a);
  }
  __syncthreads();
  // layer 2: four threads per output, each over a quarter of the 4d inputs
  for (int o = threadIdx.x; o < 4 * d; o += blockDim.x) {
    const int k = o >> 2, part = o & 3;
    const float4* w = reinterpret_cast<const float4*>(p.w2 + (int64_t)k * d4 + part * d);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
    for (int i = 0; i < d / 4; ++i) {
      const float4 wv = __ldg(w + i);
      const float4 e = *reinterpret_cast<const float4*>(s_a1 + part * d + 4 * i);
      a0 = fmaf(wv.x, e.x, a0); a1 = fmaf(wv.y, e.y, a1); a2 = fmaf(wv.z, e.z, a2); a3 = fmaf(wv.w, e.w, a3);
    }
    float a = (a0 + a1) + (a2 + a3);
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    if (part == 0) {
      a += __ldg(p.b2 + k);
      temb[(int64_t)b * d + k] = a;
      act[(int64_t)b * d + k] = mish_f(a);
    }
  }
}

constexpr int kProjSamples = 32;   // samples per CTA of the projection kernels

// walk the projection table: slab (32 output channels) -> (projection j, first channel c0)
__device__ __forceinline__ bool find_slab(const TimeProj* __restrict__ table, int n_proj, int slab, TimeProj& tp, int& c0) {
  int j = 0;
  for (; j < n_proj; ++j) {
    const int ns = (table[j].cout + 31) / 32;
    if (slab < ns) break;
    slab -= ns;
  }
  if (j >= n_proj) return false;
  tp = table[j];
  c0 = slab * 32;
  return true;
}

// grid (slabs, ceil(B / 32)): proj[b][off + c] = bias[c] + sum_k W[c][k] * act[b][k]
__global__ void __launch_bounds__(256) time_proj_fwd_kernel(const TimeProj* __restrict__ table, int n_proj,
                                                            const float* __restrict__ act, int dim, int B,
                                                            int total, float* __restrict__ proj) {
  extern __shared__ __align__(16) float ws[];   // W[32][dim+1] | act[32][dim]
  float* s_act = ws + 32 * (dim + 1);
  TimeProj tp;
  int c0;
  if (!find_slab(table, n_proj, blockIdx.x, tp, c0)) return;
  const int b0 = blockIdx.y * kProjSamples;
  for (int i = threadIdx.x; i < 32 * dim; i += blockDim.x) {
    const int c = i / dim, k = i - c * dim;
    ws[c * (dim + 1) + k] = (c0 + c < tp.cout) ? __ldg(tp.w + (int64_t)(c0 + c) * dim + k) : 0.f;
    s_act[i] = (b0 + c < B) ? __ldg(act + (int64_t)(b0 + c) * dim + k) : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x & 31, bs = threadIdx.x >> 5;
  if (c0 + c >= tp.cout) return;
  const float bias = __ldg(tp.b + c0 + c);
  float s[4] = {bias, bias, bias, bias};
  for (int k = 0; k < dim; ++k) {
    const float w = ws[c * (dim + 1) + k];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = fmaf(w, s_act[(bs + 8 * q) * dim + k], s[q]);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int b = b0 + bs + 8 * q;
    if (b < B) proj[(int64_t)b * total + tp.offset + c0 + c] = s[q];
  }
}

// (1) grid (slabs, ceil(B / 32)).  For its 32 channels x 32 samples the CTA produces
//     * the projection weight / bias gradients   gW[c][k] += sum_b dproj[b][c] * act[b][k]
//     * its share of d_act[b][k] += sum_c dproj[b][c] * W[c][k]  (fp32 atomics into a zeroed accumulator)
__global__ void __launch_bounds__(256) time_proj_bwd_kernel(const TimeProj* __restrict__ table, int n_proj,
                                                            const float* __restrict__ act, int dim, int B, int total,
                                                            const float* __restrict__ d_proj, float* __restrict__ d_act,
                                                            int slab0) {
  extern __shared__ __align__(16) float ws[];   // W[32][dim+1] | act[32][dim+1] | dp[32 samples][33]
  float* s_act = ws + 32 * (dim + 1);
  float* s_dp = s_act + 32 * (dim + 1);
  TimeProj tp;
  int c0;
  if (!find_slab(table, n_proj, blockIdx.x + slab0, tp, c0)) return;
  const int b0 = blockIdx.y * kProjSamples;
  for (int i = threadIdx.x; i < 32 * dim; i += blockDim.x) {
    const int c = i / dim, k = i - c * dim;
    ws[c * (dim + 1) + k] = (c0 + c < tp.cout) ? __ldg(tp.w + (int64_t)(c0 + c) * dim + k) : 0.f;
    s_act[c * (dim + 1) + k] = (b0 + c < B) ? __ldg(act + (int64_t)(b0 + c) * dim + k) : 0.f;
  }
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int bb = i >> 5, c = i & 31;
    s_dp[bb * 33 + c] = (b0 + bb < B && c0 + c < tp.cout) ? __ldg(d_proj + (int64_t)(b0 + bb) * total + tp.offset + c0 + c) : 0.f;
  }
  __syncthreads();
  // weight gradient: 32 x dim outputs
  for (int o = threadIdx.x; o < 32 * dim; o += blockDim.x) {
    const int c = o / dim, k = o - c * dim;
    float s = 0.f;
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) s = fmaf(s_dp[bb * 33 + c], s_act[bb * (dim + 1) + k], s);
    if (c0 + c < tp.cout) atomicAdd(tp.gw + (int64_t)(c0 + c) * dim + k, s);
  }
  if (threadIdx.x < 32 && c0 + threadIdx.x < tp.cout) {
    float s = 0.f;
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) s += s_dp[bb * 33 + threadIdx.x];
    atomicAdd(tp.gb + c0 + threadIdx.x, s);
  }
  // data gradient: 32 samples x dim outputs
  for (int o = threadIdx.x; o < 32 * dim; o += blockDim.x) {
    const int bb = o / dim, k = o - bb * dim;
    float s = 0.f;
#pragma unroll 8
    for (int c = 0; c < 32; ++c) s = fmaf(s_dp[bb * 33 + c], ws[c * (dim + 1) + k], s);
    if (b0 + bb < B) atomicAdd(d_act + (int64_t)(b0 + bb) * dim + k, s);
  }
}

// (2) one CTA per sample: d_temb = d_act * mish'(temb) (and re-zero the accumulator); a1 = mish(h1);
//     d_h1 = (d_temb . W2) * mish'(h1)
__global__ void __launch_bounds__(256) time_bwd_sample_kernel(const TimeMlpParams p, const float* __restrict__ h1,
                                                              const float* __restrict__ temb, float* __restrict__ d_act,
                                                              float* __restrict__ d_temb, float* __restrict__ d_h1,
                                                              float* __restrict__ a1) {
  extern __shared__ __align__(16) float sm[];   // dtemb[d]
  const int d = p.dim, d4 = 4 * p.dim;
  const int b = blockIdx.x;
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float g = d_act[(int64_t)b * d + k] * mish_grad_f(__ldg(temb + (int64_t)b * d + k));
    d_act[(int64_t)b * d + k] = 0.f;
    sm[k] = g;
    d_temb[(int64_t)b * d + k] = g;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d4; j += blockDim.x) {
    float a = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < d; ++kk) a = fmaf(sm[kk], __ldg(p.w2 + (int64_t)kk * d4 + j), a);
    const float h = __ldg(h1 + (int64_t)b * d4 + j);
    d_h1[(int64_t)b * d4 + j] = a * mish_grad_f(h);
    a1[(int64_t)b * d4 + j] = mish_f(h);
  }
}

// (3) time_mlp parameter gradients as batch-reduced outer products, 32 x 32 output tiles:
//     C[m][n] += sum_b A[b][m] * Bm[b][n]   (tiles 0..n2-1: gw2 = d_temb^T a1; then gw1 = d_h1^T emb); biases = column sums of A
__global__ void __launch_bounds__(256) time_mlp_wgrad_kernel(const TimeMlpParams p, int B, const float* __restrict__ emb,
                                                             const float* __restrict__ a1, const float* __restrict__ d_temb,
                                                             const float* __restrict__ d_h1) {
  __shared__ float As[32][33], Bs[32][33];
  const int d = p.dim, d4 = 4 * p.dim;
  const int tiles2 = (d / 32) * (d4 / 32);
  int tile = blockIdx.x;
  const float* A; const float* Bm; float* C; float* bias;
  int lda, ldb, ldc, tn;
  if (tile < tiles2) { A = d_temb; lda = d; Bm = a1; ldb = d4; C = p.gw2; ldc = d4; bias = p.gb2; tn = d4 / 32; }
  else { tile -= tiles2; A = d_h1; lda = d4; Bm = emb; ldb = d; C = p.gw1; ldc = d; bias = p.gb1; tn = d / 32; }
  const int m0 = (tile / tn) * 32, n0 = (tile % tn) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // thread -> column n0+tx, rows ty, ty+8, ty+16, ty+24
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bsum = 0.f;
  for (int b0 = 0; b0 < B; b0 += 32) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int bb = ty + 8 * q;
      const bool ok = b0 + bb < B;
      As[bb][tx] = ok ? __ldg(A + (int64_t)(b0 + bb) * lda + m0 + tx) : 0.f;
      Bs[bb][tx] = ok ? __ldg(Bm + (int64_t)(b0 + bb) * ldb + n0 + tx) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int bb = 0; bb < 32; ++bb) {
      const float bv = Bs[bb][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fmaf(As[bb][ty + 8 * q], bv, acc[q]);
    }
    if (n0 == 0 && ty == 0) {
#pragma unroll 8
      for (int bb = 0; bb < 32; ++bb) bsum += As[bb][tx];
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) C[(int64_t)(m0 + ty + 8 * q) * ldc + n0 + tx] += acc[q];
  if (n0 == 0 && ty == 0) bias[m0 + tx] += bsum;
}

}  // namespace

int launch_time_mlp_forward(const LaunchCtx& lc, const TimeMlpParams& p, const int64_t* t, int B, float* emb,
                            float* h1, float* temb, float* act) {
  if (p.dim > 256 || 256 % p.dim != 0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "time_mlp: dim must divide 256");
  ProfScope ps_(lc, K_TIME, 16.0 * B * p.dim * p.dim, 0.0);
  const size_t smem = (size_t)5 * p.dim * sizeof(float);
  time_mlp_fwd_kernel<<<B, 256, smem, lc.stream>>>(p, t, emb, h1, temb, act);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_time_proj_forward(const LaunchCtx& lc, const TimeProj* d_table, int n_proj, const float* act,
                             int dim, int B, int total, float* proj) {
  // every projection width is a multiple of 32, so slabs = total / 32
  const int slabs = total / 32;
  ProfScope ps_(lc, K_TIME, 2.0 * B * dim * total, 0.0);
  const size_t smem = (size_t)(32 * (dim + 1) + 32 * dim) * sizeof(float);
  time_proj_fwd_kernel<<<dim3(slabs, cdiv(B, kProjSamples)), 256, smem, lc.stream>>>(d_table, n_proj, act, dim, B, total, proj);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// Backward of the per-block time projections for the 32-channel slabs [slab_lo, slab_hi) of the projection table (all of
// them when slab_hi < 0): parameter gradients of those blocks + their share of d_act (accumulated; the accumulator is
// zero on entry of a backward pass and re-zeroed by launch_time_mlp_backward).  Split by slab range so that the
// gradients of a block group are final as soon as the backward pass has left that group (gradient buckets, unet.cu).
// Workspace layout ([Bcap, 10 d] floats, Bcap = the batch the context was planned for, NOT the batch of this call: the
// accumulator must sit at the same address whatever batch runs, or a smaller batch's d_temb / d_h1 would land on
// accumulator rows a later, larger batch expects to find zero -- the epoch-tail batch of a loader without drop_last):
//   d_act [Bcap, d] (accumulator, zero between backward passes) | d_temb [Bcap, d] | d_h1 [Bcap, 4d] | mish(h1) [Bcap, 4d]
int launch_time_proj_backward(const LaunchCtx& lc, const TimeMlpParams& p, const TimeProj* d_table, int n_proj, int total,
                              int B, int Bcap, const float* act, const float* d_proj, float* ws, int slab_lo, int slab_hi) {
  const int d = p.dim;
  if (d % 32 != 0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "time_mlp: dim must be a multiple of 32");
  if (B > Bcap) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "time_mlp: batch exceeds the planned batch");
  float* d_act = ws;
  if (slab_hi < 0) { slab_lo = 0; slab_hi = total / 32; }
  if (slab_hi <= slab_lo) return IGM_OK;
  ProfScope ps_(lc, K_TIME, 6.0 * B * d * 32.0 * (slab_hi - slab_lo), 0.0);
  const size_t smem = (size_t)(2 * 32 * (d + 1) + 32 * 33) * sizeof(float);
  time_proj_bwd_kernel<<<dim3(slab_hi - slab_lo, cdiv(B, kProjSamples)), 256, smem, lc.stream>>>(d_table, n_proj, act, d, B, total,
                                                                                                 d_proj, d_act, slab_lo);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// The time MLP itself (Linear -> Mish -> Linear, ddpm.py:188-193), after EVERY projection slab has added its d_act share.
int launch_time_mlp_backward(const LaunchCtx& lc, const TimeMlpParams& p, int B, int Bcap, const float* emb, const float* h1,
                             const float* temb, float* ws) {
  const int d = p.dim, d4 = 4 * p.dim;
  float* d_act = ws;                           // accumulator: zero on entry, re-zeroed by time_bwd_sample_kernel
  float* d_temb = d_act + (int64_t)Bcap * d;
  float* d_h1 = d_temb + (int64_t)Bcap * d;
  float* a1 = d_h1 + (int64_t)Bcap * d4;
  ProfScope ps_(lc, K_TIME, 4.0 * B * d * d4 * 2, 0.0);
  time_bwd_sample_kernel<<<B, 256, d * sizeof(float), lc.stream>>>(p, h1, temb, d_act, d_temb, d_h1, a1);
  IGM_POST_LAUNCH(lc);
  const int tiles = 2 * (d / 32) * (d4 / 32);
  time_mlp_wgrad_kernel<<<tiles, 256, 0, lc.stream>>>(p, B, emb, a1, d_temb, d_h1);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
