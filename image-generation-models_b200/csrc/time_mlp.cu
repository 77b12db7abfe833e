// Timestep-embedding path of the reference U-Net, forward and backward:
//   SinusoidalPosEmb(dim) -> Linear(dim, 4dim) -> Mish -> Linear(4dim, dim)   (src/models/ddpm.py:47-59, :188-193)
//   per ResnetBlock: Mish -> Linear(dim, C_out)                                 (:126-130, applied at :139-140)
// All the per-block projections share the input mish(t_emb), so they are
// evaluated by ONE kernel over a device table of (weight, bias, C_out, offset).
// Negligible FLOPs (<0.02 % of a forward); plain CUDA-core kernels.
#include "common.cuh"

namespace igm {
namespace {

// one CTA per sample
__global__ void __launch_bounds__(256) time_mlp_fwd_kernel(const TimeMlpParams p, const int64_t* __restrict__ t,
                                                           float* __restrict__ emb, float* __restrict__ h1,
                                                           float* __restrict__ temb, float* __restrict__ act) {
  extern __shared__ float sm[];   // emb[d] | a1[4d]
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
    const float* w = p.w1 + (int64_t)j * d;
    float a = __ldg(p.b1 + j);
    for (int i = 0; i < d; ++i) a = fmaf(__ldg(w + i), s_emb[i], a);
    h1[(int64_t)b * d4 + j] = a;
    s_a1[j] = mish_f(a);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float* w = p.w2 + (int64_t)k * d4;
    float a = __ldg(p.b2 + k);
    for (int j = 0; j < d4; ++j) a = fmaf(__ldg(w + j), s_a1[j], a);
    temb[(int64_t)b * d + k] = a;
    act[(int64_t)b * d + k] = mish_f(a);
  }
}

// grid.x enumerates (projection, 32-channel slab) pairs via a prefix table walk
__global__ void __launch_bounds__(256) time_proj_fwd_kernel(const TimeProj* __restrict__ table, int n_proj,
                                                            const float* __restrict__ act, int dim, int B,
                                                            int total, float* __restrict__ proj) {
  extern __shared__ float ws[];   // [32][dim+1]
  int slab = blockIdx.x, j = 0;
  for (; j < n_proj; ++j) {
    const int ns = (table[j].cout + 31) / 32;
    if (slab < ns) break;
    slab -= ns;
  }
  if (j >= n_proj) return;
  const TimeProj tp = table[j];
  const int c0 = slab * 32;
  for (int i = threadIdx.x; i < 32 * dim; i += blockDim.x) {
    const int c = i / dim, k = i - c * dim;
    ws[c * (dim + 1) + k] = (c0 + c < tp.cout) ? __ldg(tp.w + (int64_t)(c0 + c) * dim + k) : 0.f;
  }
  __syncthreads();
  const int c = threadIdx.x & 31, bs = threadIdx.x >> 5;
  if (c0 + c >= tp.cout) return;
  const float bias = __ldg(tp.b + c0 + c);
  for (int b = bs; b < B; b += 8) {
    const float* a = act + (int64_t)b * dim;
    float s = bias;
#pragma unroll 8
    for (int k = 0; k < dim; ++k) s = fmaf(ws[c * (dim + 1) + k], __ldg(a + k), s);
    proj[(int64_t)b * total + tp.offset + c0 + c] = s;
  }
}

// (1) projection parameter gradients
__global__ void __launch_bounds__(256) time_proj_wgrad_kernel(const TimeProj* __restrict__ table, int n_proj,
                                                              const float* __restrict__ act, int dim, int B,
                                                              int total, const float* __restrict__ d_proj) {
  int slab = blockIdx.x, j = 0;
  for (; j < n_proj; ++j) {
    const int ns = (table[j].cout + 31) / 32;
    if (slab < ns) break;
    slab -= ns;
  }
  if (j >= n_proj) return;
  const TimeProj tp = table[j];
  const int c0 = slab * 32;
  for (int i = threadIdx.x; i < 32 * dim; i += blockDim.x) {
    const int c = c0 + i / dim, k = i % dim;
    if (c >= tp.cout) continue;
    float s = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b)
      s = fmaf(__ldg(d_proj + (int64_t)b * total + tp.offset + c), __ldg(act + (int64_t)b * dim + k), s);
    tp.gw[(int64_t)c * dim + k] += s;
  }
  if (threadIdx.x < 32 && c0 + threadIdx.x < tp.cout) {
    const int c = c0 + threadIdx.x;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __ldg(d_proj + (int64_t)b * total + tp.offset + c);
    tp.gb[c] += s;
  }
}

// (2) per-sample back-propagation to d_temb [B,d] and d_h1 [B,4d]; one CTA per sample
__global__ void __launch_bounds__(256) time_bwd_sample_kernel(const TimeMlpParams p,
                                                              const TimeProj* __restrict__ table, int n_proj,
                                                              int total, const float* __restrict__ h1,
                                                              const float* __restrict__ temb,
                                                              const float* __restrict__ d_proj,
                                                              float* __restrict__ d_temb, float* __restrict__ d_h1) {
  extern __shared__ float sm[];   // dproj[total] | red[256] | dtemb[d]
  const int d = p.dim, d4 = 4 * p.dim;
  float* s_dp = sm;
  float* s_red = sm + total;
  float* s_dt = s_red + 256;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < total; i += blockDim.x) s_dp[i] = __ldg(d_proj + (int64_t)b * total + i);
  __syncthreads();
  // d_act[k] = sum_{j,c} dproj[off_j + c] * W_j[c][k]
  const int parts = blockDim.x / d;          // requires d <= 256 and 256 % d == 0
  const int k = threadIdx.x % d, part = threadIdx.x / d;
  float s = 0.f;
  if (part < parts) {
    for (int j = 0; j < n_proj; ++j) {
      const TimeProj tp = table[j];
#pragma unroll 8
      for (int c = part; c < tp.cout; c += parts) s = fmaf(s_dp[tp.offset + c], __ldg(tp.w + (int64_t)c * d + k), s);
    }
  }
  s_red[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x < d) {
    float a = 0.f;
    for (int q = 0; q < parts; ++q) a += s_red[q * d + threadIdx.x];
    const float g = a * mish_grad_f(__ldg(temb + (int64_t)b * d + threadIdx.x));
    s_dt[threadIdx.x] = g;
    d_temb[(int64_t)b * d + threadIdx.x] = g;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d4; j += blockDim.x) {
    float a = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < d; ++kk) a = fmaf(s_dt[kk], __ldg(p.w2 + (int64_t)kk * d4 + j), a);
    d_h1[(int64_t)b * d4 + j] = a * mish_grad_f(__ldg(h1 + (int64_t)b * d4 + j));
  }
}

// (3) time_mlp parameter gradients; one thread per weight element, loop over the batch
__global__ void time_mlp_wgrad_kernel(const TimeMlpParams p, int B, const float* __restrict__ emb,
                                      const float* __restrict__ h1, const float* __restrict__ d_temb,
                                      const float* __restrict__ d_h1) {
  const int d = p.dim, d4 = 4 * p.dim;
  const int n2 = d * d4;          // w2 [d][4d]
  const int n1 = d4 * d;          // w1 [4d][d]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n2) {
    const int k = i / d4, j = i - k * d4;
    float s = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b)
      s = fmaf(__ldg(d_temb + (int64_t)b * d + k), mish_f(__ldg(h1 + (int64_t)b * d4 + j)), s);
    p.gw2[i] += s;
  } else if (i < n2 + n1) {
    const int r = i - n2;
    const int j = r / d, ii = r - j * d;
    float s = 0.f;
#pragma unroll 8
    for (int b = 0; b < B; ++b)
      s = fmaf(__ldg(d_h1 + (int64_t)b * d4 + j), __ldg(emb + (int64_t)b * d + ii), s);
    p.gw1[r] += s;
  } else if (i < n2 + n1 + d) {
    const int k = i - n2 - n1;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __ldg(d_temb + (int64_t)b * d + k);
    p.gb2[k] += s;
  } else if (i < n2 + n1 + d + d4) {
    const int j = i - n2 - n1 - d;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += __ldg(d_h1 + (int64_t)b * d4 + j);
    p.gb1[j] += s;
  }
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
  const size_t smem = (size_t)32 * (dim + 1) * sizeof(float);
  time_proj_fwd_kernel<<<slabs, 256, smem, lc.stream>>>(d_table, n_proj, act, dim, B, total, proj);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_time_backward(const LaunchCtx& lc, const TimeMlpParams& p, const TimeProj* d_table, int n_proj,
                         int total, int B, const float* emb, const float* h1, const float* temb,
                         const float* act, const float* d_proj, float* ws) {
  const int d = p.dim, d4 = 4 * p.dim;
  float* d_temb = ws;
  float* d_h1 = ws + (int64_t)B * d;
  const int slabs = total / 32;
  ProfScope ps_(lc, K_TIME, 6.0 * B * d * total, 0.0);
  time_proj_wgrad_kernel<<<slabs, 256, 0, lc.stream>>>(d_table, n_proj, act, d, B, total, d_proj);
  IGM_POST_LAUNCH(lc);
  const size_t smem = (size_t)(total + 256 + d) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(time_bwd_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
  }
  time_bwd_sample_kernel<<<B, 256, smem, lc.stream>>>(p, d_table, n_proj, total, h1, temb, d_proj, d_temb, d_h1);
  IGM_POST_LAUNCH(lc);
  const int n = 2 * d * d4 + d + d4;
  time_mlp_wgrad_kernel<<<cdiv(n, 256), 256, 0, lc.stream>>>(p, B, emb, h1, d_temb, d_h1);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
