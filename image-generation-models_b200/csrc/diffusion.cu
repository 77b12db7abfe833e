// Boundary conversions and the streaming (HBM-bound) pieces of GaussianDiffusion:
//   q_sample            reference src/models/ddpm.py:433-444  (fused with NCHW->NHWC)
//   final Conv1x1       :236
//   l1 / l2 loss        :453-456 (+ its gradient seed)
//   p_sample tail       :359-364, :385, :367-376, :394-397 (one fused update kernel)
//   Adam                torch.optim.Adam as configured at :502-512
#include "common.cuh"

namespace igm {
namespace {

// ---------------------------------------------------------------------------
__global__ void input_prep_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                  const int64_t* __restrict__ t, const float* __restrict__ sa,
                                  const float* __restrict__ sb, float* __restrict__ out_nhwc,
                                  float* __restrict__ out_nchw, int B, int C, int HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b, p)
  if (i >= (int64_t)B * HW) return;
  const int b = (int)(i / HW), p = (int)(i - (int64_t)b * HW);
  float ca = 1.f, cb = 0.f;
  if (noise) {
    const int64_t tb = t[b];
    ca = __ldg(sa + tb);
    cb = __ldg(sb + tb);
  }
  for (int c = 0; c < C; ++c) {
    const int64_t src = ((int64_t)b * C + c) * HW + p;
    float v = __ldg(x + src);
    // extract(a,t)*x_start + extract(b,t)*noise with torch's op order (mul, mul, add; no fma)
    if (noise) v = __fadd_rn(__fmul_rn(ca, v), __fmul_rn(cb, __ldg(noise + src)));
    if (out_nhwc) out_nhwc[i * C + c] = v;
    if (out_nchw) out_nchw[src] = v;
  }
}

// 8 lanes per pixel; w in shared memory; writes NCHW
__global__ void __launch_bounds__(256) final_conv_kernel(const float* __restrict__ act, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out,
                                                         int B, int HW, int K, int C) {
  extern __shared__ float sw[];   // [C][K]
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = __ldg(w + i);
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int64_t M = (int64_t)B * HW;
  const bool ok = m < M;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (ok) {
    const float* a = act + m * K;
    for (int k4 = sub; k4 < (K >> 2); k4 += 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(a + k4 * 4));
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < C) {
          const float* wc = sw + c * K + k4 * 4;
          acc[c] += v.x * wc[0] + v.y * wc[1] + v.z * wc[2] + v.w * wc[3];
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 4);
    acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
    acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
  }
  if (ok && sub == 0) {
    const int b = (int)(m / HW), p = (int)(m - (int64_t)b * HW);
    for (int c = 0; c < C; ++c) out[((int64_t)b * C + c) * HW + p] = acc[c] + __ldg(bias + c);
  }
}

// ---------------------------------------------------------------------------
// loss: stage 1 partial sums (fixed grid), stage 2 fp64 combine
constexpr int LOSS_CTAS = 296;

__global__ void __launch_bounds__(256) loss_partial_kernel(const float* __restrict__ pred,
                                                           const float* __restrict__ noise, int64_t n,
                                                           int loss_type, float* __restrict__ ws) {
  __shared__ float red[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __ldg(noise + i) - __ldg(pred + i);
    s += (loss_type == 1) ? fabsf(d) : d * d;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    ws[blockIdx.x] = t;
  }
}

__global__ void loss_final_kernel(const float* __restrict__ ws, int parts, int64_t n, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < parts; ++i) t += (double)ws[i];
    *out = (float)(t / (double)n);
  }
}

// d_pred in NHWC: i enumerates NCHW elements of pred
__global__ void loss_backward_kernel(const float* __restrict__ pred, const float* __restrict__ noise,
                                     int64_t n, int C, int HW, int loss_type,
                                     const float* __restrict__ d_loss, float scale,
                                     float* __restrict__ d_nhwc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (d_loss) scale *= __ldg(d_loss);
  const float d = __ldg(pred + i) - __ldg(noise + i);   // d(loss)/d(pred) has the sign of (pred - noise)
  float g;
  if (loss_type == 1)
    g = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
  else
    g = 2.f * d;
  const int64_t bc = i / HW;
  const int p = (int)(i - bc * HW);
  const int64_t b = bc / C;
  const int c = (int)(bc - b * C);
  d_nhwc[(b * HW + p) * C + c] = g * scale;
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index
  if (i >= (int64_t)B * HW * C) return;
  const int64_t bc = i / HW;
  const int p = (int)(i - bc * HW);
  const int64_t b = bc / C;
  const int c = (int)(bc - b * C);
  dst[(b * HW + p) * C + c] = __ldg(src + i);
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int HW, int C) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW index
  if (i >= (int64_t)B * HW * C) return;
  const int64_t bc = i / HW;
  const int p = (int)(i - bc * HW);
  const int64_t b = bc / C;
  const int c = (int)(bc - b * C);
  dst[i] = __ldg(src + (b * HW + p) * C + c);
}

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011) + Box-Muller: 4 normals per counter
__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
__device__ __forceinline__ void philox4x32_10(uint64_t seed, uint64_t ctr_lo, uint32_t ctr_hi, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)ctr_lo, c1 = (uint32_t)(ctr_lo >> 32), c2 = ctr_hi, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)a + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)b * 2.3283064365386963e-10f;            // [0, 1)
  const float r = sqrtf(-2.f * logf(u1));
  float s, c;
  sincosf(6.283185307179586f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

// one thread per 4 consecutive NCHW elements
__global__ void __launch_bounds__(256) sampler_update_kernel(const SamplerStepArgs a) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i0 = i4 * 4;
  if (i0 >= a.n) return;
  const int64_t t = a.t_dev[0];
  const int step = a.step_dev[0] - 1;     // index of this step since the loop started
  const igm_schedule& S = *a.sched_dev;
  const float c1 = __ldg(S.sqrt_recip_alphas_cumprod + t);
  const float c2 = __ldg(S.sqrt_recipm1_alphas_cumprod + t);
  const float pm1 = __ldg(S.posterior_mean_coef1 + t);
  const float pm2 = __ldg(S.posterior_mean_coef2 + t);
  const float logv = __ldg(S.posterior_log_variance_clipped + t);
  // nonzero_mask * (0.5 * log_var).exp()   (ddpm.py:396-397)
  const float sigma = (t == 0) ? 0.f : expf(0.5f * logv);
  float z[4];
  if (a.noise) {
    const float* nz = a.noise + (int64_t)step * a.n + i0;
#pragma unroll
    for (int j = 0; j < 4; ++j) z[j] = (i0 + j < a.n) ? __ldg(nz + j) : 0.f;
  } else {
    uint32_t r[4];
    philox4x32_10(a.seed, (uint64_t)i4, (uint32_t)step, r);
    box_muller(r[0], r[1], z[0], z[1]);
    box_muller(r[2], r[3], z[2], z[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = i0 + j;
    if (i >= a.n) break;
    const float x = a.img[i];
    const float e = __ldg(a.eps + i);
    // torch evaluates mul, mul, sub / add as separate rounded fp32 ops: no fma contraction here
    float x0 = __fsub_rn(__fmul_rn(c1, x), __fmul_rn(c2, e));
    if (a.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
    const float mean = __fadd_rn(__fmul_rn(pm1, x0), __fmul_rn(pm2, x));
    a.img[i] = __fadd_rn(mean, __fmul_rn(sigma, z[j]));
  }
}

// state[0] = current t, state[1] = steps issued so far
__global__ void sampler_tick_kernel(int64_t* __restrict__ t_vec, int B, int* __restrict__ state) {
  const int t = state[0];
  for (int b = threadIdx.x; b < B; b += blockDim.x) t_vec[b] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    state[0] = t - 1;
    state[1] = state[1] + 1;
  }
}

__global__ void add_kernel(float* __restrict__ dst, const float* __restrict__ src, int64_t n4, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) {
    float4 d = reinterpret_cast<float4*>(dst)[i];
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
    d.x += s.x; d.y += s.y; d.z += s.z; d.w += s.w;
    reinterpret_cast<float4*>(dst)[i] = d;
  }
  if (i == 0)
    for (int64_t j = n4 * 4; j < n; ++j) dst[j] += src[j];
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                   float b1, float b2, float eps, float step_size,
                                                   float bc2_sqrt, float grad_scale) {
  // 128-bit body (arenas are 16-byte aligned: every tensor starts on a multiple of 4 floats), scalar tail
  const int64_t n4 = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(v)) & 15) ? 0 : (n >> 2);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
    float* mm = reinterpret_cast<float*>(&m4);
    float* vv = reinterpret_cast<float*>(&v4);
    float* pp = reinterpret_cast<float*>(&p4);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gi = gg[e] * grad_scale;
      const float mi = mm[e] + (gi - mm[e]) * (1.f - b1);
      const float vi = vv[e] * b2 + (1.f - b2) * gi * gi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      pp[e] = pp[e] - step_size * (mi / denom);
      mm[e] = mi;
      vv[e] = vi;
    }
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
  for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = m[i] + (gi - m[i]) * (1.f - b1);           // exp_avg.lerp_(grad, 1 - beta1)
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;          // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
  }
}

}  // namespace

int launch_input_prep(const LaunchCtx& lc, const float* x_nchw, const float* noise_nchw, const int64_t* t,
                      const float* sqrt_ac, const float* sqrt_1mac, float* out_nhwc, float* out_nchw, int B,
                      int C, int HW) {
  const int64_t n = (int64_t)B * HW;
  ProfScope ps_(lc, K_ELEM, 3.0 * n * C, 8.0 * n * C);
  input_prep_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, lc.stream>>>(x_nchw, noise_nchw, t, sqrt_ac, sqrt_1mac,
                                                                     out_nhwc, out_nchw, B, C, HW);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_final_conv(const LaunchCtx& lc, const float* act, const float* w, const float* b, float* out_nchw,
                      int B, int HW, int K, int C) {
  if (C > 4 || K % 4 != 0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "final conv: channels <= 4 and K % 4 == 0 required");
  const int64_t threads = (int64_t)B * HW * 8;
  ProfScope ps_(lc, K_CONV_FPROP, 2.0 * B * HW * (double)K * C, 4.0 * B * HW * (double)(K + C));
  final_conv_kernel<<<(unsigned)cdiv64(threads, 256), 256, (size_t)C * K * sizeof(float), lc.stream>>>(
      act, w, b, out_nchw, B, HW, K, C);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_loss(const LaunchCtx& lc, const float* pred, const float* noise, int64_t n, int loss_type, float* ws,
                float* loss_out) {
  ProfScope ps_(lc, K_ELEM, 2.0 * n, 8.0 * n);
  loss_partial_kernel<<<LOSS_CTAS, 256, 0, lc.stream>>>(pred, noise, n, loss_type, ws);
  IGM_POST_LAUNCH(lc);
  loss_final_kernel<<<1, 32, 0, lc.stream>>>(ws, LOSS_CTAS, n, loss_out);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_loss_backward_nhwc(const LaunchCtx& lc, const float* pred, const float* noise, int64_t n, int C,
                              int HW, int loss_type, const float* d_loss, float scale, float* d_nhwc) {
  ProfScope ps_(lc, K_ELEM, 2.0 * n, 12.0 * n);
  loss_backward_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, lc.stream>>>(pred, noise, n, C, HW, loss_type, d_loss,
                                                                        scale / (float)n, d_nhwc);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_nchw_to_nhwc(const LaunchCtx& lc, const float* src, float* dst, int B, int HW, int C) {
  const int64_t n = (int64_t)B * HW * C;
  ProfScope ps_(lc, K_ELEM, 0.0, 8.0 * n);
  nchw_to_nhwc_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, lc.stream>>>(src, dst, B, HW, C);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_nhwc_to_nchw(const LaunchCtx& lc, const float* src, float* dst, int B, int HW, int C) {
  const int64_t n = (int64_t)B * HW * C;
  ProfScope ps_(lc, K_ELEM, 0.0, 8.0 * n);
  nhwc_to_nchw_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, lc.stream>>>(src, dst, B, HW, C);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_sampler_update(const LaunchCtx& lc, const SamplerStepArgs& a) {
  const int64_t n4 = cdiv64(a.n, 4);
  ProfScope ps_(lc, K_ELEM, 12.0 * a.n, 4.0 * a.n * (a.noise ? 4 : 3));
  sampler_update_kernel<<<(unsigned)cdiv64(n4, 256), 256, 0, lc.stream>>>(a);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_sampler_tick(const LaunchCtx& lc, int64_t* t_vec, int B, int* state) {
  sampler_tick_kernel<<<1, 256, 0, lc.stream>>>(t_vec, B, state);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

// dst += src * scale * (*alpha_dev if given): the deferred half of an eagerly computed backward pass
__global__ void __launch_bounds__(256) axpy_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                    const float* __restrict__ alpha_dev, float scale, int64_t n4, int64_t n) {
  const float a = alpha_dev ? scale * __ldg(alpha_dev) : scale;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 s = __ldg(reinterpret_cast<const float4*>(src) + i);
    float4 d = reinterpret_cast<float4*>(dst)[i];
    d.x = fmaf(s.x, a, d.x); d.y = fmaf(s.y, a, d.y); d.z = fmaf(s.z, a, d.z); d.w = fmaf(s.w, a, d.w);
    reinterpret_cast<float4*>(dst)[i] = d;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - n4 * 4)) {
    const int64_t i = n4 * 4 + threadIdx.x;
    dst[i] = fmaf(src[i], a, dst[i]);
  }
}

int launch_axpy(const LaunchCtx& lc, float* dst, const float* src, const float* alpha_dev, float scale, int64_t n) {
  const int64_t n4 = n / 4;
  ProfScope ps_(lc, K_ELEM, 2.0 * n, 12.0 * n);
  int64_t blocks = cdiv64(n4 > 0 ? n4 : 1, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  axpy_kernel<<<(unsigned)blocks, 256, 0, lc.stream>>>(dst, src, alpha_dev, scale, n4, n);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_add(const LaunchCtx& lc, float* dst, const float* src, int64_t n) {
  const int64_t n4 = n / 4;
  ProfScope ps_(lc, K_ELEM, 1.0 * n, 12.0 * n);
  add_kernel<<<(unsigned)cdiv64(n4 > 0 ? n4 : 1, 256), 256, 0, lc.stream>>>(dst, src, n4, n);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_adam(const LaunchCtx& lc, float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1,
                float b2, float eps, int step, float grad_scale) {
  const double bc1 = 1.0 - pow((double)b1, (double)step);
  const double bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr / bc1);
  const float bc2_sqrt = (float)sqrt(bc2);
  int blocks = (int)cdiv64(n, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  ProfScope ps_(lc, K_ADAM, 12.0 * n, 28.0 * n);
  adam_kernel<<<blocks, 256, 0, lc.stream>>>(p, g, m, v, n, b1, b2, eps, step_size, bc2_sqrt, grad_scale);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
