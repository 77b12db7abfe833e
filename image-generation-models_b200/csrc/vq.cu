// Vector-quantiser of the reference VQ-VAE (src/models/vqvae.py:24-43) as one fused kernel:
// nearest code (first index on ties) + gather + squared-error partials.  The reference materialises
// the [N*h*w, K] distance matrix (268 MB at B=128, K=512) through torch.cdist; here distances live in
// registers only.  fp32 CUDA-core FMAs on purpose: the argmin must match the fp32 reference, and the
// whole lookup is 8.9 GFLOP per B=128 step.
#include "common.cuh"

namespace igm {
namespace {

constexpr int VQ_MAX_D = 64;

// Work split of the lookup.  The per-GPU problem is small (32768 vectors x 512 codes): one vector per thread is 7 warps per
// SM, each a 64-long dependent FMA chain fed by one broadcast LDS.128 per four FMAs -- latency- and LSU-bound at 14.8 TFLOP/s
// (20 % of the fp32 peak).  Here a thread owns V = 2 vectors (each code load feeds two chains) and ONE of S = 4 interleaved
// code slices (k = slice mod 4), so a 128-thread CTA covers 64 vectors and the grid is four times larger; the slices' winners
// are merged with "smaller distance, then smaller index", which is the first index over all codes because every slice scans
// its codes in increasing order with a strict '<'.  Per-(vector, code) arithmetic (order of the 64 FMAs, the distance formula)
// is unchanged, so the chosen codes are bit-identical to the one-vector kernel.  Codebook rows are padded to D + 4 floats:
// the four codes a warp reads at once then sit in different banks.
constexpr int VQ_V = 2;
constexpr int VQ_S = 4;
constexpr int VQ_THREADS = 128;
constexpr int VQ_VEC_PER_CTA = VQ_THREADS / VQ_S * VQ_V;   // 64

template <int D>
__global__ void __launch_bounds__(VQ_THREADS, 3) vq_forward_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                                int64_t* __restrict__ idx_out, float* __restrict__ quant,
                                                                float* __restrict__ ws, int N, int HW, int K, int k_tile) {
  constexpr int LD = D + 4;
  extern __shared__ __align__(16) float sm[];   // [k_tile][D + 4] codes | [k_tile] squared norms
  float* s_cb = sm;
  float* s_e2 = sm + (size_t)k_tile * LD;
  __shared__ float red[VQ_THREADS / 32];
  const int slice = threadIdx.x & (VQ_S - 1), grp = threadIdx.x >> 2;   // VQ_S == 4
  const int64_t nvec = (int64_t)N * HW;
  int64_t v[VQ_V];
  bool ok[VQ_V];
  int n[VQ_V], p[VQ_V];
  float x[VQ_V][D], x2[VQ_V], best[VQ_V];
  int best_k[VQ_V];
#pragma unroll
  for (int u = 0; u < VQ_V; ++u) {
    v[u] = (int64_t)blockIdx.x * VQ_VEC_PER_CTA + u * (VQ_THREADS / VQ_S) + grp;
    ok[u] = v[u] < nvec;
    n[u] = ok[u] ? (int)(v[u] / HW) : 0;
    p[u] = ok[u] ? (int)(v[u] - (int64_t)n[u] * HW) : 0;
    x2[u] = 0.f;
    best[u] = INFINITY;
    best_k[u] = 0x7fffffff;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      x[u][c] = ok[u] ? __ldg(z + ((int64_t)n[u] * D + c) * HW + p[u]) : 0.f;
      x2[u] = fmaf(x[u][c], x[u][c], x2[u]);
    }
  }
  for (int k0 = 0; k0 < K; k0 += k_tile) {
    const int kt = min(k_tile, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * (D / 4); i += blockDim.x) {
      const int k = i / (D / 4), c4 = i - k * (D / 4);
      *reinterpret_cast<float4*>(s_cb + k * LD + c4 * 4) = __ldg(reinterpret_cast<const float4*>(cb + (size_t)(k0 + k) * D) + c4);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kt; k += blockDim.x) {
      float e2 = 0.f;
      for (int c = 0; c < D; ++c) e2 = fmaf(s_cb[k * LD + c], s_cb[k * LD + c], e2);
      s_e2[k] = e2;
    }
    __syncthreads();
    for (int k = slice; k < kt; k += VQ_S) {
      const float4* e = reinterpret_cast<const float4*>(s_cb + k * LD);
      float dot[VQ_V];
#pragma unroll
      for (int u = 0; u < VQ_V; ++u) dot[u] = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < D / 4; ++c4) {
        const float4 ev = e[c4];
#pragma unroll
        for (int u = 0; u < VQ_V; ++u) {
          dot[u] = fmaf(x[u][c4 * 4 + 0], ev.x, dot[u]);
          dot[u] = fmaf(x[u][c4 * 4 + 1], ev.y, dot[u]);
          dot[u] = fmaf(x[u][c4 * 4 + 2], ev.z, dot[u]);
          dot[u] = fmaf(x[u][c4 * 4 + 3], ev.w, dot[u]);
        }
      }
      // torch.cdist (mm path): sqrt(clamp(|x|^2 + |e|^2 - 2 x.e, 0)); strict '<' keeps the first index of the slice on ties
      const float e2 = s_e2[k];
#pragma unroll
      for (int u = 0; u < VQ_V; ++u) {
        const float d = sqrtf(fmaxf(x2[u] + e2 - 2.f * dot[u], 0.f));
        if (d < best[u]) { best[u] = d; best_k[u] = k0 + k; }
      }
    }
  }
  // merge the four slices of a vector (adjacent lanes): smaller distance, then smaller index = first index overall
#pragma unroll
  for (int u = 0; u < VQ_V; ++u) {
#pragma unroll
    for (int o = 1; o < VQ_S; o <<= 1) {
      const float od = __shfl_xor_sync(0xffffffffu, best[u], o);
      const int ok_ = __shfl_xor_sync(0xffffffffu, best_k[u], o);
      if (od < best[u] || (od == best[u] && ok_ < best_k[u])) { best[u] = od; best_k[u] = ok_; }
    }
  }
  float err = 0.f;
#pragma unroll
  for (int u = 0; u < VQ_V; ++u) {
    if (!ok[u]) continue;
    if (slice == 0) idx_out[v[u]] = best_k[u];
    const float* q = cb + (size_t)best_k[u] * D;
    // the four lanes of a vector share the gather: lane `slice` writes channels slice, slice + 4, ...
#pragma unroll
    for (int c = 0; c < D; ++c) {
      if ((c & (VQ_S - 1)) != slice) continue;
      const float qc = __ldg(q + c);
      quant[((int64_t)n[u] * D + c) * HW + p[u]] = qc;
      const float df = x[u][c] - qc;
      err = fmaf(df, df, err);
    }
  }
  err = warp_sum(err);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < VQ_THREADS / 32; ++i) t += red[i];
    ws[blockIdx.x] = t;
  }
}

__global__ void vq_loss_final_kernel(const float* __restrict__ ws, int parts, double n_el, float beta,
                                     float* __restrict__ losses) {
  // one warp, fixed order (lane-strided partial sums, then a shuffle tree): deterministic
  double t = 0.0;
  for (int i = threadIdx.x; i < parts; i += 32) t += (double)ws[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const float mse = (float)(t / n_el);
    losses[0] = mse;          // vq_loss     = mse(z.detach(), q)        (vqvae.py:38)
    losses[1] = beta * mse;   // commit_loss = beta * mse(z, q.detach()) (vqvae.py:39)
  }
}

// dz = d_commit * beta * 2 (z - q) / n ; dE[idx] += d_vq * 2 (q - z) / n + d_quant
__global__ void vq_backward_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                   const int64_t* __restrict__ idx, const float* __restrict__ d_quant,
                                   const float* __restrict__ d_vq, const float* __restrict__ d_commit, float beta,
                                   float* __restrict__ dz, float* __restrict__ d_cb, int N, int D, int HW) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // NCHW element
  const int64_t total = (int64_t)N * D * HW;
  if (i >= total) return;
  const int64_t nc = i / HW;
  const int p = (int)(i - nc * HW);
  const int n = (int)(nc / D);
  const int c = (int)(nc - (int64_t)n * D);
  const int64_t k = idx[(int64_t)n * HW + p];
  const float q = __ldg(cb + k * D + c);
  const float diff = __ldg(z + i) - q;
  const float s = 2.f / (float)total;
  const float gv = d_vq ? __ldg(d_vq) : 0.f;
  const float gc = d_commit ? __ldg(d_commit) : 0.f;
  if (dz) dz[i] = gc * beta * s * diff;
  float ge = -gv * s * diff;
  if (d_quant) ge += __ldg(d_quant + i);
  if (d_cb) atomicAdd(d_cb + k * D + c, ge);
}

}  // namespace
}  // namespace igm

using namespace igm;

extern "C" int igm_vq_workspace_floats(int N, int HW) { return (int)cdiv64((int64_t)N * HW, VQ_VEC_PER_CTA) + 8; }

extern "C" int igm_vq_forward(const float* z, const float* codebook, int64_t* idx, float* quant, float* losses, int N,
                              int D, int HW, int K, float beta, float* ws, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!z || !codebook || !idx || !quant || !losses || !ws) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  if (N < 1 || HW < 1 || K < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad sizes");
  if (D != 16 && D != 32 && D != 64) IGM_FAIL(st, IGM_ERR_INVALID, "latent_dim must be 16, 32 or 64");
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  const int64_t nvec = (int64_t)N * HW;
  const int parts = (int)cdiv64(nvec, VQ_VEC_PER_CTA);
  // 256 codes per shared-memory chunk (66 KB at D = 64): three 128-thread CTAs per SM (168 registers per thread), i.e. six
  // independent FMA chains per scheduler
  int k_tile = K;
  const int max_codes = (64 * 1024) / (D * 4);
  if (k_tile > max_codes) k_tile = max_codes;
  const size_t smem = (size_t)k_tile * (D + 4 + 1) * sizeof(float);
  cudaError_t e = cudaSuccess;
  if (D == 64) {
    e = cudaFuncSetAttribute(vq_forward_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<64><<<parts, VQ_THREADS, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
  } else if (D == 32) {
    e = cudaFuncSetAttribute(vq_forward_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<32><<<parts, VQ_THREADS, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
  } else {
    e = cudaFuncSetAttribute(vq_forward_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<16><<<parts, VQ_THREADS, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
  }
  if (e != cudaSuccess) IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e));
  IGM_POST_LAUNCH(lc);
  vq_loss_final_kernel<<<1, 32, 0, lc.stream>>>(ws, parts, (double)nvec * D, beta, losses);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

extern "C" int igm_vq_backward(const float* z, const float* codebook, const int64_t* idx, const float* d_quant,
                               const float* d_vq, const float* d_commit, float beta, float* dz, float* d_codebook,
                               int N, int D, int HW, int K, void* stream) {
  Status& st = global_status();
  st = Status();
  (void)K;
  if (!z || !codebook || !idx) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  const int64_t total = (int64_t)N * D * HW;
  vq_backward_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, lc.stream>>>(z, codebook, idx, d_quant, d_vq, d_commit, beta,
                                                                         dz, d_codebook, N, D, HW);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}
