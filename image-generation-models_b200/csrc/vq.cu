// Vector-quantiser of the reference VQ-VAE (src/models/vqvae.py:24-43) as one fused kernel:
// nearest code (first index on ties) + gather + squared-error partials.  The reference materialises
// the [N*h*w, K] distance matrix (268 MB at B=128, K=512) through torch.cdist; here distances live in
// registers only.  fp32 CUDA-core FMAs on purpose: the argmin must match the fp32 reference, and the
// whole lookup is 8.9 GFLOP per B=128 step.
#include "common.cuh"

namespace igm {
namespace {

constexpr int VQ_MAX_D = 64;

// one thread per latent vector; the codebook (chunk) is broadcast from shared memory.
// (Round 2 tried two vectors per thread, and two vectors x four interleaved code slices per thread group with padded codebook
// rows: 0.174 / 0.169 ms against 0.145 ms for this kernel on the C5 per-GPU problem, 32768 vectors x 512 codes -- the problem
// is one wave of 128 CTAs either way, and the finer split only added a partial second wave.  Reverted.)
template <int D>
__global__ void __launch_bounds__(256) vq_forward_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                         int64_t* __restrict__ idx_out, float* __restrict__ quant,
                                                         float* __restrict__ ws, int N, int HW, int K, int k_tile) {
  extern __shared__ __align__(16) float sm[];   // [k_tile][D] codes | [k_tile] squared norms
  float* s_cb = sm;
  float* s_e2 = sm + (size_t)k_tile * D;
  __shared__ float red[8];
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nvec = (int64_t)N * HW;
  const bool ok = v < nvec;
  const int n = ok ? (int)(v / HW) : 0;
  const int p = ok ? (int)(v - (int64_t)n * HW) : 0;
  float x[D];
  float x2 = 0.f;
#pragma unroll
  for (int c = 0; c < D; ++c) {
    x[c] = ok ? __ldg(z + ((int64_t)n * D + c) * HW + p) : 0.f;
    x2 = fmaf(x[c], x[c], x2);
  }
  float best = INFINITY;
  int best_k = 0;
  for (int k0 = 0; k0 < K; k0 += k_tile) {
    const int kt = min(k_tile, K - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < kt * D / 4; i += blockDim.x)
      reinterpret_cast<float4*>(s_cb)[i] = __ldg(reinterpret_cast<const float4*>(cb + (size_t)k0 * D) + i);
    __syncthreads();
    for (int k = threadIdx.x; k < kt; k += blockDim.x) {
      float e2 = 0.f;
      for (int c = 0; c < D; ++c) e2 = fmaf(s_cb[k * D + c], s_cb[k * D + c], e2);
      s_e2[k] = e2;
    }
    __syncthreads();
    for (int k = 0; k < kt; ++k) {
      const float4* e = reinterpret_cast<const float4*>(s_cb + k * D);
      float dot = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < D / 4; ++c4) {
        const float4 ev = e[c4];
        dot = fmaf(x[c4 * 4 + 0], ev.x, dot);
        dot = fmaf(x[c4 * 4 + 1], ev.y, dot);
        dot = fmaf(x[c4 * 4 + 2], ev.z, dot);
        dot = fmaf(x[c4 * 4 + 3], ev.w, dot);
      }
      // torch.cdist (mm path): sqrt(clamp(|x|^2 + |e|^2 - 2 x.e, 0)); strict '<' keeps the first index on ties
      const float d = sqrtf(fmaxf(x2 + s_e2[k] - 2.f * dot, 0.f));
      if (d < best) { best = d; best_k = k0 + k; }
    }
  }
  float err = 0.f;
  if (ok) {
    idx_out[v] = best_k;
    const float* q = cb + (size_t)best_k * D;
#pragma unroll
    for (int c = 0; c < D; ++c) {
      const float qc = __ldg(q + c);
      quant[((int64_t)n * D + c) * HW + p] = qc;
      const float df = x[c] - qc;
      err = fmaf(df, df, err);
    }
  }
  err = warp_sum(err);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = err;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i];
    ws[blockIdx.x] = t;
  }
}

__global__ void vq_loss_final_kernel(const float* __restrict__ ws, int parts, double n_el, float beta,
                                     float* __restrict__ losses) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < parts; ++i) t += (double)ws[i];
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

extern "C" int igm_vq_workspace_floats(int N, int HW) { return (int)cdiv64((int64_t)N * HW, 256) + 8; }

extern "C" int igm_vq_forward(const float* z, const float* codebook, int64_t* idx, float* quant, float* losses, int N,
                              int D, int HW, int K, float beta, float* ws, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!z || !codebook || !idx || !quant || !losses || !ws) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  if (N < 1 || HW < 1 || K < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad sizes");
  if (D != 16 && D != 32 && D != 64) IGM_FAIL(st, IGM_ERR_INVALID, "latent_dim must be 16, 32 or 64");
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &ops_launch_counter();
  const int64_t nvec = (int64_t)N * HW;
  const int parts = (int)cdiv64(nvec, 256);
  int k_tile = K;
  const int max_codes = (160 * 1024) / ((D + 1) * 4);
  if (k_tile > max_codes) k_tile = max_codes;
  const size_t smem = (size_t)k_tile * (D + 1) * sizeof(float);
  cudaError_t e = cudaSuccess;
  if (D == 64) {
    e = cudaFuncSetAttribute(vq_forward_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<64><<<parts, 256, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
  } else if (D == 32) {
    e = cudaFuncSetAttribute(vq_forward_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<32><<<parts, 256, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
  } else {
    e = cudaFuncSetAttribute(vq_forward_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) vq_forward_kernel<16><<<parts, 256, smem, lc.stream>>>(z, codebook, idx, quant, ws, N, HW, K, k_tile);
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
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &ops_launch_counter();
  const int64_t total = (int64_t)N * D * HW;
  vq_backward_kernel<<<(unsigned)cdiv64(total, 256), 256, 0, lc.stream>>>(z, codebook, idx, d_quant, d_vq, d_commit, beta,
                                                                         dz, d_codebook, N, D, HW);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}
