// Shared declarations for libigm_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/igm_b200.h"

namespace igm {

// ---------------------------------------------------------------------------
// error plumbing: CUDA failures are recorded with file:line and surface as
// IGM_ERR_CUDA through the C ABI (never abort / throw across the boundary).
// ---------------------------------------------------------------------------
struct Status {
  int code = IGM_OK;
  std::string msg;
};

void set_error(Status& st, int code, const char* file, int line, const char* what);
Status& global_status();   // errors raised without a context (igm_unet_create, igm_debug_*)
int64_t& ops_launch_counter();   // kernels launched by the context-free entry points (igm_conv2d_*, igm_vq_*, igm_pixelcnn_run, ...)

#define IGM_CUDA(st, expr)                                                      \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) {                                                    \
      ::igm::set_error((st), IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return (st).code;                                                         \
    }                                                                           \
  } while (0)

#define IGM_FAIL(st, errc, what)                                   \
  do {                                                             \
    ::igm::set_error((st), (errc), __FILE__, __LINE__, (what));    \
    return (errc);                                                 \
  } while (0)

#define IGM_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != IGM_OK) return _r; \
  } while (0)

// Kernel classes for the built-in event profiler (igm_profile_start/stop)
enum KClass {
  K_CONV_FPROP = 0, K_CONV_DGRAD, K_CONV_WGRAD, K_NORM, K_ATTN, K_TIME, K_ELEM, K_ADAM, K_PACK, K_NCLASS
};
const char* kclass_name(int k);

// CUDA-event profiler: one (start, stop) pair around every launch while enabled.  Used by
// bench.py's roofline pass only (never during the timed region).
struct Profiler {
  bool on = false;
  struct Rec { cudaEvent_t a, b; int cls; double flops, bytes; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get();
  void reset();
};

// Launch bookkeeping: every kernel launch goes through this so that
// igm_launch_count() is exact and launch errors are caught where they happen.
struct LaunchCtx {
  cudaStream_t stream = 0;
  Status* st = nullptr;
  int64_t* counter = nullptr;
  Profiler* prof = nullptr;
};

struct ProfScope {
  Profiler* p = nullptr;
  cudaStream_t s = 0;
  size_t idx = 0;
  ProfScope(const LaunchCtx& lc, int cls, double flops, double bytes) {
    if (lc.prof && lc.prof->on) {
      p = lc.prof;
      s = lc.stream;
      Profiler::Rec r{p->get(), p->get(), cls, flops, bytes};
      idx = p->recs.size();
      p->recs.push_back(r);
      cudaEventRecord(r.a, s);
    }
  }
  ~ProfScope() {
    if (p) cudaEventRecord(p->recs[idx].b, s);
  }
};

inline int post_launch(const LaunchCtx& lc, const char* file, int line) {
  if (lc.counter) ++*lc.counter;
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(*lc.st, IGM_ERR_CUDA, file, line, cudaGetErrorString(e));
    return lc.st->code;
  }
  return IGM_OK;
}
#define IGM_POST_LAUNCH(lc) IGM_TRY(::igm::post_launch((lc), __FILE__, __LINE__))

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining
// (launch latency and the kernel's own prologue overlap the predecessor's tail); the kernel MUST execute pdl_wait()
// before it touches anything the predecessor wrote.  IGM_PDL=0 launches such kernels the ordinary way.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#ifdef __CUDACC__
// -DIGM_PDL_EARLY_TRIGGER=1: after the wait, let the successor's CTAs be scheduled as soon as every CTA of THIS grid has
// reached this point (its prologue then overlaps this kernel's body, not just its tail; the trigger comes AFTER the wait,
// so whatever a successor touches before its own wait was written at least two kernels back).  Measured slower on the
// CIFAR-10 step (196.5 -> 193.6 steps/s, sampler 77.8 -> 76.6 samples/s): waiting CTAs hold SM slots.  Off.
#ifndef IGM_PDL_EARLY_TRIGGER
#define IGM_PDL_EARLY_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#if IGM_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
#endif

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

constexpr int kGroups = 8;          // reference ddpm.py:113 / :132-133
constexpr int kHeads = 4;           // reference ddpm.py:147
constexpr int kDimHead = 32;        // reference ddpm.py:147
constexpr float kGnEps = 1e-5f;     // torch.nn.GroupNorm default
constexpr float kLnEps = 1e-5f;     // reference ddpm.py:86
constexpr int kGnChunk = 64;        // pixels per GroupNorm partial-statistics chunk (launch_gn_partial layout)
constexpr int kGnChunkMin = 16;     // smallest pixel chunk a GroupNorm CTA works on (workspace sizing)

// ---------------------------------------------------------------------------
// Generic implicit-GEMM convolution (gather form) on NHWC fp32 activations.
//   out[m, n] = bias[n] + sum_{tap, k} in[pix(m, tap), k] * w[tap][k][n] (+ add[m, n])
// m enumerates output pixels (b, oy, ox); k runs over the (optionally
// concatenated) input channels.  Two gather modes cover every conv of the
// path and its data-gradient:
//   transposed = 0:  iy = oy*stride - pad + ky*dil         (Conv2d fprop, ConvTranspose2d dgrad)
//   transposed = 1:  iy = (oy + pad - ky*dil) / stride      (ConvTranspose2d fprop, Conv2d dgrad)
// ---------------------------------------------------------------------------
struct ConvArgs {
  const float* in0 = nullptr;   // [B, IH, IW, C0]
  const float* in1 = nullptr;   // [B, IH, IW, C1] second source of a channel concat (or null)
  int C0 = 0, C1 = 0;
  int B = 0, IH = 0, IW = 0, OH = 0, OW = 0;
  int N = 0;                    // output channels
  int KH = 1, KW = 1, stride = 1, pad = 0, dil = 1;
  int pad_w = -1;               // horizontal padding when it differs from `pad` (rectangular kernels); -1: same
  int transposed = 0;
  const float* w = nullptr;     // packed [KH*KW][C0+C1][N]
  const float* bias = nullptr;  // [N] or null
  float* out0 = nullptr;        // n <  N0 -> out0[m*N0 + n]
  float* out1 = nullptr;        // n >= N0 -> out1[m*(N-N0) + n-N0]   (null when N0 == N)
  int N0 = 0;
  int kclass = K_CONV_FPROP;    // profiler class (fprop / dgrad)
  const float* add0 = nullptr;  // optional addends with the same split
  const float* add1 = nullptr;
};
int launch_conv(const LaunchCtx& lc, const ConvArgs& a);

// Weight gradient of the same convolution family.
//   g[tap][qc][pc] += sum_pix Q[gather(pix, tap), qc] * P[pix, pc]
// P is the tensor enumerated pixel by pixel, Q the gathered one
// (iy = py*stride - pad + ky*dil).  The result is ACCUMULATED (atomicAdd) at
//   grad[qc*sq + pc*sp + tap]
// which addresses PyTorch's OIHW (Conv2d) or IOHW (ConvTranspose2d) layouts directly.
struct WgradArgs {
  const float* P = nullptr;  int PC = 0, PH = 0, PW = 0;   // [B, PH, PW, PC]
  const float* Q = nullptr;  int QC = 0, QH = 0, QW = 0;   // [B, QH, QW, QC]
  int B = 0;
  int KH = 1, KW = 1, stride = 1, pad = 0, dil = 1;
  int pad_w = -1;
  float* grad = nullptr;
  int64_t sq = 0, sp = 0;
};
int launch_wgrad(const LaunchCtx& lc, const WgradArgs& a);

// colsum: out[n] += sum_m x[m, n]   (bias gradients)
int launch_colsum(const LaunchCtx& lc, const float* x, int64_t M, int N, float* out);

// Weight packing: dst[tap][k][n] = src[k*sk + n*sn + tap]
int launch_pack_weight(const LaunchCtx& lc, const float* src, float* dst, int taps, int K, int N,
                       int64_t sk, int64_t sn);

// All weight re-packs of a network in ONE launch: a device table of jobs, each either
//   fp32  dst_f[tap][k][n]              (SIMT engine)            or
//   bf16  dst_hi/lo[n][tap*K + k]       (tcgen05 engine, optional tap flip)
// with element (n, tap, k) read from src[k*sk + n*sn + tap'].
struct PackJob {
  const float* src;
  float* dst_f;
  void* dst_hi;
  void* dst_lo;
  int taps, K, N, flip;
  int64_t sk, sn;
  int64_t begin;   // prefix sum of element counts
};
int launch_pack_jobs(const LaunchCtx& lc, const PackJob* d_jobs, int n_jobs, int64_t total);
// tiled variant for the bf16 hi/lo layouts with N, K multiples of 32: job.begin = first CTA, d_cta_job[cta] = job index
int launch_pack_tiles(const LaunchCtx& lc, const PackJob* d_jobs, const int* d_cta_job, int n_ctas, int max_taps, double elems);

// ---------------------------------------------------------------------------
// normalisation / activation kernels (norm_act.cu)
// ---------------------------------------------------------------------------
// GroupNorm(8) partial statistics: part[b][chunk][g] = (sum, sumsq) over a chunk of kGnChunk pixels.
int launch_gn_partial(const LaunchCtx& lc, const float* y, int B, int HW, int C, float* part);
// Finalise statistics (stats[b][g] = mean, rstd), then out = mish(gn(y)) [+ temb[b, c]] [+ res[m, c]].
// temb has row stride temb_stride.
// out_hi / out_lo (optional): bf16 hi/lo copy of `out` for the tensor-core conv that consumes it.
int launch_gn_apply(const LaunchCtx& lc, const float* y, const float* part, const float* gamma,
                    const float* beta, const float* temb, int temb_stride, const float* res,
                    float* out, float* stats, int B, int HW, int C, __nv_bfloat16* out_hi = nullptr,
                    __nv_bfloat16* out_lo = nullptr, int nparts = 0 /* partial slots per image; 0 = 64-pixel chunks */);
// Backward of the above.  d_out: grad of `out`.  Produces dy (grad of conv output y),
// accumulates dgamma/dbeta into the grad arena, and (optionally) dtemb[b, c] (+= over pixels).
struct GnBwdArgs {
  const float* d_out; const float* y; const float* stats; const float* gamma; const float* beta;
  float* dy; float* dgamma; float* dbeta; float* dtemb; int dtemb_stride;
  float* dbias;       // optional: += column sums of dy (gradient of the bias of the conv feeding this norm)
  __nv_bfloat16* dy_hi; __nv_bfloat16* dy_lo;   // optional bf16 hi/lo copy of dy (tensor-core wgrad / dgrad operand)
  // optional by-products on d_out itself (fused kernel only, see gn_backward_is_fused): its bf16 hi/lo staging copy and
  // its column sums -- the dY operand and the bias gradient of a ResnetBlock's res_conv, which shares d_out with block2
  __nv_bfloat16* dout_hi = nullptr; __nv_bfloat16* dout_lo = nullptr; float* dout_colsum = nullptr;
  float* ws_group;    // [B][chunks][G][2]
  float* ws_chan;     // [B][chunks][C][3]
  int B, HW, C;
};
int launch_gn_backward(const LaunchCtx& lc, const GnBwdArgs& a);
bool gn_backward_is_fused(const GnBwdArgs& a);   // will launch_gn_backward take the single-kernel path for this shape?

// Channel LayerNorm of reference ddpm.py:85-95 (eps added to std).  x,out: [M, C]
int launch_ln_forward(const LaunchCtx& lc, const float* x, const float* g, const float* b, float* out,
                      int64_t M, int C, __nv_bfloat16* out_hi = nullptr, __nv_bfloat16* out_lo = nullptr);
// dx = d_res + LN'(d_out)  (the Residual branch add is fused); dg/db accumulated via ws partials.
int launch_ln_backward(const LaunchCtx& lc, const float* d_out, const float* x, const float* g,
                       const float* d_res, float* dx, float* dg, float* db, float* ws, int64_t M, int C, bool finalize = true);
int launch_ln_param_finalize(const LaunchCtx& lc, const float* ws, int64_t M, int C, float* dg, float* db);
int ln_backward_parts(int64_t M);   // CTAs (= partial rows of ws [parts][2][C]) the backward kernel uses

// ---------------------------------------------------------------------------
// linear attention (attention.cu) — reference ddpm.py:154-166
// qkv: [B, n, 384] (q | k | v, each heads*32), out: [B, n, 128]
// ctx: [B, heads, 32, 32], kstat: [B, heads, 32, 2] (max, sum of exp)
// ---------------------------------------------------------------------------
// ws: linattn_ws_floats(B, n) floats of scratch (per-chunk partial statistics / contexts)
int linattn_ws_floats(int B, int n);
int launch_linattn_forward(const LaunchCtx& lc, const float* qkv, float* out, float* ctx, float* kstat,
                           int B, int n, float* ws, __nv_bfloat16* out_hi = nullptr, __nv_bfloat16* out_lo = nullptr);
// inference shortcut (attention.cu: linattn_mb_kernel): context only, then the per-image matrix M_b = W_out ctx^T W_q
// kv_only: `qkv` is a compact [M, 256] k | v tensor (the sampler's per-image-matrix path never forms q)
int launch_linattn_ctx(const LaunchCtx& lc, const float* qkv, float* ctx, float* kstat, int B, int n, float* ws,
                       bool kv_only = false);
// tensor-core attention path (training, n >= 1024): glue kernels around the per-image 1x1 convs of conv_tc.cu; q / k / v
// and dO come as bf16 hi / lo staging pairs ([M, 384] resp. [M, 128]), d(qkv) leaves as thirds of a [M, 384] pair
int launch_linattn_ctx_hl(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, float* ctx, float* kstat,
                          int B, int n, float* ws);
float* linattn_dctx_ptr(float* ws);
int launch_linattn_dctx_hl(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, const __nv_bfloat16* d_hi,
                           const __nv_bfloat16* d_lo, int B, int n, float* ws);
// per-image block-diagonal 128 x 128 weight matrices (bf16 hi / lo, [B * 128 rows][128]) from m [B, 4, 32, 32];
// cc (nullable) [B, 128] = sum_e m[d][e] * m2[d][e]
int launch_linattn_wt(const LaunchCtx& lc, const float* m, int transpose, int B, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo,
                      const float* m2 = nullptr, float* cc = nullptr);
int launch_linattn_p(const LaunchCtx& lc, const __nv_bfloat16* q_hi, const __nv_bfloat16* q_lo, const float* kstat,
                     __nv_bfloat16* p_hi, __nv_bfloat16* p_lo, int B, int n);
int launch_linattn_dk(const LaunchCtx& lc, const __nv_bfloat16* p_hi, const __nv_bfloat16* p_lo, const float* T, const float* cc,
                      __nv_bfloat16* d_hi, __nv_bfloat16* d_lo, int B, int n);
int launch_linattn_mb(const LaunchCtx& lc, const float* ctx, const float* w_out, const float* w_q, int B, int C,
                      __nv_bfloat16* mb_hi, __nv_bfloat16* mb_lo);
int launch_linattn_backward(const LaunchCtx& lc, const float* qkv, const float* ctx, const float* kstat,
                            const float* d_out, float* d_qkv, int B, int n, float* ws,
                            __nv_bfloat16* d_hi = nullptr, __nv_bfloat16* d_lo = nullptr);
// (d_qkv may be null when d_hi / d_lo are given: the gradient is then only emitted as the bf16 hi/lo staging copy)

// ---------------------------------------------------------------------------
// time embedding MLP (time_mlp.cu) — reference ddpm.py:47-59, :188-193, :126-130
// ---------------------------------------------------------------------------
struct TimeMlpParams {
  const float* w1; const float* b1;   // [4d, d], [4d]
  const float* w2; const float* b2;   // [d, 4d], [d]
  float* gw1; float* gb1; float* gw2; float* gb2;
  int dim;
};
struct TimeProj {                      // one per ResnetBlock: Linear(dim, cout) after Mish
  const float* w; const float* b; float* gw; float* gb; int cout; int offset;
};
// emb:[B,d] h1:[B,4d] (pre-activation) temb:[B,d] act:[B,d]=mish(temb) proj:[B,total]
int launch_time_mlp_forward(const LaunchCtx& lc, const TimeMlpParams& p, const int64_t* t, int B,
                            float* emb, float* h1, float* temb, float* act);
int launch_time_proj_forward(const LaunchCtx& lc, const TimeProj* d_table, int n_proj, const float* act,
                             int dim, int B, int total, float* proj);
// ws: [Bcap, 10 d] floats, layout fixed by the PLANNED batch Bcap (see time_mlp.cu); d_act accumulator zero between passes
int launch_time_proj_backward(const LaunchCtx& lc, const TimeMlpParams& p, const TimeProj* d_table, int n_proj, int total,
                              int B, int Bcap, const float* act, const float* d_proj, float* ws, int slab_lo, int slab_hi);
int launch_time_mlp_backward(const LaunchCtx& lc, const TimeMlpParams& p, int B, int Bcap, const float* emb, const float* h1,
                             const float* temb, float* ws);

// ---------------------------------------------------------------------------
// boundary + diffusion elementwise (diffusion.cu)
// ---------------------------------------------------------------------------
// NCHW -> NHWC with optional q_sample fused: out = a[t_b]*x + s[t_b]*noise (ddpm.py:441-444)
int launch_input_prep(const LaunchCtx& lc, const float* x_nchw, const float* noise_nchw,
                      const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac, float* out_nhwc,
                      float* out_nchw, int B, int C, int HW);
// final Conv1x1(dim -> C) of ddpm.py:236, writing NCHW.
int launch_final_conv(const LaunchCtx& lc, const float* act, const float* w, const float* b,
                      float* out_nchw, int B, int HW, int K, int C);
// loss = mean |noise - pred| (l1) or mean (noise - pred)^2 (l2); deterministic two-stage reduce.
int launch_loss(const LaunchCtx& lc, const float* pred, const float* noise, int64_t n, int loss_type,
                float* ws, float* loss_out);
// d_pred (written NHWC [B*HW, C]) = scale * dloss/dpred; pred/noise are NCHW with n elements
int launch_loss_backward_nhwc(const LaunchCtx& lc, const float* pred, const float* noise, int64_t n, int C,
                              int HW, int loss_type, const float* d_loss, float scale, float* d_nhwc);
int launch_nchw_to_nhwc(const LaunchCtx& lc, const float* src, float* dst, int B, int HW, int C);
// fused p_sample tail (ddpm.py:359-364, :385, :367-376, :394-397); t read from *t_dev (int64 [B], uniform)
struct SamplerStepArgs {
  float* img;            // [B,C,H,W] in/out
  const float* eps;      // U-Net output
  const float* noise;    // injected noise for this step or null
  const int64_t* t_dev;  // [B] (all equal)
  const igm_schedule* sched_dev;  // device copy of the pointer table
  uint64_t seed;
  const int* step_dev;   // device step counter (for Philox offset)
  int64_t n; int per_sample; int clip;
};
int launch_sampler_update(const LaunchCtx& lc, const SamplerStepArgs& a);
// t[b] = *t_scalar for all b; then (*t_scalar)--, (*step)++   (device-side loop state for graph replay)
int launch_sampler_tick(const LaunchCtx& lc, int64_t* t_vec, int B, int* state /* [t, step] */);
int launch_add(const LaunchCtx& lc, float* dst, const float* src, int64_t n);
int launch_axpy(const LaunchCtx& lc, float* dst, const float* src, const float* alpha_dev, float scale, int64_t n);
int launch_adam(const LaunchCtx& lc, float* p, const float* g, float* m, float* v, int64_t n, float lr,
                float b1, float b2, float eps, int step, float grad_scale);
// NHWC -> NCHW copy (debug taps, d_x)
int launch_nhwc_to_nchw(const LaunchCtx& lc, const float* src, float* dst, int B, int HW, int C);

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
// Mish (reference ddpm.py:62-64): x * tanh(softplus(x)), F.softplus(beta=1, threshold=20).
// With w = e^x:  tanh(log(1 + w)) = ((1+w)^2 - 1) / ((1+w)^2 + 1) = n / (n + 2),  n = w (w + 2),
// so one exp and one division replace exp + log1p + tanh (the GroupNorm kernels were ALU-bound on
// those three library calls).  Above the softplus threshold tanh(x) == 1 in fp32, so mish(x) = x.
// Branch-free: ex2.approx / rcp.approx (~2 ulp, far inside the 1e-3 parity budget) and a select; for x <= 20 the
// operands stay in the normal range (n + 2 in [2, 2.4e17]), so no range handling is needed.
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mish_f(float x) {
  const float w = ex2_approx(x * 1.4426950408889634f);
  const float n = w * (w + 2.f);
  const float m = x * (n * rcp_approx(n + 2.f));
  return x > 20.f ? x : m;
}
// d/dx mish = t + x * dt/dx,  t = n/(n+2),  dt/dx = 2 n' / (n+2)^2,  n' = 2 w (w + 1)
// (kept with its early-out branch: the branch-free form lets ptxas interleave all 32 elements of the fused GroupNorm
// backward and pushes it from 128 to 176 registers, one resident CTA per SM instead of two)
__device__ __forceinline__ float mish_grad_f(float x) {
  if (x > 20.f) return 1.f;
  const float w = ex2_approx(x * 1.4426950408889634f);
  const float n = w * (w + 2.f);
  const float r = rcp_approx(n + 2.f);
  return n * r + x * (4.f * w * (w + 1.f)) * (r * r);
}
// branch-free form for the kernels whose tiles live in shared memory (registers are not the constraint there): the
// exponent argument is clamped at 20, where t = 1 - 2/n is already 1.0f and the x * dt/dx term is < 1e-15, so the value
// equals the early-out's 1.0f; without the per-element branch the four chains of a float4 interleave.
__device__ __forceinline__ float mish_grad_nb_f(float x) {
  const float w = ex2_approx(fminf(x, 20.f) * 1.4426950408889634f);
  const float n = w * (w + 2.f);
  const float r = rcp_approx(n + 2.f);
  return n * r + x * (4.f * w * (w + 1.f)) * (r * r);
}
// bf16 hi/lo staging of four consecutive fp32 values (operands of the tcgen05 bf16x3 engine)
// (a, b) -> packed bf16 hi pair and packed bf16 lo pair (element a in the low half: lower address); two values per cvt
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));
  const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - bh), "f"(a - ah));
}
__device__ __forceinline__ void store_split4(__nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                             int64_t off, const float4& o) {
  uint2 h, l;
  split_pair(o.x, o.y, h.x, l.x);
  split_pair(o.z, o.w, h.y, l.y);
  *reinterpret_cast<uint2*>(hi + off) = h;
  *reinterpret_cast<uint2*>(lo + off) = l;
}
// Packed fp32 FMA (Blackwell FFMA2): d.x += a.x * b.x, d.y += a.y * b.y in ONE issue slot.  A three-register FFMA issues
// every other cycle per scheduler on sm_100; the packed form carries two FMAs per issue, so FMA-issue-bound loops
// (linear attention) run up to twice as fast.  Each lane's rounding is that of a scalar fmaf.
__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  uint64_t dd = *reinterpret_cast<uint64_t*>(&d);
  const uint64_t aa = *reinterpret_cast<const uint64_t*>(&a), bb = *reinterpret_cast<const uint64_t*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
#endif

}  // namespace igm
