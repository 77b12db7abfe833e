// Incremental PixelCNN engine (reference src/models/pixelcnn.py): one persistent kernel, one CTA
// per image, walks the H x W raster once.
//
// The reference sampler (pixelcnn.py:167-195) re-runs the whole network on the top h+1 rows for each
// of the H*W pixels (784 forwards, 843 GFLOP/sample on MNIST).  The masked convolutions make the
// logits at (h, w) depend only on already generated pixels, so the same result is obtained by
//   * once per row h: the vertical stack of all 12 layers for that row (it only sees rows < h of the
//     image), plus the v->h link conv1x1_1 of every layer;
//   * once per pixel: one column of the horizontal stack (11 gated layers) and the 1x1 head,
// with per-layer feature rows cached in HBM/L2 (~2.1 GFLOP/sample in total).  The same walk in
// teacher-forced mode (pixels given, nothing drawn) IS PixelCNN.forward (pixelcnn.py:128-154).
// fp32 CUDA-core arithmetic throughout: sampled pixels must match the fp32 reference decisions.
#include "common.cuh"

namespace igm {
namespace {

constexpr int NLAYERS = 11;
__constant__ int c_dil[NLAYERS] = {1, 2, 1, 4, 1, 2, 1, 4, 1, 2, 1};   // pixelcnn.py:108-122
constexpr int PB = 7;         // output positions per thread and pass in the row GEMMs (28 = 4 x 7, 32 = 2 x 2 x 7 + ...)
constexpr int PADW = 4;       // zero columns on either side of a staged input row (largest dilation)
constexpr int ROWX = 2 * PADW + PB;   // extra columns of a staged row: both paddings + slack for positions past W
constexpr int MAX_W = 64;
constexpr int MAX_HD = 128;

struct PcnnOffsets {
  int64_t vs0_w, vs0_b, hs0_w, hs0_b;
  int64_t vert_w[NLAYERS], vert_b[NLAYERS], v2h_w[NLAYERS], v2h_b[NLAYERS];
  int64_t horiz_w[NLAYERS], horiz_b[NLAYERS], h2_w[NLAYERS], h2_b[NLAYERS];
  int64_t out_w, out_b, total;
};

PcnnOffsets make_offsets(int C, int Hd) {
  PcnnOffsets o;
  int64_t p = 0;
  auto take = [&](int64_t n) { int64_t r = p; p += (n + 3) & ~int64_t(3); return r; };
  o.vs0_w = take((int64_t)10 * C * Hd); o.vs0_b = take(Hd);
  o.hs0_w = take((int64_t)2 * C * Hd); o.hs0_b = take(Hd);
  for (int i = 0; i < NLAYERS; ++i) {
    o.vert_w[i] = take((int64_t)6 * Hd * 2 * Hd); o.vert_b[i] = take(2 * Hd);
    o.v2h_w[i] = take((int64_t)2 * Hd * 2 * Hd); o.v2h_b[i] = take(2 * Hd);
    o.horiz_w[i] = take((int64_t)2 * Hd * 2 * Hd); o.horiz_b[i] = take(2 * Hd);
    o.h2_w[i] = take((int64_t)Hd * Hd); o.h2_b[i] = take(Hd);
  }
  o.out_w = take((int64_t)Hd * 256 * C); o.out_b = take(256 * C);
  o.total = p;
  return o;
}

struct PcnnArgs {
  const float* Wt;           // packed weights, every matrix [K][N] (N contiguous), see make_offsets
  PcnnOffsets off;
  float* img;                // [N, C, H, W] in/out
  const float* uniforms;     // [H*W][N*C] or null
  const uint8_t* skip;       // [H*W]: 1 = keep the given pixel (reference :185) or null
  const float* cond;         // [NLAYERS][2 (vert, horiz)][N][2*Hd] class-conditioning pre-gate addends (:71,:79) or null
  float* logits;             // [N, 256, C, H, W] or null
  float* ws;                 // per-image workspace
  int64_t ws_per_img;
  uint64_t seed;
  int N, C, H, W, Hd;
  int mode;                  // 0 inverse-CDF draw, 1 greedy argmax, 2 teacher forced (no draw)
  int normalize;             // pixel value k/255 (0) or 2k/255 - 1 (1)
  int early;                 // conv1x1_2 weight slices loaded next to horiz_conv's (IGM_PCNN_EARLY=0: at their use)
  int prof;                  // IGM_PCNN_PROF=1: thread 0 of every CTA adds its row-pass / pixel-chain / head cycles to g_pcnn_prof
};

// y[n] = bias[n] + extra[n] + sum_k Wt[k][n] * x[k]   for n < N (N = 64, 128 or 256 per call), all 256 threads.
// The per-pixel chain of 22 dependent GEMVs is bound by the latency of the weight loads (the matrices stream from L2:
// 901 KB per pixel step do not fit shared memory), so the loads are made wide and independent: a thread owns FOUR
// adjacent outputs and one K slice (N/4 threads per slice, 256 / (N/4) slices), i.e. K * N / 1024 <= 16 128-bit loads,
// all in flight before the first FMA.  Slice partials are summed in a fixed order (deterministic).  Measured and dropped:
// fetching the NEXT matrix of the chain ahead of time, into registers (+1 % while the chain still had global loads of small
// operands between its barriers, see the staging arrays of the kernel) or with cp.async into thread-private shared-memory
// slots a whole layer ahead -- once those small loads were staged, both pipelines LOST 6 % against this plain form
// (1602 vs 1708 samples/s).  Keeping nine of the eleven conv1x1_2 matrices resident in the SM's spare shared memory changed
// nothing either (1753 vs 1755): the weight fetches are not what the chain waits for.
template <int N, int K>
__device__ __forceinline__ void cta_gemv_t(const float* __restrict__ Wt, const float* __restrict__ bias,
                                           const float* __restrict__ extra, const float* x_s, float* red_s /*[1024]*/,
                                           float* y_s) {
  // compile-time shape: the slice arithmetic (three integer divisions in the run-time form) folds away and the loads of a
  // slice are addressed with immediates
  constexpr int n4 = N >> 2;             // threads per K slice
  constexpr int parts = 256 / n4;        // K slices
  constexpr int klen = K / parts;        // K, N are powers of two >= 64: divides exactly
  constexpr int CH = klen < 16 ? klen : 16;   // rows per batch of loads (all in flight before the first FMA)
  static_assert(klen >= 1 && klen % CH == 0, "a thread's K slice is a multiple of its load batch");
  const int tid = threadIdx.x;
  const int part = tid / n4, q = tid % n4;
  const int kb = part * klen;
  const float4* wp = reinterpret_cast<const float4*>(Wt + (size_t)kb * N) + q;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int k0 = 0; k0 < klen; k0 += CH) {
    float4 wv[CH];
#pragma unroll
    for (int k = 0; k < CH; ++k) wv[k] = __ldg(wp + (k0 + k) * n4);
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const float x = x_s[kb + k0 + k];
      acc.x = fmaf(wv[k].x, x, acc.x); acc.y = fmaf(wv[k].y, x, acc.y);
      acc.z = fmaf(wv[k].z, x, acc.z); acc.w = fmaf(wv[k].w, x, acc.w);
    }
  }
  *reinterpret_cast<float4*>(red_s + part * N + q * 4) = acc;
  __syncthreads();
  if (tid < N) {
    float s = bias ? bias[tid] : 0.f;   // generic load: the per-pixel chain passes shared-memory copies
    if (extra) s += extra[tid];
#pragma unroll
    for (int p = 0; p < parts; ++p) s += red_s[p * N + tid];
    y_s[tid] = s;
  }
  __syncthreads();
}
// The small GEMV of a layer (conv1x1_2) with its weight slice loaded EARLY: gemv_early_load is issued next to the loads of the
// layer's large GEMV (the weights do not depend on the data), gemv_early_finish consumes the registers after the gate -- the
// second L2 round trip of the layer overlaps the first.  Same slicing and summation order as cta_gemv_t.
template <int N, int K>
struct GemvSlice {
  static constexpr int n4 = N >> 2, parts = 256 / n4, klen = K / parts;
  static_assert(klen >= 1 && klen <= 16, "an early slice lives in registers");
  float4 w[klen];
};
template <int N, int K>
__device__ __forceinline__ void gemv_early_load(const float* __restrict__ Wt, GemvSlice<N, K>& sl) {
  using S = GemvSlice<N, K>;
  const int tid = threadIdx.x;
  const int part = tid / S::n4, q = tid % S::n4;
  const float4* wp = reinterpret_cast<const float4*>(Wt + (size_t)(part * S::klen) * N) + q;
#pragma unroll
  for (int k = 0; k < S::klen; ++k) sl.w[k] = __ldg(wp + k * S::n4);
}
template <int N, int K>
__device__ __forceinline__ void gemv_early_finish(const GemvSlice<N, K>& sl, const float* __restrict__ bias,
                                                  const float* __restrict__ extra, const float* x_s, float* red_s, float* y_s) {
  using S = GemvSlice<N, K>;
  const int tid = threadIdx.x;
  const int part = tid / S::n4, q = tid % S::n4;
  const int kb = part * S::klen;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int k = 0; k < S::klen; ++k) {
    const float x = x_s[kb + k];
    acc.x = fmaf(sl.w[k].x, x, acc.x); acc.y = fmaf(sl.w[k].y, x, acc.y);
    acc.z = fmaf(sl.w[k].z, x, acc.z); acc.w = fmaf(sl.w[k].w, x, acc.w);
  }
  *reinterpret_cast<float4*>(red_s + part * N + q * 4) = acc;
  __syncthreads();
  if (tid < N) {
    float s = bias ? bias[tid] : 0.f;
    if (extra) s += extra[tid];
#pragma unroll
    for (int p = 0; p < S::parts; ++p) s += red_s[p * N + tid];
    y_s[tid] = s;
  }
  __syncthreads();
}

// out[w][n] = bias[n] + sum_{taps} sum_ci Wt[(tap*Kc + ci)][n] * in(tap, w)[ci]  for all w < W, n < N.
// `sm_base + row_off[t]` is column 0 of the smem row (ci contiguous, Kc floats per column) tap t reads (row_off < 0: zero row); `shift[t]` is
// the column offset of tap t (|shift| <= PADW: the staged rows carry PADW zero columns on either side, so the inner loop has
// no bounds test).  Register tile: a thread owns TWO adjacent output channels and PB = 7 positions (N / 2 thread columns x
// 512 / N position groups); per 4 input channels that is 4 64-bit weight loads + 7 128-bit broadcast loads for 56 FMAs.
// (The first version -- one channel x 14 positions, three bounds tests and a 64-bit address product per load -- ran the row
// pass at 12 % of the SM's FMA rate and was 51 % of the sampler; tools/pixelcnn_phases.py.)  Every output is still ONE
// accumulator walking (tap, ci) in ascending order: bit-identical to the scalar form.
// N and Kc are compile-time (N = 2 * hidden_dim; Kc = hidden_dim for vert_conv, 2 * hidden_dim for the v -> h link): with
// run-time strides ptxas spent ~60 integer instructions per 112 FMAs on 64-bit address products (SASS of the first tiled
// version), and the weight loads of only two k-chunks were in flight.
template <int N, int KC>
__device__ __forceinline__ void cta_row_gemm_t(const float* __restrict__ Wt, const float* __restrict__ bias,
                                               const float* sm_base, const int* row_off, const int* shift, int ntaps, int W,
                                               float* out_s /*[W][N]*/) {
  const int tid = threadIdx.x;
  constexpr int ncol = N >> 1;             // thread columns
  constexpr int groups = 256 / ncol;
  const int tc = tid % ncol, grp = tid / ncol;
  const int n0 = tc * 2;
  const float2 b2 = bias ? __ldg(reinterpret_cast<const float2*>(bias + n0)) : make_float2(0.f, 0.f);
  for (int w0 = grp * PB; w0 < W; w0 += groups * PB) {
    float2 acc[PB];
#pragma unroll
    for (int p = 0; p < PB; ++p) acc[p] = b2;
    for (int t = 0; t < ntaps; ++t) {
      if (row_off[t] < 0) continue;
      // offset from the kernel's shared-memory array, not a pointer out of a local array: a pointer of unknown address space
      // made ptxas emit generic LD for the operand rows (ncu source page of the previous version) instead of LDS
      const float* row = sm_base + row_off[t];
      const float* xp = row + (w0 + shift[t]) * KC;
      const float2* wp = reinterpret_cast<const float2*>(Wt + (size_t)t * KC * N + n0);
      // 16-row weight blocks, double-buffered in registers: the next block's loads are issued before the current block's
      // 448 FMAs (with the loads at the top of each block every block waited out an L2 round trip: 32 per layer)
      auto load16 = [&](float2 (&wv)[16], const float2* src) {
#pragma unroll
        for (int k = 0; k < 16; ++k) wv[k] = __ldg(src + k * (N / 2));
      };
      auto fma16 = [&](const float2 (&wv)[16], const float* xq) {
#pragma unroll
        for (int c4 = 0; c4 < 16; c4 += 4) {
#pragma unroll
          for (int p = 0; p < PB; ++p) {
            const float4 x = *reinterpret_cast<const float4*>(xq + p * KC + c4);
            acc[p].x = fmaf(wv[c4 + 0].x, x.x, acc[p].x); acc[p].y = fmaf(wv[c4 + 0].y, x.x, acc[p].y);
            acc[p].x = fmaf(wv[c4 + 1].x, x.y, acc[p].x); acc[p].y = fmaf(wv[c4 + 1].y, x.y, acc[p].y);
            acc[p].x = fmaf(wv[c4 + 2].x, x.z, acc[p].x); acc[p].y = fmaf(wv[c4 + 2].y, x.z, acc[p].y);
            acc[p].x = fmaf(wv[c4 + 3].x, x.w, acc[p].x); acc[p].y = fmaf(wv[c4 + 3].y, x.w, acc[p].y);
          }
        }
      };
      static_assert(KC % 32 == 0, "two 16-row blocks per iteration");
      float2 wA[16], wB[16];
      load16(wA, wp);
#pragma unroll 1
      for (int c32 = 0; c32 < KC; c32 += 32, wp += 32 * (N / 2), xp += 32) {
        load16(wB, wp + 16 * (N / 2));
        fma16(wA, xp);
        if (c32 + 32 < KC) load16(wA, wp + 32 * (N / 2));
        fma16(wB, xp + 16);
      }
    }
#pragma unroll
    for (int p = 0; p < PB; ++p)
      if (w0 + p < W) *reinterpret_cast<float2*>(out_s + (w0 + p) * N + n0) = acc[p];
  }
  __syncthreads();
}

__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t a, uint32_t b) {
  // Philox4x32-10 keyed by seed, counter (a, b, 0, 0) -> one uniform in [0, 1)
  uint32_t c0 = a, c1 = b, c2 = 0, c3 = 0, k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

__device__ unsigned long long g_pcnn_prof[8];   // + [4..7]: chain split (fill, horiz gemv, gate, conv1x1_2 + tail)   // row pass | per-pixel chain | head + draw | pixels   (IGM_PCNN_PROF=1)

// HD = hidden_dim as a template parameter: every index split of the kernel (staging loops, gates, slices) divides by a
// power-of-two constant instead of a run-time value.
template <int HD>
__global__ void __launch_bounds__(256, 1) pixelcnn_kernel(const PcnnArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int n_img = blockIdx.x;
  const int tid = threadIdx.x;
  constexpr int Hd = HD, N2 = 2 * HD;
  const int C = a.C, H = a.H, W = a.W;
  // shared memory carve-up
  float* in_s = sm;                         // [2][W + ROWX][Hd]  two input rows of the vertical conv, PADW zero columns either side
  float* vc_s = in_s + 2 * (W + ROWX) * Hd; // [W][2Hd]     vert_conv output of the current layer (+ PB rows of slack via t_s)
  float* t_s = vc_s + W * N2;               // [W][2Hd]     scratch row (v2h result / gated row)
  float* x_s = t_s + W * N2;                // [2Hd]        gemv input
  float* y_s = x_s + N2;                    // [256]        gemv output
  float* red_s = y_s + 256;                 // [1024]       K-slice partials of a gemv
  float* cur_s = red_s + 1024;              // [Hd]         running h-stack column
  float* lg_s = cur_s + Hd;                 // [256]        logits of one channel
  // Small operands of the per-pixel chain, staged so that no global load sits on its critical path (each one was an
  // exposed L2 round trip between two barriers: ~3 per layer, 33 per pixel): constants once per kernel, the rest per pixel
  float* hb_s = lg_s + 256;                 // [11][2Hd]    horiz_conv biases
  float* h2b_s = hb_s + NLAYERS * N2;       // [11][Hd]     conv1x1_2 biases
  float* cond_s = h2b_s + NLAYERS * Hd;     // [11][2Hd]    class-conditioning addends of the horizontal gates (zeros without)
  float* v2h_s = cond_s + NLAYERS * N2;     // [11][2Hd]    conv1x1_1(vert_conv) at (h, w) of every layer
  float* hsp_s = v2h_s + NLAYERS * N2;      // [11][Hd]     horizontal features at (h, w - dilation) of every layer
  __shared__ int s_pick;

  float* ws = a.ws + (int64_t)n_img * a.ws_per_img;
  float* Vc = ws;                                        // [12][H][W][Hd]   vertical-stack features
  float* V2H = Vc + (int64_t)12 * H * W * Hd;            // [11][W][2Hd]     conv1x1_1(vert_conv) of row h
  float* Hs = V2H + (int64_t)NLAYERS * W * N2;           // [12][W][Hd]      horizontal-stack features of row h
  float* img = a.img + (int64_t)n_img * C * H * W;
  const float* Wt = a.Wt;

  for (int i = tid; i < 2 * (W + ROWX) * Hd; i += 256) in_s[i] = 0.f;   // the padding columns stay zero for the whole kernel
  for (int i = tid; i < NLAYERS * N2; i += 256) {
    const int l = i / N2, j = i - l * N2;
    hb_s[i] = __ldg(Wt + a.off.horiz_b[l] + j);
    cond_s[i] = a.cond ? __ldg(a.cond + ((int64_t)(l * 2 + 1) * a.N + blockIdx.x) * N2 + j) : 0.f;
  }
  for (int i = tid; i < NLAYERS * Hd; i += 256) {
    const int l = i / Hd;
    h2b_s[i] = __ldg(Wt + a.off.h2_b[l] + (i - l * Hd));
  }
  __syncthreads();

  for (int h = 0; h < H; ++h) {
    long long t_row = 0;
    if (a.prof && tid == 0) t_row = clock64();
    // ================= row pass: vertical stack =================
    // layer 0: conv_vstack 5x5, rows ky = 0,1 live (image rows h-2, h-1), pad 2   (pixelcnn.py:98-100)
    for (int i = tid; i < W * Hd; i += 256) {
      const int w = i / Hd, co = i - w * Hd;
      float s = __ldg(Wt + a.off.vs0_b + co);
      for (int ky = 0; ky < 2; ++ky) {
        const int r = h - 2 + ky;
        if (r < 0) continue;
        for (int kx = 0; kx < 5; ++kx) {
          const int col = w - 2 + kx;
          if (col < 0 || col >= W) continue;
          for (int ci = 0; ci < C; ++ci)
            s = fmaf(__ldg(Wt + a.off.vs0_w + (int64_t)((ky * 5 + kx) * C + ci) * Hd + co),
                     img[((int64_t)ci * H + r) * W + col], s);
        }
      }
      Vc[(((int64_t)0 * H + h) * W + w) * Hd + co] = s;
    }
    __syncthreads();
    for (int l = 0; l < NLAYERS; ++l) {
      const int d = c_dil[l];
      // stage rows h-d and h of the previous layer's vertical features
      const float* prev = Vc + ((int64_t)l * H) * W * Hd;
      const bool top_ok = (h - d) >= 0;
      float* row0 = in_s + PADW * Hd;                        // column 0 of the staged row h - d
      float* row1 = in_s + (W + ROWX) * Hd + PADW * Hd;      // column 0 of the staged row h
      for (int i = tid; i < W * Hd / 4; i += 256) {
        reinterpret_cast<float4*>(row1)[i] = reinterpret_cast<const float4*>(prev + (int64_t)h * W * Hd)[i];
        if (top_ok) reinterpret_cast<float4*>(row0)[i] = reinterpret_cast<const float4*>(prev + (int64_t)(h - d) * W * Hd)[i];
      }
      __syncthreads();
      // vert_conv 3x3 dilated, rows ky = 0 (h-d), 1 (h); cols w-d, w, w+d      (pixelcnn.py:51-53)
      int rows[6];
      int shift[6];
      for (int ky = 0; ky < 2; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          rows[ky * 3 + kx] = (ky == 0) ? (top_ok ? (int)(row0 - sm) : -1) : (int)(row1 - sm);
          shift[ky * 3 + kx] = (kx - 1) * d;
        }
      cta_row_gemm_t<N2, Hd>(Wt + a.off.vert_w[l], Wt + a.off.vert_b[l], sm, rows, shift, 6, W, vc_s);
      // gated vertical output: tanh(a) * sigmoid(b)                            (pixelcnn.py:69)
      float* vout = Vc + (((int64_t)(l + 1) * H + h) * W) * Hd;
      for (int i = tid; i < W * Hd; i += 256) {
        const int w = i / Hd, c = i - w * Hd;
        float av = vc_s[w * N2 + c], bv = vc_s[w * N2 + Hd + c];
        if (a.cond) {
          const float* cp = a.cond + ((int64_t)(l * 2 + 0) * a.N + blockIdx.x) * N2;
          av = __fadd_rn(av, __ldg(cp + c));
          bv = __fadd_rn(bv, __ldg(cp + Hd + c));
        }
        vout[i] = tanhf(av) * (1.f / (1.f + expf(-bv)));
      }
      // v -> h link: conv1x1_1 on the PRE-gate features                         (pixelcnn.py:74)
      const int rows1[1] = {(int)(vc_s - sm)};
      const int shift1[1] = {0};
      cta_row_gemm_t<N2, N2>(Wt + a.off.v2h_w[l], Wt + a.off.v2h_b[l], sm, rows1, shift1, 1, W, t_s);
      float* v2h = V2H + (int64_t)l * W * N2;
      for (int i = tid; i < W * N2 / 4; i += 256) reinterpret_cast<float4*>(v2h)[i] = reinterpret_cast<const float4*>(t_s)[i];
      __syncthreads();
    }

    if (a.prof && tid == 0) atomicAdd(&g_pcnn_prof[0], (unsigned long long)(clock64() - t_row));
    // ================= pixel pass: horizontal stack, one column at a time =================
    for (int w = 0; w < W; ++w) {
      long long t_px = 0;
      if (a.prof && tid == 0) t_px = clock64();
      // layer 0: conv_hstack 1x5, cols kx = 0,1 live (w-2, w-1), pad 2           (pixelcnn.py:101-103)
      if (tid < Hd) {
        float s = __ldg(Wt + a.off.hs0_b + tid);
        for (int kx = 0; kx < 2; ++kx) {
          const int col = w - 2 + kx;
          if (col < 0) continue;
          for (int ci = 0; ci < C; ++ci)
            s = fmaf(__ldg(Wt + a.off.hs0_w + (int64_t)(kx * C + ci) * Hd + tid), img[((int64_t)ci * H + h) * W + col], s);
        }
        cur_s[tid] = s;
        Hs[((int64_t)0 * W + w) * Hd + tid] = s;
      }
      // this pixel's v->h link rows and the features one dilation to the left, for all layers at once (plain loads: both were
      // written by this CTA, V2H in the row pass, Hs at earlier pixels of this row -- dilations are >= 1)
      for (int i = tid; i < NLAYERS * N2 / 4; i += 256) {
        const int l = i / (N2 / 4), j4 = i - l * (N2 / 4);
        reinterpret_cast<float4*>(v2h_s)[i] = reinterpret_cast<const float4*>(V2H + ((int64_t)l * W + w) * N2)[j4];
      }
      for (int i = tid; i < NLAYERS * Hd / 4; i += 256) {
        const int l = i / (Hd / 4), j4 = i - l * (Hd / 4);
        const int col = w - c_dil[l];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col >= 0) v = reinterpret_cast<const float4*>(Hs + ((int64_t)l * W + col) * Hd)[j4];
        reinterpret_cast<float4*>(hsp_s)[i] = v;
      }
      __syncthreads();
      // weights of the chain one matrix ahead (registers): possible while a slice is <= 16 / <= 4 float4 (hidden_dim <= 64)
      for (int l = 0; l < NLAYERS; ++l) {
        const int d = c_dil[l];
        // horiz_conv 1x3 dilated, cols kx = 0 (w-d), 1 (w)                        (pixelcnn.py:48-50)
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
        if (a.prof && tid == 0) c0 = clock64();
        if (tid < Hd) {
          x_s[Hd + tid] = cur_s[tid];
          x_s[tid] = hsp_s[l * Hd + tid];
        }
        __syncthreads();
        if (a.prof && tid == 0) c1 = clock64();
        GemvSlice<Hd, Hd> h2sl;
        if (a.early) gemv_early_load<Hd, Hd>(Wt + a.off.h2_w[l], h2sl);   // conv1x1_2's slice travels with horiz_conv's
        cta_gemv_t<N2, N2>(Wt + a.off.horiz_w[l], hb_s + l * N2, v2h_s + l * N2, x_s, red_s, y_s);
        if (a.prof && tid == 0) c2 = clock64();
        // gated horizontal output: tanh(a) * tanh(b)  (sic)                      (pixelcnn.py:77)
        if (tid < Hd) {
          float ah = y_s[tid], bh = y_s[Hd + tid];
          if (a.cond) {
            ah = __fadd_rn(ah, cond_s[l * N2 + tid]);
            bh = __fadd_rn(bh, cond_s[l * N2 + Hd + tid]);
          }
          x_s[tid] = tanhf(ah) * tanhf(bh);
        }
        __syncthreads();
        if (a.prof && tid == 0) c3 = clock64();
        // conv1x1_2 + residual                                                    (pixelcnn.py:80)
        if (a.early) gemv_early_finish<Hd, Hd>(h2sl, h2b_s + l * Hd, cur_s, x_s, red_s, y_s);
        else cta_gemv_t<Hd, Hd>(Wt + a.off.h2_w[l], h2b_s + l * Hd, cur_s, x_s, red_s, y_s);
        if (tid < Hd) {
          cur_s[tid] = y_s[tid];
          Hs[((int64_t)(l + 1) * W + w) * Hd + tid] = y_s[tid];
        }
        __syncthreads();
        if (a.prof && tid == 0) {
          const long long c4 = clock64();
          atomicAdd(&g_pcnn_prof[4], (unsigned long long)(c1 - c0));
          atomicAdd(&g_pcnn_prof[5], (unsigned long long)(c2 - c1));
          atomicAdd(&g_pcnn_prof[6], (unsigned long long)(c3 - c2));
          atomicAdd(&g_pcnn_prof[7], (unsigned long long)(c4 - c3));
        }
      }
      if (a.prof && tid == 0) {
        const long long t = clock64();
        atomicAdd(&g_pcnn_prof[1], (unsigned long long)(t - t_px));
        atomicAdd(&g_pcnn_prof[3], 1ull);
        t_px = t;
      }
      // head: conv_out(elu(h_stack))                                              (pixelcnn.py:148)
      if (tid < Hd) {
        const float v = cur_s[tid];
        x_s[tid] = v > 0.f ? v : expm1f(v);
      }
      __syncthreads();
      const bool keep = a.mode == 2 || (a.skip && a.skip[h * W + w]);
      for (int ch = 0; ch < C; ++ch) {
        // logits of channel ch: conv_out channel o = cls*C + ch  (reshape at pixelcnn.py:151-153)
        {
          float s = __ldg(Wt + a.off.out_b + (int64_t)tid * C + ch);
          const float* wp = Wt + a.off.out_w + (int64_t)tid * C + ch;
          // all weights of a 64-row block in flight before the first FMA (the 16-way unrolled form made four dependent L2
          // round trips per logit); same summation order
          constexpr int HB = Hd < 64 ? Hd : 64;
#pragma unroll 1
          for (int k0 = 0; k0 < Hd; k0 += HB) {
            float wv[HB];
#pragma unroll
            for (int k = 0; k < HB; ++k) wv[k] = __ldg(wp + (int64_t)(k0 + k) * 256 * C);
#pragma unroll
            for (int k = 0; k < HB; ++k) s = fmaf(wv[k], x_s[k0 + k], s);
          }
          lg_s[tid] = s;
          if (a.logits) a.logits[((((int64_t)n_img * 256 + tid) * C + ch) * H + h) * W + w] = s;
        }
        __syncthreads();
        if (!keep && tid < 32) {
          // softmax over 256 classes + draw, one warp; lane owns classes [8*lane, 8*lane+8)
          float v[8];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = lg_s[tid * 8 + j]; mx = fmaxf(mx, v[j]); }
          mx = warp_max(mx);
          int pickk;
          if (a.mode == 1) {
            // greedy: first index of the maximum
            int best = 1 << 30;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (v[j] == mx) best = min(best, tid * 8 + j);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            pickk = best;
          } else {
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = expf(v[j] - mx); sum += v[j]; }
            const float tot = warp_sum(sum);
            // inclusive scan of lane sums -> exclusive offset of this lane
            float run = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const float t = __shfl_up_sync(0xffffffffu, run, o);
              if (tid >= o) run += t;
            }
            float cdf = run - sum;
            float u;
            if (a.uniforms) u = __ldg(a.uniforms + (int64_t)(h * W + w) * a.N * C + (int64_t)n_img * C + ch);
            else u = philox_uniform(a.seed, (uint32_t)(h * W + w), (uint32_t)(n_img * C + ch));
            const float inv = 1.f / tot;
            int cnt = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              cdf += v[j];
              cnt += (cdf * inv <= u) ? 1 : 0;   // k = #{j : cdf_j <= u}
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            pickk = min(cnt, 255);
          }
          if (tid == 0) s_pick = pickk;
        }
        __syncthreads();
        if (!keep && tid == 0) {
          float val = (float)s_pick / 255.f;
          if (a.normalize) val = val * 2.f - 1.f;
          img[((int64_t)ch * H + h) * W + w] = val;
        }
        __syncthreads();
      }
      if (a.prof && tid == 0) atomicAdd(&g_pcnn_prof[2], (unsigned long long)(clock64() - t_px));
    }
  }
}

}  // namespace
}  // namespace igm

using namespace igm;

extern "C" int64_t igm_pixelcnn_weight_floats(int C, int Hd) { return make_offsets(C, Hd).total; }

extern "C" int64_t igm_pixelcnn_workspace_floats(int N, int C, int H, int W, int Hd) {
  (void)C;
  const int64_t per = (int64_t)12 * H * W * Hd + (int64_t)NLAYERS * W * 2 * Hd + (int64_t)12 * W * Hd;
  return per * N;
}

// mode 0: inverse-CDF draws (uniforms [H*W][N*C] or Philox(seed)); 1: greedy; 2: teacher forced.
// cond (nullable): [11][2][N][2*Hd] per-image conditioning addends (cond_proj_* outputs, vert then horiz).
extern "C" int igm_pixelcnn_run(const float* weights, float* img, const float* uniforms, const uint8_t* skip,
                                const float* cond, float* logits, float* ws, uint64_t seed, int N, int C, int H, int W, int Hd,
                                int mode, int normalize, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!weights || !img || !ws) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  if (Hd != 32 && Hd != 64 && Hd != 128) IGM_FAIL(st, IGM_ERR_INVALID, "hidden_dim must be 32, 64 or 128");
  if (W < 1 || W > MAX_W || H < 1 || N < 1 || C < 1 || C > 4) IGM_FAIL(st, IGM_ERR_INVALID, "unsupported image shape");
  if (mode < 0 || mode > 2) IGM_FAIL(st, IGM_ERR_INVALID, "mode must be 0, 1 or 2");
  PcnnArgs a;
  a.Wt = weights; a.off = make_offsets(C, Hd);
  a.img = img; a.uniforms = uniforms; a.skip = skip; a.cond = cond; a.logits = logits; a.ws = ws;
  a.ws_per_img = igm_pixelcnn_workspace_floats(1, C, H, W, Hd);
  a.seed = seed; a.N = N; a.C = C; a.H = H; a.W = W; a.Hd = Hd; a.mode = mode; a.normalize = normalize;
  const size_t smem = sizeof(float) * ((size_t)2 * (W + ROWX) * Hd + (size_t)2 * W * 2 * Hd + 2 * Hd + 256 + 1024 + Hd + 256 +
                                       (size_t)NLAYERS * (3 * 2 * Hd + 2 * Hd));
  static const bool early_on = [] { const char* e = getenv("IGM_PCNN_EARLY"); return !(e && e[0] == '0'); }();
  a.early = (early_on && Hd <= 64) ? 1 : 0;   // hidden_dim 128: 16 float4 per thread would spill
  static const bool prof_on = [] { const char* e = getenv("IGM_PCNN_PROF"); return e && e[0] == '1'; }();
  a.prof = prof_on ? 1 : 0;
  if (smem > 227 * 1024) IGM_FAIL(st, IGM_ERR_INVALID, "PixelCNN: image row too wide for this hidden_dim (shared memory)");
  auto go = [&](auto kernel) -> cudaError_t {
    cudaError_t er = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (er != cudaSuccess) return er;
    kernel<<<N, 256, smem, (cudaStream_t)stream>>>(a);
    return cudaSuccess;
  };
  cudaError_t e = Hd == 64 ? go(pixelcnn_kernel<64>) : Hd == 32 ? go(pixelcnn_kernel<32>) : go(pixelcnn_kernel<128>);
  if (e != cudaSuccess) IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e));
  ++ops_launch_counter();
  e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// IGM_PCNN_PROF=1: cycles summed over all CTAs since the last call -- out[0] row pass, out[1] per-pixel chain, out[2] head +
// draw, out[3] pixels, out[4..7] the chain split into x fill / horiz GEMV / gate / conv1x1_2 + tail -- and reset.  Diagnosis only (tools/pixelcnn_phases.py).
extern "C" int igm_debug_pixelcnn_prof(unsigned long long* out) {
  unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (cudaMemcpyFromSymbol(out, g_pcnn_prof, sizeof(z)) != cudaSuccess) return IGM_ERR_CUDA;
  if (cudaMemcpyToSymbol(g_pcnn_prof, z, sizeof(z)) != cudaSuccess) return IGM_ERR_CUDA;
  return IGM_OK;
}
