// tcgen05 (5th-gen tensor core) convolution engine: declarations shared with unet.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace igm {

// One stride-1 convolution (or its data gradient) lowered to an implicit GEMM
//   D[m, n] = sum_{tap, k} A[pix(m) + tap, k] * Wt[n, tap*K + k]
// with every operand split into bf16 hi + lo parts ("bf16x3": hi*hi + hi*lo + lo*hi,
// fp32 accumulation in TMEM) so the result is fp32-accurate to ~2^-17 relative.
// One (tap) term of the implicit GEMM: the activation box is fetched at tile origin + (dx, dy) on the
// source grid, from sub-lattice (px, py) when the source is read with stride 2 ("space-to-depth"
// view), and multiplied by column block `wtap` of the packed weight matrix.
struct TcTap {
  int dx, dy, px, py, wtap;
};
constexpr int kTcMaxTaps = 16;

struct TcConv {
  bool valid = false;
  // activations may come from two tensors concatenated along channels (skip connections):
  // channels [0, K0) from source 0, [K0, K) from source 1 (a1_* = a0_* when there is one source)
  alignas(64) CUtensorMap a_hi, a_lo, a1_hi, a1_lo, b_hi, b_lo;
  int K0 = 0;
  __nv_bfloat16* w_hi = nullptr;   // [N][taps*K]
  __nv_bfloat16* w_lo = nullptr;
  int K = 0, N = 0, KH = 1, KW = 1, pad = 0;
  int H = 0, W = 0, Bmax = 0;           // tile grid (= output grid of one launch) per image
  int BH = 0, BW = 0, BB = 0, BN = 0;   // M-tile = BB images x BH rows x BW cols (<= 128 pixels)
  int ntaps = 0;
  TcTap taps[kTcMaxTaps];
  int Csrc = 0;                         // channels of source 0 (sub-lattice px selects channel block px*Csrc)
  // output pixel of tile position (oy, ox):  (oy*sy + oy_off, ox*sx + ox_off) on an out_H x out_W image
  int out_H = 0, out_W = 0, sy = 1, sx = 1, oy_off = 0, ox_off = 0;
  // all four output-parity phases in ONE launch (tc_plan_phases4): per-phase tap ranges and output offsets
  int w_img_rows = 0;   // > 0: per-image weight matrices (tc_plan_img)
  int nph = 1, ph_tap0[4] = {0, 0, 0, 0}, ph_ntaps[4] = {0, 0, 0, 0}, ph_oy[4] = {0, 0, 0, 0}, ph_ox[4] = {0, 0, 0, 0};
  // TMA-store epilogue: output tensor maps, (re-)encoded when a launch passes a new output pointer
  struct OutMaps {
    const void* p0 = nullptr; const void* p1 = nullptr; const void* ph = nullptr; const void* pl = nullptr;
    int n0 = 0, n1 = 0, nh = 0, ldh = 0;
    alignas(64) CUtensorMap m0, m1, mh, ml;
  };
  mutable OutMaps om;
};

// Stride-2 convolution read (Conv2d k3 s2 p1 forward / ConvTranspose2d k4 s2 p1 data gradient):
// source [Bmax, SH, SW, K] sampled at (2*oy - pad + ky, 2*ox - pad + kx); output grid SH/2 x SW/2.
bool tc_strided_eligible(int K, int N, int SH, int SW, int KH);
int tc_plan_strided(Status& st, TcConv& t, int K, int N, int SH, int SW, int Bmax, int KH, int pad,
                    __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo);
// One output-parity phase (py, px) of a stride-2 transposed convolution (ConvTranspose2d k4 s2 p1
// forward / Conv2d k3 s2 p1 data gradient): source grid GH x GW at unit stride, output written at
// (2*a + py, 2*b + px) of a 2GH x 2GW image; only the taps ky with (py + pad - ky) even contribute.
int tc_plan_phase(Status& st, TcConv& t, int K, int N, int GH, int GW, int Bmax, int KH, int pad, int py, int px,
                  __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo);

// the four phases (py, px) of tc_plan_phase as one plan / one launch (tiles of the four phases interleave on the SMs)
int tc_plan_phases4(Status& st, TcConv& t, int K, int N, int GH, int GW, int Bmax, int KH, int pad, __nv_bfloat16* a_hi,
                    __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo);

// 1x1 convolution whose weights differ per image: weight tensor [Bmax * N rows][K] (image b owns rows b*N .. b*N+N-1)
// a_pitch > K: the K input channels are a slice of a wider tensor (a_hi / a_lo already point at the slice's first channel)
int tc_plan_img(Status& st, TcConv& t, int K, int N, int H, int W, int Bmax, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, int a_pitch = 0);

// Can this (stride-1, non-dilated) conv run on the tensor-core engine?
bool tc_eligible(int K, int N, int H, int W, int KH);

// Fills `t` (tile shape + the TMA descriptors).  Source 0: bf16 hi/lo [Bmax, H, W, K0]; optional
// source 1: [Bmax, H, W, K - K0] (pass K0 = K and null pointers for a single source).
int tc_plan(Status& st, TcConv& t, int K, int N, int H, int W, int Bmax, int KH, int pad,
            __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo,
            int K0 = 0, __nv_bfloat16* a1_hi = nullptr, __nv_bfloat16* a1_lo = nullptr);

struct TcRun {
  int B = 0;
  const float* bias = nullptr;
  float* out0 = nullptr; float* out1 = nullptr; int N0 = 0;
  const float* add0 = nullptr; const float* add1 = nullptr;
  __nv_bfloat16* hi0 = nullptr;   // optional: also emit out0 as bf16 hi/lo (operand staging for the
  __nv_bfloat16* lo0 = nullptr;   // next tensor-core conv), same [M, N0] layout; out0 may then be null ("lean")
  int ld_hi = 0;                  // > N0: hi0 / lo0 are a channel slice of a wider [M, ld_hi] tensor (pointers at the slice)
  // optional fused GroupNorm(8) partial statistics of the output (sum, sum of squares per image,
  // 32-pixel slot and group), layout part[b][slot][8][2]; see tc_gn_fusable()
  float* gn_part = nullptr;
  int kclass = K_CONV_FPROP;
};
// Can the epilogue of plan `t` produce the GroupNorm partials (full 128-pixel tiles, warps inside one image)?
bool tc_gn_fusable(const TcConv& t, int B);
inline int tc_gn_slots(const TcConv& t) { return t.H * t.W / 32; }
int launch_conv_tc(const LaunchCtx& lc, const TcConv& t, const TcRun& r);

// CTA-pair (`cta_group::2`, M = 256) variant of the per-tap engine (conv_tc2.cu): same activation boxes and packed weights
// as `base`, each CTA of a pair stages its own M tile and half of the weight tile.  Parity-green on B200 but slower than
// the per-tap engine (profiles/r2_conv_engines.md): reachable through igm_debug_conv / igm_debug_conv_bench (engine = 3)
// and IGM_CONV_PAIR=1 only.  `base` must outlive the pair plan.
struct TcConvPair {
  bool valid = false;
  const TcConv* base = nullptr;
  alignas(64) CUtensorMap b_hi, b_lo;   // weight boxes of 64 channels x 64 rows
  mutable TcConv::OutMaps om;
};
bool tc2_eligible(const TcConv& base);
int tc2_plan(Status& st, TcConvPair& t, const TcConv& base);
int launch_conv_tc2(const LaunchCtx& lc, const TcConvPair& t, const TcRun& r);

// Weight gradients on the tensor cores (wgrad_tc.cu).  Pixels are the reduction axis:
//   G[tap][cs][cp] = sum_pixels S[pixel shifted by tap][cs] * P[pixel][cp]
// S ("shifted") is fetched with the per-tap offsets / stride-2 sub-lattices of TcTap, P ("plain") is
// walked pixel tile by pixel tile (64-pixel TMA boxes, optionally a 2-tensor channel concat).
// The result is ACCUMULATED at grad[cs*s_shift + cp*s_plain + wtap] (PyTorch OIHW / IOHW directly).
struct TcWgrad {
  bool valid = false;
  alignas(64) CUtensorMap s_hi, s_lo, p_hi, p_lo, p1_hi, p1_lo;
  int CS = 0, CP = 0, P0 = 0;          // channels of S, of P, of P's first tensor
  int GH = 0, GW = 0, Bmax = 0;        // pixel grid of P per image
  int BW = 0, BH = 0, BB = 0, rows = 0, BN = 0;
  int ntaps = 0;
  TcTap taps[kTcMaxTaps];
  int64_t s_shift = 0, s_plain = 0;
  double flops_per_image = 0;
};
bool tcw_eligible(int Cin, int Cout, int H, int W, int KH);
// stride-1 KxK conv: S = dY [Bmax,H,W,Cout] (staging), P = X [Bmax,H,W,Cin] (optionally 2 tensors)
int tcw_plan(Status& st, TcWgrad& t, int Cin, int Cout, int H, int W, int Bmax, int KH, int pad,
             __nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo, __nv_bfloat16* x_hi, __nv_bfloat16* x_lo,
             int C0 = 0, __nv_bfloat16* x1_hi = nullptr, __nv_bfloat16* x1_lo = nullptr);
// stride-2 gather (Conv2d k3 s2 p1: S = X, P = dY, grad OIHW; ConvTranspose2d k4 s2 p1: S = dY, P = X,
// grad IOHW).  S lives on the fine grid [Bmax, 2GH, 2GW, CS], P on the coarse grid [Bmax, GH, GW, CP].
bool tcw_strided_eligible(int CS, int CP, int GH, int GW, int KH);
int tcw_plan_strided(Status& st, TcWgrad& t, int CS, int CP, int GH, int GW, int Bmax, int KH, int pad,
                     __nv_bfloat16* s_hi, __nv_bfloat16* s_lo, __nv_bfloat16* p_hi, __nv_bfloat16* p_lo,
                     int64_t s_shift, int64_t s_plain);
bool tcw_batch_ok(const TcWgrad& t, int B);
// grad (fp32) += G; both operands must already be staged as bf16 hi/lo in the planned buffers
int launch_wgrad_tc(const LaunchCtx& lc, const TcWgrad& t, int B, float* grad, int variant = 0);

// Halo-reuse weight gradients of the stride-1 3x3 convs (wgrad_halo.cu): one zero-padded dY tile per pixel
// tile feeds all nine taps through shifted shared-memory descriptors; partial results are reduced into the
// per-layer workspace ws[ky][kx][ci][co] and folded into the OIHW gradient by launch_wgrad_halo_finalize.
struct TcWgradHalo {
  bool valid = false;
  alignas(64) CUtensorMap s_hi, s_lo, p_hi, p_lo, p1_hi, p1_lo;
  int Cin = 0, Cout = 0, P0 = 0, H = 0, W = 0, Bmax = 0;
  bool stacked = false;               // Cin == 64: the hi and lo copies of X form one M = 128 operand
  int BH = 0, stages = 0;
  int prows = 0, prows_pad = 0, p_blk_bytes = 0, s_box_bytes = 0, s_off = 0, stage_bytes = 0;
  float* ws = nullptr;
};
struct HaloFinJob {
  float* ws; float* grad; int Cin, Cout, tile_begin;
};
bool tcwh_eligible(int Cin, int Cout, int H, int W, int KH);
int tcwh_plan(Status& st, TcWgradHalo& t, int Cin, int Cout, int H, int W, int Bmax, __nv_bfloat16* dy_hi,
              __nv_bfloat16* dy_lo, __nv_bfloat16* x_hi, __nv_bfloat16* x_lo, int C0, __nv_bfloat16* x1_hi,
              __nv_bfloat16* x1_lo, float* ws);
int launch_wgrad_halo(const LaunchCtx& lc, const TcWgradHalo& t, int B);
// tile_begin = first CTA of the layer; d_cta_job[cta] = layer index (a layer spans (Cin/32)*(Cout/32) CTAs)
// (n_ctas, cta_base) select the CTA range: all layers (0 .. total) or one layer (its tile_begin, its CTA count)
int launch_wgrad_halo_finalize(const LaunchCtx& lc, const HaloFinJob* d_jobs, const int* d_cta_job, int n_ctas, double elems,
                               int cta_base = 0);

// fp32 [M, C] -> bf16 hi / lo written at channel offset `coff` of rows with `cdst` channels
int launch_split_bf16(const LaunchCtx& lc, const float* src, int64_t M, int C, __nv_bfloat16* hi,
                      __nv_bfloat16* lo, int cdst, int coff);

// dst = hi + lo (debug taps of tensors whose fp32 copy was skipped)
int launch_merge_bf16(const LaunchCtx& lc, const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* dst, int64_t n);
// the same split (dense rows, cdst = C) fused with colsum[c] += sum_m src[m][c]; C must satisfy split_colsum_ok
bool split_colsum_ok(int C);
int launch_split_bf16_colsum(const LaunchCtx& lc, const float* src, int64_t M, int C, __nv_bfloat16* hi,
                             __nv_bfloat16* lo, float* colsum);

// Wt[n][tap*K + k] (hi, lo) = src[k*sk + n*sn + (flip ? taps-1-tap : tap)]
int launch_pack_weight_tc(const LaunchCtx& lc, const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int taps,
                          int K, int N, int64_t sk, int64_t sn, int flip);

}  // namespace igm
