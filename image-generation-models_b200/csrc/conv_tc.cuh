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
struct TcConv {
  bool valid = false;
  // activations may come from two tensors concatenated along channels (skip connections):
  // channels [0, K0) from source 0, [K0, K) from source 1 (a1_* = a0_* when there is one source)
  alignas(64) CUtensorMap a_hi, a_lo, a1_hi, a1_lo, b_hi, b_lo;
  int K0 = 0;
  __nv_bfloat16* w_hi = nullptr;   // [N][taps*K]
  __nv_bfloat16* w_lo = nullptr;
  int K = 0, N = 0, KH = 1, KW = 1, pad = 0;
  int H = 0, W = 0, Bmax = 0;
  int BH = 0, BW = 0, BB = 0, BN = 0;   // M-tile = BB images x BH rows x BW cols (<= 128 pixels)
};

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
  __nv_bfloat16* lo0 = nullptr;   // next tensor-core conv), same [M, N0] layout
  int kclass = K_CONV_FPROP;
};
int launch_conv_tc(const LaunchCtx& lc, const TcConv& t, const TcRun& r);

// Weight gradient of the same convolutions on the tensor cores (wgrad_tc.cu): pixels are the
// reduction axis, fetched as 64-pixel TMA boxes of the bf16 hi/lo copies of dY and X.
struct TcWgrad {
  bool valid = false;
  alignas(64) CUtensorMap dy_hi, dy_lo, x_hi, x_lo, x1_hi, x1_lo;   // X may be a 2-tensor channel concat
  int C0 = 0;
  int Cin = 0, Cout = 0, KH = 1, KW = 1, pad = 0, H = 0, W = 0, Bmax = 0;
  int BW = 0, BH = 0, BB = 0, rows = 0, BN = 0;
};
bool tcw_eligible(int Cin, int Cout, int H, int W, int KH);
int tcw_plan(Status& st, TcWgrad& t, int Cin, int Cout, int H, int W, int Bmax, int KH, int pad,
             __nv_bfloat16* dy_hi, __nv_bfloat16* dy_lo, __nv_bfloat16* x_hi, __nv_bfloat16* x_lo,
             int C0 = 0, __nv_bfloat16* x1_hi = nullptr, __nv_bfloat16* x1_lo = nullptr);
bool tcw_batch_ok(const TcWgrad& t, int B);
// grad (PyTorch OIHW fp32) += dW; dY / X must already be staged as bf16 hi/lo in the planned buffers
int launch_wgrad_tc(const LaunchCtx& lc, const TcWgrad& t, int B, float* grad, int variant = 0);

// fp32 [M, C] -> bf16 hi / lo written at channel offset `coff` of rows with `cdst` channels
int launch_split_bf16(const LaunchCtx& lc, const float* src, int64_t M, int C, __nv_bfloat16* hi,
                      __nv_bfloat16* lo, int cdst, int coff);

// Wt[n][tap*K + k] (hi, lo) = src[k*sk + n*sn + (flip ? taps-1-tap : tap)]
int launch_pack_weight_tc(const LaunchCtx& lc, const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int taps,
                          int K, int N, int64_t sk, int64_t sn, int flip);

}  // namespace igm
