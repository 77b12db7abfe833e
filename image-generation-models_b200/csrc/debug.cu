// Test-only entry point: one stride-1 convolution (or its data gradient) on either engine,
// so the tcgen05 kernel can be parity-checked in isolation over many shapes.
// engine 0 = fp32 SIMT, 1 = tcgen05 per-tap (conv_tc.cu), 3 = tcgen05 CTA pair (conv_tc2.cu, comparison only).
// (engine 2 was the halo-reuse forward kernel: removed in round 2, it faulted on hardware and only covered W >= 28.)
#include "common.cuh"
#include "conv_tc.cuh"

using namespace igm;

extern "C" int igm_debug_conv(int engine, int mode, const float* x, const float* w_oihw, const float* bias,
                              const float* add, float* out, int B, int H, int W, int Cin, int Cout, int K,
                              void* stream) {
  Status& st = global_status();
  st = Status();
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  if (K != 1 && K != 3) IGM_FAIL(st, IGM_ERR_INVALID, "K must be 1 or 3");
  const int KK = K * K, pad = (K - 1) / 2;
  const int Kc = mode == 0 ? Cin : Cout;   // contraction channels
  const int N = mode == 0 ? Cout : Cin;    // output channels
  const int64_t M = (int64_t)B * H * W;
  const int64_t nw = (int64_t)KK * Cin * Cout;
  // OIHW strides of (contraction channel, output channel)
  const int64_t sk = mode == 0 ? KK : (int64_t)Cin * KK;
  const int64_t sn = mode == 0 ? (int64_t)Cin * KK : KK;
  int rc = IGM_OK;
  if (engine == 0) {
    float* wp = nullptr;
    IGM_CUDA(st, cudaMalloc(&wp, nw * sizeof(float)));
    rc = launch_pack_weight(lc, w_oihw, wp, KK, Kc, N, sk, sn);
    if (rc == IGM_OK) {
      ConvArgs a;
      a.in0 = x; a.C0 = Kc; a.B = B; a.IH = a.OH = H; a.IW = a.OW = W;
      a.N = a.N0 = N; a.KH = a.KW = K; a.stride = 1; a.pad = pad; a.transposed = mode;
      a.w = wp; a.bias = bias; a.out0 = out; a.add0 = add;
      rc = launch_conv(lc, a);
    }
    cudaStreamSynchronize(lc.stream);
    cudaFree(wp);
    return rc;
  }
  if (!tc_eligible(Kc, N, H, W, K)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the tcgen05 engine");
  __nv_bfloat16 *wh = nullptr, *wl = nullptr, *ah = nullptr, *al = nullptr;
  IGM_CUDA(st, cudaMalloc(&wh, nw * 2));
  IGM_CUDA(st, cudaMalloc(&wl, nw * 2));
  IGM_CUDA(st, cudaMalloc(&ah, M * Kc * 2));
  IGM_CUDA(st, cudaMalloc(&al, M * Kc * 2));
  TcConv t;
  if (engine != 1 && engine != 3) {
    set_error(st, IGM_ERR_INVALID, __FILE__, __LINE__, "engine must be 0 (SIMT), 1 (per-tap tcgen05) or 3 (CTA pair)");
    rc = IGM_ERR_INVALID;
  } else {
    rc = tc_plan(st, t, Kc, N, H, W, B, K, pad, ah, al, wh, wl);
  }
  TcConvPair tp;
  if (rc == IGM_OK && engine == 3) rc = tc2_plan(st, tp, t);   // CTA-pair engine (conv_tc2.cu) on the same plan
  if (rc == IGM_OK) rc = launch_pack_weight_tc(lc, w_oihw, wh, wl, KK, Kc, N, sk, sn, mode);
  if (rc == IGM_OK) rc = launch_split_bf16(lc, x, M, Kc, ah, al, Kc, 0);
  if (rc == IGM_OK) {
    TcRun r;
    r.B = B; r.bias = bias; r.out0 = out; r.N0 = N; r.add0 = add;
    rc = engine == 3 ? launch_conv_tc2(lc, tp, r) : launch_conv_tc(lc, t, r);
  }
  cudaError_t e = cudaStreamSynchronize(lc.stream);
  if (rc == IGM_OK && e != cudaSuccess) {
    set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e));
    rc = IGM_ERR_CUDA;
  }
  cudaFree(wh); cudaFree(wl); cudaFree(ah); cudaFree(al);
  return rc;
}

// Kernel-level timing of one stride-1 convolution on the tcgen05 engines: operands are staged and the plan is built
// once, then `iters` launches are timed with CUDA events on `stream` (after `warm` untimed ones).  engine 1 = per-tap
// (conv_tc.cu), 3 = CTA pair (conv_tc2.cu); gn != 0 also produces the fused GroupNorm partial statistics, as the
// Block convs of the U-Net do.  Inputs are synthetic (a fixed pattern); *ms_per_launch receives the average.
extern "C" int igm_debug_conv_bench(int engine, int mode, int B, int H, int W, int Cin, int Cout, int K, int gn, int warm,
                                    int iters, float* ms_per_launch, void* stream) {
  Status& st = global_status();
  st = Status();
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  if (K != 1 && K != 3) IGM_FAIL(st, IGM_ERR_INVALID, "K must be 1 or 3");
  if (engine != 1 && engine != 3) IGM_FAIL(st, IGM_ERR_INVALID, "engine must be 1 (per-tap) or 3 (CTA pair)");
  if (iters < 1 || warm < 0 || !ms_per_launch) IGM_FAIL(st, IGM_ERR_INVALID, "bad iteration counts");
  const int KK = K * K, pad = (K - 1) / 2;
  const int Kc = mode == 0 ? Cin : Cout, N = mode == 0 ? Cout : Cin;
  const int64_t M = (int64_t)B * H * W;
  const int64_t nw = (int64_t)KK * Cin * Cout;
  if (!tc_eligible(Kc, N, H, W, K))
    IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the requested tcgen05 engine");
  __nv_bfloat16 *wh = nullptr, *wl = nullptr, *ah = nullptr, *al = nullptr;
  float *out = nullptr, *part = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int rc = IGM_OK;
  auto fail = [&](cudaError_t e) {
    if (e != cudaSuccess && rc == IGM_OK) { set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e)); rc = IGM_ERR_CUDA; }
  };
  fail(cudaMalloc(&wh, nw * 2)); fail(cudaMalloc(&wl, nw * 2));
  fail(cudaMalloc(&ah, M * Kc * 2)); fail(cudaMalloc(&al, M * Kc * 2));
  fail(cudaMalloc(&out, M * N * 4));
  fail(cudaMalloc(&part, (size_t)B * (H * W / 16 + 64) * kGroups * 2 * 4));
  fail(cudaEventCreate(&e0)); fail(cudaEventCreate(&e1));
  if (rc == IGM_OK) {
    // bf16 1.0 / small values: the timing does not depend on the data, only finite operands are needed
    fail(cudaMemsetAsync(wh, 0x3c, nw * 2, lc.stream)); fail(cudaMemsetAsync(wl, 0x30, nw * 2, lc.stream));
    fail(cudaMemsetAsync(ah, 0x3c, M * Kc * 2, lc.stream)); fail(cudaMemsetAsync(al, 0x30, M * Kc * 2, lc.stream));
  }
  TcConv t;
  if (rc == IGM_OK) rc = tc_plan(st, t, Kc, N, H, W, B, K, pad, ah, al, wh, wl);
  TcConvPair tp;
  if (rc == IGM_OK && engine == 3) rc = tc2_plan(st, tp, t);
  TcRun r;
  r.B = B; r.out0 = out; r.N0 = N; r.kclass = mode == 0 ? K_CONV_FPROP : K_CONV_DGRAD;
  if (rc == IGM_OK && gn) {
    if (tc_gn_fusable(t, B)) r.gn_part = part;
  }
  for (int i = 0; i < warm + iters && rc == IGM_OK; ++i) {
    if (i == warm) fail(cudaEventRecord(e0, lc.stream));
    if (rc == IGM_OK) rc = engine == 3 ? launch_conv_tc2(lc, tp, r) : launch_conv_tc(lc, t, r);
  }
  if (rc == IGM_OK) {
    fail(cudaEventRecord(e1, lc.stream));
    fail(cudaStreamSynchronize(lc.stream));
    float ms = 0.f;
    if (rc == IGM_OK) fail(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / (float)iters;
  } else {
    cudaStreamSynchronize(lc.stream);
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(wh); cudaFree(wl); cudaFree(ah); cudaFree(al); cudaFree(out); cudaFree(part);
  return rc;
}

// Weight gradient of a stride-1 KxK conv: gw[Cout,Cin,K,K] += sum dY * X.  engine 0 = SIMT split-K,
// engine 1 = tcgen05 (variant selects the MN-major descriptor convention during bring-up).
extern "C" int igm_debug_wgrad(int engine, int variant, const float* x, const float* dy, float* gw, int B, int H,
                               int W, int Cin, int Cout, int K, void* stream) {
  Status& st = global_status();
  st = Status();
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  if (K != 1 && K != 3) IGM_FAIL(st, IGM_ERR_INVALID, "K must be 1 or 3");
  const int KK = K * K, pad = (K - 1) / 2;
  const int64_t M = (int64_t)B * H * W;
  int rc = IGM_OK;
  if (engine == 0) {
    WgradArgs w;
    w.P = dy; w.PC = Cout; w.PH = H; w.PW = W;
    w.Q = x; w.QC = Cin; w.QH = H; w.QW = W;
    w.B = B; w.KH = w.KW = K; w.stride = 1; w.pad = pad; w.dil = 1;
    w.grad = gw; w.sq = KK; w.sp = (int64_t)Cin * KK;
    rc = launch_wgrad(lc, w);
    cudaStreamSynchronize(lc.stream);
    return rc;
  }
  if (!tcw_eligible(Cin, Cout, H, W, K)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the tcgen05 wgrad engine");
  __nv_bfloat16 *dh = nullptr, *dl = nullptr, *xh = nullptr, *xl = nullptr;
  IGM_CUDA(st, cudaMalloc(&dh, M * Cout * 2));
  IGM_CUDA(st, cudaMalloc(&dl, M * Cout * 2));
  IGM_CUDA(st, cudaMalloc(&xh, M * Cin * 2));
  IGM_CUDA(st, cudaMalloc(&xl, M * Cin * 2));
  TcWgrad t;
  rc = tcw_plan(st, t, Cin, Cout, H, W, B, K, pad, dh, dl, xh, xl);
  if (rc == IGM_OK) rc = launch_split_bf16(lc, dy, M, Cout, dh, dl, Cout, 0);
  if (rc == IGM_OK) rc = launch_split_bf16(lc, x, M, Cin, xh, xl, Cin, 0);
  if (rc == IGM_OK) rc = launch_wgrad_tc(lc, t, B, gw, variant);
  cudaError_t e = cudaStreamSynchronize(lc.stream);
  if (rc == IGM_OK && e != cudaSuccess) {
    set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e));
    rc = IGM_ERR_CUDA;
  }
  cudaFree(dh); cudaFree(dl); cudaFree(xh); cudaFree(xl);
  return rc;
}

// Stride-2 resampling convolutions of the U-Net on either engine (kernel-level parity tests).
//   kind 0: Downsample  Conv2d(C, C2, 3, stride 2, pad 1)           weight OIHW [C2][C][3][3]
//   kind 1: Upsample    ConvTranspose2d(C, C2, 4, stride 2, pad 1)  weight IOHW [C][C2][4][4]
//   mode 0: forward      x [B,H,W,C]        -> out [B,OH,OW,C2]   (+bias)
//   mode 1: data grad    x = dY [B,OH,OW,C2] -> out = dX [B,H,W,C] (+add)
//   mode 2: weight grad  x [B,H,W,C], aux = dY [B,OH,OW,C2]       -> out = dW (accumulated, weight layout)
// (OH, OW) = (H/2, W/2) for kind 0 and (2H, 2W) for kind 1.
extern "C" int igm_debug_resample(int engine, int kind, int mode, const float* x, const float* aux, const float* w,
                                  const float* bias, const float* add, float* out, int B, int H, int W, int C, int C2,
                                  void* stream) {
  Status& st = global_status();
  st = Status();
  int64_t launches = 0;
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &launches;
  const int K = kind == 0 ? 3 : 4, KK = K * K, pad = 1;
  const int OH = kind == 0 ? H / 2 : 2 * H, OW = kind == 0 ? W / 2 : 2 * W;
  const int64_t Min = (int64_t)B * H * W, Mout = (int64_t)B * OH * OW;
  const int64_t nw = (int64_t)KK * C * C2;
  int rc = IGM_OK;
  // strides of (input channel ci in C, output channel co in C2) inside the PyTorch weight
  const int64_t s_ci = kind == 0 ? KK : (int64_t)C2 * KK;
  const int64_t s_co = kind == 0 ? (int64_t)C * KK : KK;
  if (engine == 0) {
    if (mode == 2) {
      WgradArgs g;
      if (kind == 0) { g.P = aux; g.PC = C2; g.PH = OH; g.PW = OW; g.Q = x; g.QC = C; g.QH = H; g.QW = W; g.sq = s_ci; g.sp = s_co; }
      else { g.P = x; g.PC = C; g.PH = H; g.PW = W; g.Q = aux; g.QC = C2; g.QH = OH; g.QW = OW; g.sq = s_co; g.sp = s_ci; }
      g.B = B; g.KH = g.KW = K; g.stride = 2; g.pad = pad; g.dil = 1; g.grad = out;
      rc = launch_wgrad(lc, g);
      cudaStreamSynchronize(lc.stream);
      return rc;
    }
    float* wp = nullptr;
    IGM_CUDA(st, cudaMalloc(&wp, nw * sizeof(float)));
    ConvArgs a;
    a.B = B; a.KH = a.KW = K; a.stride = 2; a.pad = pad; a.bias = bias; a.out0 = out; a.add0 = add;
    if (mode == 0) {
      rc = launch_pack_weight(lc, w, wp, KK, C, C2, s_ci, s_co);
      a.in0 = x; a.C0 = C; a.IH = H; a.IW = W; a.OH = OH; a.OW = OW; a.N = a.N0 = C2; a.transposed = kind;
    } else {
      rc = launch_pack_weight(lc, w, wp, KK, C2, C, s_co, s_ci);
      a.in0 = x; a.C0 = C2; a.IH = OH; a.IW = OW; a.OH = H; a.OW = W; a.N = a.N0 = C; a.transposed = 1 - kind;
    }
    a.w = wp;
    if (rc == IGM_OK) rc = launch_conv(lc, a);
    cudaStreamSynchronize(lc.stream);
    cudaFree(wp);
    return rc;
  }
  // ---- tcgen05 engine ----
  __nv_bfloat16 *wh = nullptr, *wl = nullptr, *xh = nullptr, *xl = nullptr, *ah = nullptr, *al = nullptr;
  IGM_CUDA(st, cudaMalloc(&wh, nw * 2));
  IGM_CUDA(st, cudaMalloc(&wl, nw * 2));
  const int64_t nx = (mode == 1) ? Mout * C2 : Min * C;
  IGM_CUDA(st, cudaMalloc(&xh, nx * 2));
  IGM_CUDA(st, cudaMalloc(&xl, nx * 2));
  rc = launch_split_bf16(lc, x, mode == 1 ? Mout : Min, mode == 1 ? C2 : C, xh, xl, mode == 1 ? C2 : C, 0);
  if (mode == 2) {
    IGM_CUDA(st, cudaMalloc(&ah, Mout * C2 * 2));
    IGM_CUDA(st, cudaMalloc(&al, Mout * C2 * 2));
    if (rc == IGM_OK) rc = launch_split_bf16(lc, aux, Mout, C2, ah, al, C2, 0);
    TcWgrad t;
    if (rc == IGM_OK) {
      if (kind == 0) rc = tcw_plan_strided(st, t, C, C2, OH, OW, B, K, pad, xh, xl, ah, al, s_ci, s_co);   // S = X (fine), P = dY
      else rc = tcw_plan_strided(st, t, C2, C, H, W, B, K, pad, ah, al, xh, xl, s_co, s_ci);                  // S = dY (fine), P = X
    }
    if (rc == IGM_OK) rc = launch_wgrad_tc(lc, t, B, out);
  } else {
    // forward contracts over C (n = C2), data gradient over C2 (n = C); taps are never flipped here
    if (mode == 0) { if (rc == IGM_OK) rc = launch_pack_weight_tc(lc, w, wh, wl, KK, C, C2, s_ci, s_co, 0); }
    else { if (rc == IGM_OK) rc = launch_pack_weight_tc(lc, w, wh, wl, KK, C2, C, s_co, s_ci, 0); }
    TcRun r;
    r.B = B; r.bias = bias; r.out0 = out; r.add0 = add;
    const bool strided = (kind == 0 && mode == 0) || (kind == 1 && mode == 1);
    if (strided) {
      TcConv t;
      const int Kc = mode == 0 ? C : C2, N = mode == 0 ? C2 : C;
      const int SH = mode == 0 ? H : OH, SW = mode == 0 ? W : OW;
      if (rc == IGM_OK) rc = tc_plan_strided(st, t, Kc, N, SH, SW, B, K, pad, xh, xl, wh, wl);
      r.N0 = N;
      if (rc == IGM_OK) rc = launch_conv_tc(lc, t, r);
    } else {
      const int Kc = mode == 0 ? C : C2, N = mode == 0 ? C2 : C;
      const int GH = mode == 0 ? H : OH, GW = mode == 0 ? W : OW;
      r.N0 = N;
      for (int ph = 0; ph < 4 && rc == IGM_OK; ++ph) {
        TcConv t;
        rc = tc_plan_phase(st, t, Kc, N, GH, GW, B, K, pad, ph / 2, ph % 2, xh, xl, wh, wl);
        if (rc == IGM_OK) rc = launch_conv_tc(lc, t, r);
      }
    }
  }
  cudaError_t e = cudaStreamSynchronize(lc.stream);
  if (rc == IGM_OK && e != cudaSuccess) {
    set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e));
    rc = IGM_ERR_CUDA;
  }
  cudaFree(wh); cudaFree(wl); cudaFree(xh); cudaFree(xl); cudaFree(ah); cudaFree(al);
  return rc;
}
