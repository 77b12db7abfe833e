// Test-only entry point: one stride-1 convolution (or its data gradient) on either engine,
// so the tcgen05 kernel can be parity-checked in isolation over many shapes.
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
  rc = tc_plan(st, t, Kc, N, H, W, B, K, pad, ah, al, wh, wl);
  if (rc == IGM_OK) rc = launch_pack_weight_tc(lc, w_oihw, wh, wl, KK, Kc, N, sk, sn, mode);
  if (rc == IGM_OK) rc = launch_split_bf16(lc, x, M, Kc, ah, al, Kc, 0);
  if (rc == IGM_OK) {
    TcRun r;
    r.B = B; r.bias = bias; r.out0 = out; r.N0 = N; r.add0 = add;
    rc = launch_conv_tc(lc, t, r);
  }
  cudaError_t e = cudaStreamSynchronize(lc.stream);
  if (rc == IGM_OK && e != cudaSuccess) {
    set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e));
    rc = IGM_ERR_CUDA;
  }
  cudaFree(wh); cudaFree(wl); cudaFree(ah); cudaFree(al);
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
