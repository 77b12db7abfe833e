// Context-free operator entry points (NHWC fp32 = torch channels_last): the generic convolution family
// of conv_simt.cu plus the small activations the VQ-VAE encoder/decoder (reference src/networks/vqvae.py)
// and the PixelCNN training path (src/models/pixelcnn.py:64-82, :148, :156-165) need.  These serve the
// secondary models; the DDPM U-Net uses the planned, fused path of unet.cu.
#include "common.cuh"

namespace igm {
namespace {

__global__ void relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaxf(x[i], 0.f);
}
// dx = dy * (y > 0)   (y = relu output; in-place friendly)
__global__ void relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}
// gated activation over the channel halves of x [M, 2C]: mode 0 tanh(a)*sigmoid(b), mode 1 tanh(a)*tanh(b)
// cond (nullable): [M / hw, 2C] per-image addend of the pre-activation (class conditioning, pixelcnn.py:71,:79)
__global__ void gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ cond, int64_t hw,
                                float* __restrict__ y, int64_t M, int C, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int64_t m = i / C;
  const int c = (int)(i - m * C);
  float a = x[m * 2 * C + c], b = x[m * 2 * C + C + c];
  if (cond) { const float* cp = cond + (m / hw) * 2 * C; a = __fadd_rn(a, cp[c]); b = __fadd_rn(b, cp[C + c]); }
  const float g = mode == 0 ? 1.f / (1.f + expf(-b)) : tanhf(b);
  y[i] = tanhf(a) * g;
}
__global__ void gate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ cond, int64_t hw,
                                const float* __restrict__ dy, float* __restrict__ dx, int64_t M, int C, int mode) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int64_t m = i / C;
  const int c = (int)(i - m * C);
  float a = x[m * 2 * C + c], b = x[m * 2 * C + C + c];
  if (cond) { const float* cp = cond + (m / hw) * 2 * C; a = __fadd_rn(a, cp[c]); b = __fadd_rn(b, cp[C + c]); }
  const float ta = tanhf(a);
  const float g = mode == 0 ? 1.f / (1.f + expf(-b)) : tanhf(b);
  const float dg = mode == 0 ? g * (1.f - g) : 1.f - g * g;
  const float d = dy[i];
  dx[m * 2 * C + c] = d * (1.f - ta * ta) * g;
  dx[m * 2 * C + C + c] = d * ta * dg;
}
__global__ void elu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; y[i] = v > 0.f ? v : expm1f(v); }
}
__global__ void elu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float v = x[i]; dx[i] = v > 0.f ? dy[i] : dy[i] * expf(v); }
}

// per-image column sums: out[b, n] = sum over the image's hw rows of x[(b*hw + r), n]   (conditioning gradient)
__global__ void __launch_bounds__(256) image_colsum_kernel(const float* __restrict__ x, int64_t hw, int N, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = threadIdx.x >> 5;
  __shared__ float red[8][33];
  float acc = 0.f;
  if (n < N)
    for (int64_t r = r0; r < hw; r += 8) acc += x[((int64_t)b * hw + r) * N + n];
  red[r0][threadIdx.x & 31] = acc;
  __syncthreads();
  if (r0 == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
    out[(int64_t)b * N + n] = t;
  }
}

// kind 0: y = a + b; kind 1: y = a + (b - a)  (the straight-through value of vqvae.py:103, rounded like torch)
__global__ void ewise_kernel(int kind, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = kind == 0 ? __fadd_rn(a[i], b[i]) : __fadd_rn(a[i], __fsub_rn(b[i], a[i]));
}

// mean squared error: out[0] += sum (a-b)^2 / n (out pre-zeroed); and its gradient da = scale[0] * 2 (a-b) / n
__global__ void __launch_bounds__(256) mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                      float inv_n, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    acc += d * d;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j];
    atomicAdd(out, t * inv_n);
  }
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ scale,
                               float k, float* __restrict__ da, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) da[i] = scale[0] * k * (a[i] - b[i]);
}

// 256-way cross entropy of PixelCNN (pixelcnn.py:163): logits NHWC [M, 256*C] with channel o = cls*C + ch,
// target [M, C] int64.  One warp per (pixel, channel): nll[m*C + ch], optionally d_logits = scale*(softmax - onehot).
__global__ void __launch_bounds__(256) ce256_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
                                                    float* __restrict__ nll, float* __restrict__ d_logits,
                                                    const float* __restrict__ d_nll_scale, int64_t MC, int C) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= MC) return;
  const int64_t m = w / C;
  const int ch = (int)(w - m * C);
  const float* lp = logits + m * 256 * C + ch;
  float v[8];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = lp[(int64_t)(lane + 32 * j) * C]; mx = fmaxf(mx, v[j]); }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += expf(v[j] - mx);
  s = warp_sum(s);
  const float lse = mx + logf(s);
  const int t = (int)target[w];
  if (nll && lane == (t & 31)) nll[w] = lse - v[t >> 5];
  if (d_logits) {
    const float sc = d_nll_scale ? d_nll_scale[w] : 1.f;
    float* dp = d_logits + m * 256 * C + ch;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int cls = lane + 32 * j;
      dp[(int64_t)cls * C] = sc * (expf(v[j] - lse) - (cls == t ? 1.f : 0.f));
    }
  }
}

LaunchCtx make_lc(Status& st, int64_t& counter, void* stream) {
  LaunchCtx lc;
  lc.stream = (cudaStream_t)stream;
  lc.st = &st;
  lc.counter = &counter;
  return lc;
}

struct Geo {
  int B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW;
};

int check_geo(Status& st, const Geo& g) {
  if (g.B < 1 || g.H < 1 || g.W < 1 || g.Cin < 1 || g.Cout < 1 || g.KH < 1 || g.KW < 1 || g.KH * g.KW > 32 || g.stride < 1 ||
      g.dil < 1 || g.OH < 1 || g.OW < 1)
    IGM_FAIL(st, IGM_ERR_INVALID, "conv2d: bad geometry");
  if (g.transposed && g.stride > 1 && g.pad_h != g.pad_w) IGM_FAIL(st, IGM_ERR_INVALID, "strided transposed conv needs square padding");
  return IGM_OK;
}

}  // namespace

// tensor-core route (ops_tc.cu)
bool ops_tc_geo_ok(int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil, int OH, int OW);
bool ops_tc_convT2_ok(int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil, int transposed,
                      int OH, int OW);
int64_t ops_tc_ws_floats(int B, int H, int W, int Cin, int Cout, int KH, int OH, int OW);
int ops_tc_convT2_forward(const LaunchCtx& lc, const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                          int Cin, int Cout, float* ws);
int ops_tc_convT2_backward(const LaunchCtx& lc, const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H,
                           int W, int Cin, int Cout, float* ws, int* did_dw);
int ops_tc_conv_forward(const LaunchCtx& lc, const float* x, const float* w, const float* bias, const float* residual, float* y,
                        int B, int H, int W, int Cin, int Cout, int KH, int transposed, float* ws);
int ops_tc_conv_backward(const LaunchCtx& lc, const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H,
                         int W, int Cin, int Cout, int KH, int transposed, float* ws, int* did_dw);
}  // namespace igm

using namespace igm;

// workspace floats a conv forward / backward call needs: the fp32 packed weights of the CUDA-core engine, followed (when the
// geometry qualifies for the tensor-core route) by the bf16 weight / activation / gradient staging pairs
extern "C" int64_t igm_conv2d_workspace_floats(int B, int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h,
                                               int pad_w, int dil, int OH, int OW) {
  int64_t n = (int64_t)KH * KW * Cin * Cout + 64;
  // sized for either tensor-core route (the transposed flag is not part of this query: both are "64-multiple channels")
  if (ops_tc_geo_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, OH, OW) ||
      ops_tc_convT2_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, 1, OH, OW))
    n += ops_tc_ws_floats(B, H, W, Cin, Cout, KH, OH, OW);
  return n;
}

// y[B,OH,OW,Cout] = conv(x[B,H,W,Cin]) (+bias).  transposed = 0: Conv2d, weight OIHW [Cout,Cin,KH,KW];
// transposed = 1: ConvTranspose2d, weight IOHW [Cin,Cout,KH,KW].  Dilation applies to Conv2d only.
// residual (nullable, y's shape) is added in the epilogue.
extern "C" int igm_conv2d_forward(const float* x, const float* w, const float* bias, const float* residual, float* y, int B,
                                  int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil,
                                  int transposed, int OH, int OW, float* ws, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!x || !w || !y || !ws) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  Geo g{B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW};
  IGM_TRY(check_geo(st, g));
  LaunchCtx lc = make_lc(st, ops_launch_counter(), stream);
  const int KK = KH * KW;
  if (ops_tc_geo_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, OH, OW))
    return ops_tc_conv_forward(lc, x, w, bias, residual, y, B, H, W, Cin, Cout, KH, transposed, ws + (int64_t)KK * Cin * Cout + 64);
  if (!residual && ops_tc_convT2_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW))
    return ops_tc_convT2_forward(lc, x, w, bias, y, B, H, W, Cin, Cout, ws + (int64_t)KK * Cin * Cout + 64);
  // packed [tap][ci][co]
  if (!transposed) IGM_TRY(launch_pack_weight(lc, w, ws, KK, Cin, Cout, KK, (int64_t)Cin * KK));
  else IGM_TRY(launch_pack_weight(lc, w, ws, KK, Cin, Cout, (int64_t)Cout * KK, KK));
  ConvArgs a;
  a.in0 = x; a.C0 = Cin; a.B = B; a.IH = H; a.IW = W; a.OH = OH; a.OW = OW;
  a.N = a.N0 = Cout; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad_h; a.pad_w = pad_w; a.dil = dil;
  a.transposed = transposed; a.w = ws; a.bias = bias; a.out0 = y; a.add0 = residual;
  return launch_conv(lc, a);
}

// dx (nullable) = data gradient; dw += weight gradient (PyTorch layout); db (nullable) += bias gradient.
extern "C" int igm_conv2d_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int B,
                                   int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil,
                                   int transposed, int OH, int OW, float* ws, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!x || !w || !dy || !ws) IGM_FAIL(st, IGM_ERR_INVALID, "null tensor");
  Geo g{B, H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW};
  IGM_TRY(check_geo(st, g));
  LaunchCtx lc = make_lc(st, ops_launch_counter(), stream);
  const int KK = KH * KW;
  if (ops_tc_geo_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, OH, OW)) {
    int did_dw = 0;
    IGM_TRY(ops_tc_conv_backward(lc, x, w, dy, dx, dw, B, H, W, Cin, Cout, KH, transposed, ws + (int64_t)KK * Cin * Cout + 64, &did_dw));
    dx = nullptr;                 // done on the tensor cores
    if (did_dw) dw = nullptr;
  } else if (ops_tc_convT2_ok(H, W, Cin, Cout, KH, KW, stride, pad_h, pad_w, dil, transposed, OH, OW)) {
    int did_dw = 0;
    IGM_TRY(ops_tc_convT2_backward(lc, x, w, dy, dx, dw, B, H, W, Cin, Cout, ws + (int64_t)KK * Cin * Cout + 64, &did_dw));
    dx = nullptr;
    if (did_dw) dw = nullptr;
  }
  if (dx) {
    // packed [tap][co][ci]; the data gradient of a Conv2d is a transposed gather and vice versa
    if (!transposed) IGM_TRY(launch_pack_weight(lc, w, ws, KK, Cout, Cin, (int64_t)Cin * KK, KK));
    else IGM_TRY(launch_pack_weight(lc, w, ws, KK, Cout, Cin, KK, (int64_t)Cout * KK));
    ConvArgs a;
    a.in0 = dy; a.C0 = Cout; a.B = B; a.IH = OH; a.IW = OW; a.OH = H; a.OW = W;
    a.N = a.N0 = Cin; a.KH = KH; a.KW = KW; a.stride = stride; a.pad = pad_h; a.pad_w = pad_w; a.dil = dil;
    a.transposed = transposed ? 0 : 1; a.kclass = K_CONV_DGRAD; a.w = ws; a.out0 = dx;
    IGM_TRY(launch_conv(lc, a));
  }
  if (dw) {
    WgradArgs q;
    q.B = B; q.KH = KH; q.KW = KW; q.stride = stride; q.pad = pad_h; q.pad_w = pad_w; q.dil = dil; q.grad = dw;
    if (!transposed) {   // P = dy (co), Q = x (ci) gathered; OIHW
      q.P = dy; q.PC = Cout; q.PH = OH; q.PW = OW; q.Q = x; q.QC = Cin; q.QH = H; q.QW = W;
      q.sq = KK; q.sp = (int64_t)Cin * KK;
    } else {             // P = x (ci), Q = dy (co) gathered; IOHW
      q.P = x; q.PC = Cin; q.PH = H; q.PW = W; q.Q = dy; q.QC = Cout; q.QH = OH; q.QW = OW;
      q.sq = KK; q.sp = (int64_t)Cout * KK;
    }
    IGM_TRY(launch_wgrad(lc, q));
  }
  if (db) IGM_TRY(launch_colsum(lc, dy, (int64_t)B * OH * OW, Cout, db));
  return IGM_OK;
}

// kind 0 relu, 1 elu (x = pre-activation), 2 gate tanh*sigmoid, 3 gate tanh*tanh (x: [M, 2C] -> y: [M, C])
// cond (nullable, gates only): [M / hw, 2C] per-image pre-activation addend.
extern "C" int igm_act_forward(int kind, const float* x, const float* cond, int64_t hw, float* y, int64_t M, int C, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!x || !y || M < 0 || C < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad activation args");
  const int64_t n = M * C;
  const unsigned grid = (unsigned)cdiv64(n > 0 ? n : 1, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == 0) relu_fwd_kernel<<<grid, 256, 0, s>>>(x, y, n);
  else if (kind == 1) elu_fwd_kernel<<<grid, 256, 0, s>>>(x, y, n);
  else if (kind == 2 || kind == 3) gate_fwd_kernel<<<grid, 256, 0, s>>>(x, cond, hw > 0 ? hw : 1, y, M, C, kind - 2);
  else IGM_FAIL(st, IGM_ERR_INVALID, "unknown activation");
  ++ops_launch_counter();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// relu: ref = the relu OUTPUT; elu / gates: ref = the pre-activation input x.  dx has x's shape.
// dcond (nullable, needs cond): [M / hw, 2C] = per-image sums of dx.
extern "C" int igm_act_backward(int kind, const float* ref, const float* cond, int64_t hw, const float* dy, float* dx,
                                float* dcond, int64_t M, int C, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!ref || !dy || !dx || M < 0 || C < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad activation args");
  const int64_t n = M * C;
  const unsigned grid = (unsigned)cdiv64(n > 0 ? n : 1, 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == 0) relu_bwd_kernel<<<grid, 256, 0, s>>>(ref, dy, dx, n);
  else if (kind == 1) elu_bwd_kernel<<<grid, 256, 0, s>>>(ref, dy, dx, n);
  else if (kind == 2 || kind == 3) {
    gate_bwd_kernel<<<grid, 256, 0, s>>>(ref, cond, hw > 0 ? hw : 1, dy, dx, M, C, kind - 2);
    if (dcond) {
      if (hw < 1 || M % hw) IGM_FAIL(st, IGM_ERR_INVALID, "conditioning needs M = images * hw");
      image_colsum_kernel<<<dim3((unsigned)cdiv64(2 * C, 32), (unsigned)(M / hw)), 256, 0, s>>>(dx, hw, 2 * C, dcond);
      ++ops_launch_counter();
    }
  } else IGM_FAIL(st, IGM_ERR_INVALID, "unknown activation");
  ++ops_launch_counter();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// PixelCNN's 256-way cross entropy on NHWC logits [M, 256*C] (channel = cls*C + ch), targets [M, C] int64:
// nll[M*C] (nullable) and/or d_logits = d_nll[M*C] (nullable: ones) * (softmax - onehot).
extern "C" int igm_ce256(const float* logits, const int64_t* target, float* nll, float* d_logits, const float* d_nll,
                         int64_t M, int C, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!logits || !target || M < 1 || C < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad cross-entropy args");
  const int64_t MC = M * C;
  ce256_kernel<<<(unsigned)cdiv64(MC * 32, 256), 256, 0, (cudaStream_t)stream>>>(logits, target, nll, d_logits, d_nll, MC, C);
  ++ops_launch_counter();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// kind 0: y = a + b; 1: y = a + (b - a) (straight-through estimator value, src/models/vqvae.py:103)
extern "C" int igm_ewise(int kind, const float* a, const float* b, float* y, int64_t n, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!a || !b || !y || n < 0 || kind < 0 || kind > 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad elementwise args");
  ewise_kernel<<<(unsigned)cdiv64(n > 0 ? n : 1, 256), 256, 0, (cudaStream_t)stream>>>(kind, a, b, y, n);
  ++ops_launch_counter();
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// F.mse_loss(a, b) (src/models/vqvae.py:106): loss (nullable) [1]; da (nullable) = d_loss[0] * 2 (a - b) / n.
extern "C" int igm_mse(const float* a, const float* b, int64_t n, float* loss, const float* d_loss, float* da, void* stream) {
  Status& st = global_status();
  st = Status();
  if (!a || !b || n < 1 || (da && !d_loss)) IGM_FAIL(st, IGM_ERR_INVALID, "bad mse args");
  cudaStream_t s = (cudaStream_t)stream;
  if (loss) {
    cudaMemsetAsync(loss, 0, sizeof(float), s);
    const unsigned grid = (unsigned)(cdiv64(n, 256) < 592 ? cdiv64(n, 256) : 592);
    mse_fwd_kernel<<<grid, 256, 0, s>>>(a, b, n, 1.f / (float)n, loss);
    ++ops_launch_counter();
  }
  if (da) {
    mse_bwd_kernel<<<(unsigned)cdiv64(n, 256), 256, 0, s>>>(a, b, d_loss, 2.f / (float)n, da, n);
    ++ops_launch_counter();
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) { cudaGetLastError(); IGM_FAIL(st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
  return IGM_OK;
}

// Kernels launched so far by the context-free entry points of this library (generic conv / activation / loss operators,
// igm_vq_*, igm_pixelcnn_run); bench.py reports the difference over its timed region as gpu_launches of the VQ-VAE line.
extern "C" int64_t igm_ops_launch_count(void) { return ops_launch_counter(); }
