// Tensor-core route of the generic convolution operator (igm_conv2d_forward / igm_conv2d_backward, ops.cu): stride-1
// 1x1 / 3x3 "same" convolutions (and their ConvTranspose2d twins) with 64-multiple channel counts run on the tcgen05
// bf16x3 engine of conv_tc.cu / wgrad_tc.cu instead of the fp32 CUDA-core implicit GEMM.  These are the layers that hold
// 82 % of the FLOPs of the reference VQ-VAE (src/networks/vqvae.py: the 3x3 convs and the tied residual stacks at 32x32)
// and the 1x1 convs of the PixelCNN training path (src/models/pixelcnn.py:58-82).
//
// The operator is context-free (torch owns every tensor), so the operand staging the DDPM engine gets for free from its
// producers is done here per call: fp32 -> bf16 hi / lo split of the input (and of dY in backward) and the weight
// re-pack, into the caller's workspace.  TMA descriptors are cached per (geometry, pointers): torch's caching allocator
// hands the same blocks back step after step, so a training loop hits the cache after its first iteration.
#include <map>
#include <string.h>

#include "common.cuh"
#include "conv_tc.cuh"

namespace igm {

bool ops_tc_enabled() {
  static const bool on = [] { const char* e = getenv("IGM_OPS_TC"); return !(e && e[0] == '0'); }();
  return on;
}

bool ops_tc_geo_ok(int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil, int OH, int OW) {
  if (!ops_tc_enabled()) return false;
  if (stride != 1 || dil != 1 || KH != KW || (KH != 1 && KH != 3)) return false;
  if (pad_h != (KH - 1) / 2 || pad_w != pad_h || OH != H || OW != W) return false;
  return tc_eligible(Cin, Cout, H, W, KH) && tc_eligible(Cout, Cin, H, W, KH);
}

// ConvTranspose2d(k = 4, s = 2, p = 1) with 64-multiple channels (the VQ-VAE decoder's first upsampling layer, 128 -> 64 at
// 32x32 -> 64x64: 13 % of the network's FLOPs): forward = four output-parity phases, data gradient = stride-2 gather over
// dY, weight gradient = stride-2 wgrad -- the plans the DDPM Upsample layer uses (conv_tc.cu / wgrad_tc.cu)
bool ops_tc_convT2_ok(int H, int W, int Cin, int Cout, int KH, int KW, int stride, int pad_h, int pad_w, int dil, int transposed,
                      int OH, int OW) {
  if (!ops_tc_enabled() || !transposed) return false;
  if (stride != 2 || dil != 1 || KH != 4 || KW != 4 || pad_h != 1 || pad_w != 1 || OH != 2 * H || OW != 2 * W) return false;
  if (Cin % 64 != 0 || Cout % 64 != 0 || W > 128) return false;
  return tc_strided_eligible(Cout, Cin, OH, OW, 4) && tcw_strided_eligible(Cout, Cin, H, W, 4);
}

static int64_t r64(int64_t n) { return (n + 63) & ~int64_t(63); }

// extra workspace floats of the tensor-core routes: packed bf16 weights | x pair | dy pair
int64_t ops_tc_ws_floats(int B, int H, int W, int Cin, int Cout, int KH, int OH, int OW) {
  const int64_t nw = (int64_t)KH * KH * Cin * Cout;
  return r64(nw) + r64((int64_t)B * H * W * Cin) + r64((int64_t)B * OH * OW * Cout) + 64;
}

namespace {

struct PlanKey {
  int kind;   // 0 conv plan, 1 wgrad plan
  int B, H, W, K, N, KH;
  const void *a, *b, *c, *d;
  bool operator<(const PlanKey& o) const { return memcmp(this, &o, sizeof(PlanKey)) < 0; }
};

template <class T>
struct PlanCache {
  std::map<PlanKey, T> m;
  T* get(const PlanKey& k, bool& fresh) {
    auto it = m.find(k);
    fresh = it == m.end();
    if (fresh) {
      if (m.size() > 512) m.clear();   // pointers churned (a different model / batch): start over
      it = m.emplace(k, T()).first;
    }
    return &it->second;
  }
};

PlanCache<TcConv>& conv_plans() { static PlanCache<TcConv> c; return c; }
PlanCache<TcWgrad>& wgrad_plans() { static PlanCache<TcWgrad> c; return c; }

PlanKey make_key(int kind, int B, int H, int W, int K, int N, int KH, const void* a, const void* b, const void* c, const void* d) {
  PlanKey k;
  memset(&k, 0, sizeof(k));
  k.kind = kind; k.B = B; k.H = H; k.W = W; k.K = K; k.N = N; k.KH = KH; k.a = a; k.b = b; k.c = c; k.d = d;
  return k;
}

int get_conv_plan(Status& st, const TcConv*& out, int K, int N, int H, int W, int B, int KH, __nv_bfloat16* ah, __nv_bfloat16* al,
                  __nv_bfloat16* wh, __nv_bfloat16* wl) {
  bool fresh;
  TcConv* t = conv_plans().get(make_key(0, B, H, W, K, N, KH, ah, al, wh, wl), fresh);
  if (fresh || !t->valid) IGM_TRY(tc_plan(st, *t, K, N, H, W, B, KH, (KH - 1) / 2, ah, al, wh, wl));
  out = t;
  return IGM_OK;
}

}  // namespace

// y = conv(x) (+bias) (+residual).  transposed: ConvTranspose2d weight IOHW (stride 1: a conv with flipped taps).
int ops_tc_conv_forward(const LaunchCtx& lc, const float* x, const float* w, const float* bias, const float* residual, float* y,
                        int B, int H, int W, int Cin, int Cout, int KH, int transposed, float* ws) {
  const int KK = KH * KH;
  const int64_t M = (int64_t)B * H * W, nw = (int64_t)KK * Cin * Cout;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* wl = wh + nw;
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw));
  __nv_bfloat16* xl = xh + M * Cin;
  // Wt[n = co][tap * Cin + ci]
  if (!transposed) IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cin, Cout, KK, (int64_t)Cin * KK, 0));
  else IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cin, Cout, (int64_t)Cout * KK, KK, 1));
  IGM_TRY(launch_split_bf16(lc, x, M, Cin, xh, xl, Cin, 0));
  const TcConv* t = nullptr;
  IGM_TRY(get_conv_plan(*lc.st, t, Cin, Cout, H, W, B, KH, xh, xl, wh, wl));
  TcRun r;
  r.B = B; r.bias = bias; r.out0 = y; r.N0 = Cout; r.add0 = residual; r.kclass = K_CONV_FPROP;
  return launch_conv_tc(lc, *t, r);
}

// dx (nullable), dw (nullable, ACCUMULATED, the weight's own layout).  Returns IGM_OK with *did_dw = 0 when the weight
// gradient has to take the CUDA-core path (batch not a multiple of the images per pixel box).
int ops_tc_conv_backward(const LaunchCtx& lc, const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H,
                         int W, int Cin, int Cout, int KH, int transposed, float* ws, int* did_dw) {
  const int KK = KH * KH;
  const int64_t M = (int64_t)B * H * W, nw = (int64_t)KK * Cin * Cout;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* wl = wh + nw;
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw));
  __nv_bfloat16* xl = xh + M * Cin;
  __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw) + r64(M * Cin));
  __nv_bfloat16* dl = dh + M * Cout;
  *did_dw = 0;
  IGM_TRY(launch_split_bf16(lc, dy, M, Cout, dh, dl, Cout, 0));
  if (dx) {
    // data gradient: contraction over co.  Wt[n = ci][tap * Cout + co], taps flipped for Conv2d, as they are for its twin
    if (!transposed) IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cout, Cin, (int64_t)Cin * KK, KK, 1));
    else IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cout, Cin, KK, (int64_t)Cout * KK, 0));
    const TcConv* t = nullptr;
    IGM_TRY(get_conv_plan(*lc.st, t, Cout, Cin, H, W, B, KH, dh, dl, wh, wl));
    TcRun r;
    r.B = B; r.out0 = dx; r.N0 = Cin; r.kclass = K_CONV_DGRAD;
    IGM_TRY(launch_conv_tc(lc, *t, r));
  }
  if (dw && tcw_eligible(Cin, Cout, H, W, KH)) {
    IGM_TRY(launch_split_bf16(lc, x, M, Cin, xh, xl, Cin, 0));
    bool fresh;
    // Conv2d: dW[co][ci][k] = sum dY[p + pad - k][co] X[p][ci].  ConvTranspose2d (IOHW): the same sum with the roles of
    // x and dy exchanged, dW[ci][co][k] = sum X[p + pad - k][ci] dY[p][co]
    TcWgrad* t = wgrad_plans().get(make_key(1, B, H, W, transposed ? Cout : Cin, transposed ? Cin : Cout, KH, dh, dl, xh, xl), fresh);
    if (fresh || !t->valid) {
      if (!transposed) IGM_TRY(tcw_plan(*lc.st, *t, Cin, Cout, H, W, B, KH, (KH - 1) / 2, dh, dl, xh, xl));
      else IGM_TRY(tcw_plan(*lc.st, *t, Cout, Cin, H, W, B, KH, (KH - 1) / 2, xh, xl, dh, dl));
    }
    if (tcw_batch_ok(*t, B)) {
      IGM_TRY(launch_wgrad_tc(lc, *t, B, dw));
      *did_dw = 1;
    }
  }
  return IGM_OK;
}

// ---- ConvTranspose2d k4 s2 p1 ---------------------------------------------------------------------------------------
int ops_tc_convT2_forward(const LaunchCtx& lc, const float* x, const float* w, const float* bias, float* y, int B, int H, int W,
                          int Cin, int Cout, float* ws) {
  const int KK = 16;
  const int64_t M = (int64_t)B * H * W, nw = (int64_t)KK * Cin * Cout;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* wl = wh + nw;
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw));
  __nv_bfloat16* xl = xh + M * Cin;
  IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cin, Cout, (int64_t)Cout * KK, KK, 0));   // IOHW, taps as they are
  IGM_TRY(launch_split_bf16(lc, x, M, Cin, xh, xl, Cin, 0));
  bool fresh;
  TcConv* t = conv_plans().get(make_key(2, B, H, W, Cin, Cout, 4, xh, xl, wh, wl), fresh);
  if (fresh || !t->valid) IGM_TRY(tc_plan_phases4(*lc.st, *t, Cin, Cout, H, W, B, 4, 1, xh, xl, wh, wl));
  TcRun r;
  r.B = B; r.bias = bias; r.out0 = y; r.N0 = Cout; r.kclass = K_CONV_FPROP;
  return launch_conv_tc(lc, *t, r);
}

int ops_tc_convT2_backward(const LaunchCtx& lc, const float* x, const float* w, const float* dy, float* dx, float* dw, int B, int H,
                           int W, int Cin, int Cout, float* ws, int* did_dw) {
  const int KK = 16;
  const int64_t M = (int64_t)B * H * W, Mo = 4 * M, nw = (int64_t)KK * Cin * Cout;
  __nv_bfloat16* wh = reinterpret_cast<__nv_bfloat16*>(ws);
  __nv_bfloat16* wl = wh + nw;
  __nv_bfloat16* xh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw));
  __nv_bfloat16* xl = xh + M * Cin;
  __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(ws + r64(nw) + r64(M * Cin));
  __nv_bfloat16* dl = dh + Mo * Cout;
  *did_dw = 0;
  IGM_TRY(launch_split_bf16(lc, dy, Mo, Cout, dh, dl, Cout, 0));
  if (dx) {
    // dx[i] = sum_k dy[2 i - 1 + k] w[ci][co][k]: a stride-2 gather over dY, contraction over co (IOHW: co stride KK)
    IGM_TRY(launch_pack_weight_tc(lc, w, wh, wl, KK, Cout, Cin, KK, (int64_t)Cout * KK, 0));
    bool fresh;
    TcConv* t = conv_plans().get(make_key(3, B, H, W, Cout, Cin, 4, dh, dl, wh, wl), fresh);
    if (fresh || !t->valid) IGM_TRY(tc_plan_strided(*lc.st, *t, Cout, Cin, 2 * H, 2 * W, B, 4, 1, dh, dl, wh, wl));
    TcRun r;
    r.B = B; r.out0 = dx; r.N0 = Cin; r.kclass = K_CONV_DGRAD;
    IGM_TRY(launch_conv_tc(lc, *t, r));
  }
  if (dw) {
    IGM_TRY(launch_split_bf16(lc, x, M, Cin, xh, xl, Cin, 0));
    bool fresh;
    // S = dY (fine grid, co), P = X (coarse grid, ci); IOHW: co stride KK, ci stride Cout * KK
    TcWgrad* t = wgrad_plans().get(make_key(4, B, H, W, Cout, Cin, 4, dh, dl, xh, xl), fresh);
    if (fresh || !t->valid) IGM_TRY(tcw_plan_strided(*lc.st, *t, Cout, Cin, H, W, B, 4, 1, dh, dl, xh, xl, KK, (int64_t)Cout * KK));
    if (tcw_batch_ok(*t, B)) {
      IGM_TRY(launch_wgrad_tc(lc, *t, B, dw));
      *did_dw = 1;
    }
  }
  return IGM_OK;
}

}  // namespace igm
