// Layer plan, forward/backward orchestration and the C ABI of libigm_b200.
//
// The plan mirrors reference Unet.__init__ (src/models/ddpm.py:170-236) and the
// forward mirrors Unet.forward (:238-261); every activation / workspace buffer is
// carved once from a single HBM arena at igm_unet_create, so a forward, a
// backward or a sampler step is a fixed launch sequence with no allocation and
// can be captured into a CUDA graph (igm_ddpm_sample_loop does).
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>

#include "common.cuh"
#include "conv_tc.cuh"

namespace igm {

void set_error(Status& st, int code, const char* file, int line, const char* what) {
  st.code = code;
  char buf[512];
  snprintf(buf, sizeof(buf), "%s:%d: %s", file, line, what ? what : "?");
  st.msg = buf;
}

const char* kclass_name(int k) {
  static const char* names[K_NCLASS] = {"conv_fprop", "conv_dgrad", "conv_wgrad", "norm_act", "attention",
                                        "time_mlp",   "elementwise", "adam",      "pack"};
  return (k >= 0 && k < K_NCLASS) ? names[k] : "?";
}

cudaEvent_t Profiler::get() {
  cudaEvent_t e;
  if (!pool.empty()) {
    e = pool.back();
    pool.pop_back();
    return e;
  }
  cudaEventCreate(&e);
  return e;
}

void Profiler::reset() {
  for (auto& r : recs) {
    pool.push_back(r.a);
    pool.push_back(r.b);
  }
  recs.clear();
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("IGM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

Status& global_status() {
  static Status s;
  return s;
}

int64_t& ops_launch_counter() {
  static int64_t n = 0;
  return n;
}

namespace {

// IGM_CONV_PAIR=1: eligible stride-1 convs (N % 128 == 0) run forward and data gradient on the cta_group::2 engine
// (conv_tc2.cu).  Parity-green on B200 but 1.7-1.9x SLOWER than the per-tap engine on every layer shape
// (profiles/r2_conv_engines.md), so it stays an off-by-default comparison switch.
bool conv_pair_on() {
  static const bool on = [] { const char* e = getenv("IGM_CONV_PAIR"); return e && e[0] == '1'; }();
  return on;
}

struct ParamInfo {
  std::string name;
  int64_t offset = 0;
  int ndim = 0;
  int64_t shape[4] = {1, 1, 1, 1};
  int64_t numel() const { return shape[0] * shape[1] * shape[2] * shape[3]; }
};

// Activation in NHWC; v = value, g = gradient (training only)
struct Act {
  float* v = nullptr;
  float* g = nullptr;
  __nv_bfloat16* hi = nullptr;   // bf16 hi/lo staging of v, written by the producing kernel when the
  __nv_bfloat16* lo = nullptr;   // tcgen05 engine is active (operand of the convs that consume v)
  int C = 0, H = 0, W = 0;
};

struct ConvL {
  int dyb = 0;   // slot of the dY staging ring this layer's gradient is staged in
  int fin_begin = 0, fin_ctas = 0;   // CTA range of this layer in the halo-wgrad finalize table (set at bind time)
  int pw = -1, pb = -1;   // parameter indices
  int Cin = 0, Cout = 0, K = 1;
  bool convT = false;
  float* w_fwd = nullptr;   // [K*K][Cin][Cout]
  float* w_bwd = nullptr;   // [K*K][Cout][Cin]
  // tcgen05 engine (stride-1 Conv2d only): bf16 hi/lo weights [N][taps*K] and TMA plans
  int H = 0, W = 0;         // spatial size the layer runs at (0: not a stride-1 conv)
  bool tc_f_ok = false, tc_b_ok = false, tc_w_ok = false;
  __nv_bfloat16 *wf_hi = nullptr, *wf_lo = nullptr, *wb_hi = nullptr, *wb_lo = nullptr;
  TcConv tc_f, tc_b;
  TcConvPair tcp_f, tcp_b;   // CTA-pair plans on top of tc_f / tc_b (conv_tc2.cu), only when IGM_CONV_PAIR=1 (bring-up, default off)
  TcWgrad tc_w;
  bool tc_wh_ok = false;       // 3x3 stride-1: halo-reuse weight-gradient engine (wgrad_halo.cu)
  TcWgradHalo tc_wh;
  float* wg_ws = nullptr;      // its [3][3][Cin][Cout] reduction workspace
  bool bias_in_norm = false;   // bias gradient is produced by the GroupNorm backward that follows this conv
  const Act* src0 = nullptr;   // input tensor(s) of the layer (wired once the plan is final)
  const Act* src1 = nullptr;
  // stride-2 resampling convs on the tcgen05 engine (ResampleL holds the plans)
  int rs_kind = 0;             // 0: not a resample conv, 1: Conv2d k3 s2 p1, 2: ConvTranspose2d k4 s2 p1
  bool rs_ok = false;          // shapes eligible (bf16 weight buffers allocated)
  bool rs_f_valid = false, rs_b_valid = false;
};

struct BlockL {
  ConvL conv;
  int gn_w = -1, gn_b = -1;
  float* raw = nullptr;     // conv output (pre-norm), [M, Cout]
  float* stats = nullptr;   // [B][8][2] mean, rstd
};

struct ResnetL {
  std::string name;
  int Cin = 0, Cout = 0, H = 0, W = 0;
  int mlp_w = -1, mlp_b = -1, temb_off = 0;
  BlockL b1, b2;
  bool has_res = false;
  ConvL res;
  Act h1, out;
  float* r = nullptr;   // res_conv output
  const Act* in0 = nullptr;   // wired inputs (in1: skip tensor concatenated along channels)
  const Act* in1 = nullptr;
};

struct AttnL {
  std::string name;
  int C = 0, H = 0, W = 0;
  ConvL qkv, outc;
  int ln_g = -1, ln_b = -1;
  Act ln;                  // [M, C]   pre-norm output (input of to_qkv)
  float* qkv_t = nullptr;  // [M, 384]
  Act att;                 // [M, 128] attention output (input of to_out)
  const Act* in = nullptr;
  float* ctx = nullptr;    // [B, 4, 32, 32]
  float* kstat = nullptr;  // [B, 4, 32, 2]
  Act out;
  // tensor-core attention path (training, n >= 1024): q / k / v only exist as the bf16 hi / lo staging pair of the
  // to_qkv output, every per-pixel contraction is a per-image 1x1 conv on the tcgen05 engine (attention.cu, bottom)
  bool tcp_ok = false;
  __nv_bfloat16 *qkv_hi = nullptr, *qkv_lo = nullptr;   // [M, 384]
  __nv_bfloat16 *w1_hi = nullptr, *w1_lo = nullptr;     // [B][128][128] block-diagonal ctx (forward out = conv(q; ctx))
  TcConv tc_att_out, tc_att_dq, tc_att_dv, tc_att_t;
  // inference shortcut (n >= 2C): to_out runs with the per-image matrix M_b on LN(x); no attention output tensor
  bool mb_ok = false;
  __nv_bfloat16 *mb_hi = nullptr, *mb_lo = nullptr;   // [B][C][C]
  TcConv tc_mb;
  TcConv tc_kv;   // to_qkv restricted to its k | v rows (the per-image-matrix path of sampler steps never forms q)
};

struct ResampleL {
  std::string name;
  ConvL conv;
  int Hin = 0, Win = 0;
  Act out;
  bool present = false;
  const Act* in = nullptr;
  // tcgen05 plans: forward / data gradient are either one strided conv or four output-parity phases
  TcConv tcf[4], tcb[4];
  int n_tcf = 0, n_tcb = 0;
  TcWgrad tcw;
};

struct Stage {
  ResnetL r1, r2;
  AttnL attn;
  ResampleL rs;
};

// bump allocator with a measuring pass
struct Arena {
  float* base = nullptr;
  int64_t used = 0;   // floats
  float* alloc(int64_t n) {
    n = (n + 63) & ~int64_t(63);   // 256-byte granularity
    float* p = base ? base + used : nullptr;
    used += n;
    return p;
  }
};

}  // namespace
}  // namespace igm

using namespace igm;

struct igm_ctx {
  igm_unet_cfg cfg{};
  int device = 0;
  Status st;
  int64_t launches = 0;
  int conv_engine = 0;
  Profiler prof;

  std::vector<ParamInfo> params;
  int64_t param_elems = 0;
  float* P = nullptr;   // bound parameter arena (caller-owned)
  float* G = nullptr;   // bound gradient arena (caller-owned)

  // plan
  std::vector<int> dims;
  std::vector<Stage> downs, ups;
  ResnetL mid1, mid2;
  AttnL mid_attn;
  BlockL final_block;
  Act final_act;
  ConvL final_conv;
  float* final_wbwd = nullptr;   // [1][C][K]
  std::vector<TimeProj> proj_host;
  TimeProj* proj_dev = nullptr;
  int n_proj = 0, proj_total = 0;
  int tm_w1 = -1, tm_b1 = -1, tm_w2 = -1, tm_b2 = -1;

  // buffers
  float* arena = nullptr;
  int64_t arena_floats = 0;
  Act x_in;                 // NHWC network input
  float *t_emb = nullptr, *t_h1 = nullptr, *t_temb = nullptr, *t_act = nullptr, *t_proj = nullptr;
  float *t_dproj = nullptr, *t_ws = nullptr;
  float* gn_part = nullptr;
  float* attn_ws = nullptr;   // per-chunk partials of the linear-attention kernels
  // shared scratch of the tensor-core attention path (one attention block runs at a time): dO and softmax(k) staging
  // pairs [maxM, 128], the three backward weight matrices [B][128][128] and the normaliser term [B][128]
  __nv_bfloat16 *att_d_hi = nullptr, *att_d_lo = nullptr, *att_p_hi = nullptr, *att_p_lo = nullptr;
  __nv_bfloat16 *att_w_hi[3] = {nullptr, nullptr, nullptr}, *att_w_lo[3] = {nullptr, nullptr, nullptr};
  float* att_cc = nullptr;
  bool attn_tc = true;        // IGM_ATTN_TC=0: CUDA-core attention kernels everywhere
  int64_t attn_tc_min_pix = 65536;   // IGM_ATTN_TC_MIN=<pixels per launch> from which the tensor-core path is taken
  float *ws_group = nullptr, *ws_chan = nullptr, *ws_ln = nullptr;
  float *scrA = nullptr, *scrB = nullptr, *scrC = nullptr;
  // bf16x2 staging of an output gradient: a small ring, each conv layer owns one slot (ConvL::dyb, assigned in backward
  // order) so that a weight-gradient GEMM on the side stream can still read slot k while the main stream stages the next
  // layer's gradient into slot k+1
  static constexpr int kDyBufs = 3;
  __nv_bfloat16 *dy_hi[kDyBufs] = {nullptr, nullptr, nullptr}, *dy_lo[kDyBufs] = {nullptr, nullptr, nullptr};
  // side stream of the tensor-core weight-gradient kernels (IGM_WGRAD_STREAM=0: everything on the caller's stream)
  bool side_on = true;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_wr = nullptr, ev_join = nullptr, ev_rd[kDyBufs] = {nullptr, nullptr, nullptr};
  bool rd_pending[kDyBufs] = {false, false, false};
  bool side_dirty = false;
  // gradient buckets (igm_unet_grad_buckets): contiguous ranges of the gradient arena in the order the backward pass
  // completes them; ev_bk_main / ev_bk_side mark "every kernel that writes bucket k has been enqueued" on the caller's /
  // the side stream, so an all-reduce of bucket k can start while the backward pass continues (igm_unet_bucket_wait)
  static constexpr int kMaxBuckets = IGM_MAX_MULTS + 2;
  int n_buckets = 0;
  int64_t bucket_lo[kMaxBuckets] = {}, bucket_hi[kMaxBuckets] = {};
  int bucket_slab_lo[kMaxBuckets] = {}, bucket_slab_hi[kMaxBuckets] = {};
  cudaEvent_t ev_bk_main[kMaxBuckets] = {}, ev_bk_side[kMaxBuckets] = {};
  bool bk_side_rec[kMaxBuckets] = {};
  bool bk_valid = false;
  cudaEvent_t ev_ln = nullptr;   // side stream has folded the LayerNorm partials workspace (ws_ln may be overwritten)
  bool ln_pending = false;
  bool fin_per_layer = false;    // this backward folded the halo workspaces layer by layer (no finalize pass at the end)
  bool tc_available = false;
  HaloFinJob* fin_dev = nullptr;      // device job table of the halo-wgrad finalize pass
  int* fin_cta_dev = nullptr;         // CTA -> job
  int fin_cta_cap = 0;
  int fin_n = 0, fin_tiles = 0;
  double fin_elems = 0;
  bool halo_on = true;
  bool attn_mb = true;                // IGM_ATTN_MB=0: always materialise the attention output
  bool attn_kv_only = true;           // IGM_ATTN_KV=0: the per-image-matrix path still computes all of q | k | v
  PackJob* pack_dev = nullptr;        // device job table of igm_unet_pack_weights
  int pack_n = 0, pack_engine = -1;
  int64_t pack_total = 0;
  PackJob* packt_dev = nullptr;       // tiled re-pack (bf16 layouts): jobs, CTA -> job table
  int* packt_cta_dev = nullptr;
  int packt_ctas = 0, packt_taps = 1, packt_cap = 0;
  double packt_elems = 0;
  float* pred = nullptr;        // [B,C,H,W] network output (NCHW)
  float* noise_copy = nullptr;  // NCHW
  float* d_pred = nullptr;      // NHWC [M, C]
  float* loss_ws = nullptr;
  int64_t* t_vec = nullptr;     // sampler timesteps [B]
  int* loop_state = nullptr;    // [t, step]
  igm_schedule sched{};
  igm_schedule* sched_dev = nullptr;
  bool have_sched = false;

  std::map<std::string, Act> taps;

  std::vector<float*> skip_g;   // gradient of the skip tensor of down stage i (i >= 1), training only
  const Act* final_in = nullptr;   // input of final_conv.0

  int last_B = 0;
  bool fwd_valid = false, loss_valid = false;

  // sampler graph cache
  cudaGraphExec_t graph_exec = nullptr;
  struct GraphKey {
    float* img = nullptr; const float* noise = nullptr; uint64_t seed = 0; int B = 0, clip = 0;
    bool operator==(const GraphKey& o) const {
      return img == o.img && noise == o.noise && seed == o.seed && B == o.B && clip == o.clip;
    }
  } graph_key;

  LaunchCtx lc(void* stream) {
    LaunchCtx l;
    l.stream = (cudaStream_t)stream;
    l.st = &st;
    l.counter = &launches;
    l.prof = &prof;
    return l;
  }
  float* Pp(int idx) const { return idx < 0 ? nullptr : P + params[idx].offset; }
  float* Gp(int idx) const { return idx < 0 ? nullptr : G + params[idx].offset; }
};

namespace igm {
namespace {

// ---------------------------------------------------------------------------
// parameter table: same order/shapes as reference Unet(...).state_dict()
// ---------------------------------------------------------------------------
struct ParamBuilder {
  std::vector<ParamInfo>& out;
  int64_t off = 0;
  int add(const std::string& name, std::initializer_list<int64_t> shape) {
    ParamInfo p;
    p.name = name;
    p.offset = off;
    p.ndim = (int)shape.size();
    int i = 0;
    for (int64_t s : shape) p.shape[i++] = s;
    // keep every tensor 16-byte aligned inside the flat arena (float4 loads of bias / gamma)
    off += (p.numel() + 3) & ~int64_t(3);
    out.push_back(p);
    return (int)out.size() - 1;
  }
};

struct PlanBuilder {
  igm_ctx& c;
  ParamBuilder pb;
  Arena ar;
  bool training;
  int B;
  int64_t maxMC = 0, maxM = 0, maxGnWs = 0, maxDy = 0, halo_w_elems = 0, maxAttTc = 0;

  PlanBuilder(igm_ctx& ctx, float* base) : c(ctx), pb{ctx.params} {
    ar.base = base;
    training = ctx.cfg.training != 0;
    B = ctx.cfg.max_batch;
  }

  int64_t M(int H, int W) const { return (int64_t)B * H * W; }

  Act act(int C, int H, int W, bool grad, bool staged = true) {
    Act a;
    a.C = C; a.H = H; a.W = W;
    a.v = ar.alloc(M(H, W) * C);
    if (grad && training) a.g = ar.alloc(M(H, W) * C);
    if (staged && C % 64 == 0) {   // may feed a tensor-core conv: room for its bf16 hi/lo copy
      a.hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((M(H, W) * C + 1) / 2));
      a.lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((M(H, W) * C + 1) / 2));
    }
    return a;
  }
  // `lean`: in tensor-core mode the fp32 copy of this tensor may be skipped (only its bf16 hi/lo staging is
  // consumed); igm_debug_read_tap then rebuilds it from hi + lo
  void tap(const std::string& n, float* v, int C, int H, int W, const Act* lean = nullptr) {
    Act a; a.v = v; a.C = C; a.H = H; a.W = W;
    if (lean) { a.hi = lean->hi; a.lo = lean->lo; }
    c.taps[n] = a;
  }

  // H, W > 0 marks a stride-1 Conv2d running at that resolution (candidate for the tcgen05 engine)
  ConvL conv(const std::string& name, int Cin, int Cout, int K, bool bias, bool convT = false, int H = 0, int W = 0,
             bool need_dgrad = true) {
    ConvL l;
    l.Cin = Cin; l.Cout = Cout; l.K = K; l.convT = convT;
    if (convT) l.pw = pb.add(name + ".weight", {Cin, Cout, K, K});
    else l.pw = pb.add(name + ".weight", {Cout, Cin, K, K});
    if (bias) l.pb = pb.add(name + ".bias", {Cout});
    const int64_t nw = (int64_t)K * K * Cin * Cout;
    l.w_fwd = ar.alloc(nw);
    if (training) l.w_bwd = ar.alloc(nw);
    l.H = H; l.W = W;
    if (H > 0 && !convT) {
      l.tc_f_ok = tc_eligible(Cin, Cout, H, W, K);
      l.tc_b_ok = training && need_dgrad && tc_eligible(Cout, Cin, H, W, K);
      if (l.tc_f_ok) {
        l.wf_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
        l.wf_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
      }
      if (l.tc_b_ok) {
        l.wb_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
        l.wb_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
        maxDy = std::max(maxDy, M(H, W) * Cout);
      }
      l.tc_w_ok = training && tcw_eligible(Cin, Cout, H, W, K);
      if (l.tc_w_ok) maxDy = std::max(maxDy, M(H, W) * Cout);
      l.tc_wh_ok = training && tcwh_eligible(Cin, Cout, H, W, K);
      if (l.tc_wh_ok) halo_w_elems += nw;   // workspaces are carved from ONE contiguous block (build_plan)
    }
    return l;
  }

  BlockL block(const std::string& name, int Cin, int Cout, int H, int W) {
    BlockL b;
    b.conv = conv(name + ".block.0", Cin, Cout, 3, true, false, H, W);
    b.conv.bias_in_norm = true;
    b.gn_w = pb.add(name + ".block.1.weight", {Cout});
    b.gn_b = pb.add(name + ".block.1.bias", {Cout});
    b.raw = ar.alloc(M(H, W) * Cout);
    b.stats = ar.alloc((int64_t)B * kGroups * 2);
    tap(name + ".conv", b.raw, Cout, H, W);
    track(H, W, Cout);
    if (training) maxDy = std::max(maxDy, M(H, W) * Cout);   // the GroupNorm backward stages dy for every block
    return b;
  }
  // bf16 weight buffers + staging room of a stride-2 resampling conv eligible for the tcgen05 engine
  void rs_alloc(ResampleL& rs, int kind) {
    ConvL& l = rs.conv;
    l.rs_kind = kind;
    const int C = l.Cin;
    const int gh = kind == 1 ? rs.out.H : rs.Hin, gw = kind == 1 ? rs.out.W : rs.Win;   // coarse grid
    l.rs_ok = (l.Cin == l.Cout) && C % 64 == 0 && gw <= 64 && (kind == 2 || (rs.Hin % 2 == 0 && rs.Win % 2 == 0));
    if (!l.rs_ok) return;
    const int64_t nw = (int64_t)l.K * l.K * l.Cin * l.Cout;
    l.wf_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
    l.wf_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
    if (training) {
      l.wb_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
      l.wb_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((nw + 1) / 2));
      maxDy = std::max(maxDy, M(rs.out.H, rs.out.W) * l.Cout);
    }
    (void)gh;
  }
  void track(int H, int W, int C) {
    maxMC = std::max(maxMC, M(H, W) * C);
    maxM = std::max(maxM, M(H, W));
    const int64_t chunks = cdiv(H * W, kGnChunkMin);
    maxGnWs = std::max(maxGnWs, (int64_t)B * chunks * 4 * C);
  }

  ResnetL resnet(const std::string& name, int Cin, int Cout, int H, int W) {
    ResnetL r;
    r.name = name; r.Cin = Cin; r.Cout = Cout; r.H = H; r.W = W;
    r.mlp_w = pb.add(name + ".mlp.1.weight", {Cout, c.cfg.dim});
    r.mlp_b = pb.add(name + ".mlp.1.bias", {Cout});
    r.temb_off = c.proj_total;
    c.proj_total += Cout;
    TimeProj tp{};
    tp.cout = Cout; tp.offset = r.temb_off;
    c.proj_host.push_back(tp);   // pointers are filled at bind time
    r.b1 = block(name + ".block1", Cin, Cout, H, W);
    r.b2 = block(name + ".block2", Cout, Cout, H, W);
    r.has_res = Cin != Cout;
    if (r.has_res) {
      r.res = conv(name + ".res_conv", Cin, Cout, 1, true, false, H, W);
      r.r = ar.alloc(M(H, W) * Cout);
    }
    r.h1 = act(Cout, H, W, true);
    r.out = act(Cout, H, W, true);
    tap(name + ".h1", r.h1.v, Cout, H, W, &r.h1);
    tap(name + ".out", r.out.v, Cout, H, W);
    return r;
  }

  AttnL attn(const std::string& name, int C, int H, int W) {
    AttnL a;
    a.name = name; a.C = C; a.H = H; a.W = W;
    const int hd = kHeads * kDimHead;
    a.qkv = conv(name + ".fn.fn.to_qkv", C, 3 * hd, 1, false, false, H, W);
    a.outc = conv(name + ".fn.fn.to_out", hd, C, 1, true, false, H, W);
    a.ln_g = pb.add(name + ".fn.norm.g", {1, C, 1, 1});
    a.ln_b = pb.add(name + ".fn.norm.b", {1, C, 1, 1});
    a.ln = act(C, H, W, false);
    a.qkv_t = ar.alloc(M(H, W) * 3 * hd);
    a.att = act(hd, H, W, false);
    a.ctx = ar.alloc((int64_t)B * kHeads * kDimHead * kDimHead);
    a.kstat = ar.alloc((int64_t)B * kHeads * kDimHead * 2);
    a.out = act(C, H, W, true);
    a.mb_ok = H * W >= 8 * C   /* measured: only pays off when the image has many more pixels than channels */ && H * W >= 128 && C % 64 == 0 && tc_eligible(C, C, H, W, 1);
    a.tcp_ok = training && H * W >= 1024 && tc_eligible(hd, hd, H, W, 1) && tc_eligible(C, 3 * hd, H, W, 1);
    if (a.tcp_ok) {
      a.qkv_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc((M(H, W) * 3 * hd + 1) / 2));
      a.qkv_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc((M(H, W) * 3 * hd + 1) / 2));
      a.w1_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc(((int64_t)B * hd * hd + 1) / 2));
      a.w1_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc(((int64_t)B * hd * hd + 1) / 2));
      maxAttTc = std::max(maxAttTc, M(H, W));
    }
    if (a.mb_ok) {
      a.mb_hi = reinterpret_cast<__nv_bfloat16*>(ar.alloc(((int64_t)B * C * C + 1) / 2));
      a.mb_lo = reinterpret_cast<__nv_bfloat16*>(ar.alloc(((int64_t)B * C * C + 1) / 2));
    }
    tap(name + ".ln", a.ln.v, C, H, W, &a.ln);
    tap(name + ".out", a.out.v, C, H, W);
    track(H, W, C);
    maxMC = std::max(maxMC, M(H, W) * 3 * hd);
    return a;
  }

  void build() {
    const igm_unet_cfg& cfg = c.cfg;
    c.params.clear();
    c.taps.clear();
    c.proj_host.clear();
    c.proj_total = 0;
    c.dims.clear();
    c.dims.push_back(cfg.channels);
    for (int i = 0; i < cfg.n_mults; ++i) c.dims.push_back(cfg.dim * cfg.dim_mults[i]);
    const int nres = cfg.n_mults;
    const int d = cfg.dim;

    c.tm_w1 = pb.add("time_mlp.1.weight", {4 * d, d});
    c.tm_b1 = pb.add("time_mlp.1.bias", {4 * d});
    c.tm_w2 = pb.add("time_mlp.3.weight", {d, 4 * d});
    c.tm_b2 = pb.add("time_mlp.3.bias", {d});

    int H = cfg.height, W = cfg.width;
    c.x_in = act(cfg.channels, H, W, false, false);
    c.downs.assign(nres, Stage());
    std::vector<int> hs(nres), ws(nres);
    for (int i = 0; i < nres; ++i) {
      const int ci = c.dims[i], co = c.dims[i + 1];
      Stage& s = c.downs[i];
      const std::string p = "downs." + std::to_string(i);
      s.r1 = resnet(p + ".0", ci, co, H, W);
      s.r2 = resnet(p + ".1", co, co, H, W);
      s.attn = attn(p + ".2", co, H, W);
      hs[i] = H; ws[i] = W;
      if (i < nres - 1) {
        s.rs.present = true;
        s.rs.name = p + ".3";
        s.rs.conv = conv(p + ".3.conv", co, co, 3, true);
        s.rs.Hin = H; s.rs.Win = W;
        H = (H + 2 - 3) / 2 + 1;
        W = (W + 2 - 3) / 2 + 1;
        s.rs.out = act(co, H, W, true);
        tap(p + ".3.out", s.rs.out.v, co, H, W);
        track(H, W, co);
        rs_alloc(s.rs, 1);
      }
    }
    // ups are registered before the mid blocks in the reference (ddpm.py:195-196)
    c.ups.assign(nres - 1, Stage());
    {
      int Hu = H, Wu = W;
      for (int j = 0; j < nres - 1; ++j) {
        const int si = nres - 1 - j;              // reversed(in_out[1:])
        const int ci = c.dims[si], co = c.dims[si + 1];
        Stage& s = c.ups[j];
        const std::string p = "ups." + std::to_string(j);
        s.r1 = resnet(p + ".0", co * 2, ci, Hu, Wu);
        s.r2 = resnet(p + ".1", ci, ci, Hu, Wu);
        s.attn = attn(p + ".2", ci, Hu, Wu);
        s.rs.present = true;
        s.rs.name = p + ".3";
        s.rs.conv = conv(p + ".3.conv", ci, ci, 4, true, /*convT=*/true);
        s.rs.Hin = Hu; s.rs.Win = Wu;
        Hu *= 2; Wu *= 2;
        s.rs.out = act(ci, Hu, Wu, true);
        tap(p + ".3.out", s.rs.out.v, ci, Hu, Wu);
        track(Hu, Wu, ci);
        rs_alloc(s.rs, 2);
      }
    }
    const int mid = c.dims.back();
    c.mid1 = resnet("mid_block1", mid, mid, H, W);
    c.mid_attn = attn("mid_attn", mid, H, W);
    c.mid2 = resnet("mid_block2", mid, mid, H, W);

    // final_conv = Block(dims[1], dims[1]) + Conv1x1(dims[1], channels) at full resolution
    const int fH = cfg.height, fW = cfg.width, fc = c.dims[1];
    // (with an odd number of halvings the up path may not return to H x W; reject in create)
    c.final_block = block("final_conv.0", fc, fc, fH, fW);
    c.final_act = act(fc, fH, fW, true);
    tap("final_conv.0.out", c.final_act.v, fc, fH, fW);
    c.final_conv = ConvL();
    c.final_conv.Cin = fc; c.final_conv.Cout = cfg.channels; c.final_conv.K = 1;
    c.final_conv.pw = pb.add("final_conv.1.weight", {cfg.channels, fc, 1, 1});
    c.final_conv.pb = pb.add("final_conv.1.bias", {cfg.channels});
    if (training) c.final_wbwd = ar.alloc((int64_t)cfg.channels * fc);
    c.param_elems = pb.off;
    c.n_proj = (int)c.proj_host.size();

    // time path buffers
    c.t_emb = ar.alloc((int64_t)B * d);
    c.t_h1 = ar.alloc((int64_t)B * 4 * d);
    c.t_temb = ar.alloc((int64_t)B * d);
    c.t_act = ar.alloc((int64_t)B * d);
    c.t_proj = ar.alloc((int64_t)B * c.proj_total);
    tap("time_mlp", c.t_temb, d, 1, 1);
    const int64_t HW0 = (int64_t)cfg.height * cfg.width;
    // 32-pixel slots (statistics fused into the conv epilogue) or 64-pixel chunks
    c.gn_part = ar.alloc((int64_t)B * cdiv((int)HW0, 32) * kGroups * 2);
    c.attn_ws = ar.alloc(linattn_ws_floats(B, (int)HW0));
    c.pred = ar.alloc((int64_t)B * cfg.channels * HW0);
    c.loss_ws = ar.alloc(1024);
    c.t_vec = reinterpret_cast<int64_t*>(ar.alloc(2 * (int64_t)B + 4));
    c.loop_state = reinterpret_cast<int*>(ar.alloc(64));
    c.sched_dev = reinterpret_cast<igm_schedule*>(ar.alloc(64));
    c.proj_dev = reinterpret_cast<TimeProj*>(ar.alloc((int64_t)(sizeof(TimeProj) * c.n_proj + 3) / 4 + 64));
    c.pack_dev = reinterpret_cast<PackJob*>(ar.alloc((int64_t)(sizeof(PackJob) * 1024) / 4));
    c.packt_dev = reinterpret_cast<PackJob*>(ar.alloc((int64_t)(sizeof(PackJob) * 1024) / 4));
    c.packt_cap = (int)(2 * pb.off / 1024) + 1024;
    c.packt_cta_dev = reinterpret_cast<int*>(ar.alloc(c.packt_cap));
    c.fin_dev = reinterpret_cast<HaloFinJob*>(ar.alloc((int64_t)(sizeof(HaloFinJob) * 256) / 4));
    c.fin_cta_cap = (int)(halo_w_elems / 1024) + 64;
    c.fin_cta_dev = reinterpret_cast<int*>(ar.alloc(c.fin_cta_cap));
    if (training) {
      c.t_dproj = ar.alloc((int64_t)B * c.proj_total);
      c.t_ws = ar.alloc((int64_t)B * 10 * d);
      c.ws_group = ar.alloc((int64_t)B * cdiv((int)HW0, kGnChunkMin) * kGroups * 2);
      c.ws_chan = ar.alloc(maxGnWs);
      c.ws_ln = ar.alloc((int64_t)ln_backward_parts(maxM) * 2 * 1024);
      c.scrA = ar.alloc(maxMC);
      c.scrB = ar.alloc(maxM * kHeads * kDimHead);
      c.scrC = ar.alloc(maxM * 3 * kHeads * kDimHead);
      c.noise_copy = ar.alloc((int64_t)B * cfg.channels * HW0);
      c.d_pred = ar.alloc((int64_t)B * cfg.channels * HW0);
    }
    if (maxAttTc > 0) {
      const int64_t hd = kHeads * kDimHead;
      auto bf = [&](int64_t n) { return reinterpret_cast<__nv_bfloat16*>(ar.alloc((n + 1) / 2)); };
      c.att_d_hi = bf(maxAttTc * hd); c.att_d_lo = bf(maxAttTc * hd);
      c.att_p_hi = bf(maxAttTc * hd); c.att_p_lo = bf(maxAttTc * hd);
      for (int k = 0; k < 3; ++k) { c.att_w_hi[k] = bf((int64_t)B * hd * hd); c.att_w_lo[k] = bf((int64_t)B * hd * hd); }
      c.att_cc = ar.alloc((int64_t)B * hd);
    }
    if (maxDy > 0) {
      for (int k = 0; k < igm_ctx::kDyBufs; ++k) {
        c.dy_hi[k] = reinterpret_cast<__nv_bfloat16*>(ar.alloc((maxDy + 1) / 2));
        c.dy_lo[k] = reinterpret_cast<__nv_bfloat16*>(ar.alloc((maxDy + 1) / 2));
      }
    }
  }
};

// ---------------------------------------------------------------------------
// launch sequences
// ---------------------------------------------------------------------------
struct Runner {
  igm_ctx& c;
  LaunchCtx lc;
  int B;
  bool infer = false;   // no backward pass will read this forward's intermediates (sampler steps)

  // ---- side stream of the tensor-core weight-gradient kernels -------------------------------------------------
  // A weight gradient feeds nothing until the optimizer step, so it leaves the critical path: the kernel is enqueued on
  // c.side behind an event that marks "dY slot staged", and the main stream only waits for it (ev_rd[slot]) when it is
  // about to overwrite that slot of the staging ring, and once at the end of backward (ev_join).  Off while profiling
  // (per-scope event timing wants kernels alone) and on the SIMT engine.
  bool side_active() const { return c.side_on && c.side && tc_on() && !(c.prof.on); }
  __nv_bfloat16* dyh(const ConvL& l) const { return c.dy_hi[l.dyb]; }
  __nv_bfloat16* dyl(const ConvL& l) const { return c.dy_lo[l.dyb]; }
  // call before any kernel that writes slot l.dyb of the staging ring
  int before_dy_write(const ConvL& l) {
    if (c.rd_pending[l.dyb]) {
      IGM_CUDA(c.st, cudaStreamWaitEvent(lc.stream, c.ev_rd[l.dyb], 0));
      c.rd_pending[l.dyb] = false;
    }
    return IGM_OK;
  }
  // launch context for a weight-gradient kernel of layer l (reads slot l.dyb, already staged on the main stream)
  int side_begin(LaunchCtx& out) {
    out = lc;
    if (!side_active()) return IGM_OK;
    IGM_CUDA(c.st, cudaEventRecord(c.ev_wr, lc.stream));
    IGM_CUDA(c.st, cudaStreamWaitEvent(c.side, c.ev_wr, 0));
    out.stream = c.side;
    return IGM_OK;
  }
  int wgrad_begin(const ConvL& l, LaunchCtx& out) { (void)l; return side_begin(out); }
  int wgrad_end(const ConvL& l, const LaunchCtx& used) {
    if (used.stream == lc.stream) return IGM_OK;
    IGM_CUDA(c.st, cudaEventRecord(c.ev_rd[l.dyb], c.side));
    c.rd_pending[l.dyb] = true;
    c.side_dirty = true;
    return IGM_OK;
  }
  // everything the side stream was given is done before the main stream continues
  int side_join() {
    if (c.side_dirty) {
      IGM_CUDA(c.st, cudaEventRecord(c.ev_join, c.side));
      IGM_CUDA(c.st, cudaStreamWaitEvent(lc.stream, c.ev_join, 0));
      c.side_dirty = false;
      c.ln_pending = false;
      for (bool& b : c.rd_pending) b = false;
    }
    return IGM_OK;
  }

  // optional by-product of the NEXT conv_dgrad on the tensor-core engine: its output also (lean: only) as a bf16 hi / lo pair
  __nv_bfloat16 *dg_hi = nullptr, *dg_lo = nullptr;
  bool dg_lean = false;
  int64_t M(int H, int W) const { return (int64_t)B * H * W; }
  bool tc_on() const { return c.conv_engine == 1; }
  // tensor-core attention path for this block?  (training forward / backward only; the sampler keeps the M_b shortcut)
  bool attn_tcp(const AttnL& a) const {
    // measured (profiles/r2_attention.md): the path pays for itself from ~64 K pixels per launch on; below that its
    // twelve launches per block lose against the four of the CUDA-core kernels
    return c.attn_tc && a.tcp_ok && tc_on() && !infer && c.cfg.training && (int64_t)B * a.H * a.W >= c.attn_tc_min_pix &&
           a.tc_att_out.valid && a.qkv.tc_f.valid && a.qkv.tc_b.valid && tcw_batch_ok(a.qkv.tc_w, B) && a.outc.tc_b.valid;
  }
  bool use_tc(const TcConv& t) const { return tc_on() && t.valid; }
  bool use_pair(const TcConvPair& t) const { return tc_on() && conv_pair_on() && t.valid; }
  // bf16 staging pointers of an activation: only handed to producers while the tcgen05 engine is on
  __nv_bfloat16* hi(const Act& a) const { return tc_on() ? a.hi : nullptr; }
  __nv_bfloat16* lo(const Act& a) const { return tc_on() ? a.lo : nullptr; }
  // producers that cannot emit the staging copy themselves (SIMT convs) are followed by a split pass
  int stage_act(const Act& a) {
    if (!tc_on() || !a.hi) return IGM_OK;
    return launch_split_bf16(lc, a.v, M(a.H, a.W), a.C, a.hi, a.lo, a.C, 0);
  }

  // May the producer of `consumer`'s input skip the fp32 copy?  Yes when every reader (forward conv, and in
  // training its weight-gradient GEMM) runs on the tensor cores, which only read the bf16 hi/lo staging.
  bool lean_ok(const ConvL& consumer) const {
    return tc_on() && consumer.tc_f.valid && (!c.cfg.training || tcw_batch_ok(consumer.tc_w, B));
  }

  // forward of a layer conv reading its wired sources.  `out_act`: when the output is an activation
  // that later feeds tensor-core convs, its bf16 hi/lo copy is produced here as well.
  int conv_fwd(const ConvL& l, int IH, int IW, int OH, int OW, int stride, int pad, float* out, const float* add,
               const Act* out_act = nullptr, float* gn_part = nullptr) {
    const Act* s0 = l.src0;
    const Act* s1 = l.src1;
    if (stride == 1 && use_tc(l.tc_f)) {
      TcRun r;
      r.B = B; r.bias = c.Pp(l.pb); r.out0 = out; r.N0 = l.Cout; r.add0 = add; r.kclass = K_CONV_FPROP;
      r.gn_part = gn_part;
      if (out_act) { r.hi0 = hi(*out_act); r.lo0 = lo(*out_act); }
      if (use_pair(l.tcp_f)) return launch_conv_tc2(lc, l.tcp_f, r);
      return launch_conv_tc(lc, l.tc_f, r);
    }
    ConvArgs a;
    a.in0 = s0->v; a.C0 = s0->C;
    a.in1 = s1 ? s1->v : nullptr; a.C1 = s1 ? s1->C : 0;
    a.B = B; a.IH = IH; a.IW = IW; a.OH = OH; a.OW = OW;
    a.N = l.Cout; a.N0 = l.Cout;
    a.KH = a.KW = l.K; a.stride = stride; a.pad = pad; a.dil = 1;
    a.transposed = l.convT ? 1 : 0;
    a.w = l.w_fwd;
    a.bias = c.Pp(l.pb);
    a.out0 = out; a.add0 = add;
    IGM_TRY(launch_conv(lc, a));
    if (out_act) IGM_TRY(stage_act(*out_act));
    return IGM_OK;
  }
  // data gradient: d_out [B,OH,OW,Cout] -> d_in split (d0: C0 channels, d1: C1 channels)
  // dy_staged: the bf16 hi/lo copy of d_out already sits in the dy staging buffers
  int conv_dgrad(const ConvL& l, const float* d_out, int OH, int OW, int IH, int IW, int stride, int pad,
                 float* d0, int C0, float* d1, int C1, const float* add0, const float* add1,
                 bool dy_staged = false) {
    if (stride == 1 && use_tc(l.tc_b) && C0 % 32 == 0) {
      if (!dy_staged) {
        IGM_TRY(before_dy_write(l));
        IGM_TRY(launch_split_bf16(lc, d_out, M(OH, OW), l.Cout, dyh(l), dyl(l), l.Cout, 0));
      }
      TcRun r;
      r.B = B; r.bias = nullptr; r.out0 = d0; r.out1 = d1; r.N0 = C0; r.add0 = add0; r.add1 = add1;
      r.kclass = K_CONV_DGRAD;
      if (dg_hi && !d1) {
        r.hi0 = dg_hi; r.lo0 = dg_lo;
        if (dg_lean) r.out0 = nullptr;
        return launch_conv_tc(lc, l.tc_b, r);
      }
      if (use_pair(l.tcp_b)) return launch_conv_tc2(lc, l.tcp_b, r);
      return launch_conv_tc(lc, l.tc_b, r);
    }
    ConvArgs a;
    a.in0 = d_out; a.C0 = l.Cout; a.C1 = 0;
    a.B = B; a.IH = OH; a.IW = OW; a.OH = IH; a.OW = IW;
    a.N = C0 + C1; a.N0 = C0;
    a.KH = a.KW = l.K; a.stride = stride; a.pad = pad; a.dil = 1;
    a.transposed = l.convT ? 0 : 1;
    a.kclass = K_CONV_DGRAD;
    a.w = l.w_bwd;
    a.bias = nullptr;
    a.out0 = d0; a.out1 = d1; a.add0 = add0; a.add1 = add1;
    return launch_conv(lc, a);
  }
  // weight + bias gradients of a layer conv (forward inputs = its wired sources)
  int conv_wgrad(const ConvL& l, int IH, int IW, const float* d_out, int OH, int OW, int stride, int pad,
                 bool dy_staged = false, bool bias_done = false) {
    const int KK = l.K * l.K;
    float* gw = c.Gp(l.pw);
    const Act* s0 = l.src0;
    const Act* s1 = l.src1;
    if (stride == 1 && tc_on() && tcw_batch_ok(l.tc_w, B)) {
      // tensor-core path: X is already staged (forward), dY is staged by the caller or here
      if (!dy_staged) {
        IGM_TRY(before_dy_write(l));
        IGM_TRY(launch_split_bf16(lc, d_out, M(OH, OW), l.Cout, dyh(l), dyl(l), l.Cout, 0));
      }
      LaunchCtx wl;
      IGM_TRY(wgrad_begin(l, wl));
      if (c.halo_on && l.tc_wh.valid) {
        IGM_TRY(launch_wgrad_halo(wl, l.tc_wh, B));   // workspace -> OIHW gradient by the finalize pass:
        if (wl.stream != lc.stream) {                   // per layer, right behind it, when the side stream carries it
          IGM_TRY(launch_wgrad_halo_finalize(wl, c.fin_dev, c.fin_cta_dev, l.fin_ctas, 9.0 * l.Cin * l.Cout, l.fin_begin));
          c.fin_per_layer = true;
        }
      } else IGM_TRY(launch_wgrad_tc(wl, l.tc_w, B, gw));
      IGM_TRY(wgrad_end(l, wl));
    } else if (!l.convT) {
      // Conv2d: P = d_out (pc = co), Q = input (qc = ci) gathered at oy*s - p + ky;  W[co][ci][tap]
      const Act* srcs[2] = {s0, s1};
      int coff = 0;
      for (int i = 0; i < 2; ++i) {
        if (!srcs[i]) continue;
        WgradArgs w;
        w.P = d_out; w.PC = l.Cout; w.PH = OH; w.PW = OW;
        w.Q = srcs[i]->v; w.QC = srcs[i]->C; w.QH = IH; w.QW = IW;
        w.B = B; w.KH = w.KW = l.K; w.stride = stride; w.pad = pad; w.dil = 1;
        w.grad = gw + (int64_t)coff * KK;
        w.sq = KK; w.sp = (int64_t)l.Cin * KK;
        IGM_TRY(launch_wgrad(lc, w));
        coff += srcs[i]->C;
      }
    } else {
      // ConvTranspose2d: oy = iy*s - p + ky.  P = input (pc = ci), Q = d_out (qc = co);  W[ci][co][tap]
      WgradArgs w;
      w.P = s0->v; w.PC = l.Cin; w.PH = IH; w.PW = IW;
      w.Q = d_out; w.QC = l.Cout; w.QH = OH; w.QW = OW;
      w.B = B; w.KH = w.KW = l.K; w.stride = stride; w.pad = pad; w.dil = 1;
      w.grad = gw;
      w.sq = KK; w.sp = (int64_t)l.Cout * KK;
      IGM_TRY(launch_wgrad(lc, w));
    }
    if (l.pb >= 0 && !l.bias_in_norm && !bias_done) IGM_TRY(launch_colsum(lc, d_out, M(OH, OW), l.Cout, c.Gp(l.pb)));
    return IGM_OK;
  }
  // stage dY as bf16 hi/lo; when the layer has its own bias gradient the column sums ride along (sets bias_done)
  int stage_dy(const ConvL& l, const float* dY, int64_t m, bool& bias_done) {
    bias_done = false;
    IGM_TRY(before_dy_write(l));
    if (l.pb >= 0 && !l.bias_in_norm && split_colsum_ok(l.Cout)) {
      bias_done = true;
      return launch_split_bf16_colsum(lc, dY, m, l.Cout, dyh(l), dyl(l), c.Gp(l.pb));
    }
    return launch_split_bf16(lc, dY, m, l.Cout, dyh(l), dyl(l), l.Cout, 0);
  }

  // does conv_bwd(l, ..., d0) stage dY as bf16 hi/lo (tensor-core wgrad and/or dgrad)?
  bool bwd_stages_dy(const ConvL& l, bool want_dgrad) const {
    return tc_on() && (tcw_batch_ok(l.tc_w, B) || (want_dgrad && l.tc_b.valid && l.src0->C % 32 == 0));
  }
  // Backward of a stride-1 conv: weight/bias gradients, then (if d0) the data gradient.  dY is staged
  // once as bf16 hi/lo (by the producer when `staged`, else here) and shared by wgrad and dgrad.
  int conv_bwd(const ConvL& l, int H, int W, const float* dY, float* d0, float* d1, const float* add0,
               const float* add1, bool staged = false, bool bias_done = false) {
    const int pad = (l.K - 1) / 2;
    const int C0 = l.src0->C, C1 = l.src1 ? l.src1->C : 0;
    if (!staged && tc_on() && (tcw_batch_ok(l.tc_w, B) || (d0 && l.tc_b.valid && C0 % 32 == 0))) {
      IGM_TRY(stage_dy(l, dY, M(H, W), bias_done));
      staged = true;
    }
    IGM_TRY(conv_wgrad(l, H, W, dY, H, W, 1, pad, staged, bias_done));
    if (d0) IGM_TRY(conv_dgrad(l, dY, H, W, H, W, 1, pad, d0, C0, d1, C1, add0, add1, staged));
    return IGM_OK;
  }

  // ---- Block: conv3x3 -> GN -> Mish (+temb) (+res) ----
  int block_fwd(BlockL& b, int H, int W, const float* temb, const float* res, const Act& out, bool lean = false) {
    // GroupNorm partial statistics come out of the conv epilogue when the tensor-core plan allows it
    const bool fused = use_tc(b.conv.tc_f) && tc_gn_fusable(b.conv.tc_f, B);
    IGM_TRY(conv_fwd(b.conv, H, W, H, W, 1, 1, b.raw, nullptr, nullptr, fused ? c.gn_part : nullptr));
    int nparts = 0;
    if (fused) nparts = tc_gn_slots(b.conv.tc_f);
    else IGM_TRY(launch_gn_partial(lc, b.raw, B, H * W, b.conv.Cout, c.gn_part));
    IGM_TRY(launch_gn_apply(lc, b.raw, c.gn_part, c.Pp(b.gn_w), c.Pp(b.gn_b), temb, c.proj_total, res, lean ? nullptr : out.v,
                            b.stats, B, H * W, b.conv.Cout, hi(out), lo(out), nparts));
    return IGM_OK;
  }
  // d_out: grad of block output; leaves dy (grad of conv output) in scrA (+ its bf16 staging copy)
  // also_stage (optional): a conv whose dY is this block's d_out (the ResnetBlock's res_conv).  When the fused kernel
  // runs, it writes the bf16 hi/lo staging of d_out into that conv's ring slot and adds d_out's column sums to its bias
  // gradient on the way, and *also_done is set: the separate split (+ column-sum) pass over d_out is not needed.
  int block_bwd_norm(BlockL& b, const float* d_out, int H, int W, float* dtemb, const ConvL* also_stage = nullptr,
                     bool* also_done = nullptr) {
    GnBwdArgs g;
    g.d_out = d_out; g.y = b.raw; g.stats = b.stats;
    g.gamma = c.Pp(b.gn_w); g.beta = c.Pp(b.gn_b);
    // the fp32 copy of dy is only read by the SIMT fallbacks; the tcgen05 wgrad / dgrad read the bf16 hi/lo staging
    const bool tc_only = tc_on() && tcw_batch_ok(b.conv.tc_w, B) && b.conv.tc_b.valid && b.conv.src0->C % 32 == 0;
    g.dy = tc_only ? nullptr : c.scrA; g.dgamma = c.Gp(b.gn_w); g.dbeta = c.Gp(b.gn_b);
    g.dtemb = dtemb; g.dtemb_stride = c.proj_total;
    g.dbias = c.Gp(b.conv.pb);
    if (tc_on()) IGM_TRY(before_dy_write(b.conv));
    g.dy_hi = tc_on() ? dyh(b.conv) : nullptr; g.dy_lo = tc_on() ? dyl(b.conv) : nullptr;
    g.ws_group = c.ws_group; g.ws_chan = c.ws_chan;
    g.B = B; g.HW = H * W; g.C = b.conv.Cout;
    if (also_done) *also_done = false;
    if (also_stage && also_done && tc_on() && also_stage->Cout == g.C && gn_backward_is_fused(g)) {
      IGM_TRY(before_dy_write(*also_stage));
      g.dout_hi = dyh(*also_stage); g.dout_lo = dyl(*also_stage);
      if (also_stage->pb >= 0 && !also_stage->bias_in_norm) g.dout_colsum = c.Gp(also_stage->pb);
      *also_done = true;
    }
    return launch_gn_backward(lc, g);
  }
  bool dy_is_staged() const { return tc_on() && c.dy_hi[0] != nullptr; }

  int resnet_fwd(ResnetL& r) {
    const int H = r.H, W = r.W;
    IGM_TRY(block_fwd(r.b1, H, W, c.t_proj + r.temb_off, nullptr, r.h1, lean_ok(r.b2.conv)));
    const float* res = r.in0->v;
    if (r.has_res) {
      IGM_TRY(conv_fwd(r.res, H, W, H, W, 1, 0, r.r, nullptr));
      res = r.r;
    }
    IGM_TRY(block_fwd(r.b2, H, W, nullptr, res, r.out));
    return IGM_OK;
  }
  int maybe_time_backward(const ResnetL& r) {
    if (time_after != &r || time_done) return IGM_OK;
    LaunchCtx sl;
    IGM_TRY(side_begin(sl));
    IGM_TRY(time_backward(sl));
    if (sl.stream != lc.stream) c.side_dirty = true;
    time_done = true;
    return IGM_OK;
  }
  // The backward pass has left parameter group k (igm_unet_grad_buckets): run the group's time-projection parameter
  // gradients (their d(temb) are complete now) and mark the bucket's gradients as final on both streams.
  int bucket_done(int k) {
    if (k < 0 || k >= c.n_buckets) return IGM_OK;
    if (k + 1 < c.n_buckets) {   // the last group's slabs run with the time MLP itself (time_backward)
      TimeMlpParams tp;
      time_params(tp);
      LaunchCtx sl;
      IGM_TRY(side_begin(sl));
      IGM_TRY(launch_time_proj_backward(sl, tp, c.proj_dev, c.n_proj, c.proj_total, B, c.cfg.max_batch, c.t_act, c.t_dproj, c.t_ws,
                                        c.bucket_slab_lo[k], c.bucket_slab_hi[k]));
      if (sl.stream != lc.stream) c.side_dirty = true;
    }
    IGM_CUDA(c.st, cudaEventRecord(c.ev_bk_main[k], lc.stream));
    c.bk_side_rec[k] = false;
    if (side_active() && c.side_dirty) {
      IGM_CUDA(c.st, cudaEventRecord(c.ev_bk_side[k], c.side));
      c.bk_side_rec[k] = true;
    }
    return IGM_OK;
  }
  // d_out = r.out.g ; writes d_in0 / d_in1 unless null
  int resnet_bwd(ResnetL& r, float* d0, float* d1) {
    const int H = r.H, W = r.W;
    const float* d_out = r.out.g;
    // block2 (its fused GroupNorm backward also stages d_out for res_conv, which shares it)
    bool res_staged = false;
    const bool res_wants = r.has_res && bwd_stages_dy(r.res, d0 != nullptr);
    IGM_TRY(block_bwd_norm(r.b2, d_out, H, W, nullptr, res_wants ? &r.res : nullptr, &res_staged));
    IGM_TRY(conv_bwd(r.b2.conv, H, W, c.scrA, r.h1.g, nullptr, nullptr, nullptr, dy_is_staged()));
    // block1 (+ time-embedding add)
    if (r.has_res) {
      // res_conv first (its dY is d_out), then block1 accumulates on top of its data gradient
      IGM_TRY(conv_bwd(r.res, H, W, d_out, d0, d1, nullptr, nullptr, res_staged, res_staged));
      IGM_TRY(block_bwd_norm(r.b1, r.h1.g, H, W, c.t_dproj + r.temb_off));
      IGM_TRY(maybe_time_backward(r));
      IGM_TRY(conv_bwd(r.b1.conv, H, W, c.scrA, d0, d1, d0, d1, dy_is_staged()));
    } else {
      IGM_TRY(block_bwd_norm(r.b1, r.h1.g, H, W, c.t_dproj + r.temb_off));
      IGM_TRY(maybe_time_backward(r));
      IGM_TRY(conv_bwd(r.b1.conv, H, W, c.scrA, d0, nullptr, d_out, nullptr, dy_is_staged()));
    }
    return IGM_OK;
  }

  int attn_fwd(AttnL& a) {
    const int H = a.H, W = a.W;
    const int64_t m = M(H, W);
    const float* x = a.in->v;
    IGM_TRY(launch_ln_forward(lc, x, c.Pp(a.ln_g), c.Pp(a.ln_b), lean_ok(a.qkv) ? nullptr : a.ln.v, m, a.C, hi(a.ln), lo(a.ln)));
    if (attn_tcp(a)) {
      const int hd = kHeads * kDimHead;
      {   // to_qkv (no bias): only the bf16 hi / lo staging pair of q | k | v is written
        TcRun r;
        r.B = B; r.N0 = 3 * hd; r.hi0 = a.qkv_hi; r.lo0 = a.qkv_lo; r.kclass = K_CONV_FPROP;
        IGM_TRY(launch_conv_tc(lc, a.qkv.tc_f, r));
      }
      IGM_TRY(launch_linattn_ctx_hl(lc, a.qkv_hi, a.qkv_lo, a.ctx, a.kstat, B, H * W, c.attn_ws));
      IGM_TRY(launch_linattn_wt(lc, a.ctx, 0, B, a.w1_hi, a.w1_lo));
      {   // out = conv(q; ctx): [M, 128] x per-image block-diagonal 128 x 128 on the tensor cores
        TcRun r;
        r.B = B; r.N0 = hd; r.out0 = lean_ok(a.outc) ? nullptr : a.att.v; r.hi0 = hi(a.att); r.lo0 = lo(a.att);
        r.kclass = K_ATTN;
        IGM_TRY(launch_conv_tc(lc, a.tc_att_out, r));
      }
      IGM_TRY(conv_fwd(a.outc, H, W, H, W, 1, 0, a.out.v, x, &a.out));
      return IGM_OK;
    }
    if (c.attn_mb && (infer || !c.cfg.training) && tc_on() && a.tc_mb.valid && use_tc(a.qkv.tc_f)) {
      // inference, n >= 8C: y = (W_out ctx^T W_q) LN(x) + b + x with one C x C matrix per image; q itself is never
      // needed, so to_qkv only produces its k | v rows ([M, 256], two thirds of the output traffic that bounds this conv)
      const bool kv_only = c.attn_kv_only && use_tc(a.tc_kv);
      if (kv_only) {
        TcRun r;
        r.B = B; r.out0 = a.qkv_t; r.N0 = 2 * kHeads * kDimHead; r.kclass = K_CONV_FPROP;
        IGM_TRY(launch_conv_tc(lc, a.tc_kv, r));
      } else {
        IGM_TRY(conv_fwd(a.qkv, H, W, H, W, 1, 0, a.qkv_t, nullptr));
      }
      IGM_TRY(launch_linattn_ctx(lc, a.qkv_t, a.ctx, a.kstat, B, H * W, c.attn_ws, kv_only));
      IGM_TRY(launch_linattn_mb(lc, a.ctx, c.Pp(a.outc.pw), c.Pp(a.qkv.pw), B, a.C, a.mb_hi, a.mb_lo));
      TcRun r;
      r.B = B; r.bias = c.Pp(a.outc.pb); r.out0 = a.out.v; r.N0 = a.C; r.add0 = x; r.kclass = K_CONV_FPROP;
      r.hi0 = hi(a.out); r.lo0 = lo(a.out);
      return launch_conv_tc(lc, a.tc_mb, r);
    }
    IGM_TRY(conv_fwd(a.qkv, H, W, H, W, 1, 0, a.qkv_t, nullptr));
    IGM_TRY(launch_linattn_forward(lc, a.qkv_t, lean_ok(a.outc) ? nullptr : a.att.v, a.ctx, a.kstat, B, H * W, c.attn_ws,
                                   hi(a.att), lo(a.att)));
    IGM_TRY(conv_fwd(a.outc, H, W, H, W, 1, 0, a.out.v, x, &a.out));
    return IGM_OK;
  }
  int attn_bwd(AttnL& a, float* dx) {
    const int H = a.H, W = a.W;
    const int64_t m = M(H, W);
    const float* d_out = a.out.g;
    if (attn_tcp(a)) {
      const int hd = kHeads * kDimHead, n = H * W;
      // to_out backward: its data gradient dO leaves only as a staging pair
      dg_hi = c.att_d_hi; dg_lo = c.att_d_lo; dg_lean = true;
      const int rc0 = conv_bwd(a.outc, H, W, d_out, c.scrB, nullptr, nullptr, nullptr);
      dg_hi = dg_lo = nullptr; dg_lean = false;
      IGM_TRY(rc0);
      IGM_TRY(launch_linattn_dctx_hl(lc, a.qkv_hi, a.qkv_lo, c.att_d_hi, c.att_d_lo, B, n, c.attn_ws));
      const float* dctx = linattn_dctx_ptr(c.attn_ws);
      IGM_TRY(launch_linattn_wt(lc, a.ctx, 1, B, c.att_w_hi[0], c.att_w_lo[0]));                       // dq = conv(dO; ctx^T)
      IGM_TRY(launch_linattn_wt(lc, dctx, 0, B, c.att_w_hi[1], c.att_w_lo[1], a.ctx, c.att_cc));       // dv = conv(p; dctx), c
      IGM_TRY(launch_linattn_wt(lc, dctx, 1, B, c.att_w_hi[2], c.att_w_lo[2]));                        // T = conv(v; dctx^T)
      IGM_TRY(launch_linattn_p(lc, a.qkv_hi, a.qkv_lo, a.kstat, c.att_p_hi, c.att_p_lo, B, n));
      IGM_TRY(before_dy_write(a.qkv));
      __nv_bfloat16 *dh = dyh(a.qkv), *dl = dyl(a.qkv);
      TcRun r;
      r.B = B; r.N0 = hd; r.ld_hi = 3 * hd; r.kclass = K_ATTN;
      r.hi0 = dh; r.lo0 = dl;
      IGM_TRY(launch_conv_tc(lc, a.tc_att_dq, r));
      r.hi0 = dh + 2 * hd; r.lo0 = dl + 2 * hd;
      IGM_TRY(launch_conv_tc(lc, a.tc_att_dv, r));
      TcRun rt;
      rt.B = B; rt.N0 = hd; rt.out0 = c.scrB; rt.kclass = K_ATTN;
      IGM_TRY(launch_conv_tc(lc, a.tc_att_t, rt));
      IGM_TRY(launch_linattn_dk(lc, c.att_p_hi, c.att_p_lo, c.scrB, c.att_cc, dh, dl, B, n));
      IGM_TRY(conv_bwd(a.qkv, H, W, c.scrC, c.scrA, nullptr, nullptr, nullptr, /*staged=*/true));
    } else {
    IGM_TRY(conv_bwd(a.outc, H, W, d_out, c.scrB, nullptr, nullptr, nullptr));
    // to_qkv has no bias: when both of its backward convs run on the tensor cores they only read the bf16 hi/lo
    // staging copy of d(qkv), which the attention backward then writes directly (no fp32 tensor, no split pass)
    const bool direct = tc_on() && tcw_batch_ok(a.qkv.tc_w, B) && a.qkv.tc_b.valid && a.qkv.pb < 0;
    if (direct) IGM_TRY(before_dy_write(a.qkv));
    IGM_TRY(launch_linattn_backward(lc, a.qkv_t, a.ctx, a.kstat, c.scrB, direct ? nullptr : c.scrC, B, H * W, c.attn_ws,
                                    direct ? dyh(a.qkv) : nullptr, direct ? dyl(a.qkv) : nullptr));
    IGM_TRY(conv_bwd(a.qkv, H, W, c.scrC, c.scrA, nullptr, nullptr, nullptr, direct));
    }
    if (!side_active())
      return launch_ln_backward(lc, c.scrA, a.in->v, c.Pp(a.ln_g), d_out, dx, c.Gp(a.ln_g), c.Gp(a.ln_b), c.ws_ln, m, a.C);
    // the fold of the per-CTA partials into d(gamma), d(beta) feeds nothing downstream: side stream; the next LayerNorm
    // backward waits for it before it overwrites the partials workspace
    if (c.ln_pending) {
      IGM_CUDA(c.st, cudaStreamWaitEvent(lc.stream, c.ev_ln, 0));
      c.ln_pending = false;
    }
    IGM_TRY(launch_ln_backward(lc, c.scrA, a.in->v, c.Pp(a.ln_g), d_out, dx, c.Gp(a.ln_g), c.Gp(a.ln_b), c.ws_ln, m, a.C, false));
    LaunchCtx sl;
    IGM_TRY(side_begin(sl));
    IGM_TRY(launch_ln_param_finalize(sl, c.ws_ln, m, a.C, c.Gp(a.ln_g), c.Gp(a.ln_b)));
    IGM_CUDA(c.st, cudaEventRecord(c.ev_ln, c.side));
    c.ln_pending = true;
    c.side_dirty = true;
    return IGM_OK;
  }

  int time_params(TimeMlpParams& p) {
    p.w1 = c.Pp(c.tm_w1); p.b1 = c.Pp(c.tm_b1); p.w2 = c.Pp(c.tm_w2); p.b2 = c.Pp(c.tm_b2);
    p.gw1 = c.G ? c.Gp(c.tm_w1) : nullptr; p.gb1 = c.G ? c.Gp(c.tm_b1) : nullptr;
    p.gw2 = c.G ? c.Gp(c.tm_w2) : nullptr; p.gb2 = c.G ? c.Gp(c.tm_b2) : nullptr;
    p.dim = c.cfg.dim;
    return IGM_OK;
  }

  int resample_fwd(ResampleL& rs) {
    if (tc_on() && rs.n_tcf > 0) {
      TcRun r;
      r.B = B; r.bias = c.Pp(rs.conv.pb); r.out0 = rs.out.v; r.N0 = rs.conv.Cout; r.kclass = K_CONV_FPROP;
      r.hi0 = hi(rs.out); r.lo0 = lo(rs.out);
      for (int i = 0; i < rs.n_tcf; ++i) IGM_TRY(launch_conv_tc(lc, rs.tcf[i], r));
      return IGM_OK;
    }
    return conv_fwd(rs.conv, rs.Hin, rs.Win, rs.out.H, rs.out.W, 2, 1, rs.out.v, nullptr, &rs.out);
  }
  // weight/bias gradients + data gradient (into rs.in->g, plus an optional addend) of a resample conv
  int resample_bwd(ResampleL& rs, const float* add) {
    const ConvL& l = rs.conv;
    const bool tcw = tc_on() && tcw_batch_ok(rs.tcw, B);
    const bool tcb = tc_on() && rs.n_tcb > 0;
    bool bias_done = false;
    if (tcw || tcb)   // stage dY once for both tensor-core kernels
      IGM_TRY(stage_dy(l, rs.out.g, M(rs.out.H, rs.out.W), bias_done));
    if (tcw) {
      LaunchCtx wl;
      IGM_TRY(wgrad_begin(l, wl));
      IGM_TRY(launch_wgrad_tc(wl, rs.tcw, B, c.Gp(l.pw)));
      IGM_TRY(wgrad_end(l, wl));
      if (l.pb >= 0 && !bias_done) IGM_TRY(launch_colsum(lc, rs.out.g, M(rs.out.H, rs.out.W), l.Cout, c.Gp(l.pb)));
    } else {
      IGM_TRY(conv_wgrad(l, rs.Hin, rs.Win, rs.out.g, rs.out.H, rs.out.W, 2, 1, false, bias_done));
    }
    if (tcb) {
      TcRun r;
      r.B = B; r.out0 = rs.in->g; r.N0 = rs.in->C; r.add0 = add; r.kclass = K_CONV_DGRAD;
      for (int i = 0; i < rs.n_tcb; ++i) IGM_TRY(launch_conv_tc(lc, rs.tcb[i], r));
      return IGM_OK;
    }
    return conv_dgrad(l, rs.out.g, rs.out.H, rs.out.W, rs.Hin, rs.Win, 2, 1, rs.in->g, rs.in->C, nullptr, 0, add,
                      nullptr);
  }

  // whole network on c.x_in -> pred_nchw   (topology: wire_plan())
  int forward(const int64_t* t, float* pred_nchw) {
    const igm_unet_cfg& cfg = c.cfg;
    TimeMlpParams tp;
    time_params(tp);
    IGM_TRY(launch_time_mlp_forward(lc, tp, t, B, c.t_emb, c.t_h1, c.t_temb, c.t_act));
    IGM_TRY(launch_time_proj_forward(lc, c.proj_dev, c.n_proj, c.t_act, cfg.dim, B, c.proj_total, c.t_proj));
    auto stage = [&](Stage& s) -> int {
      IGM_TRY(resnet_fwd(s.r1));
      IGM_TRY(resnet_fwd(s.r2));
      IGM_TRY(attn_fwd(s.attn));
      if (s.rs.present) IGM_TRY(resample_fwd(s.rs));
      return IGM_OK;
    };
    for (auto& s : c.downs) IGM_TRY(stage(s));
    IGM_TRY(resnet_fwd(c.mid1));
    IGM_TRY(attn_fwd(c.mid_attn));
    IGM_TRY(resnet_fwd(c.mid2));
    for (auto& s : c.ups) IGM_TRY(stage(s));
    IGM_TRY(block_fwd(c.final_block, cfg.height, cfg.width, nullptr, nullptr, c.final_act));
    IGM_TRY(launch_final_conv(lc, c.final_act.v, c.Pp(c.final_conv.pw), c.Pp(c.final_conv.pb), pred_nchw, B,
                              cfg.height * cfg.width, c.final_conv.Cin, cfg.channels));
    return IGM_OK;
  }

  // d_pred: NHWC [M, channels]; d_x_nhwc may be null
  // time-embedding MLP backward: needs every block's d(temb), i.e. may start once block1 of the first ResnetBlock has run
  // its GroupNorm backward; it then overlaps that block's (SIMT, few-channel) stem weight gradients on the side stream
  const ResnetL* time_after = nullptr;
  bool time_done = false;
  int time_backward(const LaunchCtx& l) {
    TimeMlpParams tp;
    time_params(tp);
    // the earlier groups' projection slabs ran in bucket_done(); the last group's (time_mlp + downs.0) and the MLP here
    const int last = c.n_buckets - 1;
    IGM_TRY(launch_time_proj_backward(l, tp, c.proj_dev, c.n_proj, c.proj_total, B, c.cfg.max_batch, c.t_act, c.t_dproj, c.t_ws,
                                      c.bucket_slab_lo[last], c.bucket_slab_hi[last]));
    return launch_time_mlp_backward(l, tp, B, c.cfg.max_batch, c.t_emb, c.t_h1, c.t_temb, c.t_ws);
  }

  int backward(const float* d_pred, float* d_x_nhwc) {
    const igm_unet_cfg& cfg = c.cfg;
    const int nres = cfg.n_mults;
    const bool side_was_active = side_active();
    time_after = side_was_active ? &c.downs[0].r1 : nullptr;
    time_done = false;
    c.bk_valid = false;
    const int H0 = cfg.height, W0 = cfg.width;
    // final 1x1: W[c][k]
    {
      WgradArgs w;
      w.P = d_pred; w.PC = cfg.channels; w.PH = H0; w.PW = W0;
      w.Q = c.final_act.v; w.QC = c.final_conv.Cin; w.QH = H0; w.QW = W0;
      w.B = B; w.grad = c.Gp(c.final_conv.pw);
      w.sq = 1; w.sp = c.final_conv.Cin;
      LaunchCtx sl;                      // parameter gradients only: off the critical path
      IGM_TRY(side_begin(sl));
      IGM_TRY(launch_wgrad(sl, w));
      IGM_TRY(launch_colsum(sl, d_pred, M(H0, W0), cfg.channels, c.Gp(c.final_conv.pb)));
      if (sl.stream != lc.stream) c.side_dirty = true;
      ConvArgs a;
      a.in0 = d_pred; a.C0 = cfg.channels; a.B = B; a.IH = a.OH = H0; a.IW = a.OW = W0;
      a.N = a.N0 = c.final_conv.Cin; a.transposed = 1; a.kclass = K_CONV_DGRAD;
      a.w = c.final_wbwd; a.out0 = c.final_act.g;
      IGM_TRY(launch_conv(lc, a));
    }
    IGM_TRY(block_bwd_norm(c.final_block, c.final_act.g, H0, W0, nullptr));
    IGM_TRY(conv_bwd(c.final_block.conv, H0, W0, c.scrA, c.final_in->g, nullptr, nullptr, nullptr, dy_is_staged()));

    for (int j = nres - 2; j >= 0; --j) {
      Stage& s = c.ups[j];
      IGM_TRY(resample_bwd(s.rs, nullptr));                      // Upsample (ConvTranspose2d 4,2,1)
      IGM_TRY(attn_bwd(s.attn, s.r2.out.g));
      IGM_TRY(resnet_bwd(s.r2, s.r1.out.g, nullptr));
      IGM_TRY(resnet_bwd(s.r1, s.r1.in0->g, c.skip_g[nres - 1 - j]));   // concat: (x, skip)
    }
    IGM_TRY(resnet_bwd(c.mid2, c.mid_attn.out.g, nullptr));
    IGM_TRY(attn_bwd(c.mid_attn, c.mid1.out.g));
    {
      Stage& last = c.downs[nres - 1];
      IGM_TRY(resnet_bwd(c.mid1, last.attn.out.g, nullptr));
      if (nres > 1) IGM_TRY(launch_add(lc, last.attn.out.g, c.skip_g[nres - 1], M(last.attn.H, last.attn.W) * last.attn.C));
    }
    IGM_TRY(bucket_done(0));                                     // ups.*, mid_*, final_conv.*
    for (int i = nres - 1; i >= 0; --i) {
      Stage& s = c.downs[i];
      // h[0] is never consumed by the up path (ddpm.py:254-259): no skip gradient for stage 0
      if (s.rs.present) IGM_TRY(resample_bwd(s.rs, (i >= 1) ? c.skip_g[i] : nullptr));
      IGM_TRY(attn_bwd(s.attn, s.r2.out.g));
      IGM_TRY(resnet_bwd(s.r2, s.r1.out.g, nullptr));
      IGM_TRY(resnet_bwd(s.r1, i > 0 ? s.r1.in0->g : d_x_nhwc, nullptr));
      if (i > 0) IGM_TRY(bucket_done(nres - i));                 // downs.i
    }
    if (!time_done) IGM_TRY(time_backward(lc));
    time_done = false;
    time_after = nullptr;
    IGM_TRY(side_join());
    const bool fin_done = c.fin_per_layer;   // every halo layer was folded right behind its wgrad on the side stream
    c.fin_per_layer = false;
    if (tc_on() && c.halo_on && !fin_done) IGM_TRY(launch_wgrad_halo_finalize(lc, c.fin_dev, c.fin_cta_dev, c.fin_n > 0 ? c.fin_tiles : 0, c.fin_elems));
    // the last bucket (time_mlp.*, downs.0.*) is final here; so is every bucket when the side stream was off (the halo
    // workspaces were folded by the pass above, not layer by layer)
    for (int k = side_was_active ? c.n_buckets - 1 : 0; k < c.n_buckets; ++k) {
      IGM_CUDA(c.st, cudaEventRecord(c.ev_bk_main[k], lc.stream));
      c.bk_side_rec[k] = false;
    }
    c.bk_valid = true;
    return IGM_OK;
  }
};

}  // namespace
}  // namespace igm

template <class F>
static int for_each_conv(igm_ctx* c, F f) {
  auto rn = [&](ResnetL& r) -> int {
    IGM_TRY(f(r.b1.conv));
    IGM_TRY(f(r.b2.conv));
    if (r.has_res) IGM_TRY(f(r.res));
    return IGM_OK;
  };
  auto at = [&](AttnL& a) -> int {
    IGM_TRY(f(a.qkv));
    IGM_TRY(f(a.outc));
    return IGM_OK;
  };
  auto stg = [&](Stage& s) -> int {
    IGM_TRY(rn(s.r1));
    IGM_TRY(rn(s.r2));
    IGM_TRY(at(s.attn));
    if (s.rs.present) IGM_TRY(f(s.rs.conv));
    return IGM_OK;
  };
  for (auto& s : c->downs) IGM_TRY(stg(s));
  for (auto& s : c->ups) IGM_TRY(stg(s));
  IGM_TRY(rn(c->mid1));
  IGM_TRY(at(c->mid_attn));
  IGM_TRY(rn(c->mid2));
  IGM_TRY(f(c->final_block.conv));
  return IGM_OK;
}

// Topology of Unet.forward (reference ddpm.py:238-261): which activation feeds which layer.
static void wire_plan(igm_ctx* c) {
  const int nres = c->cfg.n_mults;
  auto wire_resnet = [](ResnetL& r, const Act* in0, const Act* in1) {
    r.in0 = in0; r.in1 = in1;
    r.b1.conv.src0 = in0; r.b1.conv.src1 = in1;
    r.b2.conv.src0 = &r.h1; r.b2.conv.src1 = nullptr;
    r.res.src0 = in0; r.res.src1 = in1;
  };
  auto wire_attn = [](AttnL& a, const Act* in) {
    a.in = in;
    a.qkv.src0 = &a.ln; a.outc.src0 = &a.att;
  };
  const Act* cur = &c->x_in;
  for (int i = 0; i < nres; ++i) {
    Stage& s = c->downs[i];
    wire_resnet(s.r1, cur, nullptr);
    wire_resnet(s.r2, &s.r1.out, nullptr);
    wire_attn(s.attn, &s.r2.out);
    cur = &s.attn.out;
    if (s.rs.present) { s.rs.in = cur; s.rs.conv.src0 = cur; cur = &s.rs.out; }
  }
  wire_resnet(c->mid1, cur, nullptr);
  wire_attn(c->mid_attn, &c->mid1.out);
  wire_resnet(c->mid2, &c->mid_attn.out, nullptr);
  cur = &c->mid2.out;
  for (int j = 0; j < nres - 1; ++j) {
    Stage& s = c->ups[j];
    wire_resnet(s.r1, cur, &c->downs[nres - 1 - j].attn.out);   // torch.cat((x, h.pop()), dim=1)  (:255)
    wire_resnet(s.r2, &s.r1.out, nullptr);
    wire_attn(s.attn, &s.r2.out);
    s.rs.in = &s.attn.out; s.rs.conv.src0 = &s.attn.out;
    cur = &s.rs.out;
  }
  c->final_in = cur;
  c->final_block.conv.src0 = cur;
}

// Gradient buckets in the order Runner::backward completes them.  The arena is laid out in state_dict order
// (time_mlp | downs.0 .. downs.n-1 | ups.* | mid_* | final_conv.*), the backward pass walks final_conv, ups, mid, then
// downs.n-1 .. downs.0 and the time MLP: bucket 0 = [ups.0 (or mid_block1) .. end), bucket k = downs.(n-k), last =
// [0 .. downs.1) = time_mlp + downs.0.
static void plan_buckets(igm_ctx* c) {
  const int nres = c->cfg.n_mults;
  auto poff = [&](int idx) { return c->params[idx].offset; };
  const ResnetL& head0 = c->ups.empty() ? c->mid1 : c->ups[0].r1;
  int k = 0;
  c->bucket_lo[k] = poff(head0.mlp_w); c->bucket_hi[k] = c->param_elems;
  c->bucket_slab_lo[k] = head0.temb_off / 32; c->bucket_slab_hi[k] = c->proj_total / 32;
  ++k;
  for (int i = nres - 1; i >= 1; --i, ++k) {
    c->bucket_lo[k] = poff(c->downs[i].r1.mlp_w); c->bucket_hi[k] = c->bucket_lo[k - 1];
    c->bucket_slab_lo[k] = c->downs[i].r1.temb_off / 32;
    c->bucket_slab_hi[k] = (c->downs[i].r2.temb_off + c->downs[i].r2.Cout) / 32;
  }
  c->bucket_lo[k] = 0; c->bucket_hi[k] = c->bucket_lo[k - 1];
  c->bucket_slab_lo[k] = 0; c->bucket_slab_hi[k] = (c->downs[0].r2.temb_off + c->downs[0].r2.Cout) / 32;
  c->n_buckets = k + 1;
}

// Slots of the dY staging ring, handed out round-robin in the order Runner::backward visits the convs, so that
// consecutive layers of the backward pass never share a slot.
static void assign_dy_slots(igm_ctx* c) {
  int k = 0;
  auto f = [&](ConvL& l) { l.dyb = (k++) % igm_ctx::kDyBufs; };
  auto rn = [&](ResnetL& r) { f(r.b2.conv); if (r.has_res) f(r.res); f(r.b1.conv); };
  auto at = [&](AttnL& a) { f(a.outc); f(a.qkv); };
  auto stg = [&](Stage& s) { if (s.rs.present) f(s.rs.conv); at(s.attn); rn(s.r2); rn(s.r1); };
  f(c->final_block.conv);
  for (int j = (int)c->ups.size() - 1; j >= 0; --j) stg(c->ups[j]);
  rn(c->mid2); at(c->mid_attn); rn(c->mid1);
  for (int i = (int)c->downs.size() - 1; i >= 0; --i) stg(c->downs[i]);
}

// TMA descriptors of the tcgen05 engine for every eligible stride-1 conv (needs a live driver).
// Operands are the producers' bf16 hi/lo staging copies of the wired source activations.
static int plan_tc(igm_ctx* c) {
  int n_valid = 0;
  int rc = for_each_conv(c, [&](ConvL& l) -> int {
    const Act* s0 = l.src0;
    const Act* s1 = l.src1;
    const bool staged = s0 && s0->hi && (!s1 || s1->hi);
    const int pad = (l.K - 1) / 2;
    if (l.tc_f_ok && staged)
      IGM_TRY(tc_plan(c->st, l.tc_f, l.Cin, l.Cout, l.H, l.W, c->cfg.max_batch, l.K, pad, s0->hi, s0->lo, l.wf_hi,
                      l.wf_lo, s0->C, s1 ? s1->hi : nullptr, s1 ? s1->lo : nullptr));
    if (l.tc_b_ok)
      IGM_TRY(tc_plan(c->st, l.tc_b, l.Cout, l.Cin, l.H, l.W, c->cfg.max_batch, l.K, pad, c->dy_hi[l.dyb], c->dy_lo[l.dyb], l.wb_hi,
                      l.wb_lo));
    if (conv_pair_on()) {   // CTA-pair plans reference tc_f / tc_b (ConvL objects do not move after planning)
      if (l.tc_f.valid && tc2_eligible(l.tc_f)) IGM_TRY(tc2_plan(c->st, l.tcp_f, l.tc_f));
      if (l.tc_b.valid && tc2_eligible(l.tc_b)) IGM_TRY(tc2_plan(c->st, l.tcp_b, l.tc_b));
    }
    if (l.tc_w_ok && staged)
      IGM_TRY(tcw_plan(c->st, l.tc_w, l.Cin, l.Cout, l.H, l.W, c->cfg.max_batch, l.K, pad, c->dy_hi[l.dyb], c->dy_lo[l.dyb], s0->hi,
                       s0->lo, s0->C, s1 ? s1->hi : nullptr, s1 ? s1->lo : nullptr));
    if (l.tc_wh_ok && l.tc_w_ok && staged)
      IGM_TRY(tcwh_plan(c->st, l.tc_wh, l.Cin, l.Cout, l.H, l.W, c->cfg.max_batch, c->dy_hi[l.dyb], c->dy_lo[l.dyb], s0->hi, s0->lo, s0->C,
                        s1 ? s1->hi : nullptr, s1 ? s1->lo : nullptr, l.wg_ws));
    n_valid += (l.tc_f.valid ? 1 : 0) + (l.tc_b.valid ? 1 : 0) + (l.tc_w.valid ? 1 : 0);
    return IGM_OK;
  });
  // per-image to_out convs of the inference attention shortcut, and the per-image 1x1 convs of the tensor-core
  // attention path (operands: thirds of the [M, 384] qkv staging pair, pitch 384)
  auto plan_mb = [&](AttnL& a) -> int {
    if (a.tcp_ok && a.qkv_hi && c->att_d_hi) {
      const int hd = kHeads * kDimHead, Bm = c->cfg.max_batch;
      IGM_TRY(tc_plan_img(c->st, a.tc_att_out, hd, hd, a.H, a.W, Bm, a.qkv_hi, a.qkv_lo, a.w1_hi, a.w1_lo, 3 * hd));
      IGM_TRY(tc_plan_img(c->st, a.tc_att_dq, hd, hd, a.H, a.W, Bm, c->att_d_hi, c->att_d_lo, c->att_w_hi[0], c->att_w_lo[0]));
      IGM_TRY(tc_plan_img(c->st, a.tc_att_dv, hd, hd, a.H, a.W, Bm, c->att_p_hi, c->att_p_lo, c->att_w_hi[1], c->att_w_lo[1]));
      IGM_TRY(tc_plan_img(c->st, a.tc_att_t, hd, hd, a.H, a.W, Bm, a.qkv_hi + 2 * hd, a.qkv_lo + 2 * hd, c->att_w_hi[2],
                          c->att_w_lo[2], 3 * hd));
    }
    if (!a.mb_ok || !a.ln.hi) return IGM_OK;
    if (a.qkv.tc_f.valid && a.qkv.wf_hi) {
      const int hd = kHeads * kDimHead;   // packed weights are [Cout rows][Cin]: rows hd .. 3 hd - 1 are k | v
      IGM_TRY(tc_plan(c->st, a.tc_kv, a.C, 2 * hd, a.H, a.W, c->cfg.max_batch, 1, 0, a.ln.hi, a.ln.lo,
                      a.qkv.wf_hi + (int64_t)hd * a.C, a.qkv.wf_lo + (int64_t)hd * a.C));
    }
    return tc_plan_img(c->st, a.tc_mb, a.C, a.C, a.H, a.W, c->cfg.max_batch, a.ln.hi, a.ln.lo, a.mb_hi, a.mb_lo);
  };
  if (rc == IGM_OK) for (auto& s : c->downs) { rc = plan_mb(s.attn); if (rc != IGM_OK) break; }
  if (rc == IGM_OK) for (auto& s : c->ups) { rc = plan_mb(s.attn); if (rc != IGM_OK) break; }
  if (rc == IGM_OK) rc = plan_mb(c->mid_attn);
  // stride-2 resampling convs
  auto plan_rs = [&](ResampleL& rs) -> int {
    ConvL& l = rs.conv;
    if (!rs.present || !l.rs_ok || !rs.in || !rs.in->hi) return IGM_OK;
    const int Bm = c->cfg.max_batch, C = l.Cin;
    const int64_t KK = (int64_t)l.K * l.K;
    if (l.rs_kind == 1) {
      // Downsample Conv2d(C, C, 3, 2, 1): forward = strided conv; data gradient = 4 parity phases over dY
      IGM_TRY(tc_plan_strided(c->st, rs.tcf[0], C, C, rs.Hin, rs.Win, Bm, 3, 1, rs.in->hi, rs.in->lo, l.wf_hi, l.wf_lo));
      rs.n_tcf = 1;
      if (c->cfg.training) {
        IGM_TRY(tc_plan_phases4(c->st, rs.tcb[0], C, C, rs.out.H, rs.out.W, Bm, 3, 1, c->dy_hi[l.dyb], c->dy_lo[l.dyb], l.wb_hi, l.wb_lo));
        rs.n_tcb = 1;
        // S = X (fine grid, ci), P = dY (coarse grid, co); OIHW: ci stride KK, co stride C*KK
        IGM_TRY(tcw_plan_strided(c->st, rs.tcw, C, C, rs.out.H, rs.out.W, Bm, 3, 1, rs.in->hi, rs.in->lo, c->dy_hi[l.dyb], c->dy_lo[l.dyb],
                                 KK, (int64_t)C * KK));
      }
    } else {
      // Upsample ConvTranspose2d(C, C, 4, 2, 1): forward = 4 parity phases over X; data gradient = strided conv over dY
      IGM_TRY(tc_plan_phases4(c->st, rs.tcf[0], C, C, rs.Hin, rs.Win, Bm, 4, 1, rs.in->hi, rs.in->lo, l.wf_hi, l.wf_lo));
      rs.n_tcf = 1;
      if (c->cfg.training) {
        IGM_TRY(tc_plan_strided(c->st, rs.tcb[0], C, C, rs.out.H, rs.out.W, Bm, 4, 1, c->dy_hi[l.dyb], c->dy_lo[l.dyb], l.wb_hi, l.wb_lo));
        rs.n_tcb = 1;
        // S = dY (fine grid, co), P = X (coarse grid, ci); IOHW: co stride KK, ci stride C*KK
        IGM_TRY(tcw_plan_strided(c->st, rs.tcw, C, C, rs.Hin, rs.Win, Bm, 4, 1, c->dy_hi[l.dyb], c->dy_lo[l.dyb], rs.in->hi, rs.in->lo,
                                 KK, (int64_t)C * KK));
      }
    }
    l.rs_f_valid = rs.n_tcf > 0;
    l.rs_b_valid = rs.n_tcb > 0;
    n_valid += rs.n_tcf + rs.n_tcb;
    return IGM_OK;
  };
  if (rc == IGM_OK)
    for (auto& s : c->downs) { rc = plan_rs(s.rs); if (rc != IGM_OK) break; }
  if (rc == IGM_OK)
    for (auto& s : c->ups) { rc = plan_rs(s.rs); if (rc != IGM_OK) break; }
  c->tc_available = (rc == IGM_OK) && n_valid > 0;
  return rc;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int igm_version(void) { return 100; }

const char* igm_last_error(const igm_ctx* ctx) {
  const Status& st = ctx ? ctx->st : global_status();
  return st.msg.c_str();
}

static int build_plan(igm_ctx* c, float* base, int64_t* floats_out) {
  PlanBuilder pb(*c, base);
  pb.build();
  // skip-gradient buffers (training): one per down stage >= 1
  c->skip_g.assign(c->cfg.n_mults, nullptr);
  if (c->cfg.training) {
    for (int i = 1; i < c->cfg.n_mults; ++i) {
      const AttnL& a = c->downs[i].attn;
      c->skip_g[i] = pb.ar.alloc((int64_t)c->cfg.max_batch * a.H * a.W * a.C);
    }
  }
  // halo-wgrad workspaces: one contiguous block (the finalize pass walks all of them; scattered over the
  // multi-GB arena it was TLB-latency bound)
  {
    float* blk = pb.ar.alloc(pb.halo_w_elems + 64);
    int64_t off = 0;
    for_each_conv(c, [&](ConvL& l) -> int {
      if (l.tc_wh_ok) {
        l.wg_ws = blk ? blk + off : nullptr;
        off += (int64_t)l.K * l.K * l.Cin * l.Cout;
      }
      return IGM_OK;
    });
  }
  *floats_out = pb.ar.used;
  return IGM_OK;
}

int igm_unet_create(igm_ctx** out, const igm_unet_cfg* cfg, int device) {
  Status& st = global_status();
  st = Status();
  if (!out || !cfg) IGM_FAIL(st, IGM_ERR_INVALID, "null argument");
  if (cfg->dim <= 0 || cfg->dim % 32 != 0 || 256 % cfg->dim != 0)
    IGM_FAIL(st, IGM_ERR_INVALID, "dim must be 32, 64, 128 or 256");
  if (cfg->channels < 1 || cfg->channels > 4) IGM_FAIL(st, IGM_ERR_INVALID, "channels must be in [1,4]");
  if (cfg->n_mults < 1 || cfg->n_mults > IGM_MAX_MULTS) IGM_FAIL(st, IGM_ERR_INVALID, "bad n_mults");
  for (int i = 0; i < cfg->n_mults; ++i)
    if (cfg->dim_mults[i] < 1 || cfg->dim * cfg->dim_mults[i] > 1024)
      IGM_FAIL(st, IGM_ERR_INVALID, "dim * mult must be in [32, 1024]");
  if (cfg->height < 1 || cfg->width < 1 || cfg->max_batch < 1) IGM_FAIL(st, IGM_ERR_INVALID, "bad image size / batch");
  {
    // the up path doubles what the down path halved: H, W must survive n_mults-1 exact halvings
    int h = cfg->height, w = cfg->width;
    for (int i = 0; i + 1 < cfg->n_mults; ++i) {
      if (h % 2 || w % 2) IGM_FAIL(st, IGM_ERR_INVALID, "height/width must be divisible by 2^(n_mults-1)");
      h /= 2; w /= 2;
    }
  }
  if (cfg->loss_type != 1 && cfg->loss_type != 2) IGM_FAIL(st, IGM_ERR_INVALID, "loss_type must be 1 (l1) or 2 (l2)");
  IGM_CUDA(st, cudaSetDevice(device));
  igm_ctx* c = new igm_ctx();
  c->cfg = *cfg;
  c->device = device;
  int64_t floats = 0;
  build_plan(c, nullptr, &floats);
  for (auto& st2 : c->ups)
    if (!st2.r1.has_res) {
      delete c;
      IGM_FAIL(st, IGM_ERR_INVALID, "unsupported dim_mults: an up-path ResnetBlock without res_conv");
    }
  cudaError_t e = cudaMalloc(&c->arena, (size_t)floats * sizeof(float));
  if (e != cudaSuccess) {
    delete c;
    set_error(st, IGM_ERR_NOMEM, __FILE__, __LINE__, cudaGetErrorString(e));
    return st.code;
  }
  c->arena_floats = floats;
  e = cudaMemset(c->arena, 0, (size_t)floats * sizeof(float));
  if (e != cudaSuccess) {
    cudaFree(c->arena);
    delete c;
    set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(e));
    return st.code;
  }
  int64_t floats2 = 0;
  build_plan(c, c->arena, &floats2);
  wire_plan(c);
  assign_dy_slots(c);
  plan_buckets(c);
  // tensor-core engine: on by default when the shapes allow it (IGM_CONV_ENGINE=0 forces the SIMT engine)
  if (const char* ps = getenv("IGM_PREFER_SHARED")) {
    if (ps[0] == '1') cudaDeviceSetCacheConfig(cudaFuncCachePreferShared);
  }
  if (const char* mbe = getenv("IGM_ATTN_MB")) c->attn_mb = !(mbe[0] == '0');
  if (const char* kve = getenv("IGM_ATTN_KV")) c->attn_kv_only = !(kve[0] == '0');
  if (const char* ate = getenv("IGM_ATTN_TC")) c->attn_tc = !(ate[0] == '0');
  if (const char* atm = getenv("IGM_ATTN_TC_MIN")) c->attn_tc_min_pix = atoll(atm);
  const char* halo = getenv("IGM_WGRAD_HALO");
  c->halo_on = !(halo && halo[0] == '0');
  const char* eng = getenv("IGM_CONV_ENGINE");
  if (!(eng && eng[0] == '0')) {
    if (plan_tc(c) != IGM_OK) {
      st = c->st;
      cudaFree(c->arena);
      delete c;
      return st.code;
    }
    c->conv_engine = c->tc_available ? 1 : 0;
  }
  if (const char* ws = getenv("IGM_WGRAD_STREAM")) c->side_on = !(ws[0] == '0');
  if (c->cfg.training && c->side_on) {
    cudaError_t se = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_wr, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming);
    if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_ln, cudaEventDisableTiming);
    for (int k = 0; k < igm_ctx::kDyBufs && se == cudaSuccess; ++k)
      se = cudaEventCreateWithFlags(&c->ev_rd[k], cudaEventDisableTiming);

    if (se != cudaSuccess) {
      set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(se));
      igm_unet_destroy(c);
      return st.code;
    }
  }
  if (c->cfg.training) {
    cudaError_t se = cudaSuccess;
    for (int k = 0; k < c->n_buckets && se == cudaSuccess; ++k) {
      se = cudaEventCreateWithFlags(&c->ev_bk_main[k], cudaEventDisableTiming);
      if (se == cudaSuccess) se = cudaEventCreateWithFlags(&c->ev_bk_side[k], cudaEventDisableTiming);
    }
    if (se != cudaSuccess) {
      set_error(st, IGM_ERR_CUDA, __FILE__, __LINE__, cudaGetErrorString(se));
      igm_unet_destroy(c);
      return st.code;
    }
  }
  *out = c;
  return IGM_OK;
}

void igm_unet_destroy(igm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
  if (c->side) { cudaStreamSynchronize(c->side); cudaStreamDestroy(c->side); }
  if (c->ev_wr) cudaEventDestroy(c->ev_wr);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_ln) cudaEventDestroy(c->ev_ln);
  for (cudaEvent_t e : c->ev_rd) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_bk_main) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_bk_side) if (e) cudaEventDestroy(e);
  c->prof.reset();
  for (cudaEvent_t e : c->prof.pool) cudaEventDestroy(e);
  if (c->arena) cudaFree(c->arena);
  delete c;
}

int igm_unet_num_params(const igm_ctx* c) { return c ? (int)c->params.size() : IGM_ERR_INVALID; }
int64_t igm_unet_param_elems(const igm_ctx* c) { return c ? c->param_elems : IGM_ERR_INVALID; }

int igm_unet_param_info(const igm_ctx* c, int index, char* name_buf, int name_cap, int64_t* offset, int32_t* ndim,
                        int64_t shape[4]) {
  if (!c || index < 0 || index >= (int)c->params.size()) return IGM_ERR_INVALID;
  const ParamInfo& p = c->params[index];
  if (name_buf && name_cap > 0) {
    strncpy(name_buf, p.name.c_str(), name_cap - 1);
    name_buf[name_cap - 1] = 0;
  }
  if (offset) *offset = p.offset;
  if (ndim) *ndim = p.ndim;
  if (shape)
    for (int i = 0; i < 4; ++i) shape[i] = i < p.ndim ? p.shape[i] : 1;
  return IGM_OK;
}

int igm_unet_bind_params(igm_ctx* c, float* params, float* grads) {
  if (!c || !params) return IGM_ERR_INVALID;
  if (c->cfg.training && !grads) IGM_FAIL(c->st, IGM_ERR_INVALID, "training context needs a gradient arena");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  c->P = params;
  c->G = grads;
  // per-block time projections: fill the device table
  int k = 0;
  auto fill = [&](const ResnetL& r) {
    TimeProj& tp = c->proj_host[k++];
    tp.w = c->Pp(r.mlp_w); tp.b = c->Pp(r.mlp_b);
    tp.gw = grads ? c->Gp(r.mlp_w) : nullptr; tp.gb = grads ? c->Gp(r.mlp_b) : nullptr;
  };
  for (auto& s : c->downs) { fill(s.r1); fill(s.r2); }
  for (auto& s : c->ups) { fill(s.r1); fill(s.r2); }
  fill(c->mid1); fill(c->mid2);
  IGM_CUDA(c->st, cudaMemcpy(c->proj_dev, c->proj_host.data(), sizeof(TimeProj) * c->n_proj, cudaMemcpyHostToDevice));
  if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
  c->pack_engine = -1;   // job table holds raw parameter pointers: rebuild at the next pack
  // halo weight-gradient workspaces -> OIHW gradients
  c->fin_n = 0; c->fin_tiles = 0; c->fin_elems = 0;
  if (grads) {
    std::vector<HaloFinJob> jobs;
    std::vector<int> cta_job;
    for_each_conv(c, [&](ConvL& l) -> int {
      if (!l.tc_wh.valid) return IGM_OK;
      HaloFinJob j{l.wg_ws, c->Gp(l.pw), l.Cin, l.Cout, c->fin_tiles};
      const int n_ctas = (l.Cin / 32) * (l.Cout / 32);   // one CTA per 32 x 32 (ci, co) tile
      l.fin_begin = c->fin_tiles; l.fin_ctas = n_ctas;
      c->fin_tiles += n_ctas;
      c->fin_elems += 9.0 * l.Cin * l.Cout;
      cta_job.insert(cta_job.end(), n_ctas, (int)jobs.size());
      jobs.push_back(j);
      return IGM_OK;
    });
    if (jobs.size() > 256 || (int)cta_job.size() > c->fin_cta_cap) IGM_FAIL(c->st, IGM_ERR_INVALID, "too many halo-wgrad layers");
    if (!jobs.empty()) {
      IGM_CUDA(c->st, cudaMemcpy(c->fin_dev, jobs.data(), jobs.size() * sizeof(HaloFinJob), cudaMemcpyHostToDevice));
      IGM_CUDA(c->st, cudaMemcpy(c->fin_cta_dev, cta_job.data(), cta_job.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    c->fin_n = (int)jobs.size();
  }
  return IGM_OK;
}

// Job list of one full re-pack for the active engine (only the layouts that engine reads).
static int build_pack_jobs(igm_ctx* c) {
  std::vector<PackJob> jobs, tjobs;
  std::vector<int> cta_job;
  int64_t total = 0;
  c->packt_taps = 1; c->packt_elems = 0;
  auto add = [&](const float* src, float* dst_f, void* hi, void* lo, int taps, int K, int N, int64_t sk, int64_t sn,
                 int flip) {
    PackJob j{};
    j.src = src; j.dst_f = dst_f; j.dst_hi = hi; j.dst_lo = lo;
    j.taps = taps; j.K = K; j.N = N; j.flip = flip; j.sk = sk; j.sn = sn;
    const int n_ctas = (N / 32) * (K / 32);
    if (hi && N % 32 == 0 && K % 32 == 0 && taps <= 16 && (sk == taps || sn == taps) &&
        (int)cta_job.size() + n_ctas <= c->packt_cap && tjobs.size() < 1024) {
      // tiled path: whole-line accesses through shared memory
      j.begin = (int64_t)cta_job.size();
      cta_job.insert(cta_job.end(), n_ctas, (int)tjobs.size());
      tjobs.push_back(j);
      c->packt_taps = std::max(c->packt_taps, taps);
      c->packt_elems += (double)taps * K * N;
      return;
    }
    j.begin = total;
    total += (int64_t)taps * K * N;
    jobs.push_back(j);
  };
  const bool tc = c->conv_engine == 1;
  for_each_conv(c, [&](ConvL& l) -> int {
    const int KK = l.K * l.K;
    const float* w = c->Pp(l.pw);
    if (!l.convT) {
      // Conv2d OIHW: fprop contracts ci (sk = KK, sn = Cin*KK); dgrad contracts co
      // (the stride-2 Downsample keeps its taps un-flipped: its phase plans carry explicit offsets)
      if (tc && (l.tc_f.valid || l.rs_f_valid)) add(w, nullptr, l.wf_hi, l.wf_lo, KK, l.Cin, l.Cout, KK, (int64_t)l.Cin * KK, 0);
      else add(w, l.w_fwd, nullptr, nullptr, KK, l.Cin, l.Cout, KK, (int64_t)l.Cin * KK, 0);
      if (tc && l.tc_b.valid) add(w, nullptr, l.wb_hi, l.wb_lo, KK, l.Cout, l.Cin, (int64_t)l.Cin * KK, KK, 1);
      else if (tc && l.rs_b_valid) add(w, nullptr, l.wb_hi, l.wb_lo, KK, l.Cout, l.Cin, (int64_t)l.Cin * KK, KK, 0);
      else if (l.w_bwd) add(w, l.w_bwd, nullptr, nullptr, KK, l.Cout, l.Cin, (int64_t)l.Cin * KK, KK, 0);
    } else {
      // ConvTranspose2d IOHW
      if (tc && l.rs_f_valid) add(w, nullptr, l.wf_hi, l.wf_lo, KK, l.Cin, l.Cout, (int64_t)l.Cout * KK, KK, 0);
      else add(w, l.w_fwd, nullptr, nullptr, KK, l.Cin, l.Cout, (int64_t)l.Cout * KK, KK, 0);
      if (tc && l.rs_b_valid) add(w, nullptr, l.wb_hi, l.wb_lo, KK, l.Cout, l.Cin, KK, (int64_t)l.Cout * KK, 0);
      else if (l.w_bwd) add(w, l.w_bwd, nullptr, nullptr, KK, l.Cout, l.Cin, KK, (int64_t)l.Cout * KK, 0);
    }
    return IGM_OK;
  });
  if (c->final_wbwd)   // dgrad of the final 1x1: [1][k = c][n = K] = W[c][K]
    add(c->Pp(c->final_conv.pw), c->final_wbwd, nullptr, nullptr, 1, c->cfg.channels, c->final_conv.Cin,
        c->final_conv.Cin, 1, 0);
  if (jobs.size() > 1024) IGM_FAIL(c->st, IGM_ERR_INVALID, "too many pack jobs");
  if (!jobs.empty())
    IGM_CUDA(c->st, cudaMemcpy(c->pack_dev, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
  c->pack_n = (int)jobs.size();
  c->pack_total = total;
  if (!tjobs.empty()) {
    IGM_CUDA(c->st, cudaMemcpy(c->packt_dev, tjobs.data(), tjobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice));
    IGM_CUDA(c->st, cudaMemcpy(c->packt_cta_dev, cta_job.data(), cta_job.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  c->packt_ctas = (int)cta_job.size();
  c->pack_engine = c->conv_engine;
  return IGM_OK;
}

int igm_unet_pack_weights(igm_ctx* c, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!c->P) IGM_FAIL(c->st, IGM_ERR_STATE, "bind parameters first");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  if (c->pack_engine != c->conv_engine) IGM_TRY(build_pack_jobs(c));
  LaunchCtx lc = c->lc(stream);
  IGM_TRY(launch_pack_tiles(lc, c->packt_dev, c->packt_cta_dev, c->packt_ctas, c->packt_taps, c->packt_elems));
  return launch_pack_jobs(lc, c->pack_dev, c->pack_n, c->pack_total);
}

static int check_ready(igm_ctx* c, int B) {
  if (!c) return IGM_ERR_INVALID;
  if (!c->P) IGM_FAIL(c->st, IGM_ERR_STATE, "parameters not bound");
  if (B < 1 || B > c->cfg.max_batch) IGM_FAIL(c->st, IGM_ERR_INVALID, "batch exceeds max_batch");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  return IGM_OK;
}

int igm_unet_forward(igm_ctx* c, const float* x, const int64_t* t, float* out, int B, void* stream) {
  IGM_TRY(check_ready(c, B));
  if (!x || !t || !out) IGM_FAIL(c->st, IGM_ERR_INVALID, "null tensor");
  Runner r{*c, c->lc(stream), B};
  const int HW = c->cfg.height * c->cfg.width;
  IGM_TRY(launch_input_prep(r.lc, x, nullptr, nullptr, nullptr, nullptr, c->x_in.v, nullptr, B, c->cfg.channels, HW));
  IGM_TRY(r.forward(t, out));
  c->last_B = B;
  c->fwd_valid = true;
  c->loss_valid = false;
  return IGM_OK;
}

int igm_unet_backward(igm_ctx* c, const float* d_out, float* d_x, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!c->cfg.training) IGM_FAIL(c->st, IGM_ERR_STATE, "context was created with training = 0");
  if (!c->fwd_valid) IGM_FAIL(c->st, IGM_ERR_STATE, "backward without a forward");
  if (!d_out) IGM_FAIL(c->st, IGM_ERR_INVALID, "null tensor");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  const int B = c->last_B;
  Runner r{*c, c->lc(stream), B};
  const int HW = c->cfg.height * c->cfg.width;
  IGM_TRY(launch_nchw_to_nhwc(r.lc, d_out, c->d_pred, B, HW, c->cfg.channels));
  // d_x (NHWC) is staged in noise_copy, then transposed out
  float* dx_nhwc = d_x ? c->noise_copy : nullptr;
  IGM_TRY(r.backward(c->d_pred, dx_nhwc));
  if (d_x) IGM_TRY(launch_nhwc_to_nchw(r.lc, dx_nhwc, d_x, B, HW, c->cfg.channels));
  return IGM_OK;
}

int igm_unet_grad_buckets(const igm_ctx* c, int64_t* lo, int64_t* hi, int cap) {
  if (!c) return IGM_ERR_INVALID;
  for (int k = 0; k < c->n_buckets && k < cap; ++k) {
    if (lo) lo[k] = c->bucket_lo[k];
    if (hi) hi[k] = c->bucket_hi[k];
  }
  return c->n_buckets;
}

int igm_unet_bucket_wait(igm_ctx* c, int k, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (k < 0 || k >= c->n_buckets) IGM_FAIL(c->st, IGM_ERR_INVALID, "bad bucket index");
  if (!c->bk_valid || !c->ev_bk_main[k]) IGM_FAIL(c->st, IGM_ERR_STATE, "bucket_wait without a backward pass of a training context");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  IGM_CUDA(c->st, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_bk_main[k], 0));
  if (c->bk_side_rec[k]) IGM_CUDA(c->st, cudaStreamWaitEvent((cudaStream_t)stream, c->ev_bk_side[k], 0));
  return IGM_OK;
}

int igm_ddpm_set_schedule(igm_ctx* c, const igm_schedule* s) {
  if (!c || !s) return IGM_ERR_INVALID;
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  c->sched = *s;
  IGM_CUDA(c->st, cudaMemcpy(c->sched_dev, s, sizeof(igm_schedule), cudaMemcpyHostToDevice));
  c->have_sched = true;
  return IGM_OK;
}

int igm_ddpm_q_sample(igm_ctx* c, const float* x_start, const int64_t* t, const float* noise, float* out, int B,
                      void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!c->have_sched) IGM_FAIL(c->st, IGM_ERR_STATE, "schedule not set");
  if (!x_start || !t || !noise || !out) IGM_FAIL(c->st, IGM_ERR_INVALID, "null tensor");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  LaunchCtx lc = c->lc(stream);
  return launch_input_prep(lc, x_start, noise, t, c->sched.sqrt_alphas_cumprod, c->sched.sqrt_one_minus_alphas_cumprod,
                           nullptr, out, B, c->cfg.channels, c->cfg.height * c->cfg.width);
}

int igm_ddpm_p_losses(igm_ctx* c, const float* x_start, const int64_t* t, const float* noise, float* loss_out, int B,
                      void* stream) {
  IGM_TRY(check_ready(c, B));
  if (!c->have_sched) IGM_FAIL(c->st, IGM_ERR_STATE, "schedule not set");
  if (!x_start || !t || !noise || !loss_out) IGM_FAIL(c->st, IGM_ERR_INVALID, "null tensor");
  Runner r{*c, c->lc(stream), B};
  const int HW = c->cfg.height * c->cfg.width;
  const int64_t n = (int64_t)B * c->cfg.channels * HW;
  IGM_TRY(launch_input_prep(r.lc, x_start, noise, t, c->sched.sqrt_alphas_cumprod,
                            c->sched.sqrt_one_minus_alphas_cumprod, c->x_in.v, nullptr, B, c->cfg.channels, HW));
  IGM_TRY(r.forward(t, c->pred));
  IGM_TRY(launch_loss(r.lc, c->pred, noise, n, c->cfg.loss_type, c->loss_ws, loss_out));
  if (c->cfg.training) {
    IGM_CUDA(c->st, cudaMemcpyAsync(c->noise_copy, noise, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice,
                                    (cudaStream_t)stream));
  }
  c->last_B = B;
  c->fwd_valid = true;
  c->loss_valid = c->cfg.training != 0;
  return IGM_OK;
}

int igm_ddpm_p_losses_backward(igm_ctx* c, const float* d_loss, float scale, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!c->cfg.training) IGM_FAIL(c->st, IGM_ERR_STATE, "context was created with training = 0");
  if (!c->loss_valid) IGM_FAIL(c->st, IGM_ERR_STATE, "p_losses_backward without p_losses");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  const int B = c->last_B;
  Runner r{*c, c->lc(stream), B};
  const int HW = c->cfg.height * c->cfg.width;
  const int64_t n = (int64_t)B * c->cfg.channels * HW;
  IGM_TRY(launch_loss_backward_nhwc(r.lc, c->pred, c->noise_copy, n, c->cfg.channels, HW, c->cfg.loss_type,
                                    d_loss, scale, c->d_pred));
  IGM_TRY(r.backward(c->d_pred, nullptr));
  return IGM_OK;
}

static int sampler_step(igm_ctx* c, Runner& r, float* img, const float* noise, uint64_t seed, int clip) {
  const int HW = c->cfg.height * c->cfg.width;
  const int64_t n = (int64_t)r.B * c->cfg.channels * HW;
  IGM_TRY(launch_sampler_tick(r.lc, c->t_vec, r.B, c->loop_state));
  IGM_TRY(launch_input_prep(r.lc, img, nullptr, nullptr, nullptr, nullptr, c->x_in.v, nullptr, r.B, c->cfg.channels, HW));
  IGM_TRY(r.forward(c->t_vec, c->pred));
  SamplerStepArgs a;
  a.img = img; a.eps = c->pred; a.noise = noise; a.t_dev = c->t_vec; a.sched_dev = c->sched_dev;
  a.seed = seed; a.step_dev = c->loop_state + 1; a.n = n; a.per_sample = c->cfg.channels * HW; a.clip = clip;
  IGM_TRY(launch_sampler_update(r.lc, a));
  return IGM_OK;
}

int igm_ddpm_sample_loop(igm_ctx* c, float* img, const float* noise, uint64_t seed, int B, int t_start, int n_steps,
                         int clip_denoised, void* stream) {
  IGM_TRY(check_ready(c, B));
  if (!c->have_sched) IGM_FAIL(c->st, IGM_ERR_STATE, "schedule not set");
  if (!img) IGM_FAIL(c->st, IGM_ERR_INVALID, "null tensor");
  if (t_start < 0 || t_start >= c->cfg.timesteps || n_steps < 0 || n_steps > t_start + 1)
    IGM_FAIL(c->st, IGM_ERR_INVALID, "bad t_start / n_steps");
  if (n_steps == 0) return IGM_OK;
  cudaStream_t s = (cudaStream_t)stream;
  Runner r{*c, c->lc(stream), B};
  r.infer = true;
  const int init[2] = {t_start, 0};
  IGM_CUDA(c->st, cudaMemcpyAsync(c->loop_state, init, sizeof(init), cudaMemcpyHostToDevice, s));
  // first step eagerly (also resolves lazy per-kernel attributes outside of stream capture)
  const int64_t before = c->launches;
  IGM_TRY(sampler_step(c, r, img, noise, seed, clip_denoised));
  const int64_t per_step = c->launches - before;
  if (n_steps == 1) { c->fwd_valid = false; c->loss_valid = false; return IGM_OK; }
  igm_ctx::GraphKey key;
  key.img = img; key.noise = noise; key.seed = seed; key.B = B; key.clip = clip_denoised;
  if (!c->graph_exec || !(c->graph_key == key)) {
    if (c->graph_exec) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
    cudaStream_t cs = s;
    bool own = false;
    if (s == 0 || s == cudaStreamLegacy) {   // the legacy stream cannot be captured
      IGM_CUDA(c->st, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      own = true;
    }
    const int64_t saved = c->launches;
    IGM_CUDA(c->st, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    Runner rc{*c, c->lc((void*)cs), B};
    rc.infer = true;
    int rcode = sampler_step(c, rc, img, noise, seed, clip_denoised);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &graph);
    c->launches = saved;   // capture enqueues nothing
    if (own) cudaStreamDestroy(cs);
    if (rcode != IGM_OK) { if (graph) cudaGraphDestroy(graph); return rcode; }
    if (e != cudaSuccess) IGM_FAIL(c->st, IGM_ERR_CUDA, cudaGetErrorString(e));
    e = cudaGraphInstantiate(&c->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) { c->graph_exec = nullptr; IGM_FAIL(c->st, IGM_ERR_CUDA, cudaGetErrorString(e)); }
    c->graph_key = key;
  }
  for (int k = 1; k < n_steps; ++k) {
    IGM_CUDA(c->st, cudaGraphLaunch(c->graph_exec, s));
  }
  c->launches += (int64_t)(n_steps - 1) * per_step;   // each replay runs the same kernels as the eager step
  c->fwd_valid = false;
  c->loss_valid = false;
  return IGM_OK;
}

int igm_adam_step(igm_ctx* c, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) IGM_FAIL(c->st, IGM_ERR_INVALID, "bad adam args");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  LaunchCtx lc = c->lc(stream);
  return launch_adam(lc, params, grads, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step, grad_scale);
}

int igm_grad_axpy(igm_ctx* c, float* dst, const float* src, const float* alpha, float scale, int64_t n, void* stream) {
  if (!c) return IGM_ERR_INVALID;
  if (!dst || !src || n < 0) IGM_FAIL(c->st, IGM_ERR_INVALID, "bad axpy args");
  if ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) IGM_FAIL(c->st, IGM_ERR_INVALID, "axpy arenas must be 16-byte aligned");
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  LaunchCtx lc = c->lc(stream);
  return launch_axpy(lc, dst, src, alpha, scale, n);
}

int64_t igm_debug_read_tap(igm_ctx* c, const char* name, float* dst, int64_t cap, void* stream) {
  if (!c || !name) return IGM_ERR_INVALID;
  auto it = c->taps.find(name);
  if (it == c->taps.end()) { set_error(c->st, IGM_ERR_INVALID, __FILE__, __LINE__, "unknown tap"); return IGM_ERR_INVALID; }
  const Act& a = it->second;
  const int B = c->last_B > 0 ? c->last_B : c->cfg.max_batch;
  const int64_t n = (int64_t)B * a.C * a.H * a.W;
  if (!dst) return n;
  if (cap < n) { set_error(c->st, IGM_ERR_INVALID, __FILE__, __LINE__, "tap buffer too small"); return IGM_ERR_INVALID; }
  if (cudaSetDevice(c->device) != cudaSuccess) return IGM_ERR_CUDA;
  LaunchCtx lc = c->lc(stream);
  if (c->conv_engine == 1 && a.hi) {   // the fp32 copy may have been skipped: rebuild it from the bf16 hi + lo staging
    int rr = launch_merge_bf16(lc, a.hi, a.lo, a.v, n);
    if (rr != IGM_OK) return rr;
  }
  int r = launch_nhwc_to_nchw(lc, a.v, dst, B, a.H * a.W, a.C);
  return r == IGM_OK ? n : r;
}

int igm_profile_start(igm_ctx* c) {
  if (!c) return IGM_ERR_INVALID;
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  c->prof.reset();
  c->prof.on = true;
  return IGM_OK;
}

int igm_profile_stop(igm_ctx* c, igm_profile_entry* out, int cap) {
  if (!c) return IGM_ERR_INVALID;
  IGM_CUDA(c->st, cudaSetDevice(c->device));
  c->prof.on = false;
  IGM_CUDA(c->st, cudaDeviceSynchronize());
  igm_profile_entry acc[K_NCLASS];
  for (int k = 0; k < K_NCLASS; ++k) {
    memset(&acc[k], 0, sizeof(acc[k]));
    strncpy(acc[k].name, kclass_name(k), sizeof(acc[k].name) - 1);
  }
  // IGM_PROFILE_DUMP=<file>: also append one line per launch scope (index, class, ms, flops, bytes) for tuning
  FILE* dump = nullptr;
  if (const char* path = getenv("IGM_PROFILE_DUMP")) dump = fopen(path, "a");
  int ridx = 0;
  for (auto& r : c->prof.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
    if (dump) fprintf(dump, "%d %s %.4f %.4g %.4g\n", ridx++, kclass_name(r.cls), ms, r.flops, r.bytes);
    acc[r.cls].launches += 1;
    acc[r.cls].ms += ms;
    acc[r.cls].flops += r.flops;
    acc[r.cls].bytes += r.bytes;
  }
  if (dump) { fprintf(dump, "#\n"); fclose(dump); }
  c->prof.reset();
  int n = 0;
  for (int k = 0; k < K_NCLASS && n < cap; ++k)
    if (out) out[n++] = acc[k];
  return n;
}

int64_t igm_launch_count(const igm_ctx* c) { return c ? c->launches : IGM_ERR_INVALID; }

int igm_set_conv_engine(igm_ctx* c, int engine) {
  if (!c) return IGM_ERR_INVALID;
  if (engine != 0 && engine != 1) IGM_FAIL(c->st, IGM_ERR_INVALID, "conv engine must be 0 (SIMT fp32) or 1 (tcgen05 bf16x3)");
  if (engine == 1 && !c->tc_available) {
    IGM_TRY(plan_tc(c));
    if (!c->tc_available) IGM_FAIL(c->st, IGM_ERR_INVALID, "no layer of this network is eligible for the tcgen05 engine");
  }
  if (c->graph_exec && engine != c->conv_engine) { cudaGraphExecDestroy(c->graph_exec); c->graph_exec = nullptr; }
  const bool changed = engine != c->conv_engine;
  c->conv_engine = engine;
  if (changed) c->fwd_valid = c->loss_valid = false;   // staged operands of the last forward belong to the old engine
  if (changed && c->P) IGM_TRY(igm_unet_pack_weights(c, nullptr));   // the other engine reads other layouts
  return IGM_OK;
}
int igm_get_conv_engine(const igm_ctx* c) { return c ? c->conv_engine : IGM_ERR_INVALID; }

}  // extern "C"
