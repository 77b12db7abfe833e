// tcgen05 weight-gradient engine for the stride-1 3x3 / 1x1 convolutions (sm_100a).
//
//   dW[co][ci][ky][kx] = sum_{b,iy,ix} dY[b, iy+pad-ky, ix+pad-kx, co] * X[b, iy, ix, ci]
//
// Lowered to GEMMs whose reduction dimension is the PIXEL axis: both operands are NHWC, so the
// 64-channel runs are contiguous along M / N and pixels are the strided K axis -> "MN-major" UMMA
// operands.  The very same TMA boxes as the forward engine are used (64 channels x 64 pixels, 128B
// swizzle); the (ky,kx) shift is applied to the dY box coordinates and the TMA out-of-bounds zero
// fill supplies the implicit zero padding.
//
// One CTA = one accumulator: M = 128 = two (tap, 64-channel co block) pairs, N = 64/128 input
// channels, reduced over its split of the pixel tiles; bf16x3 (hi*hi + hi*lo + lo*hi) with fp32
// accumulation in TMEM; the epilogue adds the partial result into PyTorch's OIHW gradient with
// fp32 atomics (split-K across CTAs), which also gives torch's accumulate-into-.grad semantics.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>

namespace igm {
namespace {

using namespace tc;

constexpr int PIX = 64;                 // pixels (K rows) per pipeline stage
constexpr int BOX_BYTES = PIX * 128;    // one 64-channel x 64-pixel bf16 box
constexpr int UMMA_K = 16;

struct WArgs {
  int B, H, W, CS, CP, P0;
  int ntaps;
  TcTap taps[kTcMaxTaps];
  int64_t s_shift, s_plain;
  int BW, BH, BB, rows;
  int tiles_per_img, n_ptiles;
  int n_pairs, n_ci_tiles, splits, tiles_per_split;
  int cob;            // CS / 64
  int zero_smem;      // rows < 64: stale smem rows must read as zero (they are reduced over)
  int variant;        // descriptor-convention switch for bring-up tests (0 = canonical)
  float* grad;
};

template <int NB>   // NB = N / 64
struct WCfg {
  static constexpr int BN = NB * 64;
  static constexpr int STAGE_BYTES = (4 + 2 * NB) * BOX_BYTES;
  static constexpr int STAGES = (NB == 1) ? 4 : 3;
  // NB = 1: the hi and lo boxes of the plain operand (adjacent in smem) form ONE N = 128 MN-major operand, because
  // tcgen05.mma at N = 64 issues at the N = 128 rate (tools/mma_rate_probe.cu): two instructions instead of three
  static constexpr bool STACKED = (NB == 1);
  static constexpr int ACC_COLS = STACKED ? 2 * BN : BN;
  static constexpr int TMEM_COLS = ACC_COLS;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <int NB>
__global__ void __launch_bounds__(192, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tdy_hi, const __grid_constant__ CUtensorMap tdy_lo,
                const __grid_constant__ CUtensorMap tx_hi, const __grid_constant__ CUtensorMap tx_lo,
                const __grid_constant__ CUtensorMap tx1_hi, const __grid_constant__ CUtensorMap tx1_lo, const WArgs p) {
  // tdy_* = the SHIFTED operand S (A blocks), tx_* / tx1_* = the PLAIN operand P (B blocks)
  using C = WCfg<NB>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the shared array (not through an integer): the pointer keeps its address space, so
  // the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::STAGES;
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode ----
  int bid = blockIdx.x;
  const int split = bid % p.splits;
  bid /= p.splits;
  const int ci_tile = bid % p.n_ci_tiles;
  const int pair = bid / p.n_ci_tiles;
  const int ntaps = p.ntaps;
  int blk_tap[2], blk_co0[2];
  bool blk_ok[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int ab = 2 * pair + j;
    blk_ok[j] = ab < ntaps * p.cob;
    blk_tap[j] = blk_ok[j] ? ab / p.cob : 0;
    blk_co0[j] = blk_ok[j] ? (ab % p.cob) * 64 : 0;
  }
  const int pt0 = split * p.tiles_per_split;
  const int pt1 = min(pt0 + p.tiles_per_split, p.n_ptiles);
  const int n_stages_total = pt1 - pt0;

  if (p.zero_smem || !blk_ok[1]) {
    // rows the TMA never writes (ragged boxes, missing second block) are part of the reduction
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < C::STAGES * C::STAGE_BYTES / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (n_stages_total > 0) {
    if (warp == 0) {
      if (lane == 0) {
        prefetch_tmap(&tdy_hi); prefetch_tmap(&tdy_lo); prefetch_tmap(&tx_hi); prefetch_tmap(&tx_lo);
        const uint32_t tx_bytes = (uint32_t)(((blk_ok[0] ? 2 : 0) + (blk_ok[1] ? 2 : 0) + 2 * NB) * p.rows * 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int pt = pt0; pt < pt1; ++pt) {
          int b0, y0;
          if (p.BB > 1) { b0 = pt * p.BB; y0 = 0; }
          else { b0 = pt / p.tiles_per_img; y0 = (pt - b0 * p.tiles_per_img) * p.BH; }
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * C::STAGE_BYTES;
          mbar_expect_tx(&full[stage], tx_bytes);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            if (!blk_ok[j]) continue;
            const TcTap tp = p.taps[blk_tap[j]];
            // S shifted by the tap (and taken from sub-lattice (px, py) for stride-2 gathers);
            // out-of-image pixels are zero-filled by the TMA unit
            const int cc = blk_co0[j] + tp.px * p.CS;
            tma_load_5d(st + j * BOX_BYTES, &tdy_hi, &full[stage], cc, tp.dx, tp.py, y0 + tp.dy, b0);
            tma_load_5d(st + (2 + j) * BOX_BYTES, &tdy_lo, &full[stage], cc, tp.dx, tp.py, y0 + tp.dy, b0);
          }
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int c0 = ci_tile * C::BN + nb * 64;
            if (c0 < p.P0) {
              tma_load_5d(st + (4 + nb) * BOX_BYTES, &tx_hi, &full[stage], c0, 0, 0, y0, b0);
              tma_load_5d(st + (4 + NB + nb) * BOX_BYTES, &tx_lo, &full[stage], c0, 0, 0, y0, b0);
            } else {   // second tensor of a channel concat
              tma_load_5d(st + (4 + nb) * BOX_BYTES, &tx1_hi, &full[stage], c0 - p.P0, 0, 0, y0, b0);
              tma_load_5d(st + (4 + NB + nb) * BOX_BYTES, &tx1_lo, &full[stage], c0 - p.P0, 0, 0, y0, b0);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        // D = f32, A = B = bf16, both operands MN-major (bits 15, 16), N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(C::ACC_COLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t lbo = p.variant == 1 ? 1024u : (uint32_t)BOX_BYTES;
        const uint32_t sbo = p.variant == 1 ? (uint32_t)BOX_BYTES : 1024u;
        // descriptors by offset arithmetic on four descriptors built once (see conv_tc.cu: the issue loop is one thread)
        const uint32_t s0 = smem_u32(smem);
        const uint64_t dA_hi0 = make_sw128_mn_desc(s0, lbo, sbo), dA_lo0 = make_sw128_mn_desc(s0 + 2 * BOX_BYTES, lbo, sbo);
        const uint64_t dB_hi0 = make_sw128_mn_desc(s0 + 4 * BOX_BYTES, lbo, sbo);
        const uint64_t dB_lo0 = make_sw128_mn_desc(s0 + (4 + NB) * BOX_BYTES, lbo, sbo);
        constexpr uint32_t STAGE16 = (uint32_t)C::STAGE_BYTES >> 4;
        int stage = 0;
        uint32_t phase = 0;
        uint32_t accum = 0;
        for (int it = 0; it < n_stages_total; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t soff = (uint32_t)stage * STAGE16;
#pragma unroll
          for (int ks = 0; ks < PIX / UMMA_K; ++ks) {
            const uint32_t off = soff + (uint32_t)(ks * UMMA_K * 128 >> 4);   // 16 pixel rows of 128 B
            const uint64_t dah = dA_hi0 + off, dal = dA_lo0 + off;
            const uint64_t dbh = dB_hi0 + off, dbl = dB_lo0 + off;
            if (C::STACKED) {
              umma_bf16(tmem_base, dal, dbh, idesc, accum);   // [s_lo*p_hi | s_lo*p_lo]
              umma_bf16(tmem_base, dah, dbh, idesc, 1u);      // [s_hi*p_hi | s_hi*p_lo]
            } else {
              umma_bf16(tmem_base, dal, dbh, idesc, accum);
              umma_bf16(tmem_base, dah, dbl, idesc, 1u);
              umma_bf16(tmem_base, dah, dbh, idesc, 1u);
            }
            accum = 1u;
          }
          umma_commit(&empty[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
      }
    } else {
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int j = row >> 6;
      const int cs = blk_co0[j] + (row & 63);
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
      float* g = p.grad + (int64_t)cs * p.s_shift + (int64_t)(ci_tile * C::BN) * p.s_plain + p.taps[blk_tap[j]].wtap;
#pragma unroll 1
      for (int c0 = 0; c0 < C::BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_base + (uint32_t)c0, v);
        if (C::STACKED) {
          float w[32];
          tmem_ld_32x32(t_base + (uint32_t)(C::BN + c0), w);
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] += w[jj];
        }
        if (blk_ok[j]) {
          if (p.s_plain == 1) {
            // 1x1 convolutions: the plain-operand channels are contiguous in the gradient -> 128-bit reductions, and the
            // 32 x 32 chunk is transposed through (idle) pipeline memory so that eight lanes cover one row's 128 bytes
            // (4 L2 transactions per warp instruction instead of 32)
            float* tb = reinterpret_cast<float*>(smem) + q * (32 * 33);
            __syncwarp();
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) tb[lane * 33 + jj] = v[jj];
            __syncwarp();
            float* g0 = g - (int64_t)lane * p.s_shift;   // row of lane 0 (the warp's 32 rows share block j)
            const int rr = lane >> 3, cc = (lane & 7) * 4;
#pragma unroll
            for (int r0 = 0; r0 < 32; r0 += 4) {
              const float* src = tb + (r0 + rr) * 33 + cc;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(g0 + (int64_t)(r0 + rr) * p.s_shift + c0 + cc),
                           "f"(src[0]), "f"(src[1]), "f"(src[2]), "f"(src[3]) : "memory");
            }
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) atomicAdd(g + (int64_t)(c0 + jj) * p.s_plain, v[jj]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

template <int NB>
int launch_impl(const LaunchCtx& lc, const TcWgrad& t, const WArgs& a, int grid) {
  using C = WCfg<NB>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  cudaError_t le = launch_pdl(wgrad_tc_kernel<NB>, dim3(grid), dim3(192), (size_t)C::SMEM_BYTES, lc.stream, t.s_hi, t.s_lo, t.p_hi,
                              t.p_lo, t.p1_hi, t.p1_lo, a);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace

bool tcw_eligible(int Cin, int Cout, int H, int W, int KH) {
  if (Cin < 64 || Cin % 64 != 0 || Cout < 64 || Cout % 64 != 0) return false;
  if (KH != 1 && KH != 3) return false;
  if (W < 1 || W > PIX || H < 1) return false;
  return true;
}

bool tcw_strided_eligible(int CS, int CP, int GH, int GW, int KH) {
  if (CS < 64 || CS % 64 != 0 || CP < 64 || CP % 64 != 0) return false;
  if (KH != 3 && KH != 4) return false;
  return GW >= 1 && GW <= PIX && GH >= 1;
}

namespace {

void tile_shape(TcWgrad& t, int GH, int GW, int Bmax) {
  t.GH = GH; t.GW = GW; t.Bmax = Bmax;
  t.BW = GW;
  if (GH * GW <= PIX) { t.BH = GH; t.BB = PIX / (GH * GW); }
  else { t.BH = PIX / GW; t.BB = 1; }
  if (t.BB > Bmax) t.BB = Bmax;
  t.rows = t.BB * t.BH * t.BW;
}

// rank-5 descriptor (channel, x, sub-lattice row, y, image) of a [Bmax, SH, SW, C] bf16 tensor; see conv_tc.cu
int encode_act5(Status& st, CUtensorMap* m, void* ptr, int C, int SH, int SW, int Bmax, bool s2d, const TcWgrad& t) {
  auto enc = encode_fn();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t rowB = (cuuint64_t)SW * C * 2;
  cuuint64_t dims[5], strides[4];
  if (!s2d) {
    dims[0] = C; dims[1] = SW; dims[2] = 1; dims[3] = SH; dims[4] = Bmax;
    strides[0] = (cuuint64_t)C * 2; strides[1] = rowB; strides[2] = rowB; strides[3] = rowB * SH;
  } else {
    dims[0] = 2 * (cuuint64_t)C; dims[1] = SW / 2; dims[2] = 2; dims[3] = SH / 2; dims[4] = Bmax;
    strides[0] = (cuuint64_t)C * 4; strides[1] = rowB; strides[2] = 2 * rowB; strides[3] = rowB * SH;
  }
  cuuint32_t box[5] = {64u, (cuuint32_t)t.BW, 1u, (cuuint32_t)t.BH, (cuuint32_t)t.BB};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (wgrad) failed");
  return IGM_OK;
}

}  // namespace

int tcw_plan(Status& st, TcWgrad& t, int Cin, int Cout, int H, int W, int Bmax, int KH, int pad, __nv_bfloat16* dy_hi,
             __nv_bfloat16* dy_lo, __nv_bfloat16* x_hi, __nv_bfloat16* x_lo, int C0, __nv_bfloat16* x1_hi,
             __nv_bfloat16* x1_lo) {
  t.valid = false;
  if (!tcw_eligible(Cin, Cout, H, W, KH)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the tcgen05 wgrad engine");
  if (C0 <= 0 || !x1_hi) C0 = Cin;
  if (C0 % 64 != 0 || (Cin - C0) % 64 != 0) IGM_FAIL(st, IGM_ERR_INVALID, "concat split must be a multiple of 64 channels");
  tile_shape(t, H, W, Bmax);
  t.CS = Cout; t.CP = Cin; t.P0 = C0;
  t.BN = (Cin % 128 == 0) ? 128 : 64;
  const int KK = KH * KH;
  // dW[co][ci][ky][kx] = sum dY[iy + pad - ky, ix + pad - kx, co] * X[iy, ix, ci]
  t.ntaps = KK;
  for (int ky = 0; ky < KH; ++ky)
    for (int kx = 0; kx < KH; ++kx) t.taps[ky * KH + kx] = TcTap{pad - kx, pad - ky, 0, 0, ky * KH + kx};
  t.s_shift = (int64_t)Cin * KK;   // co stride in OIHW
  t.s_plain = KK;                  // ci stride
  t.flops_per_image = 2.0 * H * W * (double)Cin * Cout * KK;
  const bool two = C0 < Cin;
  IGM_TRY(encode_act5(st, &t.s_hi, dy_hi, Cout, H, W, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.s_lo, dy_lo, Cout, H, W, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.p_hi, x_hi, C0, H, W, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.p_lo, x_lo, C0, H, W, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.p1_hi, two ? (void*)x1_hi : (void*)x_hi, two ? Cin - C0 : C0, H, W, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.p1_lo, two ? (void*)x1_lo : (void*)x_lo, two ? Cin - C0 : C0, H, W, Bmax, false, t));
  t.valid = true;
  return IGM_OK;
}

int tcw_plan_strided(Status& st, TcWgrad& t, int CS, int CP, int GH, int GW, int Bmax, int KH, int pad,
                     __nv_bfloat16* s_hi, __nv_bfloat16* s_lo, __nv_bfloat16* p_hi, __nv_bfloat16* p_lo,
                     int64_t s_shift, int64_t s_plain) {
  t.valid = false;
  if (!tcw_strided_eligible(CS, CP, GH, GW, KH)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the strided tcgen05 wgrad");
  tile_shape(t, GH, GW, Bmax);
  t.CS = CS; t.CP = CP; t.P0 = CP;
  t.BN = (CP % 128 == 0) ? 128 : 64;
  // G[ky][kx][cs][cp] = sum_{a,b} S[2a - pad + ky, 2b - pad + kx, cs] * P[a, b, cp]
  t.ntaps = KH * KH;
  auto split2 = [](int v, int& d, int& ph) { ph = ((v % 2) + 2) % 2; d = (v - ph) / 2; };
  for (int ky = 0; ky < KH; ++ky)
    for (int kx = 0; kx < KH; ++kx) {
      TcTap tp;
      split2(ky - pad, tp.dy, tp.py);
      split2(kx - pad, tp.dx, tp.px);
      tp.wtap = ky * KH + kx;
      t.taps[ky * KH + kx] = tp;
    }
  t.s_shift = s_shift; t.s_plain = s_plain;
  t.flops_per_image = 2.0 * GH * GW * (double)CS * CP * KH * KH;
  IGM_TRY(encode_act5(st, &t.s_hi, s_hi, CS, 2 * GH, 2 * GW, Bmax, true, t));
  IGM_TRY(encode_act5(st, &t.s_lo, s_lo, CS, 2 * GH, 2 * GW, Bmax, true, t));
  IGM_TRY(encode_act5(st, &t.p_hi, p_hi, CP, GH, GW, Bmax, false, t));
  IGM_TRY(encode_act5(st, &t.p_lo, p_lo, CP, GH, GW, Bmax, false, t));
  t.p1_hi = t.p_hi; t.p1_lo = t.p_lo;
  t.valid = true;
  return IGM_OK;
}

bool tcw_batch_ok(const TcWgrad& t, int B) { return t.valid && B >= 1 && B <= t.Bmax && (B % t.BB) == 0; }

int launch_wgrad_tc(const LaunchCtx& lc, const TcWgrad& t, int B, float* grad, int variant) {
  if (!tcw_batch_ok(t, B)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "wgrad_tc: batch must be a multiple of the images per box");
  WArgs a;
  a.B = B; a.H = t.GH; a.W = t.GW; a.CS = t.CS; a.CP = t.CP; a.P0 = t.P0;
  a.ntaps = t.ntaps;
  for (int i = 0; i < t.ntaps; ++i) a.taps[i] = t.taps[i];
  a.s_shift = t.s_shift; a.s_plain = t.s_plain;
  a.BW = t.BW; a.BH = t.BH; a.BB = t.BB; a.rows = t.rows;
  a.tiles_per_img = (t.BB > 1) ? 1 : cdiv(t.GH, t.BH);
  a.n_ptiles = (t.BB > 1) ? B / t.BB : B * a.tiles_per_img;
  a.cob = t.CS / 64;
  a.n_pairs = cdiv(t.ntaps * a.cob, 2);
  a.n_ci_tiles = t.CP / t.BN;
  const int base = a.n_pairs * a.n_ci_tiles;
  // one CTA per SM (192 KB of smem): aim at <= 2 full waves of 148 CTAs, never a ragged third one
  // (1x1 convs: a single wave -- their cost is dominated by the reduction traffic, which grows with the CTA count)
  int splits = ((t.s_plain == 1 ? 1 : 2) * 148) / base;
  if (splits > a.n_ptiles) splits = a.n_ptiles;
  if (splits < 1) splits = 1;
  a.tiles_per_split = cdiv(a.n_ptiles, splits);
  a.splits = cdiv(a.n_ptiles, a.tiles_per_split);
  a.zero_smem = t.rows < PIX ? 1 : 0;
  a.variant = variant;
  a.grad = grad;
  const double flops = t.flops_per_image * B;
  const double bytes = 4.0 * ((double)B * t.GH * t.GW * (t.CS + t.CP) + (double)t.ntaps * t.CS * t.CP);
  ProfScope ps_(lc, K_CONV_WGRAD, flops, bytes);
  const int grid = base * a.splits;
  if (t.BN == 128) return launch_impl<2>(lc, t, a, grid);
  return launch_impl<1>(lc, t, a, grid);
}

}  // namespace igm
