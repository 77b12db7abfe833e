// tcgen05 weight-gradient engine for the stride-1 3x3 convolutions, halo-reuse version (sm_100a).
//
//   dW[co][ci][ky][kx] = sum_{b,y,x} dY[b, y + 1 - ky, x + 1 - kx, co] * X[b, y, x, ci]      (pad = 1)
//
// wgrad_tc.cu fetches one shifted dY box per tap and is bound by L2 -> shared-memory traffic (every
// operand byte is re-read 9 times).  Here one zero-padded dY tile ("halo" tile: W+2 columns, the padding
// columns and rows come from the TMA out-of-bounds fill) is staged ONCE per pixel tile and all taps read
// it through shifted shared-memory descriptors: with the 128-byte swizzle keyed on absolute address bits a
// descriptor may start at any 128-byte row (tools/desc_probe.cu), and in the padded-linear pixel order
// p = y * (W+2) + (x+1) a tap is a constant row offset  dy * (W+2) + dx.
//
// GEMM shape per CTA (the reduction axis is the padded pixel axis, both operands MN-major):
//   A (M = 128) = X tile: two 64-channel ci blocks, or for Cin = 64 the hi and lo copies stacked
//   B (N = 192) = the three kx taps of one ky: three 64-channel co blocks of the SAME dY tile whose
//                 "leading byte offset" is one pixel row (128 B)
//   D           = one 128 x 192 fp32 accumulator per ky in TMEM (a CTA owns 1 or 2 ky -> 192/384 columns)
// bf16x3 split precision as everywhere else (hi*hi + hi*lo + lo*hi; the stacked form adds lo*lo).
// Results leave through 128-bit vector reductions (red.global.add.v4.f32) into a per-layer workspace
// [ky][kx][ci][co]; wgrad_halo_finalize_kernel folds the workspaces of all layers into PyTorch's OIHW
// gradients once per backward pass and re-zeroes them.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace igm {
namespace {

using namespace tc;

constexpr int UMMA_K = 16;
constexpr int PAD_BYTES = 2048;   // zeroed guard before / after each dY tile (shifted descriptors read up to 16 rows past it)

struct HArgs {
  int B, H, W, PW, BH, SR;        // image, padded width W+2, image rows per pixel tile, dY rows per tile (BH+1)
  int prows, prows_pad, ksteps;   // P rows per tile, rounded up to 16, MMA k-steps per tile
  int Cin, Cout, P0;
  int stacked, stages;
  int p_blk_bytes, s_box_bytes, stage_bytes, s_off;   // smem layout of one stage
  int tiles_per_img, n_tiles;
  int n_items;                    // (co block, ci unit) pairs
  int n_ci_units;
  int splits_a, tps_a, n_cta_a;   // type A CTAs: ky = 0, 1
  int splits_b, tps_b;            // type B CTAs: ky = 2
  float* ws;                      // [3][3][Cin][Cout]
  int dbg;                        // bring-up switches (IGM_HALO_DEBUG): 1 = no MMAs, 2 = no reductions, 4 = one k-step
};

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// The MMA issue loop runs in ONE thread: at 96 cycles per M128 x N192 x K16 instruction there is room for only a
// few dozen scalar instructions per MMA, so every shared-memory descriptor is formed by adding a 16-byte-unit
// offset to a descriptor built once (the start-address field is the low 14 bits and never carries out: the
// whole window is below 256 KB), and the (ky, product) loops are unrolled at compile time.
template <bool STACKED, int NKY>
__device__ __forceinline__ void mma_loop(const HArgs& p, uint8_t* smem, uint64_t* full, uint64_t* empty,
                                         uint64_t* acc_full, uint32_t tmem_base, int n_my) {
  // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = 192, M = 128
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                         ((uint32_t)(192 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t s0 = smem_u32(smem);
  // stage-0 descriptors; stacked: A = [hi | lo] of one ci block (one operand); else A_hi = blocks 0,1, A_lo = blocks 2,3
  const uint64_t dA_hi = make_sw128_mn_desc(s0, (uint32_t)p.p_blk_bytes, 1024);
  const uint64_t dA_lo = make_sw128_mn_desc(s0 + 2 * p.p_blk_bytes, (uint32_t)p.p_blk_bytes, 1024);
  // dY tile one pixel row before its first row (dx = -1); the three kx taps are the N blocks, 128 B apart
  const uint64_t dB_hi = make_sw128_mn_desc(s0 + p.s_off + PAD_BYTES - 128, 128, 1024);
  const uint64_t dB_lo = make_sw128_mn_desc(s0 + p.s_off + 2 * PAD_BYTES + p.s_box_bytes - 128, 128, 1024);
  // ky = ky0 + k reads dY rows shifted by dy = 1 - ky relative to the tile start row:
  // type A (NKY = 2): ky 0 -> +PW rows, ky 1 -> 0;  type B: ky 2 -> 0
  const uint32_t row0 = (NKY == 2) ? (uint32_t)(p.PW * 128) >> 4 : 0u;
  const uint32_t stage16 = (uint32_t)p.stage_bytes >> 4;
  const int ksteps = (p.dbg & 1) ? 0 : ((p.dbg & 4) ? 1 : p.ksteps);
  int stage = 0;
  uint32_t phase = 0;
  uint32_t accum = 0;
  for (int it = 0; it < n_my; ++it) {
    mbar_wait(&full[stage], phase);
    tc_fence_after();
    uint32_t off = (uint32_t)stage * stage16;
    for (int ks = 0; ks < ksteps; ++ks, off += (UMMA_K * 128) >> 4) {
      const uint64_t dah = dA_hi + off, dal = dA_lo + off;
#pragma unroll
      for (int k = 0; k < NKY; ++k) {
        const uint32_t boff = off + ((NKY == 2 && k == 0) ? row0 : 0u);
        const uint64_t dbh = dB_hi + boff, dbl = dB_lo + boff;
        const uint32_t d = tmem_base + (uint32_t)(k * 192);
        if (STACKED) {
          umma_bf16(d, dah, dbl, idesc, accum);   // [x_hi ; x_lo] * dy_lo
          umma_bf16(d, dah, dbh, idesc, 1u);      // [x_hi ; x_lo] * dy_hi
        } else {
          umma_bf16(d, dal, dbh, idesc, accum);
          umma_bf16(d, dah, dbl, idesc, 1u);
          umma_bf16(d, dah, dbh, idesc, 1u);
        }
      }
      accum = 1u;
    }
    umma_commit(&empty[stage]);
    if (++stage == p.stages) { stage = 0; phase ^= 1; }
  }
  umma_commit(acc_full);
}

__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ CUtensorMap ts_hi, const __grid_constant__ CUtensorMap ts_lo,
                  const __grid_constant__ CUtensorMap tp_hi, const __grid_constant__ CUtensorMap tp_lo,
                  const __grid_constant__ CUtensorMap tp1_hi, const __grid_constant__ CUtensorMap tp1_lo, const HArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the shared array (not through an integer): the pointer keeps its address space, so
  // the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + 4;
  uint64_t* acc_full = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- work decode: CTA -> (type, item, split) ----
  int bid = blockIdx.x;
  const bool type_a = bid < p.n_cta_a;
  if (!type_a) bid -= p.n_cta_a;
  const int splits = type_a ? p.splits_a : p.splits_b;
  const int tps = type_a ? p.tps_a : p.tps_b;
  const int split = bid % splits;
  const int item = bid / splits;
  const int ci_unit = item % p.n_ci_units;
  const int co0 = (item / p.n_ci_units) * 64;
  const int ci0 = ci_unit * (p.stacked ? 64 : 128);
  const int nky = type_a ? 2 : 1;
  const int ky0 = type_a ? 0 : 2;
  const int t0 = split * tps, t1 = min(t0 + tps, p.n_tiles);
  const int n_my = t1 - t0;
  const int n_pblk = p.stacked ? 2 : 4;
  // Zero the whole pipeline memory once: the guard regions around the dY tiles, the rounding slack after every
  // TMA box and the X rows [prows, prows_pad) are never written afterwards, and all of them are read by
  // shifted / padded descriptors (always against a zero factor, but stale NaN bit patterns would poison it).
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = p.stages * p.stage_bytes / 16;
    if (!(p.dbg & 16))
      for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (n_my > 0) {
    if (warp == 0) {
      if (lane == 0) {
        prefetch_tmap(&ts_hi); prefetch_tmap(&ts_lo); prefetch_tmap(&tp_hi); prefetch_tmap(&tp_lo);
        const bool second = ci0 >= p.P0;   // channel concat: this ci unit lives in the second tensor
        const CUtensorMap* mp_hi = second ? &tp1_hi : &tp_hi;
        const CUtensorMap* mp_lo = second ? &tp1_lo : &tp_lo;
        const int cc = second ? ci0 - p.P0 : ci0;
        const uint32_t tx_bytes = (uint32_t)(n_pblk * p.prows * 128 + 2 * p.SR * p.PW * 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int t = t0; t < t1; ++t) {
          const int b = t / p.tiles_per_img;
          const int y0 = (t - b * p.tiles_per_img) * p.BH;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* st = smem + stage * p.stage_bytes;
          if (p.dbg & 8) {   // bring-up: no loads, the MMAs run on the zeroed tiles
            mbar_arrive(&full[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_expect_tx(&full[stage], tx_bytes);
          // X tile: padded columns -1 .. W of image rows y0 .. y0+BH-1
          if (p.stacked) {
            tma_load_5d(st, mp_hi, &full[stage], cc, -1, 0, y0, b);
            tma_load_5d(st + p.p_blk_bytes, mp_lo, &full[stage], cc, -1, 0, y0, b);
          } else {
            tma_load_5d(st, mp_hi, &full[stage], cc, -1, 0, y0, b);
            tma_load_5d(st + p.p_blk_bytes, mp_hi, &full[stage], cc + 64, -1, 0, y0, b);
            tma_load_5d(st + 2 * p.p_blk_bytes, mp_lo, &full[stage], cc, -1, 0, y0, b);
            tma_load_5d(st + 3 * p.p_blk_bytes, mp_lo, &full[stage], cc + 64, -1, 0, y0, b);
          }
          // dY halo tile: rows ys0 .. ys0+SR-1;  type A (ky 0,1 -> dy +1,0) starts at y0, type B (ky 2 -> dy -1) at y0-1
          const int ys0 = type_a ? y0 : y0 - 1;
          uint8_t* sb = st + p.s_off;
          tma_load_5d(sb + PAD_BYTES, &ts_hi, &full[stage], co0, -1, 0, ys0, b);
          tma_load_5d(sb + 2 * PAD_BYTES + p.s_box_bytes, &ts_lo, &full[stage], co0, -1, 0, ys0, b);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        if (p.stacked) { if (type_a) mma_loop<true, 2>(p, smem, full, empty, acc_full, tmem_base, n_my);
                         else mma_loop<true, 1>(p, smem, full, empty, acc_full, tmem_base, n_my); }
        else { if (type_a) mma_loop<false, 2>(p, smem, full, empty, acc_full, tmem_base, n_my);
               else mma_loop<false, 1>(p, smem, full, empty, acc_full, tmem_base, n_my); }
      }
    } else {
      // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4; lane = input channel ----
      const int q = warp & 3;
      const int row = q * 32 + lane;
      const int ci = ci0 + (p.stacked ? (row & 63) : row);
      mbar_wait(acc_full, 0);
      tc_fence_after();
      // The CTAs of one split group finish together and reduce into the same workspace: rotate the order in
      // which each CTA walks its (ky, 32-column chunk) pieces so that they do not hammer the same sectors.
      // Stacked operand: accumulator rows r and r + 64 are the hi and lo halves of the SAME input channel; the
      // lo-half warps hand their values over through the (now idle) pipeline memory, so each element is reduced once.
      float* xbuf = reinterpret_cast<float*>(smem);   // [2 buffers][2 warps][32 values][32 lanes]
      float* tbuf = xbuf + 4096;                      // [4 warps][32][33] transpose tiles
      const int nchunks = nky * 6;
      for (int it = 0; it < nchunks; ++it) {
        const int piece = (it + split) % nchunks;
        const int k = piece / 6, c0 = (piece % 6) * 32;
        const int ky = ky0 + k;
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(k * 192);
        float v[32];
        tmem_ld_32x32(t_base + (uint32_t)c0, v);
        if (p.stacked) {
          float* xb = xbuf + ((it & 1) * 2 + (q & 1)) * 1024 + lane;
          if (q >= 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) xb[j * 32] = v[j];
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (q >= 2) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += xb[j * 32];
        }
        const int kx = 2 - c0 / 64;   // N block 0 <-> dx = -1 <-> kx = 2
        // A lane owns one accumulator row (input channel) x 32 output channels; reducing straight from there makes
        // every warp instruction touch 32 different rows (32 L2 transactions).  Transpose through shared memory so that
        // eight lanes cover one row's 128 contiguous bytes: 4 transactions per instruction.
        float* tb = tbuf + q * (32 * 33);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) tb[lane * 33 + j] = v[j];
        __syncwarp();
        const int ci_w = ci - lane;   // input channel of this warp's row 0
        float* dst = p.ws + (((int64_t)(ky * 3 + kx) * p.Cin + ci_w) * p.Cout + co0 + (c0 & 63));
        const int rr = lane >> 3, cc = (lane & 7) * 4;
#pragma unroll
        for (int r0 = 0; r0 < 32; r0 += 4) {
          const float* src = tb + (r0 + rr) * 33 + cc;
          if (!(p.dbg & 2)) red_add_v4(dst + (int64_t)(r0 + rr) * p.Cout + cc, src[0], src[1], src[2], src[3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// grad[co][ci][tap] += ws[tap][ci][co]; ws = 0.   One CTA per 32(ci) x 32(co) tile of one layer (transposed
// through shared memory so that both sides move whole 128-byte lines); CTA -> layer through a per-CTA table.
__global__ void __launch_bounds__(1024) wgrad_halo_finalize_kernel(const HaloFinJob* __restrict__ jobs,
                                                                   const int* __restrict__ cta_job, int cta_base) {
  __shared__ float tile[9][32][33];
  const int cta = blockIdx.x + cta_base;   // a launch may cover the CTA range of a single layer
  const HaloFinJob j = jobs[cta_job[cta]];
  const int t = cta - j.tile_begin;
  const int tco = j.Cout / 32;
  const int ci0 = (t / tco) * 32, co0 = (t % tco) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float v[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) v[tap] = j.ws[((int64_t)tap * j.Cin + ci0 + ty) * j.Cout + co0 + tx];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    j.ws[((int64_t)tap * j.Cin + ci0 + ty) * j.Cout + co0 + tx] = 0.f;
    tile[tap][ty][tx] = v[tap];
  }
  __syncthreads();
  // OIHW gradient of this tile: 32 co rows x (32 ci x 9 taps = 288 contiguous floats), updated with 128-bit accesses
  for (int i = threadIdx.x; i < 32 * 72; i += 1024) {
    const int row = i / 72, q4 = i - row * 72;
    float4* gp = reinterpret_cast<float4*>(j.grad + ((int64_t)(co0 + row) * j.Cin + ci0) * 9) + q4;
    float4 g = *gp;
    float* ge = reinterpret_cast<float*>(&g);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = q4 * 4 + e;
      const int ci = f / 9, tap = f - ci * 9;
      ge[e] += tile[tap][ci][row];
    }
    *gp = g;
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn_h() {
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

// rank-5 view (channel, x, 1, y, image) of a [Bmax, H, W, C] bf16 tensor with a (64, PW, 1, rows, 1) box
int encode_halo(Status& st, CUtensorMap* m, void* ptr, int C, int H, int W, int Bmax, int PW, int rows) {
  auto enc = encode_fn_h();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t rowB = (cuuint64_t)W * C * 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)Bmax};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, rowB, rowB, rowB * H};
  cuuint32_t box[5] = {64u, (cuuint32_t)PW, 1u, (cuuint32_t)rows, 1u};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (halo wgrad) failed");
  return IGM_OK;
}

int round_up(int v, int a) { return (v + a - 1) / a * a; }

// shared-memory bytes of one pipeline stage for BH image rows per tile
void stage_layout(int W, int BH, bool stacked, int& prows, int& prows_pad, int& p_blk, int& s_box, int& s_off, int& stage) {
  const int PW = W + 2;
  prows = BH * PW;
  prows_pad = round_up(prows, 16);
  p_blk = round_up(prows_pad * 128, 1024);
  s_box = round_up((BH + 1) * PW * 128, 1024);
  s_off = (stacked ? 2 : 4) * p_blk;
  stage = s_off + 2 * (PAD_BYTES + s_box) + PAD_BYTES;
}

constexpr int kSmemBudget = 220 * 1024;

}  // namespace

bool tcwh_eligible(int Cin, int Cout, int H, int W, int KH) {
  if (KH != 3) return false;
  if (Cout < 64 || Cout % 64 != 0) return false;
  if (!(Cin == 64 || (Cin >= 128 && Cin % 128 == 0))) return false;
  if (W < 2 || W + 2 > 256 || H < 1) return false;
  int prows, pp, pb, sb, so, stg;
  stage_layout(W, 1, Cin == 64, prows, pp, pb, sb, so, stg);
  return 2 * stg + 1024 + 256 <= kSmemBudget;
}

int tcwh_plan(Status& st, TcWgradHalo& t, int Cin, int Cout, int H, int W, int Bmax, __nv_bfloat16* dy_hi,
              __nv_bfloat16* dy_lo, __nv_bfloat16* x_hi, __nv_bfloat16* x_lo, int C0, __nv_bfloat16* x1_hi,
              __nv_bfloat16* x1_lo, float* ws) {
  t.valid = false;
  if (!tcwh_eligible(Cin, Cout, H, W, 3)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the halo wgrad engine");
  if (C0 <= 0 || !x1_hi) C0 = Cin;
  t.stacked = Cin == 64;
  const int unit = t.stacked ? 64 : 128;
  if (C0 % unit != 0 || (Cin - C0) % unit != 0) IGM_FAIL(st, IGM_ERR_INVALID, "halo wgrad: concat split must be a multiple of the ci unit");
  t.Cin = Cin; t.Cout = Cout; t.P0 = C0; t.H = H; t.W = W; t.Bmax = Bmax; t.ws = ws;
  // image rows per tile: the largest of 8/4/2/1 (not more than needed) that leaves room for >= 2 stages, 3 when cheap
  int best = 1;
  for (int bh : {8, 4, 2, 1}) {
    if (bh > 1 && bh / 2 >= H) continue;
    int prows, pp, pb, sb, so, stg;
    stage_layout(W, bh, t.stacked, prows, pp, pb, sb, so, stg);
    if (3 * stg + 1024 + 256 <= kSmemBudget) { best = bh; break; }   // TMA latency needs >= 3 tiles in flight
  }
  t.BH = best;
  stage_layout(W, t.BH, t.stacked, t.prows, t.prows_pad, t.p_blk_bytes, t.s_box_bytes, t.s_off, t.stage_bytes);
  t.stages = (kSmemBudget - 1024 - 256) / t.stage_bytes;
  if (t.stages > 4) t.stages = 4;
  const bool two = C0 < Cin;
  const int PW = W + 2;
  IGM_TRY(encode_halo(st, &t.s_hi, dy_hi, Cout, H, W, Bmax, PW, t.BH + 1));
  IGM_TRY(encode_halo(st, &t.s_lo, dy_lo, Cout, H, W, Bmax, PW, t.BH + 1));
  IGM_TRY(encode_halo(st, &t.p_hi, x_hi, C0, H, W, Bmax, PW, t.BH));
  IGM_TRY(encode_halo(st, &t.p_lo, x_lo, C0, H, W, Bmax, PW, t.BH));
  IGM_TRY(encode_halo(st, &t.p1_hi, two ? (void*)x1_hi : (void*)x_hi, two ? Cin - C0 : C0, H, W, Bmax, PW, t.BH));
  IGM_TRY(encode_halo(st, &t.p1_lo, two ? (void*)x1_lo : (void*)x_lo, two ? Cin - C0 : C0, H, W, Bmax, PW, t.BH));
  t.valid = true;
  return IGM_OK;
}

int launch_wgrad_halo(const LaunchCtx& lc, const TcWgradHalo& t, int B) {
  if (!t.valid || B < 1 || B > t.Bmax) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "halo wgrad: bad plan or batch");
  HArgs a;
  a.B = B; a.H = t.H; a.W = t.W; a.PW = t.W + 2; a.BH = t.BH; a.SR = t.BH + 1;
  a.prows = t.prows; a.prows_pad = t.prows_pad; a.ksteps = t.prows_pad / UMMA_K;
  a.Cin = t.Cin; a.Cout = t.Cout; a.P0 = t.P0;
  a.stacked = t.stacked ? 1 : 0; a.stages = t.stages;
  a.p_blk_bytes = t.p_blk_bytes; a.s_box_bytes = t.s_box_bytes; a.stage_bytes = t.stage_bytes; a.s_off = t.s_off;
  a.tiles_per_img = cdiv(t.H, t.BH);
  a.n_tiles = B * a.tiles_per_img;
  a.n_ci_units = t.Cin / (t.stacked ? 64 : 128);
  a.n_items = (t.Cout / 64) * a.n_ci_units;
  // one wave of 148 CTAs: type A (two ky) costs twice a type B CTA per tile, so it gets 2/3 of the SMs
  auto split_for = [&](int ctas, int& splits, int& tps) {
    splits = ctas / a.n_items;
    if (splits < 1) splits = 1;
    if (splits > a.n_tiles) splits = a.n_tiles;
    tps = cdiv(a.n_tiles, splits);
    splits = cdiv(a.n_tiles, tps);
  };
  split_for(98, a.splits_a, a.tps_a);
  split_for(50, a.splits_b, a.tps_b);
  a.n_cta_a = a.n_items * a.splits_a;
  a.ws = t.ws;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("IGM_HALO_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
  }
  const int grid = a.n_cta_a + a.n_items * a.splits_b;
  const int smem = t.stages * t.stage_bytes + 1024 + 256;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = smem;
  }
  const double flops = 2.0 * B * t.H * t.W * (double)t.Cin * t.Cout * 9;
  const double bytes = 4.0 * ((double)B * t.H * t.W * (t.Cin + t.Cout) + 9.0 * t.Cin * t.Cout);
  ProfScope ps_(lc, K_CONV_WGRAD, flops, bytes);
  cudaError_t le = launch_pdl(wgrad_halo_kernel, dim3(grid), dim3(192), (size_t)smem, lc.stream, t.s_hi, t.s_lo, t.p_hi, t.p_lo,
                              t.p1_hi, t.p1_lo, a);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_wgrad_halo_finalize(const LaunchCtx& lc, const HaloFinJob* d_jobs, const int* d_cta_job, int n_ctas, double elems,
                               int cta_base) {
  if (n_ctas <= 0) return IGM_OK;
  ProfScope ps_(lc, K_CONV_WGRAD, elems, 16.0 * elems);
  wgrad_halo_finalize_kernel<<<n_ctas, 1024, 0, lc.stream>>>(d_jobs, d_cta_job, cta_base);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
