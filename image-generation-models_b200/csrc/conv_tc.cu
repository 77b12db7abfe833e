// tcgen05 convolution engine for sm_100a: TMA-fed, TMEM-accumulated implicit GEMM.
//
// Serves the stride-1 3x3 and 1x1 convolutions of the U-Net and their data gradients
// (reference src/models/ddpm.py:116, :134, :151-152 and autograd), i.e. >90 % of the FLOPs.
//
// Precision: plain bf16 operands miss the 1e-3 parity bound, so every
// operand is carried as a bf16 hi + lo pair and the product is formed as
//     hi*hi + hi*lo + lo*hi            (fp32 accumulate in TMEM)
// which is fp32-accurate to ~2^-17 relative at 1/3 of the bf16 tensor rate.
//
// Structure (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer: per (tap, 64-channel chunk) loads the A_hi / A_lo activation tiles
//               with a 4-D tensor map (C, W, H, B) whose out-of-bounds zero fill implements the
//               conv padding, and the B_hi / B_lo weight tiles with a 2-D map; 128B swizzle.
//   warp 1      MMA issuer: 12 x tcgen05.mma (M=128, N=BN, K=16) per chunk into a TMEM accumulator;
//               tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns), + bias, + optional addend, fp32 NHWC
//               stores; double-buffered accumulators let it overlap the next tile's MMAs.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace igm {
namespace {

constexpr int BM = 128;       // output pixels per tile (TMEM lanes)
constexpr int KC = 64;        // channels per K chunk: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int A_TILE_BYTES = BM * KC * 2;   // 16 KiB

using namespace tc;

struct TcArgs {
  int B, H, W, K, K0, N, N0;
  int ntaps, Csrc;
  TcTap taps[kTcMaxTaps];
  int out_H, out_W, sy, sx, oy_off, ox_off;
  int BH, BW, BB;
  int tiles_m, tiles_n, tiles_per_img;
  int stage_tx_bytes;   // bytes one pipeline stage receives: 2 x (box rows x 128 B) + 2 x weight tile
  const float* bias;
  float* out0; float* out1;
  const float* add0; const float* add1;
  __nv_bfloat16* hi0; __nv_bfloat16* lo0;
  int ldh;              // row pitch (elements) of hi0 / lo0: N0, or larger when the output is a channel slice of a wider tensor
  float* gn_part; int gn_cpg, gn_slots;
  // TMA-store epilogue: each epilogue warp stages its 32 rows x 32 channels in swizzled shared memory and one
  // lane issues cp.async.bulk.tensor stores (a warp = box wb x hb x ib pixels of the [B, H, W, C] output)
  int tma_out, wb, hb, ib;
  // output-parity phases merged into one launch: phase ph owns taps [ph_tap0, ph_tap0 + ph_ntaps) and writes at (ph_oy, ph_ox)
  int nph, ph_tap0[4], ph_ntaps[4], ph_oy[4], ph_ox[4];
  int w_img_rows;   // > 0: every image has its own weight matrix (rows b * w_img_rows + n of the weight tensor)
};

// (sum, sum of squares) of the SEG-channel segments of a 32-column chunk, reduced over the warp's
// 32 pixels and written by lane 0: dst[seg][2]
template <int SEG>
__device__ __forceinline__ void gn_chunk_stats(const float (&v)[32], bool valid, int lane, float* dst) {
#pragma unroll
  for (int s0 = 0; s0 < 32; s0 += SEG) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      s += v[s0 + j];
      ss = fmaf(v[s0 + j], v[s0 + j], ss);
    }
    if (!valid) { s = 0.f; ss = 0.f; }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0 && dst) {
      dst[(s0 / SEG) * 2 + 0] = s;
      dst[(s0 / SEG) * 2 + 1] = ss;
    }
  }
}

template <int BN>
struct Cfg {
  static constexpr int B_TILE_BYTES = BN * KC * 2;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;
  static constexpr int STAGES = (BN == 128) ? 3 : 4;
  // BN = 64: tcgen05.mma at N = 64 issues at the N = 128 rate (measured 60 vs 64 cycles, tools/mma_rate_probe.cu),
  // so the hi and lo weight tiles (adjacent in smem) are fed as ONE N = 128 operand: columns [0,64) accumulate
  // a*w_hi, columns [64,128) a*w_lo, and two instructions (a_hi, a_lo) replace three; the epilogue adds the halves.
  // BN = 128: the same stacking for the a_hi operand only -- a_hi x [w_hi ; w_lo] as ONE N = 256 instruction plus
  // a_lo x w_hi at N = 128 cost the same 192 tensor cycles per k-step as three N = 128 instructions but read 20 KB
  // instead of 24 KB of shared memory (the engine is shared-memory-bandwidth bound: operand reads + TMA fills > 128 B/clk).
  // (BN = 64 used to issue a_lo x [w_hi ; w_lo] as well; the a_lo * w_lo half is 2^-18 relative and only cost shared-memory
  // reads: a_hi x [w_hi ; w_lo] at N = 128 plus a_lo x w_hi at N = 64 reads 14 KB per k-sub-step instead of 16 KB.)
  static constexpr bool STACKED = false;
  static constexpr int ACC_COLS = 2 * BN;
  static constexpr int TMEM_COLS = 2 * ACC_COLS;   // two accumulator stages (power of two >= 32)
  static constexpr int STORE_STAGE_BYTES = 4 * (4096 + 2048 + 2048);   // per epilogue warp: fp32 | bf16 hi | bf16 lo tiles
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
               const __grid_constant__ CUtensorMap ta1_hi, const __grid_constant__ CUtensorMap ta1_lo,
               const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo,
               const __grid_constant__ CUtensorMap to0, const __grid_constant__ CUtensorMap to1,
               const __grid_constant__ CUtensorMap to_hi, const __grid_constant__ CUtensorMap to_lo, const TcArgs p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the shared array (not through an integer): the pointer keeps its address space, so
  // the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* store_stage = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(store_stage + C::STORE_STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + C::STAGES;          // [STAGES]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * C::STAGES;   // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;          // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);   // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

  const int tiles_ph = p.tiles_m * p.tiles_n;
  const int num_tiles = tiles_ph * p.nph;
  const int kchunks = p.K / KC;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&ta_hi); prefetch_tmap(&ta_lo); prefetch_tmap(&tb_hi); prefetch_tmap(&tb_lo);
      prefetch_tmap(&ta1_hi); prefetch_tmap(&ta1_lo);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int ph = tile / tiles_ph, tt = tile - ph * tiles_ph;
        const int tm = tt / p.tiles_n, tn = tt - tm * p.tiles_n;
        int b0, y0;
        if (p.BB > 1) { b0 = tm * p.BB; y0 = 0; }
        else { b0 = tm / p.tiles_per_img; y0 = (tm - b0 * p.tiles_per_img) * p.BH; }
        for (int ti = p.ph_tap0[ph]; ti < p.ph_tap0[ph] + p.ph_ntaps[ph]; ++ti) {
          const TcTap tp = p.taps[ti];
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * C::STAGE_BYTES;
            mbar_expect_tx(&full[stage], (uint32_t)p.stage_tx_bytes);
            const int ch = kc * KC;
            // activation box: (channel, x, sub-lattice row py, y, image); px selects a channel block
            if (ch < p.K0) {
              tma_load_5d(st, &ta_hi, &full[stage], ch + tp.px * p.Csrc, tp.dx, tp.py, y0 + tp.dy, b0);
              tma_load_5d(st + A_TILE_BYTES, &ta_lo, &full[stage], ch + tp.px * p.Csrc, tp.dx, tp.py, y0 + tp.dy, b0);
            } else {   // second tensor of a channel concat
              tma_load_5d(st, &ta1_hi, &full[stage], ch - p.K0, tp.dx, tp.py, y0 + tp.dy, b0);
              tma_load_5d(st + A_TILE_BYTES, &ta1_lo, &full[stage], ch - p.K0, tp.dx, tp.py, y0 + tp.dy, b0);
            }
            const int wrow = tn * BN + b0 * p.w_img_rows;
            tma_load_2d(st + 2 * A_TILE_BYTES, &tb_hi, &full[stage], tp.wtap * p.K + ch, wrow);
            tma_load_2d(st + 2 * A_TILE_BYTES + C::B_TILE_BYTES, &tb_lo, &full[stage], tp.wtap * p.K + ch, wrow);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), K-major A/B,
      // N>>3 at bits 17-22, M>>4 at bits 24-28
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C::ACC_COLS >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // One thread issues every MMA: descriptors are formed by adding 16-byte-unit offsets to four descriptors built
      // once (the 14-bit start-address field never carries out below 256 KB), so only a handful of scalar
      // instructions separate two tcgen05.mma (the issue loop, not the tensor pipe, was the bottleneck otherwise).
      const uint32_t s0 = smem_u32(smem);
      const uint64_t dA_hi0 = make_sw128_desc(s0), dA_lo0 = make_sw128_desc(s0 + A_TILE_BYTES);
      const uint64_t dB_hi0 = make_sw128_desc(s0 + 2 * A_TILE_BYTES);
      const uint64_t dB_lo0 = make_sw128_desc(s0 + 2 * A_TILE_BYTES + C::B_TILE_BYTES);
      constexpr uint32_t STAGE16 = (uint32_t)C::STAGE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * C::ACC_COLS);
        uint32_t accum = 0;
        const int iters = p.ph_ntaps[tile / tiles_ph] * kchunks;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t soff = (uint32_t)stage * STAGE16;
#pragma unroll
          for (int k = 0; k < KC / UMMA_K; ++k) {
            const uint32_t off = soff + (uint32_t)(k * UMMA_K * 2 >> 4);   // 32 bytes inside the 128-byte swizzle row
            const uint64_t dah = dA_hi0 + off, dal = dA_lo0 + off;
            const uint64_t dbh = dB_hi0 + off, dbl = dB_lo0 + off;
            if (C::STACKED) {
              umma_bf16(d_tmem, dal, dbh, idesc, accum);   // [a_lo*w_hi | a_lo*w_lo]
              umma_bf16(d_tmem, dah, dbh, idesc, 1u);      // [a_hi*w_hi | a_hi*w_lo]
            } else {
              umma_bf16(d_tmem, dah, dbh, idesc, accum);     // [a_hi*w_hi | a_hi*w_lo]   (N = 2 BN)
              umma_bf16(d_tmem, dal, dbh, idesc_n, 1u);      //  a_lo*w_hi into the first half
            }
            accum = 1u;
          }
          umma_commit(&empty[stage]);   // frees this smem stage when the MMAs above retire
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&acc_full[as]);     // accumulator complete
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4 ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int N1 = p.N - p.N0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int ph = tile / tiles_ph, tt = tile - ph * tiles_ph;
      const int tm = tt / p.tiles_n, tn = tt - tm * p.tiles_n;
      int b0, y0;
      if (p.BB > 1) { b0 = tm * p.BB; y0 = 0; }
      else { b0 = tm / p.tiles_per_img; y0 = (tm - b0 * p.tiles_per_img) * p.BH; }
      // row -> (bb, by, bx) in TMA box order
      const int bx = row % p.BW;
      const int r2 = row / p.BW;
      const int by = r2 % p.BH;
      const int bb = r2 / p.BH;
      const int oy = y0 + by, b = b0 + bb;
      const bool valid = (bb < p.BB) && (oy < p.H) && (b < p.B);
      const int64_t opix = ((int64_t)b * p.out_H + (oy * p.sy + p.ph_oy[ph])) * p.out_W + (bx * p.sx + p.ph_ox[ph]);

      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * C::ACC_COLS);
      float gs = 0.f, gss = 0.f;   // running GroupNorm sums for groups wider than one 32-column chunk
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld_32x32(t_base + (uint32_t)c0, v);
        {
          float w[32];   // second accumulator half (products with w_lo)
          tmem_ld_32x32(t_base + (uint32_t)(BN + c0), w);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += w[j];
        }
        const int n = tn * BN + c0;
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
            v[j] += bv.x; v[j + 1] += bv.y; v[j + 2] += bv.z; v[j + 3] += bv.w;
          }
        }
        if (p.gn_part) {
          // GroupNorm partials of (conv + bias): this warp's 32 pixels lie in one image and one 32-pixel slot
          const int slot = (oy * p.W + bx) >> 5;     // uniform across the warp (taken from lane 0 below)
          const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
          const int bw = __shfl_sync(0xffffffffu, b, 0);
          const bool wok = __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
          float* dst = wok ? p.gn_part + (((int64_t)bw * p.gn_slots + slot0) * kGroups + n / p.gn_cpg) * 2 : nullptr;
          if (p.gn_cpg == 8) gn_chunk_stats<8>(v, valid, lane, dst);
          else if (p.gn_cpg == 16) gn_chunk_stats<16>(v, valid, lane, dst);
          else {
            // groups of >= 32 channels: accumulate chunk sums until the group is complete
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { s += v[j]; ss = fmaf(v[j], v[j], ss); }
            if (!valid) { s = 0.f; ss = 0.f; }
            gs += warp_sum(s);
            gss += warp_sum(ss);
            if (((n + 32) % p.gn_cpg) == 0) {
              if (lane == 0 && dst) { dst[0] = gs; dst[1] = gss; }
              gs = 0.f; gss = 0.f;
            }
          }
        }
        if (p.tma_out) {
          // ---- TMA-store epilogue: registers -> swizzled smem tile -> cp.async.bulk.tensor (coalesced, asynchronous) ----
          const bool first_half = n < p.N0;
          const float* ad = nullptr;
          if (valid) {
            if (first_half) ad = p.add0 ? p.add0 + opix * p.N0 + n : nullptr;
            else ad = p.add1 ? p.add1 + opix * N1 + (n - p.N0) : nullptr;
          }
          if (ad) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(ad + j));
              v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
            }
          }
          uint8_t* st_f = store_stage + q * 8192;          // 32 rows x 128 B, SWIZZLE_128B
          uint8_t* st_h = st_f + 4096;                     // 32 rows x 64 B, SWIZZLE_64B
          uint8_t* st_l = st_f + 6144;
          if (lane == 0) tma_store_wait_read<0>();         // the previous chunk's stores have drained this staging tile
          __syncwarp();
          const bool want_f = !first_half || p.out0 != nullptr;   // "lean" output: only the bf16 hi/lo staging copy is kept
          if (want_f) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(st_f + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          const bool want_hi = p.hi0 && first_half;
          if (want_hi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              __align__(16) uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_pair(v[8 * j + 2 * e], v[8 * j + 2 * e + 1], h[e], l[e]);
              const int off = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
              *reinterpret_cast<uint4*>(st_h + off) = *reinterpret_cast<const uint4*>(h);
              *reinterpret_cast<uint4*>(st_l + off) = *reinterpret_cast<const uint4*>(l);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // first pixel of this warp's 32 rows: row q*32 of the tile in (image, y, x) order
            const int r0 = q * 32;
            const int x0 = r0 % p.BW;
            const int r2 = r0 / p.BW;
            const int yy = y0 + r2 % p.BH, bb0 = b0 + r2 / p.BH;
            if (!first_half) tma_store_4d(&to1, st_f, n - p.N0, x0, yy, bb0);
            else if (want_f) tma_store_4d(&to0, st_f, n, x0, yy, bb0);
            if (want_hi) {
              tma_store_4d(&to_hi, st_h, n, x0, yy, bb0);
              tma_store_4d(&to_lo, st_l, n, x0, yy, bb0);
            }
            tma_store_commit();
          }
        } else
        if (valid) {
          // thread-per-pixel stores: each thread owns 128 contiguous bytes of its output row.  (A shared-
          // memory transpose to 8-lanes-per-row stores was measured 15-20 % SLOWER on every layer, r1h.)
          float* o;
          const float* ad;
          if (n < p.N0) { o = p.out0 ? p.out0 + opix * p.N0 + n : nullptr; ad = p.add0 ? p.add0 + opix * p.N0 + n : nullptr; }
          else { o = p.out1 + opix * N1 + (n - p.N0); ad = p.add1 ? p.add1 + opix * N1 + (n - p.N0) : nullptr; }
          if (ad) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(ad + j));
              v[j] += av.x; v[j + 1] += av.y; v[j + 2] += av.z; v[j + 3] += av.w;
            }
          }
          if (o) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
          if (p.hi0 && n < p.N0) {
            // bf16 hi/lo copy of the output row segment: the next tensor-core conv reads it directly
            __nv_bfloat16* oh = p.hi0 + opix * p.ldh + n;
            __nv_bfloat16* ol = p.lo0 + opix * p.ldh + n;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              __align__(16) uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_pair(v[j + 2 * e], v[j + 2 * e + 1], h[e], l[e]);
              *reinterpret_cast<uint4*>(oh + j) = *reinterpret_cast<const uint4*>(h);
              *reinterpret_cast<uint4*>(ol + j) = *reinterpret_cast<const uint4*>(l);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (p.tma_out && lane == 0) tma_store_wait<0>();   // outstanding bulk stores complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ src, int64_t M, int C, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int cdst, int coff) {
  pdl_wait();
  const int c4n = C >> 2;
  const int64_t total = M * c4n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / c4n;
    const int c = (int)(i - m * c4n) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(src + m * C + c));
    store_split4(hi, lo, m * cdst + coff + c, v);
  }
}

// split + column sums in one pass over a gradient tensor (bias gradient of the conv whose dY is being staged):
// L = C/4 lanes per row, 256/L rows in flight per CTA; per-CTA partial sums -> one atomic per channel per CTA
__global__ void __launch_bounds__(256) split_colsum_kernel(const float* __restrict__ src, int64_t M, int C,
                                                           __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                                           float* __restrict__ colsum) {
  __shared__ float4 red[256];
  pdl_wait();
  const int L = C >> 2, R = 256 / L;
  const int c4 = threadIdx.x % L, slot = threadIdx.x / L;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  const int64_t stride = (int64_t)gridDim.x * R;
  for (int64_t m0 = (int64_t)blockIdx.x * R + slot; m0 < M; m0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {   // four independent 128-bit loads in flight per thread
      const int64_t m = m0 + u * stride;
      v[u] = (m < M) ? __ldg(reinterpret_cast<const float4*>(src + m * C + c4 * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t m = m0 + u * stride;
      if (m >= M) break;
      const int64_t o = m * C + c4 * 4;
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      store_split4(hi, lo, o, v[u]);
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < L) {
    float4 t = red[threadIdx.x];
    for (int r = 1; r < R; ++r) {
      const float4 v = red[r * L + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    float* d = colsum + threadIdx.x * 4;
    atomicAdd(d + 0, t.x); atomicAdd(d + 1, t.y); atomicAdd(d + 2, t.z); atomicAdd(d + 3, t.w);
  }
}

__global__ void merge_bf16_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                  float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
}

__global__ void pack_weight_tc_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, int taps, int K, int N, int64_t sk, int64_t sn,
                                      int flip) {
  const int64_t total = (int64_t)N * taps * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int64_t r = i / K;
    const int tap = (int)(r % taps);
    const int n = (int)(r / taps);
    const int ts = flip ? (taps - 1 - tap) : tap;
    const float x = src[k * sk + n * sn + ts];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

__global__ void __launch_bounds__(256) pack_jobs_kernel(const PackJob* __restrict__ jobs, int n_jobs, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int lo = 0, hi = n_jobs - 1;
    while (lo < hi) {   // last job whose begin <= i
      const int mid = (lo + hi + 1) >> 1;
      if (jobs[mid].begin <= i) lo = mid; else hi = mid - 1;
    }
    const PackJob& j = jobs[lo];
    const int64_t e = i - j.begin;
    if (j.dst_f) {
      const int n = (int)(e % j.N);
      const int64_t r = e / j.N;
      const int k = (int)(r % j.K);
      const int tap = (int)(r / j.K);
      j.dst_f[e] = __ldg(j.src + k * j.sk + n * j.sn + tap);
    } else {
      const int k = (int)(e % j.K);
      const int64_t r = e / j.K;
      const int tap = (int)(r % j.taps);
      const int n = (int)(r / j.taps);
      const int ts = j.flip ? (j.taps - 1 - tap) : tap;
      const float x = __ldg(j.src + k * j.sk + n * j.sn + ts);
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      reinterpret_cast<__nv_bfloat16*>(j.dst_hi)[e] = h;
      reinterpret_cast<__nv_bfloat16*>(j.dst_lo)[e] = __float2bfloat16_rn(x - __bfloat162float(h));
    }
  }
}

// Tiled re-pack of the tensor-core weight layouts: a CTA moves a 32(n) x 32(k) x taps tile through shared memory,
// so both the fp32 source (runs of 32*taps contiguous floats) and the bf16 hi/lo destinations (32 contiguous k per
// (n, tap) row) are accessed in whole lines.  job.begin = first CTA of the job; cta_job[cta] = job index.
__global__ void __launch_bounds__(256) pack_tiles_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ cta_job) {
  extern __shared__ float ptile[];   // [32 outer][32 * taps + 1]
  const PackJob j = jobs[cta_job[blockIdx.x]];
  const int t = blockIdx.x - (int)j.begin;
  const int tk = j.K / 32;
  const int n0 = (t / tk) * 32, k0 = (t % tk) * 32;
  const int KK = j.taps, L = 32 * KK, pitch = L + 1;
  const bool outer_n = j.sn > j.sk;            // fprop layout of a Conv2d: each n owns a contiguous run of k x taps
  const int64_t so = outer_n ? j.sn : j.sk;    // stride between outer rows
  const float* base = j.src + (outer_n ? (int64_t)n0 * j.sn + (int64_t)k0 * j.sk : (int64_t)k0 * j.sk + (int64_t)n0 * j.sn);
  // 128-bit loads, four in flight per thread (runs are 16-byte aligned: k0, n0 are multiples of 32)
  const int L4 = L >> 2;
  for (int i0 = threadIdx.x; i0 < 32 * L4; i0 += 4 * 256) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i < 32 * L4) {
        const int o = i / L4, r4 = i - o * L4;
        v[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)o * so) + r4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      if (i < 32 * L4) {
        const int o = i / L4, r4 = i - o * L4;
        float* d = ptile + o * pitch + r4 * 4;
        d[0] = v[u].x; d[1] = v[u].y; d[2] = v[u].z; d[3] = v[u].w;
      }
    }
  }
  __syncthreads();
  const int k = threadIdx.x & 31, w = threadIdx.x >> 5;
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(j.dst_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(j.dst_lo);
  for (int pr = w; pr < 32 * KK; pr += 8) {    // (n, tap) rows of the destination
    const int n = pr / KK, tap = pr - n * KK;
    const float x = outer_n ? ptile[n * pitch + k * KK + tap] : ptile[k * pitch + n * KK + tap];
    const int td = j.flip ? (KK - 1 - tap) : tap;
    const int64_t d = ((int64_t)(n0 + n) * KK + td) * j.K + k0 + k;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    hi[d] = h;
    lo[d] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

}  // namespace

static bool tile_grid_ok(int GH, int GW) { return GW >= 1 && GW <= BM && GH >= 1; }

bool tc_eligible(int K, int N, int H, int W, int KH) {
  if (K < KC || K % KC != 0) return false;
  if (N < 64 || N % 64 != 0) return false;
  if (KH != 1 && KH != 3) return false;
  return tile_grid_ok(H, W);
}

bool tc_strided_eligible(int K, int N, int SH, int SW, int KH) {
  if (K < KC || K % KC != 0 || N < 64 || N % 64 != 0) return false;
  if (KH != 3 && KH != 4) return false;
  if (SH % 2 != 0 || SW % 2 != 0) return false;
  return tile_grid_ok(SH / 2, SW / 2);
}

namespace {

// Tile shape over a GH x GW grid and the weight descriptors; shared by every plan flavour.
int plan_common(Status& st, TcConv& t, int K, int K0, int N, int GH, int GW, int Bmax, int wtaps,
                __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, bool per_image = false) {
  auto enc = get_encode_fn();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  t.K = K; t.K0 = K0; t.N = N; t.H = GH; t.W = GW; t.Bmax = Bmax;
  t.w_hi = w_hi; t.w_lo = w_lo;
  t.BW = GW;
  if (GH * GW <= BM) { t.BH = GH; t.BB = BM / (GH * GW); }
  else { t.BH = BM / GW; t.BB = 1; }
  if (t.BB > Bmax) t.BB = Bmax;
  if (per_image) t.BB = 1;   // a tile must not span images when every image has its own weights
  t.w_img_rows = per_image ? N : 0;
  t.BN = (N % 128 == 0) ? 128 : 64;
  if (t.BN == 128) {
    // one CTA per SM: prefer 64-wide tiles when 128-wide ones leave most of a wave empty
    const int64_t tiles_m = (t.BB > 1) ? cdiv(Bmax, t.BB) : (int64_t)Bmax * cdiv(GH, t.BH);
    const int64_t t128 = tiles_m * (N / 128), t64 = tiles_m * (N / 64);
    const int64_t cost128 = cdiv64(t128, 148) * 2, cost64 = cdiv64(t64, 148);
    if (cost64 < cost128) t.BN = 64;
    // experiment switch IGM_TC_BN=64: 64-wide tiles also where the 128-wide plan leaves a CTA with at most two tiles
    // (its fill and its 128-column epilogue are then not hidden behind MMAs of a following tile)
    static const int bn_mode = [] { const char* e = getenv("IGM_TC_BN"); return e ? atoi(e) : 0; }();
    if (bn_mode == 64 && t128 <= 2 * 148 && !per_image) t.BN = 64;
  }
  // weights: [N rows][wtaps*K cols], K-major
  for (int which = 0; which < 2; ++which) {
    cuuint64_t dims[2] = {(cuuint64_t)wtaps * K, (cuuint64_t)N * (per_image ? Bmax : 1)};
    cuuint64_t strides[1] = {(cuuint64_t)wtaps * K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)t.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(which ? &t.b_lo : &t.b_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, which ? (void*)w_lo : (void*)w_hi,
                     dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (weights) failed");
  }
  return IGM_OK;
}

// Rank-5 activation descriptor (channel, x, sub-lattice row, y, image) of a [Bmax, SH, SW, C] bf16 tensor.
// s2d = false: plain view (C, SW, 1, SH, B).  s2d = true: stride-2 view (2C, SW/2, 2, SH/2, B), where
// channel index px*C + c addresses pixel column 2x + px and the third coordinate py row 2y + py.
// `pitch` (elements, plain view only): distance between consecutive pixels when the C channels are a slice of a wider
// tensor (the q / k / v thirds of the [M, 384] to_qkv output); 0 = dense.
int encode_act(Status& st, CUtensorMap* m, void* ptr, int C, int SH, int SW, int Bmax, bool s2d, const TcConv& t, int pitch = 0) {
  auto enc = get_encode_fn();
  if (pitch <= 0) pitch = C;
  const cuuint64_t rowB = (cuuint64_t)SW * (s2d ? C : pitch) * 2;
  cuuint64_t dims[5], strides[4];
  if (!s2d) {
    dims[0] = C; dims[1] = SW; dims[2] = 1; dims[3] = SH; dims[4] = Bmax;
    strides[0] = (cuuint64_t)pitch * 2; strides[1] = rowB; strides[2] = rowB; strides[3] = rowB * SH;
  } else {
    dims[0] = 2 * (cuuint64_t)C; dims[1] = SW / 2; dims[2] = 2; dims[3] = SH / 2; dims[4] = Bmax;
    strides[0] = (cuuint64_t)C * 4; strides[1] = rowB; strides[2] = 2 * rowB; strides[3] = rowB * SH;
  }
  cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)t.BW, 1u, (cuuint32_t)t.BH, (cuuint32_t)t.BB};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (activations) failed");
  return IGM_OK;
}

}  // namespace

int tc_plan(Status& st, TcConv& t, int K, int N, int H, int W, int Bmax, int KH, int pad, __nv_bfloat16* a_hi,
            __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, int K0, __nv_bfloat16* a1_hi,
            __nv_bfloat16* a1_lo) {
  t.valid = false;
  if (!tc_eligible(K, N, H, W, KH)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the tcgen05 engine");
  if (K0 <= 0 || !a1_hi) K0 = K;
  if (K0 % KC != 0 || (K - K0) % KC != 0) IGM_FAIL(st, IGM_ERR_INVALID, "concat split must be a multiple of 64 channels");
  IGM_TRY(plan_common(st, t, K, K0, N, H, W, Bmax, KH * KH, w_hi, w_lo));
  t.KH = t.KW = KH; t.pad = pad;
  t.ntaps = KH * KH;
  for (int ky = 0; ky < KH; ++ky)
    for (int kx = 0; kx < KH; ++kx) t.taps[ky * KH + kx] = TcTap{kx - pad, ky - pad, 0, 0, ky * KH + kx};
  t.Csrc = K0;
  t.out_H = H; t.out_W = W; t.sy = t.sx = 1; t.oy_off = t.ox_off = 0;
  const bool two = K0 < K;
  IGM_TRY(encode_act(st, &t.a_hi, a_hi, K0, H, W, Bmax, false, t));
  IGM_TRY(encode_act(st, &t.a_lo, a_lo, K0, H, W, Bmax, false, t));
  IGM_TRY(encode_act(st, &t.a1_hi, two ? (void*)a1_hi : (void*)a_hi, two ? K - K0 : K0, H, W, Bmax, false, t));
  IGM_TRY(encode_act(st, &t.a1_lo, two ? (void*)a1_lo : (void*)a_lo, two ? K - K0 : K0, H, W, Bmax, false, t));
  t.valid = true;
  return IGM_OK;
}

int tc_plan_img(Status& st, TcConv& t, int K, int N, int H, int W, int Bmax, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
                __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, int a_pitch) {
  t.valid = false;
  if (!tc_eligible(K, N, H, W, 1)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the tcgen05 engine");
  IGM_TRY(plan_common(st, t, K, K, N, H, W, Bmax, 1, w_hi, w_lo, /*per_image=*/true));
  t.KH = t.KW = 1; t.pad = 0;
  t.ntaps = 1;
  t.taps[0] = TcTap{0, 0, 0, 0, 0};
  t.Csrc = K;
  t.out_H = H; t.out_W = W; t.sy = t.sx = 1; t.oy_off = t.ox_off = 0;
  IGM_TRY(encode_act(st, &t.a_hi, a_hi, K, H, W, Bmax, false, t, a_pitch));
  IGM_TRY(encode_act(st, &t.a_lo, a_lo, K, H, W, Bmax, false, t, a_pitch));
  t.a1_hi = t.a_hi; t.a1_lo = t.a_lo;
  t.valid = true;
  return IGM_OK;
}

int tc_plan_strided(Status& st, TcConv& t, int K, int N, int SH, int SW, int Bmax, int KH, int pad,
                    __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo) {
  t.valid = false;
  if (!tc_strided_eligible(K, N, SH, SW, KH)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the strided tcgen05 conv");
  const int GH = SH / 2, GW = SW / 2;
  IGM_TRY(plan_common(st, t, K, K, N, GH, GW, Bmax, KH * KH, w_hi, w_lo));
  t.KH = t.KW = KH; t.pad = pad;
  t.ntaps = KH * KH;
  // source row 2*oy - pad + ky = 2*(oy + dy) + py  with  py = (ky - pad) mod 2,  dy = floor((ky - pad) / 2)
  auto split2 = [](int v, int& d, int& ph) { ph = ((v % 2) + 2) % 2; d = (v - ph) / 2; };
  for (int ky = 0; ky < KH; ++ky)
    for (int kx = 0; kx < KH; ++kx) {
      TcTap tp;
      split2(ky - pad, tp.dy, tp.py);
      split2(kx - pad, tp.dx, tp.px);
      tp.wtap = ky * KH + kx;
      t.taps[ky * KH + kx] = tp;
    }
  t.Csrc = K;
  t.out_H = GH; t.out_W = GW; t.sy = t.sx = 1; t.oy_off = t.ox_off = 0;
  IGM_TRY(encode_act(st, &t.a_hi, a_hi, K, SH, SW, Bmax, true, t));
  IGM_TRY(encode_act(st, &t.a_lo, a_lo, K, SH, SW, Bmax, true, t));
  t.a1_hi = t.a_hi; t.a1_lo = t.a_lo;
  t.valid = true;
  return IGM_OK;
}

int tc_plan_phase(Status& st, TcConv& t, int K, int N, int GH, int GW, int Bmax, int KH, int pad, int py, int px,
                  __nv_bfloat16* a_hi, __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo) {
  t.valid = false;
  if (K < KC || K % KC != 0 || N < 64 || N % 64 != 0 || !tile_grid_ok(GH, GW) || (KH != 3 && KH != 4))
    IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the phase tcgen05 conv");
  IGM_TRY(plan_common(st, t, K, K, N, GH, GW, Bmax, KH * KH, w_hi, w_lo));
  t.KH = t.KW = KH; t.pad = pad;
  // output row 2a + py receives source row (2a + py + pad - ky) / 2 = a + (py + pad - ky) / 2 for the ky of right parity
  t.ntaps = 0;
  for (int ky = 0; ky < KH; ++ky) {
    if (((py + pad - ky) % 2 + 2) % 2 != 0) continue;
    for (int kx = 0; kx < KH; ++kx) {
      if (((px + pad - kx) % 2 + 2) % 2 != 0) continue;
      TcTap tp;
      tp.dy = (py + pad - ky) / 2; tp.dx = (px + pad - kx) / 2; tp.py = tp.px = 0;
      tp.wtap = ky * KH + kx;
      t.taps[t.ntaps++] = tp;
    }
  }
  if (t.ntaps == 0) IGM_FAIL(st, IGM_ERR_INVALID, "phase without taps");
  t.Csrc = K;
  t.out_H = 2 * GH; t.out_W = 2 * GW; t.sy = t.sx = 2; t.oy_off = py; t.ox_off = px;
  IGM_TRY(encode_act(st, &t.a_hi, a_hi, K, GH, GW, Bmax, false, t));
  IGM_TRY(encode_act(st, &t.a_lo, a_lo, K, GH, GW, Bmax, false, t));
  t.a1_hi = t.a_hi; t.a1_lo = t.a_lo;
  t.valid = true;
  return IGM_OK;
}

// [Bmax, H, W, C] output tensor as a rank-4 map (C, W, H, B) whose box is one epilogue warp's 32 pixels x 32 channels
static int encode_out(Status& st, CUtensorMap* m, const void* ptr, int C, int H, int W, int Bmax, int wb, int hb, int ib,
                      bool bf16, int pitch = 0) {
  auto enc = get_encode_fn();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es_b = bf16 ? 2 : 4;
  if (pitch <= 0) pitch = C;   // > C: the output is a channel slice of a wider [.., pitch] tensor
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bmax};
  cuuint64_t strides[3] = {(cuuint64_t)pitch * es_b, (cuuint64_t)W * pitch * es_b, (cuuint64_t)H * W * pitch * es_b};
  cuuint32_t box[4] = {32u, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)ib};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (conv output) failed");
  return IGM_OK;
}

// Can the epilogue of plan `t` leave through TMA stores?  A warp's 32 tile rows must form one box of the output.
static bool tma_out_geometry(const TcConv& t, int& wb, int& hb, int& ib) {
  if (t.sy != 1 || t.sx != 1 || t.oy_off != 0 || t.ox_off != 0 || t.out_H != t.H || t.out_W != t.W) return false;
  wb = t.BW < 32 ? t.BW : 32;
  if (32 % wb != 0 || t.BW % wb != 0) return false;
  hb = 32 / wb < t.BH ? 32 / wb : t.BH;
  if (t.BH % hb != 0) return false;
  ib = 32 / (wb * hb);
  if (wb * hb * ib != 32) return false;
  if (ib > 1 && (t.BB % ib != 0)) return false;
  if (ib == 1 && hb < 32 / wb) return false;   // a warp would straddle two images inside one box row range
  return true;
}

int tc_plan_phases4(Status& st, TcConv& t, int K, int N, int GH, int GW, int Bmax, int KH, int pad, __nv_bfloat16* a_hi,
                    __nv_bfloat16* a_lo, __nv_bfloat16* w_hi, __nv_bfloat16* w_lo) {
  t.valid = false;
  if (K < KC || K % KC != 0 || N < 64 || N % 64 != 0 || !tile_grid_ok(GH, GW) || (KH != 3 && KH != 4))
    IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the phase tcgen05 conv");
  IGM_TRY(plan_common(st, t, K, K, N, GH, GW, Bmax, KH * KH, w_hi, w_lo));
  t.KH = t.KW = KH; t.pad = pad;
  t.ntaps = 0;
  t.nph = 4;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph / 2, px = ph % 2;
    t.ph_tap0[ph] = t.ntaps; t.ph_oy[ph] = py; t.ph_ox[ph] = px;
    for (int ky = 0; ky < KH; ++ky) {
      if (((py + pad - ky) % 2 + 2) % 2 != 0) continue;
      for (int kx = 0; kx < KH; ++kx) {
        if (((px + pad - kx) % 2 + 2) % 2 != 0) continue;
        if (t.ntaps >= kTcMaxTaps) IGM_FAIL(st, IGM_ERR_INVALID, "too many taps");
        TcTap tp;
        tp.dy = (py + pad - ky) / 2; tp.dx = (px + pad - kx) / 2; tp.py = tp.px = 0;
        tp.wtap = ky * KH + kx;
        t.taps[t.ntaps++] = tp;
      }
    }
    t.ph_ntaps[ph] = t.ntaps - t.ph_tap0[ph];
    if (t.ph_ntaps[ph] == 0) IGM_FAIL(st, IGM_ERR_INVALID, "phase without taps");
  }
  t.Csrc = K;
  t.out_H = 2 * GH; t.out_W = 2 * GW; t.sy = t.sx = 2; t.oy_off = t.ox_off = 0;
  IGM_TRY(encode_act(st, &t.a_hi, a_hi, K, GH, GW, Bmax, false, t));
  IGM_TRY(encode_act(st, &t.a_lo, a_lo, K, GH, GW, Bmax, false, t));
  t.a1_hi = t.a_hi; t.a1_lo = t.a_lo;
  t.valid = true;
  return IGM_OK;
}

template <int BN>
static int launch_tc_impl(const LaunchCtx& lc, const TcConv& t, const TcArgs& a, int num_tiles) {
  using C = Cfg<BN>;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  int grid = num_tiles < 148 ? num_tiles : 148;
  cudaError_t le = launch_pdl(conv_tc_kernel<BN>, dim3(grid), dim3(192), (size_t)C::SMEM_BYTES, lc.stream, t.a_hi, t.a_lo, t.a1_hi,
                              t.a1_lo, t.b_hi, t.b_lo, t.om.m0, t.om.m1, t.om.mh, t.om.ml, a);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

bool tc_gn_fusable(const TcConv& t, int B) {
  (void)B;
  if (!t.valid || t.sy != 1 || t.sx != 1 || t.out_H != t.H || t.out_W != t.W) return false;
  if (t.BB * t.BH * t.BW != BM) return false;                 // full tiles only
  const int hw = t.H * t.W;
  if (hw % 32 != 0) return false;                              // a warp's 32 pixels stay inside one image
  if (t.BB == 1 && (t.H % t.BH) != 0) return false;
  const int cpg = t.N / kGroups;
  if (t.N % kGroups != 0 || (cpg != 8 && cpg != 16 && cpg % 32 != 0) || cpg > t.BN) return false;
  return true;
}

int launch_conv_tc(const LaunchCtx& lc, const TcConv& t, const TcRun& r) {
  if (!t.valid) IGM_FAIL(*lc.st, IGM_ERR_STATE, "tcgen05 conv plan not initialised");
  if (r.B < 1 || r.B > t.Bmax) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc: bad batch");
  if (r.N0 <= 0 || r.N0 > t.N || r.N0 % 32 != 0 || (r.N0 < t.N && !r.out1))
    IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc: bad output split");
  TcArgs a;
  a.B = r.B; a.H = t.H; a.W = t.W; a.K = t.K; a.K0 = t.K0; a.N = t.N; a.N0 = r.N0;
  a.ntaps = t.ntaps; a.Csrc = t.Csrc;
  for (int i = 0; i < t.ntaps; ++i) a.taps[i] = t.taps[i];
  a.out_H = t.out_H; a.out_W = t.out_W; a.sy = t.sy; a.sx = t.sx; a.oy_off = t.oy_off; a.ox_off = t.ox_off;
  a.BH = t.BH; a.BW = t.BW; a.BB = t.BB;
  a.nph = t.nph;
  a.w_img_rows = t.w_img_rows;
  for (int i = 0; i < 4; ++i) { a.ph_tap0[i] = t.ph_tap0[i]; a.ph_ntaps[i] = t.ph_ntaps[i]; a.ph_oy[i] = t.ph_oy[i]; a.ph_ox[i] = t.ph_ox[i]; }
  if (t.nph == 1) { a.ph_tap0[0] = 0; a.ph_ntaps[0] = t.ntaps; a.ph_oy[0] = t.oy_off; a.ph_ox[0] = t.ox_off; }
  a.tiles_per_img = (t.BB > 1) ? 1 : cdiv(t.H, t.BH);
  a.tiles_m = (t.BB > 1) ? cdiv(r.B, t.BB) : r.B * a.tiles_per_img;
  a.tiles_n = t.N / t.BN;
  a.stage_tx_bytes = 2 * (t.BB * t.BH * t.BW * KC * 2) + 2 * (t.BN * KC * 2);
  a.bias = r.bias; a.out0 = r.out0; a.out1 = r.out1; a.add0 = r.add0; a.add1 = r.add1;
  a.hi0 = r.hi0; a.lo0 = r.lo0;
  a.ldh = r.ld_hi > 0 ? r.ld_hi : r.N0;
  if (!r.out0 && !r.hi0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc: no output tensor");
  if (r.ld_hi > 0 && (r.ld_hi < r.N0 || r.ld_hi % 8 != 0)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc: bad hi/lo output pitch");
  a.gn_part = nullptr; a.gn_cpg = 0; a.gn_slots = 0;
  if (r.gn_part) {
    if (!tc_gn_fusable(t, r.B) || r.N0 != t.N) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc: GroupNorm statistics cannot be fused for this plan");
    a.gn_part = r.gn_part; a.gn_cpg = t.N / kGroups; a.gn_slots = tc_gn_slots(t);
  }
  // TMA-store epilogue when the tile geometry allows it (IGM_TMA_STORE=0 keeps the thread-per-row stores)
  static const bool tma_store_off = [] { const char* e = getenv("IGM_TMA_STORE"); return e && e[0] == '0'; }();
  a.tma_out = 0; a.wb = a.hb = a.ib = 0;
  int wb, hb, ib;
  if (!tma_store_off && tma_out_geometry(t, wb, hb, ib)) {
    TcConv::OutMaps& om = t.om;
    const int N1 = t.N - r.N0;
    if (r.out0 && (om.p0 != r.out0 || om.n0 != r.N0)) {
      IGM_TRY(encode_out(*lc.st, &om.m0, r.out0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, false));
      om.p0 = r.out0; om.n0 = r.N0;
      if (!om.p1) om.m1 = om.m0;
      if (!om.ph) { om.mh = om.m0; om.ml = om.m0; }
    }
    if (N1 > 0 && (om.p1 != r.out1 || om.n1 != N1)) {
      IGM_TRY(encode_out(*lc.st, &om.m1, r.out1, N1, t.out_H, t.out_W, t.Bmax, wb, hb, ib, false));
      om.p1 = r.out1; om.n1 = N1;
    }
    if (r.hi0 && (om.ph != r.hi0 || om.pl != r.lo0 || om.nh != r.N0 || om.ldh != a.ldh)) {
      IGM_TRY(encode_out(*lc.st, &om.mh, r.hi0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, true, a.ldh));
      IGM_TRY(encode_out(*lc.st, &om.ml, r.lo0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, true, a.ldh));
      om.ph = r.hi0; om.pl = r.lo0; om.nh = r.N0; om.ldh = a.ldh;
      if (!om.p0) { om.m0 = om.mh; if (!om.p1) om.m1 = om.mh; }   // lean output: the fp32 maps are never dereferenced
    }
    a.tma_out = 1; a.wb = wb; a.hb = hb; a.ib = ib;
  }
  const double flops = 2.0 * r.B * t.H * t.W * (double)t.N * t.K * t.ntaps;   // (merged phases: ntaps = all taps of all phases)
  const double bytes = 4.0 * ((double)r.B * t.H * t.W * (t.K + t.N * (r.add0 ? 2 : 1)) + (double)t.ntaps * t.K * t.N);
  ProfScope ps_(lc, r.kclass, flops, bytes);
  const int num_tiles = a.tiles_m * a.tiles_n * a.nph;
  if (t.BN == 128) return launch_tc_impl<128>(lc, t, a, num_tiles);
  return launch_tc_impl<64>(lc, t, a, num_tiles);
}

int launch_split_bf16(const LaunchCtx& lc, const float* src, int64_t M, int C, __nv_bfloat16* hi, __nv_bfloat16* lo,
                      int cdst, int coff) {
  if (C % 4 != 0 || coff % 4 != 0 || cdst % 4 != 0) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "split: channels must be multiples of 4");
  const int64_t total = M * (C / 4);
  int blocks = (int)cdiv64(total, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps_(lc, K_ELEM, 2.0 * M * C, 8.0 * M * C);
  { cudaError_t le_ = launch_pdl(split_bf16_kernel, dim3(blocks), dim3(256), (size_t)(0), lc.stream, src, M, C, hi, lo, cdst, coff); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_merge_bf16(const LaunchCtx& lc, const __nv_bfloat16* hi, const __nv_bfloat16* lo, float* dst, int64_t n) {
  int blocks = (int)cdiv64(n, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  merge_bf16_kernel<<<blocks, 256, 0, lc.stream>>>(hi, lo, dst, n);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

bool split_colsum_ok(int C) { return C >= 32 && C <= 1024 && (C & (C - 1)) == 0; }

int launch_split_bf16_colsum(const LaunchCtx& lc, const float* src, int64_t M, int C, __nv_bfloat16* hi,
                             __nv_bfloat16* lo, float* colsum) {
  if (!split_colsum_ok(C)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "split+colsum: channels must be a power of two in [32, 1024]");
  const int R = 256 / (C >> 2);
  int blocks = (int)cdiv64(M, (int64_t)R * 8);   // >= 8 rows per thread
  if (blocks > 148 * 2) blocks = 148 * 2;        // one atomic per channel and CTA: keep the same-address count low
  if (blocks < 1) blocks = 1;
  ProfScope ps_(lc, K_ELEM, 3.0 * M * C, 8.0 * M * C);
  { cudaError_t le_ = launch_pdl(split_colsum_kernel, dim3(blocks), dim3(256), (size_t)(0), lc.stream, src, M, C, hi, lo, colsum); if (le_ != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le_)); }
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_pack_jobs(const LaunchCtx& lc, const PackJob* d_jobs, int n_jobs, int64_t total) {
  if (n_jobs <= 0 || total <= 0) return IGM_OK;
  int blocks = (int)cdiv64(total, 256 * 4);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps_(lc, K_PACK, 0.0, 8.0 * total);
  pack_jobs_kernel<<<blocks, 256, 0, lc.stream>>>(d_jobs, n_jobs, total);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_pack_tiles(const LaunchCtx& lc, const PackJob* d_jobs, const int* d_cta_job, int n_ctas, int max_taps, double elems) {
  if (n_ctas <= 0) return IGM_OK;
  const int smem = 32 * (32 * max_taps + 1) * (int)sizeof(float);
  static int attr_smem = 0;
  if (smem > 48 * 1024 && smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(pack_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = smem;
  }
  ProfScope ps_(lc, K_PACK, 0.0, 8.0 * elems);
  pack_tiles_kernel<<<n_ctas, 256, smem, lc.stream>>>(d_jobs, d_cta_job);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

int launch_pack_weight_tc(const LaunchCtx& lc, const float* src, __nv_bfloat16* hi, __nv_bfloat16* lo, int taps, int K,
                          int N, int64_t sk, int64_t sn, int flip) {
  const int64_t total = (int64_t)N * taps * K;
  int blocks = (int)cdiv64(total, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  ProfScope ps_(lc, K_PACK, 0.0, 8.0 * total);
  pack_weight_tc_kernel<<<blocks, 256, 0, lc.stream>>>(src, hi, lo, taps, K, N, sk, sn, flip);
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
