// tcgen05 convolution engine, CTA-pair version (`cta_group::2`, M = 256 per pair) for the stride-1 3x3 / 1x1
// convolutions with N % 128 == 0: reference src/models/ddpm.py:116, :134, :151-152 and their data gradients.
//
// STATUS: bring-up.  Compiled into the library and reachable ONLY through igm_debug_conv / igm_debug_conv_bench
// (engine = 3); no network path selects it.  Written after this round's GPU budget was spent: it has never run on
// hardware (gated tests: IGM_TEST_CONV_PAIR=1, tests/test_gpu_conv_tc.py::test_conv_pair_*).
//
// Why: an M = 128 tcgen05.mma of width N reads 4 KB of A + N * 32 B of B per K = 16 step in
// N / 2 cycles, i.e. 96-128 B/clk of the SM's 128 B/clk of shared-memory bandwidth before any TMA fill.  In a CTA pair
// each SM reads its own 128 rows of A but only HALF of B, and stages only half of the weight tile.
//
// Pair tile = two vertically adjacent 128-pixel M tiles x 128 output channels.  CTA r of the pair stages
//     A_hi, A_lo  of its own M tile (the per-tap TMA boxes of conv_tc.cu)
//     B'_r = [ w_hi[64 r .. 64 r + 63] ; w_lo[64 r .. 64 r + 63] ]          (128 rows)
// and the leader issues, per K = 16 step,
//     MMA 1:  a_hi x B'      M = 256, N = 256  ->  columns [hh(0..63) | hl(0..63) | hh(64..127) | hl(64..127)]
//     MMA 2:  a_lo x w_hi    M = 256, N = 128, the same B descriptor (the first 64 rows of each CTA's tile are its half
//             of w_hi), accumulator address + 64 columns  ->  lands on [hl(0..63) | hh(64..127)]: columns of the right
//             output channels; the epilogue adds a channel's hh and hl columns anyway.
// (Operand partitioning of cta_group::2 as CUTLASS encodes it for SM100_MMA_F16BF16_2x1SM_SS, cute/atom/mma_traits_sm100.hpp:
// CTA r supplies A rows r*M/2.., B rows r*N/2.. from the same descriptor offset, and holds D rows r*M/2.. x all N columns.)
// Barrier protocol (DeepGEMM-style): both CTAs run a TMA producer whose loads complete on the LEADER's `full` barrier
// (count 2: the leader's arrive.expect_tx for both CTAs' bytes + the peer's plain arrive); the leader's
// tcgen05.commit multicasts to both CTAs' `empty` / `acc_full` barriers; all eight epilogue warps arrive on the leader's
// `acc_empty`.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace igm {
namespace {

using namespace tc;

constexpr int BM = 128;
constexpr int KC = 64;
constexpr int UMMA_K = 16;
constexpr int BN = 128;                          // output channels per pair tile
constexpr int A_TILE_BYTES = BM * KC * 2;        // 16 KiB
constexpr int B_TILE_BYTES = 2 * 64 * KC * 2;    // B'_r: 64 rows of w_hi + 64 rows of w_lo = 16 KiB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + B_TILE_BYTES;
constexpr int STAGES = 4;
constexpr int ACC_COLS = 256;
constexpr int STORE_STAGE_BYTES = 4 * (4096 + 2048 + 2048);
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STORE_STAGE_BYTES + 1024 + 256;

struct PArgs {
  int B, H, W, K, K0, N, N0;
  int ntaps, Csrc;
  TcTap taps[kTcMaxTaps];
  int BH, BW, BB;
  int tiles_m, tiles_n, tiles_per_img, pair_tiles;
  int stage_tx_bytes;   // bytes ONE CTA's loads deliver per stage
  const float* bias;
  const float* add0; const float* add1;
  int want_split;
  float* gn_part; int gn_cpg, gn_slots;
  int dbg;   // IGM_PAIR_DEBUG bring-up bits: 1 = skip the a_lo x w_hi MMA (result = a_hi x (w_hi + w_lo): isolates the N split)
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-SM TMA loads: the bytes complete on the barrier at cluster address `bar` (the leader's)
__device__ __forceinline__ void tma2_load_5d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(void* dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2_512(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst_smem)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2_512(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in BOTH CTAs when all previously issued MMAs have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

template <int SEG>
__device__ __forceinline__ void chunk_stats32(const float (&v)[32], bool valid, int lane, float* dst) {
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

__global__ void __launch_bounds__(192, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
                const __grid_constant__ CUtensorMap ta1_hi, const __grid_constant__ CUtensorMap ta1_lo,
                const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo,
                const __grid_constant__ CUtensorMap to0, const __grid_constant__ CUtensorMap to1,
                const __grid_constant__ CUtensorMap to_hi, const __grid_constant__ CUtensorMap to_lo, const PArgs p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment as an OFFSET from the shared array (not through an integer): the pointer keeps its address space, so
  // the epilogue's staging accesses compile to LDS / STS instead of generic LD / ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* store_stage = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(store_stage + STORE_STAGE_BYTES);
  uint64_t* full = bars;                    // [STAGES]  both producers -> leader's MMA   (the peer's copy is unused)
  uint64_t* empty = bars + STAGES;          // [STAGES]  leader's MMA -> each CTA's producer
  uint64_t* acc_full = bars + 2 * STAGES;   // [2]       leader's MMA -> each CTA's epilogue
  uint64_t* acc_empty = acc_full + 2;       // [2]       all 8 epilogue warps -> leader's MMA (the peer's copy is unused)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 2);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc2_512(tmem_slot);
  tc_fence_before();
  cluster_sync_all();   // barrier inits and the TMEM allocation of BOTH CTAs are visible before anything is signalled
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int pair = (int)blockIdx.x >> 1, npairs = (int)gridDim.x >> 1;
  const int kchunks = p.K / KC;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&ta_hi); prefetch_tmap(&ta_lo); prefetch_tmap(&tb_hi); prefetch_tmap(&tb_lo);
      prefetch_tmap(&ta1_hi); prefetch_tmap(&ta1_lo);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < p.pair_tiles; tile += npairs) {
        const int tmp = tile / p.tiles_n, tn = tile - tmp * p.tiles_n;
        const int tm = 2 * tmp + (int)rank;      // a pair past the last M tile loads out-of-range boxes: zeros
        int b0, y0;
        if (p.BB > 1) { b0 = tm * p.BB; y0 = 0; }
        else { b0 = tm / p.tiles_per_img; y0 = (tm - b0 * p.tiles_per_img) * p.BH; }
        for (int ti = 0; ti < p.ntaps; ++ti) {
          const TcTap tp = p.taps[ti];
          for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * STAGE_BYTES;
            const uint32_t lbar = mapa_rank(smem_u32(&full[stage]), 0);   // the leader's barrier
            if (leader) mbar_expect_tx(&full[stage], 2u * (uint32_t)p.stage_tx_bytes);
            const int ch = kc * KC;
            if (ch < p.K0) {
              tma2_load_5d(st, &ta_hi, lbar, ch + tp.px * p.Csrc, tp.dx, tp.py, y0 + tp.dy, b0);
              tma2_load_5d(st + A_TILE_BYTES, &ta_lo, lbar, ch + tp.px * p.Csrc, tp.dx, tp.py, y0 + tp.dy, b0);
            } else {
              tma2_load_5d(st, &ta1_hi, lbar, ch - p.K0, tp.dx, tp.py, y0 + tp.dy, b0);
              tma2_load_5d(st + A_TILE_BYTES, &ta1_lo, lbar, ch - p.K0, tp.dx, tp.py, y0 + tp.dy, b0);
            }
            const int wrow = tn * BN + 64 * (int)rank;   // this CTA's half of the output channels
            tma2_load_2d(st + 2 * A_TILE_BYTES, &tb_hi, lbar, tp.wtap * p.K + ch, wrow);
            tma2_load_2d(st + 2 * A_TILE_BYTES + 64 * KC * 2, &tb_lo, lbar, tp.wtap * p.K + ch, wrow);
            if (!leader) mbar_arrive_cluster(lbar);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      // D = f32, A = B = bf16, K-major; N >> 3 at bits 17-22, M >> 4 at bits 24-28 (M = 256 across the pair)
      const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t s0 = smem_u32(smem);
      const uint64_t dA_hi0 = make_sw128_desc(s0), dA_lo0 = make_sw128_desc(s0 + A_TILE_BYTES);
      const uint64_t dB0 = make_sw128_desc(s0 + 2 * A_TILE_BYTES);
      constexpr uint32_t STAGE16 = (uint32_t)STAGE_BYTES >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      const int iters = p.ntaps * kchunks;
      for (int tile = pair; tile < p.pair_tiles; tile += npairs) {
        mbar_wait(&acc_empty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * ACC_COLS);
        uint32_t accum = 0;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t soff = (uint32_t)stage * STAGE16;
#pragma unroll
          for (int k = 0; k < KC / UMMA_K; ++k) {
            const uint32_t off = soff + (uint32_t)(k * UMMA_K * 2 >> 4);
            umma2_bf16(d_tmem, dA_hi0 + off, dB0 + off, idesc_2n, accum);      // [hh(0..63) | hl(0..63) | hh(64..127) | hl(64..127)]
            if (!(p.dbg & 1))
              umma2_bf16(d_tmem + 64u, dA_lo0 + off, dB0 + off, idesc_n, 1u);  // lh(0..63) onto hl(0..63), lh(64..127) onto hh(64..127)
            accum = 1u;
          }
          umma2_commit_both(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma2_commit_both(&acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else {
    // ---- epilogue warps 2..5 of BOTH CTAs: each CTA drains its own 128 accumulator rows ----
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int N1 = p.N - p.N0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = pair; tile < p.pair_tiles; tile += npairs) {
      const int tmp = tile / p.tiles_n, tn = tile - tmp * p.tiles_n;
      const int tm = 2 * tmp + (int)rank;
      int b0, y0;
      if (p.BB > 1) { b0 = tm * p.BB; y0 = 0; }
      else { b0 = tm / p.tiles_per_img; y0 = (tm - b0 * p.tiles_per_img) * p.BH; }
      const int bx = row % p.BW;
      const int r2 = row / p.BW;
      const int by = r2 % p.BH;
      const int bb = r2 / p.BH;
      const int oy = y0 + by, b = b0 + bb;
      const bool live = tm < p.tiles_m;            // the odd pair's second tile does not exist
      const bool valid = live && (bb < p.BB) && (oy < p.H) && (b < p.B);
      const int64_t opix = ((int64_t)b * p.H + oy) * p.W + bx;

      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * ACC_COLS);
      float gs = 0.f, gss = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BN && live; c0 += 32) {
        const uint32_t col = (c0 < 64) ? (uint32_t)c0 : (uint32_t)(128 + c0 - 64);   // hh columns of channels c0 .. c0+31
        float v[32];
        tmem_ld_32x32(t_base + col, v);
        {
          float w[32];   // hl columns of the same channels
          tmem_ld_32x32(t_base + col + 64u, w);
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
          const int slot = (oy * p.W + bx) >> 5;
          const int slot0 = __shfl_sync(0xffffffffu, slot, 0);
          const int bw = __shfl_sync(0xffffffffu, b, 0);
          const bool wok = __shfl_sync(0xffffffffu, valid ? 1 : 0, 0) != 0;
          float* dst = wok ? p.gn_part + (((int64_t)bw * p.gn_slots + slot0) * kGroups + n / p.gn_cpg) * 2 : nullptr;
          if (p.gn_cpg == 8) chunk_stats32<8>(v, valid, lane, dst);
          else if (p.gn_cpg == 16) chunk_stats32<16>(v, valid, lane, dst);
          else {
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
        // ---- TMA-store epilogue (the geometry check of the plan guarantees a warp's 32 rows are one output box) ----
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
        uint8_t* st_f = store_stage + q * 8192;
        uint8_t* st_h = st_f + 4096;
        uint8_t* st_l = st_f + 6144;
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st_f + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        const bool want_hi = p.want_split && first_half;
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
          const int r0 = q * 32;
          const int x0 = r0 % p.BW;
          const int rr = r0 / p.BW;
          const int yy = y0 + rr % p.BH, bb0 = b0 + rr / p.BH;
          if (first_half) tma_store_4d(&to0, st_f, n, x0, yy, bb0);
          else tma_store_4d(&to1, st_f, n - p.N0, x0, yy, bb0);
          if (want_hi) {
            tma_store_4d(&to_hi, st_h, n, x0, yy, bb0);
            tma_store_4d(&to_lo, st_l, n, x0, yy, bb0);
          }
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_rank(smem_u32(&acc_empty[as]), 0));   // the leader's barrier, from either CTA
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  cluster_sync_all();   // neither CTA frees TMEM or retires while the other may still signal it or its MMAs read its smem
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2_512(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn_p() {
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

// output [Bmax, H, W, C] as (C, W, H, B) with one epilogue warp's 32 pixels x 32 channels as the box (as conv_tc.cu)
int encode_out_p(Status& st, CUtensorMap* m, const void* ptr, int C, int H, int W, int Bmax, int wb, int hb, int ib, bool bf16) {
  auto enc = encode_fn_p();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es_b = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bmax};
  cuuint64_t strides[3] = {(cuuint64_t)C * es_b, (cuuint64_t)W * C * es_b, (cuuint64_t)H * W * C * es_b};
  cuuint32_t box[4] = {32u, (cuuint32_t)wb, (cuuint32_t)hb, (cuuint32_t)ib};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (pair conv output) failed");
  return IGM_OK;
}

bool out_geometry(const TcConv& t, int& wb, int& hb, int& ib) {
  if (t.sy != 1 || t.sx != 1 || t.oy_off != 0 || t.ox_off != 0 || t.out_H != t.H || t.out_W != t.W) return false;
  wb = t.BW < 32 ? t.BW : 32;
  if (32 % wb != 0 || t.BW % wb != 0) return false;
  hb = 32 / wb < t.BH ? 32 / wb : t.BH;
  if (t.BH % hb != 0) return false;
  ib = 32 / (wb * hb);
  if (wb * hb * ib != 32) return false;
  if (ib > 1 && (t.BB % ib != 0)) return false;
  if (ib == 1 && hb < 32 / wb) return false;
  return true;
}

}  // namespace

bool tc2_eligible(const TcConv& base) {
  int wb, hb, ib;
  return base.valid && base.nph == 1 && base.w_img_rows == 0 && base.N % BN == 0 && base.K % KC == 0 &&
         base.BB * base.BH * base.BW == BM && out_geometry(base, wb, hb, ib);
}

int tc2_plan(Status& st, TcConvPair& t, const TcConv& base) {
  t.valid = false;
  if (!tc2_eligible(base)) IGM_FAIL(st, IGM_ERR_INVALID, "plan not eligible for the CTA-pair tcgen05 conv");
  t.base = &base;
  auto enc = encode_fn_p();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  for (int which = 0; which < 2; ++which) {   // weights [N rows][ntaps*K cols]: boxes of 64 channels x 64 rows
    cuuint64_t dims[2] = {(cuuint64_t)base.KH * base.KW * base.K, (cuuint64_t)base.N};
    cuuint64_t strides[1] = {(cuuint64_t)base.KH * base.KW * base.K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, 64u};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(which ? &t.b_lo : &t.b_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, which ? (void*)base.w_lo : (void*)base.w_hi,
                     dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (pair conv weights) failed");
  }
  t.valid = true;
  return IGM_OK;
}

int launch_conv_tc2(const LaunchCtx& lc, const TcConvPair& tp, const TcRun& r) {
  if (!tp.valid || !tp.base || !tp.base->valid) IGM_FAIL(*lc.st, IGM_ERR_STATE, "pair conv plan not initialised");
  const TcConv& t = *tp.base;
  if (r.B < 1 || r.B > t.Bmax) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc2: bad batch");
  if (r.N0 <= 0 || r.N0 > t.N || r.N0 % 32 != 0 || (r.N0 < t.N && !r.out1)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc2: bad output split");
  PArgs a;
  a.B = r.B; a.H = t.H; a.W = t.W; a.K = t.K; a.K0 = t.K0; a.N = t.N; a.N0 = r.N0;
  a.ntaps = t.ntaps; a.Csrc = t.Csrc;
  for (int i = 0; i < t.ntaps; ++i) a.taps[i] = t.taps[i];
  a.BH = t.BH; a.BW = t.BW; a.BB = t.BB;
  a.tiles_per_img = (t.BB > 1) ? 1 : cdiv(t.H, t.BH);
  a.tiles_m = (t.BB > 1) ? cdiv(r.B, t.BB) : r.B * a.tiles_per_img;
  a.tiles_n = t.N / BN;
  a.pair_tiles = cdiv(a.tiles_m, 2) * a.tiles_n;
  a.stage_tx_bytes = 2 * A_TILE_BYTES + B_TILE_BYTES;
  a.bias = r.bias; a.add0 = r.add0; a.add1 = r.add1;
  a.want_split = r.hi0 ? 1 : 0;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("IGM_PAIR_DEBUG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
  }
  a.gn_part = nullptr; a.gn_cpg = 0; a.gn_slots = 0;
  if (r.gn_part) {
    if (!tc_gn_fusable(t, r.B) || r.N0 != t.N) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc2: GroupNorm statistics cannot be fused for this plan");
    a.gn_part = r.gn_part; a.gn_cpg = t.N / kGroups; a.gn_slots = tc_gn_slots(t);
  }
  int wb, hb, ib;
  if (!out_geometry(t, wb, hb, ib)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_tc2: output tile is not a TMA box");
  TcConv::OutMaps& om = tp.om;
  const int N1 = t.N - r.N0;
  if (om.p0 != r.out0 || om.n0 != r.N0) {
    IGM_TRY(encode_out_p(*lc.st, &om.m0, r.out0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, false));
    om.p0 = r.out0; om.n0 = r.N0;
    if (!om.p1) om.m1 = om.m0;
    if (!om.ph) { om.mh = om.m0; om.ml = om.m0; }
  }
  if (N1 > 0 && (om.p1 != r.out1 || om.n1 != N1)) {
    IGM_TRY(encode_out_p(*lc.st, &om.m1, r.out1, N1, t.out_H, t.out_W, t.Bmax, wb, hb, ib, false));
    om.p1 = r.out1; om.n1 = N1;
  }
  if (r.hi0 && (om.ph != r.hi0 || om.pl != r.lo0)) {
    IGM_TRY(encode_out_p(*lc.st, &om.mh, r.hi0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, true));
    IGM_TRY(encode_out_p(*lc.st, &om.ml, r.lo0, r.N0, t.out_H, t.out_W, t.Bmax, wb, hb, ib, true));
    om.ph = r.hi0; om.pl = r.lo0;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_done = true;
  }
  const double flops = 2.0 * r.B * t.H * t.W * (double)t.N * t.K * t.ntaps;
  const double bytes = 4.0 * ((double)r.B * t.H * t.W * (t.K + t.N * (r.add0 ? 2 : 1)) + (double)t.ntaps * t.K * t.N);
  ProfScope ps_(lc, r.kclass, flops, bytes);
  const int pairs = a.pair_tiles < 74 ? a.pair_tiles : 74;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = lc.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, conv_tc2_kernel, t.a_hi, t.a_lo, t.a1_hi, t.a1_lo, tp.b_hi, tp.b_lo, om.m0, om.m1, om.mh,
                                      om.ml, a);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
