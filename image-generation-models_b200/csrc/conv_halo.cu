// tcgen05 convolution engine, halo-reuse version, for the stride-1 3x3 convolutions (and their data gradients)
// on wide images (W >= 28): reference src/models/ddpm.py:116 (Block's Conv2d 3x3 p1) and its autograd.
//
// STATUS: bring-up.  Compiled into the library and reachable ONLY through igm_debug_conv(engine = 2); no
// network path selects it yet (it was written after this round's GPU budget was spent; the gated test
// tests/test_gpu_conv_tc.py::test_conv_halo_* runs it when IGM_TEST_CONV_HALO=1).
//
// conv_tc.cu fetches one shifted activation box per (tap, 64-channel chunk): nine TMA fills of 32 KB per
// k-step and tile, which together with the tensor core's own operand reads oversubscribes the 128 B/clk of
// shared-memory bandwidth.  Here ONE zero-padded tile - rows y0-1 .. y0+BH, columns
// -1 .. W of the NHWC activation, the padding supplied by the TMA out-of-bounds fill - is staged per
// (band, 64-channel chunk) in padded-linear order s = r * (W+2) + c, and tap (ky, kx) is the constant row
// offset ky * (W+2) + kx applied to the start address of the K-major shared-memory descriptor
// (tools/desc_probe.cu: with SWIZZLE_128B a descriptor may start at any 128-byte row).  The index algebra
// is the executable model tools/halo_fprop_model.py (tests/test_halo_model.py).
//
//   GEMM M index m in [0, BH * (W+2)): output pixel (y0 + m / (W+2), m % (W+2)); columns W, W+1 are junk.
//   A band is two M = 128 MMA tiles (BH = 256 / (W+2) rows); both share every weight tile.
//   N = 64 output channels per work item, weights stacked [w_hi ; w_lo] as in conv_tc.cu's BN = 64 recipe:
//       a_hi x [w_hi ; w_lo] (N = 128)  +  a_lo x w_hi (N = 64)      per K = 16 sub-step
//   TMEM: 2 accumulator stages x 2 tiles x 128 columns = 512 columns.
//
// One persistent CTA per SM, 192 threads: warp 0 = TMA producer (activation ring of 2 tiles, weight ring of
// 2-4 taps), warp 1 = MMA issuer, warps 2-5 = epilogue.  A junk accumulator row depends only on its own
// (junk) operand row, so reads past the tile need no zeroed guard; junk rows never leave the SM: the
// epilogue stores each image row of a warp's 32 accumulator rows with its own TMA store whose x start may be
// negative - the out-of-range part of the box (x < 0, x >= W) is clipped by the hardware.
#include "conv_tc.cuh"
#include "tc_ptx.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace igm {
namespace {

using namespace tc;

constexpr int KC = 64;            // channels per K chunk = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int BN = 64;            // output channels per work item
constexpr int TILE_COLS = 128;    // TMEM columns per M tile: [a*w_hi | a*w_lo]
constexpr int W_STAGE_BYTES = 2 * BN * KC * 2;   // w_hi tile followed by w_lo tile
constexpr int CW = 16;            // accumulator columns per epilogue chunk
constexpr int STORE_F32_BYTES = 32 * CW * 4;    // per epilogue warp: 32 rows x 64 B fp32 staging (SWIZZLE_64B)
constexpr int STORE_BF16_BYTES = 32 * CW * 2;   // 32 rows x 32 B bf16 staging (hi, then lo; no swizzle)
constexpr int kSmemLimit = 227 * 1024;

struct HFArgs {
  int B, H, W, PW, BH;
  int K, K0, N;
  int bands_per_img, n_items, tiles_n, kchunks;
  int a_tile_bytes, a_stages, w_stages, store_warp_bytes;
  uint32_t a_tx_bytes;
  const float* bias;
  const float* add0;
  int want_split;           // also emit the output as bf16 hi/lo
  int tma_rows;             // 1: clipped per-image-row TMA stores; 0 (IGM_HALO_TMA_STORE=0): thread-per-row stores of the valid rows
  int base_offset;          // IGM_HALO_BASE_OFFSET=1: fill the descriptor's matrix-base-offset field ((start >> 7) & 7) for the
                            // shifted activation descriptors (tools/desc_probe.cu found 0 correct; bring-up alternative)
  float* out0; __nv_bfloat16* hi0; __nv_bfloat16* lo0;
  float* gn_part; int gn_cpg, gn_slots;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[CW]) {
  uint32_t r[CW];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < CW; ++i) v[i] = __uint_as_float(r[i]);
}

// (sum, sum of squares) of the SEG-channel segments of a CW-column chunk over the warp's valid rows, written by lane 0
// (the helper of conv_tc.cu's epilogue for narrower chunks)
template <int SEG>
__device__ __forceinline__ void chunk_stats(const float (&v)[CW], bool valid, int lane, float* dst) {
#pragma unroll
  for (int s0 = 0; s0 < CW; s0 += SEG) {
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < SEG; ++j) {
      s += v[s0 + j];
      ss = fmaf(v[s0 + j], v[s0 + j], ss);
    }
    if (!valid) { s = 0.f; ss = 0.f; }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) {
      dst[(s0 / SEG) * 2 + 0] = s;
      dst[(s0 / SEG) * 2 + 1] = ss;
    }
  }
}

__global__ void __launch_bounds__(192, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap ta_hi, const __grid_constant__ CUtensorMap ta_lo,
                 const __grid_constant__ CUtensorMap ta1_hi, const __grid_constant__ CUtensorMap ta1_lo,
                 const __grid_constant__ CUtensorMap tb_hi, const __grid_constant__ CUtensorMap tb_lo,
                 const __grid_constant__ CUtensorMap to0, const __grid_constant__ CUtensorMap to_hi,
                 const __grid_constant__ CUtensorMap to_lo, const HFArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int a_stage_bytes = 2 * p.a_tile_bytes;                 // hi tile, lo tile
  uint8_t* w_ring = smem + p.a_stages * a_stage_bytes;
  uint8_t* store_stage = w_ring + p.w_stages * W_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(store_stage + 4 * p.store_warp_bytes);
  uint64_t* a_full = bars;            // [2]
  uint64_t* a_empty = bars + 2;       // [2]
  uint64_t* w_full = bars + 4;        // [4]
  uint64_t* w_empty = bars + 8;       // [4]
  uint64_t* acc_full = bars + 12;     // [2]
  uint64_t* acc_empty = bars + 14;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);   // one arrival per epilogue warp
    }
    for (int s = 0; s < 4; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int my_items = (p.n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int my_chunks = my_items * p.kchunks;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&ta_hi); prefetch_tmap(&ta_lo); prefetch_tmap(&tb_hi); prefetch_tmap(&tb_lo);
      prefetch_tmap(&ta1_hi); prefetch_tmap(&ta1_lo);
      // chunk j of this CTA = (its (j / kchunks)-th work item, K chunk j % kchunks)
      auto load_a = [&](int j) {
        const int item = (int)blockIdx.x + (j / p.kchunks) * (int)gridDim.x;
        const int ch = (j % p.kchunks) * KC;
        const int band = item / p.tiles_n;
        const int b = band / p.bands_per_img;
        const int y0 = (band - b * p.bands_per_img) * p.BH;
        const int s = j % p.a_stages;
        const uint32_t use = (uint32_t)(j / p.a_stages);
        mbar_wait(&a_empty[s], (use & 1u) ^ 1u);
        uint8_t* st = smem + s * a_stage_bytes;
        mbar_expect_tx(&a_full[s], p.a_tx_bytes);
        // box (64 channels, W+2 columns from -1, 1, BH+2 rows from y0-1, 1 image); out-of-image elements arrive as zeros
        if (ch < p.K0) {
          tma_load_5d(st, &ta_hi, &a_full[s], ch, -1, 0, y0 - 1, b);
          tma_load_5d(st + p.a_tile_bytes, &ta_lo, &a_full[s], ch, -1, 0, y0 - 1, b);
        } else {   // second tensor of a channel concat
          tma_load_5d(st, &ta1_hi, &a_full[s], ch - p.K0, -1, 0, y0 - 1, b);
          tma_load_5d(st + p.a_tile_bytes, &ta1_lo, &a_full[s], ch - p.K0, -1, 0, y0 - 1, b);
        }
      };
      int ws = 0;
      uint32_t wphase = 0;
      if (my_chunks > 0) load_a(0);
      for (int j = 0; j < my_chunks; ++j) {
        const int item = (int)blockIdx.x + (j / p.kchunks) * (int)gridDim.x;
        const int ch = (j % p.kchunks) * KC;
        const int tn = item % p.tiles_n;
        for (int tap = 0; tap < 9; ++tap) {
          // the next chunk's activation tile goes out once this chunk's first weight taps are in flight: its
          // stage frees when the PREVIOUS chunk's MMAs retire, so the fill overlaps this whole chunk
          if (tap == 2 && j + 1 < my_chunks) load_a(j + 1);
          mbar_wait(&w_empty[ws], wphase ^ 1u);
          uint8_t* st = w_ring + ws * W_STAGE_BYTES;
          mbar_expect_tx(&w_full[ws], (uint32_t)W_STAGE_BYTES);
          tma_load_2d(st, &tb_hi, &w_full[ws], tap * p.K + ch, tn * BN);
          tma_load_2d(st + BN * KC * 2, &tb_lo, &w_full[ws], tap * p.K + ch, tn * BN);
          if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // D = f32, A = B = bf16, K-major; N >> 3 at bits 17-22, M >> 4 at bits 24-28
      const uint32_t idesc_2n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(2 * BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc_n = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t s0 = smem_u32(smem);
      const uint64_t dA_hi0 = make_sw128_desc(s0), dA_lo0 = make_sw128_desc(s0 + (uint32_t)p.a_tile_bytes);
      const uint64_t dB0 = make_sw128_desc(smem_u32(w_ring));
      const uint32_t a_stage16 = (uint32_t)a_stage_bytes >> 4;
      int ws = 0;
      uint32_t wphase = 0;
      int as = 0;
      uint32_t aphase = 0;
      int j = 0;
      for (int it = 0; it < my_items; ++it) {
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        const int band = item / p.tiles_n;
        const int y0 = (band % p.bands_per_img) * p.BH;
        const int bh = min(p.BH, p.H - y0);
        const int n_t = (bh * p.PW > 128) ? 2 : 1;     // a short last band may live in the first tile alone
        mbar_wait(&acc_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * 2 * TILE_COLS);
        for (int kc = 0; kc < p.kchunks; ++kc, ++j) {
          const int s = j % p.a_stages;
          mbar_wait(&a_full[s], (uint32_t)(j / p.a_stages) & 1u);
          tc_fence_after();
          const uint32_t a_off = (uint32_t)s * a_stage16;
          uint32_t shift = 0;   // (ky * PW + kx) rows of 128 B, in 16-byte units
          for (int ky = 0; ky < 3; ++ky, shift += (uint32_t)(p.PW - 3) * 8u) {
            for (int kx = 0; kx < 3; ++kx, shift += 8u) {
              mbar_wait(&w_full[ws], wphase);
              tc_fence_after();
              const uint64_t db = dB0 + (uint32_t)ws * (uint32_t)(W_STAGE_BYTES >> 4);
              // start row of the shifted tile modulo the 8-row swizzle atom (stage and tile offsets are multiples of 8 rows)
              const uint64_t bo = p.base_offset ? ((uint64_t)((shift >> 3) & 7u) << 49) : 0ull;
              const uint32_t first = (kc | ky | kx) ? 1u : 0u;
              for (int t = 0; t < n_t; ++t) {
                const uint32_t off = a_off + shift + (uint32_t)t * (128u * 128u >> 4);
                const uint32_t d = d_tmem + (uint32_t)(t * TILE_COLS);
#pragma unroll
                for (int k = 0; k < KC / UMMA_K; ++k) {
                  const uint32_t ko = (uint32_t)(k * UMMA_K * 2 >> 4);   // 32 bytes inside the 128-byte swizzle row
                  umma_bf16(d, (dA_hi0 + off + ko) | bo, db + ko, idesc_2n, (k == 0) ? first : 1u);   // [a_hi*w_hi | a_hi*w_lo]
                  umma_bf16(d, (dA_lo0 + off + ko) | bo, db + ko, idesc_n, 1u);                       //  a_lo*w_hi into the first half
                }
              }
              umma_commit(&w_empty[ws]);
              if (++ws == p.w_stages) { ws = 0; wphase ^= 1u; }
            }
          }
          umma_commit(&a_empty[s]);
        }
        umma_commit(&acc_full[as]);
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    }
  } else {
    // ---- epilogue warps 2..5: TMEM lane quarter = warp % 4 ----
    const int q = warp & 3;
    uint8_t* st_f = store_stage + q * p.store_warp_bytes;   // 32 rows x 64 B fp32, SWIZZLE_64B
    uint8_t* st_h = st_f + STORE_F32_BYTES;                 // 32 rows x 32 B bf16 (only when want_split)
    uint8_t* st_l = st_h + STORE_BF16_BYTES;
    int as = 0;
    uint32_t aphase = 0;
    for (int it = 0; it < my_items; ++it) {
      const int item = (int)blockIdx.x + it * (int)gridDim.x;
      const int band = item / p.tiles_n, tn = item - band * p.tiles_n;
      const int b = band / p.bands_per_img;
      const int band_i = band - b * p.bands_per_img;
      const int y0 = band_i * p.BH;
      const int bh = min(p.BH, p.H - y0);
      const int n_t = (bh * p.PW > 128) ? 2 : 1;
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      for (int t = 0; t < 2; ++t) {
        const int m0w = t * 128 + q * 32;        // first accumulator row of this warp in the band's padded-linear space
        const int m = m0w + lane;
        const int yy = m / p.PW, xx = m - yy * p.PW;
        const bool valid = (t < n_t) && (xx < p.W) && (yy < bh);
        const int64_t opix = ((int64_t)b * p.H + (y0 + yy)) * p.W + xx;
        // GroupNorm partial slot of this warp: one per (band, tile, warp); every slot is written on every launch
        float* gn_dst = p.gn_part
            ? p.gn_part + ((int64_t)b * p.gn_slots + (band_i * 2 + t) * 4 + q) * (kGroups * 2) : nullptr;
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 2 * TILE_COLS + t * TILE_COLS);
        // image rows this warp's 32 accumulator rows touch (inside the band)
        const int yr0 = m0w / p.PW, yr1 = min((m0w + 31) / p.PW, bh - 1);
        float gs = 0.f, gss = 0.f;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += CW) {
          float v[CW];
          if (t < n_t) {
            tmem_ld_32x16(t_base + (uint32_t)c0, v);
            float w[CW];   // products with w_lo
            tmem_ld_32x16(t_base + (uint32_t)(BN + c0), w);
#pragma unroll
            for (int jj = 0; jj < CW; ++jj) v[jj] += w[jj];
          } else {
#pragma unroll
            for (int jj = 0; jj < CW; ++jj) v[jj] = 0.f;
          }
          const int n = tn * BN + c0;
          if (p.bias) {
#pragma unroll
            for (int jj = 0; jj < CW; jj += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n + jj));
              v[jj] += bv.x; v[jj + 1] += bv.y; v[jj + 2] += bv.z; v[jj + 3] += bv.w;
            }
          }
          if (gn_dst) {
            float* dst = gn_dst + (n / p.gn_cpg) * 2;
            if (p.gn_cpg == 8) chunk_stats<8>(v, valid, lane, dst);
            else if (p.gn_cpg == 16) chunk_stats<16>(v, valid, lane, dst);
            else {   // groups of 32 or 64 channels: accumulate chunk sums until the group is complete
              float s = 0.f, ss = 0.f;
#pragma unroll
              for (int jj = 0; jj < CW; ++jj) { s += v[jj]; ss = fmaf(v[jj], v[jj], ss); }
              if (!valid) { s = 0.f; ss = 0.f; }
              gs += warp_sum(s);
              gss += warp_sum(ss);
              if (((n + CW) % p.gn_cpg) == 0) {
                if (lane == 0) { dst[0] = gs; dst[1] = gss; }
                gs = 0.f; gss = 0.f;
              }
            }
          }
          if (t >= n_t || yr1 < yr0) continue;   // nothing of this warp's rows lies inside the band
          if (p.add0 && valid) {
            const float* ad = p.add0 + opix * p.N + n;
#pragma unroll
            for (int jj = 0; jj < CW; jj += 4) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(ad + jj));
              v[jj] += av.x; v[jj + 1] += av.y; v[jj + 2] += av.z; v[jj + 3] += av.w;
            }
          }
          if (!p.tma_rows) {
            // fallback that does not rely on TMA stores clipping a box with a negative start: each valid thread writes
            // its own row segment (64 contiguous bytes of its output pixel)
            if (valid) {
              float* o = p.out0 + opix * p.N + n;
#pragma unroll
              for (int jj = 0; jj < CW; jj += 4)
                *reinterpret_cast<float4*>(o + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
              if (p.want_split) {
#pragma unroll
                for (int jj = 0; jj < CW; jj += 8) {
                  __align__(16) uint32_t h[4], l[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) split_pair(v[jj + 2 * e], v[jj + 2 * e + 1], h[e], l[e]);
                  *reinterpret_cast<uint4*>(p.hi0 + opix * p.N + n + jj) = *reinterpret_cast<const uint4*>(h);
                  *reinterpret_cast<uint4*>(p.lo0 + opix * p.N + n + jj) = *reinterpret_cast<const uint4*>(l);
                }
              }
            }
            continue;
          }
          // registers -> staging tile (row = lane) -> one TMA store per image row the warp touches
          if (lane == 0) tma_store_wait_read<0>();   // the previous chunk's stores have drained the staging tiles
          __syncwarp();
#pragma unroll
          for (int jj = 0; jj < CW / 4; ++jj)
            *reinterpret_cast<float4*>(st_f + lane * 64 + ((jj ^ ((lane >> 1) & 3)) << 4)) =
                make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
          if (p.want_split) {
#pragma unroll
            for (int jj = 0; jj < CW / 8; ++jj) {
              __align__(16) uint32_t h[4], l[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) split_pair(v[8 * jj + 2 * e], v[8 * jj + 2 * e + 1], h[e], l[e]);
              *reinterpret_cast<uint4*>(st_h + lane * 32 + jj * 16) = *reinterpret_cast<const uint4*>(h);
              *reinterpret_cast<uint4*>(st_l + lane * 32 + jj * 16) = *reinterpret_cast<const uint4*>(l);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // staging row r holds accumulator row m0w + r = image row yr, column m0w + r - yr * PW: a box that starts at
            // x = m0w - yr * PW (negative from the second row on) puts every row of image row yr in place, and the
            // rows that belong to other image rows or to the junk columns fall outside [0, W) and are clipped
            for (int yr = yr0; yr <= yr1; ++yr) {
              const int xs = m0w - yr * p.PW;
              if (xs >= p.W) continue;            // only junk columns of this image row
              tma_store_4d(&to0, st_f, n, xs, y0 + yr, b);
              if (p.want_split) {
                tma_store_4d(&to_hi, st_h, n, xs, y0 + yr, b);
                tma_store_4d(&to_lo, st_l, n, xs, y0 + yr, b);
              }
            }
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();   // outstanding bulk stores complete before the CTA retires
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn_f() {
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

// rank-5 view (channel, x, 1, y, image) of a [Bmax, H, W, C] bf16 tensor, box (64, W+2, 1, rows, 1)
int encode_band(Status& st, CUtensorMap* m, void* ptr, int C, int H, int W, int Bmax, int rows) {
  auto enc = encode_fn_f();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t rowB = (cuuint64_t)W * C * 2;
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, 1, (cuuint64_t)H, (cuuint64_t)Bmax};
  cuuint64_t strides[4] = {(cuuint64_t)C * 2, rowB, rowB, rowB * H};
  cuuint32_t box[5] = {(cuuint32_t)KC, (cuuint32_t)(W + 2), 1u, (cuuint32_t)rows, 1u};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (halo conv activations) failed");
  return IGM_OK;
}

// [Bmax, H, W, C] output as (C, W, H, B) with a box of CW channels x 32 columns of ONE image row
int encode_row_out(Status& st, CUtensorMap* m, const void* ptr, int C, int H, int W, int Bmax, bool bf16) {
  auto enc = encode_fn_f();
  if (!enc) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t es_b = bf16 ? 2 : 4;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Bmax};
  cuuint64_t strides[3] = {(cuuint64_t)C * es_b, (cuuint64_t)W * C * es_b, (cuuint64_t)H * W * C * es_b};
  cuuint32_t box[4] = {(cuuint32_t)CW, 32u, 1u, 1u};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(ptr), dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (halo conv output) failed");
  return IGM_OK;
}

int round_up_i(int v, int a) { return (v + a - 1) / a * a; }

// shared-memory plan: two activation stages, the store staging, then as many weight stages (<= 4) as fit
bool smem_plan(int W, bool split, int& BH, int& a_tile_bytes, int& w_stages, int& store_warp_bytes, int& total) {
  const int PW = W + 2;
  BH = 256 / PW;
  if (BH < 1) return false;
  a_tile_bytes = round_up_i((BH + 2) * PW * 128, 1024);
  store_warp_bytes = round_up_i(STORE_F32_BYTES + (split ? 2 * STORE_BF16_BYTES : 0), 1024);
  const int fixed = 2 * 2 * a_tile_bytes + 4 * store_warp_bytes + 1024 /*align slack*/ + 256 /*barriers*/;
  w_stages = (kSmemLimit - fixed) / W_STAGE_BYTES;
  if (w_stages > 4) w_stages = 4;
  if (w_stages < 2) return false;
  // shifted descriptors of the second tile read up to row 255 + 2 PW + 2 of the LAST lo tile: that must stay inside the
  // CTA's window (the weight ring and the store staging follow the activation ring)
  const int overrun = (256 + 2 * PW + 2) * 128 - a_tile_bytes;
  if (overrun > w_stages * W_STAGE_BYTES + 4 * store_warp_bytes) return false;
  total = fixed + w_stages * W_STAGE_BYTES;
  return true;
}

}  // namespace

bool tch_eligible(int K, int N, int H, int W) {
  if (K < KC || K % KC != 0 || N < BN || N % BN != 0) return false;
  if (W < 28 || W + 2 > 128 || H < 1) return false;   // narrower images leave too many junk accumulator rows
  int BH, at, ws, sw, tot;
  return smem_plan(W, true, BH, at, ws, sw, tot);
}

int tch_plan(Status& st, TcConvHalo& t, int K, int N, int H, int W, int Bmax, __nv_bfloat16* a_hi, __nv_bfloat16* a_lo,
             __nv_bfloat16* w_hi, __nv_bfloat16* w_lo, int K0, __nv_bfloat16* a1_hi, __nv_bfloat16* a1_lo) {
  t.valid = false;
  if (!tch_eligible(K, N, H, W)) IGM_FAIL(st, IGM_ERR_INVALID, "shape not eligible for the halo-reuse tcgen05 conv");
  if (K0 <= 0 || !a1_hi) K0 = K;
  if (K0 % KC != 0 || (K - K0) % KC != 0) IGM_FAIL(st, IGM_ERR_INVALID, "concat split must be a multiple of 64 channels");
  t.K = K; t.K0 = K0; t.N = N; t.H = H; t.W = W; t.Bmax = Bmax;
  int ws, sw, tot;
  smem_plan(W, true, t.BH, t.a_tile_bytes, ws, sw, tot);
  const bool two = K0 < K;
  IGM_TRY(encode_band(st, &t.a_hi, a_hi, K0, H, W, Bmax, t.BH + 2));
  IGM_TRY(encode_band(st, &t.a_lo, a_lo, K0, H, W, Bmax, t.BH + 2));
  IGM_TRY(encode_band(st, &t.a1_hi, two ? (void*)a1_hi : (void*)a_hi, two ? K - K0 : K0, H, W, Bmax, t.BH + 2));
  IGM_TRY(encode_band(st, &t.a1_lo, two ? (void*)a1_lo : (void*)a_lo, two ? K - K0 : K0, H, W, Bmax, t.BH + 2));
  auto enc = encode_fn_f();
  for (int which = 0; which < 2; ++which) {   // weights: [N rows][9*K cols], K-major, the layout conv_tc.cu uses
    cuuint64_t dims[2] = {(cuuint64_t)9 * K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)9 * K * 2};
    cuuint32_t box[2] = {(cuuint32_t)KC, (cuuint32_t)BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(which ? &t.b_lo : &t.b_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, which ? (void*)w_lo : (void*)w_hi, dims,
                     strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) IGM_FAIL(st, IGM_ERR_CUDA, "cuTensorMapEncodeTiled (halo conv weights) failed");
  }
  t.valid = true;
  return IGM_OK;
}

int tch_gn_slots(const TcConvHalo& t) { return cdiv(t.H, t.BH) * 8; }

bool tch_gn_fusable(const TcConvHalo& t) {
  const int cpg = t.N / kGroups;
  return t.valid && t.N % kGroups == 0 && (cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64);
}

int launch_conv_halo(const LaunchCtx& lc, const TcConvHalo& t, const TcRun& r) {
  if (!t.valid) IGM_FAIL(*lc.st, IGM_ERR_STATE, "halo conv plan not initialised");
  if (r.B < 1 || r.B > t.Bmax) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_halo: bad batch");
  if (r.N0 != t.N || r.out1 || r.add1) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_halo: split outputs are not supported");
  if (r.gn_part && !tch_gn_fusable(t)) IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_halo: GroupNorm statistics cannot be fused for this plan");
  HFArgs a;
  a.B = r.B; a.H = t.H; a.W = t.W; a.PW = t.W + 2; a.BH = t.BH;
  a.K = t.K; a.K0 = t.K0; a.N = t.N;
  a.bands_per_img = cdiv(t.H, t.BH);
  a.tiles_n = t.N / BN;
  a.n_items = r.B * a.bands_per_img * a.tiles_n;
  a.kchunks = t.K / KC;
  int bh_, at_, smem_bytes = 0;
  if (!smem_plan(t.W, r.hi0 != nullptr, bh_, at_, a.w_stages, a.store_warp_bytes, smem_bytes))
    IGM_FAIL(*lc.st, IGM_ERR_INVALID, "conv_halo: shared-memory plan does not fit");
  a.a_tile_bytes = t.a_tile_bytes; a.a_stages = 2;
  a.a_tx_bytes = (uint32_t)(2 * (t.BH + 2) * (t.W + 2) * 128);
  a.bias = r.bias; a.add0 = r.add0;
  a.want_split = r.hi0 ? 1 : 0;
  static const bool tma_rows_off = [] { const char* e = getenv("IGM_HALO_TMA_STORE"); return e && e[0] == '0'; }();
  a.tma_rows = tma_rows_off ? 0 : 1;
  static const bool base_off = [] { const char* e = getenv("IGM_HALO_BASE_OFFSET"); return e && e[0] == '1'; }();
  a.base_offset = base_off ? 1 : 0;
  a.out0 = r.out0; a.hi0 = r.hi0; a.lo0 = r.lo0;
  a.gn_part = r.gn_part; a.gn_cpg = r.gn_part ? t.N / kGroups : 0; a.gn_slots = tch_gn_slots(t);
  TcConv::OutMaps& om = t.om;
  if (om.p0 != r.out0) {
    IGM_TRY(encode_row_out(*lc.st, &om.m0, r.out0, t.N, t.H, t.W, t.Bmax, false));
    om.p0 = r.out0;
    if (!om.ph) { om.mh = om.m0; om.ml = om.m0; }
  }
  if (r.hi0 && (om.ph != r.hi0 || om.pl != r.lo0)) {
    IGM_TRY(encode_row_out(*lc.st, &om.mh, r.hi0, t.N, t.H, t.W, t.Bmax, true));
    IGM_TRY(encode_row_out(*lc.st, &om.ml, r.lo0, t.N, t.H, t.W, t.Bmax, true));
    om.ph = r.hi0; om.pl = r.lo0;
  }
  static int attr_smem = 0;
  if (smem_bytes > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = smem_bytes;
  }
  const double flops = 2.0 * r.B * t.H * t.W * (double)t.N * t.K * 9;
  const double bytes = 4.0 * ((double)r.B * t.H * t.W * (t.K + t.N * (r.add0 ? 2 : 1)) + 9.0 * t.K * t.N);
  ProfScope ps_(lc, r.kclass, flops, bytes);
  const int grid = a.n_items < 148 ? a.n_items : 148;
  cudaError_t le = launch_pdl(conv_halo_kernel, dim3(grid), dim3(192), (size_t)smem_bytes, lc.stream, t.a_hi, t.a_lo, t.a1_hi,
                              t.a1_lo, t.b_hi, t.b_lo, om.m0, om.mh, om.ml, a);
  if (le != cudaSuccess) IGM_FAIL(*lc.st, IGM_ERR_CUDA, cudaGetErrorString(le));
  IGM_POST_LAUNCH(lc);
  return IGM_OK;
}

}  // namespace igm
