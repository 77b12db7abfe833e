// Hardware probe (not product code): can a SWIZZLE_128B shared-memory matrix descriptor start at a
// 128-byte row that is NOT 1024-byte aligned?  Answers whether one halo tile in shared memory can feed
// all (ky,kx) taps of a convolution by shifting the descriptor start address.
//   test K: K-major A [rows=M][64 k], start shifted by s rows:   D[m][n]       = sum_k A[m+s][k] * B[n][k]
//   test M: MN-major A [rows=K][64 m], two 64-blocks LBO apart:  D[m+64j][n]   = sum_k A[k+s_j][m] * B[k][n]
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_out/desc_probe tools/desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../image-generation-models_b200/csrc/tc_ptx.cuh"

using namespace igm::tc;

__host__ __device__ inline float aval(int r, int c) { return (float)(((r * 7 + c * 3) % 11) - 5); }
__host__ __device__ inline float bval(int r, int c) { return (float)(((r * 5 + c * 2) % 7) - 3); }

// logical (row r, element c of 64) -> byte offset in a 128B-swizzled buffer whose base is 1024-aligned
__device__ inline uint32_t sw_off(int r, int c) { return (uint32_t)(r * 128 + ((((c >> 3) ^ (r & 7)) << 4) | ((c & 7) << 1))); }

struct Params { int mode, s0, s1, bo; };   // mode 0 = K-major shift, 1 = MN-major two-block shift; bo = use base_offset field

__global__ void __launch_bounds__(128) probe(Params p, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* A = smem;                 // 320 rows x 128 B
  uint8_t* Bm = smem + 320 * 128;    // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(Bm + 64 * 128);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 320 * 64; i += 128) {
    const int r = i >> 6, c = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(A + sw_off(r, c)) = __float2bfloat16(aval(r, c));
  }
  for (int i = tid; i < 64 * 64; i += 128) {
    const int r = i >> 6, c = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(Bm + sw_off(r, c)) = __float2bfloat16(bval(r, c));
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (tid == 0) {
    if (p.mode == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      for (int k = 0; k < 4; ++k) {
        const uint32_t a = smem_u32(A) + p.s0 * 128 + k * 32;
        uint64_t da = make_sw128_desc(a);
        if (p.bo) da |= (uint64_t)((a >> 7) & 7) << 49;
        const uint64_t db = make_sw128_desc(smem_u32(Bm) + k * 32);
        umma_bf16(tmem, da, db, idesc, k ? 1u : 0u);
      }
    } else {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo = (uint32_t)(p.s1 - p.s0) * 128;
      for (int k = 0; k < 4; ++k) {
        const uint32_t a = smem_u32(A) + (p.s0 + k * 16) * 128;
        uint64_t da = make_sw128_mn_desc(a, lbo, 1024);
        if (p.bo) da |= (uint64_t)((a >> 7) & 7) << 49;
        const uint64_t db = make_sw128_mn_desc(smem_u32(Bm) + k * 16 * 128, 8192, 1024);
        umma_bf16(tmem, da, db, idesc, k ? 1u : 0u);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  float v[32];
  for (int c0 = 0; c0 < 64; c0 += 32) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j) out[(warp * 32 + (tid & 31)) * 64 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  float* h = (float*)malloc(128 * 64 * 4);
  const int smem = 320 * 128 + 64 * 128 + 64 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[][2] = {{0, 0}, {1, 0}, {3, 0}, {8, 0}, {13, 0}, {35, 0}, {70, 0}};
  for (int bo = 0; bo < 2; ++bo)
    for (auto& s : shifts) {
      Params p{0, s[0], 0, bo};
      cudaMemset(d, 0xff, 128 * 64 * 4);
      probe<<<1, 128, smem>>>(p, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += aval(m + s[0], k) * bval(n, k);
          if (h[m * 64 + n] != ref) ++bad;
        }
      printf("K-major   shift %3d base_offset_field=%d : %s (%d mismatches) %s\n", s[0], bo, bad ? "FAIL" : "ok", bad,
             e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  const int pairs[][2] = {{0, 8}, {0, 1}, {1, 3}, {3, 35}, {5, 70}, {33, 34}, {0, 64}};
  for (int bo = 0; bo < 2; ++bo)
    for (auto& s : pairs) {
      Params p{1, s[0], s[1], bo};
      cudaMemset(d, 0xff, 128 * 64 * 4);
      probe<<<1, 128, smem>>>(p, d);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, d, 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
          const int sj = s[m >> 6];
          float ref = 0;
          for (int k = 0; k < 64; ++k) ref += aval(k + sj, m & 63) * bval(k, n);
          if (h[m * 64 + n] != ref) ++bad;
        }
      printf("MN-major  shifts (%3d,%3d) base_offset_field=%d : %s (%d mismatches) %s\n", s[0], s[1], bo,
             bad ? "FAIL" : "ok", bad, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
