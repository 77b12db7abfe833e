// Hardware probe (not product code): shared-memory fill rate per SM through TMA (cp.async.bulk, 1-D) from an
// L2-resident buffer, as a function of how many SMs pull at once.  Decides whether fewer, fatter CTAs can beat the
// ~37 B/clk/SM that 148 CTAs see (chip-wide L2 output cap) on the small-M convolutions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/tma_ingest_probe tools/tma_ingest_probe.cu
#include <cstdio>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../image-generation-models_b200/csrc/tc_ptx.cuh"
using namespace igm::tc;

constexpr int STAGE = 32 * 1024, STAGES = 4;

__global__ void __launch_bounds__(64) pull(const uint8_t* __restrict__ src, size_t span, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    size_t off = ((size_t)blockIdx.x * 7919 * STAGE) % span;
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&bars[s], ((it / STAGES) - 1) & 1);
      if (it < iters) {
        mbar_expect_tx(&bars[s], STAGE);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(smem + s * STAGE)), "l"(src + off), "r"(STAGE), "r"(smem_u32(&bars[s])) : "memory");
        off = (off + (size_t)gridDim.x * STAGE) % span;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

// same pull loop with 2-D tiled tensor-map loads: box = 64 bf16 (128 B) x 128 rows = 16 KB, SWIZZLE_128B, two per stage
__global__ void __launch_bounds__(64) pull_tiled(const __grid_constant__ CUtensorMap tm, int rows_total, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  if (threadIdx.x == 0) { for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    int row = (blockIdx.x * 7919 * 256) % rows_total;
    for (int it = 0; it < iters + STAGES; ++it) {
      const int s = it % STAGES;
      if (it >= STAGES) mbar_wait(&bars[s], ((it / STAGES) - 1) & 1);
      if (it < iters) {
        mbar_expect_tx(&bars[s], STAGE);
        tma_load_2d(smem + s * STAGE, &tm, &bars[s], 0, row);
        tma_load_2d(smem + s * STAGE + 16384, &tm, &bars[s], 0, row + 128);
        row = (row + gridDim.x * 256) % rows_total;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t span = 64ull << 20;   // 64 MB: L2-resident after the first pass
  uint8_t* d; cudaMalloc(&d, span + STAGE); cudaMemset(d, 1, span + STAGE);
  long long* o; cudaMalloc(&o, 148 * 8);
  const int smem = STAGES * STAGE + 1024 + 64;
  cudaFuncSetAttribute(pull, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2000;
  for (int ctas : {8, 32, 64, 96, 128, 148}) {
    for (int rep = 0; rep < 2; ++rep) { pull<<<ctas, 64, smem>>>(d, span, iters, o); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, o, ctas * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%3d CTAs: %.1f B/clk per SM, %.0f B/clk chip-wide (%s)\n", ctas, (double)iters * STAGE / mx,
           (double)iters * STAGE / mx * ctas, cudaGetErrorString(cudaGetLastError()));
  }
  // tiled variant
  void* fp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fp);
  const int rows_total = (int)(span / 128) - 512;
  CUtensorMap tm;
  cuuint64_t dims[2] = {64, (cuuint64_t)(span / 128)};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, 128};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  cudaFuncSetAttribute(pull_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int ctas : {64, 128, 148}) {
    for (int rep = 0; rep < 2; ++rep) { pull_tiled<<<ctas, 64, smem>>>(tm, rows_total, iters, o); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, o, ctas * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("tiled 2-D, %3d CTAs: %.1f B/clk per SM, %.0f B/clk chip-wide (%s)\n", ctas, (double)iters * STAGE / mx,
           (double)iters * STAGE / mx * ctas, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
