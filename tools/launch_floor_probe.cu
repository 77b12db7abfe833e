// Hardware probe (not product code): per-launch floor of a 148-CTA kernel as the tcgen05 prologue pieces are
// added one by one (large dynamic smem, mbarrier init, TMEM alloc/dealloc, smem zeroing).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/launch_floor_probe tools/launch_floor_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../image-generation-models_b200/csrc/tc_ptx.cuh"
using namespace igm::tc;

__global__ void __launch_bounds__(192, 1) k(int mode, int zero_bytes, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 128);
  const int warp = threadIdx.x >> 5;
  if (zero_bytes) {
    uint4* z = reinterpret_cast<uint4*>(smem + 1024);
    for (int i = threadIdx.x; i < zero_bytes / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (mode & 1) {
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) mbar_init(&bars[s], 1); fence_barrier_init(); }
  }
  if (mode & 2) {
    if (warp == 1) tmem_alloc<512>(slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t t = *slot;
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(t); }
  }
  if (out && threadIdx.x == 0 && blockIdx.x == 0) out[0] = 1.f;
}

int main() {
  float* d; cudaMalloc(&d, 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  struct { int smem, mode, zero; const char* name; } cfgs[] = {
      {2048, 0, 0, "2 KB smem, nothing"}, {200 * 1024, 0, 0, "200 KB smem, nothing"}, {200 * 1024, 1, 0, "+ mbarrier init"},
      {200 * 1024, 2, 0, "+ tmem alloc/dealloc 512"}, {200 * 1024, 3, 0, "+ both"}, {200 * 1024, 3, 176 * 1024, "+ both + zero 176 KB"},
      {2048, 2, 0, "2 KB smem + tmem alloc/dealloc"}};
  for (auto& c : cfgs) {
    for (int i = 0; i < 20; ++i) k<<<148, 192, c.smem>>>(c.mode, c.zero, d);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 200; ++i) k<<<148, 192, c.smem>>>(c.mode, c.zero, d);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("%-36s %.2f us per launch (%s)\n", c.name, ms * 1000 / 200, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
