// Hardware probe (not product code): issue rate of tcgen05.mma kind::f16 (SS mode) at M=128 for N=64/128/256,
// K-major and MN-major operands.  One CTA per SM on all SMs so the power/clock state is realistic.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o build/mma_rate_probe tools/mma_rate_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../image-generation-models_b200/csrc/tc_ptx.cuh"
using namespace igm::tc;

// halo-wgrad operand shapes: MN-major, N = 192 as three 64-blocks `lbo` bytes apart, start shifted by `shift` bytes
__global__ void __launch_bounds__(128) rate_halo(int lbo, int shift, int a_lbo, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(192 >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 40 * 1024 + 2048;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t da = make_sw128_mn_desc(a0 + k * 2048, a_lbo, 1024);
        const uint64_t db = make_sw128_mn_desc(b0 + shift + k * 2048, lbo, 1024);
        umma_bf16(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

__global__ void __launch_bounds__(128) rate(int N, int mn, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(slot);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (mn) idesc |= (1u << 15) | (1u << 16);
    const uint32_t a0 = smem_u32(smem), b0 = a0 + 32 * 1024;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        uint64_t da, db;
        if (mn) { da = make_sw128_mn_desc(a0 + k * 2048, 8192, 1024); db = make_sw128_mn_desc(b0 + k * 2048, 8192, 1024); }
        else { da = make_sw128_desc(a0 + k * 32); db = make_sw128_desc(b0 + k * 32); }
        umma_bf16(tmem + (uint32_t)((it & 1) * 256), da, db, idesc, 1u);
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 96 * 1024 + 64 + 1024;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 4000;
  for (int mn = 0; mn < 2; ++mn)
    for (int N : {64, 128, 256}) {
      for (int rep = 0; rep < 2; ++rep) { rate<<<148, 128, smem>>>(N, mn, iters, d); cudaDeviceSynchronize(); }
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      double cyc = (double)h / (iters * 4);
      printf("%s M=128 N=%3d K=16: %.1f cycles/MMA -> %.0f flop/clk/SM (%s)\n", mn ? "MN-major" : "K-major ", N, cyc,
             2.0 * 128 * N * 16 / cyc, cudaGetErrorString(cudaGetLastError()));
    }
  cudaFuncSetAttribute(rate_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int cfgs[][3] = {{8192, 0, 18432}, {128, 0, 18432}, {128, -128, 18432}, {128, 4224, 18432}, {8192, -128, 18432}, {128, -128, 10240}};
  for (auto& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) { rate_halo<<<148, 128, smem>>>(c[0], c[1], c[2], iters, d); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    double cyc = (double)h / (iters * 4);
    printf("MN-major M=128 N=192 B.lbo=%5d B.shift=%5d A.lbo=%d: %.1f cycles/MMA (%s)\n", c[0], c[1], c[2], cyc,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
