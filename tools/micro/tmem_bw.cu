// Micro-benchmark: TMEM -> register read bandwidth of tcgen05.ld.32x32b.x32 per SM, for 4 / 8 / 16 reader warps.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pytorch_generative_b200/csrc/pg_common.cuh"
void pg_set_error(const char*, ...) {}
int pg_check_launch(const char*) { return 0; }

template <int X>
__global__ void __launch_bounds__(512, 1) tmem_read(int iters, unsigned* sink, long long* cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (X == 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(base + ((warp >> 2) * 128 + c * 32) % 512, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= v[i];
      } else {
        uint32_t v[16];
        tmem_ld_32x32b_x16(base + ((warp >> 2) * 128 + c * 32) % 512, v);
        tmem_ld_32x32b_x16(base + ((warp >> 2) * 128 + c * 32 + 16) % 512, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc ^= v[i];
      }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

int main() {
  unsigned* sink; long long* cyc;
  cudaMalloc(&sink, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    tmem_read<32><<<148, warps * 32>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 4 * warps * 32 * 32 * 4;
    printf("x32: %2d warps: %lld cycles, %.1f B/cycle/SM  (%s)\n", warps, h[0], bytes / h[0], cudaGetErrorString(cudaGetLastError()));
  }
  for (int warps : {4, 8}) {
    tmem_read<16><<<148, warps * 32>>>(iters, sink, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double bytes = (double)iters * 4 * warps * 32 * 32 * 4;
    printf("x16x2: %2d warps: %lld cycles, %.1f B/cycle/SM\n", warps, h[0], bytes / h[0]);
  }
  return 0;
}
