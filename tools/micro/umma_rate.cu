// Micro-benchmark: sustained issue rate of tcgen05.mma (kind::f16, bf16 operands from shared memory, M = 128) for the
// instruction shapes of the attention kernels: N in {64, 128, 256}, K-major / MN-major operands, one thread issuing
// `batch` instructions (K = 16 each) per commit.  Prints cycles per instruction against the math floor M*N/256... (N/2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pytorch_generative_b200/csrc/pg_common.cuh"
void pg_set_error(const char*, ...) {}
int pg_check_launch(const char*) { return 0; }

constexpr int ATOM = 128 * 128;  // [128 rows][64 bf16] swizzle atom

__device__ __forceinline__ uint64_t d_k(uint32_t addr, int kk) { return umma_desc_sw128(addr + (kk >> 2) * ATOM + (kk & 3) * 32, 16, 1024); }
__device__ __forceinline__ uint64_t d_mn(uint32_t addr, int kk) { return umma_desc_sw128(addr + kk * 2048, ATOM, 1024); }

template <int N, int A_MN, int B_MN, int ELECTED>
__global__ void __launch_bounds__(128, 1) umma_rate(int iters, int batch, int ncommit_wait, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (2 * ATOM + 4 * ATOM) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); fence_proxy_async_smem(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (ELECTED ? (warp == 0) : (threadIdx.x == 0)) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N, A_MN, B_MN);
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 2 * ATOM);
    uint32_t ph = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int kk = 0; kk < batch; ++kk) {
        const int k8 = kk & 7;
        const uint64_t da = A_MN ? d_mn(a_addr, k8) : d_k(a_addr, k8);
        const uint64_t db = B_MN ? d_mn(b_addr, k8) : d_k(b_addr, k8 & 3);
        if (ELECTED) umma_bf16_ss_w(tmem + ((it & 1) * 256), da, db, idesc, kk > 0);
        else umma_bf16_ss(tmem + ((it & 1) * 256), da, db, idesc, kk > 0);
      }
      if (ncommit_wait || it == iters - 1) {
        if (ELECTED) umma_commit_w(&bar);
        else umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(slot); }
}

template <int N, int A_MN, int B_MN, int ELECTED = 0>
void run(const char* name, long long* cyc) {
  const int smem = 6 * ATOM + 1024;
  cudaFuncSetAttribute(umma_rate<N, A_MN, B_MN, ELECTED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int batch : {4, 8, 32}) {
    for (int wait : {0, 1}) {
      const int iters = 2000;
      umma_rate<N, A_MN, B_MN, ELECTED><<<148, 128, smem>>>(iters, batch, wait, cyc);
      cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("%-28s N=%3d batch=%2d %s: %7.1f cycles / instruction (math floor %d)  [%s]\n", name, N, batch,
             wait ? "wait each commit " : "back-to-back     ", (double)h[0] / ((double)iters * batch), N / 2,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  run<64, 0, 0>("A K-major, B K-major", cyc);
  run<64, 0, 1>("A K-major, B MN-major", cyc);
  run<64, 1, 1>("A MN-major, B MN-major", cyc);
  run<128, 0, 0>("A K-major, B K-major", cyc);
  run<128, 1, 1>("A MN-major, B MN-major", cyc);
  run<256, 0, 0>("A K-major, B K-major", cyc);
  printf("---- warp-converged issue (elect.sync), same shapes ----\n");
  run<64, 0, 0, 1>("elected: A K, B K", cyc);
  run<64, 1, 1, 1>("elected: A MN, B MN", cyc);
  run<128, 0, 0, 1>("elected: A K, B K", cyc);
  run<256, 0, 0, 1>("elected: A K, B K", cyc);
  return 0;
}
