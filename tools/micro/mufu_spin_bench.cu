// Micro-benchmark (B200): MUFU.EX2 throughput of compute warps while other warps of the CTA spin on
// mbarrier.try_wait (what the TMA / MMA-issuer warps of the attention kernel do).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// warps [0, n_spin) spin, warps [n_spin, n_spin + n_comp) run ex2 chains; mode: 0 = try_wait spin, 1 = try_wait + nanosleep(20),
// 2 = test_wait spin (non-blocking), 3 = plain smem volatile flag spin
__global__ void k(uint32_t* out, int n_spin, int iters, int mode) {
  __shared__ uint64_t bar;
  __shared__ volatile int flag;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    flag = 0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  if (warp < n_spin) {
    if (mode == 3) { while (flag == 0) {} return; }
    uint32_t ok = 0;
    while (!ok) {
      if (mode == 2)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
      if (mode == 1 && !ok) __nanosleep(20);
    }
    return;
  }
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = 12345u + threadIdx.x * 8 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (warp == n_spin && (threadIdx.x & 31) == 0) {
    if (blockIdx.x == 0) out[0] = static_cast<uint32_t>(t1 - t0);
  }
  __syncwarp();
  // release the spinners: the last compute warp to finish arrives / sets the flag (approximation: every compute warp does)
  if ((threadIdx.x & 31) == 0 && warp == n_spin) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
    flag = 1;
  }
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 2048;
  for (int mode = 0; mode < 4; ++mode)
    for (int n_spin : {0, 1, 2, 5}) {
      for (int n_comp : {4, 8}) {
        k<<<148, (n_spin + n_comp) * 32>>>(d, n_spin, iters, mode);
        cudaDeviceSynchronize();
        k<<<148, (n_spin + n_comp) * 32>>>(d, n_spin, iters, mode);
        cudaError_t e = cudaDeviceSynchronize();
        uint32_t clk;
        cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
        printf("mode %d (%s) spinners %d, compute warps %d: %6.2f ex2/clk/SM (%s)\n", mode,
               mode == 0 ? "try_wait" : mode == 1 ? "try_wait+nanosleep" : mode == 2 ? "test_wait" : "smem flag", n_spin, n_comp,
               double(iters) * 8 * n_comp * 32 / clk, cudaGetErrorString(e));
      }
    }
  return 0;
}
