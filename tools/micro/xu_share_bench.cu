// Micro-benchmark (B200): do MUFU.EX2 and the fp32 -> bf16x2 pack (cvt.rn.bf16x2.f32 = F2FP) share an execution pipe?
// Even warps run op A, odd warps run op B (so every SM sub-partition hosts both); the per-SM rates of the pair are
// compared with each op alone. If A + B together take the SUM of their solo times, they share a pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu_share_bench xu_share_bench.cu && ./xu_share_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { NONE = 0, EX2, F2FP, PRMT, FADD2, IADD, FFMA, F2F16 };

template <int OP>
__device__ __forceinline__ void body(uint32_t (&r)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
    if (OP == F2FP) asm volatile("cvt.rn.bf16x2.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
    if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r[i]));
    if (OP == F2F16) asm volatile("cvt.rn.f16x2.f32 %0, %0, %1;" : "+r"(r[i]) : "r"(r[(i + 1) & 7]));
    if (OP == FADD2) { if ((i & 1) == 0) asm volatile("{.reg .b64 t; mov.b64 t, {%0,%1}; add.rn.f32x2 t, t, t; mov.b64 {%0,%1}, t;}" : "+r"(r[i]), "+r"(r[i + 1])); }
  }
}

template <int A, int B>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
  const bool odd = (threadIdx.x >> 5) & 4;   // warps 0-3 op A, 4-7 op B: each sub-partition gets one of each
  __syncthreads();
  long long t0 = clock64();
  if (!odd) { for (int it = 0; it < iters; ++it) body<A>(r); }
  else      { for (int it = 0; it < iters; ++it) body<B>(r); }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= r[i];
  out[1024 + blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) out[threadIdx.x >> 5] = static_cast<uint32_t>(t1 - t0);
}

template <int A, int B>
void run(const char* name) {
  uint32_t* d;
  cudaMalloc(&d, (1024 + 148 * 256) * 4);
  const int iters = 4096;
  for (int rep = 0; rep < 2; ++rep) { k<A, B><<<148, 256>>>(d, iters, 12345u); cudaDeviceSynchronize(); }
  uint32_t clk[8];
  cudaMemcpy(clk, d, 32, cudaMemcpyDeviceToHost);
  // clk per warp-level instruction, for a warp of the A group and of the B group (one of each per sub-partition)
  printf("%-22s A-warp %6.2f clk/instr   B-warp %6.2f clk/instr   (%s)\n", name, double(clk[0]) / (iters * 8.0), double(clk[4]) / (iters * 8.0),
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<EX2, NONE>("ex2 | idle");
  run<F2FP, NONE>("f2fp.bf16x2 | idle");
  run<F2F16, NONE>("f2fp.f16x2 | idle");
  run<PRMT, NONE>("prmt | idle");
  run<IADD, NONE>("iadd | idle");
  run<FFMA, NONE>("ffma | idle");
  run<FADD2, NONE>("fadd2(4 per 8) | idle");
  run<EX2, EX2>("ex2 | ex2");
  run<EX2, F2FP>("ex2 | f2fp.bf16x2");
  run<EX2, F2F16>("ex2 | f2fp.f16x2");
  run<EX2, PRMT>("ex2 | prmt");
  run<EX2, IADD>("ex2 | iadd");
  run<EX2, FFMA>("ex2 | ffma");
  run<F2FP, F2FP>("f2fp | f2fp");
  run<F2FP, PRMT>("f2fp | prmt");
  run<F2FP, FFMA>("f2fp | ffma");
  return 0;
}
