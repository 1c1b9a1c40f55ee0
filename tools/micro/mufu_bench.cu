// Micro-benchmark (B200): per-SM throughput of the exponential flavours the attention softmax could use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + threadIdx.x * 8 + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(r[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r[i]));
      if (MODE == 3) { if ((i & 1) == 0) asm volatile("{.reg .b64 t; mov.b64 t, {%0,%1}; fma.rn.f32x2 t, t, t, t; mov.b64 {%0,%1}, t;}" : "+r"(r[i]), "+r"(r[i + 1])); }
      if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r[i]));
      if (MODE == 5) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo,hi}, %0; ex2.approx.ftz.bf16 lo, lo; mov.b32 %0, {lo,hi};}" : "+r"(r[i]));
      if (MODE == 6) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(r[i]));
    }
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = static_cast<uint32_t>(t1 - t0);
}

template <int MODE>
void run(const char* name, int warps_per_sm, int results_per_op) {
  uint32_t* d;
  cudaMalloc(&d, 148 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148, warps_per_sm * 32>>>(d, iters, 12345u);
  cudaDeviceSynchronize();
  k<MODE><<<148, warps_per_sm * 32>>>(d, iters, 12345u);
  cudaDeviceSynchronize();
  uint32_t clk;
  cudaMemcpy(&clk, d, 4, cudaMemcpyDeviceToHost);
  const double ops = double(iters) * 8 * warps_per_sm * 32;   // thread-level instructions per SM
  printf("%-28s warps/SM %2d: %7.2f thread-instr/clk/SM  -> %7.2f results/clk/SM  (err %s)\n", name, warps_per_sm, ops / clk,
         ops * results_per_op / clk, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  for (int w : {4, 8, 16}) {
    run<0>("ex2.approx.ftz.f32", w, 1);
    run<1>("ex2.approx.ftz.bf16x2", w, 2);
    run<2>("ex2.approx.f16x2", w, 2);
    run<5>("ex2.approx.ftz.bf16", w, 1);
    run<6>("tanh.approx.f32", w, 1);
    run<3>("fma.rn.f32x2 (4 per 8 slots)", w, 1);
    run<4>("fma.rn.f32", w, 1);
  }
  return 0;
}
