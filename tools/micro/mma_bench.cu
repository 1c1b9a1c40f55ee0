// Micro-benchmark (B200): issue/execute rate of tcgen05.mma kind::f16 at the shapes the attention kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../vfmseg_b200/csrc -o mma_bench mma_bench.cu
#include <cstdio>
#include "sm100_ptx.cuh"
using namespace vfm;

// mode bits: 0 = SS / 1 = TS (A from TMEM); n = MMA N; nacc = number of accumulators cycled; b_mn = B MN-major
__global__ void __launch_bounds__(128, 1) k(int ts, int n, int nacc, int b_mn, int iters, int k_advance, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0) {
    const uint64_t da = make_sw128_desc(smem_u32(smem));
    const uint64_t db = make_sw128_desc(smem_u32(smem + 16384));
    const uint32_t idesc = make_idesc_bf16(128, n, 0, b_mn);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      t0 = clock64();
      if (elect_one_sync()) {
        const uint32_t kb = b_mn ? 128 : 2;
        const uint32_t acc1 = tmem + (nacc > 1 ? 256u : 0u);
        const uint32_t ta = tmem + 384u;
        for (int i = 0; i < iters; i += 8) {
          if (ts) {
            umma_ts(tmem, ta, db, idesc, 1); umma_ts(tmem, ta + 8, db + kb, idesc, 1);
            umma_ts(tmem, ta + 16, db + 2 * kb, idesc, 1); umma_ts(tmem, ta + 24, db + 3 * kb, idesc, 1);
            umma_ts(acc1, ta, db, idesc, 1); umma_ts(acc1, ta + 8, db + kb, idesc, 1);
            umma_ts(acc1, ta + 16, db + 2 * kb, idesc, 1); umma_ts(acc1, ta + 24, db + 3 * kb, idesc, 1);
          } else {
            umma_ss(tmem, da, db, idesc, 1); umma_ss(tmem, da + 2, db + kb, idesc, 1);
            umma_ss(tmem, da + 4, db + 2 * kb, idesc, 1); umma_ss(tmem, da + 6, db + 3 * kb, idesc, 1);
            umma_ss(acc1, da, db, idesc, 1); umma_ss(acc1, da + 2, db + kb, idesc, 1);
            umma_ss(acc1, da + 4, db + 2 * kb, idesc, 1); umma_ss(acc1, da + 6, db + 3 * kb, idesc, 1);
          }
        }
        tc_commit(&bar);
      }
      __syncwarp();
      long long ti = clock64();
      mbar_wait(&bar, rep & 1);
      t1 = clock64();
      if (rep == 1 && lane_id() == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = ti - t0; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  struct C { const char* name; int ts, n, nacc, b_mn, kadv; } cs[] = {
      {"SS N=64  1 acc", 0, 64, 1, 0, 1},  {"SS N=64  2 acc", 0, 64, 2, 0, 1},  {"SS N=64 1 acc same k", 0, 64, 1, 0, 0},
      {"SS N=128 1 acc", 0, 128, 1, 0, 1}, {"SS N=128 2 acc", 0, 128, 2, 0, 1}, {"SS N=256 1 acc", 0, 256, 1, 0, 1},
      {"SS N=32  1 acc", 0, 32, 1, 0, 1},  {"SS N=16  1 acc", 0, 16, 1, 0, 1},
      {"TS N=64  1 acc (B MN-major)", 1, 64, 1, 1, 1}, {"TS N=64  1 acc (B K-major)", 1, 64, 1, 0, 1},
      {"TS N=128 1 acc (B K-major)", 1, 128, 1, 0, 1}, {"TS N=16 1 acc (B K-major)", 1, 16, 1, 0, 1},
      {"SS N=64 1 acc (B MN-major)", 0, 64, 1, 1, 1},
  };
  for (int grid : {1, 148}) {
    for (auto& c : cs) {
      const int iters = 512;
      k<<<grid, 128, 64 * 1024>>>(c.ts, c.n, c.nacc, c.b_mn, iters, c.kadv, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[2];
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("grid %3d  %-30s: %7.1f clk/MMA executed, %7.1f clk/MMA issue   (%s)\n", grid, c.name, double(h[0]) / iters, double(h[1]) / iters,
             cudaGetErrorString(e));
    }
  }
  return 0;
}
