// Memory-bound kernels specific to the EVA02 backbone (BASELINE config 4; rein/models/backbones/eva_02.py):
// 2-D rotary position embedding of q and k (:119-160, applied at :362-369 to the patch tokens only) and the SwiGLU
// feed-forward's elementwise part fused with its inner LayerNorm (:204-242). The dense parts reuse the tcgen05 GEMMs
// and the fused attention kernel.
#pragma once
#include "sm100_ptx.cuh"

namespace vfm {

// In-place RoPE on the packed qkv activations [M, 3C] (bf16): for every patch token (token 0 of each sequence, the cls
// token, is skipped, eva_02.py:363,367) and every head, t * cos + rotate_half(t) * sin with
// rotate_half(x)[2i] = -x[2i+1], rotate_half(x)[2i+1] = x[2i] (:54-58). cos/sin: fp32 [tokens_per_seq - 1, 64].
__global__ void __launch_bounds__(256)
rope_qk_kernel(__nv_bfloat16* __restrict__ qkv, long long M, int C, int heads, FastDiv tokens_per_seq,
               const float* __restrict__ cos_t, const float* __restrict__ sin_t) {
  const long long total = M * 2 * heads * 8;   // one thread = 8 consecutive dims of one (row, q|k, head)
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int chunk = static_cast<int>(idx & 7);
    long long t = idx >> 3;
    const int head = static_cast<int>(t % heads); t /= heads;
    const int which = static_cast<int>(t & 1);
    const long long row = t >> 1;
    int seq, tok;
    tokens_per_seq.divmod(static_cast<int>(row), seq, tok);
    if (tok == 0) continue;
    uint4* ptr = reinterpret_cast<uint4*>(qkv + row * 3 * C + which * C + head * 64 + chunk * 8);
    const uint4 u = *ptr;
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float4* cp = reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(tok - 1) * 64 + chunk * 8);
    const float4* sp = reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(tok - 1) * 64 + chunk * 8);
    const float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
    const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = __bfloat1622float2(h[k]);
      o[k] = pack_bf16x2(x.x * cs[2 * k] - x.y * sn[2 * k], x.y * cs[2 * k + 1] + x.x * sn[2 * k + 1]);
    }
    *ptr = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// SwiGLU.forward between its GEMMs (eva_02.py:234-239): in [M, 2*Hp] = (w1 x | w2 x) bf16 -> hidden = silu(x1) * x2 ->
// LayerNorm over the H real columns (ffn_ln, subln=True) -> out [M, Hp] bf16; columns [H, Hp) (the padding that makes
// the 2730-wide hidden layer TMA-addressable) are written as zeros. One warp per row, row in registers.
// 128 threads per row, two rows per CTA: each thread holds at most three 8-element chunks, so all of a row's loads are in
// flight at once and 16 rows are resident per SM (one warp per row with the whole row in 96 registers left each warp with
// one or two loads outstanding: 246 us per launch against 95 us of HBM traffic).
constexpr int SWIGLU_TPR = 128;
constexpr int SWIGLU_MAX_ITERS = 3;    // Hp <= 3 * 128 * 8 = 3072
__global__ void __launch_bounds__(256)
swiglu_layernorm_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
                        const float* __restrict__ beta, long long M, int H, int Hp, float eps) {
  __shared__ float red[2][2][SWIGLU_TPR / 32];   // [row slot][sum | sum of squares][warp]
  const int slot = threadIdx.x / SWIGLU_TPR, t = threadIdx.x % SWIGLU_TPR;
  const int lane = threadIdx.x & 31, wr = t >> 5;
  const long long row = blockIdx.x * 2LL + slot;
  const bool live = row < M;
  const int chunks = Hp / 8;
  const __nv_bfloat16* r1 = in + (live ? row : 0) * 2 * Hp;
  const __nv_bfloat16* r2 = r1 + Hp;
  float v[SWIGLU_MAX_ITERS][8];
  uint4 a[SWIGLU_MAX_ITERS], b[SWIGLU_MAX_ITERS];
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * SWIGLU_TPR + t;
    if (live && c < chunks) { a[it] = __ldg(reinterpret_cast<const uint4*>(r1) + c); b[it] = __ldg(reinterpret_cast<const uint4*>(r2) + c); }
  }
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * SWIGLU_TPR + t;
#pragma unroll
    for (int e = 0; e < 8; ++e) v[it][e] = 0.f;
    if (live && c < chunks) {
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a[it]);
      const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b[it]);
      const bool whole = c * 8 + 8 <= H;   // every chunk but the last: no per-element column test
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x1 = __bfloat1622float2(ah[k]), x2 = __bfloat1622float2(bh[k]);
        // silu(x) = x sigmoid(x) = u + u tanh(u), u = x / 2: ONE MUFU op (tanh.approx, |error| < 2^-10.9) instead of ex2 + rcp
        const float u0 = 0.5f * x1.x, u1 = 0.5f * x1.y;
        const float h0 = fmaf(u0, fast_tanh(u0), u0) * x2.x;
        const float h1 = fmaf(u1, fast_tanh(u1), u1) * x2.y;
        v[it][2 * k] = (whole || c * 8 + 2 * k < H) ? h0 : 0.f;
        v[it][2 * k + 1] = (whole || c * 8 + 2 * k + 1 < H) ? h1 : 0.f;
        s += v[it][2 * k] + v[it][2 * k + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[slot][0][wr] = s;
  __syncthreads();
  const float mean = (red[slot][0][0] + red[slot][0][1] + red[slot][0][2] + red[slot][0][3]) / H;
  float q = 0.f;
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * SWIGLU_TPR + t;
    if (c < chunks) {
      const bool whole = c * 8 + 8 <= H;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = (whole || c * 8 + e < H) ? v[it][e] - mean : 0.f;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if (lane == 0) red[slot][1][wr] = q;
  __syncthreads();
  if (!live) return;
  const float rstd = rsqrtf((red[slot][1][0] + red[slot][1][1] + red[slot][1][2] + red[slot][1][3]) / H + eps);
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15) == 0;
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * SWIGLU_TPR + t;
    if (c < chunks) {
      float y[8];
      if (c * 8 + 8 <= H && vec_ok) {
        // whole chunk inside the row: gamma / beta as two 16-byte loads each (the scalar form touched 8 cache lines per
        // load instruction)
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = (v[it][e] - mean) * rstd * g[e] + bb[e];
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int col = c * 8 + e;
          y[e] = col < H ? (v[it][e] - mean) * rstd * __ldg(gamma + col) + __ldg(beta + col) : 0.f;
        }
      }
      reinterpret_cast<uint4*>(out + row * Hp)[c] =
          make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    }
  }
}

}  // namespace vfm
