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
constexpr int SWIGLU_MAX_ITERS = 12;   // Hp <= 12 * 32 * 8 = 3072
__global__ void __launch_bounds__(256)
swiglu_layernorm_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
                        const float* __restrict__ beta, long long M, int H, int Hp, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int chunks = Hp / 8;
  const __nv_bfloat16* r1 = in + row * 2 * Hp;
  const __nv_bfloat16* r2 = r1 + Hp;
  float v[SWIGLU_MAX_ITERS][8];
  float s = 0.f;
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * 32 + lane;
    if (c < chunks) {
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(r1) + c), b = __ldg(reinterpret_cast<const uint4*>(r2) + c);
      const __nv_bfloat162* ah = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* bh = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 x1 = __bfloat1622float2(ah[k]), x2 = __bfloat1622float2(bh[k]);
        const float h0 = x1.x * fast_rcp(1.f + fast_exp2(-1.4426950408889634f * x1.x)) * x2.x;
        const float h1 = x1.y * fast_rcp(1.f + fast_exp2(-1.4426950408889634f * x1.y)) * x2.y;
        v[it][2 * k] = (c * 8 + 2 * k < H) ? h0 : 0.f;
        v[it][2 * k + 1] = (c * 8 + 2 * k + 1 < H) ? h1 : 0.f;
        s += v[it][2 * k] + v[it][2 * k + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / H;
  float q = 0.f;
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * 32 + lane;
    if (c < chunks) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = (c * 8 + e < H) ? v[it][e] - mean : 0.f;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / H + eps);
#pragma unroll
  for (int it = 0; it < SWIGLU_MAX_ITERS; ++it) {
    const int c = it * 32 + lane;
    if (c < chunks) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = c * 8 + e;
        y[e] = col < H ? (v[it][e] - mean) * rstd * __ldg(gamma + col) + __ldg(beta + col) : 0.f;
      }
      reinterpret_cast<uint4*>(out + row * Hp)[c] =
          make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
    }
  }
}

}  // namespace vfm
