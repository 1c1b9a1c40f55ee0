// HBM-bound kernels either side of the tensor-core work: patch gather (+ pixel normalisation),
// cls row, LayerNorm, GroupNorm+ReLU. All coalesced 16-byte accesses; no tensor cores.
#pragma once
#include "sm100_ptx.cuh"

namespace vfm {

// ---------------------------------------------------------------------------------------------
// Patch gather ("im2col" for a k = s = 16 conv; patch_embed.py:66,75-77) fused with the crop
// slicing of slide_inference (Ms_VFM_encoder_decoder.py:433-442) and, for uint8 input, with
// mmseg SegDataPreProcessor (channel flip + (x - mean) / std; lora_dinov2_linear.py:13-21).
// out[(crop * gh * gw + gy * gw + gx), c * 256 + py * 16 + px] = img[b, c, y1 + gy*16 + py, x1 + gx*16 + px]
// crops[i] = {image index, y1, x1, 0}.
struct PixelNorm {  // only used by the uint8 path
  float mean[3];
  float inv_std[3];
  int flip;  // 1: input channel order is BGR, network wants RGB
};

template <typename T>
__global__ void __launch_bounds__(256)
patch_gather_kernel(const T* __restrict__ img, int img_h, int img_w, const int4* __restrict__ crops, int n_crops,
                    int gh, int gw, PixelNorm nrm, __nv_bfloat16* __restrict__ out) {
  // one thread = 8 consecutive pixels of one patch row
  const long long total = static_cast<long long>(n_crops) * gh * gw * 3 * 16 * 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    int half = static_cast<int>(idx & 1);
    int py = static_cast<int>((idx >> 1) & 15);
    long long t = idx >> 5;
    int c = static_cast<int>(t % 3);
    t /= 3;
    int gx = static_cast<int>(t % gw);
    t /= gw;
    int gy = static_cast<int>(t % gh);
    int crop = static_cast<int>(t / gh);
    int4 cb = __ldg(crops + crop);
    int y = cb.y + gy * 16 + py, x = cb.z + gx * 16 + half * 8;
    float v[8];
    if constexpr (sizeof(T) == 4) {
      const float* src = reinterpret_cast<const float*>(img) + ((static_cast<size_t>(cb.x) * 3 + c) * img_h + y) * img_w + x;
      if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
      }
    } else {
      int cs = nrm.flip ? 2 - c : c;  // output channel c reads stored channel cs
      const uint8_t* src = reinterpret_cast<const uint8_t*>(img) + ((static_cast<size_t>(cb.x) * 3 + cs) * img_h + y) * img_w + x;
      // normalisation constants are indexed by the *output* (RGB) channel, as in mmseg
      float mu = nrm.mean[c], is = nrm.inv_std[c];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (static_cast<float>(__ldg(src + i)) - mu) * is;
    }
    size_t row = (static_cast<size_t>(crop) * gh + gy) * gw + gx;
    uint4* dst = reinterpret_cast<uint4*>(out + row * 768 + c * 256 + py * 16 + half * 8);
    *dst = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// x[crop * tokens, :] = cls_token + pos_embed[0]   (dino_v2.py:225-226)
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                int n_crops, int tokens, int C) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_crops * C) return;
  int crop = i / C, c = i - crop * C;
  x[static_cast<size_t>(crop) * tokens * C + c] = cls[c] + pos[c];
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over the channel dim of the fp32 residual stream, bf16 out (block.py:63,75 with
// eps = 1e-6 from dino_v2.py:104). One warp per row, row held in registers, two-pass variance.
// Optionally also snapshots the un-normalised row as a bf16 feature tap with the cls row dropped
// (dino_v2.py:261-267): the residual stream is read here anyway, so the tap costs only its own write and every
// residual GEMM can use the TMA reduce-add epilogue. `out` may be null (tap only, after the last block).
template <int ITERS>  // C = 128 * ITERS
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                 __nv_bfloat16* __restrict__ out, int M, float eps, __nv_bfloat16* __restrict__ tap, int tap_ld,
                 int tap_col0, FastDiv tokens_per_crop, int cls_rows, const int* __restrict__ out_map) {
  constexpr int C = 128 * ITERS;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
  float4 v[ITERS];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  if (tap != nullptr) {
    int crop = 0, tok = 1;
    if (cls_rows) tokens_per_crop.divmod(row, crop, tok);   // cls_rows = 0 (SAM ViT: no cls token): every row is tapped
    if (tok != 0) {
      uint2* trow = reinterpret_cast<uint2*>(tap + static_cast<size_t>(row - (cls_rows ? crop + 1 : 0)) * tap_ld + tap_col0);
#pragma unroll
      for (int i = 0; i < ITERS; ++i)
        trow[i * 32 + lane] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
    }
  }
  if (out == nullptr) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q * (1.f / C) + eps);
  // out_map (optional): destination row of each input row — SAM's window partition writes the normalised tokens straight
  // into window order (the zero rows of the padding are never written: the caller keeps them zero)
  uint2* orow = reinterpret_cast<uint2*>(out + static_cast<size_t>(out_map ? __ldg(out_map + row) : row) * C);
#pragma unroll
  for (int i = 0; i < ITERS; ++i) {
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    float y0 = (v[i].x - mean) * rstd * g.x + b.x, y1 = (v[i].y - mean) * rstd * g.y + b.y;
    float y2 = (v[i].z - mean) * rstd * g.z + b.z, y3 = (v[i].w - mean) * rstd * g.w + b.w;
    orow[i * 32 + lane] = make_uint2(pack_bf16x2(y0, y1), pack_bf16x2(y2, y3));
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(groups, C) + ReLU over token-major activations [n_crops * P, C] (mmcv ConvModule
// order conv -> norm -> act, linear_head.py:36-40; torch GroupNorm: biased variance, eps 1e-5).
// One CTA per (crop, group); the (P x C/groups) slab is L1/L2 resident across the three passes.
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256)
groupnorm_relu_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ gamma,
                      const float* __restrict__ beta, int P, int C, int groups, float eps, int relu) {
  __shared__ float red[8];
  const int cg = C / groups;        // channels per group (multiple of 8)
  const int vec_per_tok = cg / 8;   // uint4 per token
  const int crop = blockIdx.x / groups, g = blockIdx.x - crop * groups;
  const __nv_bfloat16* base = in + static_cast<size_t>(crop) * P * C + g * cg;
  __nv_bfloat16* obase = out + static_cast<size_t>(crop) * P * C + g * cg;
  const int nvec = P * vec_per_tok;
  float s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    int tok = i / vec_per_tok, j = i - tok * vec_per_tok;
    uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(tok) * C + j * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); s += f.x + f.y; }
  }
  const float n = static_cast<float>(P) * cg;
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    int tok = i / vec_per_tok, j = i - tok * vec_per_tok;
    uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(tok) * C + j * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { float2 f = __bfloat1622float2(h[k]); float a = f.x - mean, b = f.y - mean; q += a * a + b * b; }
  }
  const float rstd = rsqrtf(block_sum(q, red) / n + eps);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    int tok = i / vec_per_tok, j = i - tok * vec_per_tok;
    uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(tok) * C + j * 8);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float2 f = __bfloat1622float2(h[k]);
      int ch = g * cg + j * 8 + 2 * k;
      float y0 = (f.x - mean) * rstd * __ldg(gamma + ch) + __ldg(beta + ch);
      float y1 = (f.y - mean) * rstd * __ldg(gamma + ch + 1) + __ldg(beta + ch + 1);
      if (relu) { y0 = fmaxf(y0, 0.f); y1 = fmaxf(y1, 0.f); }
      o[k] = pack_bf16x2(y0, y1);
    }
    *reinterpret_cast<uint4*>(obase + static_cast<size_t>(tok) * C + j * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace vfm
