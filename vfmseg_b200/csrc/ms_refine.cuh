// Kernels of the coarse-to-fine path (BASELINE config 3): MsVFMEncoderDecoder.ms_inference
// (rein/models/segmentors/Ms_VFM_encoder_decoder.py:400-466) and the memory-bound pieces of VFMHead / its
// transformer decoder (rein/models/heads/VFMHead.py:61-89, heads/Transformer.py:52-59,91-92,228-283).
// The dense pieces (1x1 / k2s2 convs as GEMMs, q/k/v/out/ff projections, attention) reuse gemm_sm100.cuh and
// attention_sm100.cuh. Everything here is HBM/L2-bound elementwise or gather work: coalesced accesses, no tensor cores.
//
// Stage 0 of ms_inference produces coarse logits L0 [B, nc, lh, lw] for the whole (down-scaled) image; the reference
// then materialises their bilinear upsampling to the full image (159 MB per 1024x2048 image) and slices 512x512
// "context" windows out of it. Here the upsampled field U is never stored: every consumer (confidence gate, context
// embedding, final merge) samples L0 with PyTorch's upsample_bilinear2d(align_corners=False) arithmetic on the fly.
#pragma once
#include "elementwise.cuh"
#include "sm100_ptx.cuh"

namespace vfm {

// PyTorch area_pixel_compute_source_index + the four taps/weights of upsample_bilinear2d (align_corners=False)
struct Bilerp {
  int o00, dx, dy;      // offset of the top-left tap, +1 column (or 0 at the border), +1 row (or 0)
  float h0, h1, w0, w1;
};
__device__ __forceinline__ Bilerp bilerp_setup(int oy, int ox, float scale_h, float scale_w, int ih, int iw) {
  float sy = scale_h * (oy + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
  float sx = scale_w * (ox + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
  int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  y0 = y0 < ih - 1 ? y0 : ih - 1;
  x0 = x0 < iw - 1 ? x0 : iw - 1;
  Bilerp b;
  b.o00 = y0 * iw + x0;
  b.dy = (y0 < ih - 1) ? iw : 0;
  b.dx = (x0 < iw - 1) ? 1 : 0;
  b.h1 = sy - y0; b.h0 = 1.f - b.h1;
  b.w1 = sx - x0; b.w0 = 1.f - b.w1;
  return b;
}
__device__ __forceinline__ float bilerp_eval(const float* __restrict__ p, const Bilerp& b) {
  return bilerp_rn(b.h0, b.h1, b.w0, b.w1, __ldg(p + b.o00), __ldg(p + b.o00 + b.dx), __ldg(p + b.o00 + b.dy),
                   __ldg(p + b.o00 + b.dy + b.dx));
}

// ---------------------------------------------------------------------------------------------
// resize(inputs, size=(h, w), 'bilinear', align_corners=False) of the network input (Ms_VFM_encoder_decoder.py:413),
// fused with mmseg SegDataPreProcessor for uint8 input (channel flip + (x - mean) / std). Output fp32 [B,3,h,w].
template <typename T>
__global__ void __launch_bounds__(256)
image_resize_norm_kernel(const T* __restrict__ img, int B, int H, int W, PixelNorm nrm, float* __restrict__ out, int h, int w) {
  const float scale_h = static_cast<float>(H) / h, scale_w = static_cast<float>(W) / w;
  const long long total = static_cast<long long>(B) * 3 * h * w;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % w);
    const int y = static_cast<int>((idx / w) % h);
    const int c = static_cast<int>((idx / (static_cast<long long>(w) * h)) % 3);
    const int b = static_cast<int>(idx / (static_cast<long long>(w) * h * 3));
    const Bilerp bl = bilerp_setup(y, x, scale_h, scale_w, H, W);
    float v;
    if constexpr (sizeof(T) == 4) {
      v = bilerp_eval(reinterpret_cast<const float*>(img) + (static_cast<size_t>(b) * 3 + c) * H * W, bl);
    } else {
      const int cs = nrm.flip ? 2 - c : c;
      const uint8_t* p = reinterpret_cast<const uint8_t*>(img) + (static_cast<size_t>(b) * 3 + cs) * H * W;
      const float mu = nrm.mean[c], is = nrm.inv_std[c];
      const float a00 = (static_cast<float>(__ldg(p + bl.o00)) - mu) * is, a01 = (static_cast<float>(__ldg(p + bl.o00 + bl.dx)) - mu) * is;
      const float a10 = (static_cast<float>(__ldg(p + bl.o00 + bl.dy)) - mu) * is;
      const float a11 = (static_cast<float>(__ldg(p + bl.o00 + bl.dy + bl.dx)) - mu) * is;
      v = bilerp_rn(bl.h0, bl.h1, bl.w0, bl.w1, a00, a01, a10, a11);
    }
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// Confidence gate of ms_inference (Ms_VFM_encoder_decoder.py:446-448): for every window k of every image, the number of
// pixels whose max softmax probability of the (upsampled) coarse logits exceeds `thr`. The host turns the counts into
// the reference's `confidence < conf` decision with ONE device->host read per batch (the reference does one .item()
// per window). One CTA per (image, 4 rows); shared counters, one global atomic per window per CTA.
template <int NC_MAX>
__global__ void __launch_bounds__(256)
ms_confidence_kernel(const float* __restrict__ low0, const int2* __restrict__ boxes, int n_crops, int nc, int crop_h,
                     int crop_w, int lh, int lw, int H, int W, float thr, int rows_per_cta, int* __restrict__ counts) {
  extern __shared__ int s_cnt[];   // [n_crops] counters, then the boxes
  int2* s_boxes = reinterpret_cast<int2*>(s_cnt + n_crops);
  for (int i = threadIdx.x; i < n_crops; i += blockDim.x) { s_cnt[i] = 0; s_boxes[i] = boxes[i]; }
  __syncthreads();
  const int ctas_per_img = (H + rows_per_cta - 1) / rows_per_cta;
  const int b = blockIdx.x / ctas_per_img;
  const int y_base = (blockIdx.x - b * ctas_per_img) * rows_per_cta;
  const float scale_h = static_cast<float>(lh) / H, scale_w = static_cast<float>(lw) / W;
  const size_t plane = static_cast<size_t>(lh) * lw;
  const float* img0 = low0 + static_cast<size_t>(b) * nc * plane;
  const int n_pix = rows_per_cta * W;
  for (int i = threadIdx.x; i < n_pix; i += blockDim.x) {
    const int y = y_base + i / W, x = i % W;
    if (y >= H) break;
    const Bilerp bl = bilerp_setup(y, x, scale_h, scale_w, lh, lw);
    float u[NC_MAX];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c) {
      if (c < nc) { u[c] = bilerp_eval(img0 + c * plane, bl); m = fmaxf(m, u[c]); }
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c)
      if (c < nc) s += expf(u[c] - m);
    if (1.f / s > thr) {   // max softmax = exp(0) / sum
      for (int k = 0; k < n_crops; ++k) {
        const int cy = y - s_boxes[k].x, cx = x - s_boxes[k].y;
        if (cy >= 0 && cy < crop_h && cx >= 0 && cx < crop_w) atomicAdd(&s_cnt[k], 1);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_crops; i += blockDim.x)
    if (s_cnt[i]) atomicAdd(&counts[b * n_crops + i], s_cnt[i]);
}

// ---------------------------------------------------------------------------------------------
// Operand of VFMHead.seg_logits_embed[0] (Conv2d(nc -> C/4, k=2, s=2); VFMHead.py:38-39) for the windows that are
// refined: context = crop(U, box) -> resize to (ctx_h, ctx_w) (VFMHead.py:63-67) -> non-overlapping 2x2 patches.
// out[(r * ctx_h/2 + oy) * ctx_w/2 + ox, cin * 4 + dy * 2 + dx] (bf16, row pitch kpad >= 4 nc, zero padded) —
// the column order of a flattened Conv2d weight [C_out, nc, 2, 2].  crops[r] = {image, y1, x1, 0}.
__global__ void __launch_bounds__(256)
ms_context_im2col_kernel(const float* __restrict__ low0, const int4* __restrict__ crops, int n_ref, int nc, int crop_h,
                         int crop_w, int lh, int lw, int H, int W, int ctx_h, int ctx_w, __nv_bfloat16* __restrict__ out, int kpad) {
  const int oh = ctx_h / 2, ow = ctx_w / 2, cin_slots = kpad / 4;
  const float up_h = static_cast<float>(lh) / H, up_w = static_cast<float>(lw) / W;             // L0 -> full image
  const float dn_h = static_cast<float>(crop_h) / ctx_h, dn_w = static_cast<float>(crop_w) / ctx_w;   // window -> context
  const size_t plane = static_cast<size_t>(lh) * lw;
  const long long total = static_cast<long long>(n_ref) * oh * ow * cin_slots;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cin = static_cast<int>(idx % cin_slots);
    long long t = idx / cin_slots;
    const int ox = static_cast<int>(t % ow); t /= ow;
    const int oy = static_cast<int>(t % oh);
    const int r = static_cast<int>(t / oh);
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (cin < nc) {
      const int4 cb = __ldg(crops + r);
      const float* p = low0 + (static_cast<size_t>(cb.x) * nc + cin) * plane;
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int i = 2 * oy + (d >> 1), j = 2 * ox + (d & 1);     // context pixel
        // bilinear sample of the window (its own border clamps) at context pixel (i, j)
        float sy = dn_h * (i + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
        float sx = dn_w * (j + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
        int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
        y0 = y0 < crop_h - 1 ? y0 : crop_h - 1;
        x0 = x0 < crop_w - 1 ? x0 : crop_w - 1;
        const int y1 = y0 < crop_h - 1 ? y0 + 1 : y0, x1 = x0 < crop_w - 1 ? x0 + 1 : x0;
        const float h1 = sy - y0, h0 = 1.f - h1, w1 = sx - x0, w0 = 1.f - w1;
        const float u00 = bilerp_eval(p, bilerp_setup(cb.y + y0, cb.z + x0, up_h, up_w, lh, lw));
        const float u01 = bilerp_eval(p, bilerp_setup(cb.y + y0, cb.z + x1, up_h, up_w, lh, lw));
        const float u10 = bilerp_eval(p, bilerp_setup(cb.y + y1, cb.z + x0, up_h, up_w, lh, lw));
        const float u11 = bilerp_eval(p, bilerp_setup(cb.y + y1, cb.z + x1, up_h, up_w, lh, lw));
        v[d] = bilerp_rn(h0, h1, w0, w1, u00, u01, u10, u11);
      }
    }
    const size_t row = (static_cast<size_t>(r) * oh + oy) * ow + ox;
    *reinterpret_cast<uint2*>(out + row * kpad + cin * 4) = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
  }
}

// ---------------------------------------------------------------------------------------------
// Operand of a Conv2d(C -> C', k=2, s=2) over token-major activations (VFMHead.py:42): in [n*h*w, C] ->
// out [n*(h/2)*(w/2), 4C], column = (dy * 2 + dx) * C + c. 16-byte moves (C % 8 == 0).
__global__ void __launch_bounds__(256)
space_to_depth2_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int n, int h, int w, int C) {
  const int vecs = C / 8, oh = h / 2, ow = w / 2;
  const long long total = static_cast<long long>(n) * oh * ow * 4 * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(idx % vecs);
    long long t = idx / vecs;
    const int d = static_cast<int>(t % 4); t /= 4;
    const int ox = static_cast<int>(t % ow); t /= ow;
    const int oy = static_cast<int>(t % oh);
    const int i = static_cast<int>(t / oh);
    const size_t src = ((static_cast<size_t>(i) * h + 2 * oy + (d >> 1)) * w + 2 * ox + (d & 1)) * C + v * 8;
    const size_t dst = ((static_cast<size_t>(i) * oh + oy) * ow + ox) * (4 * static_cast<size_t>(C)) + d * C + v * 8;
    *reinterpret_cast<uint4*>(out + dst) = __ldg(reinterpret_cast<const uint4*>(in + src));
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm(groups, C) (+ activation) over token-major activations [n * P, C] for any channels-per-group
// (VFMHead.py:30-32,40-47: 2, 4 and 8 channels per group; Transformer.py:91-92 Normalize with eps 1e-6 whose output
// is the fp32 residual stream of the decoder). act: 0 none, 1 ReLU, 2 exact-erf GELU. One CTA per (sample, group).
template <typename OutT>
__global__ void __launch_bounds__(256)
groupnorm_act_kernel(const __nv_bfloat16* __restrict__ in, OutT* __restrict__ out, const float* __restrict__ gamma,
                     const float* __restrict__ beta, int P, int C, int groups, float eps, int act) {
  __shared__ float red[8];
  const int cg = C / groups;
  const int smp = blockIdx.x / groups, g = blockIdx.x - smp * groups;
  const __nv_bfloat16* base = in + static_cast<size_t>(smp) * P * C + g * cg;
  OutT* obase = out + static_cast<size_t>(smp) * P * C + g * cg;
  const int n_el = P * cg;
  float s = 0.f;
  for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
    const int tok = i / cg, c = i - tok * cg;
    s += __bfloat162float(base[static_cast<size_t>(tok) * C + c]);
  }
  const float n = static_cast<float>(n_el);
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
    const int tok = i / cg, c = i - tok * cg;
    const float d = __bfloat162float(base[static_cast<size_t>(tok) * C + c]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / n + eps);
  for (int i = threadIdx.x; i < n_el; i += blockDim.x) {
    const int tok = i / cg, c = i - tok * cg;
    float y = (__bfloat162float(base[static_cast<size_t>(tok) * C + c]) - mean) * rstd * __ldg(gamma + g * cg + c) + __ldg(beta + g * cg + c);
    if (act == 1) y = fmaxf(y, 0.f);
    else if (act == 2) y = gelu_erf(y);
    if constexpr (sizeof(OutT) == 4) obase[static_cast<size_t>(tok) * C + c] = y;
    else obase[static_cast<size_t>(tok) * C + c] = __float2bfloat16(y);
  }
}

// ---------------------------------------------------------------------------------------------
// GEGLU (Transformer.py:52-59): in [M, 2I] = (x | gate) -> out [M, I] = x * gelu(gate), exact-erf GELU.
__global__ void __launch_bounds__(256)
geglu_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, long long M, int I) {
  const int vecs = I / 8;
  const long long total = M * vecs;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long row = idx / vecs;
    const int v = static_cast<int>(idx - row * vecs);
    const uint4 xa = __ldg(reinterpret_cast<const uint4*>(in + row * 2 * I + v * 8));
    const uint4 ga = __ldg(reinterpret_cast<const uint4*>(in + row * 2 * I + I + v * 8));
    const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xa);
    const __nv_bfloat162* gh = reinterpret_cast<const __nv_bfloat162*>(&ga);
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = __bfloat1622float2(xh[k]), g = __bfloat1622float2(gh[k]);
      o[k] = pack_bf16x2(x.x * gelu_erf(g.x), x.y * gelu_erf(g.y));
    }
    *reinterpret_cast<uint4*>(out + row * I + v * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// fp32 -> bf16 (the decoder's fp32 residual stream feeding the conv_seg GEMM); n % 4 == 0
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(in) + i);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  }
}

// ---------------------------------------------------------------------------------------------
// Stage 1 merge of ms_inference (Ms_VFM_encoder_decoder.py:449-461) + postprocess argmax: per pixel, every covering
// window contributes, in row-major window order, either its refined logits (aux decoder output [rh, rw] resized to
// the window) or the context value U(y, x) (the window was confident enough to be skipped); sum / count; argmax.
// ref_index[b * n_crops + k] = row of `refined` holding that window's logits, or -1.
template <int NC_MAX>
__global__ void __launch_bounds__(256)
ms_merge_argmax_kernel(const float* __restrict__ low0, const float* __restrict__ refined, const int* __restrict__ ref_index,
                       const int2* __restrict__ boxes, int n_crops, int nc, int crop_h, int crop_w, int lh, int lw, int rh,
                       int rw, int H, int W, int n_img, uint8_t* __restrict__ labels, float* __restrict__ logits_out) {
  extern __shared__ int2 s_boxes[];
  for (int i = threadIdx.x; i < n_crops; i += blockDim.x) s_boxes[i] = boxes[i];
  __syncthreads();
  const float up_h = static_cast<float>(lh) / H, up_w = static_cast<float>(lw) / W;
  const float r_h = static_cast<float>(rh) / crop_h, r_w = static_cast<float>(rw) / crop_w;
  const size_t plane0 = static_cast<size_t>(lh) * lw, plane_r = static_cast<size_t>(rh) * rw;
  const long long total = static_cast<long long>(n_img) * H * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int b = static_cast<int>(idx / (static_cast<long long>(W) * H));
    float acc[NC_MAX], ctx[NC_MAX];
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c) acc[c] = 0.f;
    bool have_ctx = false;
    int count = 0;
    for (int k = 0; k < n_crops; ++k) {
      const int cy = y - s_boxes[k].x, cx = x - s_boxes[k].y;
      if (cy < 0 || cy >= crop_h || cx < 0 || cx >= crop_w) continue;
      ++count;
      const int ri = __ldg(ref_index + b * n_crops + k);
      if (ri >= 0) {
        const Bilerp bl = bilerp_setup(cy, cx, r_h, r_w, rh, rw);
        const float* p = refined + static_cast<size_t>(ri) * nc * plane_r;
#pragma unroll
        for (int c = 0; c < NC_MAX; ++c)
          if (c < nc) acc[c] = __fadd_rn(acc[c], bilerp_eval(p + c * plane_r, bl));
      } else {
        if (!have_ctx) {
          const Bilerp bl = bilerp_setup(y, x, up_h, up_w, lh, lw);
          const float* p = low0 + static_cast<size_t>(b) * nc * plane0;
#pragma unroll
          for (int c = 0; c < NC_MAX; ++c)
            if (c < nc) ctx[c] = bilerp_eval(p + c * plane0, bl);
          have_ctx = true;
        }
#pragma unroll
        for (int c = 0; c < NC_MAX; ++c)
          if (c < nc) acc[c] = __fadd_rn(acc[c], ctx[c]);
      }
    }
    const float cnt = static_cast<float>(count);
    int best = 0;
    float bestv = acc[0] / cnt;
    const size_t pix = static_cast<size_t>(y) * W + x;
    float* lo = logits_out ? logits_out + static_cast<size_t>(b) * nc * H * W + pix : nullptr;
    if (lo) lo[0] = bestv;
#pragma unroll
    for (int c = 1; c < NC_MAX; ++c) {
      if (c < nc) {
        const float v = acc[c] / cnt;
        if (lo) lo[static_cast<size_t>(c) * H * W] = v;
        if (v > bestv) { bestv = v; best = c; }
      }
    }
    labels[idx] = static_cast<uint8_t>(best);
  }
}

}  // namespace vfm
