// The memory-bound tail of slide inference: overlapped-window logit merge + argmax, and the
// confusion-matrix histogram behind mIoU.
#pragma once
#include "ms_refine.cuh"
#include "sm100_ptx.cuh"

namespace vfm {

// ---------------------------------------------------------------------------------------------
// slide_merge_argmax: one pass replaces, per image,
//   * the bilinear resize of every crop's low-res logits to the crop size
//     (mmseg BaseDecodeHead.predict_by_feat -> resize(align_corners=False); rein/utils/wrappers.py:9-28),
//   * preds += F.pad(crop_logit, ...); count_mat[box] += 1; preds / count_mat
//     (Ms_VFM_encoder_decoder.py:455-461, the in-repo copy of mmseg slide_inference),
//   * argmax(dim=0), first maximum wins (mmseg BaseSegmentor.postprocess_result).
// The 159 MB/image fp32 full-resolution logits are only written when `logits_out` is given.
// Accumulation order = row-major crop order, as in the reference loop.
//
// lowres : [n_img * n_crops, NC, lh, lw] fp32 (crop k of image b at index b * n_crops + k)
// boxes  : n_crops x {y1, x1} (top-left corner of each crop window, full-res pixels)
template <int NC_MAX>
__global__ void __launch_bounds__(256)
slide_merge_argmax_kernel(const float* __restrict__ lowres, const int2* __restrict__ boxes, int n_crops, int nc,
                          int crop_h, int crop_w, int lh, int lw, int H, int W, int n_img,
                          uint8_t* __restrict__ labels, float* __restrict__ logits_out) {
  extern __shared__ int2 s_boxes[];
  for (int i = threadIdx.x; i < n_crops; i += blockDim.x) s_boxes[i] = boxes[i];
  __syncthreads();
  const float scale_h = static_cast<float>(lh) / crop_h, scale_w = static_cast<float>(lw) / crop_w;
  const long long total = static_cast<long long>(n_img) * H * W;
  const size_t plane = static_cast<size_t>(lh) * lw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % W);
    const int y = static_cast<int>((idx / W) % H);
    const int b = static_cast<int>(idx / (static_cast<long long>(W) * H));
    float acc[NC_MAX];
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c) acc[c] = 0.f;
    int count = 0;
    for (int k = 0; k < n_crops; ++k) {
      const int cy = y - s_boxes[k].x, cx = x - s_boxes[k].y;
      if (cy < 0 || cy >= crop_h || cx < 0 || cx >= crop_w) continue;
      ++count;
      // PyTorch upsample_bilinear2d, align_corners=False
      float sy = scale_h * (cy + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
      float sx = scale_w * (cx + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
      const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
      const int yp = (y0 < lh - 1) ? lw : 0, xp = (x0 < lw - 1) ? 1 : 0;
      const float h1 = sy - y0, h0 = 1.f - h1, w1 = sx - x0, w0 = 1.f - w1;
      const float* p = lowres + (static_cast<size_t>(b) * n_crops + k) * nc * plane + static_cast<size_t>(y0) * lw + x0;
#pragma unroll
      for (int c = 0; c < NC_MAX; ++c) {
        if (c < nc) {
          const float* q = p + c * plane;
          acc[c] = __fadd_rn(acc[c], bilerp_rn(h0, h1, w0, w1, __ldg(q), __ldg(q + xp), __ldg(q + yp), __ldg(q + yp + xp)));
        }
      }
    }
    const float cnt = static_cast<float>(count);  // >= 1: the grid covers the image (count_mat assert in the reference)
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

// Same result, four horizontally adjacent pixels per thread (W % 4 == 0): the four pixels lie within one low-res step of
// every window, so their bilinear taps come from at most three low-res columns x two rows — 6 loads per class and window
// instead of 16 — and the window tests run once per strip. Each pixel evaluates exactly the expression of the kernel
// above (same operands, same order); labels leave as one 32-bit store, the optional logits as float4.
template <int NC_MAX>
__global__ void __launch_bounds__(256)
slide_merge_argmax4_kernel(const float* __restrict__ lowres, const int2* __restrict__ boxes, int n_crops, int nc,
                           int crop_h, int crop_w, int lh, int lw, int H, int W, int n_img,
                           uint8_t* __restrict__ labels, float* __restrict__ logits_out) {
  extern __shared__ int2 s_boxes[];
  for (int i = threadIdx.x; i < n_crops; i += blockDim.x) s_boxes[i] = boxes[i];
  __syncthreads();
  const float scale_h = static_cast<float>(lh) / crop_h, scale_w = static_cast<float>(lw) / crop_w;
  const int W4 = W >> 2;
  const long long total = static_cast<long long>(n_img) * H * W4;
  const size_t plane = static_cast<size_t>(lh) * lw;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xs = static_cast<int>(idx % W4) * 4;
    const int y = static_cast<int>((idx / W4) % H);
    const int b = static_cast<int>(idx / (static_cast<long long>(W4) * H));
    float acc[4][NC_MAX];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < NC_MAX; ++c) acc[j][c] = 0.f;
    int count[4] = {0, 0, 0, 0};
    for (int k = 0; k < n_crops; ++k) {
      const int cy = y - s_boxes[k].x, cx0 = xs - s_boxes[k].y;
      if (cy < 0 || cy >= crop_h || cx0 + 3 < 0 || cx0 >= crop_w) continue;
      // PyTorch upsample_bilinear2d, align_corners=False
      float sy = scale_h * (cy + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
      const int y0 = static_cast<int>(sy);
      const int yp = (y0 < lh - 1) ? lw : 0;
      const float h1 = sy - y0, h0 = 1.f - h1;
      bool in[4]; int x0[4]; float w0[4], w1[4];
      int base = -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cx = cx0 + j;
        in[j] = cx >= 0 && cx < crop_w;
        float sx = scale_w * (cx + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
        x0[j] = static_cast<int>(sx);
        w1[j] = sx - x0[j]; w0[j] = 1.f - w1[j];
        if (in[j]) { ++count[j]; if (base < 0) base = x0[j]; }
      }
      bool narrow = true;   // all x0 within {base, base + 1}: true whenever the window is upsampled by >= 4 (it is x4 here)
#pragma unroll
      for (int j = 0; j < 4; ++j) narrow = narrow && (!in[j] || (x0[j] - base) <= 1);
      const int c1 = min(base + 1, lw - 1), c2 = min(base + 2, lw - 1);
      const float* p = lowres + (static_cast<size_t>(b) * n_crops + k) * nc * plane + static_cast<size_t>(y0) * lw;
#pragma unroll
      for (int c = 0; c < NC_MAX; ++c) {
        if (c < nc) {
          const float* q = p + c * plane;
          if (narrow) {
            const float a0 = __ldg(q + base), a1 = __ldg(q + c1), a2 = __ldg(q + c2);
            const float b0 = __ldg(q + yp + base), b1 = __ldg(q + yp + c1), b2 = __ldg(q + yp + c2);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (in[j]) {
                const bool first = x0[j] == base;
                const float tl = first ? a0 : a1, tr = first ? a1 : a2, bl = first ? b0 : b1, br = first ? b1 : b2;
                acc[j][c] = __fadd_rn(acc[j][c], bilerp_rn(h0, h1, w0[j], w1[j], tl, tr, bl, br));
              }
            }
          } else {   // general scale: per-pixel taps
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (in[j]) {
                const int xp = (x0[j] < lw - 1) ? 1 : 0;
                const float* r = q + x0[j];
                acc[j][c] = __fadd_rn(acc[j][c], bilerp_rn(h0, h1, w0[j], w1[j], __ldg(r), __ldg(r + xp), __ldg(r + yp), __ldg(r + yp + xp)));
              }
            }
          }
        }
      }
    }
    const size_t pix = static_cast<size_t>(y) * W + xs;
    uint32_t lab = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float cnt = static_cast<float>(count[j]);  // >= 1: the grid covers the image
      int best = 0;
      acc[j][0] = acc[j][0] / cnt;
      float bestv = acc[j][0];
#pragma unroll
      for (int c = 1; c < NC_MAX; ++c) {
        if (c < nc) {
          acc[j][c] = acc[j][c] / cnt;
          if (acc[j][c] > bestv) { bestv = acc[j][c]; best = c; }
        }
      }
      lab |= static_cast<uint32_t>(best) << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(labels + static_cast<size_t>(b) * H * W + pix) = lab;
    if (logits_out) {
      float* lo = logits_out + static_cast<size_t>(b) * nc * H * W + pix;
#pragma unroll
      for (int c = 0; c < NC_MAX; ++c)
        if (c < nc) *reinterpret_cast<float4*>(lo + static_cast<size_t>(c) * H * W) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
    }
  }
}

// Same result, staged through shared memory: one CTA per 64 x 16 tile of output pixels; for every window that overlaps
// the tile (in row-major window order, as the reference accumulates) the CTA copies the low-res footprint of the tile —
// at most 6 rows x 18 columns per class — into shared memory with all loads in flight, then every thread resamples its
// strip of four pixels from there. The per-pixel gather kernels above spend their time in chains of L2-latency loads
// (0.65 ms per image against ~4 us of traffic); this one reads each low-res value once per tile.
// Requires W % 4 == 0 and an exact x4 upsampling (crop = 4 x low-res), which is what LinearHead produces; tiles that hang
// over the right / bottom edge (e.g. the 1820 x 1024 BDD100K test size, configs/_base_/datasets/bdd100k_1024x1024.py:15)
// run with their outside strips idle.
constexpr int MERGE_TW = 64, MERGE_TH = 16, MERGE_FR = 6, MERGE_FC = 20;   // tile, footprint rows / padded columns
template <int NC_MAX>
__global__ void __launch_bounds__(256)
slide_merge_tile_kernel(const float* __restrict__ lowres, const int2* __restrict__ boxes, int n_crops, int nc,
                        int crop_h, int crop_w, int lh, int lw, int H, int W,
                        uint8_t* __restrict__ labels, float* logits_out, const float* flip_a) {
  __shared__ float fp[NC_MAX * MERGE_FR * MERGE_FC];
  const int tx0 = blockIdx.x * MERGE_TW, ty0 = blockIdx.y * MERGE_TH, b = blockIdx.z;
  const int t = threadIdx.x;
  const int xs = tx0 + (t & 15) * 4, y = ty0 + (t >> 4);
  const bool in_img = xs < W && y < H;   // W % 4 == 0: a strip is inside or outside as a whole
  const size_t plane = static_cast<size_t>(lh) * lw;
  float acc[4][NC_MAX];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c) acc[j][c] = 0.f;
  int count[4] = {0, 0, 0, 0};
  auto src = [](int cpix) { float v = 0.25f * (cpix + 0.5f) - 0.5f; return v < 0.f ? 0.f : v; };   // align_corners=False, x4
  for (int k = 0; k < n_crops; ++k) {
    const int by = __ldg(&boxes[k].x), bx = __ldg(&boxes[k].y);
    // tile rect intersected with the window (CTA-uniform)
    const int cy_lo = max(ty0 - by, 0), cy_hi = min(ty0 + MERGE_TH - 1 - by, crop_h - 1);
    const int cx_lo = max(tx0 - bx, 0), cx_hi = min(tx0 + MERGE_TW - 1 - bx, crop_w - 1);
    if (cy_lo > cy_hi || cx_lo > cx_hi) continue;
    const int ly_lo = static_cast<int>(src(cy_lo)), ly_hi = min(static_cast<int>(src(cy_hi)) + 1, lh - 1);
    const int lx_lo = static_cast<int>(src(cx_lo)), lx_hi = min(static_cast<int>(src(cx_hi)) + 1, lw - 1);
    const int fr = ly_hi - ly_lo + 1, fc = lx_hi - lx_lo + 1;   // <= 6 x 18
    __syncthreads();   // the previous window's footprint has been consumed
    const float* base = lowres + (static_cast<size_t>(b) * n_crops + k) * nc * plane + static_cast<size_t>(ly_lo) * lw + lx_lo;
    for (int i = t; i < nc * fr * fc; i += 256) {
      const int c = i / (fr * fc), rem = i - c * (fr * fc);
      const int r = rem / fc, col = rem - r * fc;
      fp[(c * MERGE_FR + r) * MERGE_FC + col] = __ldg(base + c * plane + static_cast<size_t>(r) * lw + col);
    }
    __syncthreads();
    const int cy = y - by;
    if (!in_img || cy < 0 || cy >= crop_h) continue;   // (no barrier below this point in the iteration)
    const float sy = src(cy);
    const int y0 = static_cast<int>(sy);
    const int yp = (y0 < lh - 1) ? MERGE_FC : 0;
    const float h1 = sy - y0, h0 = 1.f - h1;
    const float* row = fp + (y0 - ly_lo) * MERGE_FC - lx_lo;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cx = xs + j - bx;
      if (cx < 0 || cx >= crop_w) continue;
      ++count[j];
      const float sx = src(cx);
      const int x0 = static_cast<int>(sx);
      const int xp = (x0 < lw - 1) ? 1 : 0;
      const float w1 = sx - x0, w0 = 1.f - w1;
      const float* q = row + x0;
#pragma unroll
      for (int c = 0; c < NC_MAX; ++c) {
        if (c < nc) {
          const float* qc = q + c * (MERGE_FR * MERGE_FC);
          acc[j][c] = __fadd_rn(acc[j][c], bilerp_rn(h0, h1, w0, w1, qc[0], qc[xp], qc[yp], qc[yp + xp]));
        }
      }
    }
  }
  if (!in_img) return;
  if (flip_a != nullptr) {
    // Second pass of the horizontal-flip test-time augmentation (hrda_encoder_decoder.py:196-229, scales = [1]): the windows
    // were computed on the mirrored image, so this strip is b = slide(flip(img)) at columns xs .. xs + 3 and belongs to output
    // columns W - 1 - xs - j. flip_a = slide(img) at full resolution; result = (a + flip(b)) / 2 with the arithmetic of
    // tta_flip_mean_argmax_kernel (one rounded add, exact halving), first maximum wins. logits_out may alias flip_a: every
    // element is read and written by the same thread.
    const size_t pixm = static_cast<size_t>(y) * W + (W - 4 - xs);
    float best[4];
    int arg[4] = {0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c) {
      if (c < nc) {
        const size_t off = (static_cast<size_t>(b) * nc + c) * H * W + pixm;
        const float4 av = *reinterpret_cast<const float4*>(flip_a + off);
        const float a4[4] = {av.x, av.y, av.z, av.w};
        float m[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {   // output column W - 4 - xs + i  <-  strip pixel j = 3 - i
          const float bv = acc[3 - i][c] / static_cast<float>(count[3 - i]);
          m[i] = __fadd_rn(a4[i], bv) * 0.5f;
          if (c == 0 || m[i] > best[i]) { best[i] = m[i]; arg[i] = c; }
        }
        if (logits_out) *reinterpret_cast<float4*>(logits_out + off) = make_float4(m[0], m[1], m[2], m[3]);
      }
    }
    *reinterpret_cast<uint32_t*>(labels + static_cast<size_t>(b) * H * W + pixm) =
        static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1]) << 8) | (static_cast<uint32_t>(arg[2]) << 16) | (static_cast<uint32_t>(arg[3]) << 24);
    return;
  }
  const size_t pix = static_cast<size_t>(y) * W + xs;
  uint32_t lab = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float cnt = static_cast<float>(count[j]);  // >= 1: the grid covers the image
    int best = 0;
    acc[j][0] = acc[j][0] / cnt;
    float bestv = acc[j][0];
#pragma unroll
    for (int c = 1; c < NC_MAX; ++c) {
      if (c < nc) {
        acc[j][c] = acc[j][c] / cnt;
        if (acc[j][c] > bestv) { bestv = acc[j][c]; best = c; }
      }
    }
    lab |= static_cast<uint32_t>(best) << (8 * j);
  }
  *reinterpret_cast<uint32_t*>(labels + static_cast<size_t>(b) * H * W + pix) = lab;
  if (logits_out) {
    float* lo = logits_out + static_cast<size_t>(b) * nc * H * W + pix;
#pragma unroll
    for (int c = 0; c < NC_MAX; ++c)
      if (c < nc) *reinterpret_cast<float4*>(lo + static_cast<size_t>(c) * H * W) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
  }
}

// Same result again, one class at a time (round 2; VERDICT r1 "What's weak": the tile kernel above issues 6.8 k instructions per
// warp at 124 registers, 2 CTAs per SM). What changed:
//  * the footprints of ALL windows that overlap the tile (<= MERGE_MAXW, in window order; found with one ballot over the window
//    list) are staged first, every load in flight at once, behind ONE barrier; border taps are replicated into the footprint
//    (source row / column index clamped while staging), so the resampling code never tests for the last low-res row or column;
//  * the class loop is outermost: four accumulators per thread instead of 76; the division by the window count, the argmax
//    update, the optional logit store and the flip-TTA combination happen per class;
//  * the exact x4 upsampling makes the horizontal geometry integer: for a window pixel cx >= 2, x0 = (cx - 2) >> 2 and
//    w1 = 0.125 + 0.25 * ((cx - 2) & 3) (the values the float expression of the kernels above produces, exactly), and the phase
//    (xs - bx - 2) & 3 of a strip is the same for every thread of the CTA (xs % 4 == 0). A strip that lies inside the window
//    with cx >= 2 takes its 4-6 taps per class from shared memory once and evaluates the four pixels with immediate weights;
//    strips on a window's left / right border use the per-pixel expression, out of line;
//  * a warp owns two strips x 16 rows (not 16 strips x 2 rows): a window's left / right border then crosses one warp of the
//    tile instead of all eight (the first version of this kernel spent 12 % of its instructions in the border path);
//  * count in {1, 2, 4, 8, ...}: multiplying by the exact reciprocal equals the IEEE division, other counts divide.
// Every pixel evaluates the same expression in the same window order as slide_merge_argmax_kernel: bit-identical results
// (tests/test_ops_gpu.py compares the kernels). A tile overlapped by more than MERGE_MAXW windows (stride < crop / 2) gathers
// its taps per pixel from global memory inside the same class loop.
constexpr int MERGE_MAXW = 4;
constexpr int MERGE_SLOT = MERGE_FR * MERGE_FC;   // floats per class and window footprint

__device__ __forceinline__ float merge_lds(uint32_t addr) {   // 32-bit shared address: no generic-pointer arithmetic per tap
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

template <int P>
__device__ __forceinline__ void merge_strip_phase(uint32_t q, float h0, float h1, float (&acc)[4]) {
  const float a0 = merge_lds(q), a1 = merge_lds(q + 4), b0 = merge_lds(q + 4 * MERGE_FC), b1 = merge_lds(q + 4 * MERGE_FC + 4);
  float a2 = 0.f, b2 = 0.f;
  if (P > 0) { a2 = merge_lds(q + 8); b2 = merge_lds(q + 4 * MERGE_FC + 8); }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ph = (P + j) & 3, o = (P + j) >> 2;
    const float w1 = 0.125f + 0.25f * ph, w0 = 1.f - w1;   // exact: 0.125 / 0.375 / 0.625 / 0.875
    acc[j] = __fadd_rn(acc[j], bilerp_rn(h0, h1, w0, w1, o ? a1 : a0, o ? a2 : a1, o ? b1 : b0, o ? b2 : b1));
  }
}

// A strip on a window's left / right border (some of its pixels outside the window, or cx < 2 where the source coordinate
// clamps to 0): the per-pixel expression of slide_merge_argmax_kernel on the staged footprint (q = shared address of the
// strip's tap row at low-res column 0). Kept out of line so that its per-pixel geometry is not hoisted out of the class
// loop into registers.
__device__ __noinline__ float4 merge_strip_border(uint32_t q, int cx0, int crop_w, float h1, float4 acc) {
  const float h0 = 1.f - h1;
  float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cx = cx0 + j;
    if (cx < 0 || cx >= crop_w) continue;
    float sx = 0.25f * (cx + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
    const int x0 = static_cast<int>(sx);
    const float w1 = sx - x0, w0 = 1.f - w1;
    const uint32_t r = q + 4 * x0;   // taps beyond the last column / row are replicated in the footprint
    a[j] = __fadd_rn(a[j], bilerp_rn(h0, h1, w0, w1, merge_lds(r), merge_lds(r + 4), merge_lds(r + 4 * MERGE_FC), merge_lds(r + 4 * MERGE_FC + 4)));
  }
  return make_float4(a[0], a[1], a[2], a[3]);
}

// More windows over a tile than footprint slots: one class of one strip from global memory, windows in order.
__device__ __noinline__ float4 merge_strip_gather(const float* __restrict__ cls_plane, size_t win_stride, const int2* __restrict__ boxes,
                                                  int n_crops, int crop_h, int crop_w, int lh, int lw, int xs, int y) {
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < n_crops; ++k) {
    const int cy = y - __ldg(&boxes[k].x), cx0 = xs - __ldg(&boxes[k].y);
    if (cy < 0 || cy >= crop_h || cx0 + 3 < 0 || cx0 >= crop_w) continue;
    float sy = 0.25f * (cy + 0.5f) - 0.5f; sy = sy < 0.f ? 0.f : sy;
    const int y0 = static_cast<int>(sy);
    const int yp = (y0 < lh - 1) ? lw : 0;
    const float h1 = sy - y0, h0 = 1.f - h1;
    const float* p = cls_plane + k * win_stride + static_cast<size_t>(y0) * lw;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cx = cx0 + j;
      if (cx < 0 || cx >= crop_w) continue;
      float sx = 0.25f * (cx + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
      const int x0 = static_cast<int>(sx);
      const int xp = (x0 < lw - 1) ? 1 : 0;
      const float w1 = sx - x0, w0 = 1.f - w1;
      const float* r = p + x0;
      a[j] = __fadd_rn(a[j], bilerp_rn(h0, h1, w0, w1, __ldg(r), __ldg(r + xp), __ldg(r + yp), __ldg(r + yp + xp)));
    }
  }
  return make_float4(a[0], a[1], a[2], a[3]);
}

template <int NC_MAX, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
slide_merge_class_kernel(const float* __restrict__ lowres, const int2* __restrict__ boxes, int n_crops, int nc,
                         int crop_h, int crop_w, int lh, int lw, int H, int W,
                         uint8_t* __restrict__ labels, float* logits_out, const float* flip_a) {
  __shared__ float fp[MERGE_MAXW * NC_MAX * MERGE_SLOT];
  const int tx0 = blockIdx.x * MERGE_TW, ty0 = blockIdx.y * MERGE_TH, b = blockIdx.z;
  const int t = threadIdx.x, lane = t & 31;
  const int xs = tx0 + ((t >> 5) * 2 + (lane >> 4)) * 4, y = ty0 + (lane & 15);   // warp = 2 strips x 16 rows
  const bool in_img = xs < W && y < H;   // W % 4 == 0: a strip is inside or outside as a whole
  const int iplane = lh * lw;
  auto src = [](int cpix) { float v = 0.25f * (cpix + 0.5f) - 0.5f; return v < 0.f ? 0.f : v; };   // align_corners=False, x4
  const uint32_t fp_addr = static_cast<uint32_t>(__cvta_generic_to_shared(fp));

  // per-thread state of the (at most MERGE_MAXW) windows over this tile. off: shared address of the strip's first tap of class 0;
  // mode: -1 strip outside the window (or slot unused), 0..3 strip inside with cx >= 2 (value = phase), 4 strip on the window's
  // left / right border
  uint32_t off[MERGE_MAXW];
  int mode[MERGE_MAXW], wcx0[MERGE_MAXW];
  float wh1[MERGE_MAXW];
#pragma unroll
  for (int s = 0; s < MERGE_MAXW; ++s) { off[s] = 0; mode[s] = -1; wcx0[s] = 0; wh1[s] = 0.f; }
  int count[4] = {0, 0, 0, 0};
  int n_ov = 0;
  for (int k0 = 0; k0 < n_crops; k0 += 32) {
    // one window per lane: does it overlap the tile? (CTA-uniform result: every warp computes the same mask)
    int by_l = 0, bx_l = 0;
    bool ov = false;
    if (k0 + lane < n_crops) {
      by_l = __ldg(&boxes[k0 + lane].x); bx_l = __ldg(&boxes[k0 + lane].y);
      ov = max(ty0 - by_l, 0) <= min(ty0 + MERGE_TH - 1 - by_l, crop_h - 1) && max(tx0 - bx_l, 0) <= min(tx0 + MERGE_TW - 1 - bx_l, crop_w - 1);
    }
    unsigned m = __ballot_sync(0xffffffffu, ov);
    while (m) {   // overlapping windows in window order
      const int kl = __ffs(m) - 1;
      m &= m - 1;
      const int k = k0 + kl;
      const int by = __shfl_sync(0xffffffffu, by_l, kl), bx = __shfl_sync(0xffffffffu, bx_l, kl);
      const int cy = y - by, cx0 = xs - bx;
      const bool rowok = in_img && cy >= 0 && cy < crop_h;
#pragma unroll
      for (int j = 0; j < 4; ++j) count[j] += (rowok && cx0 + j >= 0 && cx0 + j < crop_w) ? 1 : 0;
      if (n_ov < MERGE_MAXW) {
        // tile rect intersected with the window and its low-res footprint (CTA-uniform)
        const int cy_lo = max(ty0 - by, 0), cy_hi = min(ty0 + MERGE_TH - 1 - by, crop_h - 1);
        const int cx_lo = max(tx0 - bx, 0), cx_hi = min(tx0 + MERGE_TW - 1 - bx, crop_w - 1);
        const int ly_lo = static_cast<int>(src(cy_lo)), fr = static_cast<int>(src(cy_hi)) + 2 - ly_lo;   // <= 6 rows
        const int lx_lo = static_cast<int>(src(cx_lo)), fc = static_cast<int>(src(cx_hi)) + 2 - lx_lo;   // <= 18 columns
        const int e = t & 127;
        if (e < fr * fc) {   // two thread groups take the even / odd classes of one footprint element each
          const int r = e / fc, col = e - r * fc;
          const float* g = lowres + (static_cast<size_t>(b) * n_crops + k) * nc * iplane + (min(ly_lo + r, lh - 1) * lw + min(lx_lo + col, lw - 1));
          float* d = fp + n_ov * (NC_MAX * MERGE_SLOT) + r * MERGE_FC + col;
#pragma unroll
          for (int c = 0; c < NC_MAX; c += 2) {
            const int cc = c + (t >> 7);
            if (cc < nc) d[cc * MERGE_SLOT] = __ldg(g + cc * iplane);
          }
        }
        const bool anyin = rowok && cx0 + 3 >= 0 && cx0 < crop_w;
        const bool fast = anyin && cx0 >= 2 && cx0 + 3 < crop_w;
        const float sy = src(cy);
        const int y0 = static_cast<int>(sy);
        const uint32_t o = fp_addr + 4u * static_cast<uint32_t>(n_ov * (NC_MAX * MERGE_SLOT) + (y0 - ly_lo) * MERGE_FC - lx_lo + (fast ? ((cx0 - 2) >> 2) : 0));
        const int md = !anyin ? -1 : (fast ? ((cx0 - 2) & 3) : 4);
#pragma unroll
        for (int s = 0; s < MERGE_MAXW; ++s)
          if (s == n_ov) { off[s] = o; mode[s] = md; wcx0[s] = cx0; wh1[s] = sy - y0; }
      }
      ++n_ov;
    }
  }
  __syncthreads();
  if (!in_img) return;

  bool pow2 = true;
  float inv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {   // count >= 1: the grid covers the image (count_mat assert in the reference)
    pow2 = pow2 && count[j] > 0 && (count[j] & (count[j] - 1)) == 0;
    inv[j] = __int_as_float((128 - __ffs(count[j])) << 23);   // 2^-k for count = 2^k
  }
  const size_t plane_out = static_cast<size_t>(H) * W;
  const size_t pix = static_cast<size_t>(y) * W + (flip_a != nullptr ? (W - 4 - xs) : xs);
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  int arg[4] = {0, 0, 0, 0};
  // division by the window count, argmax update (strict >: the first maximum wins), optional logit store of one class
  auto finish_class = [&](int c, float (&acc)[4]) {
    if (pow2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] *= inv[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = acc[j] / static_cast<float>(count[j]);
    }
    const size_t o = (static_cast<size_t>(b) * nc + c) * plane_out + pix;
    if (flip_a != nullptr) {
      // second pass of the horizontal-flip test-time augmentation: see slide_merge_tile_kernel. Output column W - 4 - xs + i
      // <- strip pixel 3 - i; (a + flip(b)) / 2 with one rounded add and an exact halving; logits_out may alias flip_a.
      const float4 av = *reinterpret_cast<const float4*>(flip_a + o);
      float m[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        m[i] = __fadd_rn(m[i], acc[3 - i]) * 0.5f;
        if (c == 0 || m[i] > best[i]) { best[i] = m[i]; arg[i] = c; }
      }
      if (logits_out) *reinterpret_cast<float4*>(logits_out + o) = make_float4(m[0], m[1], m[2], m[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c == 0 || acc[j] > best[j]) { best[j] = acc[j]; arg[j] = c; }
      if (logits_out) *reinterpret_cast<float4*>(logits_out + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
  };
  if (n_ov <= MERGE_MAXW) {
    for (int c = 0; c < nc; ++c) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t coff = static_cast<uint32_t>(c) * (4u * MERGE_SLOT);
#pragma unroll
      for (int s = 0; s < MERGE_MAXW; ++s) {
        const int md = mode[s];
        if (md < 0) continue;
        const uint32_t q = off[s] + coff;
        const float h1 = wh1[s], h0 = 1.f - h1;
        if (md < 2) {
          if (md == 0) merge_strip_phase<0>(q, h0, h1, acc);
          else merge_strip_phase<1>(q, h0, h1, acc);
        } else if (md < 4) {
          if (md == 2) merge_strip_phase<2>(q, h0, h1, acc);
          else merge_strip_phase<3>(q, h0, h1, acc);
        } else {   // window border
          const float4 r = merge_strip_border(q, wcx0[s], crop_w, h1, make_float4(acc[0], acc[1], acc[2], acc[3]));
          acc[0] = r.x; acc[1] = r.y; acc[2] = r.z; acc[3] = r.w;
        }
      }
      finish_class(c, acc);
    }
  } else {   // more windows over this tile than footprint slots
    for (int c = 0; c < nc; ++c) {
      const float4 r = merge_strip_gather(lowres + (static_cast<size_t>(b) * n_crops * nc + c) * iplane, static_cast<size_t>(nc) * iplane,
                                          boxes, n_crops, crop_h, crop_w, lh, lw, xs, y);
      float acc[4] = {r.x, r.y, r.z, r.w};
      finish_class(c, acc);
    }
  }
  *reinterpret_cast<uint32_t*>(labels + static_cast<size_t>(b) * plane_out + pix) =
      static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1]) << 8) | (static_cast<uint32_t>(arg[2]) << 16) | (static_cast<uint32_t>(arg[3]) << 24);
}

// ---------------------------------------------------------------------------------------------
// Stage-1 merge of ms_inference (Ms_VFM_encoder_decoder.py:449-461) + argmax in the class-major tile form of
// slide_merge_class_kernel: same results, bit for bit, as ms_merge_argmax_kernel (ms_refine.cuh; tests compare the two). Sources
// per 64 x 16 output tile: the coarse logits low0 (context value U(y, x), upsampled by S0 = H / lh) and, for every REFINED window
// over the tile, its aux-decoder logits (upsampled by Sr = crop / rh); windows that were confident enough to be skipped add the
// context value. Both upsampling factors must be powers of two >= 4 (8 and 16 in the shipped config): for a source pixel
// coordinate c >= S / 2, x0 = (c - S / 2) >> log2 S and w1 = (((c - S / 2) & (S - 1)) + 0.5) / S exactly as the float expression
// of bilerp_setup, a strip of four pixels touches at most three source columns, and w1 of pixel j is w1 of pixel 0 plus j / S
// (minus one past the column step) without rounding. Per class a strip costs 6 shared loads per source instead of 16 L2-latency
// loads per pixel and source, with 4 accumulators per thread instead of 19 + 19.
struct MsStrip {        // per-thread view of one source (context or one refined window) for this thread's strip
  uint32_t off;         // shared address of the tap (row y0, column x0 of pixel 0 | footprint column 0 for border strips), class 0
  int meta;             // bits 0-2: 0 = border strip (per-pixel path), 1..4 = index of the first pixel one source column further (4: none);
                        // bits 4-7: pixel j lies inside the window; bit 8: source used by this strip; bit 9: window refined;
                        // bits 16-31: source-local x of pixel 0 plus 8 (border strips)
  float w1_0, h1;
};

// Out of line on purpose: inlined five times (context + four window slots) the compiler hoists every slot's per-pixel weights and
// select masks out of the class loop and spills them (192 bytes of stack at 128 registers).
__device__ __noinline__ float4 ms_strip_eval(uint32_t q, int kx, float w1_0, float inv_s, float h1) {
  const float h0 = 1.f - h1;
  const float a0 = merge_lds(q), a1 = merge_lds(q + 4), a2 = merge_lds(q + 8);
  const float b0 = merge_lds(q + 4 * MERGE_FC), b1 = merge_lds(q + 4 * MERGE_FC + 4), b2 = merge_lds(q + 4 * MERGE_FC + 8);
  float v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const bool o = j >= kx;
    const float w1 = __fadd_rn(w1_0, static_cast<float>(j) * inv_s) - (o ? 1.f : 0.f);   // exact
    const float w0 = 1.f - w1;
    v[j] = bilerp_rn(h0, h1, w0, w1, o ? a1 : a0, o ? a2 : a1, o ? b1 : b0, o ? b2 : b1);
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

// Border strip of a source (pixels outside [0, limit), or c < S / 2 where the source coordinate clamps to 0): bilerp_setup's float
// expression per pixel on the staged footprint (q = shared address of footprint row y0, source column 0). Returns the four
// values; pixels outside the range return 0 (the caller masks them).
__device__ __noinline__ float4 ms_strip_border(uint32_t q, int c0, int limit, float inv_s, float h1) {
  const float h0 = 1.f - h1;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = c0 + j;
    if (c < 0 || c >= limit) continue;
    float sx = inv_s * (c + 0.5f) - 0.5f; sx = sx < 0.f ? 0.f : sx;
    const int x0 = static_cast<int>(sx);
    const float w1 = sx - x0, w0 = 1.f - w1;
    const uint32_t r = q + 4 * x0;   // taps beyond the last column / row are replicated in the footprint
    v[j] = bilerp_rn(h0, h1, w0, w1, merge_lds(r), merge_lds(r + 4), merge_lds(r + 4 * MERGE_FC), merge_lds(r + 4 * MERGE_FC + 4));
  }
  return make_float4(v[0], v[1], v[2], v[3]);
}

// More windows over a tile than footprint slots: one class of one strip from global memory, the loop of ms_merge_argmax_kernel.
__device__ __noinline__ float4 ms_strip_gather(const float* __restrict__ low0_c, const float* __restrict__ refined_c, size_t ref_stride,
                                               const int* __restrict__ ref_index, const int2* __restrict__ boxes, int n_crops,
                                               int crop_h, int crop_w, int lh, int lw, int rh, int rw, int H, int W, int xs, int y) {
  const float up_h = static_cast<float>(lh) / H, up_w = static_cast<float>(lw) / W;
  const float r_h = static_cast<float>(rh) / crop_h, r_w = static_cast<float>(rw) / crop_w;
  float a[4] = {0.f, 0.f, 0.f, 0.f}, ctx[4] = {0.f, 0.f, 0.f, 0.f};
  bool have_ctx = false;
  for (int k = 0; k < n_crops; ++k) {
    const int cy = y - __ldg(&boxes[k].x), cx0 = xs - __ldg(&boxes[k].y);
    if (cy < 0 || cy >= crop_h || cx0 + 3 < 0 || cx0 >= crop_w) continue;
    const int ri = __ldg(ref_index + k);
    if (ri < 0 && !have_ctx) {
#pragma unroll
      for (int j = 0; j < 4; ++j) ctx[j] = bilerp_eval(low0_c, bilerp_setup(y, xs + j, up_h, up_w, lh, lw));
      have_ctx = true;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cx = cx0 + j;
      if (cx < 0 || cx >= crop_w) continue;
      const float v = ri >= 0 ? bilerp_eval(refined_c + ri * ref_stride, bilerp_setup(cy, cx, r_h, r_w, rh, rw)) : ctx[j];
      a[j] = __fadd_rn(a[j], v);
    }
  }
  return make_float4(a[0], a[1], a[2], a[3]);
}

template <int NC_MAX, int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS)
ms_merge_class_kernel(const float* __restrict__ low0, const float* __restrict__ refined, const int* __restrict__ ref_index,
                      const int2* __restrict__ boxes, int n_crops, int nc, int crop_h, int crop_w, int lh, int lw, int rh, int rw,
                      int H, int W, int lg0, int lgr, uint8_t* __restrict__ labels, float* __restrict__ logits_out) {
  __shared__ float fp[(MERGE_MAXW + 1) * NC_MAX * MERGE_SLOT];   // slot 0: context footprint, 1..4: refined windows
  const int tx0 = blockIdx.x * MERGE_TW, ty0 = blockIdx.y * MERGE_TH, b = blockIdx.z;
  const int t = threadIdx.x, lane = t & 31;
  const int xs = tx0 + ((t >> 5) * 2 + (lane >> 4)) * 4, y = ty0 + (lane & 15);   // warp = 2 strips x 16 rows
  const bool in_img = xs < W && y < H;   // W % 4 == 0
  const int S0 = 1 << lg0, Sr = 1 << lgr;
  const float inv_s0 = 1.f / static_cast<float>(S0), inv_sr = 1.f / static_cast<float>(Sr);   // = lw / W, rw / crop_w (exact)
  const uint32_t fp_addr = static_cast<uint32_t>(__cvta_generic_to_shared(fp));
  auto src = [](int cpix, float inv_s) { float v = inv_s * (cpix + 0.5f) - 0.5f; return v < 0.f ? 0.f : v; };
  // footprint of a source over output pixels [c_lo, c_hi] x [r_lo, r_hi] (source-local coordinates), staged with all loads in flight
  auto stage = [&](int slot, const float* __restrict__ base, int plane, int ih, int iw, int r_lo, int r_hi, int c_lo, int c_hi, float inv_s,
                   int& ly_lo, int& lx_lo) {
    ly_lo = static_cast<int>(src(r_lo, inv_s));
    lx_lo = static_cast<int>(src(c_lo, inv_s));
    const int fr = static_cast<int>(src(r_hi, inv_s)) + 2 - ly_lo, fc = static_cast<int>(src(c_hi, inv_s)) + 2 - lx_lo;   // <= 6 x 18 for S >= 4
    const int e = t & 127;
    if (e < fr * fc) {   // two thread groups take the even / odd classes of one footprint element each
      const int r = e / fc, col = e - r * fc;
      const float* g = base + (min(ly_lo + r, ih - 1) * iw + min(lx_lo + col, iw - 1));
      float* d = fp + slot * (NC_MAX * MERGE_SLOT) + r * MERGE_FC + col;
#pragma unroll
      for (int c = 0; c < NC_MAX; c += 2) {
        const int cc = c + (t >> 7);
        if (cc < nc) d[cc * MERGE_SLOT] = __ldg(g + cc * plane);
      }
    }
  };
  // this thread's strip in a source: c0 = source-local x of pixel 0, [0, limit) the valid range, (ly_lo, lx_lo) the footprint origin
  auto strip = [&](int slot, int row, int c0, int limit, int lg, float inv_s, int ly_lo, int lx_lo, bool rowok, bool refined_w) {
    MsStrip st;
    const int S = 1 << lg;
    const bool anyin = rowok && c0 + 3 >= 0 && c0 < limit;
    const bool fast = anyin && c0 >= (S >> 1) && c0 + 3 < limit;
    const float sy = src(row, inv_s);
    const int y0 = static_cast<int>(sy);
    const int t0 = c0 - (S >> 1);
    st.off = fp_addr + 4u * static_cast<uint32_t>(slot * (NC_MAX * MERGE_SLOT) + (y0 - ly_lo) * MERGE_FC - lx_lo + (fast ? (t0 >> lg) : 0));
    const int ph0 = t0 & (S - 1);
    int inmask = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) inmask |= (rowok && c0 + j >= 0 && c0 + j < limit) ? (16 << j) : 0;
    st.meta = (fast ? min(S - ph0, 4) : 0) | inmask | (anyin ? 256 : 0) | (refined_w ? 512 : 0) | (anyin ? ((c0 + 8) << 16) : 0);
    st.w1_0 = (static_cast<float>(ph0) + 0.5f) * inv_s;
    st.h1 = sy - y0;
    return st;
  };

  // context footprint: the tile itself in image coordinates
  MsStrip cst;
  {
    int ly_lo, lx_lo;
    stage(0, low0 + static_cast<size_t>(b) * nc * lh * lw, lh * lw, lh, lw, ty0, min(ty0 + MERGE_TH - 1, H - 1), tx0, min(tx0 + MERGE_TW - 1, W - 1),
          inv_s0, ly_lo, lx_lo);
    cst = strip(0, y, xs, W, lg0, inv_s0, ly_lo, lx_lo, in_img, false);
  }
  MsStrip wst[MERGE_MAXW];
#pragma unroll
  for (int s = 0; s < MERGE_MAXW; ++s) { wst[s].off = 0; wst[s].meta = 0; wst[s].w1_0 = 0.f; wst[s].h1 = 0.f; }
  int count[4] = {0, 0, 0, 0};
  int n_ov = 0;
  for (int k0 = 0; k0 < n_crops; k0 += 32) {
    int by_l = 0, bx_l = 0, ri_l = -1;
    bool ov = false;
    if (k0 + lane < n_crops) {
      by_l = __ldg(&boxes[k0 + lane].x); bx_l = __ldg(&boxes[k0 + lane].y);
      ri_l = __ldg(ref_index + static_cast<size_t>(b) * n_crops + k0 + lane);
      ov = max(ty0 - by_l, 0) <= min(ty0 + MERGE_TH - 1 - by_l, crop_h - 1) && max(tx0 - bx_l, 0) <= min(tx0 + MERGE_TW - 1 - bx_l, crop_w - 1);
    }
    unsigned m = __ballot_sync(0xffffffffu, ov);
    while (m) {   // overlapping windows in window order
      const int kl = __ffs(m) - 1;
      m &= m - 1;
      const int by = __shfl_sync(0xffffffffu, by_l, kl), bx = __shfl_sync(0xffffffffu, bx_l, kl), ri = __shfl_sync(0xffffffffu, ri_l, kl);
      const int cy = y - by, cx0 = xs - bx;
      const bool rowok = in_img && cy >= 0 && cy < crop_h;
#pragma unroll
      for (int j = 0; j < 4; ++j) count[j] += (rowok && cx0 + j >= 0 && cx0 + j < crop_w) ? 1 : 0;
      if (n_ov < MERGE_MAXW) {
        int ly_lo = 0, lx_lo = 0;
        if (ri >= 0)
          stage(1 + n_ov, refined + static_cast<size_t>(ri) * nc * rh * rw, rh * rw, rh, rw, max(ty0 - by, 0), min(ty0 + MERGE_TH - 1 - by, crop_h - 1),
                max(tx0 - bx, 0), min(tx0 + MERGE_TW - 1 - bx, crop_w - 1), inv_sr, ly_lo, lx_lo);
        const MsStrip st = strip(1 + n_ov, cy, cx0, crop_w, lgr, inv_sr, ly_lo, lx_lo, rowok, ri >= 0);
#pragma unroll
        for (int s = 0; s < MERGE_MAXW; ++s)
          if (s == n_ov) wst[s] = st;
      }
      ++n_ov;
    }
  }
  __syncthreads();
  if (!in_img) return;

  bool pow2 = true, need_ctx = false;
  float inv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {   // count >= 1: the grid covers the image (count_mat assert in the reference)
    pow2 = pow2 && count[j] > 0 && (count[j] & (count[j] - 1)) == 0;
    inv[j] = __int_as_float((128 - __ffs(count[j])) << 23);   // 2^-k for count = 2^k
  }
#pragma unroll
  for (int s = 0; s < MERGE_MAXW; ++s) need_ctx = need_ctx || ((wst[s].meta & 256) && !(wst[s].meta & 512));
  const size_t plane_out = static_cast<size_t>(H) * W;
  const size_t pix = static_cast<size_t>(y) * W + xs;
  float best[4] = {0.f, 0.f, 0.f, 0.f};
  int arg[4] = {0, 0, 0, 0};
  for (int c = 0; c < nc; ++c) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (n_ov <= MERGE_MAXW) {
      const uint32_t coff = static_cast<uint32_t>(c) * (4u * MERGE_SLOT);
      float ctx[4] = {0.f, 0.f, 0.f, 0.f};
      if (need_ctx) {   // U(y, xs + j): the context value of the strip, evaluated once per class
        const float4 r = (cst.meta & 7) ? ms_strip_eval(cst.off + coff, cst.meta & 7, cst.w1_0, inv_s0, cst.h1)
                                        : ms_strip_border(cst.off + coff, xs, W, inv_s0, cst.h1);
        ctx[0] = r.x; ctx[1] = r.y; ctx[2] = r.z; ctx[3] = r.w;
      }
#pragma unroll
      for (int s = 0; s < MERGE_MAXW; ++s) {
        const int meta = wst[s].meta;
        if (!(meta & 256)) continue;
        float v[4];
        if (!(meta & 512)) {   // window skipped by the confidence gate: its logits are the context
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = ctx[j];
        } else {
          const float4 r = (meta & 7) ? ms_strip_eval(wst[s].off + coff, meta & 7, wst[s].w1_0, inv_sr, wst[s].h1)
                                      : ms_strip_border(wst[s].off + coff, (meta >> 16) - 8, crop_w, inv_sr, wst[s].h1);
          v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (meta & (16 << j)) acc[j] = __fadd_rn(acc[j], v[j]);
      }
    } else {   // more windows over this tile than footprint slots
      const float4 r = ms_strip_gather(low0 + (static_cast<size_t>(b) * nc + c) * lh * lw, refined ? refined + static_cast<size_t>(c) * rh * rw : nullptr,
                                       static_cast<size_t>(nc) * rh * rw, ref_index + static_cast<size_t>(b) * n_crops, boxes, n_crops, crop_h, crop_w,
                                       lh, lw, rh, rw, H, W, xs, y);
      acc[0] = r.x; acc[1] = r.y; acc[2] = r.z; acc[3] = r.w;
    }
    if (pow2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] *= inv[j];
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = acc[j] / static_cast<float>(count[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (c == 0 || acc[j] > best[j]) { best[j] = acc[j]; arg[j] = c; }   // strict >: the first maximum wins
    if (logits_out)
      *reinterpret_cast<float4*>(logits_out + (static_cast<size_t>(b) * nc + c) * plane_out + pix) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  *reinterpret_cast<uint32_t*>(labels + static_cast<size_t>(b) * plane_out + pix) =
      static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1]) << 8) | (static_cast<uint32_t>(arg[2]) << 16) | (static_cast<uint32_t>(arg[3]) << 24);
}

// ---------------------------------------------------------------------------------------------
// Confusion matrix cm[label, pred] ((nc + 1) x nc, int64) over non-ignored pixels. Integer restatement of mmseg
// IoUMetric.intersect_and_union as called from rein/dg_metrics.py:50-52 (three float32 histc
// calls): area_intersect = diag(cm), area_pred = cm.sum(0), area_label = cm.sum(1).
// Per-CTA shared int32 histogram, per-thread run-length aggregation (segmentation maps are
// piecewise constant, so most of a thread's 16 pixels share one bin), then one int64 global
// atomic per non-empty bin per CTA.
__global__ void __launch_bounds__(256)
confusion_matrix_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ label, long long n, int nc,
                        int ignore_index, unsigned long long* __restrict__ cm) {
  extern __shared__ int s_hist[];
  const int bins = (nc + 1) * nc;  // extra row nc: label outside [0, nc) but not ignored (still counts in area_pred)
  for (int i = threadIdx.x; i < bins; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  const long long nvec = n >> 4;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint4 pv = __ldg(reinterpret_cast<const uint4*>(pred) + v);
    const uint4 lv = __ldg(reinterpret_cast<const uint4*>(label) + v);
    const uint8_t* pb = reinterpret_cast<const uint8_t*>(&pv);
    const uint8_t* lb = reinterpret_cast<const uint8_t*>(&lv);
    int run_key = -1, run_len = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int l = lb[i], p = pb[i];
      const int key = (l == ignore_index || p >= nc) ? -1 : (l < nc ? l : nc) * nc + p;
      if (key == run_key) { ++run_len; }
      else {
        if (run_key >= 0) atomicAdd(&s_hist[run_key], run_len);
        run_key = key; run_len = 1;
      }
    }
    if (run_key >= 0) atomicAdd(&s_hist[run_key], run_len);
  }
  // scalar tail (n not a multiple of 16), first CTA only
  if (blockIdx.x == 0) {
    for (long long i = (nvec << 4) + threadIdx.x; i < n; i += blockDim.x) {
      const int l = label[i], p = pred[i];
      if (l != ignore_index && p < nc) atomicAdd(&s_hist[(l < nc ? l : nc) * nc + p], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < bins; i += blockDim.x) {
    const int c = s_hist[i];
    if (c) atomicAdd(&cm[i], static_cast<unsigned long long>(c));
  }
}

// ---------------------------------------------------------------------------------------------
// tta_flip_mean_argmax: horizontal-flip test-time augmentation, the combining step of
// rein/models/segmentors/hrda_encoder_decoder.py:196-229 with scales = [1]:
//   res = 0; res += slide(img); res += flip(slide(flip(img)), [3]); res / 2   -> argmax(dim=0), first maximum wins.
// a = slide(img), b = slide(flip(img)) (NOT yet flipped back), both fp32 [n_img, nc, H, W]. 0 + a is exact and / 2 is an
// exact halving, so (a + b_mirrored) * 0.5f is bit-identical to the reference's three statements. HBM-bound: two reads and
// at most one write of the logits; a thread owns VEC consecutive pixels of one row for every class (16-byte accesses when
// VEC = 4; the mirrored read is the same 128-byte lines in reverse lane order, so it coalesces as well).
template <int VEC>
__global__ void __launch_bounds__(256)
tta_flip_mean_argmax_kernel(const float* a, const float* __restrict__ b, int n_img, int nc, int H, int W,
                            uint8_t* __restrict__ labels, float* logits_out /* may alias a */) {
  constexpr int CH = 4;   // classes whose loads are issued together (8 x 16 B in flight per thread)
  const int Wv = W / VEC;
  const long long total = static_cast<long long>(n_img) * H * Wv;
  const size_t plane = static_cast<size_t>(H) * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int xv = static_cast<int>(idx % Wv);
    const long long r = idx / Wv;
    const int y = static_cast<int>(r % H);
    const int img = static_cast<int>(r / H);
    const int x = xv * VEC, xm = W - VEC - x;          // mirrored strip starts at W - 1 - (x + VEC - 1)
    const size_t row = static_cast<size_t>(img) * nc * plane + static_cast<size_t>(y) * W;
    float best[VEC];
    int arg[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) { best[i] = -INFINITY; arg[i] = 0; }
    for (int c0 = 0; c0 < nc; c0 += CH) {
      float va[CH][VEC], vb[CH][VEC];
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        if (c0 + k < nc) {
          const size_t off = row + static_cast<size_t>(c0 + k) * plane;
          if constexpr (VEC == 4) {
            const float4 fa = __ldcs(reinterpret_cast<const float4*>(a + off + x));
            const float4 fb = __ldcs(reinterpret_cast<const float4*>(b + off + xm));
            va[k][0] = fa.x; va[k][1] = fa.y; va[k][2] = fa.z; va[k][3] = fa.w;
            vb[k][0] = fb.w; vb[k][1] = fb.z; vb[k][2] = fb.y; vb[k][3] = fb.x;
          } else {
            va[k][0] = __ldcs(a + off + x); vb[k][0] = __ldcs(b + off + xm);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int c = c0 + k;
        if (c < nc) {
          const size_t off = row + static_cast<size_t>(c) * plane;
          float m[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            m[i] = __fadd_rn(va[k][i], vb[k][i]) * 0.5f;
            if (m[i] > best[i] || c == 0) { best[i] = m[i]; arg[i] = c; }   // strict >: the first maximum wins
          }
          if (logits_out != nullptr) {
            if constexpr (VEC == 4) __stcs(reinterpret_cast<float4*>(logits_out + off + x), make_float4(m[0], m[1], m[2], m[3]));
            else __stcs(logits_out + off + x, m[0]);
          }
        }
      }
    }
    uint8_t* lp = labels + static_cast<size_t>(img) * plane + static_cast<size_t>(y) * W + x;
    if constexpr (VEC == 4)
      *reinterpret_cast<uint32_t*>(lp) = static_cast<uint32_t>(arg[0]) | (static_cast<uint32_t>(arg[1]) << 8) |
                                         (static_cast<uint32_t>(arg[2]) << 16) | (static_cast<uint32_t>(arg[3]) << 24);
    else
      *lp = static_cast<uint8_t>(arg[0]);
  }
}

}  // namespace vfm
