// The memory-bound tail of slide inference: overlapped-window logit merge + argmax, and the
// confusion-matrix histogram behind mIoU.
#pragma once
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
          acc[c] += h0 * (w0 * __ldg(q) + w1 * __ldg(q + xp)) + h1 * (w0 * __ldg(q + yp) + w1 * __ldg(q + yp + xp));
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

}  // namespace vfm
