// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = A[M,K] * W[N,K]^T
//   A: activations, row-major bf16 (K contiguous)        -> TMA box {64, 128}, SWIZZLE_128B
//   W: nn.Linear / 1x1-conv weight, row-major bf16 [N,K] -> TMA box {64, BLOCK_N}
//   D: fp32 accumulators in TMEM (2 stages), drained by 8 epilogue warps through an
//      epilogue functor (bias / GELU / LayerScale+residual / pixel-shuffle / NCHW logits).
//
// Replaces, on the reference's hot path, every nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s=2):
//   rein/models/backbones/dino_layers/attention.py:51,67 (qkv, proj), mlp.py:26-28 (fc1, fc2),
//   patch_embed.py:66 (proj conv), rein/models/heads/linear_head.py:36-48 (fusion conv,
//   transposed convs, conv_seg).
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane), warp 2 = TMEM
// allocator, warp 3 idle, warps 4..11 = epilogue (lane quadrant = warp % 4, column half = (warp-4)/4).
#pragma once
#include "sm100_ptx.cuh"

namespace vfm {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;   // 64 bf16 = one 128-B swizzle span
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 384;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int kStages = (BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 6 : 8);
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : 2 * BLOCK_N;  // 2 accumulator stages
  static constexpr int kEpiSplit = (BLOCK_N >= 64) ? 2 : 1;                // column halves
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// ------------------------------------------------------------------ epilogues
// Each functor sees one accumulator row fragment: 32 consecutive columns [col0, col0+32) of
// global row `row` (row < M guaranteed by the caller; col0 + 32 <= padded N).

struct EpiBiasBf16 {            // out = bf16(acc + bias)            (qkv; fusion conv with bias=null)
  __nv_bfloat16* out; int ldo; const float* bias; int n_valid;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    uint32_t p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float b0 = bias ? __ldg(bias + col0 + 2 * i) : 0.f, b1 = bias ? __ldg(bias + col0 + 2 * i + 1) : 0.f;
      p[i] = pack_bf16x2(v[2 * i] + b0, v[2 * i + 1] + b1);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
  }
};

struct EpiBiasGeluBf16 {        // out = bf16(gelu_erf(acc + bias))  (fc1; mlp.py:35-36)
  __nv_bfloat16* out; int ldo; const float* bias;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    uint32_t p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = gelu_erf(v[2 * i] + __ldg(bias + col0 + 2 * i));
      float b = gelu_erf(v[2 * i + 1] + __ldg(bias + col0 + 2 * i + 1));
      p[i] = pack_bf16x2(a, b);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * ldo + col0);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
  }
};

// x += gamma * (acc + bias) on the fp32 residual stream (block.py:112-113 with LayerScale,
// layer_scale.py:27); optionally snapshots the new residual as a bf16 feature tap with the cls
// row dropped (dino_v2.py:261-267) into a token-major [n_crops*patches, tap_ld] buffer.
struct EpiResidual {
  float* x; int ldx; const float* bias; const float* gamma;
  __nv_bfloat16* tap; int tap_ld; int tap_col0; int tokens_per_crop;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    float4* xp = reinterpret_cast<float4*>(x + static_cast<size_t>(row) * ldx + col0);
    float r[32];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 o = xp[i];
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col0) + i);
      float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      o.x += g.x * (v[4 * i] + b.x);
      o.y += g.y * (v[4 * i + 1] + b.y);
      o.z += g.z * (v[4 * i + 2] + b.z);
      o.w += g.w * (v[4 * i + 3] + b.w);
      xp[i] = o;
      r[4 * i] = o.x; r[4 * i + 1] = o.y; r[4 * i + 2] = o.z; r[4 * i + 3] = o.w;
    }
    if (tap != nullptr) {
      int crop = row / tokens_per_crop, tok = row - crop * tokens_per_crop;
      if (tok != 0) {
        size_t trow = static_cast<size_t>(row - crop - 1);
        uint4* dst = reinterpret_cast<uint4*>(tap + trow * tap_ld + tap_col0 + col0);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_uint4(pack_bf16x2(r[8 * i], r[8 * i + 1]), pack_bf16x2(r[8 * i + 2], r[8 * i + 3]),
                              pack_bf16x2(r[8 * i + 4], r[8 * i + 5]), pack_bf16x2(r[8 * i + 6], r[8 * i + 7]));
      }
    }
  }
};

// Patch-embed: x[crop, 1 + p, :] = acc + bias + pos_embed[1 + p, :]   (dino_v2.py:219,225-226;
// GEMM row = crop * patches + p; the cls row is written by a separate tiny kernel).
struct EpiPatchEmbed {
  float* x; int ldx; const float* bias; const float* pos; int patches;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    int crop = row / patches, p = row - crop * patches;
    size_t xrow = static_cast<size_t>(crop) * (patches + 1) + 1 + p;
    float4* xp = reinterpret_cast<float4*>(x + xrow * ldx + col0);
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<size_t>(1 + p) * ldx + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      float4 q = __ldg(pp + i);
      xp[i] = make_float4(v[4 * i] + b.x + q.x, v[4 * i + 1] + b.y + q.y, v[4 * i + 2] + b.z + q.z,
                          v[4 * i + 3] + b.w + q.w);
    }
  }
};

// ConvTranspose2d(k=2,s=2) as GEMM + pixel shuffle (+ folded BatchNorm) + GELU
// (linear_head.py:42-48). GEMM column n = (dy*2+dx)*c_out + co; GEMM row = crop*h*w + y*w + x;
// destination row = crop*4hw + (2y+dy)*2w + (2x+dx), destination column = co.
struct EpiConvT2x2Gelu {
  __nv_bfloat16* out; const float* bias; int c_out; int h; int w;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    int quad = col0 / c_out, co0 = col0 - quad * c_out;
    int dy = quad >> 1, dx = quad & 1;
    int hw = h * w;
    int crop = row / hw, rem = row - crop * hw;
    int y = rem / w, xx = rem - y * w;
    size_t orow = static_cast<size_t>(crop) * (4 * hw) + static_cast<size_t>(2 * y + dy) * (2 * w) + (2 * xx + dx);
    uint32_t p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = gelu_erf(v[2 * i] + __ldg(bias + col0 + 2 * i));
      float b = gelu_erf(v[2 * i + 1] + __ldg(bias + col0 + 2 * i + 1));
      p[i] = pack_bf16x2(a, b);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + orow * c_out + co0);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
  }
};

// conv_seg (1x1, channels -> num_classes; mmseg BaseDecodeHead.cls_seg) writing the reference's
// NCHW low-resolution logits: out[crop, cls, pix] = acc + bias[cls]. A warp's 32 rows are 32
// consecutive pixels, so each class store is one coalesced 128-B line.
struct EpiClsNCHW {
  float* out; const float* bias; int num_classes; int pix_per_crop;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    int crop = row / pix_per_crop, pix = row - crop * pix_per_crop;
    float* base = out + static_cast<size_t>(crop) * num_classes * pix_per_crop + pix;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      int cls = col0 + i;
      if (cls < num_classes) base[static_cast<size_t>(cls) * pix_per_crop] = v[i] + __ldg(bias + cls);
    }
  }
};

// Plain fp32 store (tests / generic use): out[row, col] = acc (+ bias).
struct EpiF32 {
  float* out; int ldo; const float* bias;
  __device__ __forceinline__ void operator()(int row, int col0, const float (&v)[32]) const {
    float4* dst = reinterpret_cast<float4*>(out + static_cast<size_t>(row) * ldo + col0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
      if (bias) { b0 = __ldg(bias + col0 + 4 * i); b1 = __ldg(bias + col0 + 4 * i + 1); b2 = __ldg(bias + col0 + 4 * i + 2); b3 = __ldg(bias + col0 + 4 * i + 3); }
      dst[i] = make_float4(v[4 * i] + b0, v[4 * i + 1] + b1, v[4 * i + 2] + b2, v[4 * i + 3] + b3);
    }
  }
};

// ------------------------------------------------------------------ kernel
template <int BLOCK_N, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    int M, int N, int K, Epi epi) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;    // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / GEMM_BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * Cfg::kEpiSplit); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, m_blk * GEMM_BLOCK_M);
          tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, n_blk * BLOCK_N);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(GEMM_BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t da = make_sw128_desc(smem_u32(smem_a + stage * Cfg::kABytes));
          const uint64_t db = make_sw128_desc(smem_u32(smem_b + stage * Cfg::kBBytes));
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k) {
            // +32 B per UMMA_K inside the 128-B swizzle span -> +2 in the (addr >> 4) field
            umma_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    if (half < Cfg::kEpiSplit) {
      constexpr int kColsPerSplit = BLOCK_N / Cfg::kEpiSplit;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const int row = m_blk * GEMM_BLOCK_M + quad * 32 + lane;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerSplit;
#pragma unroll 1
        for (int c = 0; c < kColsPerSplit; c += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          const int col0 = n_blk * BLOCK_N + half * kColsPerSplit + c;
          if (row < M && col0 < N) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            epi(row, col0, v);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

}  // namespace vfm
