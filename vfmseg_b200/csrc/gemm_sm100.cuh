// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = A[M,K] * W[N,K]^T
//   A: activations, row-major bf16 (K contiguous)        -> TMA box {64, 128}, SWIZZLE_128B
//   W: nn.Linear / 1x1-conv weight, row-major bf16 [N,K] -> TMA box {64, 128} (CTA pair) or {64, BLOCK_N}
//   D: fp32 accumulators in TMEM (2 stages), drained by 8 epilogue warps through an epilogue functor
//      (bias / GELU / LayerScale+residual / pixel-shuffle / NCHW logits).
//
// Replaces, on the reference's hot path, every nn.Linear / Conv2d(k=s) / ConvTranspose2d(k=s=2):
//   rein/models/backbones/dino_layers/attention.py:51,67 (qkv, proj), mlp.py:26-28 (fc1, fc2),
//   patch_embed.py:66 (proj conv), rein/models/heads/linear_head.py:36-48 (fusion conv,
//   transposed convs, conv_seg).
//
// Two shapes of the same kernel:
//   CTA_GROUP = 2 : a cluster of two CTAs (one SM pair) owns a 256 x 256 output tile. Each CTA stages its own
//                   128 rows of A and its own 128 rows (half of N) of W, the leader issues
//                   tcgen05.mma.cta_group::2 (UMMA 256x256x16) and each CTA drains its own 128 accumulator rows.
//                   Per SM this halves the W bytes staged per MMA, which buys a 6-stage ring (3 us of TMA
//                   look-ahead) inside 227 KB and leaves shared-memory bandwidth for the tensor core.
//   CTA_GROUP = 1 : one CTA, 128 x BLOCK_N tile (used with BLOCK_N = 32 for the 19-class conv_seg).
//
// Roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one lane, leader CTA only), warp 2 = TMEM
// allocator, warp 3 idle, warps 4..11 = epilogue (TMEM lane quadrant = warp % 4, column half = (warp-4)/4).
//
// Epilogue data path: tcgen05.ld gives every thread one accumulator ROW (32 columns at a time). Row-per-thread
// global access is 32 scattered lines per instruction, which made the first version of this kernel LSU-bound
// (profiles/r1_prof_proj.txt: tensor pipe 29 %). So each warp transposes its 32x32 fragment through a padded
// shared-memory tile and then works column-per-lane: one coalesced 128-byte line per instruction, and per-column
// constants (bias, LayerScale gamma) are loaded once per lane instead of once per element.
#pragma once
#include <type_traits>

#include "sm100_ptx.cuh"

namespace vfm {

#ifdef VFM_EPI_TIMING
__device__ unsigned long long g_epi_dbg[8];   // [0]=wait tmem_full, [1]=tcgen05.ld, [2]=transpose, [3]=elementwise+global, [4]=tiles
#define VFM_TICK(var) const long long var = clock64()
#define VFM_ACC(i, a, b) dbg[i] += static_cast<unsigned long long>((b) - (a))
#else
#define VFM_TICK(var)
#define VFM_ACC(i, a, b)
#endif

constexpr int GEMM_BLOCK_M = 128;   // rows per CTA
constexpr int GEMM_BLOCK_K = 64;    // 64 bf16 = one 128-B swizzle span
constexpr int GEMM_UMMA_K = 16;
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_WARPS = 8;
// per-warp 32x32 fp32 transpose tile, XOR-swizzled (word (r, c) lives at r*32 + (c ^ r)): conflict-free for the
// row-per-thread writes and for both column-per-lane read patterns, and exactly the 4 KB of one TMA box
__device__ __forceinline__ int stg_idx(int r, int c) { return r * 32 + (c ^ r); }

enum EpiMode { EPI_F32 = 0, EPI_BF16X2 = 1, EPI_DIRECT = 2, EPI_TMA_BF16 = 3, EPI_TMA_RED_F32 = 4, EPI_TMA_RES_STATS = 5 };

template <int BLOCK_N, int CTA_GROUP, int EPI_WARP_BYTES = 4096>
struct GemmCfg {
  static constexpr int kBRows = BLOCK_N / CTA_GROUP;                  // W rows staged per CTA
  static constexpr int kABytes = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  static constexpr int kBBytes = kBRows * GEMM_BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEpiWarpBytes = EPI_WARP_BYTES;           // per warp: 32x32 fp32 transpose tile or one 4 KB TMA box (12 KB: EPI_TMA_RES_STATS)
  static constexpr int kEpiStageBytes = GEMM_EPI_WARPS * EPI_WARP_BYTES;
  static constexpr int kBarBytes = 512;
  static constexpr int kBudget = 232448 - 1024 - kBarBytes - kEpiStageBytes;
  static constexpr int kStages = (kBudget / kStageBytes) > 8 ? 8 : (kBudget / kStageBytes);
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : 2 * BLOCK_N;   // 2 accumulator stages
  static constexpr int kEpiSplit = (BLOCK_N >= 64) ? 2 : 1;                  // column halves
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiStageBytes + 1024 /*align slack*/ + kBarBytes;
  static_assert((2 * kStages + 6 + 2 * GEMM_EPI_WARPS) * 8 <= kBarBytes, "barrier block too small");
};

// ------------------------------------------------------------------ epilogues
// EPI_F32   : elem(row, col, v, cc)          lane = one column; 32 consecutive columns per instruction (128 B)
// EPI_BF16X2: elem2(row, col, v0, v1, cc)    lane = two adjacent columns; two rows x 64 B per instruction
// EPI_DIRECT: direct(row, col0, v[32])       row per thread (only where that is already coalesced)
// col_setup(col) loads the per-column constants of the lane once per 32-column chunk.

struct EpiBiasBf16 {            // out = bf16(acc + bias)            (qkv; fusion conv with bias = null)
  static constexpr int kMode = EPI_BF16X2;
  __nv_bfloat16* out; int ldo; const float* bias;
  struct Col { float b0, b1; };
  __device__ __forceinline__ Col col_setup(int col) const {
    if (!bias) return {0.f, 0.f};
    const float2 b = __ldg(reinterpret_cast<const float2*>(bias + col));
    return {b.x, b.y};
  }
  __device__ __forceinline__ void elem2(int row, int col, float v0, float v1, const Col& c) const {
    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(row) * ldo + col) = pack_bf16x2(v0 + c.b0, v1 + c.b1);
  }
};

struct EpiBiasGeluBf16 {        // out = bf16(gelu_erf(acc + bias))  (fc1; mlp.py:35-36)
  static constexpr int kMode = EPI_BF16X2;
  __nv_bfloat16* out; int ldo; const float* bias;
  struct Col { float b0, b1; };
  __device__ __forceinline__ Col col_setup(int col) const {
    const float2 b = __ldg(reinterpret_cast<const float2*>(bias + col));
    return {b.x, b.y};
  }
  __device__ __forceinline__ void elem2(int row, int col, float v0, float v1, const Col& c) const {
    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(row) * ldo + col) =
        pack_bf16x2(gelu_erf(v0 + c.b0), gelu_erf(v1 + c.b1));
  }
};

// TMA-store epilogues: math stays row-per-thread in registers, the 32-row fragment is written to a
// SWIZZLE_128B shared-memory box with conflict-free 16-byte stores and one elected lane hands it to the TMA
// engine, which clips the M / N tails. No per-element global instructions at all.
// epilogues whose apply needs the output row (not just the column) declare kNeedsRow and implement apply_row
template <class E, class = void> struct epi_needs_row : std::false_type {};
template <class E> struct epi_needs_row<E, std::enable_if_t<E::kNeedsRow>> : std::true_type {};

// EPI_TMA_RED_F32 epilogues that declare kPlainStore write the box with a plain TMA store at row out_row(row_base)
// instead of reduce-adding it at row_base
template <class E, class = void> struct epi_plain_store : std::false_type {};
template <class E> struct epi_plain_store<E, std::enable_if_t<E::kPlainStore>> : std::true_type {};

// epilogues that need more than one 4 KB box of staging per warp declare kStageBytesPerWarp
template <class E, class = void> struct epi_warp_bytes : std::integral_constant<int, 4096> {};
template <class E> struct epi_warp_bytes<E, std::enable_if_t<(E::kStageBytesPerWarp > 0)>> : std::integral_constant<int, E::kStageBytesPerWarp> {};

template <class E, class = void> struct epi_has_row_ctx : std::false_type {};
template <class E> struct epi_has_row_ctx<E, std::enable_if_t<E::kHasRowCtx>> : std::true_type {};
template <class E, bool = epi_has_row_ctx<E>::value> struct epi_row_ctx { struct type {}; };
template <class E> struct epi_row_ctx<E, true> { using type = typename E::RowCtx; };
template <class E, class = void> struct epi_deep : std::false_type {};
template <class E> struct epi_deep<E, std::enable_if_t<E::kDeepStaging>> : std::true_type {};

template <bool kGelu>
struct EpiTmaBf16 {             // out = bf16(act(acc + bias)); act = exact-erf GELU (fc1) or identity (qkv, fusion conv)
  static constexpr int kMode = EPI_TMA_BF16;
  static constexpr bool kStore4D = false;
  const float* bias;            // may be null when !kGelu
  __device__ __forceinline__ void apply(int col0, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
      if constexpr (kGelu) {
        v[4 * i] = gelu_erf(v[4 * i]); v[4 * i + 1] = gelu_erf(v[4 * i + 1]);
        v[4 * i + 2] = gelu_erf(v[4 * i + 2]); v[4 * i + 3] = gelu_erf(v[4 * i + 3]);
      }
    }
  }
};

// qkv Linear + bias + 2-D rotary position embedding on the q and k thirds (EVA02, eva_02.py:337-369): columns below
// rope_cols of every patch token (token 0 of a sequence, the cls token, is skipped) get t * cos + rotate_half(t) * sin with
// rotate_half(x)[2i] = -x[2i+1], rotate_half(x)[2i+1] = x[2i] (:54-58), on the fp32 accumulator (one rounding to bf16 instead
// of two). cos / sin: fp32 [tokens_per_seq - 1, 64]; head_dim 64, so a 32-column fragment is half a head.
struct EpiTmaBf16Rope {
  static constexpr int kMode = EPI_TMA_BF16;
  static constexpr bool kStore4D = false;
  static constexpr bool kNeedsRow = true;
  const float* bias; const float* cos_t; const float* sin_t; int rope_cols; FastDiv tokens_per_seq;
  __device__ __forceinline__ void apply_row(int row, int col0, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + col0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
    if (col0 >= rope_cols) return;   // warp-uniform: the v third
    int seq, tok;
    tokens_per_seq.divmod(row, seq, tok);
    if (tok == 0) return;
    const float4* cp = reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(tok - 1) * 64 + (col0 & 63));
    const float4* sp = reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(tok - 1) * 64 + (col0 & 63));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = __ldg(cp + i), sn = __ldg(sp + i);
      const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
      v[4 * i] = x0 * c.x - x1 * sn.x; v[4 * i + 1] = x1 * c.y + x0 * sn.y;
      v[4 * i + 2] = x2 * c.z - x3 * sn.z; v[4 * i + 3] = x3 * c.w + x2 * sn.w;
    }
  }
};

// ConvTranspose2d(k=2,s=2) (+ folded BatchNorm) + GELU with the pixel shuffle done by the TMA engine: the output
// [n*2h*2w, c_out] is described as a 4-D tensor (co, dx, x, R = crop*2h + 2y + dy), so the 32 consecutive x of a
// warp's fragment (one image row y; needs w % 32 == 0 and c_out % 64 == 0) are one box {64, 1, 32, 1}.
// linear_head.py:42-48.
struct EpiTmaConvT {
  static constexpr int kMode = EPI_TMA_BF16;
  static constexpr bool kStore4D = true;
  const float* bias; int c_out; int h; FastDiv div_hw, div_w;
  __device__ __forceinline__ void apply(int col0, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      v[4 * i] = gelu_erf(v[4 * i] + b.x); v[4 * i + 1] = gelu_erf(v[4 * i + 1] + b.y);
      v[4 * i + 2] = gelu_erf(v[4 * i + 2] + b.z); v[4 * i + 3] = gelu_erf(v[4 * i + 3] + b.w);
    }
  }
  __device__ __forceinline__ void coords(int row_base, int col0, int& c0, int& c1, int& c2, int& c3) const {
    const int quad = col0 / c_out;
    int crop, rem, y, x0;
    div_hw.divmod(row_base, crop, rem);
    div_w.divmod(rem, y, x0);
    c0 = col0 - quad * c_out; c1 = quad & 1; c2 = x0; c3 = crop * 2 * h + 2 * y + (quad >> 1);
  }
};

// x += gamma * (acc + bias) as a TMA fp32 reduce-add: the SM never reads the residual stream
// (block.py:112-113 + layer_scale.py:27 for the 44 of 48 residual GEMMs that do not emit a feature tap).
struct EpiTmaResidual {
  static constexpr int kMode = EPI_TMA_RED_F32;
  const float* bias; const float* gamma;
  __device__ __forceinline__ void apply(int col0, float (&v)[32]) const {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col0) + i);
      v[4 * i] = __fmul_rn(g.x, v[4 * i] + b.x); v[4 * i + 1] = __fmul_rn(g.y, v[4 * i + 1] + b.y);
      v[4 * i + 2] = __fmul_rn(g.z, v[4 * i + 2] + b.z); v[4 * i + 3] = __fmul_rn(g.w, v[4 * i + 3] + b.w);
    }
  }
};

// The residual update that also prepares the NEXT LayerNorm (block.py:89-114: x = x + ls(f(norm(x))) is always followed
// by another norm of the new x): x_new = x + gamma * (acc + bias) is formed in the SM (old x arrives through the TMA
// engine, one 32 x 32 box a chunk ahead, its L2 prefetch a tile ahead), and leaves three ways:
//   x      fp32, plain TMA store (same bits as the reduce-add of EpiTmaResidual: one fp32 add, round to nearest);
//   xb     bf16(x_new), TMA store — the A operand of the next GEMM, whose weights carry the LayerNorm's gamma and whose
//          epilogue (EpiTmaBf16LN) applies mean / rstd per row: the layernorm_kernel pass (read 4 B + write 2 B per
//          element) disappears, at the price of this 2 B write and of reading x in the SM instead of in L2;
//   stats  (sum, sum of squares) of every row over the 128 columns this warp owns -> stats[row][n_slots] with
//          slot = column / 128, written once, summed in slot order by the consumer (no atomics: bit-reproducible).
// Two staging shapes. kDeep: 12 KB per epilogue warp (two fp32 boxes, the next chunk's old x in flight while this one is
// processed, + a bf16 box of two chunks) — leaves the main loop a 4-stage ring. !kDeep: ONE 4 KB box per warp, everything
// serial (load -> modify in place -> store fp32 -> same buffer: store bf16 -> next load; ~2.5 k clk per chunk) — the main
// loop keeps its 6 stages; for the K = 4096 GEMM (mlp.fc2: 36 k clk of MMA per tile) the epilogue has that time to spare,
// and measured with the 4-stage ring that GEMM lost 28 % (0.250 against 0.195 ms: three stages in flight do not cover
// the L2 latency of the operand stream).
template <bool kDeep>
struct EpiTmaResidualStats {
  static constexpr int kMode = EPI_TMA_RES_STATS;
  static constexpr bool kDeepStaging = kDeep;
  static constexpr int kStageBytesPerWarp = kDeep ? 12288 : 4096;
  const float* bias; const float* gamma; float2* stats; int n_slots; int l2_prefetch;
  alignas(64) CUtensorMap tmap_xb;   // bf16 [M, N]; kDeep: box {64 cols, 32 rows} SWIZZLE_128B, else {32 cols, 32 rows} SWIZZLE_64B
};

// out = bf16(act(rstd_r * (acc - mean_r * colsum) + bias)): Linear(LayerNorm(x)) with the norm folded away
// (dino_layers/block.py:89-90,112-113 norm1 -> attn.qkv, norm2 -> mlp.fc1). A = bf16(x) un-normalised, W' = bf16(ln_w * W)
// (folded along K on the host), colsum[n] = sum_k W'[n, k] (of the ROUNDED weights, so the mean term cancels what the
// tensor core summed), bias = b + W ln_b (fp32). mean / rstd come from the (sum, sum of squares) slots EpiTmaResidualStats
// wrote for the row, added in slot order; biased variance and eps as torch.nn.LayerNorm.
template <bool kGelu>
struct EpiTmaBf16LN {
  static constexpr int kMode = EPI_TMA_BF16;
  static constexpr bool kStore4D = false;
  static constexpr bool kNeedsRow = true;
  static constexpr bool kHasRowCtx = true;   // mean / rstd of the row are computed once per tile, not once per 32-column chunk
  const float* bias; const float* colsum; const float2* stats; int n_slots; int M; float inv_n; float eps;
  struct RowCtx { float rstd, nm; };
  __device__ __forceinline__ RowCtx row_ctx(int row) const {
    const float2* sp = stats + static_cast<size_t>(row < M ? row : M - 1) * n_slots;
    float sum = 0.f, sq = 0.f;
    for (int i = 0; i < n_slots; i += 2) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(sp + i));
      sum += t.x; sq += t.y; sum += t.z; sq += t.w;
    }
    const float mean = sum * inv_n;
    return {rsqrtf(fmaxf(sq * inv_n - mean * mean, 0.f) + eps), -mean};
  }
  __device__ __forceinline__ void apply_ctx(const RowCtx& rc, int row, int col0, float (&v)[32]) const {
    const float rstd = rc.rstd, nm = rc.nm;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      const float4 c = __ldg(reinterpret_cast<const float4*>(colsum + col0) + i);
      v[4 * i] = fmaf(rstd, fmaf(nm, c.x, v[4 * i]), b.x); v[4 * i + 1] = fmaf(rstd, fmaf(nm, c.y, v[4 * i + 1]), b.y);
      v[4 * i + 2] = fmaf(rstd, fmaf(nm, c.z, v[4 * i + 2]), b.z); v[4 * i + 3] = fmaf(rstd, fmaf(nm, c.w, v[4 * i + 3]), b.w);
      if constexpr (kGelu) {
        v[4 * i] = gelu_erf(v[4 * i]); v[4 * i + 1] = gelu_erf(v[4 * i + 1]);
        v[4 * i + 2] = gelu_erf(v[4 * i + 2]); v[4 * i + 3] = gelu_erf(v[4 * i + 3]);
      }
    }
  }
};

// EpiTmaBf16LN followed by EpiTmaBf16Rope's rotation: norm1 -> attn.qkv + 2-D RoPE of EVA02 (eva_02.py:337-369) with the
// LayerNorm folded into the weights.
struct EpiTmaBf16LNRope {
  static constexpr int kMode = EPI_TMA_BF16;
  static constexpr bool kStore4D = false;
  static constexpr bool kNeedsRow = true;
  static constexpr bool kHasRowCtx = true;
  EpiTmaBf16LN<false> ln; const float* cos_t; const float* sin_t; int rope_cols; FastDiv tokens_per_seq;
  using RowCtx = EpiTmaBf16LN<false>::RowCtx;
  __device__ __forceinline__ RowCtx row_ctx(int row) const { return ln.row_ctx(row); }
  __device__ __forceinline__ void apply_ctx(const RowCtx& rc, int row, int col0, float (&v)[32]) const {
    ln.apply_ctx(rc, row, col0, v);
    if (col0 >= rope_cols) return;   // warp-uniform: the v third
    int seq, tok;
    tokens_per_seq.divmod(row, seq, tok);
    if (tok == 0) return;
    const float4* cp = reinterpret_cast<const float4*>(cos_t + static_cast<size_t>(tok - 1) * 64 + (col0 & 63));
    const float4* sp = reinterpret_cast<const float4*>(sin_t + static_cast<size_t>(tok - 1) * 64 + (col0 & 63));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 c = __ldg(cp + i), sn = __ldg(sp + i);
      const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
      v[4 * i] = x0 * c.x - x1 * sn.x; v[4 * i + 1] = x1 * c.y + x0 * sn.y;
      v[4 * i + 2] = x2 * c.z - x3 * sn.z; v[4 * i + 3] = x3 * c.w + x2 * sn.w;
    }
  }
};

// x += gamma * (acc + bias) on the fp32 residual stream (block.py:112-113 with LayerScale,
// layer_scale.py:27); optionally snapshots the new residual as a bf16 feature tap with the cls
// row dropped (dino_v2.py:261-267) into a token-major [n_crops*patches, tap_ld] buffer.
struct EpiResidual {
  static constexpr int kMode = EPI_F32;
  float* x; int ldx; const float* bias; const float* gamma;
  __nv_bfloat16* tap; int tap_ld; int tap_col0; FastDiv tokens_per_crop;
  static constexpr bool kNeedsOld = true;   // all 32 old values of a chunk are fetched before the first store
  struct Col { float b, g; };
  __device__ __forceinline__ Col col_setup(int col) const { return {__ldg(bias + col), __ldg(gamma + col)}; }
  __device__ __forceinline__ float load_old(int row, int col) const { return x[static_cast<size_t>(row) * ldx + col]; }
  __device__ __forceinline__ void elem(int row, int col, float v, float old, const Col& c) const {
    float* p = x + static_cast<size_t>(row) * ldx + col;
    const float o = __fadd_rn(old, __fmul_rn(c.g, v + c.b));   // unfused, bit-identical to the TMA reduce-add path
    *p = o;
    if (tap != nullptr) {
      int crop, tok;
      tokens_per_crop.divmod(row, crop, tok);   // warp-uniform
      if (tok != 0) tap[static_cast<size_t>(row - crop - 1) * tap_ld + tap_col0 + col] = __float2bfloat16_rn(o);
    }
  }
};

// Patch-embed with fp32 TMA stores (patches % 32 == 0, so a warp's 32-row box never straddles two crops):
// x[crop, cls + p, :] = acc + bias + pos_embed[cls + p, :]. Math row-per-thread in registers (each thread reads its pos-embed
// row as 8 x 16 B, the table is 4 MB and L1/L2-resident), the box leaves through the TMA engine at the row shifted by the
// cls rows in front of it. The column-per-lane version below spent 370 us per 36-window launch waiting on one pos-embed load
// per row (ptxas serialised the 32 loads at the 168-register cap; profiles/r1_prof_patch_embed.txt).
struct EpiTmaPatchEmbed {
  static constexpr int kMode = EPI_TMA_RED_F32;
  static constexpr bool kNeedsRow = true;
  static constexpr bool kPlainStore = true;
  const float* bias; const float* pos; int ldp; FastDiv patches; int cls;
  __device__ __forceinline__ void apply_row(int row, int col0, float (&v)[32]) const {
    const int p = row - patches.div(row) * patches.d;
    const float4* pp = reinterpret_cast<const float4*>(pos + static_cast<size_t>(cls + p) * ldp + col0);
    float4 pe[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pe[i] = __ldg(pp + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col0) + i);
      v[4 * i] = v[4 * i] + b.x + pe[i].x; v[4 * i + 1] = v[4 * i + 1] + b.y + pe[i].y;
      v[4 * i + 2] = v[4 * i + 2] + b.z + pe[i].z; v[4 * i + 3] = v[4 * i + 3] + b.w + pe[i].w;
    }
  }
  __device__ __forceinline__ int out_row(int row_base) const { return row_base + (patches.div(row_base) + 1) * cls; }
};

// Patch-embed: x[crop, 1 + p, :] = acc + bias + pos_embed[1 + p, :]   (dino_v2.py:219,225-226;
// GEMM row = crop * patches + p; the cls row is written by a separate tiny kernel).
struct EpiPatchEmbed {
  static constexpr int kMode = EPI_F32;
  float* x; int ldx; const float* bias; const float* pos; FastDiv patches;
  int cls = 1;   // 1: a cls row precedes each crop's patch rows in x and in pos (DINOv2, EVA02); 0: none (SAM ViT)
  static constexpr bool kNeedsOld = true;   // "old" = the pos-embed value: all 32 are fetched before the first store
  struct Col { float b; };
  __device__ __forceinline__ Col col_setup(int col) const { return {__ldg(bias + col)}; }
  __device__ __forceinline__ float load_old(int row, int col) const {
    int crop, p;
    patches.divmod(row, crop, p);
    return __ldg(pos + static_cast<size_t>(cls + p) * ldx + col);
  }
  __device__ __forceinline__ void elem(int row, int col, float v, float pe, const Col& c) const {
    int crop, p;
    patches.divmod(row, crop, p);
    const size_t xrow = static_cast<size_t>(crop) * (patches.d + cls) + cls + p;
    x[xrow * ldx + col] = v + c.b + pe;
  }
};

// ConvTranspose2d(k=2,s=2) as GEMM + pixel shuffle (+ folded BatchNorm) + GELU
// (linear_head.py:42-48). GEMM column n = (dy*2+dx)*c_out + co; GEMM row = crop*h*w + y*w + x;
// destination row = crop*4hw + (2y+dy)*2w + (2x+dx), destination column = co.
struct EpiConvT2x2Gelu {
  static constexpr int kMode = EPI_BF16X2;
  __nv_bfloat16* out; const float* bias; int c_out; int h; int w; FastDiv div_hw, div_w;
  struct Col { float b0, b1; int co, dy, dx; };
  __device__ __forceinline__ Col col_setup(int col) const {
    const float2 b = __ldg(reinterpret_cast<const float2*>(bias + col));
    const int quad = col / c_out;
    return {b.x, b.y, col - quad * c_out, quad >> 1, quad & 1};
  }
  __device__ __forceinline__ void elem2(int row, int col, float v0, float v1, const Col& c) const {
    const int hw = h * w;
    int crop, rem, y, xx;
    div_hw.divmod(row, crop, rem);
    div_w.divmod(rem, y, xx);
    const size_t orow = static_cast<size_t>(crop) * (4 * hw) + static_cast<size_t>(2 * y + c.dy) * (2 * w) + (2 * xx + c.dx);
    *reinterpret_cast<uint32_t*>(out + orow * c_out + c.co) = pack_bf16x2(gelu_erf(v0 + c.b0), gelu_erf(v1 + c.b1));
  }
};

// conv_seg (1x1, channels -> num_classes; mmseg BaseDecodeHead.cls_seg) writing the reference's
// NCHW low-resolution logits: out[crop, cls, pix] = acc + bias[cls]. A warp's 32 rows are 32
// consecutive pixels, so row-per-thread stores are already one coalesced 128-B line per class.
struct EpiClsNCHW {
  static constexpr int kMode = EPI_DIRECT;
  float* out; const float* bias; int num_classes; FastDiv pix_per_crop;
  __device__ __forceinline__ void direct(int row, int col0, const float (&v)[32]) const {
    int crop, pix;
    pix_per_crop.divmod(row, crop, pix);
    const int ppc = pix_per_crop.d;
    float* base = out + static_cast<size_t>(crop) * num_classes * ppc + pix;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int cls = col0 + i;
      if (cls < num_classes) base[static_cast<size_t>(cls) * ppc] = v[i] + __ldg(bias + cls);
    }
  }
};

// Plain fp32 store (tests / generic use): out[row, col] = acc (+ bias).
struct EpiF32 {
  static constexpr int kMode = EPI_F32;
  float* out; int ldo; const float* bias;
  static constexpr bool kNeedsOld = false;
  struct Col { float b; };
  __device__ __forceinline__ Col col_setup(int col) const { return {bias ? __ldg(bias + col) : 0.f}; }
  __device__ __forceinline__ float load_old(int, int) const { return 0.f; }
  __device__ __forceinline__ void elem(int row, int col, float v, float, const Col& c) const {
    out[static_cast<size_t>(row) * ldo + col] = v + c.b;
  }
};

// ------------------------------------------------------------------ kernel
template <int BLOCK_N, int CTA_GROUP, class Epi>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, int M, int N, int K, const __grid_constant__ Epi epi) {
  using Cfg = GemmCfg<BLOCK_N, CTA_GROUP, epi_warp_bytes<Epi>::value>;
  constexpr int kStages = Cfg::kStages;
  constexpr int TILE_M = GEMM_BLOCK_M * CTA_GROUP;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  // epilogue staging, 1024-B aligned (SWIZZLE_128B boxes): stage bytes are a multiple of 1024
  float* epi_stage = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes + Cfg::kEpiStageBytes);
  uint64_t* full_bar = bars;                   // [kStages]  TMA -> MMA      (leader's copy is the live one)
  uint64_t* empty_bar = bars + kStages;        // [kStages]  MMA -> TMA      (each CTA has its own)
  uint64_t* tmem_full = bars + 2 * kStages;    // [2]        MMA -> epilogue (each CTA has its own)
  uint64_t* tmem_empty = bars + 2 * kStages + 2;  // [2]     epilogue -> MMA (leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_ld = bars + 2 * kStages + 6;   // [8 epilogue warps][2]  TMA -> epilogue warp (EPI_TMA_RES_STATS: old residual boxes)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CTA_GROUP == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int m_tiles = (M + TILE_M - 1) / TILE_M;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = K / GEMM_BLOCK_K;
  const int first_tile = blockIdx.x / CTA_GROUP;
  const int tile_step = gridDim.x / CTA_GROUP;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    // full: one arrival (the leader producer's expect_tx); the peer CTA contributes only transaction bytes
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 4 * Cfg::kEpiSplit * CTA_GROUP); }
    if constexpr (Epi::kMode == EPI_TMA_RES_STATS) {
      for (int s = 0; s < 2 * GEMM_EPI_WARPS; ++s) mbar_init(&epi_ld[s], 1);
    }
    fence_barrier_init();
  }
  pdl_launch_dependents();   // the next kernel on the stream may set itself up on SMs this grid has left
  if (warp == 2) tmem_alloc<Cfg::kTmemCols, CTA_GROUP>(tmem_slot);
  tc_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                // everything above touched no global data; the operands may come from the previous kernel

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own A rows and its own W rows) =====================
    // whole warp converged; one elected lane issues
    int stage = 0; uint32_t phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      const int a_row = m_blk * TILE_M + static_cast<int>(cta_rank) * GEMM_BLOCK_M;
      const int b_row = n_blk * BLOCK_N + static_cast<int>(cta_rank) * Cfg::kBRows;
      for (int kb = 0; kb < k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one_sync()) {
          if constexpr (CTA_GROUP == 2) {
            // transaction bytes of both CTAs land on the leader's barrier (its tx-count may dip below zero
            // until the leader's expect_tx of the same phase arrives; the pending arrival keeps the phase open)
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
            tma_load_2d_2sm(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, a_row);
            tma_load_2d_2sm(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, b_row);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
            tma_load_2d(smem_a + stage * Cfg::kABytes, &tmap_a, &full_bar[stage], kb * GEMM_BLOCK_K, a_row);
            tma_load_2d(smem_b + stage * Cfg::kBBytes, &tmap_b, &full_bar[stage], kb * GEMM_BLOCK_K, b_row);
          }
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; whole warp converged, one elected lane issues) ==========
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(TILE_M, BLOCK_N, 0, 0);
      const uint64_t da0 = make_sw128_desc(smem_u32(smem_a));
      const uint64_t db0 = make_sw128_desc(smem_u32(smem_b));
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one_sync()) {
            // stage offset and the +32 B per UMMA_K step both go into the (addr >> 4) field of the descriptor
            const uint64_t da = da0 + static_cast<uint64_t>(stage * (Cfg::kABytes >> 4));
            const uint64_t db = db0 + static_cast<uint64_t>(stage * (Cfg::kBBytes >> 4));
#pragma unroll
            for (int k = 0; k < GEMM_BLOCK_K / GEMM_UMMA_K; ++k)
              umma_ss<CTA_GROUP>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            tc_commit<CTA_GROUP>(&empty_bar[stage]);                         // frees the slot in both CTAs
            if (kb == k_blocks - 1) tc_commit<CTA_GROUP>(&tmem_full[acc]);   // accumulator ready in both CTAs
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    const int half = (warp - 4) >> 2;
    if (half < Cfg::kEpiSplit) {
      constexpr int kColsPerSplit = BLOCK_N / Cfg::kEpiSplit;
      float* stg = epi_stage + (warp - 4) * (Cfg::kEpiWarpBytes / 4);   // 4096 B per warp (12288: EPI_TMA_RES_STATS)
      int acc = 0; uint32_t acc_phase = 0;
#ifdef VFM_EPI_TIMING
      unsigned long long dbg[5] = {0, 0, 0, 0, 0};
#endif
      if constexpr (Epi::kMode == EPI_TMA_RES_STATS && !epi_deep<Epi>::value) {
        // ---- residual + LayerNorm statistics, serial single-box staging (see EpiTmaResidualStats<false>)
        constexpr int kChunks = kColsPerSplit / 32;
        uint8_t* buf = reinterpret_cast<uint8_t*>(stg);
        uint64_t* ldb = epi_ld + (warp - 4) * 2;
        const int sw = lane & 7;             // SWIZZLE_128B (fp32 box, 128-byte rows): 16-byte chunk index ^= row % 8
        const int sw64 = (lane >> 1) & 3;    // SWIZZLE_64B (bf16 box, 64-byte rows): 16-byte chunk index ^= (row / 2) % 4
        auto tile_rc = [&](int tile, int& row, int& col) {
          const int mb = tile / n_tiles, nb = tile - mb * n_tiles;
          row = mb * TILE_M + static_cast<int>(cta_rank) * GEMM_BLOCK_M + quad * 32;
          col = nb * BLOCK_N + half * kColsPerSplit;
        };
        auto issue_load = [&](int col, int row) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(ldb, 4096);
            tma_load_2d(buf, &tmap_out, ldb, col, row);
          }
          __syncwarp();
        };
        uint32_t q = 0;
        if (first_tile < num_tiles) {
          int r0, c0;
          tile_rc(first_tile, r0, c0);
          issue_load(c0, r0);
        }
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
          int row_base, colw;
          tile_rc(tile, row_base, colw);
          const int next_tile = tile + tile_step;
          int nrow = 0, ncol = 0;
          if (next_tile < num_tiles) {   // the next tile's old x: into L2 now
            tile_rc(next_tile, nrow, ncol);
            if (epi.l2_prefetch && lane < kChunks) tma_prefetch_l2_2d(&tmap_out, ncol + 32 * lane, nrow);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerSplit;
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c, ++q) {
            uint32_t r[32];
            tmem_ld32(taddr + 32 * c, r);
            mbar_wait(ldb, q & 1u);
            tmem_ld_wait();
            if (c == kChunks - 1) {   // the tile's last accumulator columns are in registers: the MMA issuer may have the TMEM stage back now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CTA_GROUP == 2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
              }
            }
            const int col0 = colw + 32 * c;
            uint8_t* rowp = buf + lane * 128;
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 o = *reinterpret_cast<const float4*>(rowp + ((j ^ sw) << 4));
              const float4 b = __ldg(reinterpret_cast<const float4*>(epi.bias + col0) + j);
              const float4 g = __ldg(reinterpret_cast<const float4*>(epi.gamma + col0) + j);
              v[4 * j] = __fadd_rn(o.x, __fmul_rn(g.x, __uint_as_float(r[4 * j]) + b.x));
              v[4 * j + 1] = __fadd_rn(o.y, __fmul_rn(g.y, __uint_as_float(r[4 * j + 1]) + b.y));
              v[4 * j + 2] = __fadd_rn(o.z, __fmul_rn(g.z, __uint_as_float(r[4 * j + 2]) + b.z));
              v[4 * j + 3] = __fadd_rn(o.w, __fmul_rn(g.w, __uint_as_float(r[4 * j + 3]) + b.w));
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one_sync()) { tma_store_2d(&tmap_out, buf, col0, row_base); tma_store_commit(); }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 32; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(v[2 * i], v[2 * i + 1]);
            tma_store_wait_read<0>();   // the fp32 box has left: the same buffer takes the bf16 box
            __syncwarp();
            uint8_t* rowb = buf + lane * 64;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(rowb + ((j ^ sw64) << 4)) = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one_sync()) { tma_store_2d(&epi.tmap_xb, buf, col0, row_base); tma_store_commit(); }
            __syncwarp();
            tma_store_wait_read<0>();
            __syncwarp();
            if (c + 1 < kChunks) issue_load(col0 + 32, row_base);
            else if (next_tile < num_tiles) issue_load(ncol, nrow);
          }
          if (row_base + lane < M)
            epi.stats[static_cast<size_t>(row_base + lane) * epi.n_slots + colw / kColsPerSplit] =
                make_float2((s4[0] + s4[1]) + (s4[2] + s4[3]), (q4[0] + q4[1]) + (q4[2] + q4[3]));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        tma_store_wait<0>();
      } else if constexpr (Epi::kMode == EPI_TMA_RES_STATS) {
        // ---- residual + LayerNorm statistics (see EpiTmaResidualStats). Per warp: bx[2] = fp32 boxes (old x in, new x out,
        // in place), bb = bf16 box (two 32-column chunks side by side). q counts the chunks of this warp's stream:
        // buffer q & 1, barrier parity (q >> 1) & 1.
        static_assert(kColsPerSplit % 64 == 0, "EPI_TMA_RES_STATS: 64-column bf16 boxes");
        constexpr int kChunks = kColsPerSplit / 32;
        uint8_t* bx = reinterpret_cast<uint8_t*>(stg);
        uint8_t* bb = bx + 8192;
        uint64_t* ldb = epi_ld + (warp - 4) * 2;
        const int sw = lane & 7;   // SWIZZLE_128B: 16-byte chunk index ^= row % 8
        auto tile_rc = [&](int tile, int& row, int& col) {
          const int mb = tile / n_tiles, nb = tile - mb * n_tiles;
          row = mb * TILE_M + static_cast<int>(cta_rank) * GEMM_BLOCK_M + quad * 32;
          col = nb * BLOCK_N + half * kColsPerSplit;
        };
        auto issue_load = [&](uint32_t qq, int col, int row) {
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&ldb[qq & 1u], 4096);
            tma_load_2d(bx + (qq & 1u) * 4096, &tmap_out, &ldb[qq & 1u], col, row);
          }
          __syncwarp();
        };
        uint32_t q = 0;
        if (first_tile < num_tiles) {
          int r0, c0;
          tile_rc(first_tile, r0, c0);
          issue_load(0, c0, r0);
        }
        for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
          int row_base, colw;
          tile_rc(tile, row_base, colw);
          const int next_tile = tile + tile_step;
          int nrow = 0, ncol = 0;
          if (next_tile < num_tiles) {   // the next tile's old x: into L2 now (a whole tile ahead of its first use)
            tile_rc(next_tile, nrow, ncol);
            if (epi.l2_prefetch && lane < kChunks) tma_prefetch_l2_2d(&tmap_out, ncol + 32 * lane, nrow);
          }
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerSplit;
          float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};   // four interleaved partial sums (fixed order)
#pragma unroll 1
          for (int c = 0; c < kChunks; ++c, ++q) {
            uint32_t r[32];
            tmem_ld32(taddr + 32 * c, r);
            mbar_wait(&ldb[q & 1u], (q >> 1) & 1u);
            tmem_ld_wait();
            if (c == kChunks - 1) {   // the tile's last accumulator columns are in registers: the MMA issuer may have the TMEM stage back now
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                if constexpr (CTA_GROUP == 2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
              }
            }
            const int col0 = colw + 32 * c;
            uint8_t* rowp = bx + (q & 1u) * 4096 + lane * 128;
            float v[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 o = *reinterpret_cast<const float4*>(rowp + ((j ^ sw) << 4));
              const float4 b = __ldg(reinterpret_cast<const float4*>(epi.bias + col0) + j);
              const float4 g = __ldg(reinterpret_cast<const float4*>(epi.gamma + col0) + j);
              v[4 * j] = __fadd_rn(o.x, __fmul_rn(g.x, __uint_as_float(r[4 * j]) + b.x));
              v[4 * j + 1] = __fadd_rn(o.y, __fmul_rn(g.y, __uint_as_float(r[4 * j + 1]) + b.y));
              v[4 * j + 2] = __fadd_rn(o.z, __fmul_rn(g.z, __uint_as_float(r[4 * j + 2]) + b.z));
              v[4 * j + 3] = __fadd_rn(o.w, __fmul_rn(g.w, __uint_as_float(r[4 * j + 3]) + b.w));
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) { s4[i & 3] += v[i]; q4[i & 3] = fmaf(v[i], v[i], q4[i & 3]); }
            // the other fp32 box (stored a chunk ago) and the bf16 box have been read by the TMA engine
            tma_store_wait_read<0>();
            __syncwarp();
            if (c + 1 < kChunks) issue_load(q + 1, col0 + 32, row_base);
            else if (next_tile < num_tiles) issue_load(q + 1, ncol, nrow);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            const int hbox = c & 1;   // even chunk -> left 64 B of the bf16 rows, odd chunk -> right 64 B
            uint8_t* rowb = bb + lane * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(rowb + (((hbox * 4 + j) ^ sw) << 4)) =
                  make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                             pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one_sync()) {
              tma_store_2d(&tmap_out, bx + (q & 1u) * 4096, col0, row_base);
              if (hbox == 1) tma_store_2d(&epi.tmap_xb, bb, col0 - 32, row_base);
              tma_store_commit();
            }
            __syncwarp();
          }
          if (row_base + lane < M)
            epi.stats[static_cast<size_t>(row_base + lane) * epi.n_slots + colw / kColsPerSplit] =
                make_float2((s4[0] + s4[1]) + (s4[2] + s4[3]), (q4[0] + q4[1]) + (q4[2] + q4[3]));
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        tma_store_wait<0>();
      } else
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
        const int row_base = m_blk * TILE_M + static_cast<int>(cta_rank) * GEMM_BLOCK_M + quad * 32;
        typename epi_row_ctx<Epi>::type row_ctx{};   // per-row constants (LayerNorm mean / rstd), fetched while the main loop still runs
        if constexpr (epi_has_row_ctx<Epi>::value) row_ctx = epi.row_ctx(row_base + lane);
        VFM_TICK(t_w0);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        VFM_TICK(t_w1);
        VFM_ACC(0, t_w0, t_w1);
#ifdef VFM_EPI_TIMING
        dbg[4] += 1;
#endif
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BLOCK_N + half * kColsPerSplit;
#pragma unroll 1
        for (int c = 0; c < kColsPerSplit; c += 32) {
          uint32_t r[32];
          VFM_TICK(t0);
          tmem_ld32(taddr + c, r);
          tmem_ld_wait();
          if (c + 32 >= kColsPerSplit) {   // the tile's last accumulator columns are in registers: the MMA issuer may have the TMEM stage back
            tc_fence_before();             // now, a chunk of epilogue math and stores earlier than at the end of the tile
            __syncwarp();
            if (lane == 0) {
              if constexpr (CTA_GROUP == 2) mbar_arrive_leader(&tmem_empty[acc]); else mbar_arrive(&tmem_empty[acc]);
            }
          }
          VFM_TICK(t1);
          VFM_ACC(1, t0, t1);
          const int col0 = n_blk * BLOCK_N + half * kColsPerSplit + c;
          if constexpr (Epi::kMode == EPI_TMA_BF16 || Epi::kMode == EPI_TMA_RED_F32) {
            if (col0 < N && (!epi_plain_store<Epi>::value || row_base < M)) {   // warp-uniform
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
              if constexpr (epi_has_row_ctx<Epi>::value) epi.apply_ctx(row_ctx, row_base + lane, col0, v);
              else if constexpr (epi_needs_row<Epi>::value) epi.apply_row(row_base + lane, col0, v);
              else epi.apply(col0, v);
              uint8_t* box = reinterpret_cast<uint8_t*>(stg);
              uint8_t* rowp = box + lane * 128;
              const int sw = lane & 7;   // SWIZZLE_128B: 16-byte chunk index ^= row % 8
              if constexpr (Epi::kMode == EPI_TMA_RED_F32) {
                tma_store_wait_read<0>();        // previous box of this warp has been read by the TMA engine
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  *reinterpret_cast<float4*>(rowp + ((j ^ sw) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (elect_one_sync()) {
                  if constexpr (epi_plain_store<Epi>::value) tma_store_2d(&tmap_out, box, col0, epi.out_row(row_base));
                  else tma_reduce_add_2d(&tmap_out, box, col0, row_base);
                  tma_store_commit();
                }
              } else {
                const int hbox = (c >> 5) & 1;   // even chunk -> left 64 B of the 128-B rows, odd chunk -> right 64 B
                if (hbox == 0) { tma_store_wait_read<0>(); __syncwarp(); }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(rowp + (((hbox * 4 + j) ^ sw) << 4)) =
                      make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                 pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                if (hbox == 1 || col0 + 32 >= N) {
                  fence_proxy_async_smem();
                  __syncwarp();
                  if (elect_one_sync()) {
                    if constexpr (Epi::kStore4D) {
                      int c0, c1, c2, c3;
                      epi.coords(row_base, col0 - hbox * 32, c0, c1, c2, c3);
                      tma_store_4d(&tmap_out, box, c0, c1, c2, c3);
                    } else {
                      tma_store_2d(&tmap_out, box, col0 - hbox * 32, row_base);
                    }
                    tma_store_commit();
                  }
                }
              }
            }
          } else if constexpr (Epi::kMode == EPI_DIRECT) {
            if (row_base + lane < M && col0 < N) {
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
              epi.direct(row_base + lane, col0, v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) stg[stg_idx(lane, i)] = __uint_as_float(r[i]);
            __syncwarp();
            VFM_TICK(t2);
            VFM_ACC(2, t1, t2);
            if (col0 < N) {   // warp-uniform
              // Fast path (every tile but the M tail): no per-row predicate, so the fully unrolled loops carry no
              // branches and the shared-memory reads / global loads of all rows are in flight together.
              const bool full = row_base + 32 <= M;
              if constexpr (Epi::kMode == EPI_F32) {
                const auto cc = epi.col_setup(col0 + lane);
                auto body = [&](auto check) {
                  constexpr bool kCheck = decltype(check)::value;
                  float old[32];
                  if constexpr (Epi::kNeedsOld) {
#pragma unroll
                    for (int rr = 0; rr < 32; ++rr)
                      old[rr] = (!kCheck || row_base + rr < M) ? epi.load_old(row_base + rr, col0 + lane) : 0.f;
                  }
#pragma unroll
                  for (int rr = 0; rr < 32; ++rr) {
                    if (!kCheck || row_base + rr < M)
                      epi.elem(row_base + rr, col0 + lane, stg[stg_idx(rr, lane)], Epi::kNeedsOld ? old[rr] : 0.f, cc);
                  }
                };
                if (full) body(std::false_type{}); else body(std::true_type{});
              } else {
                const int l16 = lane & 15, hi = lane >> 4;
                const auto cc = epi.col_setup(col0 + 2 * l16);
                auto body = [&](auto check) {
                  constexpr bool kCheck = decltype(check)::value;
#pragma unroll
                  for (int it = 0; it < 16; ++it) {
                    const int rr = 2 * it + hi;
                    const float v0 = stg[stg_idx(rr, 2 * l16)], v1 = stg[stg_idx(rr, 2 * l16 + 1)];
                    if (!kCheck || row_base + rr < M) epi.elem2(row_base + rr, col0 + 2 * l16, v0, v1, cc);
                  }
                };
                if (full) body(std::false_type{}); else body(std::true_type{});
              }
            }
            __syncwarp();
            VFM_TICK(t3);
            VFM_ACC(3, t2, t3);
          }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
      if constexpr (Epi::kMode == EPI_TMA_BF16 || Epi::kMode == EPI_TMA_RED_F32) tma_store_wait<0>();
#ifdef VFM_EPI_TIMING
      if (lane == 0 && warp == 4) for (int i = 0; i < 5; ++i) atomicAdd(&g_epi_dbg[i], dbg[i]);
#endif
    }
  }

  tc_fence_before();
  if constexpr (CTA_GROUP == 2) cluster_sync(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols, CTA_GROUP>(tmem_base);
  }
}

}  // namespace vfm
