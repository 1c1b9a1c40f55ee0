// Window attention with the decomposed rel-pos bias on tcgen05 (SAM ViT windowed blocks: 196 tokens, head_dim 80).
//
// Replaces Attention.forward's core for one window, rein/models/backbones/sam_vit.py:272-287 with
// add_decomposed_rel_pos :391-428, when the whole key set fits ONE score tile (seq_len <= 208):
//   S[128 x 208] = [q(0:64) | q(64:80) | rel / scale] x [k(0:64) | k(64:80) | onehot(kh) onehot(kw)]^T
//        seven tcgen05.mma (M 128, N 208, K 16) into TMEM: four over the first 64 head dims, one over the last 16, and
//        ceil((k_h + k_w) / 16) over the bias columns — the bias is added BY the tensor core: rel_h[q, kh] + rel_w[q, kw]
//        = [rel_h | rel_w] . [onehot(kh) | onehot(kw)], exact in fp32 (bf16 x 1.0). rel comes from the table terms the
//        qkv GEMM wrote behind q | k | v (see vfm_attention_relpos_ex), pre-divided by the score scale that is applied to
//        the whole sum afterwards.
//   softmax: one query row per thread (TMEM lane), two passes over the 208 columns (max, then exp2 / sum / pack);
//        P bf16 goes to its own TMEM columns.
//   O[128 x 80] = P V: 13 k-steps x (N 64 over head dims 0:64, N 16 over 64:80), A = P from TMEM, B = V as MN-major
//        SWIZZLE_128B tiles.
// One CTA per (window, head, 128-query tile), two CTAs per SM; operands arrive by TMA straight out of the packed qkv
// rows. Shared memory holds each operand at its real width: the 64-dim parts as SWIZZLE_128B tiles, the 16-dim
// remainders as SWIZZLE_32B tiles (one k-step), the 32 bias columns as SWIZZLE_64B tiles — 106 KB per CTA. TMEM: the
// probabilities overwrite the scores in place (pass 2 has consumed a 32-column chunk before it writes 16 packed columns
// at half the offset) and O lands on score columns 112..191, so one CTA needs 208 of its 256 columns.
// No software pipeline inside the CTA: the second CTA of the SM fills the gaps. The mma.sync kernel in sam_ops.cuh stays
// for the global blocks (1024 / 4096 keys).
#pragma once
#include "sm100_ptx.cuh"

namespace vfm {

constexpr int WIN_BLOCK_Q = 128;
constexpr int WIN_KEYS = 208;                                  // 13 x 16: one score tile
constexpr int WIN_D = 80;
constexpr int WIN_THREADS = 192;                               // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2..5 softmax
constexpr int WIN_Q64 = WIN_BLOCK_Q * 128, WIN_K64 = WIN_KEYS * 128;   // [rows x 64 bf16] SWIZZLE_128B: 16 KB / 26 KB
constexpr int WIN_Q32 = WIN_BLOCK_Q * 64, WIN_K32 = WIN_KEYS * 64;     // [rows x 32 bf16] SWIZZLE_64B (bias columns)
constexpr int WIN_Q16 = WIN_BLOCK_Q * 32, WIN_K16 = WIN_KEYS * 32;     // [rows x 16 bf16] SWIZZLE_32B (head dims 64..79)
constexpr int WIN_SMEM_BYTES = WIN_Q64 + 2 * WIN_K64 + WIN_Q32 + WIN_K32 + WIN_Q16 + 2 * WIN_K16 + 1024 + 128;
constexpr uint32_t WIN_TMEM_COLS = 256;
constexpr uint32_t WIN_COL_S = 0, WIN_COL_P = 0, WIN_COL_O = 112;      // P in place over S; O over consumed S columns

struct WinParams {
  int seq_len, heads, k_h, k_w;
  int ld, g_col0;               // row pitch of the qkv buffer (elements); first table-term column (-1: no bias)
  float scale;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
  const int* out_map;           // optional: output row of each (window-order) query row, < 0 = padding row, not written —
                                // window_unpartition (sam_vit.py:335-346) folded into the store
};

// Shared-memory matrix descriptors for the narrower swizzle modes (cf. make_sw128_desc): rows of 64 B / 32 B, 8-row groups
// 512 B / 256 B apart; layout type 4 = SWIZZLE_64B, 6 = SWIZZLE_32B. K-major: advance K by 16 -> +32 B; MN-major (rows =
// K): advance K by 16 rows.
__device__ __forceinline__ uint64_t make_sw_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout_type) << 61;
  return d;
}

__global__ void __launch_bounds__(WIN_THREADS, 2)
attention_win_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                     const __grid_constant__ CUtensorMap tmap_q16, const __grid_constant__ CUtensorMap tmap_kv16, const WinParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
  uint8_t* sQ0 = smem;                      // SWIZZLE_128B tiles first (1024-byte aligned)
  uint8_t* sK0 = sQ0 + WIN_Q64;
  uint8_t* sV0 = sK0 + WIN_K64;
  uint8_t* sQB = sV0 + WIN_K64;             // SWIZZLE_64B: bias terms of the query rows [128][32] (rel_h | 0 | rel_w | 0)
  uint8_t* sE = sQB + WIN_Q32;              //              one-hot key tile [208][32]
  uint8_t* sQ1 = sE + WIN_K32;              // SWIZZLE_32B: head dims 64..79
  uint8_t* sK1 = sQ1 + WIN_Q16;
  uint8_t* sV1 = sK1 + WIN_K16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV1 + WIN_K16);
  uint64_t* load_full = bars;               // TMA -> MMA
  uint64_t* s_full = bars + 1;              // MMA -> softmax
  uint64_t* p_full = bars + 2;              // softmax -> MMA
  uint64_t* o_full = bars + 3;              // MMA -> softmax
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int C = p.heads * WIN_D;
  const int row0 = seq * p.seq_len;
  const int q0 = qt * WIN_BLOCK_Q;
  const int kk = p.g_col0 >= 0 ? p.k_h + p.k_w : 0;
  const int bias_steps = kk > 0 ? 2 : 0;

  if (tid == 0) {
    mbar_init(load_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
    // the loads go out at once: their latency runs under the bias-operand build below
    mbar_arrive_expect_tx(load_full, WIN_Q64 + 2 * WIN_K64 + WIN_Q16 + 2 * WIN_K16);
    const int qc = head * WIN_D, kc = C + head * WIN_D, vc = 2 * C + head * WIN_D;
    for (int h = 0; h < 2; ++h) {   // 64-row boxes of the Q tiles, 104-row boxes of the K / V tiles
      tma_load_2d(sQ0 + h * (WIN_Q64 / 2), &tmap_q, load_full, qc, row0 + q0 + 64 * h);
      tma_load_2d(sQ1 + h * (WIN_Q16 / 2), &tmap_q16, load_full, qc + 64, row0 + q0 + 64 * h);
      tma_load_2d(sK0 + h * (WIN_K64 / 2), &tmap_kv, load_full, kc, row0 + 104 * h);
      tma_load_2d(sK1 + h * (WIN_K16 / 2), &tmap_kv16, load_full, kc + 64, row0 + 104 * h);
      tma_load_2d(sV0 + h * (WIN_K64 / 2), &tmap_kv, load_full, vc, row0 + 104 * h);
      tma_load_2d(sV1 + h * (WIN_K16 / 2), &tmap_kv16, load_full, vc + 64, row0 + 104 * h);
    }
  }
  if (warp == 1) tmem_alloc<WIN_TMEM_COLS>(tmem_slot);
  // ---- bias operands, built by all threads while nothing else is running. Fixed column layout (k_h, k_w <= 16):
  // k-step 0 = rel_h | zeros, k-step 1 = rel_w | zeros; the one-hot tile has its ones at columns kh and 16 + kw.
  if (kk > 0) {
    const float inv_scale = 1.f / p.scale;
    const int Lh = 2 * p.k_h - 1, Lw = 2 * p.k_w - 1;
    if (tid < WIN_BLOCK_Q) {
      // one thread per query row: 2 x 16 contiguous table terms read backwards (all 32 loads in flight together), scaled,
      // written as four 16-byte chunks of the swizzled row. (Half a warp per run — two cache lines per load instruction
      // instead of 32 — measured slower, 242 vs 197 us: more steps, each with its own memory round trip.)
      const int r = tid;
      const int q = q0 + r;
      __nv_bfloat16 v[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) v[c] = __float2bfloat16(0.f);
      if (q < p.seq_len) {
        const int qh = q / p.k_w, qw = q - qh * p.k_w;
        const __nv_bfloat16* gp = p.qkv + static_cast<size_t>(row0 + q) * p.ld + p.g_col0;
        const __nv_bfloat16* gh = gp + head * Lh + qh + p.k_h - 1;
        const __nv_bfloat16* gw = gp + p.heads * Lh + head * Lw + qw + p.k_w - 1;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < p.k_h) v[c] = gh[-c];
          if (c < p.k_w) v[16 + c] = gw[-c];
        }
      }
      uint32_t w[16];
#pragma unroll
      for (int c = 0; c < 16; ++c)
        w[c] = pack_bf16x2(__bfloat162float(v[2 * c]) * inv_scale, __bfloat162float(v[2 * c + 1]) * inv_scale);
      uint8_t* rowp = sQB + r * 64;        // SWIZZLE_64B: 16-byte chunk index ^ ((row >> 1) & 3)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        *reinterpret_cast<uint4*>(rowp + ((ch ^ ((r >> 1) & 3)) << 4)) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
    } else {
      // the other 64 threads: one-hot rows, columns 0..31 (four chunks)
      for (int key = tid - WIN_BLOCK_Q; key < WIN_KEYS; key += WIN_THREADS - WIN_BLOCK_Q) {
        uint32_t w[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) w[c] = 0u;
        if (key < p.seq_len) {
          const int h = key / p.k_w, ww = key - h * p.k_w;
          w[h >> 1] |= 0x3f80u << ((h & 1) * 16);                 // bf16 1.0
          w[8 + (ww >> 1)] |= 0x3f80u << ((ww & 1) * 16);
        }
        uint8_t* rowp = sE + key * 64;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
          *reinterpret_cast<uint4*>(rowp + ((ch ^ ((key >> 1) & 3)) << 4)) = make_uint4(w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
      }
    }
    fence_proxy_async_smem();   // the tensor core reads these tiles through the async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // (loads were issued in the prologue)
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(WIN_BLOCK_Q, WIN_KEYS, 0, 0);
    constexpr uint32_t idesc_pv64 = make_idesc_bf16(WIN_BLOCK_Q, 64, 0, 1);   // B = V is MN-major
    constexpr uint32_t idesc_pv16 = make_idesc_bf16(WIN_BLOCK_Q, 16, 0, 1);
    const uint32_t tmem_s = tmem_base + WIN_COL_S, tmem_p = tmem_base + WIN_COL_P, tmem_o = tmem_base + WIN_COL_O;
    mbar_wait(load_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint64_t dq0 = make_sw128_desc(smem_u32(sQ0)), dk0 = make_sw128_desc(smem_u32(sK0));
      const uint64_t dq1 = make_sw_desc(smem_u32(sQ1), 256, 6), dk1 = make_sw_desc(smem_u32(sK1), 256, 6);
      const uint64_t dqb = make_sw_desc(smem_u32(sQB), 512, 4), de = make_sw_desc(smem_u32(sE), 512, 4);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_ss(tmem_s, dq0 + 2 * k, dk0 + 2 * k, idesc_s, k != 0);   // head dims 0..63
      umma_ss(tmem_s, dq1, dk1, idesc_s, true);                                                  // head dims 64..79
      for (int k = 0; k < bias_steps; ++k) umma_ss(tmem_s, dqb + 2 * k, de + 2 * k, idesc_s, true);
      tc_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_full, 0);
    tc_fence_after();
    if (elect_one_sync()) {
      const uint64_t dv0 = make_sw128_desc(smem_u32(sV0)), dv1 = make_sw_desc(smem_u32(sV1), 256, 6);
#pragma unroll 1
      for (int k = 0; k < WIN_KEYS / 16; ++k) {
        // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
        umma_ts(tmem_o, tmem_p + 8 * k, dv0 + 128 * k, idesc_pv64, k != 0);
        umma_ts(tmem_o + 64, tmem_p + 8 * k, dv1 + 32 * k, idesc_pv16, k != 0);   // 16 key rows of 32 B
      }
      tc_commit(o_full);
    }
    __syncwarp();
  } else {
    // ===================== softmax + output (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + WIN_COL_S;
    const uint32_t tmem_p = tmem_base + lane_base + WIN_COL_P;
    const uint32_t tmem_o = tmem_base + lane_base + WIN_COL_O;
    const float sc = p.scale * 1.4426950408889634f;
    mbar_wait(s_full, 0);
    tc_fence_after();
    // pass 1: row max over the real keys
    float m = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < WIN_KEYS / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_s + 32 * c, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (32 * c + i < p.seq_len) m = fmaxf(m, __uint_as_float(r[i]));
    }
    {
      uint32_t r[16];
      tmem_ld16(tmem_s + (WIN_KEYS / 32) * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if ((WIN_KEYS / 32) * 32 + i < p.seq_len) m = fmaxf(m, __uint_as_float(r[i]));
    }
    const float ms = m * sc;
    // pass 2: P = exp2(s * sc - ms) rounded to bf16; the row sum adds up the rounded values the PV MMA consumes
    float l = 0.f;
#pragma unroll 1
    for (int c = 0; c < (WIN_KEYS + 31) / 32; ++c) {   // 32 keys = 16 packed columns per store; the last step holds 16 keys
      uint32_t pk[16];
      uint32_t r2[2][16];
      tmem_ld16(tmem_s + 32 * c, r2[0]);
      if (32 * c + 16 < WIN_KEYS) tmem_ld16(tmem_s + 32 * c + 16, r2[1]);
      tmem_ld_wait();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int key0 = 32 * c + 16 * hh;
        if (key0 < WIN_KEYS) {
          const uint32_t (&r)[16] = r2[hh];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int key = key0 + 2 * i;
            const float e0 = key < p.seq_len ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), sc, -ms)) : 0.f;
            const float e1 = key + 1 < p.seq_len ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), sc, -ms)) : 0.f;
            const __nv_bfloat162 bb = __floats2bfloat162_rn(e0, e1);
            l += __low2float(bb) + __high2float(bb);
            pk[8 * hh + i] = *reinterpret_cast<const uint32_t*>(&bb);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[8 * hh + i] = 0u;   // columns past the last k-step (inside the P region)
        }
      }
      tmem_st16(tmem_p + 16 * c, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_full);
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int q = q0 + row;
    int orow = row0 + q;
    if (p.out_map != nullptr) orow = q < p.seq_len ? __ldg(p.out_map + row0 + q) : -1;
    uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(orow < 0 ? 0 : orow) * C + head * WIN_D);
#pragma unroll 1
    for (int c = 0; c < WIN_D / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tmem_o + 16 * c, r);
      tmem_ld_wait();
      if (q < p.seq_len && orow >= 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(r[8 * i + e]) * inv;
          dst[2 * c + i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<WIN_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
