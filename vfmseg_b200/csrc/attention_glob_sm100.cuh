// Attention with the decomposed rel-pos bias over a whole token grid on tcgen05 (SAM ViT global blocks: 1024 or 4096
// tokens, head_dim 80). Replaces Attention.forward's core, rein/models/backbones/sam_vit.py:272-287 with
// add_decomposed_rel_pos :391-428, for sequences that need a key-tile loop.
//
// One CTA per (sequence, head, 128-query tile), two CTAs per SM, 64-key tiles, online softmax. Per key tile t:
//   S_t[128 x 64] = [q(0:64) | q(64:80) | rel / scale] x [k(0:64) | k(64:80) | onehot(kh) onehot(kw)]^T
//        4 + 1 + (BH + BW) / 16 tcgen05.mma (M 128, N 64, K 16); the bias is added by the tensor core exactly as in
//        attention_win_sm100.cuh. The one-hot rows of ALL keys are a constant of the grid ([keys, BH + BW] bf16, built once
//        by the host side) and arrive by TMA like K.
//   softmax: one query row per thread; the 64 scores go to registers, P = exp2(s - m_ref) rounded to bf16 overwrites the
//        first 32 score columns, the row sum stays in a register; O (TMEM columns 128..207) is rescaled only when the
//        running max grows by more than 2^8.
//   O += P_t V_t: 4 k-steps x (N 64 over head dims 0:64, N 16 over 64:80).
// K-side tiles (K, one-hot) and V-side tiles have one buffer each but are needed in different phases: the K buffer is
// released as soon as S_t has executed, so tile t+1's K arrives under softmax(t) and PV_t, and V_{t+1} under S_{t+1} and
// softmax(t+1). Shared memory 64 KB (one 64-column bias atom: grids up to 32 x 32) or 88 KB (two: up to 64 x 64).
#pragma once
#include "attention_win_sm100.cuh"

namespace vfm {

constexpr int GLB_BLOCK_Q = 128;
constexpr int GLB_KT = 64;
constexpr int GLB_D = 80;
constexpr int GLB_THREADS = 192;                 // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2..5 softmax
constexpr int GLB_Q64 = GLB_BLOCK_Q * 128;       // [128 x 64] SWIZZLE_128B
constexpr int GLB_Q16 = GLB_BLOCK_Q * 32;        // [128 x 16] SWIZZLE_32B
constexpr int GLB_K64 = GLB_KT * 128;            // [64 x 64]
constexpr int GLB_K16 = GLB_KT * 32;             // [64 x 16]
constexpr uint32_t GLB_TMEM_COLS = 256;
constexpr uint32_t GLB_COL_S = 0, GLB_COL_O = 128;

template <int NA>
constexpr int glb_smem_bytes() { return GLB_Q64 * (1 + NA) + GLB_Q16 + GLB_K64 * (2 + NA) + 2 * GLB_K16 + 1024 + 128; }

struct GlobParams {
  int seq_len, heads, k_h, k_w;
  int bh, bw;                   // k_h, k_w rounded up to 16: column blocks of the bias operands
  int ld, g_col0;               // row pitch of the qkv buffer (elements); first table-term column
  float scale;
  const __nv_bfloat16* qkv;
  __nv_bfloat16* out;
};

template <int NA>
__global__ void __launch_bounds__(GLB_THREADS, 2)
attention_glob_kernel(const __grid_constant__ CUtensorMap tmap64, const __grid_constant__ CUtensorMap tmap16,
                      const __grid_constant__ CUtensorMap tmap_e, const GlobParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
  uint8_t* sQ0 = smem;                       // SWIZZLE_128B tiles first
  uint8_t* sQB = sQ0 + GLB_Q64;              // [NA][128 x 64] bias terms of the query rows: rel_h (bh cols) | rel_w (bw cols)
  uint8_t* sK0 = sQB + NA * GLB_Q64;
  uint8_t* sE = sK0 + GLB_K64;               // [NA][64 x 64] one-hot rows of this key tile
  uint8_t* sV0 = sE + NA * GLB_K64;
  uint8_t* sQ1 = sV0 + GLB_K64;              // SWIZZLE_32B tiles
  uint8_t* sK1 = sQ1 + GLB_Q16;
  uint8_t* sV1 = sK1 + GLB_K16;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV1 + GLB_K16);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;               // TMA -> MMA        (K, one-hot of tile t)
  uint64_t* k_empty = bars + 2;              // MMA -> TMA        (S_t executed)
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;              // MMA -> TMA        (PV_t executed)
  uint64_t* s_full = bars + 5;               // MMA -> softmax    (S_t, and every MMA before it, executed)
  uint64_t* p_full = bars + 6;               // softmax -> MMA
  uint64_t* o_full = bars + 7;               // MMA -> softmax    (last PV executed)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int C = p.heads * GLB_D;
  const int row0 = seq * p.seq_len;
  const int q0 = qt * GLB_BLOCK_Q;
  const int n_tiles = (p.seq_len + GLB_KT - 1) / GLB_KT;
  const int bias_steps = (p.bh + p.bw) >> 4;
  const int kc = C + head * GLB_D, vc = 2 * C + head * GLB_D;

  if (tid == 0) {
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(v_empty, 1);
    mbar_init(s_full, 1); mbar_init(p_full, 4); mbar_init(o_full, 1);
    fence_barrier_init();
    // Q and the first K / V tiles go out at once: their latency runs under the bias-operand build below
    mbar_arrive_expect_tx(q_full, GLB_Q64 + GLB_Q16);
    for (int h = 0; h < 2; ++h) {
      tma_load_2d(sQ0 + h * (GLB_Q64 / 2), &tmap64, q_full, head * GLB_D, row0 + q0 + 64 * h);
      tma_load_2d(sQ1 + h * (GLB_Q16 / 2), &tmap16, q_full, head * GLB_D + 64, row0 + q0 + 64 * h);
    }
    mbar_arrive_expect_tx(k_full, GLB_K64 * (1 + NA) + GLB_K16);
    tma_load_2d(sK0, &tmap64, k_full, kc, row0);
    tma_load_2d(sK1, &tmap16, k_full, kc + 64, row0);
    for (int a = 0; a < NA; ++a) tma_load_2d(sE + a * GLB_K64, &tmap_e, k_full, 64 * a, 0);
    mbar_arrive_expect_tx(v_full, GLB_K64 + GLB_K16);
    tma_load_2d(sV0, &tmap64, v_full, vc, row0);
    tma_load_2d(sV1, &tmap16, v_full, vc + 64, row0);
  }
  if (warp == 1) tmem_alloc<GLB_TMEM_COLS>(tmem_slot);
  // ---- bias terms of the query rows (built by all threads): one task = (row, 16-column chunk): 16 contiguous table terms
  // read backwards (all loads in flight together), pre-divided by the score scale, two 16-byte chunks of the swizzled row
  {
    const float inv_scale = 1.f / p.scale;
    const int Lh = 2 * p.k_h - 1, Lw = 2 * p.k_w - 1;
    const int chunks = (p.bh + p.bw) >> 4, h_chunks = p.bh >> 4;
    for (int t = tid; t < GLB_BLOCK_Q * chunks; t += GLB_THREADS) {
      const int r = t / chunks, cc = t - r * chunks;
      const int q = q0 + r;
      const bool is_w = cc >= h_chunks;
      const int k0 = (is_w ? cc - h_chunks : cc) * 16;   // first kh / kw of this chunk
      const int kn = is_w ? p.k_w : p.k_h;
      __nv_bfloat16 v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = __float2bfloat16(0.f);
      if (q < p.seq_len) {
        const int qh = q / p.k_w, qw = q - qh * p.k_w;
        const __nv_bfloat16* gp = p.qkv + static_cast<size_t>(row0 + q) * p.ld + p.g_col0 +
                                  (is_w ? p.heads * Lh + head * Lw + qw + p.k_w - 1 : head * Lh + qh + p.k_h - 1) - k0;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (k0 + c < kn) v[c] = gp[-c];
      }
      uint32_t w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c)
        w[c] = pack_bf16x2(__bfloat162float(v[2 * c]) * inv_scale, __bfloat162float(v[2 * c + 1]) * inv_scale);
      const int col = cc * 16;                             // column inside the [128 x 64 NA] operand
      uint8_t* rowp = sQB + (col >> 6) * GLB_Q64 + (r >> 3) * 1024 + (r & 7) * 128;
      const int ch = (col & 63) >> 3;                      // 16-byte chunk of the 128-byte row
      *reinterpret_cast<uint4*>(rowp + ((ch ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<uint4*>(rowp + (((ch + 1) ^ (r & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
    }
    // columns of the last atom past bh + bw are never multiplied (bias_steps stops before them)
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: tiles 1.. (tile 0 went out in the prologue) =====================
    for (int t = 1; t < n_tiles; ++t) {
      mbar_wait(k_empty, (t - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(k_full, GLB_K64 * (1 + NA) + GLB_K16);
        tma_load_2d(sK0, &tmap64, k_full, kc, row0 + t * GLB_KT);
        tma_load_2d(sK1, &tmap16, k_full, kc + 64, row0 + t * GLB_KT);
        for (int a = 0; a < NA; ++a) tma_load_2d(sE + a * GLB_K64, &tmap_e, k_full, 64 * a, t * GLB_KT);
      }
      __syncwarp();
      mbar_wait(v_empty, (t - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(v_full, GLB_K64 + GLB_K16);
        tma_load_2d(sV0, &tmap64, v_full, vc, row0 + t * GLB_KT);
        tma_load_2d(sV1, &tmap16, v_full, vc + 64, row0 + t * GLB_KT);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_s = make_idesc_bf16(GLB_BLOCK_Q, GLB_KT, 0, 0);
    constexpr uint32_t idesc_pv64 = make_idesc_bf16(GLB_BLOCK_Q, 64, 0, 1);   // B = V is MN-major
    constexpr uint32_t idesc_pv16 = make_idesc_bf16(GLB_BLOCK_Q, 16, 0, 1);
    const uint32_t tmem_s = tmem_base + GLB_COL_S, tmem_o = tmem_base + GLB_COL_O;
    const uint64_t dq0 = make_sw128_desc(smem_u32(sQ0)), dk0 = make_sw128_desc(smem_u32(sK0));
    const uint64_t dq1 = make_sw_desc(smem_u32(sQ1), 256, 6), dk1 = make_sw_desc(smem_u32(sK1), 256, 6);
    const uint64_t dqb = make_sw128_desc(smem_u32(sQB)), de = make_sw128_desc(smem_u32(sE));
    const uint64_t dv0 = make_sw128_desc(smem_u32(sV0)), dv1 = make_sw_desc(smem_u32(sV1), 256, 6);
    mbar_wait(q_full, 0);
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(k_full, t & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_ss(tmem_s, dq0 + 2 * k, dk0 + 2 * k, idesc_s, k != 0);   // head dims 0..63
        umma_ss(tmem_s, dq1, dk1, idesc_s, true);                                                  // head dims 64..79
        for (int b = 0; b < bias_steps; ++b) {   // atom b / 4 (64 columns each), k-step b % 4 inside it
          const uint64_t off_q = static_cast<uint64_t>((b >> 2) * (GLB_Q64 >> 4) + 2 * (b & 3));
          const uint64_t off_e = static_cast<uint64_t>((b >> 2) * (GLB_K64 >> 4) + 2 * (b & 3));
          umma_ss(tmem_s, dqb + off_q, de + off_e, idesc_s, true);
        }
        tc_commit(k_empty);   // the K-side buffer is free as soon as S_t has executed
        tc_commit(s_full);
      }
      __syncwarp();
      mbar_wait(v_full, t & 1);
      mbar_wait(p_full, t & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const int ksteps = (min(GLB_KT, p.seq_len - t * GLB_KT) + 15) >> 4;
        for (int k = 0; k < ksteps; ++k) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V (2048 B / 512 B)
          umma_ts(tmem_o, tmem_s + 8 * k, dv0 + 128 * k, idesc_pv64, (t | k) != 0);
          umma_ts(tmem_o + 64, tmem_s + 8 * k, dv1 + 32 * k, idesc_pv16, (t | k) != 0);
        }
        tc_commit(v_empty);
        if (t == n_tiles - 1) tc_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + output (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + GLB_COL_S;
    const uint32_t tmem_o = tmem_base + lane_base + GLB_COL_O;
    const float sc = p.scale * 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;   // log2 units: P stays <= 2^8 relative to m_ref
    float m_ref = -INFINITY, l = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(s_full, t & 1);                 // S_t is there, and PV_{t-1} has executed (commits are ordered)
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      const int valid = p.seq_len - t * GLB_KT;
      if (valid < GLB_KT) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;   // -inf: keys past the end
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[i]));
      const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * sc;
      const bool jump = m_tile > m_ref + (t == 0 ? 0.f : kRescaleThreshold);
      if (__any_sync(0xffffffffu, jump)) {
        const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;
        if (jump) { m_ref = m_tile; l *= alpha; }
        if (t > 0) {
#pragma unroll 1
          for (int c = 0; c < GLB_D / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem_o + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(tmem_o + c * 16, r);
          }
        }
      }
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const __nv_bfloat162 b = __floats2bfloat162_rn(fast_exp2(fmaf(__uint_as_float(s[2 * i]), sc, -m_ref)),
                                                       fast_exp2(fmaf(__uint_as_float(s[2 * i + 1]), sc, -m_ref)));
        l += __low2float(b) + __high2float(b);
        pk[i] = *reinterpret_cast<const uint32_t*>(&b);
      }
      tmem_st16(tmem_s, *reinterpret_cast<const uint32_t(*)[16]>(&pk[0]));       // P_t over the first 32 score columns
      tmem_st16(tmem_s + 16, *reinterpret_cast<const uint32_t(*)[16]>(&pk[16]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int q = q0 + row;
    uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(row0 + q) * C + head * GLB_D);
#pragma unroll 1
    for (int c = 0; c < GLB_D / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tmem_o + 16 * c, r);
      tmem_ld_wait();
      if (q < p.seq_len) {
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
    tmem_dealloc<GLB_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
