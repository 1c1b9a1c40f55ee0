// Fused multi-head attention (non-causal, head_dim 64) on tcgen05 for sm_100a.
//
// Replaces rein/models/backbones/dino_layers/attention.py:56-66 (reshape/permute of the fused
// qkv output, q*scale, q@k^T, softmax, @v, transpose back) — the N x N score tensor never
// leaves the SM. The 1/sqrt(d) scale is folded into W_q/b_q on the host (exact: 0.125 = 2^-3).
//
// Input : packed qkv activations [n_seq * seq_len, 3 * heads * 64] bf16, column =
//         which * (heads*64) + head * 64 + d  (exactly what qkv Linear emits, attention.py:58).
// Output: [n_seq * seq_len, heads * 64] bf16 (attention.py:66, "transpose(1,2).reshape(B,N,C)").
//
// One CTA = one (sequence, head, 128-query tile). Per 128-key tile:
//   S = Q K^T     tcgen05.mma SS, fp32 S in TMEM cols [0,128)
//   softmax       4 warps, one query row per thread (TMEM lane); online max/sum in registers;
//                 P written back to TMEM as packed bf16 (cols [128,192))
//   O_j = P V     tcgen05.mma TS (A = P from TMEM, B = V tile MN-major), fp32 in cols [192,256)
//   O += O_j      running output in registers, rescaled by exp2(m_old - m_new)
// 256 TMEM columns and ~82 KB smem per CTA, so two CTAs share an SM and overlap each other's
// MMA and softmax phases.
//
// Roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator,
// warps 2..5 = softmax/output (TMEM lane quadrant = warp % 4).
#pragma once
#include "sm100_ptx.cuh"

namespace vfm {

constexpr int ATT_BLOCK_Q = 128;
constexpr int ATT_BLOCK_KV = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = ATT_BLOCK_KV * ATT_D * 2;  // 16 KB (Q, K and V tiles alike)
constexpr int ATT_SMEM_BYTES = (1 + 2 * ATT_KV_STAGES) * ATT_TILE_BYTES + 1024 + 256;
constexpr uint32_t ATT_TMEM_COLS = 256;
constexpr uint32_t ATT_COL_S = 0, ATT_COL_P = 128, ATT_COL_O = 192;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                     int seq_len, int heads, int q_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + ATT_TILE_BYTES;
  uint8_t* smem_v = smem + (1 + ATT_KV_STAGES) * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (1 + 2 * ATT_KV_STAGES) * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                          // TMA -> MMA
  uint64_t* kv_full = bars + 1;                     // [stages] TMA -> MMA
  uint64_t* kv_empty = bars + 1 + ATT_KV_STAGES;    // [stages] MMA (PV done) -> TMA
  uint64_t* s_full = bars + 1 + 2 * ATT_KV_STAGES;  // MMA -> softmax   (S_j ready)
  uint64_t* p_full = s_full + 1;                    // softmax -> MMA   (P_j written, O_{j-1} drained)
  uint64_t* o_full = s_full + 2;                    // MMA -> softmax   (O_j = P_j V_j ready)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work unit
  const int unit = blockIdx.x;
  const int qt = unit % q_tiles;
  const int head = (unit / q_tiles) % heads;
  const int seq = unit / (q_tiles * heads);
  const int C = heads * ATT_D;
  const int row0 = seq * seq_len;  // first row of this sequence in the packed qkv matrix
  const int kv_tiles = (seq_len + ATT_BLOCK_KV - 1) / ATT_BLOCK_KV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(smem_q, &tmap_qkv, q_full, head * ATT_D, row0 + qt * ATT_BLOCK_Q);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < kv_tiles; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_arrive_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
        tma_load_2d(smem_k + stage * ATT_TILE_BYTES, &tmap_qkv, &kv_full[stage], C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
        tma_load_2d(smem_v + stage * ATT_TILE_BYTES, &tmap_qkv, &kv_full[stage], 2 * C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // S: M=128 (queries), N=kv width, K-major A and B.  PV: M=128, N=64 (d), B = V is MN-major.
      constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK_Q, ATT_D, 0, 1);
      const uint32_t tmem_s = tmem_base + ATT_COL_S, tmem_p = tmem_base + ATT_COL_P, tmem_o = tmem_base + ATT_COL_O;
      auto kv_width = [&](int j) {  // keys in tile j, rounded up to a multiple of 32 (masked in softmax)
        int w = seq_len - j * ATT_BLOCK_KV;
        w = w > ATT_BLOCK_KV ? ATT_BLOCK_KV : w;
        return (w + 31) & ~31;
      };
      auto issue_s = [&](int j, int stage) {
        const uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK_Q, kv_width(j), 0, 0);
        const uint64_t dq = make_sw128_desc(smem_u32(smem_q));
        const uint64_t dk = make_sw128_desc(smem_u32(smem_k + stage * ATT_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(s_full);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&kv_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      int stage = 0; uint32_t phase = 0;
      for (int j = 0; j < kv_tiles; ++j) {
        // P_j is in TMEM (and O_{j-1} has been drained): O_j = P_j V_j
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dv = make_sw128_desc(smem_u32(smem_v + stage * ATT_TILE_BYTES));
        const int ksteps = kv_width(j) / 16;
        for (int k = 0; k < ksteps; ++k) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
          umma_ts(tmem_o, tmem_p + 8 * k, dv + 128 * k, idesc_pv, k != 0);
        }
        tc_commit(&kv_empty[stage]);  // K_j and V_j are free once PV_j (and S_j before it) finished
        tc_commit(o_full);
        if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
        if (j + 1 < kv_tiles) {
          mbar_wait(&kv_full[stage], phase);
          tc_fence_after();
          issue_s(j + 1, stage);
        }
      }
    }
  } else {
    // ===================== softmax + output (warps 2..5) =====================
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + ATT_COL_S;
    const uint32_t tmem_p = tmem_base + lane_base + ATT_COL_P;
    const uint32_t tmem_o = tmem_base + lane_base + ATT_COL_O;
    constexpr float kLog2e = 1.4426950408889634f;

    float o[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;

    for (int j = 0; j < kv_tiles; ++j) {
      int valid = seq_len - j * ATT_BLOCK_KV;
      valid = valid > ATT_BLOCK_KV ? ATT_BLOCK_KV : valid;
      const int chunks = (valid + 31) >> 5;

      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // pass 1: row max of this tile
      float m_tile = -INFINITY;
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(r[i]);
          if (c * 32 + i < valid) m_tile = fmaxf(m_tile, s);
        }
      }
      // drain O_{j-1} before P_j may trigger the MMA that overwrites it
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < ATT_D / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_o + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(r[i]);
        }
      }
      const float m_new = fmaxf(m_run, m_tile);  // finite: every tile has >= 1 valid key
      const float alpha = fast_exp2((m_run - m_new) * kLog2e);
      m_run = m_new;
      l_run *= alpha;
      if (alpha != 1.f) {
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o[i] *= alpha;
      }
      // pass 2: p = exp(s - m), row sum, pack to bf16, store as the A operand of the PV MMA
      const float mb = m_new * kLog2e;
      float l_tile = 0.f;
      for (int c = 0; c < chunks; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = fast_exp2(fmaf(__uint_as_float(r[2 * i]), kLog2e, -mb));
          float p1 = fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), kLog2e, -mb));
          if (c * 32 + 2 * i >= valid) p0 = 0.f;
          if (c * 32 + 2 * i + 1 >= valid) p1 = 0.f;
          l_tile += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st16(tmem_p + c * 16, pk);
      }
      l_run += l_tile;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // last tile's O
    mbar_wait(o_full, (kv_tiles - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < ATT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_o + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(r[i]);
    }
    const int q_idx = qt * ATT_BLOCK_Q + quad * 32 + lane;
    if (q_idx < seq_len) {
      const float inv = 1.f / l_run;
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + q_idx) * C + head * ATT_D);
#pragma unroll
      for (int i = 0; i < ATT_D / 8; ++i)
        dst[i] = make_uint4(pack_bf16x2(o[8 * i] * inv, o[8 * i + 1] * inv), pack_bf16x2(o[8 * i + 2] * inv, o[8 * i + 3] * inv),
                            pack_bf16x2(o[8 * i + 4] * inv, o[8 * i + 5] * inv), pack_bf16x2(o[8 * i + 6] * inv, o[8 * i + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
