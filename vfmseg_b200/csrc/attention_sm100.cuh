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
//   O += P V      tcgen05.mma TS (A = P from TMEM, B = V tile MN-major), fp32 accumulator in cols [192,256);
//                 O stays in TMEM for the whole KV loop and is rescaled lazily (only when the running
//                 max grows by more than 2^8), so the steady-state tile costs no O traffic at all
//   The score MMA of tile j+1 is issued as soon as every softmax thread holds S_j in registers, so
//   it runs underneath the exponentials of tile j.
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
  uint64_t* kv_empty = bars + 1 + ATT_KV_STAGES;    // [stages] MMA (PV_j done) -> TMA
  uint64_t* s_full = bars + 1 + 2 * ATT_KV_STAGES;  // MMA -> softmax   (S_j in TMEM)
  uint64_t* s_free = s_full + 1;                    // softmax -> MMA   (S_j copied to registers)
  uint64_t* p_full = s_full + 2;                    // softmax -> MMA   (P_j in TMEM, O rescaled if needed)
  uint64_t* pv_done = s_full + 3;                   // MMA -> softmax   (O += P_j V_j finished)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int unit = blockIdx.x;
  const int qt = unit % q_tiles;
  const int head = (unit / q_tiles) % heads;
  const int seq = unit / (q_tiles * heads);
  const int C = heads * ATT_D;
  const int row0 = seq * seq_len;
  const int kv_tiles = (seq_len + ATT_BLOCK_KV - 1) / ATT_BLOCK_KV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_qkv);
    mbar_init(q_full, 1);
    for (int s = 0; s < ATT_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp converged, one elected lane issues) =====================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
      tma_load_2d(smem_q, &tmap_qkv, q_full, head * ATT_D, row0 + qt * ATT_BLOCK_Q);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int j = 0; j < kv_tiles; ++j) {
      mbar_wait(&kv_empty[stage], phase ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&kv_full[stage], 2 * ATT_TILE_BYTES);
        tma_load_2d(smem_k + stage * ATT_TILE_BYTES, &tmap_qkv, &kv_full[stage], C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
        tma_load_2d(smem_v + stage * ATT_TILE_BYTES, &tmap_qkv, &kv_full[stage], 2 * C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
      }
      __syncwarp();
      if (++stage == ATT_KV_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp converged, one elected lane issues) =====================
    // Issue order: S_0, [S_1, PV_0], [S_2, PV_1], ... so the score MMA of tile j+1 runs while the softmax
    // warps exponentiate tile j. A commit covers every earlier MMA, so s_full(j+1) also implies PV_{j-1} done.
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK_Q, ATT_D, 0, 1);   // B = V is MN-major
    const uint32_t tmem_s = tmem_base + ATT_COL_S, tmem_p = tmem_base + ATT_COL_P, tmem_o = tmem_base + ATT_COL_O;
    const uint64_t dq = make_sw128_desc(smem_u32(smem_q));
    const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
    const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
    auto kv_width = [&](int j) {  // keys in tile j rounded up to 32 (the excess is masked by the softmax)
      int w = seq_len - j * ATT_BLOCK_KV;
      w = w > ATT_BLOCK_KV ? ATT_BLOCK_KV : w;
      return (w + 31) & ~31;
    };
    auto issue_s = [&](int j, int stage) {   // called by the elected lane only
      const uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK_Q, kv_width(j), 0, 0);
      const uint64_t dk = dk0 + static_cast<uint64_t>(stage * (ATT_TILE_BYTES >> 4));
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      tc_commit(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&kv_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) issue_s(0, 0);
    __syncwarp();
    for (int j = 0; j < kv_tiles; ++j) {
      const int stage = j % ATT_KV_STAGES;
      if (j + 1 < kv_tiles) {
        const int nstage = (j + 1) % ATT_KV_STAGES;
        mbar_wait(&kv_full[nstage], ((j + 1) / ATT_KV_STAGES) & 1);
        mbar_wait(s_free, j & 1);          // every softmax thread holds S_j in registers
        tc_fence_after();
        if (elect_one_sync()) issue_s(j + 1, nstage);
        __syncwarp();
      }
      mbar_wait(p_full, j & 1);            // P_j stored (and O rescaled when the running max jumped)
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dv = dv0 + static_cast<uint64_t>(stage * (ATT_TILE_BYTES >> 4));
        const int ksteps = kv_width(j) / 16;
        for (int k = 0; k < ksteps; ++k) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
          umma_ts(tmem_o, tmem_p + 8 * k, dv + 128 * k, idesc_pv, (j | k) != 0);
        }
        tc_commit(&kv_empty[stage]);
        tc_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax + output (warps 2..5) =====================
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + ATT_COL_S;
    const uint32_t tmem_p = tmem_base + lane_base + ATT_COL_P;
    const uint32_t tmem_o = tmem_base + lane_base + ATT_COL_O;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;   // in log2 units: P stays <= 2^8 relative to the reference max

    // m_ref: the max (times log2e) that P and the O accumulator in TMEM are currently relative to.
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < kv_tiles; ++j) {
      int valid = seq_len - j * ATT_BLOCK_KV;
      valid = valid > ATT_BLOCK_KV ? ATT_BLOCK_KV : valid;
      const int chunks = (valid + 31) >> 5;   // warp-uniform

      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t s[128];
      if (chunks > 0) tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      if (chunks > 1) tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      if (chunks > 2) tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
      if (chunks > 3) tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                   // the tensor core may overwrite S with tile j+1 now

      if (valid < ATT_BLOCK_KV) {            // tail tile only: mask keys past the sequence end
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      float m_tile = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; ++i) m_tile = fmaxf(m_tile, __uint_as_float(s[i]));
      m_tile *= kLog2e;

      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);     // O and the P buffer are quiescent
        tc_fence_after();
        const bool jump = m_tile > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, jump)) { // rare after the first tiles: rescale O in TMEM
          const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;
          if (jump) { m_ref = m_tile; l_run *= alpha; }
#pragma unroll 1
          for (int c = 0; c < ATT_D / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem_o + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st16(tmem_o + c * 16, r);
          }
        }
      } else {
        m_ref = m_tile;
      }

      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < chunks) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(s[c * 32 + 2 * i]), kLog2e, -m_ref));
            const float p1 = fast_exp2(fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), kLog2e, -m_ref));
            l_tile += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st16(tmem_p + c * 16, pk);
        }
      }
      l_run += l_tile;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    mbar_wait(pv_done, (kv_tiles - 1) & 1);
    tc_fence_after();
    const int q_idx = qt * ATT_BLOCK_Q + quad * 32 + lane;
    const float inv = 1.f / l_run;
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + q_idx) * C + head * ATT_D);
#pragma unroll
    for (int c = 0; c < ATT_D / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_o + c * 32, r);
      tmem_ld_wait();
      if (q_idx < seq_len) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float v[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[8 * i + t]) * inv;
          dst[c * 4 + i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
      }
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
