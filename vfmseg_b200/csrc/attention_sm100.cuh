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
// One CTA = one (sequence, head, 128-query tile); 64-key tiles. Per key tile j:
//   S_j = Q K_j^T   tcgen05.mma SS, fp32 scores in TMEM buffer j % 2 (2 x 64 columns) — issued two tiles ahead
//   softmax         4 warps, one query row per thread (one TMEM lane): the row is copied to 64 registers, the S
//                   buffer released, running max kept in registers, P_j = exp2(s - m) truncated to bf16 and
//                   written to TMEM buffer j % 2 (2 x 32 columns)
//   O += P_j V_j    tcgen05.mma TS (A = P from TMEM, B = V tile as an MN-major SWIZZLE_128B operand); O (64 fp32
//                   columns) stays in TMEM for the whole KV loop and is rescaled lazily, only when a row's
//                   running max grows by more than 2^8
// Because S and P are both double-buffered, the exponentials of tile j+1 never wait for the PV MMA of tile j
// (with single buffers the two serialised: ~1700 clk of MMA round trip + ~1650 clk of MUFU per tile,
// profiles/r1_attention_timeline.txt). 256 TMEM columns and ~82 KB smem per CTA -> two CTAs per SM.
// The last key tile of a 1025-token sequence holds one key: it runs a 32-wide MMA and masks 31 columns.
//
// Roles (192 threads): warp 0 = TMA producer (K ring 4 x 8 KB released right after S_j, V ring 4 x 8 KB),
// warp 1 = MMA issuer + TMEM allocator, warps 2..5 = softmax/output (TMEM lane quadrant = warp % 4).
#pragma once
#include <type_traits>

#include "sm100_ptx.cuh"

namespace vfm {

constexpr int ATT_BLOCK_Q = 128;
constexpr int ATT_BLOCK_KV = 64;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;           // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2..5 softmax
constexpr int ATT_K_STAGES = 4;            // K_j is released as soon as S_j = Q K_j^T has executed
constexpr int ATT_V_STAGES = 4;            // V_j is held until O += P_j V_j has executed
constexpr int ATT_Q_BYTES = ATT_BLOCK_Q * ATT_D * 2;      // 16 KB
constexpr int ATT_KV_BYTES = ATT_BLOCK_KV * ATT_D * 2;    //  8 KB
constexpr int ATT_SMEM_BYTES = ATT_Q_BYTES + (ATT_K_STAGES + ATT_V_STAGES) * ATT_KV_BYTES + 1024 + 256;
constexpr uint32_t ATT_TMEM_COLS = 256;
// TMEM columns: S0 [0,64) S1 [64,128) fp32 scores, P0 [128,160) P1 [160,192) packed bf16 probabilities, O [192,256)
constexpr uint32_t ATT_COL_S = 0, ATT_COL_P = 128, ATT_COL_O = 192;
constexpr int kPolyExpEvery = 0;   // 0: all exponentials on MUFU; n: every n-th pair computes its odd element with poly_exp2

#ifdef VFM_EPI_TIMING
__device__ long long g_att_trace[16][12];
#define ATT_TRACE(j, ev) do { if (blockIdx.x == 300 && lane == 0 && (j) < 16) g_att_trace[j][ev] = clock64(); } while (0)
__device__ unsigned long long g_att_dbg[8];
#else
#define ATT_TRACE(j, ev)
#endif

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                     int seq_len, int heads, int q_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + ATT_Q_BYTES;
  uint8_t* smem_v = smem_k + ATT_K_STAGES * ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + ATT_V_STAGES * ATT_KV_BYTES);
  uint64_t* q_full = bars;                          // TMA -> MMA
  uint64_t* k_full = bars + 1;                      // [K stages] TMA -> MMA
  uint64_t* k_empty = k_full + ATT_K_STAGES;        // [K stages] MMA (S_j executed) -> TMA
  uint64_t* v_full = k_empty + ATT_K_STAGES;        // [V stages] TMA -> MMA
  uint64_t* v_empty = v_full + ATT_V_STAGES;        // [V stages] MMA (PV_j executed) -> TMA
  uint64_t* s_full = v_empty + ATT_V_STAGES;        // [2] MMA -> softmax   (S_j in TMEM buffer j % 2)
  uint64_t* s_free = s_full + 2;                    // [2] unused (p_full implies the S buffer was read)
  uint64_t* p_full = s_full + 4;                    // [2] softmax -> MMA   (P_j in TMEM buffer j % 2)
  uint64_t* pv_done = s_full + 6;                   // [2] MMA -> softmax   (O += P_j V_j executed; frees P buffer j % 2)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 8);

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
    for (int s = 0; s < ATT_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < ATT_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 1); mbar_init(&p_full[s], 128); mbar_init(&pv_done[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (warp converged, one elected lane issues) =====================
    if (elect_one_sync()) {   // Q tile = two 64-row boxes
      mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_2d(smem_q, &tmap_qkv, q_full, head * ATT_D, row0 + qt * ATT_BLOCK_Q);
      tma_load_2d(smem_q + ATT_KV_BYTES, &tmap_qkv, q_full, head * ATT_D, row0 + qt * ATT_BLOCK_Q + 64);
    }
    __syncwarp();
    auto load_k = [&](int j) {
      const int st = j % ATT_K_STAGES;
      mbar_wait(&k_empty[st], ((j / ATT_K_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[st], ATT_KV_BYTES);
        tma_load_2d(smem_k + st * ATT_KV_BYTES, &tmap_qkv, &k_full[st], C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
      }
      __syncwarp();
    };
    auto load_v = [&](int j) {
      const int st = j % ATT_V_STAGES;
      mbar_wait(&v_empty[st], ((j / ATT_V_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[st], ATT_KV_BYTES);
        tma_load_2d(smem_v + st * ATT_KV_BYTES, &tmap_qkv, &v_full[st], 2 * C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
      }
      __syncwarp();
    };
    // load order follows the order of use: K_0, K_1, K_2, V_0, K_3, V_1, ...
    load_k(0);
    if (kv_tiles > 1) load_k(1);
    for (int j = 0; j < kv_tiles; ++j) {
      if (j + 2 < kv_tiles) load_k(j + 2);
      load_v(j);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp converged, one elected lane issues) =====================
    // S and P are double-buffered, so neither the exponentials of tile j+1 nor the PV MMA of tile j ever wait for
    // each other. Issue order:  S_0, S_1, [PV_0, S_2], [PV_1, S_3], ...
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK_Q, ATT_D, 0, 1);   // B = V is MN-major
    const uint32_t tmem_o = tmem_base + ATT_COL_O;
    const uint64_t dq = make_sw128_desc(smem_u32(smem_q));
    const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
    const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
    auto kv_width = [&](int j) {  // keys in tile j rounded up to 32 (the excess is masked by the softmax)
      int w = seq_len - j * ATT_BLOCK_KV;
      w = w > ATT_BLOCK_KV ? ATT_BLOCK_KV : w;
      return (w + 31) & ~31;
    };
    auto issue_s = [&](int j) {   // whole warp; waits for K_j, then the elected lane issues
      const int st = j % ATT_K_STAGES;
      mbar_wait(&k_full[st], (j / ATT_K_STAGES) & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK_Q, kv_width(j), 0, 0);
        const uint64_t dk = dk0 + static_cast<uint64_t>(st * (ATT_KV_BYTES >> 4));
        const uint32_t tmem_s = tmem_base + ATT_COL_S + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(&k_empty[st]);       // K_j can be overwritten as soon as S_j has executed
        tc_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    if (kv_tiles > 1) issue_s(1);
    for (int j = 0; j < kv_tiles; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int st = j % ATT_V_STAGES;
      mbar_wait(&v_full[st], (j / ATT_V_STAGES) & 1);
      mbar_wait(&p_full[b], ph);         // P_j stored (and O rescaled when the running max jumped); S_j is in registers
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t dv = dv0 + static_cast<uint64_t>(st * (ATT_KV_BYTES >> 4));
        const uint32_t tmem_p = tmem_base + ATT_COL_P + b * 32;
        const int ksteps = kv_width(j) / 16;
        for (int k = 0; k < ksteps; ++k) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
          umma_ts(tmem_o, tmem_p + 8 * k, dv + 128 * k, idesc_pv, (j | k) != 0);
        }
        tc_commit(&v_empty[st]);
        tc_commit(&pv_done[b]);
      }
      __syncwarp();
      // S_{j+2} reuses S buffer b (its previous contents are in registers since before p_full(j)). Its commit also
      // covers PV_j, so s_full(j+2) tells the softmax that P buffer b is free again: one wait per tile, not two.
      if (j + 2 < kv_tiles) issue_s(j + 2);
    }
  } else {
    // ===================== softmax + output (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_o = tmem_base + lane_base + ATT_COL_O;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;   // in log2 units: P stays <= 2^8 relative to the reference max

    // m_ref: the max (times log2e) that P and the O accumulator in TMEM are currently relative to.
    float m_ref = -INFINITY, l_run = 0.f;

    for (int j = 0; j < kv_tiles; ++j) {
      const int b = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      int valid = seq_len - j * ATT_BLOCK_KV;
      valid = valid > ATT_BLOCK_KV ? ATT_BLOCK_KV : valid;
      const int chunks = (valid + 31) >> 5;   // warp-uniform: 1 or 2
      const uint32_t tmem_s = tmem_base + lane_base + ATT_COL_S + b * 64;
      const uint32_t tmem_p = tmem_base + lane_base + ATT_COL_P + b * 32;

      if (warp == 2) ATT_TRACE(j, 3);
      mbar_wait(&s_full[b], ph);
      tc_fence_after();
      if (warp == 2) ATT_TRACE(j, 4);
      uint32_t s[64];
      tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      if (chunks > 1) tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      if (warp == 2) ATT_TRACE(j, 5);

      if (valid < ATT_BLOCK_KV) {            // tail tile only: mask keys past the sequence end
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
      for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[i]));
      const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kLog2e;
      if (warp == 2) ATT_TRACE(j, 6);

      if (j == 0) {
        m_ref = m_tile;
      } else {
        const bool jump = m_tile > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, jump)) { // rare after the first tiles: rescale O in TMEM
          // every PV up to tile j-1 must have executed before O is touched (pv_done[(j-1)%2] is the newest phase
          // of that barrier, so waiting on it now and again later as a P-buffer guard is harmless)
          mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
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
      }
      // P buffer b was last read by PV_{j-2}, which executed before S_j (see the MMA issue order): no wait needed
      if (warp == 2) ATT_TRACE(j, 7);

      float l_tile = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c < chunks) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x0 = fmaf(__uint_as_float(s[c * 32 + 2 * i]), kLog2e, -m_ref);
            const float x1 = fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), kLog2e, -m_ref);
            // P is truncated to bf16 with integer ops (LOP3/PRMT) and the row sum is taken over the truncated
            // values, so P / l is an exact softmax of the weights the tensor core actually uses (per-weight
            // relative perturbation <= 2^-8, no bias after normalisation); saves the F2FP conversions.
            const uint32_t b0 = __float_as_uint(fast_exp2(x0)) & 0xffff0000u;
            const float e1 = (kPolyExpEvery != 0 && valid == ATT_BLOCK_KV && (i % (kPolyExpEvery ? kPolyExpEvery : 1)) == 0) ? poly_exp2(x1) : fast_exp2(x1);
            const uint32_t b1 = __float_as_uint(e1) & 0xffff0000u;
            l_tile += __uint_as_float(b0) + __uint_as_float(b1);
            pk[i] = __byte_perm(b0, b1, 0x7632);
          }
          tmem_st16(tmem_p + c * 16, pk);
        }
      }
      l_run += l_tile;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[b]);
      if (warp == 2) ATT_TRACE(j, 8);
    }
    // the last PV implies all earlier ones (commits are ordered)
    mbar_wait(&pv_done[(kv_tiles - 1) & 1], ((kv_tiles - 1) >> 1) & 1);
    tc_fence_after();
    const int q_idx = qt * ATT_BLOCK_Q + row;
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
