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
// Roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator,
// warps 2..9 = softmax/output (TMEM lane quadrant = warp % 4, key half = (warp - 2) / 4).
#pragma once
#include <type_traits>

#include "sm100_ptx.cuh"

namespace vfm {

constexpr int ATT_BLOCK_Q = 128;
constexpr int ATT_BLOCK_KV = 128;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 320;           // warp 0 TMA, warp 1 MMA, warps 2..9 softmax
constexpr int ATT_SOFTMAX_THREADS = 256;
constexpr int ATT_K_STAGES = 3;    // K_j is released as soon as S_j = Q K_j^T has executed
constexpr int ATT_V_STAGES = 2;    // V_j is held until O += P_j V_j has executed
constexpr int ATT_TILE_BYTES = ATT_BLOCK_KV * ATT_D * 2;  // 16 KB (Q, K and V tiles alike)
constexpr int ATT_NUM_TILES = 1 + ATT_K_STAGES + ATT_V_STAGES;
constexpr int ATT_XCHG_BYTES = 3 * 256 * 4;   // row-max exchange (2 tile parities x 2 halves x 128 rows) + row-sum exchange
constexpr int ATT_SMEM_BYTES = ATT_NUM_TILES * ATT_TILE_BYTES + 1024 + 256 + ATT_XCHG_BYTES;
constexpr uint32_t ATT_TMEM_COLS = 256;
constexpr uint32_t ATT_COL_S = 0, ATT_COL_P = 128, ATT_COL_O = 192;
constexpr int kPolyExpEvery = 0;   // 0: all exponentials on MUFU; n: every n-th pair computes its odd element with poly_exp2

#ifdef VFM_EPI_TIMING
__device__ long long g_att_trace[16][12];   // [tile][event] clock64 of CTA 300 (a steady-state CTA sharing its SM)
#define ATT_TRACE(j, ev) do { if (blockIdx.x == 300 && lane == 0 && (j) < 16) g_att_trace[j][ev] = clock64(); } while (0)
__device__ unsigned long long g_att_dbg[8];
#define ATT_TICK(var) const long long var = clock64()
#define ATT_ACC(i, a, b) adbg[i] += static_cast<unsigned long long>((b) - (a))
#else
#define ATT_TICK(var)
#define ATT_ACC(i, a, b)
#define ATT_TRACE(j, ev)
#endif

__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_qkv, __nv_bfloat16* __restrict__ out,
                     int seq_len, int heads, int q_tiles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + ATT_TILE_BYTES;
  uint8_t* smem_v = smem + (1 + ATT_K_STAGES) * ATT_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_NUM_TILES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;                          // TMA -> MMA
  uint64_t* k_full = bars + 1;                      // [K stages] TMA -> MMA
  uint64_t* k_empty = k_full + ATT_K_STAGES;        // [K stages] MMA (S_j done) -> TMA
  uint64_t* v_full = k_empty + ATT_K_STAGES;        // [V stages] TMA -> MMA
  uint64_t* v_empty = v_full + ATT_V_STAGES;        // [V stages] MMA (PV_j done) -> TMA
  uint64_t* s_full = v_empty + ATT_V_STAGES;        // MMA -> softmax   (S_j in TMEM)
  uint64_t* s_free = s_full + 1;                    // softmax -> MMA   (S_j copied to registers)
  uint64_t* p_full = s_full + 2;                    // softmax -> MMA   (P_j in TMEM, O rescaled if needed)
  uint64_t* pv_done = s_full + 3;                   // MMA -> softmax   (O += P_j V_j finished)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 4);
  float* max_x = reinterpret_cast<float*>(smem + ATT_NUM_TILES * ATT_TILE_BYTES + 256);

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
    mbar_init(s_full, 1);
    mbar_init(s_free, ATT_SOFTMAX_THREADS);
    mbar_init(p_full, ATT_SOFTMAX_THREADS);
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
    // load order follows the order of use: K_0, K_1, V_0, K_2, V_1, ...
    auto load_k = [&](int j) {
      const int st = j % ATT_K_STAGES;
      mbar_wait(&k_empty[st], ((j / ATT_K_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[st], ATT_TILE_BYTES);
        tma_load_2d(smem_k + st * ATT_TILE_BYTES, &tmap_qkv, &k_full[st], C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
      }
      __syncwarp();
    };
    auto load_v = [&](int j) {
      const int st = j % ATT_V_STAGES;
      mbar_wait(&v_empty[st], ((j / ATT_V_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[st], ATT_TILE_BYTES);
        tma_load_2d(smem_v + st * ATT_TILE_BYTES, &tmap_qkv, &v_full[st], 2 * C + head * ATT_D, row0 + j * ATT_BLOCK_KV);
      }
      __syncwarp();
    };
    load_k(0);
    for (int j = 0; j < kv_tiles; ++j) {
      if (j + 1 < kv_tiles) load_k(j + 1);
      load_v(j);
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
    auto issue_s = [&](int j) {   // called by the elected lane only
      const int st = j % ATT_K_STAGES;
      const uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK_Q, kv_width(j), 0, 0);
      const uint64_t dk = dk0 + static_cast<uint64_t>(st * (ATT_TILE_BYTES >> 4));
#pragma unroll
      for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
      tc_commit(&k_empty[st]);   // K_j can be overwritten as soon as S_j has executed
      tc_commit(s_full);
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one_sync()) issue_s(0);
    __syncwarp();
    for (int j = 0; j < kv_tiles; ++j) {
      const int stage = j % ATT_V_STAGES;
      if (j + 1 < kv_tiles) {
        mbar_wait(&k_full[(j + 1) % ATT_K_STAGES], ((j + 1) / ATT_K_STAGES) & 1);
        mbar_wait(s_free, j & 1);          // every softmax thread holds S_j in registers
        tc_fence_after();
        if (elect_one_sync()) issue_s(j + 1);
        __syncwarp();
        ATT_TRACE(j + 1, 0);
      }
      mbar_wait(&v_full[stage], (j / ATT_V_STAGES) & 1);
      mbar_wait(p_full, j & 1);            // P_j stored (and O rescaled when the running max jumped)
      tc_fence_after();
      ATT_TRACE(j, 1);
      if (elect_one_sync()) {
        const uint64_t dv = dv0 + static_cast<uint64_t>(stage * (ATT_TILE_BYTES >> 4));
        const int ksteps = kv_width(j) / 16;
        for (int k = 0; k < ksteps; ++k) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
          umma_ts(tmem_o, tmem_p + 8 * k, dv + 128 * k, idesc_pv, (j | k) != 0);
        }
        tc_commit(&v_empty[stage]);
        tc_commit(pv_done);
      }
      __syncwarp();
      ATT_TRACE(j, 2);
    }
  } else {
    // ===================== softmax + output (warps 2..9) =====================
    // Two warps per TMEM lane quadrant: warp pair (q, q+4) shares 32 query rows, each thread owns one row and
    // one 64-key half of the tile (64 live score registers instead of 128 -> room for the exponentials to overlap,
    // and four softmax warps per scheduler with two CTAs per SM). The pair exchanges its partial row max through
    // shared memory once per tile and its partial row sum once at the end.
    const int sw = warp - 2;
    const int quad = warp & 3;
    const int half = sw >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + ATT_COL_S + half * 64;
    const uint32_t tmem_p = tmem_base + lane_base + ATT_COL_P + half * 32;
    const uint32_t tmem_o = tmem_base + lane_base + ATT_COL_O + half * 32;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;   // in log2 units: P stays <= 2^8 relative to the reference max

    // m_ref: the max (times log2e) that P and the O accumulator in TMEM are currently relative to.
    float m_ref = -INFINITY, l_run = 0.f;
#ifdef VFM_EPI_TIMING
    unsigned long long adbg[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = clock64();
#endif

    for (int j = 0; j < kv_tiles; ++j) {
      int valid = seq_len - j * ATT_BLOCK_KV - half * 64;   // valid keys in this thread's 64-key half
      valid = valid > 64 ? 64 : (valid < 0 ? 0 : valid);
      const int chunks = (valid + 31) >> 5;                 // warp-uniform: 0, 1 or 2

      ATT_TICK(a0);
      if (warp == 2) ATT_TRACE(j, 3);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (warp == 2) ATT_TRACE(j, 4);
      ATT_TICK(a1);
      ATT_ACC(0, a0, a1);
      uint32_t s[64];
      if (chunks > 0) tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      if (chunks > 1) tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                   // the tensor core may overwrite S with tile j+1 now
      ATT_TICK(a2);
      ATT_ACC(1, a1, a2);
      if (warp == 2) ATT_TRACE(j, 5);

      if (valid < 64) {                      // tail tile only: mask keys past the sequence end
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;  // -inf
      }
      float m_part = -INFINITY;
#pragma unroll
      for (int i = 0; i < 64; ++i) m_part = fmaxf(m_part, __uint_as_float(s[i]));
      float* mx = max_x + (j & 1) * 256;
      mx[half * 128 + row] = m_part;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");   // the two warps of this quadrant
      const float m_tile = fmaxf(m_part, mx[(half ^ 1) * 128 + row]) * kLog2e;

      ATT_TICK(a3);
      ATT_ACC(2, a2, a3);
      if (warp == 2) ATT_TRACE(j, 6);
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);     // O and the P buffer are quiescent
        tc_fence_after();
        const bool jump = m_tile > m_ref + kRescaleThreshold;   // identical in both threads of a row
        if (__any_sync(0xffffffffu, jump)) { // rare after the first tiles: rescale this thread's 32 O columns
          const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;
          if (jump) { m_ref = m_tile; l_run *= alpha; }
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
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

      ATT_TICK(a4);
      ATT_ACC(3, a3, a4);
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
            const uint32_t b1 = __float_as_uint(fast_exp2(x1)) & 0xffff0000u;
            l_tile += __uint_as_float(b0) + __uint_as_float(b1);
            pk[i] = __byte_perm(b0, b1, 0x7632);
          }
          tmem_st16(tmem_p + c * 16, pk);
        }
      }
      l_run += l_tile;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      if (warp == 2) ATT_TRACE(j, 8);
      ATT_TICK(a5);
      ATT_ACC(4, a4, a5);
    }
#ifdef VFM_EPI_TIMING
    adbg[5] = static_cast<unsigned long long>(clock64() - t_begin);
    if (lane == 0 && warp == 2) { for (int i = 0; i < 6; ++i) atomicAdd(&g_att_dbg[i], adbg[i]); atomicAdd(&g_att_dbg[6], 1ull); }
#endif
    // total row sum = sum of the two halves
    float* lx = max_x + 512;
    lx[half * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
    const float inv = 1.f / (l_run + lx[(half ^ 1) * 128 + row]);
    mbar_wait(pv_done, (kv_tiles - 1) & 1);
    tc_fence_after();
    const int q_idx = qt * ATT_BLOCK_Q + row;
    uint32_t r[32];
    tmem_ld32(tmem_o, r);
    tmem_ld_wait();
    if (q_idx < seq_len) {
      uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(row0 + q_idx) * C + head * ATT_D + half * 32);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(r[8 * i + t]) * inv;
        dst[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
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
