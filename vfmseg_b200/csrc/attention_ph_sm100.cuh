// Fused multi-head attention (non-causal, head_dim 64) on tcgen05 for sm_100a — "per-half" pipeline, round 2.
//
// Same contract, units, shared-memory / TMEM layout and service warps as attention_pp_kernel (attention_pp_sm100.cuh; replaces
// the q@k^T / softmax / @v core of rein/models/backbones/dino_layers/attention.py:56-66). What differs is the softmax warp's
// schedule. There, a warp handles a 128-key tile in one piece: MUFU pass (~1300 clk) -> pack / tcgen05.st of P / tcgen05.ld of
// S(t+1) -> wait for both -> row max: ~1800 clk outside the pass, most of it TMEM and mbarrier round-trip latency, so the tile
// period of the two alternating warpgroups is pass + rest ~ 3100 clk against the 2 x 1024 clk the MUFU pipe needs
// (profiles/r2_attn_pp_trace_*.txt). Here the tile is handled as two 64-key halves whose registers are refilled right after they
// are consumed, and NOTHING is waited for in the half-iteration that issued it:
//   half-iteration (t, h) of warpgroup X:  token -> MUFU pass on half h (32 FFMA2 + 64 MUFU.EX2) -> token to the partner ->
//     complete what the PREVIOUS half-iteration issued (tcgen05.wait::st / ::ld are free by now: P(prev) goes to the PV issuer,
//     the S columns of that half go back to the S issuer, its row maximum is taken) -> pack + tcgen05.st of P(t, h) ->
//     tcgen05.ld of S(t+1, h) into the registers just consumed.
// tcgen05.wait::ld waits for ALL outstanding loads, so the wait sits BEFORE the new loads are issued: every load has the partner's
// pass and this warp's next pass to complete (the P store is completed at once: see the note at its tcgen05.wait::st). For that the S and PV MMAs are issued per half as well (S: N = 64,
// PV: four k-steps), each with its own full / free barrier pair; the S issuer may write half h of S(t+1) as soon as half h of S(t)
// is in registers, half a tile before the other half.
#pragma once
#include <type_traits>

#include "attention_pp_sm100.cuh"

namespace vfm {

constexpr int APH_HALF = APP_BLOCK_KV / 2;   // keys per half-iteration

__global__ void __launch_bounds__(APP_THREADS, 1)
attention_ph_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_q = smem;                                               // [stage][tile A | tile B]
  uint8_t* smem_k = smem_q + 2 * APP_Q_STAGES * APP_TILE_BYTES;
  uint8_t* smem_v = smem_k + APP_K_STAGES * APP_TILE_BYTES;
  uint8_t* smem_x = smem_v + APP_V_STAGES * APP_TILE_BYTES;             // [2] K / V chunks of the extra-query warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_x + 2 * APP_TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* q_empty = q_full + APP_Q_STAGES;
  uint64_t* k_full = q_empty + APP_Q_STAGES;
  uint64_t* k_empty = k_full + APP_K_STAGES;
  uint64_t* v_full = k_empty + APP_K_STAGES;
  uint64_t* v_empty = v_full + APP_V_STAGES;
  uint64_t* s_full = v_empty + APP_V_STAGES;        // [X][half]   S issuer -> softmax: half h of S_X(t) is in TMEM
  uint64_t* s_free = s_full + 4;                    // [X][half]   softmax -> S issuer: half h of S_X(t) is in registers
  uint64_t* p_full = s_free + 4;                    // [X][half]   softmax -> PV issuer: half h of P_X(t) is stored
  uint64_t* p_free = p_full + 4;                    // [X][half]   PV issuer -> softmax: PV_X(t, h) executed
  uint64_t* o_ready = p_free + 4;                   // [2]
  uint64_t* o_free = o_ready + 2;                   // [2]
  uint64_t* x_full = o_free + 2;                    // [2] TMA -> extra-query warp
  uint64_t* l_full = x_full + 2;                    // [unit parity][2][4]
  uint64_t* e_full = l_full + 16;                   // (unused here; the shared output-warp code names it)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(l_full + 16);
  static_assert((2 * APP_Q_STAGES + 2 * APP_K_STAGES + 2 * APP_V_STAGES + 4 * 4 + 3 * 2 + 16) * 8 + 8 <= APP_BAR_BYTES, "barrier block too small");
  float* extra_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + APP_BAR_BYTES);   // [APP_MAX_EXTRA_KEYS] (+ 64 for q)
  float2* lw_smem = reinterpret_cast<float2*>(extra_sc + APP_MAX_EXTRA_KEYS + 64);   // [X][unit parity][row]: (row sum, extra key's weight)
  float* es_smem = reinterpret_cast<float*>(lw_smem + 2 * 2 * 128);
  uint8_t* smem_xr = reinterpret_cast<uint8_t*>(es_smem + 2 * 2 * 128);    // [APP_XR_STAGES][extra key row | extra value row]
  (void)e_full; (void)es_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int kv_tiles = (p.kv_len + APP_BLOCK_KV - 1) / APP_BLOCK_KV;
  const int q_pairs = p.q_tiles;   // units per (sequence, head)
  const int first_unit = blockIdx.x, unit_step = gridDim.x;
  const int n_my = (p.n_units - first_unit + unit_step - 1) / unit_step;   // >= 1 (grid <= units)
  const int total_tiles = n_my * kv_tiles;
  const int tail_valid = p.kv_len - (kv_tiles - 1) * APP_BLOCK_KV;         // keys in the last tile of a sequence (1..128)

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < APP_Q_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], p.extra ? 9 : 1); }
    for (int s = 0; s < APP_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < APP_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4); mbar_init(&p_full[i], 4); mbar_init(&p_free[i], 1); }
    for (int x = 0; x < 2; ++x) { mbar_init(&o_ready[x], 1); mbar_init(&o_free[x], 4); mbar_init(&x_full[x], 1); }
    for (int i = 0; i < 16; ++i) mbar_init(&l_full[i], 1);
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 9) tmem_alloc<APP_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
#ifdef VFM_APP_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_app_clk[0] = clock64(); g_app_clk[1] = ns;
  }
#endif

  struct Unit { int q_row0, kv_row0, head, seq, qp; };
  auto unit_of = [&](int k) {
    const int u = first_unit + k * unit_step;
    Unit r;
    r.qp = u % q_pairs;
    r.head = (u / q_pairs) % p.heads;
    r.seq = u / (q_pairs * p.heads);
    r.q_row0 = r.seq * p.q_seq_rows + p.q_row_off + r.qp * APP_UNIT_Q;
    r.kv_row0 = r.seq * p.kv_seq_rows + p.kv_row_off;
    return r;
  };

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(VFM_APP_SERVICE_REGS));
    if (warp == 8) {
      // ===================== TMA producer =====================
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        const Unit un = unit_of(k);
        const int qs = k % APP_Q_STAGES;
        APP_WAIT_BG(&q_empty[qs], ((k / APP_Q_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&q_full[qs], 2 * APP_TILE_BYTES + (p.extra ? 256 : 0));
          tma_load_2d(smem_q + (2 * qs) * APP_TILE_BYTES, &tmap_q, &q_full[qs], p.q_col0 + un.head * ATT_D, un.q_row0);
          tma_load_2d(smem_q + (2 * qs + 1) * APP_TILE_BYTES, &tmap_q, &q_full[qs], p.q_col0 + un.head * ATT_D, un.q_row0 + APP_TILE_Q);
          if (p.extra) {   // row 0 of the sequence's K and V (this head): read from shared memory by the softmax / helper threads
            uint8_t* xr = smem_xr + (k % APP_XR_STAGES) * 256;
            bulk_load_1d(xr, p.k_ptr + static_cast<size_t>(un.seq) * p.kv_seq_rows * p.k_ld + p.k_col0 + un.head * ATT_D, 128, &q_full[qs]);
            bulk_load_1d(xr + 128, p.v_ptr + static_cast<size_t>(un.seq) * p.kv_seq_rows * p.v_ld + p.v_col0 + un.head * ATT_D, 128, &q_full[qs]);
          }
        }
        __syncwarp();
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int ks = t % APP_K_STAGES, vs = t % APP_V_STAGES;
          APP_WAIT_BG(&k_empty[ks], ((t / APP_K_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&k_full[ks], APP_TILE_BYTES);
            tma_load_2d(smem_k + ks * APP_TILE_BYTES, &tmap_k, &k_full[ks], p.k_col0 + un.head * ATT_D, un.kv_row0 + j * APP_BLOCK_KV);
          }
          __syncwarp();
          APP_WAIT_BG(&v_empty[vs], ((t / APP_V_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&v_full[vs], APP_TILE_BYTES);
            tma_load_2d(smem_v + vs * APP_TILE_BYTES, &tmap_v, &v_full[vs], p.v_col0 + un.head * ATT_D, un.kv_row0 + j * APP_BLOCK_KV);
          }
          __syncwarp();
        }
      }
    } else if (warp == 9) {
      // ===================== S issuer: S_X(t, h) = Q_X K(t)[64 h .. 64 h + 63]^T, per tile (h0: A, B), (h1: A, B) =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(APP_TILE_Q, APH_HALF, 0, 0);
      const uint64_t dq0 = make_sw128_desc(smem_u32(smem_q));
      const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        const int qs = k % APP_Q_STAGES;
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int ks = t % APP_K_STAGES;
          const bool last = j == kv_tiles - 1;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (h == 0 && x == 0) {
                if (j == 0) mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
                mbar_wait(&k_full[ks], (t / APP_K_STAGES) & 1);
              }
              APP_TRACE(2, 2 * t + h, 2 * x);
              if (t > 0) mbar_wait(&s_free[2 * x + h], (t - 1) & 1);   // warpgroup X has half h of S_X(t-1) in registers
              tc_fence_after();
              APP_TRACE(2, 2 * t + h, 2 * x + 1);
              if (elect_one_sync()) {
                const uint64_t dq = dq0 + static_cast<uint64_t>((2 * qs + x) * (APP_TILE_BYTES >> 4));
                const uint64_t dk = dk0 + static_cast<uint64_t>(ks * (APP_TILE_BYTES >> 4) + h * (APH_HALF * 128 >> 4));   // key rows 64 h ..
                const uint32_t tmem_s = tmem_base + APP_COL_S + x * APP_BLOCK_KV + h * APH_HALF;
#pragma unroll
                for (int kk = 0; kk < ATT_D / 16; ++kk) umma_ss(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
                tc_commit(&s_full[2 * x + h]);
                if (h == 1 && x == 1) {
                  tc_commit(&k_empty[ks]);
                  if (last) tc_commit(&q_empty[qs]);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    } else if (warp == 10) {
      // ===================== PV issuer: O_X += P_X(t, h) V(t)[64 h ..], per tile (h0: A, B), (h1: A, B) =====================
      constexpr uint32_t idesc_pv = make_idesc_bf16(APP_TILE_Q, ATT_D, 0, 1);   // B = V is MN-major
      const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int vs = t % APP_V_STAGES;
          const bool last = j == kv_tiles - 1;
          mbar_wait(&v_full[vs], (t / APP_V_STAGES) & 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // k-steps of 16 keys that hold at least one key of the sequence (rows past its end may belong to the next one)
            const int valid = last ? min(max(tail_valid - h * APH_HALF, 0), APH_HALF) : APH_HALF;
            const int ksteps = (valid + 15) >> 4;
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (j == 0 && h == 0 && k > 0) mbar_wait(&o_free[x], (k - 1) & 1);   // the previous unit's O_X has been copied out
              APP_TRACE(3, 2 * t + h, 2 * x);
              mbar_wait(&p_full[2 * x + h], t & 1);
              tc_fence_after();
              APP_TRACE(3, 2 * t + h, 2 * x + 1);
              if (elect_one_sync()) {
                const uint64_t dv = dv0 + static_cast<uint64_t>(vs * (APP_TILE_BYTES >> 4) + h * (APH_HALF * 128 >> 4));
                const uint32_t tmem_o = tmem_base + APP_COL_O + x * ATT_D;
                const uint32_t tmem_p = tmem_base + APP_COL_P + x * (APP_BLOCK_KV / 2) + h * (APH_HALF / 2);
                // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
                for (int kk = 0; kk < ksteps; ++kk) umma_ts(tmem_o, tmem_p + 8 * kk, dv + 128 * kk, idesc_pv, (j | h | kk) != 0);
                tc_commit(&p_free[2 * x + h]);
                if (last && h == 1) tc_commit(&o_ready[x]);
                if (x == 1 && h == 1) tc_commit(&v_empty[vs]);
              }
              __syncwarp();
            }
          }
        }
      }
    } else if (warp == 11) {
      // ===================== extra-token query rows (CUDA cores, background) =====================
      if (p.extra) {
        const int pairs = p.n_units / q_pairs;   // (sequence, head) pairs
        uint32_t ld = 0, use = 0;
        for (int i = blockIdx.x; i < pairs; i += gridDim.x)
          attention_extra_query_warp(p, &tmap_k, &tmap_v, i / p.heads, i % p.heads, extra_sc, extra_sc + APP_MAX_EXTRA_KEYS, smem_x,
                                     x_full, ld, use);
      }
    } else {
      // ===================== output warps: write every finished unit out =====================
      const int quad = warp & 3;
      const int row = quad * 32 + lane;
      const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
      // Output of unit k: O_X out of TMEM in 16-column chunks, + the extra key's value row, normalised, stored; then the
      // accumulator goes back to the PV issuer. (A TMEM load round trip costs ~400 clk here while the MMA pipe streams its
      // accumulators through TMEM — ~1900 clk for the four dependent rounds of one tile, trace p5 — but two 32-column loads
      // instead of four of 16 measured slower, 0.275 against 0.263 ms: the 56-register budget of these warps spills.)
      auto unit_output = [&](int k) {
        const Unit un = unit_of(k);
        const uint4* vx = reinterpret_cast<const uint4*>(smem_xr + (k % APP_XR_STAGES) * 256 + 128);
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          if (quad == 0 && x == 0) APP_TRACE(3, k, 6);
          APP_WAIT_BG(&l_full[(k & 1) * 8 + x * 4 + quad], (k >> 1) & 1);
          const float2 lw = lw_smem[(x * 2 + (k & 1)) * 128 + row];
          if (quad == 0 && x == 0) APP_TRACE(3, k, 7);
          APP_WAIT_BG(&o_ready[x], k & 1);
          tc_fence_after();
          if (quad == 0 && x == 0) APP_TRACE(2, k, 6);
          const float inv = 1.f / lw.x;
          const int q_idx = un.qp * APP_UNIT_Q + x * APP_TILE_Q + row;   // body index of this thread's query row
          const bool live = q_idx < p.q_len;
          uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(un.seq * p.q_seq_rows + p.q_row_off + q_idx) * p.out_ld + un.head * ATT_D);
          const uint32_t tmem_o = tmem_base + lane_base + APP_COL_O + x * ATT_D;
#if VFM_APP_OUT32
          // two 32-column rounds instead of four of 16 (needs the 64+ register budget of VFM_APP_SERVICE_REGS); the accumulator
          // goes back to the PV issuer as soon as the second round is in registers, before its global stores
#pragma unroll 1
          for (int c = 0; c < ATT_D / 32; ++c) {
            uint32_t o[32];
            tmem_ld32(tmem_o + c * 32, o);
            tmem_ld_wait();
            if (c == ATT_D / 32 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&o_free[x]);
            }
            if (live) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * i + e]);
                if (p.extra) {
                  const uint4 xv = vx[4 * c + i];
                  v[0] = fmaf(lw.y, bf16lo(xv.x), v[0]); v[1] = fmaf(lw.y, bf16hi(xv.x), v[1]);
                  v[2] = fmaf(lw.y, bf16lo(xv.y), v[2]); v[3] = fmaf(lw.y, bf16hi(xv.y), v[3]);
                  v[4] = fmaf(lw.y, bf16lo(xv.z), v[4]); v[5] = fmaf(lw.y, bf16hi(xv.z), v[5]);
                  v[6] = fmaf(lw.y, bf16lo(xv.w), v[6]); v[7] = fmaf(lw.y, bf16hi(xv.w), v[7]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= inv;
                dst[4 * c + i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              }
            }
          }
          if (quad == 0 && x == 0) APP_TRACE(2, k, 7);
          continue;
#endif
#pragma unroll 1
          for (int c = 0; c < ATT_D / 16; ++c) {
            uint32_t o[16];
            tmem_ld16(tmem_o + c * 16, o);
            tmem_ld_wait();
            if (live) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * i + e]);
                if (p.extra) {
                  const uint4 xv = vx[2 * c + i];
                  v[0] = fmaf(lw.y, bf16lo(xv.x), v[0]); v[1] = fmaf(lw.y, bf16hi(xv.x), v[1]);
                  v[2] = fmaf(lw.y, bf16lo(xv.y), v[2]); v[3] = fmaf(lw.y, bf16hi(xv.y), v[3]);
                  v[4] = fmaf(lw.y, bf16lo(xv.z), v[4]); v[5] = fmaf(lw.y, bf16hi(xv.z), v[5]);
                  v[6] = fmaf(lw.y, bf16lo(xv.w), v[6]); v[7] = fmaf(lw.y, bf16hi(xv.w), v[7]);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] *= inv;
                dst[2 * c + i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[x]);   // the next unit's first PV_X may overwrite O_X now
          if (quad == 0 && x == 0) APP_TRACE(2, k, 7);
        }
      };
      // The extra key's score of every query row of unit k (q_row . k_extra on the CUDA cores, both rows from shared
      // memory), published to the softmax warp of the same quadrant a whole unit ahead: on the softmax warps it sat
      // between two units (~1000 clk per unit with the wait for the Q tile, trace p5).
      auto extra_scores = [&](int k) {
        const int qs = k % APP_Q_STAGES;
        APP_WAIT_BG(&q_full[qs], (k / APP_Q_STAGES) & 1);
        const uint4* kx = reinterpret_cast<const uint4*>(smem_xr + (k % APP_XR_STAGES) * 256);
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          const uint8_t* qrow = smem_q + (2 * qs + x) * APP_TILE_BYTES + row * 128;
          float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
            const uint4 kv = kx[c];
            acc0 = fmaf(bf16lo(qv.x), bf16lo(kv.x), acc0); acc1 = fmaf(bf16hi(qv.x), bf16hi(kv.x), acc1);
            acc0 = fmaf(bf16lo(qv.y), bf16lo(kv.y), acc0); acc1 = fmaf(bf16hi(qv.y), bf16hi(kv.y), acc1);
            acc0 = fmaf(bf16lo(qv.z), bf16lo(kv.z), acc0); acc1 = fmaf(bf16hi(qv.z), bf16hi(kv.z), acc1);
            acc0 = fmaf(bf16lo(qv.w), bf16lo(kv.w), acc0); acc1 = fmaf(bf16hi(qv.w), bf16hi(kv.w), acc1);
          }
          es_smem[(x * 2 + (k & 1)) * 128 + row] = (acc0 + acc1) * 1.4426950408889634f;
          __syncwarp();
          if (lane == 0) mbar_arrive(&e_full[(k & 1) * 8 + x * 4 + quad]);
        }
        if (lane == 0) mbar_arrive(&q_empty[qs]);   // this warp no longer reads the stage's Q tiles
      };
#if VFM_APP_SCORES_ON_OUTPUT
      if (p.extra) extra_scores(0);
      for (int k = 0; k < n_my; ++k) {
        if (p.extra && k + 1 < n_my) extra_scores(k + 1);
        unit_output(k);
      }
#else
      for (int k = 0; k < n_my; ++k) unit_output(k);
#endif
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(VFM_APP_SOFTMAX_REGS));
    // ===================== softmax: warpgroup X = A (warps 0..3) or B (warps 4..7); see the file header =====================
    const int x = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + APP_COL_S + x * APP_BLOCK_KV;
    const uint32_t tmem_p = tmem_base + lane_base + APP_COL_P + x * (APP_BLOCK_KV / 2);
    const uint32_t tmem_o = tmem_base + lane_base + APP_COL_O + x * ATT_D;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 24.0f;   // log2 units (see attention_sm100.cuh)
    const uint32_t sched_mask = static_cast<uint32_t>(p.sched_mask);
    const int bar_mine = 1 + x, bar_other = 2 - x;
    constexpr int kBarThreads = 256;

    uint32_t s[128];          // scores of the current tile, then its exponentials; half h = s[64 h .. 64 h + 63]
    float m_half[2];          // row maximum (x log2e) of the scores each half holds
    bool pend_st = false, pend_ld = false;   // the previous half-iteration's P store / S loads have not been completed yet

    // keys of half h of tile j (of kv_tiles) that belong to the sequence
    auto half_valid = [&](int j, int h) { return j == kv_tiles - 1 ? min(max(tail_valid - h * APH_HALF, 0), APH_HALF) : APH_HALF; };
    // row max (times log2e) of half h; keys past the end of the sequence are masked to -inf first
    auto half_max = [&](auto hc, int valid) {
      constexpr int o = decltype(hc)::value * APH_HALF;
      if (valid < APH_HALF) {
#pragma unroll
        for (int i = 0; i < APH_HALF; ++i)
          if (i >= valid) s[o + i] = 0xff800000u;
      }
      float m8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m8[c] = fmaxf(__uint_as_float(s[o + 2 * c]), __uint_as_float(s[o + 2 * c + 1]));
#pragma unroll
      for (int i = 16; i < APH_HALF; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) m8[c] = fmax3(m8[c], __uint_as_float(s[o + i + 2 * c]), __uint_as_float(s[o + i + 2 * c + 1]));
      }
      return fmaxf(fmax3(m8[0], m8[1], m8[2]), fmaxf(fmax3(m8[3], m8[4], m8[5]), fmaxf(m8[6], m8[7]))) * kLog2e;
    };
    // What half-iteration (., hp) issued and did not wait for: its P store goes to the PV issuer, and (when it also loaded the
    // next scores of its half) those columns go back to the S issuer and their row maximum is taken. valid = keys of that half.
    auto complete = [&](auto hpc, int valid) {
      constexpr int hp = decltype(hpc)::value;
      if (pend_st) tmem_st_wait();
      if (pend_ld) tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pend_st) mbar_arrive(&p_full[2 * x + hp]);
        if (pend_ld) mbar_arrive(&s_free[2 * x + hp]);
      }
      if (pend_ld) m_half[hp] = half_max(hpc, valid);
      pend_st = pend_ld = false;
    };

    // ---- prologue: the first tile of this CTA's stream
    mbar_wait(&s_full[2 * x], 0);
    mbar_wait(&s_full[2 * x + 1], 0);
    tc_fence_after();
    tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
    tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
    tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
    tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(&s_free[2 * x]); mbar_arrive(&s_free[2 * x + 1]); }
    m_half[0] = half_max(std::integral_constant<int, 0>{}, half_valid(0, 0));
    m_half[1] = half_max(std::integral_constant<int, 1>{}, half_valid(0, 1));
    // MUFU passes alternate A(t, 0), B(t, 0), A(t, 1), B(t, 1), A(t + 1, 0), ...: B opens the first A pass
    if (x == 1) named_bar_arrive(bar_other, kBarThreads);

    int t = 0;
    for (int k = 0; k < n_my; ++k) {
      float m_ref = -INFINITY, w_extra = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;   // row sum = l0 + l1 + l2 + l3
      if (p.extra) {
        // the extra key starts the running softmax with weight 1: q_row . k_extra on the CUDA cores, both rows from shared memory
        const int qs = k % APP_Q_STAGES;
        mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
        const uint8_t* qrow = smem_q + (2 * qs + x) * APP_TILE_BYTES + row * 128;
        const uint4* kx = reinterpret_cast<const uint4*>(smem_xr + (k % APP_XR_STAGES) * 256);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
          const uint4 kv = kx[c];
          acc0 = fmaf(bf16lo(qv.x), bf16lo(kv.x), acc0); acc1 = fmaf(bf16hi(qv.x), bf16hi(kv.x), acc1);
          acc0 = fmaf(bf16lo(qv.y), bf16lo(kv.y), acc0); acc1 = fmaf(bf16hi(qv.y), bf16hi(kv.y), acc1);
          acc0 = fmaf(bf16lo(qv.z), bf16lo(kv.z), acc0); acc1 = fmaf(bf16hi(qv.z), bf16hi(kv.z), acc1);
          acc0 = fmaf(bf16lo(qv.w), bf16lo(kv.w), acc0); acc1 = fmaf(bf16hi(qv.w), bf16hi(kv.w), acc1);
        }
        m_ref = (acc0 + acc1) * kLog2e;
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qs]);   // this warp no longer reads the stage's Q tile
        w_extra = 1.f;
        l0 = 1.f;
      }

      for (int j = 0; j < kv_tiles; ++j, ++t) {
        const bool more = t + 1 < total_tiles;
        const int jn = j + 1 == kv_tiles ? 0 : j + 1;   // position of tile t + 1 inside its sequence
        // One half-iteration. hc = this half; the previous half-iteration was (t, 0) for h = 1 and (t - 1, 1) for h = 0.
        auto half_iter = [&](auto hc) {
          constexpr int h = decltype(hc)::value;
          constexpr int o = h * APH_HALF;
          using prev_t = std::integral_constant<int, 1 - h>;
          const bool first = h == 0 && j == 0;                        // first half-iteration of the unit: O_X holds nothing yet
          const int prev_valid = h == 0 ? half_valid(j, 1) : half_valid(jn, 0);   // keys of the half the previous half-iteration loaded
          if (quad == 0) APP_TRACE(x, 2 * t + h, 0);
          // ---- reference check: m_ref moves (and O_X is rescaled) only when this half's maximum exceeds it by > 2^24
          {
            const float m_tile = m_half[h];
            const bool jump = m_tile > m_ref + (first ? 0.f : kRescaleThreshold);
            if (__any_sync(0xffffffffu, jump)) {
              const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;   // exp2(-inf) = 0 on the very first half
              if (jump) { m_ref = m_tile; w_extra *= alpha; l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha; }
              if (!first) {   // rare: every PV_X issued so far must have executed before O_X is rescaled in TMEM
                if (pend_st || pend_ld) complete(prev_t{}, prev_valid);
                mbar_wait(&p_free[2 * x + (1 - h)], (h == 0 ? t - 1 : t) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < ATT_D / 16; ++c) {
                  uint32_t r[16];
                  tmem_ld16(tmem_o + c * 16, r);
                  tmem_ld_wait();
#pragma unroll
                  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                  tmem_st16(tmem_o + c * 16, r);
                }
                tmem_st_wait();
              }
            }
          }
          named_bar_sync(bar_mine, kBarThreads);   // the partner has issued the last exponential of its pass
          if (quad == 0) APP_TRACE(x, 2 * t + h, 1);
          const float neg_m = -m_ref;
          // the two passes sit under branches on bits of a kernel parameter (all ones at run time) so that ptxas cannot
          // interleave them: the pass must be nothing but FFMA2 + MUFU, 8 clk apart (see attention_pp_sm100.cuh)
          if (sched_mask & 1u) {
#pragma unroll
            for (int i = 0; i < APH_HALF / 2; ++i) {
              float x0, x1;
              ffma2_bc(x0, x1, __uint_as_float(s[o + 2 * i]), __uint_as_float(s[o + 2 * i + 1]), kLog2e, neg_m);
              s[o + 2 * i] = __float_as_uint(fast_exp2(x0));
              s[o + 2 * i + 1] = __float_as_uint(fast_exp2(x1));
            }
          }
          if (sched_mask & 2u) {
            named_bar_arrive(bar_other, kBarThreads);   // the pipe goes to the partner
            if (quad == 0) APP_TRACE(x, 2 * t + h, 2);
            if (pend_st || pend_ld) complete(prev_t{}, prev_valid);
            if (quad == 0) APP_TRACE(x, 2 * t + h, 3);
            if (t > 0) mbar_wait(&p_free[2 * x + h], (t - 1) & 1);       // PV_X(t-1, h) has read this half of P_X
            if (more) mbar_wait(&s_full[2 * x + h], (t + 1) & 1);         // half h of S_X(t+1): issued half a tile ago
            tc_fence_after();
            if (quad == 0) APP_TRACE(x, 2 * t + h, 4);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t pk[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int kk = o / 2 + 16 * c + i;
                if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
                else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
                pk[i] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
              }
              tmem_st16(tmem_p + o / 2 + c * 16, pk);
              if (more) tmem_ld32(tmem_s + o + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[o + 32 * c]));   // into the registers just consumed
            }
            // The P store is completed here, not deferred like the loads: p_full needs the arrival of all four warps of the
            // warpgroup, and a warp in the (rare) rescale path above waits for the PV behind it BEFORE the token barrier its three
            // siblings must pass to reach their deferred arrival — a dead-lock (found by the peaked-score cases of the test matrix).
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[2 * x + h]);
            pend_st = false;
            pend_ld = more;
            if (quad == 0) APP_TRACE(x, 2 * t + h, 5);
          }
        };
        half_iter(std::integral_constant<int, 0>{});
        half_iter(std::integral_constant<int, 1>{});
      }

      // ---- end of the unit: its last P half goes to the PV issuer now (the output warps wait for that PV), the row sums to the
      // output warp of this quadrant; the loads of the next unit's first scores stay in flight
      if (pend_st) {
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[2 * x + 1]);
        pend_st = false;
      }
      const float l_total = (l0 + l1) + (l2 + l3);
      lw_smem[(x * 2 + (k & 1)) * 128 + row] = make_float2(l_total, w_extra);
      __syncwarp();
      if (lane == 0) mbar_arrive(&l_full[(k & 1) * 8 + x * 4 + quad]);
    }
  }

  tc_fence_before();
  __syncthreads();
#ifdef VFM_APP_TRACE
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_app_clk[2] = clock64(); g_app_clk[3] = ns;
  }
#endif
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<APP_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
