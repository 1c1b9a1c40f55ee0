// Fused multi-head attention (non-causal, head_dim 64) on tcgen05 for sm_100a — "ping-pong" kernel, round 2.
//
// Same contract as attention_fwd_kernel (attention_sm100.cuh): replaces the q@k^T / softmax / @v core of
// rein/models/backbones/dino_layers/attention.py:56-66 and rein/models/heads/Transformer.py:113-136.
//
// What changed against the round-1 kernel, and why (profiles/r1_attention_timeline.txt): that kernel ran two CTAs per
// SM with 64-key tiles; a softmax warp spent ~620 clk of fixed cost per 64-key tile (barrier wake-up, fences, row max,
// reference check, store + arrive) around ~700 clk of exponentials, the MUFU pipe (the real ceiling at head_dim 64:
// 1024 clk of ex2 per 128 x 128 scores against 512 clk of MMA) was busy ~half of the time.
// Here ONE persistent CTA per SM owns all 512 TMEM columns and works on units of 256 query rows:
//   * two softmax warpgroups A / B (128 rows each, one row per thread) share every K / V tile and run half a tile
//     period apart, so the fixed per-tile cost of one hides under the exponentials of the other;
//   * 128-key tiles: half as many barrier round trips per score, S = Q K^T as N = 128 MMAs (64.5 clk per k-step
//     against 2 x 48.5 for two N = 64 ones: the N = 64 form is bound by the shared-memory read of A);
//   * S, P and O all have their own TMEM columns (S_A S_B 2 x 128 fp32, P_A P_B 2 x 64 packed bf16, O_A O_B 2 x 64
//     fp32 = 512), so S_X(t+1) is issued the moment warpgroup X has copied S_X(t) to registers — not behind
//     P_X(t) -> PV_X(t) as the aliased layout forces — and is ready long before X comes back for it;
//   * the row sums are kept in registers (packed FADD2), the row max uses the 3-input FMNMX, the scale/subtract is a
//     packed FFMA2, P is rounded to nearest (cvt.rn.bf16x2), not truncated;
//   * the S MMAs and the PV MMAs are issued by two different warps (tcgen05.mma blocks its issuing thread while the
//     pipe is busy), each serving A then B in a fixed order.
// "Extra token" mode (ViT sequences = 1 cls token + n x 128 patch tokens): the tensor tiles cover the body tokens; the
// cls KEY is one dot product per query row on the CUDA cores (before the loop) and one AXPY in the epilogue; the cls
// QUERY row of each (sequence, head) is computed by warp 3 of the CTAs straight from global memory while the tensor
// pipeline runs (1/1025 of the work; no second launch).
//
// Warps (384 threads): warpgroups 0, 1 = softmax A, B (TMEM lane quadrant = warp % 4); warpgroup 2 = {TMA producer,
// S issuer (+ TMEM allocator), PV issuer, extra-query rows}, setmaxnreg-trimmed. The service warps have the HIGHEST warp
// ids on purpose: the scheduler of an SM sub-partition prefers the higher warp id among ready warps, and an issuer warp
// that is late by a few hundred clocks stalls a whole softmax warpgroup.
#pragma once
#include "attention_sm100.cuh"

namespace vfm {

constexpr int APP_TILE_Q = 128;                       // query rows per softmax warpgroup
constexpr int APP_UNIT_Q = 2 * APP_TILE_Q;            // query rows per unit
constexpr int APP_BLOCK_KV = 128;
constexpr int APP_THREADS = 384;
constexpr int APP_Q_STAGES = 2, APP_K_STAGES = 3, APP_V_STAGES = 3;
constexpr int APP_TILE_BYTES = 128 * ATT_D * 2;       // 16 KB: one Q tile, one K tile, one V tile
constexpr int APP_MAX_EXTRA_KEYS = 4224;              // score scratch of the extra-query warp (floats)
constexpr int APP_BAR_BYTES = 512;
constexpr int APP_SMEM_BYTES = 1024 /* alignment slack */ + (2 * APP_Q_STAGES + APP_K_STAGES + APP_V_STAGES + 2 /* extra-query staging */) * APP_TILE_BYTES +
                               APP_BAR_BYTES + APP_MAX_EXTRA_KEYS * 4 + 64 * 4 /* extra query row */ + 2 * 2 * 2 * 128 * 4 /* row max / sum exchange */;
constexpr uint32_t APP_TMEM_COLS = 512;
constexpr uint32_t APP_COL_S = 0, APP_COL_P = 256, APP_COL_O = 384;   // + X * 128 / 64 / 64 for warpgroup X

#ifndef VFM_APP_POLY
#define VFM_APP_POLY 0      // n > 0: one exponential in n is evaluated on the FMA pipe (poly_exp2) instead of MUFU
#endif
#ifndef VFM_APP_ALUPACK
#define VFM_APP_ALUPACK 0   // 1: pack P with two integer adds + PRMT instead of F2FP (measured slower: 0.325 vs 0.301 ms)
#endif
#ifndef VFM_APP_HANDOFF
#define VFM_APP_HANDOFF 1   // the two softmax warpgroups take turns on the MUFU pipe (named-barrier token)
#endif

#ifdef VFM_APP_TRACE
// debug build: event timeline of CTA 0 — rows 0/1 softmax A/B (quad 0, lane 0), 2 S issuer, 3 PV issuer; [tile][event]
__device__ long long g_app_trace[4][64][8];
__device__ unsigned long long g_app_clk[4];   // clock64 / globaltimer at kernel start and end of CTA 0
#define APP_TRACE(who, tile, ev) do { if (blockIdx.x == 0 && lane == 0 && (tile) < 64) g_app_trace[who][tile][ev] = clock64(); } while (0)
#else
#define APP_TRACE(who, tile, ev)
#endif

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// v | (later & 0x80000000): v itself whenever `later` is non-negative, but data-dependent on `later` (one LOP3)
__device__ __forceinline__ uint32_t sign_gate(uint32_t v, uint32_t later) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, 0x80000000, 0xF8;" : "=r"(d) : "r"(v), "r"(later));
  return d;
}
// Two non-negative fp32 bit patterns -> packed bf16x2 (lo, hi), rounded half up, on the integer pipe: two adds and one
// PRMT. cvt.rn.bf16x2.f32 (F2FP) would be one instruction, but it executes on the same quarter-rate unit as MUFU.EX2:
// with it the 64 packs of a tile cost the MUFU pipe another 512 clk on top of the 1024 of the 128 exponentials
// (second GPU trace: 790 clk for the 64 FADD2 + 64 F2FP of a tile while the other warpgroup ran its exponentials).
// Half-up differs from round-to-nearest-even only on exact ties (probability 2^-16 per value).
__device__ __forceinline__ uint32_t pack_bf16x2_alu(uint32_t lo_bits, uint32_t hi_bits) {
  return __byte_perm(lo_bits + 0x8000u, hi_bits + 0x8000u, 0x7632u);
}
// (d0, d1) = (a0, a1) * b + c as one packed FFMA2
__device__ __forceinline__ void ffma2_bc(float& d0, float& d1, float a0, float a1, float b, float c) {
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b), "f"(c));
}
// (d0, d1) += (a0, a1) as one packed FADD2
__device__ __forceinline__ void fadd2_acc(float& d0, float& d1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rd, {%0, %1};\n\t"
      "add.rn.f32x2 rd, rd, ra;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1));
}

// The query row of the extra token (cls) of one (sequence, head) against all kv_len + 1 keys, by ONE warp on the CUDA
// cores (2 x 64 x (kv_len + 1) MACs: 1/1025 of the attention work of a ViT window) while the tensor pipeline of the CTA
// runs. The K rows, then the V rows of the pair stream through a private double buffer of 128-row TMA boxes (same tensor
// maps as the main pipeline, SWIZZLE_128B), so the warp never waits out a global-memory round trip per step: the first
// version read K / V with plain loads (8 x 16 B per lane in flight) and took ~137 k clk per pair — four pairs per CTA
// made it the longest chain of the kernel (0.349 ms per 36-window launch with it, ~0.25 ms worth of tile work).
// Pass 1: lane owns keys lane, lane + 32, lane + 64, lane + 96 of a chunk: scores (log2 units) to shared memory.
// Pass 2: lane owns output columns 2 lane, 2 lane + 1 and walks the 128 rows of a V chunk.
// ld / use: running counts of chunk loads issued / consumed by this warp (buffer = count & 1, parity = (count >> 1) & 1).
__device__ __forceinline__ void attention_extra_query_warp(const AttParams& p, const CUtensorMap* tmap_k, const CUtensorMap* tmap_v,
                                                           int seq, int head, float* sc, float* qs, uint8_t* stage,
                                                           uint64_t* full, uint32_t& ld, uint32_t& use) {
  const int lane = threadIdx.x & 31;
  const int kv_total = p.kv_len + 1;
  const int n_chunks = (kv_total + 127) >> 7;
  const int row0 = seq * p.kv_seq_rows;
  auto issue = [&](int c) {   // chunk c of the pair's stream: K chunks, then V chunks
    if (c < 2 * n_chunks) {
      const uint32_t b = ld & 1u;
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&full[b], APP_TILE_BYTES);
        if (c < n_chunks) tma_load_2d(stage + b * APP_TILE_BYTES, tmap_k, &full[b], p.k_col0 + head * ATT_D, row0 + c * 128);
        else tma_load_2d(stage + b * APP_TILE_BYTES, tmap_v, &full[b], p.v_col0 + head * ATT_D, row0 + (c - n_chunks) * 128);
      }
      __syncwarp();
      ++ld;
    }
  };
  issue(0);
  issue(1);
  {
    const uint32_t qv = __ldg(reinterpret_cast<const uint32_t*>(p.q_ptr + static_cast<size_t>(seq) * p.q_seq_rows * p.q_ld + p.q_col0 + head * ATT_D) + lane);
    qs[2 * lane] = bf16lo(qv);
    qs[2 * lane + 1] = bf16hi(qv);
  }
  __syncwarp();
  float m = -INFINITY;
  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t b = use & 1u;
    mbar_wait(&full[b], (use >> 1) & 1u);
    const uint8_t* tile = stage + b * APP_TILE_BYTES;
#pragma unroll
    for (int r4 = 0; r4 < 4; ++r4) {
      const int r = lane + 32 * r4;
      const uint8_t* krow = tile + r * 128;
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 kv = *reinterpret_cast<const uint4*>(krow + ((ch ^ (r & 7)) << 4));
        const float4 q0 = *reinterpret_cast<const float4*>(qs + 8 * ch);
        const float4 q1 = *reinterpret_cast<const float4*>(qs + 8 * ch + 4);
        a0 = fmaf(q0.x, bf16lo(kv.x), a0); a1 = fmaf(q0.y, bf16hi(kv.x), a1);
        a0 = fmaf(q0.z, bf16lo(kv.y), a0); a1 = fmaf(q0.w, bf16hi(kv.y), a1);
        a0 = fmaf(q1.x, bf16lo(kv.z), a0); a1 = fmaf(q1.y, bf16hi(kv.z), a1);
        a0 = fmaf(q1.z, bf16lo(kv.w), a0); a1 = fmaf(q1.w, bf16hi(kv.w), a1);
      }
      const int key = c * 128 + r;
      const float sv = (a0 + a1) * 1.4426950408889634f;
      if (key < kv_total) { sc[key] = sv; m = fmaxf(m, sv); }
    }
    ++use;
    __syncwarp();          // every lane has read the buffer: it may be refilled
    issue(c + 2);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float l = 0.f;
  for (int key = lane; key < kv_total; key += 32) {
    const float w = fast_exp2(sc[key] - m);
    sc[key] = w;
    l += w;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  for (int c = 0; c < n_chunks; ++c) {
    const uint32_t b = use & 1u;
    mbar_wait(&full[b], (use >> 1) & 1u);
    const uint8_t* tile = stage + b * APP_TILE_BYTES;
    const int n_valid = min(128, kv_total - c * 128);
    const float* w = sc + c * 128;
    // the lane's 4 bytes of row r: 16-byte chunk lane >> 2 (swizzled with r & 7), word lane & 3
    const int cw = lane >> 2, wo = (lane & 3) * 4;
    if (n_valid == 128) {
#pragma unroll 8
      for (int r = 0; r < 128; r += 2) {
        const uint32_t v0 = *reinterpret_cast<const uint32_t*>(tile + r * 128 + ((cw ^ (r & 7)) << 4) + wo);
        const uint32_t v1 = *reinterpret_cast<const uint32_t*>(tile + (r + 1) * 128 + ((cw ^ ((r + 1) & 7)) << 4) + wo);
        const float2 ww = *reinterpret_cast<const float2*>(w + r);
        o0 = fmaf(ww.x, bf16lo(v0), o0); o1 = fmaf(ww.x, bf16hi(v0), o1);
        o2 = fmaf(ww.y, bf16lo(v1), o2); o3 = fmaf(ww.y, bf16hi(v1), o3);
      }
    } else {
      for (int r = 0; r < n_valid; ++r) {
        const uint32_t v0 = *reinterpret_cast<const uint32_t*>(tile + r * 128 + ((cw ^ (r & 7)) << 4) + wo);
        o0 = fmaf(w[r], bf16lo(v0), o0); o1 = fmaf(w[r], bf16hi(v0), o1);
      }
    }
    ++use;
    __syncwarp();
    issue(n_chunks + c + 2);
  }
  const float inv = 1.f / l;
  reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(seq) * p.q_seq_rows * p.out_ld + head * ATT_D)[lane] =
      pack_bf16x2((o0 + o2) * inv, (o1 + o3) * inv);
  __syncwarp();
}

__device__ __forceinline__ bool named_bar_red_or(int id, int count, bool pred) {
  uint32_t out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.u32 q, %1, 0;\n\t"
      "barrier.cta.red.or.pred p, %2, %3, q;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"(static_cast<uint32_t>(pred)), "r"(id), "r"(count)
      : "memory");
  return out != 0;
}

// Softmax + output role of attention_pp_kernel<2>: 16 warps. Warp w: TMEM lane quadrant quad = w & 3, query tile X =
// (w >> 2) & 1 (A or B), column half h = w >> 3. A thread owns one query row of tile X and the key columns [64 h, 64 h + 64)
// of every key tile; its partner (same lane, warp w ^ 8) owns the other half of the same row. Both keep the same
// reference max m_ref: a tile's two half maxima are only exchanged (shared memory + a 64-thread named barrier) when one
// of the 64 rows x 2 halves of the warp pair needs the reference moved, which one barrier.red.or per tile finds out.
// Why: with one row per thread (kSplit = 1) two softmax warps share an SM sub-partition and each warp's per-tile chain —
// barrier polls, TMEM load, row max, 64 packs, TMEM store, arrives: ~1700 clk around a 1050 clk MUFU pass — is mostly
// latency that only the ONE other warp can cover (ncu: issue slots 38 % busy, MUFU 50 %, nothing saturated). Four
// warps per sub-partition cover it the way a GPU is meant to.
__device__ __forceinline__ void attention_pp_softmax_split(const AttParams& p, uint32_t tmem_base, const uint8_t* smem_q, uint64_t* q_full,
                                                           uint64_t* q_empty, uint64_t* s_full, uint64_t* s_free, uint64_t* p_full,
                                                           uint64_t* p_free, uint64_t* o_free, float* xch, int first_unit, int unit_step,
                                                           int n_my, int kv_tiles, int q_pairs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int quad = warp & 3, x = (warp >> 2) & 1, h = warp >> 3;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
  const uint32_t tmem_s = tmem_base + lane_base + APP_COL_S + x * APP_BLOCK_KV + h * 64;
  const uint32_t tmem_p = tmem_base + lane_base + APP_COL_P + x * (APP_BLOCK_KV / 2) + h * 32;
  const uint32_t tmem_o = tmem_base + lane_base + APP_COL_O + x * ATT_D + h * 32;   // this half's 32 output columns
  const int pair_bar = 1 + x * 4 + quad;                                           // named barrier of the warp pair (64 threads)
  float* xm_mine = xch + ((0 * 2 + x) * 2 + h) * 128 + row;
  float* xm_other = xch + ((0 * 2 + x) * 2 + (h ^ 1)) * 128 + row;
  float* xl_mine = xch + ((1 * 2 + x) * 2 + h) * 128 + row;
  float* xl_other = xch + ((1 * 2 + x) * 2 + (h ^ 1)) * 128 + row;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr float kRescaleThreshold = 24.0f;   // log2 units (see attention_sm100.cuh)
  const int tail_valid = p.kv_len - (kv_tiles - 1) * APP_BLOCK_KV;   // keys in the last tile of a sequence (1..128)
  int t = 0;
  for (int k = 0; k < n_my; ++k) {
    const int u = first_unit + k * unit_step;
    const int qp = u % q_pairs, head = (u / q_pairs) % p.heads, seq = u / (q_pairs * p.heads);
    const int q_idx = qp * APP_UNIT_Q + x * APP_TILE_Q + row;   // body index of this thread's query row
    float m_ref = -INFINITY, w_extra = 0.f, l0 = 0.f, l1 = 0.f;
    if (p.extra) {
      // the extra key: s = q_row . k_extra on the CUDA cores (both halves compute it; half 0 carries its weight 1)
      const int qs = k % APP_Q_STAGES;
      mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
      const uint8_t* qrow = smem_q + (2 * qs + x) * APP_TILE_BYTES + row * 128;
      const uint4* kx = reinterpret_cast<const uint4*>(p.k_ptr + static_cast<size_t>(seq) * p.kv_seq_rows * p.k_ld + p.k_col0 + head * ATT_D);
      float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
        const uint4 kv = __ldg(kx + c);
        acc0 = fmaf(bf16lo(qv.x), bf16lo(kv.x), acc0); acc1 = fmaf(bf16hi(qv.x), bf16hi(kv.x), acc1);
        acc0 = fmaf(bf16lo(qv.y), bf16lo(kv.y), acc0); acc1 = fmaf(bf16hi(qv.y), bf16hi(kv.y), acc1);
        acc0 = fmaf(bf16lo(qv.z), bf16lo(kv.z), acc0); acc1 = fmaf(bf16hi(qv.z), bf16hi(kv.z), acc1);
        acc0 = fmaf(bf16lo(qv.w), bf16lo(kv.w), acc0); acc1 = fmaf(bf16hi(qv.w), bf16hi(kv.w), acc1);
      }
      m_ref = (acc0 + acc1) * kLog2e;
      w_extra = 1.f;
      if (h == 0) l0 = 1.f;
      __syncwarp();
      if (lane == 0) mbar_arrive(&q_empty[qs]);   // this warp no longer reads the Q tile from shared memory
    }

    for (int j = 0; j < kv_tiles; ++j, ++t) {
      mbar_wait(&s_full[x], t & 1);
      tc_fence_after();
      uint32_t s[64];
      tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[x]);   // S_X(t+1) may overwrite the score columns now
      const int valid = (j == kv_tiles - 1 ? tail_valid : APP_BLOCK_KV) - 64 * h;   // live keys in this half (<= 0: none)
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = 0xff800000u;   // -inf
      }
      float m4[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) m4[c] = fmaxf(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1]));
#pragma unroll
      for (int i = 8; i < 64; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) m4[c] = fmax3(m4[c], __uint_as_float(s[i + 2 * c]), __uint_as_float(s[i + 2 * c + 1]));
      }
      const float m_half = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kLog2e;
      // does any row of the warp pair need its reference moved? (always on a unit's first tile)
      const bool want = m_half > m_ref + (j == 0 ? 0.f : kRescaleThreshold);
      if (named_bar_red_or(pair_bar, 64, want)) {
        *xm_mine = m_half;
        named_bar_sync(pair_bar, 64);
        const float m_tile = fmaxf(m_half, *xm_other);
        const bool jump = m_tile > m_ref + (j == 0 ? 0.f : kRescaleThreshold);
        const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;   // exp2(-inf) = 0 on the very first tile
        if (jump) { m_ref = m_tile; w_extra *= alpha; l0 *= alpha; l1 *= alpha; }
        if (j > 0) {   // rescale this half's 32 columns of O_X (every PV_X up to tile t-1 has executed)
          mbar_wait(&p_free[x], (t - 1) & 1);
          tc_fence_after();
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
      }
      if (t > 0) {   // PV_X(t-1) has read P_X: the buffer may be rewritten
        mbar_wait(&p_free[x], (t - 1) & 1);
        tc_fence_after();
      }
      const float neg_m = -m_ref;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0, x1;
          ffma2_bc(x0, x1, __uint_as_float(s[32 * c + 2 * i]), __uint_as_float(s[32 * c + 2 * i + 1]), kLog2e, neg_m);
          const float e0 = fast_exp2(x0), e1 = fast_exp2(x1);
          fadd2_acc(l0, l1, e0, e1);
          pk[i] = pack_bf16x2(e0, e1);
        }
        tmem_st16(tmem_p + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[x]);
    }

    // ---- unit epilogue: row sum = both halves' partial sums; this half normalises and stores 32 of the 64 columns
    *xl_mine = l0 + l1;
    named_bar_sync(pair_bar, 64);
    const float inv = 1.f / ((l0 + l1) + *xl_other);
    mbar_wait(&p_free[x], (t - 1) & 1);   // the last PV_X of the unit (commits are ordered: all earlier ones too)
    tc_fence_after();
    uint32_t o[32];
    tmem_ld32(tmem_o, o);
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&o_free[x]);   // the next unit's first PV_X may overwrite O_X now
    if (q_idx < p.q_len) {
      const uint4* vx = reinterpret_cast<const uint4*>(p.v_ptr + static_cast<size_t>(seq) * p.kv_seq_rows * p.v_ld + p.v_col0 + head * ATT_D) + 4 * h;
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(seq * p.q_seq_rows + p.q_row_off + q_idx) * p.out_ld + head * ATT_D) + 4 * h;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * i + e]);
        if (p.extra) {
          const uint4 xv = __ldg(vx + i);
          v[0] = fmaf(w_extra, bf16lo(xv.x), v[0]); v[1] = fmaf(w_extra, bf16hi(xv.x), v[1]);
          v[2] = fmaf(w_extra, bf16lo(xv.y), v[2]); v[3] = fmaf(w_extra, bf16hi(xv.y), v[3]);
          v[4] = fmaf(w_extra, bf16lo(xv.z), v[4]); v[5] = fmaf(w_extra, bf16hi(xv.z), v[5]);
          v[6] = fmaf(w_extra, bf16lo(xv.w), v[6]); v[7] = fmaf(w_extra, bf16hi(xv.w), v[7]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= inv;
        dst[i] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
      }
    }
  }
}


// kSplit = 1: 8 softmax warps, one query row (128 scores per key tile) per thread.
// kSplit = 2: 16 softmax warps, two threads per query row — warps w and w + 8 own the same 32 rows, key columns
//             [0, 64) and [64, 128) of every tile — so four softmax warps share each SM sub-partition instead of two.
template <int kSplit>
__global__ void __launch_bounds__(32 * (8 * kSplit + 4), 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                               // [stage][tile A | tile B]
  uint8_t* smem_k = smem_q + 2 * APP_Q_STAGES * APP_TILE_BYTES;
  uint8_t* smem_v = smem_k + APP_K_STAGES * APP_TILE_BYTES;
  uint8_t* smem_x = smem_v + APP_V_STAGES * APP_TILE_BYTES;             // [2] K / V chunks of the extra-query warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_x + 2 * APP_TILE_BYTES);
  uint64_t* q_full = bars;                          // [2] TMA -> S issuer, softmax (extra key)
  uint64_t* q_empty = q_full + APP_Q_STAGES;        // [2] S issuer (last S of the unit executed) [+ softmax warps] -> TMA
  uint64_t* k_full = q_empty + APP_Q_STAGES;        // [3] TMA -> S issuer
  uint64_t* k_empty = k_full + APP_K_STAGES;        // [3] S issuer (S_A, S_B of the tile executed) -> TMA
  uint64_t* v_full = k_empty + APP_K_STAGES;        // [3] TMA -> PV issuer
  uint64_t* v_empty = v_full + APP_V_STAGES;        // [3] PV issuer (PV_A, PV_B of the tile executed) -> TMA
  uint64_t* s_full = v_empty + APP_V_STAGES;        // [2] S issuer -> softmax X   (S_X(t) in TMEM)
  uint64_t* s_free = s_full + 2;                    // [2] softmax X -> S issuer   (S_X(t) copied to registers)
  uint64_t* p_full = s_free + 2;                    // [2] softmax X -> PV issuer  (P_X(t) in TMEM, O_X rescaled if needed)
  uint64_t* p_free = p_full + 2;                    // [2] PV issuer -> softmax X  (O_X += P_X(t) V(t) executed)
  uint64_t* o_free = p_free + 2;                    // [2] softmax X -> PV issuer  (O_X of the finished unit copied out)
  uint64_t* x_full = o_free + 2;                   // [2] TMA -> extra-query warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_full + 2);
  float* extra_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + APP_BAR_BYTES);   // [APP_MAX_EXTRA_KEYS] (+ 64 for q)

  float* xch = extra_sc + APP_MAX_EXTRA_KEYS + 64;   // [2 kinds][2 X][2 halves][128 rows]: row max / row sum exchange (kSplit = 2)

  constexpr int kSW = 8 * kSplit;                   // softmax warps; the four service warps follow
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int kv_tiles = (p.kv_len + APP_BLOCK_KV - 1) / APP_BLOCK_KV;
  const int q_pairs = p.q_tiles;   // units per (sequence, head)
  const int first_unit = blockIdx.x, unit_step = gridDim.x;
  const int n_my = (p.n_units - first_unit + unit_step - 1) / unit_step;   // >= 1 (grid <= units)

  if (warp == kSW && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < APP_Q_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], p.extra ? 1 + kSW : 1); }
    for (int s = 0; s < APP_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < APP_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1); mbar_init(&s_free[x], 4 * kSplit); mbar_init(&p_full[x], 4 * kSplit); mbar_init(&p_free[x], 1);
      mbar_init(&o_free[x], 4 * kSplit);
    }
    mbar_init(&x_full[0], 1); mbar_init(&x_full[1], 1);
    fence_barrier_init();
  }
  if (warp == kSW + 1) tmem_alloc<APP_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
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

  if (warp >= kSW) {
    if constexpr (kSplit == 1) asm volatile("setmaxnreg.dec.sync.aligned.u32 96;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == kSW) {
      // ===================== TMA producer =====================
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        const Unit un = unit_of(k);
        const int qs = k % APP_Q_STAGES;
        mbar_wait(&q_empty[qs], ((k / APP_Q_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&q_full[qs], 2 * APP_TILE_BYTES);
          tma_load_2d(smem_q + (2 * qs) * APP_TILE_BYTES, &tmap_q, &q_full[qs], p.q_col0 + un.head * ATT_D, un.q_row0);
          tma_load_2d(smem_q + (2 * qs + 1) * APP_TILE_BYTES, &tmap_q, &q_full[qs], p.q_col0 + un.head * ATT_D, un.q_row0 + APP_TILE_Q);
        }
        __syncwarp();
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int ks = t % APP_K_STAGES, vs = t % APP_V_STAGES;
          mbar_wait(&k_empty[ks], ((t / APP_K_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&k_full[ks], APP_TILE_BYTES);
            tma_load_2d(smem_k + ks * APP_TILE_BYTES, &tmap_k, &k_full[ks], p.k_col0 + un.head * ATT_D, un.kv_row0 + j * APP_BLOCK_KV);
          }
          __syncwarp();
          mbar_wait(&v_empty[vs], ((t / APP_V_STAGES) & 1) ^ 1);
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(&v_full[vs], APP_TILE_BYTES);
            tma_load_2d(smem_v + vs * APP_TILE_BYTES, &tmap_v, &v_full[vs], p.v_col0 + un.head * ATT_D, un.kv_row0 + j * APP_BLOCK_KV);
          }
          __syncwarp();
        }
      }
    } else if (warp == kSW + 1) {
      // ===================== S issuer: S_X(t) = Q_X K(t)^T, X = A then B =====================
      constexpr uint32_t idesc_s = make_idesc_bf16(APP_TILE_Q, APP_BLOCK_KV, 0, 0);
      const uint64_t dq0 = make_sw128_desc(smem_u32(smem_q));
      const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        const int qs = k % APP_Q_STAGES;
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int ks = t % APP_K_STAGES;
          const bool last = j == kv_tiles - 1;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (x == 0) {
              if (j == 0) mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
              mbar_wait(&k_full[ks], (t / APP_K_STAGES) & 1);
            }
            APP_TRACE(2, t, 3 * x);
            if (t > 0) mbar_wait(&s_free[x], (t - 1) & 1);   // warpgroup X has S_X(t-1) in registers
            tc_fence_after();
            APP_TRACE(2, t, 3 * x + 1);
            if (elect_one_sync()) {
              const uint64_t dq = dq0 + static_cast<uint64_t>((2 * qs + x) * (APP_TILE_BYTES >> 4));
              const uint64_t dk = dk0 + static_cast<uint64_t>(ks * (APP_TILE_BYTES >> 4));
              const uint32_t tmem_s = tmem_base + APP_COL_S + x * APP_BLOCK_KV;
#pragma unroll
              for (int kk = 0; kk < ATT_D / 16; ++kk) umma_ss(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
              tc_commit(&s_full[x]);
              if (x == 1) {
                tc_commit(&k_empty[ks]);
                if (last) tc_commit(&q_empty[qs]);
              }
            }
            __syncwarp();
            APP_TRACE(2, t, 3 * x + 2);
          }
        }
      }
    } else if (warp == kSW + 2) {
      // ===================== PV issuer: O_X += P_X(t) V(t), X = A then B =====================
      constexpr uint32_t idesc_pv = make_idesc_bf16(APP_TILE_Q, ATT_D, 0, 1);   // B = V is MN-major
      const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
      const int tail_ksteps = (p.kv_len - (kv_tiles - 1) * APP_BLOCK_KV + 15) >> 4;
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int vs = t % APP_V_STAGES;
          const int ksteps = j == kv_tiles - 1 ? tail_ksteps : APP_BLOCK_KV / 16;
          mbar_wait(&v_full[vs], (t / APP_V_STAGES) & 1);
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            APP_TRACE(3, t, 3 * x);
            if (j == 0 && k > 0) mbar_wait(&o_free[x], (k - 1) & 1);   // the previous unit's O_X has been copied out
            mbar_wait(&p_full[x], t & 1);
            tc_fence_after();
            APP_TRACE(3, t, 3 * x + 1);
            if (elect_one_sync()) {
              const uint64_t dv = dv0 + static_cast<uint64_t>(vs * (APP_TILE_BYTES >> 4));
              const uint32_t tmem_o = tmem_base + APP_COL_O + x * ATT_D;
              const uint32_t tmem_p = tmem_base + APP_COL_P + x * (APP_BLOCK_KV / 2);
              // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
              for (int kk = 0; kk < ksteps; ++kk) umma_ts(tmem_o, tmem_p + 8 * kk, dv + 128 * kk, idesc_pv, (j | kk) != 0);
              tc_commit(&p_free[x]);
              if (x == 1) tc_commit(&v_empty[vs]);
            }
            __syncwarp();
            APP_TRACE(3, t, 3 * x + 2);
          }
        }
      }
    } else if (p.extra) {
      // ===================== extra-token query rows (CUDA cores, background) =====================
      const int pairs = p.n_units / q_pairs;   // (sequence, head) pairs
      uint32_t ld = 0, use = 0;
      for (int i = blockIdx.x; i < pairs; i += gridDim.x)
        attention_extra_query_warp(p, &tmap_k, &tmap_v, i / p.heads, i % p.heads, extra_sc, extra_sc + APP_MAX_EXTRA_KEYS, smem_x,
                                   x_full, ld, use);
    }
  } else if constexpr (kSplit == 2) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    attention_pp_softmax_split(p, tmem_base, smem_q, q_full, q_empty, s_full, s_free, p_full, p_free, o_free, xch, first_unit, unit_step,
                               n_my, kv_tiles, q_pairs);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ===================== softmax + output: warpgroup X = A (warps 0..3) or B (warps 4..7) =====================
    // Per key tile t (one query row per thread, its 128 scores in registers s[]):
    //   reference check   m_ref moves (and O_X is rescaled) only when the tile max exceeds it by > 2^24
    //   token             named barrier: the MUFU pipe is ours
    //   MUFU pass         64 FFMA2 + 128 MUFU.EX2, in place in s[]; nothing else, so the pipe runs at 8 clk / instruction
    //   token release
    //   consumer pass     FADD2 row sums + F2FP packs + 4 x tcgen05.st of P(t) — and, chunk by chunk into the registers
    //                     this frees, the tcgen05.ld of S(t+1), which the S issuer finished long ago; then its row max.
    // so the only work of a tile that is not hidden behind the OTHER warpgroup's MUFU pass is what sits between the
    // token release and the next token request. (First version of this loop: barrier wake-up, TMEM load, row max and
    // P-buffer wait were all in front of the MUFU pass, ~1700 clk per tile against a 1240 clk MUFU pass of the other
    // warpgroup: ncu source view of that build, profiles/r2_attn_pp_notes.txt.)
    // Warps whose 32 rows lie past the end of the sequence are not special-cased: their rows are computed and never
    // stored (MMA rows are independent); ViT shapes have none.
    const int x = warp >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base + APP_COL_S + x * APP_BLOCK_KV;
    const uint32_t tmem_p = tmem_base + lane_base + APP_COL_P + x * (APP_BLOCK_KV / 2);
    const uint32_t tmem_o = tmem_base + lane_base + APP_COL_O + x * ATT_D;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 24.0f;   // log2 units (see attention_sm100.cuh)
    constexpr int kPoly = VFM_APP_POLY;
    const uint32_t sched_mask = static_cast<uint32_t>(p.sched_mask);
    const int total_tiles = n_my * kv_tiles;
    const int tail_valid = p.kv_len - (kv_tiles - 1) * APP_BLOCK_KV;   // keys in the last tile of a sequence (1..128)

    uint32_t s[128];
    // row max (times log2e) of the tile in s[]; keys past the end of the sequence (tail tile) are masked to -inf first
    auto row_max = [&](int valid) {
      if (valid < APP_BLOCK_KV) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = 0xff800000u;
      }
      float m8[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) m8[c] = fmaxf(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1]));
#pragma unroll
      for (int i = 16; i < 128; i += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) m8[c] = fmax3(m8[c], __uint_as_float(s[i + 2 * c]), __uint_as_float(s[i + 2 * c + 1]));
      }
      return fmaxf(fmax3(m8[0], m8[1], m8[2]), fmaxf(fmax3(m8[3], m8[4], m8[5]), fmaxf(m8[6], m8[7]))) * kLog2e;
    };

    // ---- prologue: the first tile of this CTA's stream
    mbar_wait(&s_full[x], 0);
    tc_fence_after();
    tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
    tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
    tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
    tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_free[x]);
    float m_next = row_max(kv_tiles == 1 ? tail_valid : APP_BLOCK_KV);
#if VFM_APP_HANDOFF
    // MUFU passes alternate A(t), B(t), A(t+1), ...: warpgroup X enters its pass through named barrier 1 + X (256
    // threads: its own 128 bar.sync + the other warpgroup's 128 bar.arrive at the end of ITS pass). B opens the first
    // A pass here. Without the token the two warpgroups drift into phase: both in their exponentials (sharing the
    // pipe), then both in their bookkeeping (pipe idle) — 3300 clk per 256 x 128 scores in the first traces.
    if (x == 1) named_bar_arrive(1, 256);
#endif
    int t = 0;
    for (int k = 0; k < n_my; ++k) {
      const Unit un = unit_of(k);
      const int q_tile0 = un.qp * APP_UNIT_Q + x * APP_TILE_Q;   // body index of this warpgroup's first query row
      float m_ref = -INFINITY, w_extra = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;   // row sum = l0 + l1 + l2 + l3
      if (p.extra) {
        // the extra key: s = q_row . k_extra on the CUDA cores; it starts the running softmax with weight 1
        const int qs = k % APP_Q_STAGES;
        mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
        const uint8_t* qrow = smem_q + (2 * qs + x) * APP_TILE_BYTES + row * 128;
        const uint4* kx = reinterpret_cast<const uint4*>(p.k_ptr + static_cast<size_t>(un.seq) * p.kv_seq_rows * p.k_ld + p.k_col0 + un.head * ATT_D);
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c ^ (row & 7)) << 4));
          const uint4 kv = __ldg(kx + c);
          acc0 = fmaf(bf16lo(qv.x), bf16lo(kv.x), acc0); acc1 = fmaf(bf16hi(qv.x), bf16hi(kv.x), acc1);
          acc0 = fmaf(bf16lo(qv.y), bf16lo(kv.y), acc0); acc1 = fmaf(bf16hi(qv.y), bf16hi(kv.y), acc1);
          acc0 = fmaf(bf16lo(qv.z), bf16lo(kv.z), acc0); acc1 = fmaf(bf16hi(qv.z), bf16hi(kv.z), acc1);
          acc0 = fmaf(bf16lo(qv.w), bf16lo(kv.w), acc0); acc1 = fmaf(bf16hi(qv.w), bf16hi(kv.w), acc1);
        }
        m_ref = (acc0 + acc1) * kLog2e;
        w_extra = 1.f;
        l0 = 1.f;
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qs]);   // this warp no longer reads the Q tile from shared memory
      }

      for (int j = 0; j < kv_tiles; ++j, ++t) {
        if (quad == 0) APP_TRACE(x, t, 0);
        const float m_tile = m_next;
        {
          const bool jump = m_tile > m_ref + (j == 0 ? 0.f : kRescaleThreshold);
          if (__any_sync(0xffffffffu, jump)) {   // rare after the first tile: rescale O in TMEM
            const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;   // exp2(-inf) = 0 on the very first tile
            if (jump) { m_ref = m_tile; w_extra *= alpha; l0 *= alpha; l1 *= alpha; l2 *= alpha; l3 *= alpha; }
            if (j > 0) {
              mbar_wait(&p_free[x], (t - 1) & 1);   // every PV_X up to tile t-1 has executed
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
        if (quad == 0) APP_TRACE(x, t, 3);
#if VFM_APP_HANDOFF
        named_bar_sync(1 + x, 256);         // the other warpgroup has issued the last exponential of its pass
#endif
        if (quad == 0) APP_TRACE(x, t, 4);
        const float neg_m = -m_ref;
        // The two passes sit under branches on bits of a kernel parameter (AttParams::sched_mask, all ones at run time).
        // Left alone, ptxas schedules every consumer of an exponential directly behind it (the FADD2 chain looks
        // critical to it) and the in-order warp sits out the MUFU latency once per pair: measured 28 clk per pair
        // instead of the 16 the pipe needs. A fake data dependence on a later exponential (round 1's trick) only
        // moves the wait, and warp-level barriers are scheduled across; real control flow is what ptxas respects.
        const bool more = t + 1 < total_tiles;
        if (sched_mask & 1u) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            float x0, x1;
            ffma2_bc(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), kLog2e, neg_m);
            s[2 * i] = __float_as_uint(fast_exp2(x0));
            s[2 * i + 1] = __float_as_uint((kPoly != 0 && (i % (kPoly ? kPoly : 1)) == 0) ? poly_exp2(x1) : fast_exp2(x1));
          }
        }
        if (sched_mask & 2u) {
#if VFM_APP_HANDOFF
          named_bar_arrive(2 - x, 256);     // the pipe goes to the other warpgroup
#endif
          if (quad == 0) APP_TRACE(x, t, 7);
          if (more) mbar_wait(&s_full[x], (t + 1) & 1);   // S_X(t+1): issued as soon as S_X(t) had been copied out, a tile ago
          if (t > 0) mbar_wait(&p_free[x], (t - 1) & 1);   // PV_X(t-1) has read P_X: the buffer may be rewritten
          tc_fence_after();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int kk = 16 * c + i;
              if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
              else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#if VFM_APP_ALUPACK
              pk[i] = pack_bf16x2_alu(s[2 * kk], s[2 * kk + 1]);
#else
              pk[i] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#endif
            }
            tmem_st16(tmem_p + c * 16, pk);
            if (more) tmem_ld32(tmem_s + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));   // into the registers just consumed
          }
          if (quad == 0) APP_TRACE(x, t, 5);
          if (more) {
            // the row max of tile t+1 is what stands between this warpgroup and its next token request: first
            tmem_ld_wait();
            const int jn = j + 1 == kv_tiles ? 0 : j + 1;
            m_next = row_max(jn == kv_tiles - 1 ? tail_valid : APP_BLOCK_KV);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&p_full[x]);
            if (more) mbar_arrive(&s_free[x]);   // S_X(t+2) may overwrite the score columns now
          }
          if (quad == 0) APP_TRACE(x, t, 6);
        }
      }

      // ---- unit epilogue: O_X out of TMEM in 16-column chunks (s[] already holds the next unit's first tile),
      // normalise, store; then hand the accumulator back
      mbar_wait(&p_free[x], (t - 1) & 1);   // the last PV_X of the unit (commits are ordered: all earlier ones too)
      tc_fence_after();
      const float inv = 1.f / ((l0 + l1) + (l2 + l3));
      const int q_idx = q_tile0 + row;
      const uint4* vx = reinterpret_cast<const uint4*>(p.v_ptr + static_cast<size_t>(un.seq) * p.kv_seq_rows * p.v_ld + p.v_col0 + un.head * ATT_D);
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(un.seq * p.q_seq_rows + p.q_row_off + q_idx) * p.out_ld + un.head * ATT_D);
#pragma unroll 1
      for (int c = 0; c < ATT_D / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(tmem_o + c * 16, o);
        tmem_ld_wait();
        if (q_idx < p.q_len) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * i + e]);
            if (p.extra) {
              const uint4 xv = __ldg(vx + 2 * c + i);
              v[0] = fmaf(w_extra, bf16lo(xv.x), v[0]); v[1] = fmaf(w_extra, bf16hi(xv.x), v[1]);
              v[2] = fmaf(w_extra, bf16lo(xv.y), v[2]); v[3] = fmaf(w_extra, bf16hi(xv.y), v[3]);
              v[4] = fmaf(w_extra, bf16lo(xv.z), v[4]); v[5] = fmaf(w_extra, bf16hi(xv.z), v[5]);
              v[6] = fmaf(w_extra, bf16lo(xv.w), v[6]); v[7] = fmaf(w_extra, bf16hi(xv.w), v[7]);
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
  if (warp == kSW + 1) {
    tc_fence_after();
    tmem_dealloc<APP_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
