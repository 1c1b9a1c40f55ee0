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
//   * two softmax warpgroups A / B (128 rows each, one row per thread) share every K / V tile and take turns on the MUFU
//     pipe (a named-barrier token per SM sub-partition), so the bookkeeping of one hides under the exponentials of the
//     other;
//   * 128-key tiles: half as many barrier round trips per score, S = Q K^T as N = 128 MMAs (64.5 clk per k-step
//     against 2 x 48.5 for two N = 64 ones: the N = 64 form is bound by the shared-memory read of A);
//   * S, P and O all have their own TMEM columns (S_A S_B 2 x 128 fp32, P_A P_B 2 x 64 packed bf16, O_A O_B 2 x 64
//     fp32 = 512), so S_X(t+1) is issued the moment S_X(t) has been copied to registers — not behind
//     P_X(t) -> PV_X(t) as the aliased layout forces — and is ready long before X comes back for it;
//   * four OUTPUT warps (one per TMEM lane quadrant) write the finished units out (O from TMEM, + the extra key's value
//     row, x 1 / row sum, store) while the softmax warps are already in the next unit. In the first ping-pong version the
//     softmax warps did it themselves between two units: wait for the last PV, four dependent TMEM load -> global load ->
//     store rounds, then the next unit's key row from global memory — ~7900 clk per unit, 21 % of the kernel
//     (profiles/r2_attn_pp_trace_*). Tried and dropped: the same warps also computing the row maxima of every S tile
//     straight from TMEM for the softmax warps (profiles/r2_attn_pp_trace_helper_wg.txt): S is then read from TMEM twice,
//     the TMEM read port (64 B / clk per sub-partition, 256 clk per warp and tile) is shared with the softmax warps'
//     own loads and the maxima arrive late — 0.324 ms against 0.285;
//   * the row sums are kept in registers (packed FADD2), the scale/subtract is a packed FFMA2, P is rounded to
//     nearest (cvt.rn.bf16x2), not truncated;
//   * the S MMAs and the PV MMAs are issued by two different warps (tcgen05.mma blocks its issuing thread while the
//     pipe is busy), each serving A then B in a fixed order.
// "Extra token" mode (ViT sequences = 1 cls token + n x 128 patch tokens): the tensor tiles cover the body tokens; the
// cls KEY is one dot product per query row on the CUDA cores (key row staged in shared memory by the TMA warp) and one
// AXPY in the epilogue; the cls QUERY row of each (sequence, head) is computed by a service warp on the CUDA cores while
// the tensor pipeline runs (1/1025 of the work; no second launch).
//
// Warps (512 threads): 0-3 softmax A, 4-7 softmax B (TMEM lane quadrant = warp % 4); 8 TMA producer, 9 S issuer
// (+ TMEM allocator), 10 PV issuer, 11 extra-query rows; 12-15 output (quadrant = warp % 4). setmaxnreg: softmax 200,
// everything else 56 (256 x 200 + 256 x 56 = 65536 = the pool of a 512-thread CTA launched at 128 registers).
#pragma once
#include "attention_sm100.cuh"

namespace vfm {

constexpr int APP_TILE_Q = 128;                       // query rows per softmax warpgroup
constexpr int APP_UNIT_Q = 2 * APP_TILE_Q;            // query rows per unit
constexpr int APP_BLOCK_KV = 128;
constexpr int APP_THREADS = 512;
constexpr int APP_Q_STAGES = 2, APP_K_STAGES = 3, APP_V_STAGES = 3;
constexpr int APP_XR_STAGES = 8;                     // ring of (extra key row | extra value row) pairs, one per unit (see the helper's epilogue)
constexpr int APP_TILE_BYTES = 128 * ATT_D * 2;       // 16 KB: one Q tile, one K tile, one V tile
constexpr int APP_MAX_EXTRA_KEYS = 4224;              // score scratch of the extra-query warp (floats)
constexpr int APP_BAR_BYTES = 512;
constexpr int APP_XCH_BYTES = 2 * 2 * 128 * 8 /* (row sum, extra weight) [X][unit parity][row] */ + 2 * 2 * 128 * 4 /* extra key's score [X][unit parity][row] */;
constexpr int APP_SMEM_BYTES = 1024 /* alignment slack */ + (2 * APP_Q_STAGES + APP_K_STAGES + APP_V_STAGES + 2 /* extra-query staging */) * APP_TILE_BYTES +
                               APP_BAR_BYTES + APP_MAX_EXTRA_KEYS * 4 + 64 * 4 /* extra query row */ + APP_XCH_BYTES +
                               APP_XR_STAGES * 256 /* extra key row | extra value row of a unit */;
constexpr uint32_t APP_TMEM_COLS = 512;
constexpr uint32_t APP_COL_S = 0, APP_COL_P = 256, APP_COL_O = 384;   // + X * 128 / 64 / 64 for warpgroup X

#ifndef VFM_APP_POLY
#define VFM_APP_POLY 0      // n > 0: one exponential in n is evaluated on the FMA pipe (poly_exp2) instead of MUFU
#endif
#ifndef VFM_APP_ALUPACK
#define VFM_APP_ALUPACK 0   // 1: pack P with two integer adds + PRMT instead of F2FP (measured slower: 0.325 vs 0.301 ms)
#endif                      // 2: one PRMT (truncation) of exponentials pre-scaled by 1 + kTruncEps (see kTruncEps)
// ALUPACK 2: P = trunc_bf16(e (1 + eps)) with the factor folded into the exponent argument (free), so the truncation's
// mean loss of half a bf16 ulp (2^-8 / (1 + f) relative, f the mantissa fraction) is cancelled in the mean: weighted by
// the value itself E[ulp / 2] / E[x] = 2^-8 / 1.5. The fp32 row sums are accumulated from the SAME pre-scaled values and
// divided by 1 + eps in the epilogue. Error per weight: uniform in about (-0.5, 0.5] ulp like round-to-nearest.
// Measured (tools/micro/xu_share_bench.cu): F2FP does NOT share the MUFU pipe (ex2 8.09 clk / warp instruction with or
// without a co-resident F2FP stream at 2.08), so the variant only saves issue slots: 0.299 vs 0.297 ms, not the default.
constexpr float kTruncEps = 0.00390625f / 1.5f;
constexpr float kTruncLog2 = 0.0037521f;   // log2(1 + kTruncEps)
#ifndef VFM_APP_BACKOFF
#define VFM_APP_BACKOFF 0   // 1: the TMA / output warps sleep between polls of their barriers
#endif
#ifndef VFM_APP_BACKOFF_NS
#define VFM_APP_BACKOFF_NS 32
#endif
#if VFM_APP_BACKOFF
#define APP_WAIT_BG(bar, parity) mbar_wait_backoff(bar, parity, VFM_APP_BACKOFF_NS)
#else
#define APP_WAIT_BG mbar_wait
#endif
#ifndef VFM_APP_SCORES_ON_OUTPUT
#define VFM_APP_SCORES_ON_OUTPUT 0   // the extra key's scores are computed by the softmax warps at unit start (0) or, a unit
#endif                               // ahead, by the output warps (1: measured slower, 0.277 against 0.263 ms)
#ifndef VFM_APP_EARLY_PROBE
#define VFM_APP_EARLY_PROBE 0   // 1: probe the consumer pass's two barriers (non-blocking test_wait) from the end of the MUFU pass
#endif                          // (measured slower: 0.266 against 0.260 ms per 36-window launch, same box — the probes lengthen the exclusive pass; probes from the
                                // MIDDLE of the pass, with or without a fake dependence pinning them there: 552-556 against 600 TFLOP/s)
#ifndef VFM_APP_HANDOFF
#define VFM_APP_HANDOFF 1   // the two softmax warpgroups take turns on the MUFU pipe (named-barrier token):
#endif                      // 1 = one token per warpgroup pair (256 threads), 2 = one per SM sub-partition (warp pair, 64 threads)
// setmaxnreg moves registers inside the pool the CTA was LAUNCHED with (threads x the register count ptxas reports for
// the kernel), not the 64 K of the SM; a sum above the pool dead-locks the last warp's setmaxnreg.inc.
#ifndef VFM_APP_SOFTMAX_REGS
#define VFM_APP_SOFTMAX_REGS 200
#define VFM_APP_SERVICE_REGS 56
#endif
// VFM_APP_OUT32 = 1: the output warps read O in two 32-column rounds and hand the accumulator back before their stores. Measured
// with the register budget moved their way (softmax 192 / service 64: 586, 184 / 72: 592 TFLOP/s) against 595.5 as shipped: the
// unit output is not what bounds the kernel.
#ifndef VFM_APP_OUT32
#define VFM_APP_OUT32 0
#endif
#ifndef VFM_APP_SUM_IN_PASS
#define VFM_APP_SUM_IN_PASS 16  // n > 0: the FADD2 row sums of pair i - n are issued inside the MUFU pass at iteration i (n = 16: 256 clk behind)
#endif
#ifndef VFM_APP_PACK_IN_PASS
#define VFM_APP_PACK_IN_PASS 1   // 1 (with VFM_APP_SUM_IN_PASS > 0): those pairs are also packed to bf16 inside the pass, compacted in place.
                                 // Measured (36 windows, same box): neither 596, sums only 600-603, sums + packs at distance 12 / 16 / 20 / 24: 591 / 618-625 / 620 / 586 TFLOP/s
#endif
#if VFM_APP_PACK_IN_PASS && (!VFM_APP_SUM_IN_PASS || VFM_APP_ALUPACK || VFM_APP_POLY)
#error "VFM_APP_PACK_IN_PASS needs VFM_APP_SUM_IN_PASS > 0 and the default pack / exponential"
#endif
#ifndef VFM_APP_ONE_ISSUER
#define VFM_APP_ONE_ISSUER 0   // 1: one warp issues S_X(t+2) then PV_X(t) behind one commit; the softmax warp waits for / arrives on one barrier per tile
                               // (correct at the first run; measured slower, 588 against 610 TFLOP/s on that box: the scores wait for the PV behind them)
#endif
#if VFM_APP_ONE_ISSUER && VFM_APP_EARLY_PROBE
#error "VFM_APP_EARLY_PROBE probes the two-barrier protocol"
#endif
#ifndef VFM_APP_ELECT_WAIT
#define VFM_APP_ELECT_WAIT 0   // 1: lane 0 alone polls the consumer pass's two mbarriers, the warp follows through __syncwarp
#endif
static_assert(256 * VFM_APP_SOFTMAX_REGS + 256 * VFM_APP_SERVICE_REGS <= 512 * 128, "setmaxnreg budget exceeds the CTA's register pool");

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


// See the file header for the roles. Barrier protocol (all mbarriers; parity = use count of the slot & 1):
//   q_full / q_empty [2]   TMA -> S issuer, softmax (extra key) / S issuer (last S of the unit) [+ 8 softmax warps] -> TMA
//   k_full / k_empty [3]   TMA -> S issuer / S issuer (S_A, S_B of the tile executed) -> TMA;   v_full / v_empty likewise (PV)
//   s_full [X] / s_free [X]  S issuer -> softmax (S_X(t) is in TMEM) / softmax -> S issuer (S_X(t) is in registers)
//   p_full [X] / p_free [X]  softmax -> PV issuer (P_X(t) stored) / PV issuer -> softmax (PV_X(t) executed)
//   l_full [unit parity][X][quad]   softmax warp -> output warp   the unit's row sums are in shared memory. TWO barrier sets,
//                          alternating by unit: with one key tile per unit a softmax warp can finish unit k + 1 before the
//                          output warp has looked at unit k (its P store only needs PV(k), which needs the output of
//                          k - 1), and a single barrier's parity would then alias (observed as a hang). e_full likewise.
//   o_ready [X] / o_free [X]  PV issuer -> output warps (last PV_X of the unit executed) / output warps -> PV issuer (O_X copied out)
// The extra key / value rows live in a ring of APP_XR_STAGES = 8 units: the TMA warp refills slot k % 8 for unit k after
// the softmax warps have entered unit k - 2, which implies (through p_free -> o_free) that unit k - 5 has been written
// out even with one key tile per unit; an output-warp arrival on q_empty instead would dead-lock that case (the
// output of unit k would wait for the last PV of unit k while the Q tile of unit k + 2 waits for the output).
__global__ void __launch_bounds__(APP_THREADS, 1)
attention_pp_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
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
  uint64_t* s_full = v_empty + APP_V_STAGES;        // [2]
  uint64_t* s_free = s_full + 2;                    // [2]
  uint64_t* p_full = s_free + 2;                    // [2]
  uint64_t* p_free = p_full + 2;                    // [2]
  uint64_t* o_ready = p_free + 2;                   // [2]
  uint64_t* o_free = o_ready + 2;                   // [2]
  uint64_t* x_full = o_free + 2;                    // [2] TMA -> extra-query warp
  uint64_t* l_full = x_full + 2;                    // [unit parity][2][4]
  uint64_t* e_full = l_full + 16;                   // [unit parity][2][4] output warp -> softmax warp: the unit's extra-key scores are in shared memory
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(e_full + 16);
  static_assert((2 * APP_Q_STAGES + 2 * APP_K_STAGES + 2 * APP_V_STAGES + 7 * 2 + 32) * 8 + 8 <= APP_BAR_BYTES, "barrier block too small");
  float* extra_sc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + APP_BAR_BYTES);   // [APP_MAX_EXTRA_KEYS] (+ 64 for q)
  float2* lw_smem = reinterpret_cast<float2*>(extra_sc + APP_MAX_EXTRA_KEYS + 64);   // [X][unit parity][row]: (row sum, extra key's weight)
  float* es_smem = reinterpret_cast<float*>(lw_smem + 2 * 2 * 128);        // [X][unit parity][row]: q_row . k_extra (log2 units)
  uint8_t* smem_xr = reinterpret_cast<uint8_t*>(es_smem + 2 * 2 * 128);    // [APP_XR_STAGES][extra key row | extra value row] (64 bf16 each)

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
    for (int s = 0; s < APP_Q_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], p.extra ? (VFM_APP_SCORES_ON_OUTPUT ? 5 : 9) : 1); }
    for (int s = 0; s < APP_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < APP_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1); mbar_init(&s_free[x], 4); mbar_init(&p_full[x], 4); mbar_init(&p_free[x], 1);
      mbar_init(&o_ready[x], 1); mbar_init(&o_free[x], 4); mbar_init(&x_full[x], 1);
    }
    for (int i = 0; i < 16; ++i) { mbar_init(&l_full[i], 1); mbar_init(&e_full[i], 1); }
    fence_barrier_init();
  }
  pdl_launch_dependents();
  if (warp == 9) tmem_alloc<APP_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // programmatic dependent launch: the qkv rows come from the previous kernel (see sm100_ptx.cuh)
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
#if VFM_APP_ONE_ISSUER
    } else if (warp == 9) {
      // ===================== one MMA issuer: after consumer pass tau of warpgroup X, S_X(tau + 2) then PV_X(tau), ONE commit =====================
      // The softmax warp then waits for a single barrier per tile ("ready": s_full[x]) instead of s_full + p_free, and arrives on a
      // single one ("done": s_free[x]) instead of p_full + s_free. ready[x] completes once at the start (S_X(0)) and once per step n:
      // step n follows done[x] #n (n = 0: S_X(0) is in registers; n = tau + 1: consumer pass tau is over) and issues S_X(n + 1), PV_X(n - 1).
      constexpr uint32_t idesc_s = make_idesc_bf16(APP_TILE_Q, APP_BLOCK_KV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(APP_TILE_Q, ATT_D, 0, 1);   // B = V is MN-major
      const uint64_t dq0 = make_sw128_desc(smem_u32(smem_q));
      const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
      const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
      const int tail_ksteps = (tail_valid + 15) >> 4;
      auto issue_s = [&](int ts, int x) {   // S_X(ts); the warp has waited for what it needs
        const int k = ts / kv_tiles, j = ts - k * kv_tiles;
        const int qs = k % APP_Q_STAGES, ks = ts % APP_K_STAGES;
        if (x == 0) {
          if (j == 0) mbar_wait(&q_full[qs], (k / APP_Q_STAGES) & 1);
          mbar_wait(&k_full[ks], (ts / APP_K_STAGES) & 1);
          tc_fence_after();
        }
        if (elect_one_sync()) {
          const uint64_t dq = dq0 + static_cast<uint64_t>((2 * qs + x) * (APP_TILE_BYTES >> 4));
          const uint64_t dk = dk0 + static_cast<uint64_t>(ks * (APP_TILE_BYTES >> 4));
          const uint32_t tmem_s = tmem_base + APP_COL_S + x * APP_BLOCK_KV;
#pragma unroll
          for (int kk = 0; kk < ATT_D / 16; ++kk) umma_ss(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
          if (x == 1) {
            tc_commit(&k_empty[ks]);
            if (j == kv_tiles - 1) tc_commit(&q_empty[qs]);
          }
        }
        __syncwarp();
      };
      auto issue_pv = [&](int tp, int x) {
        const int k = tp / kv_tiles, j = tp - k * kv_tiles;
        const int vs = tp % APP_V_STAGES;
        const bool last = j == kv_tiles - 1;
        const int ksteps = last ? tail_ksteps : APP_BLOCK_KV / 16;
        if (x == 0) mbar_wait(&v_full[vs], (tp / APP_V_STAGES) & 1);
        if (j == 0 && k > 0) mbar_wait(&o_free[x], (k - 1) & 1);   // the previous unit's O_X has been copied out
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t dv = dv0 + static_cast<uint64_t>(vs * (APP_TILE_BYTES >> 4));
          const uint32_t tmem_o = tmem_base + APP_COL_O + x * ATT_D;
          const uint32_t tmem_p = tmem_base + APP_COL_P + x * (APP_BLOCK_KV / 2);
          for (int kk = 0; kk < ksteps; ++kk) umma_ts(tmem_o, tmem_p + 8 * kk, dv + 128 * kk, idesc_pv, (j | kk) != 0);
          if (last) tc_commit(&o_ready[x]);
          if (x == 1) tc_commit(&v_empty[vs]);
        }
        __syncwarp();
      };
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        issue_s(0, x);
        if (elect_one_sync()) tc_commit(&s_full[x]);   // ready #0
        __syncwarp();
      }
      for (int n = 0; n <= total_tiles; ++n) {
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          mbar_wait(&s_free[x], n & 1);   // done[x] #n
          tc_fence_after();
          if (n + 1 < total_tiles) issue_s(n + 1, x);
          if (n >= 1) issue_pv(n - 1, x);
          if (elect_one_sync()) tc_commit(&s_full[x]);   // ready #(n + 1): covers both
          __syncwarp();
        }
      }
    } else if (warp == 10) {
      // idle in this variant
#else
    } else if (warp == 9) {
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
    } else if (warp == 10) {
      // ===================== PV issuer: O_X += P_X(t) V(t), X = A then B =====================
      constexpr uint32_t idesc_pv = make_idesc_bf16(APP_TILE_Q, ATT_D, 0, 1);   // B = V is MN-major
      const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
      const int tail_ksteps = (tail_valid + 15) >> 4;
      int t = 0;
      for (int k = 0; k < n_my; ++k) {
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int vs = t % APP_V_STAGES;
          const bool last = j == kv_tiles - 1;
          const int ksteps = last ? tail_ksteps : APP_BLOCK_KV / 16;
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
              if (last) tc_commit(&o_ready[x]);
              if (x == 1) tc_commit(&v_empty[vs]);
            }
            __syncwarp();
            APP_TRACE(3, t, 3 * x + 2);
          }
        }
      }
#endif
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
    // ===================== softmax: warpgroup X = A (warps 0..3) or B (warps 4..7) =====================
    // Per key tile t (one query row per thread, its 128 scores in registers s[]):
    //   reference check   m_ref moves (and O_X is rescaled) only when the tile max exceeds it by > 2^24
    //   token             named barrier: the MUFU pipe is ours
    //   MUFU pass         64 FFMA2 + 128 MUFU.EX2, in place in s[], 8 clk / MUFU instruction; the pass's spare issue slots take the FADD2 row
    //                     sums and the F2FP packs of the pairs exponentiated 16 iterations (256 clk) earlier (VFM_APP_SUM_IN_PASS /
    //                     VFM_APP_PACK_IN_PASS: consumers directly behind their exponential would stall the in-order warp)
    //   token release
    //   consumer pass     FADD2 row sums + F2FP packs + 4 x tcgen05.st of P(t) — and, chunk by chunk into the registers
    //                     this frees, the tcgen05.ld of S(t+1), which the S issuer finished long ago.
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
#if VFM_APP_HANDOFF == 2
    const int bar_mine = 1 + 2 * quad + x, bar_other = 1 + 2 * quad + (x ^ 1);
    constexpr int kBarThreads = 64;
#else
    const int bar_mine = 1 + x, bar_other = 2 - x;
    constexpr int kBarThreads = 256;
#endif

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
    // MUFU passes alternate A(t), B(t), A(t+1), ...: a warp enters its pass through a named barrier (its own bar.sync +
    // the bar.arrive of its partner at the end of ITS pass). B opens the first A pass here. Without the token the two
    // warpgroups drift into phase: both in their exponentials (sharing the pipe), then both in their bookkeeping (pipe
    // idle) — 3300 clk per 256 x 128 scores in the first traces. HANDOFF 2: the token is per SM sub-partition (the A and B
    // warps of one TMEM lane quadrant share a scheduler and a MUFU unit; the four pairs need not wait for each other).
    if (x == 1) named_bar_arrive(bar_other, kBarThreads);
#endif
    int t = 0;
    for (int k = 0; k < n_my; ++k) {
      float m_ref = -INFINITY, w_extra = 0.f, l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;   // row sum = l0 + l1 + l2 + l3
      if (p.extra) {
        // the extra key starts the running softmax with weight 1; its score comes from the output warp of this quadrant
        if (quad == 0) APP_TRACE(x, t, 2);
#if VFM_APP_SCORES_ON_OUTPUT
        mbar_wait(&e_full[(k & 1) * 8 + x * 4 + quad], (k >> 1) & 1);
        m_ref = es_smem[(x * 2 + (k & 1)) * 128 + row];
#else
        {
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
        }
#endif
        w_extra = 1.f;
#if VFM_APP_ALUPACK != 2
        l0 = 1.f;
#endif
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
#if VFM_APP_ONE_ISSUER
              mbar_wait(&s_full[x], (t + 1) & 1);   // ready #(t + 1): PV_X(t-1) (and S_X(t+1)) have executed
#else
              mbar_wait(&p_free[x], (t - 1) & 1);   // every PV_X up to tile t-1 has executed
#endif
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
        named_bar_sync(bar_mine, kBarThreads);   // the partner has issued the last exponential of its pass
#endif
        if (quad == 0) APP_TRACE(x, t, 4);
#if VFM_APP_ALUPACK == 2
        const float neg_m = kTruncLog2 - m_ref;
#else
        const float neg_m = -m_ref;
#endif
        // The two passes sit under branches on bits of a kernel parameter (AttParams::sched_mask, all ones at run time).
        // Left alone, ptxas schedules every consumer of an exponential directly behind it (the FADD2 chain looks
        // critical to it) and the in-order warp sits out the MUFU latency once per pair: measured 28 clk per pair
        // instead of the 16 the pipe needs. A fake data dependence on a later exponential (round 1's trick) only
        // moves the wait, and warp-level barriers are scheduled across; real control flow is what ptxas respects.
        const bool more = t + 1 < total_tiles;
        bool ready = false;   // both barriers of the consumer pass were seen complete from inside the MUFU pass
        if (sched_mask & 1u) {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            float x0, x1;
            ffma2_bc(x0, x1, __uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]), kLog2e, neg_m);
            s[2 * i] = __float_as_uint(fast_exp2(x0));
            s[2 * i + 1] = __float_as_uint((kPoly != 0 && (i % (kPoly ? kPoly : 1)) == 0) ? poly_exp2(x1) : fast_exp2(x1));
#if VFM_APP_SUM_IN_PASS
            // row sums of the pairs exponentiated VFM_APP_SUM_IN_PASS iterations (x 16 clk) ago: ready for sure, and the pass has
            // issue slots to spare (one MUFU per 8 clk); the last pairs are added in the consumer pass. Same pairing as there.
            if (i >= VFM_APP_SUM_IN_PASS) {
              constexpr int kD = VFM_APP_SUM_IN_PASS;
              const int kk = i - kD;
              if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
              else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#if VFM_APP_PACK_IN_PASS
              // ... and their bf16 pack, compacted in place: packed pair kk lives in s[kk] (the exponentials s[kk] held belong to
              // pair kk / 2 <= kk, summed and packed before), so P chunk c is the 16 consecutive registers s[16 c ..] for tcgen05.st
              s[kk] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#endif
            }
#endif
          }
#if VFM_APP_EARLY_PROBE
          // A completed mbarrier wait still costs ~100 clk of round trip, twice, at the head of the consumer pass — on the
          // chain that decides the tile period. Probe both barriers here instead (non-blocking test_wait), where the warp
          // issues one instruction in eight: the parity operands carry a fake dependence on two of the LAST exponentials
          // (their sign bit, always 0), so ptxas cannot hoist the probes to the top of the pass, where S(t+1) / PV(t-1)
          // have not completed yet. A miss falls back to the blocking waits below.
          const bool ok_s = !more || mbar_test_wait(&s_full[x], sign_gate(static_cast<uint32_t>((t + 1) & 1), s[118]));
          const bool ok_p = t == 0 || mbar_test_wait(&p_free[x], sign_gate(static_cast<uint32_t>((t - 1) & 1), s[119]));
          ready = ok_s && ok_p;
#endif
        }
        if (sched_mask & 2u) {
#if VFM_APP_HANDOFF
          named_bar_arrive(bar_other, kBarThreads);   // the pipe goes to the partner
#endif
          if (quad == 0) APP_TRACE(x, t, 7);
          if (!ready) {
#if VFM_APP_ELECT_WAIT
            if (lane == 0) {   // one lane polls, the warp follows through __syncwarp (32 lanes polling one mbarrier are 32 shared-memory requests)
              if (more) mbar_wait(&s_full[x], (t + 1) & 1);
              if (t > 0) mbar_wait(&p_free[x], (t - 1) & 1);
            }
            __syncwarp();
#elif VFM_APP_ONE_ISSUER
            if (more || t > 0) mbar_wait(&s_full[x], (t + 1) & 1);   // ready #(t + 1): S_X(t+1) is in TMEM and PV_X(t-1) has read P_X
#else
            if (more) mbar_wait(&s_full[x], (t + 1) & 1);   // S_X(t+1): issued as soon as S_X(t) had been copied out, a tile ago
            if (t > 0) mbar_wait(&p_free[x], (t - 1) & 1);   // PV_X(t-1) has read P_X: the buffer may be rewritten
#endif
          }
          tc_fence_after();
          if (quad == 0) APP_TRACE(x, t, 1);
#if VFM_APP_PACK_IN_PASS
          // the pairs the pass did not get to (the last VFM_APP_SUM_IN_PASS), then P leaves from s[0 .. 63] and S(t+1) arrives:
          // chunk 0 into s[0 .. 31] once P chunks 0 and 1 have been stored from there, chunk 1 into s[32 .. 63] after P chunks 2 and 3
#if VFM_APP_PACK_IN_PASS == 2
          // variant: the first 32 packed pairs leave (and score columns 64 .. 95 arrive, into registers the pass has already
          // consumed when VFM_APP_SUM_IN_PASS <= 16) before the leftover pairs are summed and packed; P leaves as two x32 stores
          tmem_st32(tmem_p + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          if (more) {
            tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
            tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          }
#pragma unroll
          for (int kk = 64 - VFM_APP_SUM_IN_PASS; kk < 64; ++kk) {
            if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
            else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
            s[kk] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
          }
          tmem_st32(tmem_p + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
          if (more) {
            tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
            tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
          }
#else
#pragma unroll
          for (int kk = 64 - VFM_APP_SUM_IN_PASS; kk < 64; ++kk) {
            if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
            else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
            s[kk] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
          }
          if (more) {   // the upper half of s[] is free already
            tmem_ld32(tmem_s + 64, *reinterpret_cast<uint32_t(*)[32]>(&s[64]));
            tmem_ld32(tmem_s + 96, *reinterpret_cast<uint32_t(*)[32]>(&s[96]));
          }
          tmem_st16(tmem_p + 0, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
          tmem_st16(tmem_p + 16, *reinterpret_cast<uint32_t(*)[16]>(&s[16]));
          if (more) tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
          tmem_st16(tmem_p + 32, *reinterpret_cast<uint32_t(*)[16]>(&s[32]));
          tmem_st16(tmem_p + 48, *reinterpret_cast<uint32_t(*)[16]>(&s[48]));
          if (more) tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
#endif
#else
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int kk = 16 * c + i;
#if VFM_APP_SUM_IN_PASS
              if (kk >= 64 - VFM_APP_SUM_IN_PASS) {
#endif
              if (kk & 1) fadd2_acc(l2, l3, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
              else fadd2_acc(l0, l1, __uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#if VFM_APP_SUM_IN_PASS
              }
#endif
#if VFM_APP_ALUPACK == 2
              pk[i] = __byte_perm(s[2 * kk], s[2 * kk + 1], 0x7632u);
#elif VFM_APP_ALUPACK
              pk[i] = pack_bf16x2_alu(s[2 * kk], s[2 * kk + 1]);
#else
              pk[i] = pack_bf16x2(__uint_as_float(s[2 * kk]), __uint_as_float(s[2 * kk + 1]));
#endif
            }
            tmem_st16(tmem_p + c * 16, pk);
            if (more) tmem_ld32(tmem_s + 32 * c, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));   // into the registers just consumed
          }
#endif
          if (quad == 0) APP_TRACE(x, t, 5);
          tmem_st_wait();
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
#if VFM_APP_ONE_ISSUER
            mbar_arrive(&s_free[x]);             // done #(t + 1): P_X(t) is stored, S_X(t+1) is in registers
#else
            mbar_arrive(&p_full[x]);
            if (more) mbar_arrive(&s_free[x]);   // S_X(t+2) may overwrite the score columns now
#endif
          }
          if (more) {   // the row max of tile t+1 is what stands between this warp and its next token request
            const int jn = j + 1 == kv_tiles ? 0 : j + 1;
            m_next = row_max(jn == kv_tiles - 1 ? tail_valid : APP_BLOCK_KV);
          }
          if (quad == 0) APP_TRACE(x, t, 6);
        }
      }

      // ---- the unit's row sums go to the output warp of this quadrant, which writes the unit out
#if VFM_APP_ALUPACK == 2
      const float l_total = ((l0 + l1) + (l2 + l3)) * (1.f / (1.f + kTruncEps)) + w_extra;
#else
      const float l_total = (l0 + l1) + (l2 + l3);
#endif
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
