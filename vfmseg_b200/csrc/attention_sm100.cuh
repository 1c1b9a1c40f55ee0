// Fused multi-head attention (non-causal, head_dim 64) on tcgen05 for sm_100a.
//
// Replaces rein/models/backbones/dino_layers/attention.py:56-66 (reshape/permute of the fused
// qkv output, q*scale, q@k^T, softmax, @v, transpose back) — the N x N score tensor never
// leaves the SM — and, for the refinement head, the q/k/v attention core of
// rein/models/heads/Transformer.py:113-136 (self- and cross-attention, 8 heads x 64).
// The 1/sqrt(d) scale is folded into W_q/b_q on the host (exact: 0.125 = 2^-3).
//
// Operands: Q [rows, *] / K [rows, *] / V [rows, *] bf16 matrices addressed through three TMA maps (for the ViT
// all three are the packed qkv GEMM output [n_seq * seq_len, 3 * heads * 64], column = which * C + head * 64 + d,
// exactly what the qkv Linear emits, attention.py:58). Output [rows, heads * 64] bf16 (attention.py:66).
//
// Unit of work = one (sequence, head, 128-query tile); 64-key tiles. Persistent CTAs (grid = min(units, 2 x SMs), two
// per SM) run their units back to back as one stream of key tiles. Per key tile t:
//   S_t = Q K_t^T      tcgen05.mma SS, fp32 scores in TMEM buffer t % 2 — issued two tiles ahead
//   softmax            four warps, one query row per thread (one TMEM lane): the row is copied to 64 registers, running
//                      max kept in registers, P = exp2(s - m) truncated to bf16 and written over the first 32 columns
//                      of S_t
//   O += P_t V_t       tcgen05.mma TS (A = P from TMEM, B = V tile as an MN-major SWIZZLE_128B operand); O stays in TMEM
//                      for the whole KV loop and is rescaled lazily, only when a row's running max grows by more than
//                      2^24; the row sums come from the tensor core too (L += P_t * ones)
// The next unit's Q / K tiles are fetched and its first two score MMAs issued while the softmax warps finish the current
// unit; the only bubble between units is the O hand-over (o_free). What bounds the kernel, and every variant that was
// measured against this one (de-phasing the two CTAs of an SM, range guard after the exponentials, three score
// buffers, fp16 P, eight softmax warps per CTA, P in its own TMEM columns): profiles/r1_attention_timeline.txt.
//
// "Extra token" mode (ViT sequences = 1 cls token + 2^k patch tokens, e.g. 1025): the tensor-core tiles cover the
// body tokens only (1024 = 8 query tiles, 16 key tiles, no tail tile); the cls KEY is handled by each softmax
// thread on the CUDA cores (one 64-long dot product per query row before the loop, one AXPY in the epilogue),
// and the cls QUERY row by a second small kernel (attention_extra_query_kernel). Without it the 1025th token costs a
// ninth query tile (+12.5 %) and a seventeenth key tile (+6 %).
//
// Roles (192 threads): warp 0 = TMA producer (Q double-buffered, K ring released right after S_t, V ring), warp 1 =
// MMA issuer + TMEM allocator, warps 2..5 = softmax / output (TMEM lane quadrant = warp % 4).
#pragma once
#include <type_traits>

#include "sm100_ptx.cuh"

namespace vfm {

constexpr int ATT_BLOCK_Q = 128;
constexpr int ATT_BLOCK_KV = 64;
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;                           // warp 0 TMA, warp 1 MMA + TMEM allocator, warps 2..5 softmax
constexpr int ATT_K_STAGES = 4;                            // K_j is released as soon as S_j = Q K_j^T has executed
constexpr int ATT_V_STAGES = 4;                            // V_j is held until O += P_j V_j has executed
constexpr int ATT_Q_BYTES = ATT_BLOCK_Q * ATT_D * 2;       // 16 KB
constexpr int ATT_KV_BYTES = ATT_BLOCK_KV * ATT_D * 2;     //  8 KB
constexpr int ATT_ONES_BYTES = ATT_KV_BYTES;               // a V-shaped tile of bf16 1.0 (row sums on the tensor core)
constexpr int ATT_Q_STAGES = 2;                            // the next unit's Q tile is fetched while the current one runs
constexpr int ATT_SMEM_BYTES = ATT_Q_STAGES * ATT_Q_BYTES + (ATT_K_STAGES + ATT_V_STAGES) * ATT_KV_BYTES + ATT_ONES_BYTES + 1024 + 256;
constexpr uint32_t ATT_TMEM_COLS = 256;
// TMEM columns: S0 [0,64) S1 [64,128) fp32 scores; the packed bf16 probabilities P_b overwrite the first 32 columns of
// S_b (each thread has its S row in registers by then); O [128,192); L [192,208) = row sums of P, accumulated by a
// second small MMA against a tile of ones (every column holds the same sum).
constexpr uint32_t ATT_COL_S = 0, ATT_COL_O = 128, ATT_COL_L = 192;

#ifndef VFM_ATT_POLY
#define VFM_ATT_POLY 0   // measured: 0 -> 0.169 ms, 4 -> 0.172, 2 -> 0.179, 1 -> 0.195 (18 windows x 16 heads x 1025)
#endif

struct AttParams {
  int q_len, kv_len;            // body tokens per sequence (handled by the tensor-core tiles)
  int q_seq_rows, kv_seq_rows;  // rows between consecutive sequences in the Q / K,V matrices
  int q_row_off, kv_row_off;    // first body row inside a sequence (1 in extra-token mode)
  int heads, q_tiles;
  int q_col0, k_col0, v_col0;   // column of (head 0, d 0) in the Q / K / V matrices
  int extra;                    // 1: row 0 of each K/V sequence is one more key, handled on the CUDA cores
  int n_units;                  // (sequence, head, query tile) units, dealt round-robin to the persistent CTAs
  int sched_mask;               // all ones; the ping-pong kernel branches on its bits to pin its instruction schedule
  const __nv_bfloat16* q_ptr; int q_ld;
  const __nv_bfloat16* k_ptr;   // raw pointers for the extra key / value row
  const __nv_bfloat16* v_ptr;
  int k_ld, v_ld;
  __nv_bfloat16* out; int out_ld;
};

#ifdef VFM_EPI_TIMING
// [slot][tile][event]: the two co-resident first-wave CTAs of SM 5 (slot = parity of a per-SM ticket); events 0..7 MMA
// warp, 8..15 softmax warp with quad == 2. Same SM -> same clock, so the two timelines can be laid side by side
// (tools/att_trace.py).
__device__ long long g_att_trace[2][20][16];
__device__ unsigned int g_att_ticket[1024];
#define ATT_TRACE(j, ev) do { if (trace_slot >= 0 && lane == 0 && (j) < 20) g_att_trace[trace_slot][j][ev] = clock64(); } while (0)
__device__ unsigned long long g_att_dbg[8];
#else
#define ATT_TRACE(j, ev)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// The query row of the extra token (ViT cls token) against all kv_len + 1 keys of its sequence, on the CUDA cores:
// 2 x 64 x (kv_len + 1) MACs per (sequence, head) — 1/1025 of the attention work, not worth a 128-row tensor tile.
// Runs in the CTAs appended behind the tensor-tile units of the same grid (they land in the partly empty last wave).
// 8 lanes per key (one 16-byte chunk of the 128-byte K / V row each), so every warp load covers 4 whole rows.
// smem: kv_total floats (scores, then weights) + [warps][64] floats (partial outputs).
__device__ __forceinline__ void attention_extra_query(const AttParams& p, int seq, int head, float* sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = ATT_THREADS / 32;
  const int kv_total = p.kv_len + 1;
  float* sc = sm;                                   // [kv_total]
  float* part = sm + ((kv_total + 3) & ~3);         // [kWarps][64]
  float* red = part + kWarps * 64;                  // [kWarps]
  const int sub = lane >> 3, chunk = lane & 7;      // key within the group of 4, 16-byte chunk of the row
  const size_t kv_row = static_cast<size_t>(seq) * p.kv_seq_rows;
  const uint4* kb = reinterpret_cast<const uint4*>(p.k_ptr + kv_row * p.k_ld + p.k_col0 + head * ATT_D) + chunk;
  const uint4* vb = reinterpret_cast<const uint4*>(p.v_ptr + kv_row * p.v_ld + p.v_col0 + head * ATT_D) + chunk;
  const size_t k_pitch = static_cast<size_t>(p.k_ld) / 8, v_pitch = static_cast<size_t>(p.v_ld) / 8;   // in uint4
  float qf[8];
  {
    const uint4 qv = __ldg(reinterpret_cast<const uint4*>(p.q_ptr + static_cast<size_t>(seq) * p.q_seq_rows * p.q_ld + p.q_col0 + head * ATT_D) + chunk);
    qf[0] = bf16lo(qv.x); qf[1] = bf16hi(qv.x); qf[2] = bf16lo(qv.y); qf[3] = bf16hi(qv.y);
    qf[4] = bf16lo(qv.z); qf[5] = bf16hi(qv.z); qf[6] = bf16lo(qv.w); qf[7] = bf16hi(qv.w);
  }
  // Keys are visited in batches of kIt * kWarps * 4; all loads of a batch are issued before the first use (indices
  // past the end are clamped, their results discarded), so a batch costs one memory round trip, not kIt.
  constexpr int kIt = 8;
  // scores (already in log2 units)
  float m = -INFINITY;
  for (int base = 0; base < kv_total; base += kIt * kWarps * 4) {
    uint4 kv[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int key = min(base + (it * kWarps + warp) * 4 + sub, kv_total - 1);
      kv[it] = __ldg(kb + static_cast<size_t>(key) * k_pitch);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int key = base + (it * kWarps + warp) * 4 + sub;
      float a = qf[0] * bf16lo(kv[it].x);
      a = fmaf(qf[1], bf16hi(kv[it].x), a); a = fmaf(qf[2], bf16lo(kv[it].y), a); a = fmaf(qf[3], bf16hi(kv[it].y), a);
      a = fmaf(qf[4], bf16lo(kv[it].z), a); a = fmaf(qf[5], bf16hi(kv[it].z), a); a = fmaf(qf[6], bf16lo(kv[it].w), a);
      a = fmaf(qf[7], bf16hi(kv[it].w), a);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a *= 1.4426950408889634f;
      if (key < kv_total) {
        if (chunk == 0) sc[key] = a;
        m = fmaxf(m, a);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int w = 1; w < kWarps; ++w) m = fmaxf(m, red[w]);
  // weighted values; each lane also accumulates the weights of its keys (every key is seen by 8 lanes)
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float l = 0.f;
  for (int base = 0; base < kv_total; base += kIt * kWarps * 4) {
    uint4 vv[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int key = min(base + (it * kWarps + warp) * 4 + sub, kv_total - 1);
      vv[it] = __ldg(vb + static_cast<size_t>(key) * v_pitch);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int key = base + (it * kWarps + warp) * 4 + sub;
      const float pw = key < kv_total ? fast_exp2(sc[min(key, kv_total - 1)] - m) : 0.f;
      l += pw;
      acc[0] = fmaf(pw, bf16lo(vv[it].x), acc[0]); acc[1] = fmaf(pw, bf16hi(vv[it].x), acc[1]);
      acc[2] = fmaf(pw, bf16lo(vv[it].y), acc[2]); acc[3] = fmaf(pw, bf16hi(vv[it].y), acc[3]);
      acc[4] = fmaf(pw, bf16lo(vv[it].z), acc[4]); acc[5] = fmaf(pw, bf16hi(vv[it].z), acc[5]);
      acc[6] = fmaf(pw, bf16lo(vv[it].w), acc[6]); acc[7] = fmaf(pw, bf16hi(vv[it].w), acc[7]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
  __syncthreads();   // red[] (max) has been read by everyone
  if (sub == 0) {
#pragma unroll
    for (int e = 0; e < 8; ++e) part[warp * 64 + chunk * 8 + e] = acc[e];
  }
  if (lane == 0) red[warp] = l;
  __syncthreads();
  if (tid < ATT_D) {
    float o = 0.f, lt = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) { o += part[w * 64 + tid]; lt += red[w]; }
    p.out[static_cast<size_t>(seq) * p.q_seq_rows * p.out_ld + head * ATT_D + tid] = __float2bfloat16(o / lt);
  }
}

// Extra-token mode only: one CTA per (sequence, head) computes the extra token's query row.
__global__ void __launch_bounds__(ATT_THREADS)
attention_extra_query_kernel(const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  attention_extra_query(p, blockIdx.x / p.heads, blockIdx.x % p.heads, reinterpret_cast<float*>(smem_raw));
}

// PERSISTENT: grid = min(units, 2 x SMs); CTA c runs units c, c + grid, c + 2 grid, ... as ONE stream of key tiles
// t = 0, 1, 2, ... (all ring indices, TMEM buffer indices and barrier phases are functions of t), so the next unit's Q
// and first K tiles are fetched and its first two score MMAs issued while the softmax warps finish the current unit. A
// non-persistent CTA spent ~3000 clk waiting for Q/K -> S_0 and ~3000 more in epilogue, exit and relaunch, out of ~33 k
// (profiles/r1_attention_timeline.txt, (4)). The only bubble left between units is the O hand-over: the first PV MMA of
// unit k+1 waits until the softmax warps have copied unit k's O and L out of TMEM (o_free).
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + ATT_Q_STAGES * ATT_Q_BYTES;
  uint8_t* smem_v = smem_k + ATT_K_STAGES * ATT_KV_BYTES;
  uint8_t* smem_ones = smem_v + ATT_V_STAGES * ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + ATT_ONES_BYTES);
  uint64_t* q_full = bars;                          // [2] TMA -> MMA, softmax (extra key)
  uint64_t* q_empty = q_full + ATT_Q_STAGES;        // [2] MMA (last S of the unit executed) [+ softmax in extra mode] -> TMA
  uint64_t* k_full = q_empty + ATT_Q_STAGES;        // [K stages] TMA -> MMA
  uint64_t* k_empty = k_full + ATT_K_STAGES;        // [K stages] MMA (S_t executed) -> TMA
  uint64_t* v_full = k_empty + ATT_K_STAGES;        // [V stages] TMA -> MMA
  uint64_t* v_empty = v_full + ATT_V_STAGES;        // [V stages] MMA (PV_t executed) -> TMA
  uint64_t* s_full = v_empty + ATT_V_STAGES;        // [2] MMA -> softmax   (S_t in TMEM buffer t % 2)
  uint64_t* p_full = s_full + 2;                    // [2] softmax -> MMA   (P_t in the columns of S_t)
  uint64_t* pv_done = s_full + 4;                   // [2] MMA -> softmax   (O += P_t V_t executed)
  uint64_t* o_free = s_full + 6;                    // [1] softmax -> MMA   (O, L of the finished unit copied out of TMEM)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 7);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int kv_tiles = (p.kv_len + ATT_BLOCK_KV - 1) / ATT_BLOCK_KV;
  const int first_unit = blockIdx.x, unit_step = gridDim.x;
  const int n_my = (p.n_units - first_unit + unit_step - 1) / unit_step;   // >= 1 (grid <= units)
  const int total_tiles = n_my * kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < ATT_Q_STAGES; ++s) { mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], p.extra ? 5 : 1); }
    for (int s = 0; s < ATT_K_STAGES; ++s) { mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1); }
    for (int s = 0; s < ATT_V_STAGES; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 4); mbar_init(&pv_done[s], 1); }
    mbar_init(o_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
#ifdef VFM_EPI_TIMING
  if (threadIdx.x == 64) {   // debug build: pick the two first-wave CTAs of SM 5 for the timeline
    uint32_t tr = 0xffffffffu, smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const uint32_t tk = atomicAdd(&g_att_ticket[smid & 1023], 1u) & 1u;
    if (smid == 5) tr = tk;
    tmem_slot[1] = tr;
  }
#endif
  for (int i = threadIdx.x; i < ATT_ONES_BYTES / 16; i += ATT_THREADS)
    reinterpret_cast<uint4*>(smem_ones)[i] = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  fence_proxy_async_smem();   // the tensor core reads the ones tile through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef VFM_EPI_TIMING
  const int trace_slot = static_cast<int>(tmem_slot[1]);
#endif

  // unit k of this CTA -> (sequence, head, query tile); consecutive units share (sequence, head), so the CTAs running
  // at the same time read the same K/V through L2
  struct Unit { int q_row0, kv_row0, head, seq, qt; };
  auto unit_of = [&](int k) {
    const int u = first_unit + k * unit_step;
    Unit r;
    r.qt = u % p.q_tiles;
    r.head = (u / p.q_tiles) % p.heads;
    r.seq = u / (p.q_tiles * p.heads);
    r.q_row0 = r.seq * p.q_seq_rows + p.q_row_off + r.qt * ATT_BLOCK_Q;
    r.kv_row0 = r.seq * p.kv_seq_rows + p.kv_row_off;
    return r;
  };

  if (warp == 0) {
    // ===================== TMA producer (warp converged, one elected lane issues) =====================
    // Two cursors over the tile stream (K runs two tiles ahead of V); no divisions on the per-tile path.
    int kt = 0, kk = 0, kj = 0;      // K cursor: stream index, unit, tile within the unit
    int vt = 0, vk = 0, vj = 0;      // V cursor
    Unit uk = unit_of(0), uv = uk;
    auto load_k = [&]() {
      if (kj == 0) {   // first K tile of a unit: its Q tile (two 64-row boxes) goes first
        if (kk > 0) uk = unit_of(kk);
        const int qs = kk % ATT_Q_STAGES;
        mbar_wait(&q_empty[qs], ((kk / ATT_Q_STAGES) & 1) ^ 1);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(&q_full[qs], ATT_Q_BYTES);
          tma_load_2d(smem_q + qs * ATT_Q_BYTES, &tmap_q, &q_full[qs], p.q_col0 + uk.head * ATT_D, uk.q_row0);
          tma_load_2d(smem_q + qs * ATT_Q_BYTES + ATT_KV_BYTES, &tmap_q, &q_full[qs], p.q_col0 + uk.head * ATT_D, uk.q_row0 + 64);
        }
        __syncwarp();
      }
      const int st = kt % ATT_K_STAGES;
      mbar_wait(&k_empty[st], ((kt / ATT_K_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&k_full[st], ATT_KV_BYTES);
        tma_load_2d(smem_k + st * ATT_KV_BYTES, &tmap_k, &k_full[st], p.k_col0 + uk.head * ATT_D, uk.kv_row0 + kj * ATT_BLOCK_KV);
      }
      __syncwarp();
      ++kt;
      if (++kj == kv_tiles) { kj = 0; ++kk; }
    };
    auto load_v = [&]() {
      if (vj == 0 && vk > 0) uv = unit_of(vk);
      const int st = vt % ATT_V_STAGES;
      mbar_wait(&v_empty[st], ((vt / ATT_V_STAGES) & 1) ^ 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&v_full[st], ATT_KV_BYTES);
        tma_load_2d(smem_v + st * ATT_KV_BYTES, &tmap_v, &v_full[st], p.v_col0 + uv.head * ATT_D, uv.kv_row0 + vj * ATT_BLOCK_KV);
      }
      __syncwarp();
      ++vt;
      if (++vj == kv_tiles) { vj = 0; ++vk; }
    };
    // load order follows the order of use: K_0, K_1, K_2, V_0, K_3, V_1, ...
    load_k();
    if (total_tiles > 1) load_k();
    for (int t = 0; t < total_tiles; ++t) {
      if (t + 2 < total_tiles) load_k();
      load_v();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp converged, one elected lane issues) =====================
    // S is double-buffered and issued two tiles ahead, so the exponentials of tile t+1 never wait for the PV MMA of
    // tile t. Issue order:  S_0, S_1, [PV_0, S_2], [PV_1, S_3], ...  — tcgen05.mma executes in issue order, which is
    // what lets P_t live in the columns of S_t: S_{t+2} cannot overwrite them before PV_t has read them. The stream
    // runs across unit boundaries: S_{t+2} may already belong to the next unit.
    // (Issuing S_{t+2} early, right after the softmax warps have copied S_t to registers, with a separate single P
    // buffer, measured SLOWER: 0.177 vs 0.167 ms — the S MMAs then compete with the exponential phase.)
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK_Q, ATT_D, 0, 1);   // B = V is MN-major
    constexpr uint32_t idesc_l = make_idesc_bf16(ATT_BLOCK_Q, 16, 0, 1);       // B = ones, 16 identical columns
    constexpr uint32_t idesc_s_full = make_idesc_bf16(ATT_BLOCK_Q, ATT_BLOCK_KV, 0, 0);
    const uint32_t tmem_o = tmem_base + ATT_COL_O;
    const uint32_t tmem_l = tmem_base + ATT_COL_L;
    const uint64_t dq0 = make_sw128_desc(smem_u32(smem_q));
    const uint64_t dk0 = make_sw128_desc(smem_u32(smem_k));
    const uint64_t dv0 = make_sw128_desc(smem_u32(smem_v));
    const uint64_t d1 = make_sw128_desc(smem_u32(smem_ones));
    // keys in the last tile of a sequence rounded up to 32 (the excess is masked by the softmax); all others are full
    const int tail_width = ((p.kv_len - (kv_tiles - 1) * ATT_BLOCK_KV) + 31) & ~31;
    const uint32_t idesc_s_tail = make_idesc_bf16(ATT_BLOCK_Q, tail_width, 0, 0);
    const int tail_ksteps = tail_width / 16;
    // S cursor (runs two tiles ahead of the PV cursor): stream index, unit, tile within the unit. No divisions here.
    int st_ = 0, sk = 0, sj = 0;
    auto wait_s_operands = [&]() {   // whole warp: K tile of the S cursor and, on a unit's first tile, its Q
      if (sj == 0) mbar_wait(&q_full[sk % ATT_Q_STAGES], (sk / ATT_Q_STAGES) & 1);
      mbar_wait(&k_full[st_ % ATT_K_STAGES], (st_ / ATT_K_STAGES) & 1);
    };
    auto issue_s_elected = [&]() {   // elected lane only: S at the cursor
      const int ks = st_ % ATT_K_STAGES, qs = sk % ATT_Q_STAGES;
      const bool last = sj == kv_tiles - 1;
      const uint32_t idesc_s = last ? idesc_s_tail : idesc_s_full;
      const uint64_t dq = dq0 + static_cast<uint64_t>(qs * (ATT_Q_BYTES >> 4));
      const uint64_t dk = dk0 + static_cast<uint64_t>(ks * (ATT_KV_BYTES >> 4));
      const uint32_t tmem_s = tmem_base + ATT_COL_S + (st_ & 1) * 64;
#pragma unroll
      for (int kk = 0; kk < ATT_D / 16; ++kk) umma_ss(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, kk != 0);
      tc_commit(&k_empty[ks]);       // K_t can be overwritten as soon as S_t has executed
      tc_commit(&s_full[st_ & 1]);
      if (last) tc_commit(&q_empty[qs]);   // last use of this unit's Q tile
    };
    auto advance_s = [&]() {         // whole warp (the cursor is warp-uniform state)
      ++st_;
      if (++sj == kv_tiles) { sj = 0; ++sk; }
    };
    for (int t = 0; t < 2 && t < total_tiles; ++t) {
      wait_s_operands();
      tc_fence_after();
      if (elect_one_sync()) issue_s_elected();
      __syncwarp();
      advance_s();
    }
    int k = 0, j = 0;                // PV cursor
    for (int t = 0; t < total_tiles; ++t) {
      const int b = t & 1;
      const uint32_t ph = (t >> 1) & 1;
      const int st = t % ATT_V_STAGES;
      const bool more = t + 2 < total_tiles;
      ATT_TRACE(t, 0);
      // one issue block per key tile: PV_t, the row-sum MMA and S_{t+2} behind a single fence / election (two separate
      // blocks cost ~630 clk each on this warp, mostly fixed overhead, against ~370 clk of tensor-pipe work per tile)
      if (more) wait_s_operands();
      mbar_wait(&v_full[st], (t / ATT_V_STAGES) & 1);
      if (j == 0 && k > 0) mbar_wait(o_free, (k - 1) & 1);   // the previous unit's O and L have been copied out
      mbar_wait(&p_full[b], ph);         // P_t stored (and O rescaled when the running max jumped); S_t is in registers
      tc_fence_after();
      ATT_TRACE(t, 1);
      if (elect_one_sync()) {
        const uint64_t dv = dv0 + static_cast<uint64_t>(st * (ATT_KV_BYTES >> 4));
        const uint32_t tmem_p = tmem_base + ATT_COL_S + b * 64;
        const int ksteps = j == kv_tiles - 1 ? tail_ksteps : ATT_BLOCK_KV / 16;
        for (int kk = 0; kk < ksteps; ++kk) {
          // A: 16 bf16 of P per step = 8 TMEM columns; B: 16 key rows of V = 2048 B
          umma_ts(tmem_o, tmem_p + 8 * kk, dv + 128 * kk, idesc_pv, (j | kk) != 0);
        }
        for (int kk = 0; kk < ksteps; ++kk) umma_ts(tmem_l, tmem_p + 8 * kk, d1 + 128 * kk, idesc_l, (j | kk) != 0);
        tc_commit(&v_empty[st]);
        tc_commit(&pv_done[b]);
        if (more) issue_s_elected();
      }
      __syncwarp();
      if (more) advance_s();
      if (++j == kv_tiles) { j = 0; ++k; }
      ATT_TRACE(t, 3);
    }
  } else {
    // ===================== softmax + output (warps 2..5): one query row per thread =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_tile = tmem_base + lane_base;
    const uint32_t tmem_o = tmem_tile + ATT_COL_O;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 24.0f;  // in log2 units: P stays <= 2^24 relative to the reference max (P is
                                                // bf16 and O/l are fp32: range, not precision, is what a large P costs)
    int t = 0;   // this CTA's tile stream position
#ifdef VFM_EPI_TIMING
    if (trace_slot == 0 && quad == 2 && lane == 0) {   // SM clock vs wall clock over this CTA's life (power capping)
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      g_att_dbg[0] = clock64(); g_att_dbg[1] = ns;
    }
#endif
    for (int k = 0; k < n_my; ++k) {
      const Unit un = unit_of(k);
      const int seq = un.seq, head = un.head, qt = un.qt;
      // m_ref: the max (times log2e) that P and the O accumulator in TMEM are currently relative to.
      float m_ref = -INFINITY, w_extra = 0.f;
      if (p.extra) {
        // the extra key: s = q_row . k_extra on the CUDA cores; it starts the running softmax with weight 1
        const int qs = k % ATT_Q_STAGES;
        mbar_wait(&q_full[qs], (k / ATT_Q_STAGES) & 1);
        const uint8_t* qrow = smem_q + qs * ATT_Q_BYTES + row * 128;
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
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_empty[qs]);   // this warp no longer reads the Q tile from shared memory
      }

      // A warp whose 32 query rows all lie past the end of the sequence (three of the four warps in the one-row ninth
      // query tile of a 1025-token ViT window) only keeps the barrier phases moving: it leaves the MUFU pipe to the
      // co-resident CTA. Its TMEM lanes hold garbage P / O / L rows that are never stored (MMA rows are independent).
      const bool warp_live = qt * ATT_BLOCK_Q + quad * 32 < p.q_len;
      if (!warp_live) {
        for (int j = 0; j < kv_tiles; ++j, ++t) {
          const int b = t & 1;
          mbar_wait(&s_full[b], (t >> 1) & 1);   // keeps this warp in step with the others: S_{t+2} follows p_full_t
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[b]);
        }
        mbar_wait(&pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);
        continue;
      }

      for (int j = 0; j < kv_tiles; ++j, ++t) {
        const int b = t & 1;
        const uint32_t ph = (t >> 1) & 1;
        int valid = p.kv_len - j * ATT_BLOCK_KV;
        valid = valid > ATT_BLOCK_KV ? ATT_BLOCK_KV : valid;
        const int chunks = (valid + 31) >> 5;   // warp-uniform: 1 or 2
        const uint32_t tmem_s = tmem_tile + ATT_COL_S + b * 64;

        if (quad == 2) ATT_TRACE(t, 8);
        mbar_wait(&s_full[b], ph);
        tc_fence_after();
        if (quad == 2) ATT_TRACE(t, 9);
        uint32_t s[64];
        tmem_ld32(tmem_s + 0, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        if (chunks > 1) tmem_ld32(tmem_s + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
        tmem_ld_wait();
        if (quad == 2) ATT_TRACE(t, 10);

        if (valid < ATT_BLOCK_KV) {            // tail tile only: mask keys past the sequence end
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= valid) s[i] = 0xff800000u;  // -inf
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
#pragma unroll
        for (int i = 0; i < 64; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[i]));
        const float m_tile = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])) * kLog2e;
        if (quad == 2) ATT_TRACE(t, 11);

        {
          // on a unit's first tile nothing has been accumulated yet, so moving the reference is free: take the larger of
          // the extra key's score and the tile max (with the threshold the early tiles of extra mode kept paying rescales)
          const bool jump = m_tile > m_ref + (j == 0 ? 0.f : kRescaleThreshold);
          if (__any_sync(0xffffffffu, jump)) { // rare after the first tiles: rescale O in TMEM
            const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;   // exp2(-inf) = 0 on the very first tile
            if (jump) { m_ref = m_tile; w_extra *= alpha; }
            if (j > 0) {
              // every PV up to tile t-1 must have executed before O is touched (pv_done[(t-1)%2] is the newest phase
              // of that barrier)
              mbar_wait(&pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
              tc_fence_after();
#pragma unroll 1
              for (int c = 0; c < ATT_D / 16 + 1; ++c) {   // O and, right behind it, the row sums L
                uint32_t r[16];
                tmem_ld16(tmem_o + c * 16, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                tmem_st16(tmem_o + c * 16, r);
              }
            }
          }
        }
        if (quad == 2) ATT_TRACE(t, 12);

        // P = exp2(s log2e - m_ref), truncated to bf16 by the PRMT that packs two of them; the row sum of exactly those
        // bf16 values comes from the tensor core (L += P * ones), so P / l is an exact softmax of the weights the PV MMA
        // uses, and the CUDA cores spend 2.5 instructions per element (FFMA, MUFU, half a PRMT).
        // Software-pipelined by hand: the pair (2k, 2k+1) is packed kExpDist pairs after its exponentials were issued.
        // ptxas otherwise puts each PRMT right behind its two MUFU.EX2 and the warp then waits out the full MUFU latency
        // (~40 clk) for every pair. The PRMT selector carries a data dependence on a LATER exponential (its sign bit,
        // always 0), which ptxas cannot hoist over.
        constexpr int kExpDist = 4;
        constexpr int kPolyEvery = VFM_ATT_POLY;   // 0: all exponentials on the MUFU pipe; n: one in 2n on the FMA pipe
        if (valid <= 2) {
          // tail tile of one or two keys (the 1025th token of a ViT window): two exponentials instead of 32; the masked
          // columns of the 32-key MMA step are exact zeros
          uint32_t pk[16];
          const float e0 = fast_exp2(fmaf(__uint_as_float(s[0]), kLog2e, -m_ref));
          const float e1 = fast_exp2(fmaf(__uint_as_float(s[1]), kLog2e, -m_ref));   // s[1] = -inf when masked
          pk[0] = __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632u);
#pragma unroll
          for (int i = 1; i < 16; ++i) pk[i] = 0u;
          tmem_st16(tmem_s, pk);
        } else
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (c < chunks) {
            uint32_t pk[16];
            uint32_t* sc = &s[c * 32];
#pragma unroll
            for (int i = 0; i < 16 + kExpDist; ++i) {
              if (i < 16) {
                sc[2 * i] = __float_as_uint(fast_exp2(fmaf(__uint_as_float(sc[2 * i]), kLog2e, -m_ref)));
                // every kPolyEvery-th pair computes its odd element on the FMA/ALU pipes (degree-3 polynomial, relative
                // error 1e-4, 40x below bf16 resolution): the MUFU pipe (16 ex2/clk/SM) is the busiest unit of this kernel
                const float x1 = fmaf(__uint_as_float(sc[2 * i + 1]), kLog2e, -m_ref);
                sc[2 * i + 1] = __float_as_uint(kPolyEvery != 0 && (i % (kPolyEvery ? kPolyEvery : 1)) == 0 ? poly_exp2(x1) : fast_exp2(x1));
              }
              if (i >= kExpDist) {
                const int kk = i - kExpDist;
                const uint32_t sel = i < 16 ? __umulhi(sc[2 * i], 2u) + 0x7632u : 0x7632u;   // 0x7632 + sign bit (= 0)
                pk[kk] = __byte_perm(sc[2 * kk], sc[2 * kk + 1], sel);
              }
            }
            tmem_st16(tmem_s + c * 16, pk);   // P_t overwrites the first half of S_t (last read by this thread itself)
          }
        }
        if (quad == 2) ATT_TRACE(t, 15);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[b]);   // one arrival per warp
        if (quad == 2) ATT_TRACE(t, 13);
      }
      // ---- unit epilogue: copy L and O out of TMEM, hand the accumulator back, then normalise and store
      // the last PV implies all earlier ones (commits are ordered)
      mbar_wait(&pv_done[(t - 1) & 1], ((t - 1) >> 1) & 1);
      tc_fence_after();
      const int q_idx = qt * ATT_BLOCK_Q + row;   // body index of this thread's query
      uint32_t o[64];
      float inv;
      {
        uint32_t r[16];
        tmem_ld16(tmem_tile + ATT_COL_L, r);
        tmem_ld32(tmem_o + 0, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld32(tmem_o + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_free);   // the next unit's first PV MMA may overwrite O and L now
        inv = 1.f / (__uint_as_float(r[0]) + w_extra);
      }
      const uint4* vx = reinterpret_cast<const uint4*>(p.v_ptr + static_cast<size_t>(seq) * p.kv_seq_rows * p.v_ld + p.v_col0 + head * ATT_D);
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(seq * p.q_seq_rows + p.q_row_off + q_idx) * p.out_ld + head * ATT_D);
      if (q_idx < p.q_len) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
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

#ifdef VFM_EPI_TIMING
  if (trace_slot == 0 && warp == 4 && lane == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_att_dbg[2] = clock64(); g_att_dbg[3] = ns;
  }
#endif
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Experimental alternative (mode 3 of vfm_attention_fwd_ex): the SERIAL kernel shape of attention_glob_sm100.cuh at
// head_dim 64 — one CTA per (sequence, head, 128-query tile), single score buffer with P in place, O in TMEM, row sum
// in a register, K and V single-buffered but released in different phases — slimmed to 128 TMEM columns, 33 KB of shared
// memory and <= 84 registers (two passes over the score columns, 32 at a time) so that FOUR CTAs share an SM: four
// independent serial chains instead of two pipelined ones.
constexpr int ATS_THREADS = 192;
constexpr int ATS_SMEM_BYTES = ATT_Q_BYTES + 2 * ATT_KV_BYTES + 1024 + 128;
constexpr uint32_t ATS_TMEM_COLS = 128;

__global__ void __launch_bounds__(ATS_THREADS, 4)
attention_serial_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic keeps the shared address space (an integer round trip makes every access generic)
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATT_Q_BYTES;
  uint8_t* sV = sK + ATT_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ATT_KV_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;   // S_t executed
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;   // PV_t executed
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int qt = unit % p.q_tiles;
  const int head = (unit / p.q_tiles) % p.heads;
  const int seq = unit / (p.q_tiles * p.heads);
  const int q_row0 = seq * p.q_seq_rows + p.q_row_off + qt * ATT_BLOCK_Q;
  const int kv_row0 = seq * p.kv_seq_rows + p.kv_row_off;
  const int n_tiles = (p.kv_len + ATT_BLOCK_KV - 1) / ATT_BLOCK_KV;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], i == 6 ? 4 : 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(q_full, ATT_Q_BYTES);
    tma_load_2d(sQ, &tmap_q, q_full, p.q_col0 + head * ATT_D, q_row0);
    tma_load_2d(sQ + ATT_KV_BYTES, &tmap_q, q_full, p.q_col0 + head * ATT_D, q_row0 + 64);
    mbar_arrive_expect_tx(k_full, ATT_KV_BYTES);
    tma_load_2d(sK, &tmap_k, k_full, p.k_col0 + head * ATT_D, kv_row0);
    mbar_arrive_expect_tx(v_full, ATT_KV_BYTES);
    tma_load_2d(sV, &tmap_v, v_full, p.v_col0 + head * ATT_D, kv_row0);
  }
  if (warp == 1) tmem_alloc<ATS_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    for (int t = 1; t < n_tiles; ++t) {
      mbar_wait(k_empty, (t - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(k_full, ATT_KV_BYTES);
        tma_load_2d(sK, &tmap_k, k_full, p.k_col0 + head * ATT_D, kv_row0 + t * ATT_BLOCK_KV);
      }
      __syncwarp();
      mbar_wait(v_empty, (t - 1) & 1);
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(v_full, ATT_KV_BYTES);
        tma_load_2d(sV, &tmap_v, v_full, p.v_col0 + head * ATT_D, kv_row0 + t * ATT_BLOCK_KV);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BLOCK_Q, ATT_BLOCK_KV, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(ATT_BLOCK_Q, ATT_D, 0, 1);
    const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 64;
    const uint64_t dq = make_sw128_desc(smem_u32(sQ)), dk = make_sw128_desc(smem_u32(sK)), dv = make_sw128_desc(smem_u32(sV));
    mbar_wait(q_full, 0);
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(k_full, t & 1);
      tc_fence_after();
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
        tc_commit(k_empty);
        tc_commit(s_full);
      }
      __syncwarp();
      mbar_wait(v_full, t & 1);
      mbar_wait(p_full, t & 1);
      tc_fence_after();
      if (elect_one_sync()) {
        const int ksteps = (min(ATT_BLOCK_KV, p.kv_len - t * ATT_BLOCK_KV) + 15) >> 4;
        for (int k = 0; k < ksteps; ++k) umma_ts(tmem_o, tmem_s + 8 * k, dv + 128 * k, idesc_pv, (t | k) != 0);
        tc_commit(v_empty);
        if (t == n_tiles - 1) tc_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tmem_s = tmem_base + lane_base, tmem_o = tmem_base + lane_base + 64;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr float kRescaleThreshold = 8.0f;
    float m_ref = -INFINITY, l = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      mbar_wait(s_full, t & 1);
      tc_fence_after();
      const int valid = p.kv_len - t * ATT_BLOCK_KV;
      // pass 1: row max, 32 columns at a time
      float m = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + 32 * c, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (32 * c + i < valid) m = fmaxf(m, __uint_as_float(r[i]));
      }
      const float m_tile = m * kLog2e;
      const bool jump = m_tile > m_ref + (t == 0 ? 0.f : kRescaleThreshold);
      if (__any_sync(0xffffffffu, jump)) {
        const float alpha = jump ? fast_exp2(m_ref - m_tile) : 1.f;
        if (jump) { m_ref = m_tile; l *= alpha; }
        if (t > 0) {
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
      // pass 2: P over the scores, in place (a 32-column chunk is in registers before 16 packed columns are written)
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_s + 32 * c, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int key = 32 * c + 2 * i;
          const float e0 = key < valid ? fast_exp2(fmaf(__uint_as_float(r[2 * i]), kLog2e, -m_ref)) : 0.f;
          const float e1 = key + 1 < valid ? fast_exp2(fmaf(__uint_as_float(r[2 * i + 1]), kLog2e, -m_ref)) : 0.f;
          const __nv_bfloat162 b = __floats2bfloat162_rn(e0, e1);
          l += __low2float(b) + __high2float(b);
          pk[i] = *reinterpret_cast<const uint32_t*>(&b);
        }
        tmem_st16(tmem_s + 16 * c, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv = 1.f / l;
    const int q_idx = qt * ATT_BLOCK_Q + row;
    uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(seq * p.q_seq_rows + p.q_row_off + q_idx) * p.out_ld + head * ATT_D);
#pragma unroll 1
    for (int c = 0; c < ATT_D / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(tmem_o + 16 * c, r);
      tmem_ld_wait();
      if (q_idx < p.q_len) {
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
    tmem_dealloc<ATS_TMEM_COLS>(tmem_base);
  }
}

}  // namespace vfm
