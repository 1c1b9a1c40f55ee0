// SAM ViT backbone pieces (BASELINE config 5; rein/models/backbones/sam_vit.py): attention with the decomposed
// relative-position bias for head_dim 80 and windows of 196 tokens, the rel-pos terms it consumes, and the row gather
// behind window_partition / window_unpartition.
//
// The attention kernel here is a FIRST version on warp-level tensor-core instructions (mma.sync m16n8k16 bf16, FA2
// layout: 16 query rows per warp, online softmax in registers): the tcgen05 kernel in attention_sm100.cuh is built
// around head_dim 64 (one SWIZZLE_128B atom per operand row, 64 O columns in TMEM) and has no additive-bias path; a
// tcgen05 port with K-augmented operands ([q | q.Rh | q.Rw] x [k | onehot(kh) | onehot(kw)]) is the planned successor.
#pragma once
#include <cuda_bf16.h>

#include "sm100_ptx.cuh"

namespace vfm {

// ---------------------------------------------------------------------------------------------------------------
// rel[seq][head][token][0..k_h) = q . Rh[qh, kh, :],  rel[..][k_h + kw] = q . Rw[qw, kw, :]   (UNSCALED q)
// add_decomposed_rel_pos, sam_vit.py:417-421 (the two einsums). Rh [q_h][k_h][D], Rw [q_w][k_w][D] fp32 are the
// tables get_rel_pos (:358-388) gathers (interpolated on the host at load time).
// One CTA per (query row qh | query column qw, sequence): its [k][D] table slice sits in shared memory while all
// (token, head) vectors of that row / column stream through; one warp per vector, lanes over k.
template <int D>
__global__ void __launch_bounds__(256)
relpos_terms_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, const float* __restrict__ Rh, const float* __restrict__ Rw,
                    float* __restrict__ rel, int seq_len, int heads, int q_h, int q_w, int k_h, int k_w) {
  extern __shared__ float sm_rel[];
  const int a = blockIdx.x, seq = blockIdx.y;
  const bool is_h = a < q_h;
  const int fixed = is_h ? a : a - q_h;            // qh or qw
  const int kn = is_h ? k_h : k_w;                 // outputs per vector
  const int other = is_h ? q_w : q_h;              // tokens along the other axis
  const float* R = (is_h ? Rh : Rw) + static_cast<size_t>(fixed) * kn * D;
  float* Rs = sm_rel;                              // [kn][D + 1]
  float* qs = sm_rel + kn * (D + 1);               // [8 warps][D]
  for (int i = threadIdx.x; i < kn * D; i += blockDim.x) Rs[(i / D) * (D + 1) + (i % D)] = R[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* q = qs + warp * D;
  const int kk = k_h + k_w;
  for (int v = warp; v < other * heads; v += 8) {
    const int o = v / heads, head = v - o * heads;
    const int token = is_h ? fixed * q_w + o : o * q_w + fixed;
    const __nv_bfloat16* qp = qkv + (static_cast<size_t>(seq) * seq_len + token) * ld + head * D;
    __syncwarp();
    for (int d = lane; d < D; d += 32) q[d] = __bfloat162float(qp[d]);
    __syncwarp();
    float* out = rel + ((static_cast<size_t>(seq) * heads + head) * seq_len + token) * kk + (is_h ? 0 : k_h);
    for (int k = lane; k < kn; k += 32) {
      const float* r = Rs + k * (D + 1);
      float acc = 0.f;
#pragma unroll 8
      for (int d = 0; d < D; ++d) acc = fmaf(q[d], r[d], acc);
      out[k] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// dst[i, :] = map[i] >= 0 ? src[map[i], :] : 0   (bf16 rows of C elements, C % 8 == 0). window_partition (zero padding
// to a multiple of the window, sam_vit.py:306-316) and window_unpartition (:335-346) are both row gathers.
__global__ void rows_gather_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                   const int* __restrict__ map, long long n_rows, int C) {
  const int chunks = C >> 3;
  const long long total = n_rows * chunks;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / chunks;
    const int c = static_cast<int>(i - r * chunks);
    const int m = __ldg(map + r);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (m >= 0) v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(m) * C) + c);
    reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * C)[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// softmax(scale * q k^T + rel_h[q, kh(k)] + rel_w[q, kw(k)]) v  per (sequence, head); Attention.forward sam_vit.py:263-289.
// qkv bf16 [n_seq * seq_len, ld] with q | k | v in the first 3 * heads * D columns as the qkv Linear emits them (:266-270);
// key index k = kh * k_w + kw. out bf16 [n_seq * seq_len, heads * D]. The bias comes from one of
//   rel   fp32 [n_seq][heads][seq_len][k_h + k_w] (relpos_terms_kernel), or
//   table terms in the qkv rows themselves (g_col0 >= 0): G_h[token][head][r] = q . T_h[r], r in [0, 2 k_h - 1), at column
//         g_col0 + head * (2 k_h - 1) + r and G_w behind all heads' G_h; rel_h[q, kh] = G_h[qh - kh + k_h - 1] (the gather of
//         get_rel_pos, :382-388). G is linear in the block input, so the engine gets it from the qkv GEMM as extra
//         output columns (weights T . W_q folded on the host) instead of a separate pass over q.
// CTA = 64 query rows of one (sequence, head); 4 warps x 16 rows; 64-key tiles double-buffered with cp.async.
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_ptr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(smem_ptr)));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Tile configuration: WARPS x 16 query rows per CTA, KV keys per tile. Windows of 196 tokens use <7, 112>: two query
// tiles (112 + 84 rows: 13 of 14 warps carry rows) and two key tiles (112 + 84 keys: 26 eight-key blocks for 196 keys),
// against four 64-row tiles x four 64-key tiles (a quarter of them padding) with the generic <4, 64> shape.
template <int D, int WARPS, int KV>
struct RelposCfg {
  static constexpr int kBQ = 16 * WARPS;
  static constexpr int kThreads = 32 * WARPS;
  static constexpr int kPitch = D + 8;                 // bf16 elements; (D + 8) * 2 B keeps ldmatrix rows conflict-free
  static constexpr int kQBytes = kBQ * kPitch * 2;
  static constexpr int kKVBytes = KV * kPitch * 2;     // one K or V tile
  static size_t bytes(int kk) {
    return static_cast<size_t>(kQBytes) + 4 * kKVBytes + static_cast<size_t>(kBQ) * kk * 2 + 2 * (KV / 2) * 4;
  }
};

template <int D, int WARPS, int KV>
__global__ void __launch_bounds__(32 * WARPS, 2)
attention_relpos_kernel(const __nv_bfloat16* __restrict__ qkv, int ld, int g_col0, const float* __restrict__ rel,
                        __nv_bfloat16* __restrict__ out, int seq_len, int heads, int k_h, int k_w, float scale) {
  using Cfg = RelposCfg<D, WARPS, KV>;
  constexpr int P = Cfg::kPitch, BQ = Cfg::kBQ, NT = Cfg::kThreads;
  constexpr int KS = D / 16;                            // k-steps of Q K^T
  constexpr int NO = D / 8;                             // n-blocks of O
  constexpr int NB = KV / 8;                            // 8-key blocks per tile
  constexpr int GRP = KV / 16;                          // 16-key groups per tile
  static_assert(KV % 16 == 0 && D % 16 == 0, "tile shape");
  extern __shared__ __align__(16) uint8_t sm_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(sm_raw);
  __nv_bfloat16* Ks = Qs + BQ * P;                      // [2][KV][P]
  __nv_bfloat16* Vs = Ks + 2 * KV * P;                  // [2][KV][P]
  const bool has_bias = rel != nullptr || g_col0 >= 0;
  const int kk = has_bias ? k_h + k_w : 0;                           // even: k_h, k_w are checked even by the launcher
  // bias terms of this CTA's query rows as bf16 (the table terms ARE bf16; 64 KB of fp32 for a 64 x 64 key grid would
  // cost the second CTA per SM)
  __nv_bfloat16* rel_s = Vs + 2 * KV * P;                            // [BQ][kk]
  int* col_hw = reinterpret_cast<int*>(rel_s + BQ * kk);             // [2][KV/2] per PAIR of keys: kh | (k_h + kw) << 16

  const int qt = blockIdx.x, head = blockIdx.y, seq = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = heads * D;
  const size_t row0 = static_cast<size_t>(seq) * seq_len;
  const int q0 = qt * BQ;
  const int kv_tiles = (seq_len + KV - 1) / KV;
  constexpr int CH = D / 8;                             // 16-byte chunks per row

  auto load_tile = [&](__nv_bfloat16* dst, int which, int r0, int rows) {   // rows of Q / K / V (which = 0 / 1 / 2)
    for (int i = tid; i < rows * CH; i += NT) {
      const int r = i / CH, c = i - r * CH;
      __nv_bfloat16* d = dst + r * P + c * 8;
      if (r0 + r < seq_len) cp_async16(d, qkv + (row0 + r0 + r) * ld + which * C + head * D + c * 8);
      else *reinterpret_cast<uint4*>(d) = make_uint4(0u, 0u, 0u, 0u);   // rows past the end: zeros (0 * garbage = NaN)
    }
  };
  // k_w is even and tiles start at even keys, so the keys (2i, 2i + 1) of a pair share kh and have adjacent kw
  auto load_cols = [&](int buf, int j) {
    if (tid < KV / 2) {
      const int c = j * KV + 2 * tid;
      int h = c / k_w;
      const int w = c - h * k_w;
      h = h < k_h ? h : 0;                     // keys past the end are masked anyway
      col_hw[buf * (KV / 2) + tid] = h | ((k_h + w) << 16);
    }
  };

  load_tile(Qs, 0, q0, BQ);
  load_tile(Ks, 1, 0, KV);
  load_tile(Vs, 2, 0, KV);
  cp_async_commit();
  if (rel) {
    const float* rp = rel + ((static_cast<size_t>(seq) * heads + head) * seq_len + q0) * kk;
    for (int i = tid; i < BQ * kk; i += NT) rel_s[i] = __float2bfloat16((q0 + i / kk) < seq_len ? rp[i] : 0.f);
    load_cols(0, 0);
  } else if (g_col0 >= 0) {
    // row r of the tile: G_h is read backwards (qh - kh + k_h - 1 for kh = 0..k_h-1), G_w likewise — two contiguous runs.
    // Four rows per warp step with all loads issued before the first store (one memory round trip per step, not four).
    const int Lh = 2 * k_h - 1, Lw = 2 * k_w - 1;
    constexpr int RB = 4;
    for (int r = warp; r < BQ; r += WARPS * RB) {
      const __nv_bfloat16* gh[RB];
      const __nv_bfloat16* gw[RB];
      bool live[RB];
#pragma unroll
      for (int u = 0; u < RB; ++u) {
        const int q = q0 + r + u * WARPS;
        live[u] = r + u * WARPS < BQ && q < seq_len;
        const int qq = live[u] ? q : 0;
        const int qh = qq / k_w, qw = qq - qh * k_w;
        const __nv_bfloat16* gp = qkv + (row0 + qq) * ld + g_col0;
        gh[u] = gp + head * Lh + qh + k_h - 1;
        gw[u] = gp + heads * Lh + head * Lw + qw + k_w - 1;
      }
      for (int c = lane; c < kk; c += 32) {
        __nv_bfloat16 v[RB];
#pragma unroll
        for (int u = 0; u < RB; ++u) v[u] = c < k_h ? gh[u][-c] : gw[u][-(c - k_h)];
#pragma unroll
        for (int u = 0; u < RB; ++u)
          if (r + u * WARPS < BQ) rel_s[(r + u * WARPS) * kk + c] = live[u] ? v[u] : __float2bfloat16(0.f);
      }
    }
    load_cols(0, 0);
  }

  uint32_t qf[KS][4];
  float o[NO][4];
#pragma unroll
  for (int n = 0; n < NO; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // rows g and g + 8 of this warp's 16
  constexpr float kLog2e = 1.4426950408889634f;
  const int g = lane >> 2, t4 = lane & 3;
  const float sc = scale * kLog2e;
  const bool warp_live = q0 + warp * 16 < seq_len;   // a warp whose 16 rows lie past the end only helps with the loads

  for (int j = 0; j < kv_tiles; ++j) {
    const int buf = j & 1;
    if (j + 1 < kv_tiles) {   // prefetch the next K / V tile into the other buffer (freed by the barrier at the loop end)
      load_tile(Ks + (buf ^ 1) * KV * P, 1, (j + 1) * KV, KV);
      load_tile(Vs + (buf ^ 1) * KV * P, 2, (j + 1) * KV, KV);
      if (has_bias) load_cols(buf ^ 1, j + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (warp_live) {
      // wide key tiles (14 score blocks in registers) re-read the Q fragments from shared memory every tile instead of
      // holding 20 registers across the loop: 158 -> <= 146 registers is the difference between one and two CTAs per SM
      if (j == 0 || KV > 64) {
#pragma unroll
        for (int k = 0; k < KS; ++k)
          ldmatrix_x4(qf[k], Qs + (warp * 16 + (lane & 15)) * P + k * 16 + (lane >> 4) * 8);
      }
      const __nv_bfloat16* Kt = Ks + buf * KV * P;
      const __nv_bfloat16* Vt = Vs + buf * KV * P;
      // ragged last tile: only the 16-key groups that hold real keys are multiplied (CTA-uniform bound)
      const int keys_here = min(KV, seq_len - j * KV);
      const int groups = (keys_here + 15) >> 4;
      // ---- S = Q K^T for 16 rows x KV keys
      float s[NB][4];
#pragma unroll
      for (int n = 0; n < NB; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
#pragma unroll
      for (int n2 = 0; n2 < GRP; ++n2) {   // two 8-key blocks per ldmatrix.x4
        if (n2 < groups) {
#pragma unroll
          for (int k = 0; k < KS; ++k) {
            uint32_t b[4];
            ldmatrix_x4(b, Kt + (n2 * 16 + (lane & 7) + (lane >> 4) * 8) * P + k * 16 + ((lane >> 3) & 1) * 8);
            mma_bf16_16816(s[2 * n2], qf[k], b[0], b[1]);
            mma_bf16_16816(s[2 * n2 + 1], qf[k], b[2], b[3]);
          }
        }
      }
      // ---- scale, bias, mask (everything in log2 units)
      const int r_lo = warp * 16 + g, r_hi = r_lo + 8;
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const int col = n * 8 + 2 * t4;        // this thread's key pair (col, col + 1) of block n
        float2 w_lo = make_float2(0.f, 0.f), w_hi = w_lo;
        float h_lo = 0.f, h_hi = 0.f;
        if (has_bias) {
          const int hw = col_hw[buf * (KV / 2) + (col >> 1)];
          const int ch = hw & 0xffff, cw = hw >> 16;
          h_lo = __bfloat162float(rel_s[r_lo * kk + ch]); h_hi = __bfloat162float(rel_s[r_hi * kk + ch]);
          w_lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rel_s + r_lo * kk + cw));
          w_hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(rel_s + r_hi * kk + cw));
        }
        const int c_abs = j * KV + col;
        s[n][0] = c_abs < seq_len ? fmaf(s[n][0], sc, (h_lo + w_lo.x) * kLog2e) : -INFINITY;
        s[n][1] = c_abs + 1 < seq_len ? fmaf(s[n][1], sc, (h_lo + w_lo.y) * kLog2e) : -INFINITY;
        s[n][2] = c_abs < seq_len ? fmaf(s[n][2], sc, (h_hi + w_hi.x) * kLog2e) : -INFINITY;
        s[n][3] = c_abs + 1 < seq_len ? fmaf(s[n][3], sc, (h_hi + w_hi.y) * kLog2e) : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[n][0], s[n][1]));
        mx1 = fmaxf(mx1, fmaxf(s[n][2], s[n][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: every tile has at least one valid key
      const float a0 = fast_exp2(m0 - mn0), a1 = fast_exp2(m1 - mn1);
      m0 = mn0; m1 = mn1;
      l0 *= a0; l1 *= a1;
#pragma unroll
      for (int n = 0; n < NO; ++n) { o[n][0] *= a0; o[n][1] *= a0; o[n][2] *= a1; o[n][3] *= a1; }
      // ---- P = exp2(s - m) as bf16 A fragments; the row sums add up the rounded values the MMA consumes
      uint32_t pa[GRP][4];
#pragma unroll
      for (int n = 0; n < NB; ++n) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(fast_exp2(s[n][0] - m0), fast_exp2(s[n][1] - m0));
        const __nv_bfloat162 hi = __floats2bfloat162_rn(fast_exp2(s[n][2] - m1), fast_exp2(s[n][3] - m1));
        l0 += __low2float(lo) + __high2float(lo);
        l1 += __low2float(hi) + __high2float(hi);
        pa[n >> 1][(n & 1) * 2 + 0] = *reinterpret_cast<const uint32_t*>(&lo);
        pa[n >> 1][(n & 1) * 2 + 1] = *reinterpret_cast<const uint32_t*>(&hi);
      }
      // ---- O += P V
#pragma unroll
      for (int k = 0; k < GRP; ++k) {          // 16 keys per step
        if (k < groups) {
#pragma unroll
          for (int n2 = 0; n2 < NO / 2; ++n2) {  // two 8-wide blocks of head dims per ldmatrix.x4.trans
            uint32_t b[4];
            ldmatrix_x4_trans(b, Vt + (k * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * P + n2 * 16 + (lane >> 4) * 8);
            mma_bf16_16816(o[2 * n2], pa[k], b[0], b[1]);
            mma_bf16_16816(o[2 * n2 + 1], pa[k], b[2], b[3]);
          }
        }
      }
    }
    __syncthreads();   // everyone is done with this buffer before the next prefetch overwrites it
  }
  if (!warp_live) return;
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int q_lo = q0 + warp * 16 + g, q_hi = q_lo + 8;
#pragma unroll
  for (int n = 0; n < NO; ++n) {
    const int col = head * D + n * 8 + 2 * t4;
    if (q_lo < seq_len)
      *reinterpret_cast<uint32_t*>(out + (row0 + q_lo) * C + col) = pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
    if (q_hi < seq_len)
      *reinterpret_cast<uint32_t*>(out + (row0 + q_hi) * C + col) = pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
  }
}

}  // namespace vfm
