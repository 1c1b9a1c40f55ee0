// C ABI of libvfmseg_b200.so (see include/vfmseg_b200.h). Host-side only: argument checks,
// TMA tensor-map encoding, kernel launches and the two fused drivers (ViT backbone, LinearHead).
#include "../../include/vfmseg_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "attention_sm100.cuh"
#include "attention_pp_sm100.cuh"
#include "attention_ph_sm100.cuh"
#include "sam_ops.cuh"
#include "attention_win_sm100.cuh"
#include "attention_glob_sm100.cuh"
#include "elementwise.cuh"
#include "eva_ops.cuh"
#include "gemm_sm100.cuh"
#include "ms_refine.cuh"
#include "slide_tail.cuh"

using namespace vfm;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define VFM_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e_ = (expr);                                                                    \
    if (e_ != cudaSuccess) return fail(VFM_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

// ---- optional per-launch timing (bench.py's roofline leg): CUDA events on the launch stream ----
struct ProfRec { const char* name; cudaEvent_t start, stop; };
std::mutex g_prof_mu;
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;

// Brackets one kernel launch: counts it and, when profiling is on, records start/stop events.
struct LaunchScope {
  const char* name; cudaStream_t st; cudaEvent_t start = nullptr, stop = nullptr;
  LaunchScope(const char* n, cudaStream_t s) : name(n), st(s) {
    if (!g_prof_on) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_pool.empty()) { start = g_prof_pool.back().first; stop = g_prof_pool.back().second; g_prof_pool.pop_back(); }
    else if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) { start = stop = nullptr; return; }
    cudaEventRecord(start, st);
  }
  ~LaunchScope() {
    if (!start) return;
    cudaEventRecord(stop, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_recs.push_back({name, start, stop});
  }
};

#define VFM_LAUNCH_CHECK(name)                                                                  \
  do {                                                                                          \
    cudaError_t e_ = cudaGetLastError();                                                        \
    if (e_ != cudaSuccess) return fail(VFM_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e_)); \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                         \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int make_tmap_ex(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, CUtensorMapDataType dt, uint32_t esize);

// bf16 row-major [rows, cols] with leading dimension ld (elements); box = {64 cols, box_rows}; 128-B swizzle.
int make_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  return make_tmap_ex(m, ptr, rows, cols, ld, box_rows, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2);
}

// bf16 row-major [rows, cols], box = {box_cols, box_rows} with box_cols * 2 B == the swizzle span (32 / 64 / 128 B)
int make_tmap_sw(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) % 16) != 0)
    return fail(VFM_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte row pitch");
  const CUtensorMapSwizzle sw = box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return VFM_OK;
}

// row-major [rows, cols] tensor of `esize`-byte elements; box rows x box_cols with box_cols * esize == 128 B.
int make_tmap_ex(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                 uint32_t box_cols, CUtensorMapDataType dt, uint32_t esize) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * esize) % 16) != 0)
    return fail(VFM_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte row pitch (ptr=%p ld=%llu)", ptr,
                static_cast<unsigned long long>(ld));
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
  return VFM_OK;
}

// bf16 ConvT output [n_rows_out (= crops*2h), 2w, c_out] as the 4-D tensor (co, dx, x, R); box {64, 1, 32, 1}, 128-B swizzle.
int make_tmap_convt(CUtensorMap* m, const void* ptr, uint64_t R_total, uint64_t w, uint64_t c_out) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(VFM_ERR_INVALID, "TMA operand must be 16-byte aligned");
  cuuint64_t dims[4] = {c_out, 2, w, R_total};
  cuuint64_t strides[3] = {c_out * 2, 2 * c_out * 2, 2 * w * c_out * 2};   // bytes, dims 1..3
  cuuint32_t box[4] = {64, 1, 32, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VFM_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed (%d)", static_cast<int>(r));
  return VFM_OK;
}

int sm_count() {
  static int n = [] {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return v;
  }();
  return n;
}

// VFM_PDL=0 switches programmatic dependent launch off (every kernel then starts after its predecessor has drained)
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("VFM_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

struct OutDesc { const void* ptr = nullptr; int ld = 0; int convt_w = 0; int convt_rows = 0; int rows = 0; };   // destination of the TMA-store epilogues

template <int BLOCK_N, int CTA_GROUP, class Epi>
int launch_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const Epi& epi, cudaStream_t st,
                const char* name, OutDesc od = OutDesc()) {
  if (!A || !W) return fail(VFM_ERR_INVALID, "%s: null operand", name);
  if (M <= 0 || N <= 0 || K <= 0 || (N % 32) != 0 || (K % 8) != 0)
    return fail(VFM_ERR_INVALID, "%s: need M,N,K > 0, N %% 32 == 0, K %% 8 == 0 (M=%d N=%d K=%d)", name, M, N, K);
  using Cfg = GemmCfg<BLOCK_N, CTA_GROUP, epi_warp_bytes<Epi>::value>;
  CUtensorMap ta, tb;
  int rc = make_tmap(&ta, A, M, K, lda, GEMM_BLOCK_M);
  if (rc) return rc;
  rc = make_tmap(&tb, W, N, K, ldw, Cfg::kBRows);
  if (rc) return rc;
  CUtensorMap tout = ta;   // placeholder for epilogues that store with plain instructions
  if constexpr (Epi::kMode == EPI_TMA_BF16) {
    if constexpr (Epi::kStore4D) {
      if ((rc = make_tmap_convt(&tout, od.ptr, od.convt_rows, od.convt_w, od.ld))) return rc;
    } else {
      if ((rc = make_tmap_ex(&tout, od.ptr, M, N, od.ld, 32, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2))) return rc;
    }
  } else if constexpr (Epi::kMode == EPI_TMA_RED_F32 || Epi::kMode == EPI_TMA_RES_STATS) {
    if ((rc = make_tmap_ex(&tout, od.ptr, od.rows > 0 ? od.rows : M, N, od.ld, 32, 32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4))) return rc;
  }
  if constexpr (Epi::kMode == EPI_TMA_RES_STATS) {
    if (N % BLOCK_N) return fail(VFM_ERR_INVALID, "%s: N must be a multiple of %d", name, BLOCK_N);
  }
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, CTA_GROUP, Epi>;
  static bool attr_done = false;  // per template instantiation
  if (!attr_done) {
    VFM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_done = true;
  }
  constexpr int TILE_M = GEMM_BLOCK_M * CTA_GROUP;
  const int m_tiles = (M + TILE_M - 1) / TILE_M, n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = m_tiles * n_tiles;
  int sms = sm_count();
  if (sms <= 0) return fail(VFM_ERR_CUDA, "%s: no CUDA device", name);
  const int max_groups = sms / CTA_GROUP;
  const int groups = tiles < max_groups ? tiles : max_groups;
  // K tail (K % 64 != 0) is covered by TMA zero fill of both operands.
  const int K_pad = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K * GEMM_BLOCK_K;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(groups * CTA_GROUP);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CTA_GROUP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see pdl_wait() in sm100_ptx.cuh
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  {
    LaunchScope scope(name, st);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, tout, M, N, K_pad, epi);
    if (e != cudaSuccess) return fail(VFM_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
  }
  VFM_LAUNCH_CHECK(name);
  return VFM_OK;
}

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline __nv_bfloat16* BF(void* p) { return reinterpret_cast<__nv_bfloat16*>(p); }
inline const __nv_bfloat16* BF(const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

extern "C" {

const char* vfm_last_error(void) { return g_err; }
int vfm_abi_version(void) { return VFM_ABI_VERSION; }
long long vfm_launch_count(void) { return g_launches.load(); }

int vfm_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on = on != 0;
  return VFM_OK;
}

// Synchronises the device, folds all recorded launches into "name,count,total_ms" lines, clears the records.
int vfm_prof_report(char* buf, size_t buf_bytes) {
  if (!buf || buf_bytes == 0) return fail(VFM_ERR_INVALID, "prof_report: null buffer");
  VFM_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_prof_mu);
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : g_prof_recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      auto& a = agg[r.name];
      a.first += 1;
      a.second += ms;
    }
    g_prof_pool.emplace_back(r.start, r.stop);
  }
  g_prof_recs.clear();
  std::string out;
  char line[160];
  for (auto& kv : agg) {
    snprintf(line, sizeof(line), "%s,%lld,%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
    out += line;
  }
  if (out.size() + 1 > buf_bytes) return fail(VFM_ERR_INVALID, "prof_report: buffer too small (%zu needed)", out.size() + 1);
  memcpy(buf, out.c_str(), out.size() + 1);
  return VFM_OK;
}

#ifdef VFM_APP_TRACE
// debug build only: timeline of CTA 0 of the last attention_pp_kernel launch: 4 x 64 x 8 clock64 values + 4 clock / ns values
extern "C" int vfm_debug_app_trace(long long* out2052) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out2052, vfm::g_app_trace, 4 * 64 * 8 * sizeof(long long));
  cudaMemcpyFromSymbol(out2052 + 2048, vfm::g_app_clk, 4 * sizeof(long long));
  return 0;
}
#endif
#ifdef VFM_EPI_TIMING
extern "C" int vfm_debug_att_trace(long long* out640) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out640, vfm::g_att_trace, 2 * 20 * 16 * sizeof(long long));
  cudaMemcpyFromSymbol(out640 + 640, vfm::g_att_dbg, 4 * sizeof(long long));   // CTA life: clk0, ns0, clk1, ns1
  return 0;
}
// debug build only: read-and-clear the epilogue phase cycle counters (warp 4 of every CTA)
extern "C" int vfm_debug_epi(unsigned long long* out5) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out5, vfm::g_epi_dbg, 5 * sizeof(unsigned long long));
  unsigned long long z[8] = {0};
  cudaMemcpyToSymbol(vfm::g_epi_dbg, z, sizeof(z));
  return 0;
}
#endif

int vfm_device_check(void) {
  int dev = 0, major = 0;
  VFM_CUDA(cudaGetDevice(&dev));
  VFM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return fail(VFM_ERR_DEVICE, "device compute capability %d.x, need 10.x (sm_100a)", major);
  return VFM_OK;
}

int vfm_gemm_bias_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M,
                       int N, int K, void* stream) {
  if (!out || (ldo % 8)) return fail(VFM_ERR_INVALID, "gemm_bias_bf16: bad out/ldo");
  EpiTmaBf16<false> e{bias};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_bf16", OutDesc{out, ldo});
}

int vfm_gemm_bias_rope_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo, int M,
                            int N, int K, const float* cos_t, const float* sin_t, int rope_cols, int tokens_per_seq, void* stream) {
  if (!out || (ldo % 8) || !cos_t || !sin_t || tokens_per_seq < 2 || rope_cols < 0 || (rope_cols % 64) || rope_cols > N ||
      ((reinterpret_cast<uintptr_t>(cos_t) | reinterpret_cast<uintptr_t>(sin_t)) & 15))
    return fail(VFM_ERR_INVALID, "gemm_bias_rope_bf16: bad args (rope_cols %% 64 == 0, 16-byte aligned tables)");
  EpiTmaBf16Rope e{bias, cos_t, sin_t, rope_cols, FastDiv(tokens_per_seq)};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_rope_bf16", OutDesc{out, ldo});
}

int vfm_gemm_bias_gelu_bf16(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int ldo,
                            int M, int N, int K, void* stream) {
  if (!out || !bias || (ldo % 8)) return fail(VFM_ERR_INVALID, "gemm_bias_gelu_bf16: bad out/bias/ldo");
  EpiTmaBf16<true> e{bias};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_gelu_bf16", OutDesc{out, ldo});
}

int vfm_gemm_bias_ls_residual(const void* A, int lda, const void* W, int ldw, const float* bias, const float* gamma,
                              float* x, int ldx, void* tap, int tap_ld, int tap_col0, int tokens_per_crop, int M,
                              int N, int K, void* stream) {
  if (!x || !bias || !gamma || (ldx % 4)) return fail(VFM_ERR_INVALID, "gemm_bias_ls_residual: bad x/bias/gamma/ldx");
  if (tap && ((tap_ld % 8) || (tap_col0 % 8) || tokens_per_crop <= 0))
    return fail(VFM_ERR_INVALID, "gemm_bias_ls_residual: bad tap layout");
  if (!tap) {   // no feature tap: the residual update is a TMA fp32 reduce-add, x is never read by the SM
    EpiTmaResidual e{bias, gamma};
    return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_ls_residual", OutDesc{x, ldx});
  }
  EpiResidual e{x, ldx, bias, gamma, BF(tap), tap_ld, tap_col0, FastDiv(tokens_per_crop > 0 ? tokens_per_crop : 1)};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_ls_residual_tap");
}

int vfm_gemm_bias_ls_residual_stats(const void* A, int lda, const void* W, int ldw, const float* bias, const float* gamma,
                                    float* x, int ldx, void* xb, int ldxb, float* stats, int M, int N, int K, void* stream) {
  if (!x || !bias || !gamma || !xb || !stats || (ldx % 4) || (ldxb % 8) || (N % 256) ||
      (reinterpret_cast<uintptr_t>(stats) & 15))
    return fail(VFM_ERR_INVALID, "gemm_bias_ls_residual_stats: bad x/xb/stats/bias/gamma, or N %% 256 != 0");
  // staging shape (gemm_sm100.cuh, EpiTmaResidualStats): serial single box when the main loop is long enough to hide it
  // (K >= 2048: mlp.fc2), deep otherwise; VFM_RS_MODE=1 / 2 forces deep / serial (experiments)
  static const int rs_mode = [] { const char* e = getenv("VFM_RS_MODE"); return e ? atoi(e) : 0; }();
  const bool deep = rs_mode == 1 || (rs_mode == 0 && K < 2048);
  static const int l2pf = [] { const char* e = getenv("VFM_RS_L2PF"); return e ? atoi(e) : 0; }();   // L2 prefetch of the next tile's old x: measured slower (x read from DRAM twice: +165 MB per fc2 launch)
  int rc;
  if (deep) {
    EpiTmaResidualStats<true> e{bias, gamma, reinterpret_cast<float2*>(stats), N / 128, l2pf, {}};
    if ((rc = make_tmap_ex(&e.tmap_xb, xb, M, N, ldxb, 32, 64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2))) return rc;
    return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_ls_residual_stats", OutDesc{x, ldx});
  }
  EpiTmaResidualStats<false> e{bias, gamma, reinterpret_cast<float2*>(stats), N / 128, l2pf, {}};
  if ((rc = make_tmap_sw(&e.tmap_xb, xb, M, N, ldxb, 32, 32))) return rc;
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_bias_ls_residual_stats_k", OutDesc{x, ldx});
}

int vfm_gemm_lnfold_bf16(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                         const float* stats, float eps, int gelu, void* out, int ldo, int M, int N, int K, void* stream) {
  if (!out || !bias_f || !colsum || !stats || (ldo % 8) || (K % 256) || (reinterpret_cast<uintptr_t>(stats) & 15))
    return fail(VFM_ERR_INVALID, "gemm_lnfold_bf16: bad out/bias/colsum/stats/ldo, or K %% 256 != 0");
  const float2* st = reinterpret_cast<const float2*>(stats);
  if (gelu) {
    EpiTmaBf16LN<true> e{bias_f, colsum, st, K / 128, M, 1.f / static_cast<float>(K), eps};
    return launch_gemm<256, 2>(A, lda, Wf, ldw, M, N, K, e, S(stream), "gemm_lnfold_gelu_bf16", OutDesc{out, ldo});
  }
  EpiTmaBf16LN<false> e{bias_f, colsum, st, K / 128, M, 1.f / static_cast<float>(K), eps};
  return launch_gemm<256, 2>(A, lda, Wf, ldw, M, N, K, e, S(stream), "gemm_lnfold_bf16", OutDesc{out, ldo});
}

int vfm_gemm_lnfold_rope_bf16(const void* A, int lda, const void* Wf, int ldw, const float* bias_f, const float* colsum,
                              const float* stats, float eps, void* out, int ldo, int M, int N, int K, const float* cos_t,
                              const float* sin_t, int rope_cols, int tokens_per_seq, void* stream) {
  if (!out || !bias_f || !colsum || !stats || (ldo % 8) || (K % 256) || (reinterpret_cast<uintptr_t>(stats) & 15) || !cos_t || !sin_t ||
      tokens_per_seq < 2 || rope_cols < 0 || (rope_cols % 64) || rope_cols > N ||
      ((reinterpret_cast<uintptr_t>(cos_t) | reinterpret_cast<uintptr_t>(sin_t)) & 15))
    return fail(VFM_ERR_INVALID, "gemm_lnfold_rope_bf16: bad args (K %% 256 == 0, rope_cols %% 64 == 0, 16-byte aligned tables / stats)");
  EpiTmaBf16LNRope e{{bias_f, colsum, reinterpret_cast<const float2*>(stats), K / 128, M, 1.f / static_cast<float>(K), eps},
                     cos_t, sin_t, rope_cols, FastDiv(tokens_per_seq)};
  return launch_gemm<256, 2>(A, lda, Wf, ldw, M, N, K, e, S(stream), "gemm_lnfold_rope_bf16", OutDesc{out, ldo});
}

int vfm_gemm_patch_embed_ex(const void* A, int lda, const void* W, int ldw, const float* bias, const float* pos, float* x,
                            int patches, int cls_rows, int M, int N, int K, void* stream) {
  if (!x || !bias || !pos || patches <= 0 || (M % patches) || (cls_rows & ~1)) return fail(VFM_ERR_INVALID, "gemm_patch_embed: bad args");
  if ((patches % 32) == 0 && (N % 4) == 0) {   // a 32-row TMA box stays inside one crop: fp32 TMA stores
    EpiTmaPatchEmbed e{bias, pos, N, FastDiv(patches), cls_rows};
    OutDesc od{x, N};
    od.rows = M / patches * (patches + cls_rows);
    return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_patch_embed", od);
  }
  EpiPatchEmbed e{x, N, bias, pos, FastDiv(patches), cls_rows};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_patch_embed");
}

int vfm_gemm_patch_embed(const void* A, int lda, const void* W, int ldw, const float* bias, const float* pos, float* x,
                         int patches, int M, int N, int K, void* stream) {
  return vfm_gemm_patch_embed_ex(A, lda, W, ldw, bias, pos, x, patches, 1, M, N, K, stream);
}

int vfm_gemm_convt2x2_gelu(const void* A, int lda, const void* W, int ldw, const float* bias, void* out, int c_out,
                           int h, int w, int M, int K, void* stream) {
  if (!out || !bias || c_out <= 0 || (c_out % 32) || h <= 0 || w <= 0 || (M % (h * w)))
    return fail(VFM_ERR_INVALID, "gemm_convt2x2_gelu: bad args (c_out %% 32 == 0, M %% (h*w) == 0)");
  if (w % 32 == 0 && c_out % 64 == 0) {   // pixel shuffle by the TMA engine (4-D store)
    EpiTmaConvT e{bias, c_out, h, FastDiv(h * w), FastDiv(w)};
    OutDesc od;
    od.ptr = out; od.ld = c_out; od.convt_w = w; od.convt_rows = (M / (h * w)) * 2 * h;
    return launch_gemm<256, 2>(A, lda, W, ldw, M, 4 * c_out, K, e, S(stream), "gemm_convt2x2_gelu", od);
  }
  EpiConvT2x2Gelu e{BF(out), bias, c_out, h, w, FastDiv(h * w), FastDiv(w)};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, 4 * c_out, K, e, S(stream), "gemm_convt2x2_gelu");
}

int vfm_gemm_cls_nchw(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int num_classes,
                      int pix_per_crop, int M, int K, void* stream) {
  if (!out || !bias || num_classes <= 0 || num_classes > 32 || pix_per_crop <= 0 || (M % pix_per_crop))
    return fail(VFM_ERR_INVALID, "gemm_cls_nchw: bad args (num_classes <= 32, M %% pix_per_crop == 0)");
  EpiClsNCHW e{out, bias, num_classes, FastDiv(pix_per_crop)};
  return launch_gemm<32, 1>(A, lda, W, ldw, M, 32, K, e, S(stream), "gemm_cls_nchw");
}

int vfm_gemm_f32(const void* A, int lda, const void* W, int ldw, const float* bias, float* out, int ldo, int M, int N,
                 int K, void* stream) {
  if (!out || (ldo % 4)) return fail(VFM_ERR_INVALID, "gemm_f32: bad out/ldo");
  EpiF32 e{out, ldo, bias};
  return launch_gemm<256, 2>(A, lda, W, ldw, M, N, K, e, S(stream), "gemm_f32");
}

// Generic launcher: Q [*, q_ld], K [*, k_ld], V [*, v_ld] bf16; sequence s owns rows [s*q_seq_rows, +q_total) of Q/out
// and [s*kv_seq_rows, +kv_total) of K/V. mode: 0 auto, 1 tensor tiles over every token, 2 extra-token split
// (token 0 of each sequence handled on the CUDA cores; needs q_total == kv_total).
static int launch_attention(const void* q, int q_ld, int q_col0, const void* k, int k_ld, int k_col0, const void* v, int v_ld,
                     int v_col0, void* out, int out_ld, int n_seq, int q_seq_rows, int kv_seq_rows, int q_total,
                     int kv_total, int heads, int mode, cudaStream_t st) {
  if (!q || !k || !v || !out || n_seq <= 0 || q_total <= 0 || kv_total <= 0 || heads <= 0)
    return fail(VFM_ERR_INVALID, "attention: bad args");
  if (mode < 0 || mode > 7) return fail(VFM_ERR_INVALID, "attention: mode must be 0..7");
  if ((out_ld % 8) || (reinterpret_cast<uintptr_t>(out) & 15)) return fail(VFM_ERR_INVALID, "attention: out must be 16-byte aligned");
  if (mode == 0) {   // experiment knob (tools/): VFM_ATT_MODE=1|2|3 overrides the automatic choice
    static const int forced = [] { const char* e = std::getenv("VFM_ATT_MODE"); return e ? std::atoi(e) : 0; }();
    if (forced >= 1 && forced <= 7 && !((forced == 2 || forced == 5 || forced == 7) && (q_total != kv_total || kv_total < 2))) mode = forced;
  }
  if (mode == 0) {
    // Automatic choice. ViT windows (1 cls token + a multiple of 256 patch tokens, self attention): the ping-pong kernel in
    // extra-token mode, whatever the number of windows (the choice must not depend on the batch: a window gives the same
    // bits alone and inside a pass of 36, tests/test_e2e_gpu.py) — 0.263 ms per 36 x 16 x 1025 launch against 0.303
    // (mode 1), 0.320 (mode 4) and 0.269 for cuDNN's SDPA on the same box (profiles/r2_attn_library_comparator.json).
    // Everything else: the round-1 kernel over every token.
    const int body = kv_total - 1;
    if (q_total == kv_total && body >= 512 && body % APP_UNIT_Q == 0 && kv_total <= APP_MAX_EXTRA_KEYS) mode = 5;
  }
  if (mode >= 4) {
    // ping-pong kernel (attention_pp_sm100.cuh): 256-query units, 128-key tiles, one persistent CTA per SM; 5 = extra-token split;
    // 6 / 7 = the same units with the per-half softmax pipeline (attention_ph_sm100.cuh), 7 = extra-token split
    const bool per_half = mode >= 6;
    const int ex = mode & 1;
    if (ex && (q_total != kv_total || kv_total < 2)) return fail(VFM_ERR_INVALID, "attention: extra-token mode needs q_total == kv_total >= 2");
    if (ex && kv_total > APP_MAX_EXTRA_KEYS) return fail(VFM_ERR_INVALID, "attention: sequence too long for extra-token mode (%d keys)", kv_total);
    AttParams p{};
    p.extra = ex;
    p.sched_mask = -1;
    p.q_len = q_total - ex; p.kv_len = kv_total - ex;
    p.q_seq_rows = q_seq_rows; p.kv_seq_rows = kv_seq_rows;
    p.q_row_off = ex; p.kv_row_off = ex;
    p.heads = heads;
    p.q_tiles = (p.q_len + APP_UNIT_Q - 1) / APP_UNIT_Q;   // units per (sequence, head)
    p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
    p.q_ptr = BF(q); p.q_ld = q_ld;
    p.k_ptr = BF(k); p.v_ptr = BF(v); p.k_ld = k_ld; p.v_ld = v_ld;
    p.out = BF(out); p.out_ld = out_ld;
    const uint64_t q_rows = static_cast<uint64_t>(n_seq - 1) * q_seq_rows + q_total;
    const uint64_t kv_rows = static_cast<uint64_t>(n_seq - 1) * kv_seq_rows + kv_total;
    CUtensorMap tq, tk, tv;   // 128-row boxes
    int rc;
    if ((rc = make_tmap(&tq, q, q_rows, q_col0 + heads * ATT_D, q_ld, 128))) return rc;
    if ((rc = make_tmap(&tk, k, kv_rows, k_col0 + heads * ATT_D, k_ld, 128))) return rc;
    if ((rc = make_tmap(&tv, v, kv_rows, v_col0 + heads * ATT_D, v_ld, 128))) return rc;
    static bool attr_pp = false;
    if (!attr_pp) {
      VFM_CUDA(cudaFuncSetAttribute(attention_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, APP_SMEM_BYTES));
      VFM_CUDA(cudaFuncSetAttribute(attention_ph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, APP_SMEM_BYTES));
      attr_pp = true;
    }
    const long long units = static_cast<long long>(n_seq) * heads * p.q_tiles;
    if (units > 0x7fffffffLL) return fail(VFM_ERR_INVALID, "attention: too many units");
    p.n_units = static_cast<int>(units);
    const int n_sm = sm_count();
    if (n_sm <= 0) return fail(VFM_ERR_CUDA, "attention: no CUDA device");
    const unsigned grid = static_cast<unsigned>(units < n_sm ? units : n_sm);
    {
      LaunchScope scope("attention_fwd", st);   // same family as the round-1 kernel in the per-launch profile (bench.py roofline.families)
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(grid);
      cfg.blockDim = dim3(APP_THREADS);
      cfg.dynamicSmemBytes = APP_SMEM_BYTES;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled() ? 1 : 0;
      cudaError_t e = cudaLaunchKernelEx(&cfg, per_half ? attention_ph_kernel : attention_pp_kernel, tq, tk, tv, p);
      if (e != cudaSuccess) return fail(VFM_ERR_CUDA, "launch of attention_pp failed: %s", cudaGetErrorString(e));
    }
    VFM_LAUNCH_CHECK("attention_pp");
    return VFM_OK;
  }
  bool extra = mode == 2;
  // mode 0: plain tiles. The split saves a ninth query tile and a seventeenth key tile at 1025 tokens, but measured
  // 0.189 vs 0.167 ms on B200: its per-CTA prologue/epilogue loads and the appended query CTAs cost more than the tiles.
  if (mode == 0) extra = false;   // (unreachable with mode 5 chosen above; kept for the non-ViT shapes: mode 0 == mode 1)
  if (extra && (q_total != kv_total || kv_total < 2)) return fail(VFM_ERR_INVALID, "attention: extra-token mode needs q_total == kv_total >= 2");
  AttParams p{};
  p.extra = extra ? 1 : 0;
  p.q_len = q_total - p.extra; p.kv_len = kv_total - p.extra;
  p.q_seq_rows = q_seq_rows; p.kv_seq_rows = kv_seq_rows;
  p.q_row_off = p.extra; p.kv_row_off = p.extra;
  p.heads = heads;
  p.q_tiles = (p.q_len + ATT_BLOCK_Q - 1) / ATT_BLOCK_Q;
  p.q_col0 = q_col0; p.k_col0 = k_col0; p.v_col0 = v_col0;
  p.k_ptr = BF(k); p.v_ptr = BF(v); p.k_ld = k_ld; p.v_ld = v_ld;
  p.out = BF(out); p.out_ld = out_ld;
  const uint64_t q_rows = static_cast<uint64_t>(n_seq - 1) * q_seq_rows + q_total;
  const uint64_t kv_rows = static_cast<uint64_t>(n_seq - 1) * kv_seq_rows + kv_total;
  CUtensorMap tq, tk, tv;   // 64-row boxes (K/V tiles; a Q tile = two boxes)
  int rc;
  if ((rc = make_tmap(&tq, q, q_rows, q_col0 + heads * ATT_D, q_ld, 64))) return rc;
  if ((rc = make_tmap(&tk, k, kv_rows, k_col0 + heads * ATT_D, k_ld, 64))) return rc;
  if ((rc = make_tmap(&tv, v, kv_rows, v_col0 + heads * ATT_D, v_ld, 64))) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    VFM_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES));
    attr_done = true;
  }
  const long long units = static_cast<long long>(n_seq) * heads * p.q_tiles;
  if (units > 0x7fffffffLL) return fail(VFM_ERR_INVALID, "attention: too many units");
  p.n_units = static_cast<int>(units);
  p.q_ptr = BF(q); p.q_ld = q_ld;
  static int n_sm = 0;
  if (!n_sm) {
    int dev = 0;
    VFM_CUDA(cudaGetDevice(&dev));
    VFM_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  if (mode == 3) {   // experimental: serial kernel, four CTAs per SM (attention_serial_kernel)
    static bool attr3 = false;
    if (!attr3) {
      VFM_CUDA(cudaFuncSetAttribute(attention_serial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATS_SMEM_BYTES));
      attr3 = true;
    }
    {
      LaunchScope scope("attention_fwd_serial", st);
      attention_serial_kernel<<<static_cast<unsigned>(units), ATS_THREADS, ATS_SMEM_BYTES, st>>>(tq, tk, tv, p);
    }
    VFM_LAUNCH_CHECK("attention_fwd_serial");
    return VFM_OK;
  }
  const unsigned grid = static_cast<unsigned>(units < 2LL * n_sm ? units : 2LL * n_sm);   // persistent: two CTAs per SM
  {
    LaunchScope scope("attention_fwd", st);
    attention_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tq, tk, tv, p);
  }
  if (extra) {
    // one CTA per (sequence, head) for the extra token's query row (CUDA cores)
    const size_t need = (static_cast<size_t>((kv_total + 3) & ~3) + (ATT_THREADS / 32) * 65) * sizeof(float);
    if (need > 200 * 1024) return fail(VFM_ERR_INVALID, "attention: sequence too long for extra-token mode (%d keys)", kv_total);
    static size_t attr_need = 0;
    if (need > attr_need) {
      VFM_CUDA(cudaFuncSetAttribute(attention_extra_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(need)));
      attr_need = need;
    }
    LaunchScope scope("attention_extra_query", st);
    attention_extra_query_kernel<<<static_cast<unsigned>(n_seq * heads), ATT_THREADS, need, st>>>(p);
  }
  VFM_LAUNCH_CHECK("attention_fwd");
  return VFM_OK;
}

int vfm_attention_fwd(const void* qkv, void* out, int n_seq, int seq_len, int heads, void* stream) {
  return vfm_attention_fwd_ex(qkv, out, n_seq, seq_len, heads, 0, stream);
}

int vfm_attention_fwd_ex(const void* qkv, void* out, int n_seq, int seq_len, int heads, int mode, void* stream) {
  const int C = heads * ATT_D;
  const __nv_bfloat16* base = BF(qkv);
  return launch_attention(base, 3 * C, 0, base, 3 * C, C, base, 3 * C, 2 * C, out, C, n_seq, seq_len, seq_len, seq_len,
                          seq_len, heads, mode, S(stream));
}

// ---------------------------------------------------------------------------------------------------------------
// SAM ViT pieces (sam_ops.cuh)
int vfm_relpos_terms(const void* qkv, const float* Rh, const float* Rw, float* rel, int n_seq, int heads, int head_dim,
                     int q_h, int q_w, int k_h, int k_w, void* stream) {
  if (!qkv || !Rh || !Rw || !rel || n_seq <= 0 || heads <= 0 || q_h <= 0 || q_w <= 0 || k_h <= 0 || k_w <= 0)
    return fail(VFM_ERR_INVALID, "relpos_terms: bad args");
  if (head_dim != 80 && head_dim != 64) return fail(VFM_ERR_INVALID, "relpos_terms: head_dim must be 64 or 80 (got %d)", head_dim);
  const int kn = k_h > k_w ? k_h : k_w;
  const size_t smem = (static_cast<size_t>(kn) * (head_dim + 1) + 8 * head_dim) * sizeof(float);
  if (smem > 48 * 1024) return fail(VFM_ERR_INVALID, "relpos_terms: key grid too large (%d x %d)", k_h, k_w);
  const dim3 grid(q_h + q_w, n_seq);
  {
    LaunchScope scope("relpos_terms", S(stream));
    if (head_dim == 80)
      relpos_terms_kernel<80><<<grid, 256, smem, S(stream)>>>(BF(qkv), 3 * heads * head_dim, Rh, Rw, rel, q_h * q_w, heads, q_h, q_w, k_h, k_w);
    else
      relpos_terms_kernel<64><<<grid, 256, smem, S(stream)>>>(BF(qkv), 3 * heads * head_dim, Rh, Rw, rel, q_h * q_w, heads, q_h, q_w, k_h, k_w);
  }
  VFM_LAUNCH_CHECK("relpos_terms");
  return VFM_OK;
}

int vfm_rows_gather(const void* src, void* dst, const int* map, long long n_rows, int C, void* stream) {
  if (!src || !dst || !map || n_rows <= 0 || C <= 0 || (C % 8)) return fail(VFM_ERR_INVALID, "rows_gather: bad args (C %% 8 == 0)");
  long long blocks = (n_rows * (C / 8) + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  {
    LaunchScope scope("rows_gather", S(stream));
    rows_gather_kernel<<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(BF(src), const_cast<__nv_bfloat16*>(BF(dst)), map, n_rows, C);
  }
  VFM_LAUNCH_CHECK("rows_gather");
  return VFM_OK;
}

extern "C++" {
template <int D, int WARPS, int KV>
static int launch_attention_relpos_cfg(const void* qkv, int ld, int g_col0, const float* rel, void* out, int n_seq, int seq_len,
                                       int heads, int k_h, int k_w, float scale, cudaStream_t st) {
  using Cfg = RelposCfg<D, WARPS, KV>;
  const int kk = (rel || g_col0 >= 0) ? k_h + k_w : 0;
  const size_t smem = Cfg::bytes(kk);
  if (smem > 220 * 1024) return fail(VFM_ERR_INVALID, "attention_relpos: key grid too large (%d x %d)", k_h, k_w);
  static size_t attr = 0;
  if (smem > attr) {
    VFM_CUDA(cudaFuncSetAttribute(attention_relpos_kernel<D, WARPS, KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr = smem;
  }
  const dim3 grid((seq_len + Cfg::kBQ - 1) / Cfg::kBQ, heads, n_seq);
  {
    LaunchScope scope("attention_relpos", st);
    attention_relpos_kernel<D, WARPS, KV><<<grid, Cfg::kThreads, smem, st>>>(BF(qkv), ld, g_col0, rel, const_cast<__nv_bfloat16*>(BF(out)),
                                                                           seq_len, heads, k_h, k_w, scale);
  }
  VFM_LAUNCH_CHECK("attention_relpos");
  return VFM_OK;
}
template <int D>
static int launch_attention_relpos(const void* qkv, int ld, int g_col0, const float* rel, void* out, int n_seq, int seq_len,
                                   int heads, int k_h, int k_w, float scale, cudaStream_t st) {
  // short sequences (SAM's 14 x 14 windows): tiles cut to fit 196 tokens; long ones: 128 query rows share each K / V tile
  if (seq_len <= 224) return launch_attention_relpos_cfg<D, 7, 112>(qkv, ld, g_col0, rel, out, n_seq, seq_len, heads, k_h, k_w, scale, st);
  return launch_attention_relpos_cfg<D, 8, 64>(qkv, ld, g_col0, rel, out, n_seq, seq_len, heads, k_h, k_w, scale, st);
}
}  // extern "C++"

int vfm_attention_relpos_ex(const void* qkv, int ld, int g_col0, const float* rel, void* out, int n_seq, int seq_len, int heads,
                            int head_dim, int k_h, int k_w, float scale, void* stream) {
  if (!qkv || !out || n_seq <= 0 || seq_len <= 0 || heads <= 0) return fail(VFM_ERR_INVALID, "attention_relpos: bad args");
  const bool bias = rel != nullptr || g_col0 >= 0;
  if (rel && g_col0 >= 0) return fail(VFM_ERR_INVALID, "attention_relpos: give either rel or g_col0, not both");
  if (bias && (k_h <= 0 || k_w <= 0 || k_h * k_w != seq_len || (k_h & 1) || (k_w & 1)))
    return fail(VFM_ERR_INVALID, "attention_relpos: with a bias k_h * k_w must equal seq_len and both be even (%d x %d vs %d)", k_h, k_w, seq_len);
  if (ld < 3 * heads * head_dim || (ld % 8)) return fail(VFM_ERR_INVALID, "attention_relpos: bad row pitch %d", ld);
  if (g_col0 >= 0 && (g_col0 < 3 * heads * head_dim || g_col0 + heads * (2 * k_h - 1 + 2 * k_w - 1) > ld))
    return fail(VFM_ERR_INVALID, "attention_relpos: table-term columns [%d, ...) do not fit the row pitch %d", g_col0, ld);
  if (n_seq > 65535 || heads > 65535) return fail(VFM_ERR_INVALID, "attention_relpos: grid too large");
  if (head_dim == 80) return launch_attention_relpos<80>(qkv, ld, g_col0, rel, out, n_seq, seq_len, heads, k_h, k_w, scale, S(stream));
  if (head_dim == 64) return launch_attention_relpos<64>(qkv, ld, g_col0, rel, out, n_seq, seq_len, heads, k_h, k_w, scale, S(stream));
  return fail(VFM_ERR_INVALID, "attention_relpos: head_dim must be 64 or 80 (got %d)", head_dim);
}

int vfm_attention_window_tc(const void* qkv, int ld, int g_col0, void* out, int n_seq, int seq_len, int heads, int head_dim,
                            int k_h, int k_w, float scale, void* stream) {
  return vfm_attention_window_tc_map(qkv, ld, g_col0, out, nullptr, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, stream);
}

int vfm_attention_window_tc_map(const void* qkv, int ld, int g_col0, void* out, const int* out_map, int n_seq, int seq_len, int heads,
                                int head_dim, int k_h, int k_w, float scale, void* stream) {
  if (!qkv || !out || n_seq <= 0 || seq_len <= 0 || heads <= 0) return fail(VFM_ERR_INVALID, "attention_window_tc: bad args");
  if (head_dim != WIN_D) return fail(VFM_ERR_INVALID, "attention_window_tc: head_dim must be %d (got %d)", WIN_D, head_dim);
  if (seq_len > WIN_KEYS) return fail(VFM_ERR_INVALID, "attention_window_tc: at most %d tokens per window (got %d)", WIN_KEYS, seq_len);
  if (ld < 3 * heads * head_dim || (ld % 8)) return fail(VFM_ERR_INVALID, "attention_window_tc: bad row pitch %d", ld);
  if (g_col0 >= 0 && (k_h <= 0 || k_w <= 0 || k_h * k_w != seq_len || k_h > 16 || k_w > 16 || g_col0 < 3 * heads * head_dim ||
                      g_col0 + heads * (2 * k_h - 1 + 2 * k_w - 1) > ld))
    return fail(VFM_ERR_INVALID, "attention_window_tc: bad key grid %d x %d / table-term columns from %d (row pitch %d)", k_h, k_w, g_col0, ld);
  if (n_seq > 65535 || heads > 65535) return fail(VFM_ERR_INVALID, "attention_window_tc: grid too large");
  if ((reinterpret_cast<uintptr_t>(out) & 15)) return fail(VFM_ERR_INVALID, "attention_window_tc: out must be 16-byte aligned");
  const uint64_t rows = static_cast<uint64_t>(n_seq) * seq_len;
  CUtensorMap tq, tkv, tq16, tkv16;
  int rc;
  if ((rc = make_tmap(&tq, qkv, rows, ld, ld, 64))) return rc;
  if ((rc = make_tmap(&tkv, qkv, rows, ld, ld, WIN_KEYS / 2))) return rc;
  if ((rc = make_tmap_sw(&tq16, qkv, rows, ld, ld, 64, 16))) return rc;
  if ((rc = make_tmap_sw(&tkv16, qkv, rows, ld, ld, WIN_KEYS / 2, 16))) return rc;
  static bool attr_done = false;
  if (!attr_done) {
    VFM_CUDA(cudaFuncSetAttribute(attention_win_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WIN_SMEM_BYTES));
    attr_done = true;
  }
  WinParams p{};
  p.seq_len = seq_len; p.heads = heads; p.k_h = k_h; p.k_w = k_w; p.ld = ld; p.g_col0 = g_col0; p.scale = scale;
  p.qkv = BF(qkv); p.out = const_cast<__nv_bfloat16*>(BF(out)); p.out_map = out_map;
  const dim3 grid((seq_len + WIN_BLOCK_Q - 1) / WIN_BLOCK_Q, heads, n_seq);
  {
    LaunchScope scope("attention_window_tc", S(stream));
    attention_win_kernel<<<grid, WIN_THREADS, WIN_SMEM_BYTES, S(stream)>>>(tq, tkv, tq16, tkv16, p);
  }
  VFM_LAUNCH_CHECK("attention_window_tc");
  return VFM_OK;
}

extern "C++" {
template <int NA>
static int launch_attention_glob(const CUtensorMap& t64, const CUtensorMap& t16, const CUtensorMap& te, const GlobParams& p,
                                 int n_seq, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    VFM_CUDA(cudaFuncSetAttribute(attention_glob_kernel<NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, glb_smem_bytes<NA>()));
    attr_done = true;
  }
  const dim3 grid((p.seq_len + GLB_BLOCK_Q - 1) / GLB_BLOCK_Q, p.heads, n_seq);
  {
    LaunchScope scope("attention_global_tc", st);
    attention_glob_kernel<NA><<<grid, GLB_THREADS, glb_smem_bytes<NA>(), st>>>(t64, t16, te, p);
  }
  VFM_LAUNCH_CHECK("attention_global_tc");
  return VFM_OK;
}
}  // extern "C++"

int vfm_attention_global_tc(const void* qkv, int ld, int g_col0, const void* onehot, int onehot_rows, void* out, int n_seq,
                            int seq_len, int heads, int head_dim, int k_h, int k_w, float scale, void* stream) {
  if (!qkv || !out || !onehot || n_seq <= 0 || seq_len <= 0 || heads <= 0) return fail(VFM_ERR_INVALID, "attention_global_tc: bad args");
  if (head_dim != GLB_D) return fail(VFM_ERR_INVALID, "attention_global_tc: head_dim must be %d (got %d)", GLB_D, head_dim);
  if (ld < 3 * heads * head_dim || (ld % 8)) return fail(VFM_ERR_INVALID, "attention_global_tc: bad row pitch %d", ld);
  const int bh = (k_h + 15) & ~15, bw = (k_w + 15) & ~15;
  const int na = (bh + bw + 63) / 64;
  if (k_h <= 0 || k_w <= 0 || k_h * k_w != seq_len || na > 2 || g_col0 < 3 * heads * head_dim ||
      g_col0 + heads * (2 * k_h - 1 + 2 * k_w - 1) > ld)
    return fail(VFM_ERR_INVALID, "attention_global_tc: bad key grid %d x %d (at most 128 bias columns) / table-term columns from %d (row pitch %d)",
                k_h, k_w, g_col0, ld);
  if (onehot_rows < seq_len) return fail(VFM_ERR_INVALID, "attention_global_tc: one-hot matrix has %d rows, needs %d", onehot_rows, seq_len);
  if (n_seq > 65535 || heads > 65535) return fail(VFM_ERR_INVALID, "attention_global_tc: grid too large");
  if ((reinterpret_cast<uintptr_t>(out) & 15)) return fail(VFM_ERR_INVALID, "attention_global_tc: out must be 16-byte aligned");
  const uint64_t rows = static_cast<uint64_t>(n_seq) * seq_len;
  CUtensorMap t64, t16, te;
  int rc;
  if ((rc = make_tmap(&t64, qkv, rows, ld, ld, 64))) return rc;
  if ((rc = make_tmap_sw(&t16, qkv, rows, ld, ld, 64, 16))) return rc;
  if ((rc = make_tmap(&te, onehot, onehot_rows, 64 * na, 64 * na, 64))) return rc;
  GlobParams p{};
  p.seq_len = seq_len; p.heads = heads; p.k_h = k_h; p.k_w = k_w; p.bh = bh; p.bw = bw; p.ld = ld; p.g_col0 = g_col0; p.scale = scale;
  p.qkv = BF(qkv); p.out = const_cast<__nv_bfloat16*>(BF(out));
  return na == 1 ? launch_attention_glob<1>(t64, t16, te, p, n_seq, S(stream)) : launch_attention_glob<2>(t64, t16, te, p, n_seq, S(stream));
}

int vfm_attention_relpos(const void* qkv, const float* rel, void* out, int n_seq, int seq_len, int heads, int head_dim,
                         int k_h, int k_w, float scale, void* stream) {
  return vfm_attention_relpos_ex(qkv, 3 * heads * head_dim, -1, rel, out, n_seq, seq_len, heads, head_dim, k_h, k_w, scale, stream);
}

int vfm_attention_cross(const void* q, int q_ld, const void* kv, int kv_ld, void* out, int out_ld, int n_seq, int q_len,
                        int kv_len, int heads, void* stream) {
  const int C = heads * ATT_D;
  const __nv_bfloat16* kvb = BF(kv);
  return launch_attention(q, q_ld, 0, kvb, kv_ld, 0, kvb, kv_ld, C, out, out_ld, n_seq, q_len, kv_len, q_len, kv_len, heads,
                          1, S(stream));
}

int vfm_patch_gather(const void* img, int is_u8, const VfmPixelNorm* nrm, int img_h, int img_w, const int* crops,
                     int n_crops, int gh, int gw, void* out, void* stream) {
  if (!img || !crops || !out || n_crops <= 0 || gh <= 0 || gw <= 0) return fail(VFM_ERR_INVALID, "patch_gather: bad args");
  if (is_u8 && !nrm) return fail(VFM_ERR_INVALID, "patch_gather: uint8 input needs a VfmPixelNorm");
  PixelNorm pn{};
  if (nrm) {
    for (int i = 0; i < 3; ++i) { pn.mean[i] = nrm->mean[i]; pn.inv_std[i] = nrm->inv_std[i]; }
    pn.flip = nrm->flip;
  }
  const long long total = static_cast<long long>(n_crops) * gh * gw * 96;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (blocks > cap) blocks = cap;
{
    LaunchScope scope("patch_gather", S(stream));
    if (is_u8)
      patch_gather_kernel<uint8_t><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(
          reinterpret_cast<const uint8_t*>(img), img_h, img_w, reinterpret_cast<const int4*>(crops), n_crops, gh, gw, pn, BF(out));
    else
      patch_gather_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(
          reinterpret_cast<const float*>(img), img_h, img_w, reinterpret_cast<const int4*>(crops), n_crops, gh, gw, pn, BF(out));
  }
  VFM_LAUNCH_CHECK("patch_gather");
  return VFM_OK;
}

int vfm_cls_rows(float* x, const float* cls_token, const float* pos, int n_crops, int tokens, int C, void* stream) {
  if (!x || !cls_token || !pos) return fail(VFM_ERR_INVALID, "cls_rows: null");
  const int n = n_crops * C;
{
    LaunchScope scope("cls_rows", S(stream));
    cls_rows_kernel<<<(n + 255) / 256, 256, 0, S(stream)>>>(x, cls_token, pos, n_crops, tokens, C);
  }
  VFM_LAUNCH_CHECK("cls_rows");
  return VFM_OK;
}

int vfm_layernorm_tap(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                      void* tap, int tap_ld, int tap_col0, int tokens_per_crop, void* stream) {
  return vfm_layernorm_tap_ex(x, gamma, beta, out, M, C, eps, tap, tap_ld, tap_col0, tokens_per_crop, 1, stream);
}

int vfm_layernorm_tap_ex(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                         void* tap, int tap_ld, int tap_col0, int tokens_per_crop, int cls_rows, void* stream) {
  return vfm_layernorm_tap_map(x, gamma, beta, out, M, C, eps, tap, tap_ld, tap_col0, tokens_per_crop, cls_rows, nullptr, stream);
}

int vfm_layernorm_tap_map(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                          void* tap, int tap_ld, int tap_col0, int tokens_per_crop, int cls_rows, const int* out_map, void* stream) {
  if (!x || M <= 0) return fail(VFM_ERR_INVALID, "layernorm: bad args");
  if (!out && !tap) return fail(VFM_ERR_INVALID, "layernorm: neither an output nor a tap was given");
  if (out && (!gamma || !beta)) return fail(VFM_ERR_INVALID, "layernorm: null gamma/beta");
  if (C % 128 || C < 128 || C > 1536) return fail(VFM_ERR_INVALID, "layernorm: C must be a multiple of 128 in [128,1536] (C=%d)", C);
  if (tap && ((tap_ld % 8) || (tap_col0 % 8) || tokens_per_crop <= 0)) return fail(VFM_ERR_INVALID, "layernorm: bad tap layout");
  const int grid = (M + 7) / 8;
  cudaStream_t st = S(stream);
  const FastDiv fd(tokens_per_crop > 0 ? tokens_per_crop : 1);
  {
    LaunchScope scope("layernorm", st);
    switch (C / 128) {
#define VFM_LN_CASE(I) case I: layernorm_kernel<I><<<grid, 256, 0, st>>>(x, gamma, beta, BF(out), M, eps, BF(tap), tap_ld, tap_col0, fd, cls_rows, out_map); break;
      VFM_LN_CASE(1) VFM_LN_CASE(2) VFM_LN_CASE(3) VFM_LN_CASE(4) VFM_LN_CASE(5) VFM_LN_CASE(6) VFM_LN_CASE(7) VFM_LN_CASE(8)
      VFM_LN_CASE(9) VFM_LN_CASE(10) VFM_LN_CASE(11) VFM_LN_CASE(12)
#undef VFM_LN_CASE
    }
  }
  VFM_LAUNCH_CHECK("layernorm");
  return VFM_OK;
}

int vfm_layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int C, float eps,
                  void* stream) {
  if (!out) return fail(VFM_ERR_INVALID, "layernorm: null output");
  return vfm_layernorm_tap(x, gamma, beta, out, M, C, eps, nullptr, 0, 0, 1, stream);
}

int vfm_groupnorm_relu(const void* in, void* out, const float* gamma, const float* beta, int n_crops, int P, int C,
                       int groups, float eps, int relu, void* stream) {
  if (!in || !out || !gamma || !beta || n_crops <= 0 || P <= 0 || groups <= 0 || C % groups || (C / groups) % 8)
    return fail(VFM_ERR_INVALID, "groupnorm_relu: bad args (C/groups must be a multiple of 8)");
{
    LaunchScope scope("groupnorm_relu", S(stream));
    groupnorm_relu_kernel<<<n_crops * groups, 256, 0, S(stream)>>>(BF(in), BF(out), gamma, beta, P, C, groups, eps, relu);
  }
  VFM_LAUNCH_CHECK("groupnorm_relu");
  return VFM_OK;
}

// VFM_MERGE_MODE=1: the round-1 window-major tile kernel (76 accumulators per thread) instead of the class-major one; 2 / 3: the
// class-major kernel compiled for 2 / 3 resident CTAs per SM instead of 4 (tools/bench_merge.py: 215 / 190 / 185 us per 2 images). Read on every call (one getenv
// per merge launch) so that a test can compare the kernels bit for bit inside one process.
static int merge_mode() {
  const char* e = getenv("VFM_MERGE_MODE");
  return e ? atoi(e) : 0;
}
static void launch_merge_tiles(const float* lowres, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh, int lw,
                               int H, int W, int n_img, uint8_t* labels, float* logits_out, const float* flip_a, cudaStream_t st) {
  const dim3 grid((W + MERGE_TW - 1) / MERGE_TW, (H + MERGE_TH - 1) / MERGE_TH, n_img);
  const int2* bx = reinterpret_cast<const int2*>(boxes);
  switch (merge_mode()) {
    case 1: slide_merge_tile_kernel<19><<<grid, 256, 0, st>>>(lowres, bx, n_crops, nc, crop_h, crop_w, lh, lw, H, W, labels, logits_out, flip_a); break;
    case 2: slide_merge_class_kernel<19, 2><<<grid, 256, 0, st>>>(lowres, bx, n_crops, nc, crop_h, crop_w, lh, lw, H, W, labels, logits_out, flip_a); break;
    case 3: slide_merge_class_kernel<19, 3><<<grid, 256, 0, st>>>(lowres, bx, n_crops, nc, crop_h, crop_w, lh, lw, H, W, labels, logits_out, flip_a); break;
    default: slide_merge_class_kernel<19, 4><<<grid, 256, 0, st>>>(lowres, bx, n_crops, nc, crop_h, crop_w, lh, lw, H, W, labels, logits_out, flip_a);
  }
}

int vfm_slide_merge_argmax(const float* lowres, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh,
                           int lw, int H, int W, int n_img, uint8_t* labels, float* logits_out, void* stream) {
  if (!lowres || !boxes || !labels || n_crops <= 0 || nc <= 0 || nc > 32 || n_img <= 0)
    return fail(VFM_ERR_INVALID, "slide_merge_argmax: bad args (num_classes <= 32)");
  const long long total = static_cast<long long>(n_img) * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 32;
  if (blocks > cap) blocks = cap;
  const size_t smem = sizeof(int2) * n_crops;
{
    LaunchScope scope("slide_merge_argmax", S(stream));
    const bool strips = (W % 4) == 0 && nc <= 19 && ((reinterpret_cast<uintptr_t>(labels) & 3) == 0) &&
                        (!logits_out || (reinterpret_cast<uintptr_t>(logits_out) & 15) == 0);
    const bool tiles = strips && crop_h == 4 * lh && crop_w == 4 * lw && n_img <= 65535;
    if (tiles) {    // low-res footprints staged through shared memory, one CTA per 64 x 16 output tile (edge tiles partly idle)
      launch_merge_tiles(lowres, boxes, n_crops, nc, crop_h, crop_w, lh, lw, H, W, n_img, labels, logits_out, nullptr, S(stream));
    } else if (strips) {   // four pixels per thread sharing their bilinear taps
      long long b4 = (total / 4 + 255) / 256;
      if (b4 > cap) b4 = cap;
      slide_merge_argmax4_kernel<19><<<static_cast<unsigned>(b4), 256, smem, S(stream)>>>(
          lowres, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h, crop_w, lh, lw, H, W, n_img, labels, logits_out);
    } else if (nc <= 19)
      slide_merge_argmax_kernel<19><<<static_cast<unsigned>(blocks), 256, smem, S(stream)>>>(
          lowres, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h, crop_w, lh, lw, H, W, n_img, labels, logits_out);
    else
      slide_merge_argmax_kernel<32><<<static_cast<unsigned>(blocks), 256, smem, S(stream)>>>(
          lowres, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h, crop_w, lh, lw, H, W, n_img, labels, logits_out);
  }
  VFM_LAUNCH_CHECK("slide_merge_argmax");
  return VFM_OK;
}

int vfm_slide_merge_flip_argmax(const float* lowres, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh,
                                int lw, int H, int W, int n_img, const float* a, uint8_t* labels, float* logits_out, void* stream) {
  if (!lowres || !boxes || !labels || !a || n_crops <= 0 || nc <= 0 || nc > 19 || n_img <= 0 || n_img > 65535)
    return fail(VFM_ERR_INVALID, "slide_merge_flip_argmax: bad args (num_classes <= 19)");
  if ((W % 4) || crop_h != 4 * lh || crop_w != 4 * lw ||
      ((reinterpret_cast<uintptr_t>(labels) & 3) | (reinterpret_cast<uintptr_t>(a) & 15) | (reinterpret_cast<uintptr_t>(logits_out) & 15)))
    return fail(VFM_ERR_INVALID, "slide_merge_flip_argmax: needs W %% 4 == 0, windows upsampled x4 and 16-byte aligned logits "
                                 "(use vfm_slide_merge_argmax + vfm_tta_flip_mean_argmax otherwise)");
  {
    LaunchScope scope("slide_merge_flip_argmax", S(stream));
    launch_merge_tiles(lowres, boxes, n_crops, nc, crop_h, crop_w, lh, lw, H, W, n_img, labels, logits_out, a, S(stream));
  }
  VFM_LAUNCH_CHECK("slide_merge_flip_argmax");
  return VFM_OK;
}

int vfm_tta_flip_mean_argmax(const float* a, const float* b, int n_img, int nc, int H, int W, uint8_t* labels, float* logits_out,
                             void* stream) {
  if (!a || !b || !labels || n_img <= 0 || nc <= 0 || nc > 255 || H <= 0 || W <= 0)
    return fail(VFM_ERR_INVALID, "tta_flip_mean_argmax: bad args (num_classes <= 255)");
  if (logits_out == b) return fail(VFM_ERR_INVALID, "tta_flip_mean_argmax: logits_out may alias a, not b (b is read mirrored)");
  const bool vec4 = (W % 4) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                                      reinterpret_cast<uintptr_t>(logits_out)) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(labels) & 3) == 0;
  const long long items = static_cast<long long>(n_img) * H * (vec4 ? W / 4 : W);
  long long blocks = (items + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;   // a multiple of the SM count, grid-stride beyond it
  if (blocks > cap) blocks = cap;
  {
    LaunchScope scope("tta_flip_mean_argmax", S(stream));
    if (vec4)
      tta_flip_mean_argmax_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(a, b, n_img, nc, H, W, labels, logits_out);
    else
      tta_flip_mean_argmax_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, S(stream)>>>(a, b, n_img, nc, H, W, labels, logits_out);
  }
  VFM_LAUNCH_CHECK("tta_flip_mean_argmax");
  return VFM_OK;
}

// ------------------------------------------------------------------------------ coarse-to-fine path (config 3)
static unsigned grid_for(long long work_items, int per_sm = 16) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<unsigned>(blocks);
}

int vfm_image_resize_norm(const void* img, int is_u8, const VfmPixelNorm* nrm, int B, int H, int W, float* out, int h, int w,
                          void* stream) {
  if (!img || !out || B <= 0 || H <= 0 || W <= 0 || h <= 0 || w <= 0) return fail(VFM_ERR_INVALID, "image_resize_norm: bad args");
  if (is_u8 && !nrm) return fail(VFM_ERR_INVALID, "image_resize_norm: uint8 input needs a VfmPixelNorm");
  PixelNorm pn{};
  if (nrm) {
    for (int i = 0; i < 3; ++i) { pn.mean[i] = nrm->mean[i]; pn.inv_std[i] = nrm->inv_std[i]; }
    pn.flip = nrm->flip;
  }
  const unsigned grid = grid_for(static_cast<long long>(B) * 3 * h * w);
  {
    LaunchScope scope("image_resize_norm", S(stream));
    if (is_u8)
      image_resize_norm_kernel<uint8_t><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const uint8_t*>(img), B, H, W, pn, out, h, w);
    else
      image_resize_norm_kernel<float><<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const float*>(img), B, H, W, pn, out, h, w);
  }
  VFM_LAUNCH_CHECK("image_resize_norm");
  return VFM_OK;
}

int vfm_ms_confidence(const float* low0, const int* boxes, int n_crops, int nc, int crop_h, int crop_w, int lh, int lw, int H,
                      int W, int n_img, float thr, int* counts, void* stream) {
  if (!low0 || !boxes || !counts || n_crops <= 0 || nc <= 0 || nc > 32 || n_img <= 0) return fail(VFM_ERR_INVALID, "ms_confidence: bad args");
  VFM_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * static_cast<size_t>(n_img) * n_crops, S(stream)));
  const int rows = 4;
  const unsigned grid = static_cast<unsigned>(n_img) * ((H + rows - 1) / rows);
  const size_t smem = (sizeof(int) + sizeof(int2)) * n_crops + 8;
  {
    LaunchScope scope("ms_confidence", S(stream));
    if (nc <= 19)
      ms_confidence_kernel<19><<<grid, 256, smem, S(stream)>>>(low0, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h, crop_w,
                                                               lh, lw, H, W, thr, rows, counts);
    else
      ms_confidence_kernel<32><<<grid, 256, smem, S(stream)>>>(low0, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h, crop_w,
                                                               lh, lw, H, W, thr, rows, counts);
  }
  VFM_LAUNCH_CHECK("ms_confidence");
  return VFM_OK;
}

int vfm_ms_context_im2col(const float* low0, const int* crops, int n_ref, int nc, int crop_h, int crop_w, int lh, int lw, int H,
                          int W, int ctx_h, int ctx_w, void* out, int kpad, void* stream) {
  if (!low0 || !crops || !out || n_ref <= 0 || nc <= 0 || (ctx_h & 1) || (ctx_w & 1) || kpad < 4 * nc || (kpad % 8))
    return fail(VFM_ERR_INVALID, "ms_context_im2col: bad args (kpad >= 4*nc, kpad %% 8 == 0, even context size)");
  const long long total = static_cast<long long>(n_ref) * (ctx_h / 2) * (ctx_w / 2) * (kpad / 4);
  {
    LaunchScope scope("ms_context_im2col", S(stream));
    ms_context_im2col_kernel<<<grid_for(total), 256, 0, S(stream)>>>(low0, reinterpret_cast<const int4*>(crops), n_ref, nc, crop_h,
                                                                     crop_w, lh, lw, H, W, ctx_h, ctx_w, BF(out), kpad);
  }
  VFM_LAUNCH_CHECK("ms_context_im2col");
  return VFM_OK;
}

int vfm_space_to_depth2(const void* in, void* out, int n, int h, int w, int C, void* stream) {
  if (!in || !out || n <= 0 || (h & 1) || (w & 1) || C <= 0 || (C % 8)) return fail(VFM_ERR_INVALID, "space_to_depth2: bad args (even h, w; C %% 8 == 0)");
  const long long total = static_cast<long long>(n) * (h / 2) * (w / 2) * 4 * (C / 8);
  {
    LaunchScope scope("space_to_depth2", S(stream));
    space_to_depth2_kernel<<<grid_for(total), 256, 0, S(stream)>>>(BF(in), BF(out), n, h, w, C);
  }
  VFM_LAUNCH_CHECK("space_to_depth2");
  return VFM_OK;
}

int vfm_groupnorm_act(const void* in, void* out, int out_f32, const float* gamma, const float* beta, int n, int P, int C, int groups,
                      float eps, int act, void* stream) {
  if (!in || !out || !gamma || !beta || n <= 0 || P <= 0 || groups <= 0 || C % groups || act < 0 || act > 2)
    return fail(VFM_ERR_INVALID, "groupnorm_act: bad args");
  {
    LaunchScope scope("groupnorm_act", S(stream));
    if (out_f32)
      groupnorm_act_kernel<float><<<n * groups, 256, 0, S(stream)>>>(BF(in), reinterpret_cast<float*>(out), gamma, beta, P, C, groups, eps, act);
    else
      groupnorm_act_kernel<__nv_bfloat16><<<n * groups, 256, 0, S(stream)>>>(BF(in), BF(out), gamma, beta, P, C, groups, eps, act);
  }
  VFM_LAUNCH_CHECK("groupnorm_act");
  return VFM_OK;
}

int vfm_geglu(const void* in, void* out, long long M, int I, void* stream) {
  if (!in || !out || M <= 0 || I <= 0 || (I % 8)) return fail(VFM_ERR_INVALID, "geglu: bad args (I %% 8 == 0)");
  {
    LaunchScope scope("geglu", S(stream));
    geglu_kernel<<<grid_for(M * (I / 8)), 256, 0, S(stream)>>>(BF(in), BF(out), M, I);
  }
  VFM_LAUNCH_CHECK("geglu");
  return VFM_OK;
}

int vfm_cast_f32_bf16(const float* in, void* out, long long n, void* stream) {
  if (!in || !out || n <= 0 || (n % 4)) return fail(VFM_ERR_INVALID, "cast_f32_bf16: bad args (n %% 4 == 0)");
  {
    LaunchScope scope("cast_f32_bf16", S(stream));
    cast_f32_bf16_kernel<<<grid_for(n / 4), 256, 0, S(stream)>>>(in, BF(out), n / 4);
  }
  VFM_LAUNCH_CHECK("cast_f32_bf16");
  return VFM_OK;
}

int vfm_ms_merge_argmax(const float* low0, const float* refined, const int* ref_index, const int* boxes, int n_crops, int nc,
                        int crop_h, int crop_w, int lh, int lw, int rh, int rw, int H, int W, int n_img, uint8_t* labels,
                        float* logits_out, void* stream) {
  if (!low0 || !ref_index || !boxes || !labels || n_crops <= 0 || nc <= 0 || nc > 32 || n_img <= 0)
    return fail(VFM_ERR_INVALID, "ms_merge_argmax: bad args (num_classes <= 32)");
  const unsigned grid = grid_for(static_cast<long long>(n_img) * H * W, 32);
  const size_t smem = sizeof(int2) * n_crops;
  // class-major tile kernel (slide_tail.cuh): both upsampling factors powers of two >= 4 (8 and 16 in the shipped config)
  auto log2_factor = [](int big_h, int big_w, int small_h, int small_w) {
    for (int lg = 2; lg < 16; ++lg)
      if ((small_h << lg) == big_h && (small_w << lg) == big_w) return lg;
    return -1;
  };
  const int lg0 = (lh > 0 && lw > 0) ? log2_factor(H, W, lh, lw) : -1;
  const int lgr = refined ? ((rh > 0 && rw > 0) ? log2_factor(crop_h, crop_w, rh, rw) : -1) : 2;
  const bool tiles = merge_mode() != 1 && nc <= 19 && lg0 >= 2 && lgr >= 2 && (W % 4) == 0 && n_img <= 65535 &&
                     (reinterpret_cast<uintptr_t>(labels) & 3) == 0 && (!logits_out || (reinterpret_cast<uintptr_t>(logits_out) & 15) == 0);
  {
    LaunchScope scope("ms_merge_argmax", S(stream));
    if (tiles) {
      const dim3 tgrid((W + MERGE_TW - 1) / MERGE_TW, (H + MERGE_TH - 1) / MERGE_TH, n_img);
      if (merge_mode() == 3)   // 3 resident CTAs per SM (80 registers, loop-invariant slot state partly spilled): 335 us against 309 us per 2 images
        ms_merge_class_kernel<19, 3><<<tgrid, 256, 0, S(stream)>>>(low0, refined, ref_index, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h,
                                                                  crop_w, lh, lw, rh, rw, H, W, lg0, lgr, labels, logits_out);
      else
        ms_merge_class_kernel<19, 2><<<tgrid, 256, 0, S(stream)>>>(low0, refined, ref_index, reinterpret_cast<const int2*>(boxes), n_crops, nc, crop_h,
                                                                  crop_w, lh, lw, rh, rw, H, W, lg0, lgr, labels, logits_out);
    } else if (nc <= 19)
      ms_merge_argmax_kernel<19><<<grid, 256, smem, S(stream)>>>(low0, refined, ref_index, reinterpret_cast<const int2*>(boxes), n_crops,
                                                                 nc, crop_h, crop_w, lh, lw, rh, rw, H, W, n_img, labels, logits_out);
    else
      ms_merge_argmax_kernel<32><<<grid, 256, smem, S(stream)>>>(low0, refined, ref_index, reinterpret_cast<const int2*>(boxes), n_crops,
                                                                 nc, crop_h, crop_w, lh, lw, rh, rw, H, W, n_img, labels, logits_out);
  }
  VFM_LAUNCH_CHECK("ms_merge_argmax");
  return VFM_OK;
}

// ------------------------------------------------------------------------------ EVA02 (config 4)
int vfm_rope_qk(void* qkv, long long M, int C, int heads, int tokens_per_seq, const float* cos_t, const float* sin_t, void* stream) {
  if (!qkv || !cos_t || !sin_t || M <= 0 || heads <= 0 || C != heads * 64 || tokens_per_seq < 2 || M > 0x7fffffffLL)
    return fail(VFM_ERR_INVALID, "rope_qk: bad args (head_dim 64)");
  {
    LaunchScope scope("rope_qk", S(stream));
    rope_qk_kernel<<<grid_for(M * 2 * heads * 8), 256, 0, S(stream)>>>(BF(qkv), M, C, heads, FastDiv(tokens_per_seq), cos_t, sin_t);
  }
  VFM_LAUNCH_CHECK("rope_qk");
  return VFM_OK;
}

int vfm_swiglu_layernorm(const void* in, void* out, const float* gamma, const float* beta, long long M, int H, int Hp, float eps,
                         void* stream) {
  if (!in || !out || !gamma || !beta || M <= 0 || H <= 0 || Hp < H || (Hp % 8) || Hp > SWIGLU_MAX_ITERS * SWIGLU_TPR * 8)
    return fail(VFM_ERR_INVALID, "swiglu_layernorm: bad args (Hp %% 8 == 0, Hp <= %d)", SWIGLU_MAX_ITERS * SWIGLU_TPR * 8);
  {
    LaunchScope scope("swiglu_layernorm", S(stream));
    swiglu_layernorm_kernel<<<static_cast<unsigned>((M + 1) / 2), 256, 0, S(stream)>>>(BF(in), BF(out), gamma, beta, M, H, Hp, eps);
  }
  VFM_LAUNCH_CHECK("swiglu_layernorm");
  return VFM_OK;
}

int vfm_confusion_matrix(const uint8_t* pred, const uint8_t* label, long long n, int nc, int ignore_index,
                         long long* cm, void* stream) {
  if (n == 0) return VFM_OK;
  if (!pred || !label || !cm || n < 0 || nc <= 0 || nc > 64) return fail(VFM_ERR_INVALID, "confusion_matrix: bad args");
  if ((reinterpret_cast<uintptr_t>(pred) & 15) || (reinterpret_cast<uintptr_t>(label) & 15))
    return fail(VFM_ERR_INVALID, "confusion_matrix: pred/label must be 16-byte aligned");
  long long blocks = ((n >> 4) + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
{
    LaunchScope scope("confusion_matrix", S(stream));
    confusion_matrix_kernel<<<static_cast<unsigned>(blocks), 256, sizeof(int) * (nc + 1) * nc, S(stream)>>>(
        pred, label, n, nc, ignore_index, reinterpret_cast<unsigned long long*>(cm));
  }
  VFM_LAUNCH_CHECK("confusion_matrix");
  return VFM_OK;
}

// ------------------------------------------------------------------------------ fused drivers
// ViT workspace layout (all 256-byte aligned):
//   x     fp32 [M, C]        residual stream          M = n_crops * (gh*gw + 1)
//   xn    bf16 [M, C]        LayerNorm output / attention output (disjoint lifetimes -> two buffers)
//   att   bf16 [M, C]
//   qkv   bf16 [M, 3C]
//   hid   bf16 [M, hidden]   (the patch operand [n_crops*gh*gw, 768] aliases this buffer)
size_t vfm_vit_workspace_bytes(const VfmVitParams* p, int n_crops, int gh, int gw) {
  if (!p) return 0;
  const size_t M = static_cast<size_t>(n_crops) * (static_cast<size_t>(gh) * gw + 1);
  const size_t C = p->embed_dim;
  size_t hid = M * p->mlp_hidden * 2;
  const size_t patches = static_cast<size_t>(n_crops) * gh * gw * 768 * 2;
  if (patches > hid) hid = patches;
  return align_up(M * C * 4, 256) + 2 * align_up(M * C * 2, 256) + align_up(M * 3 * C * 2, 256) + align_up(hid, 256) +
         align_up(M * (C / 128 + 1) * 8, 256);   // LayerNorm statistics: (sum, sum of squares) per row and 128-column slot
}

int vfm_vit_forward(const VfmVitParams* p, const void* img, int is_u8, const VfmPixelNorm* nrm, int img_h, int img_w,
                    const int* crops, int n_crops, int gh, int gw, void* taps, void* workspace, size_t workspace_bytes,
                    void* stream) {
  if (!p || !p->blocks || !img || !crops || !taps || !workspace) return fail(VFM_ERR_INVALID, "vit_forward: null argument");
  const int C = p->embed_dim, Hd = p->mlp_hidden;
  if (C != p->heads * 64) return fail(VFM_ERR_INVALID, "vit_forward: head_dim must be 64 (embed_dim=%d heads=%d)", C, p->heads);
  if (p->n_taps <= 0 || p->n_taps > 8) return fail(VFM_ERR_INVALID, "vit_forward: n_taps out of range");
  const size_t need = vfm_vit_workspace_bytes(p, n_crops, gh, gw);
  if (workspace_bytes < need) return fail(VFM_ERR_WORKSPACE, "vit_forward: workspace %zu < %zu", workspace_bytes, need);
  const int P = gh * gw, T = P + 1;
  const int M = n_crops * T;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws);            ws += align_up(static_cast<size_t>(M) * C * 4, 256);
  void* xn = ws;                                      ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* att = ws;                                     ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* qkv = ws;                                     ws += align_up(static_cast<size_t>(M) * 3 * C * 2, 256);
  void* hid = ws;
  {
    size_t hb = static_cast<size_t>(M) * Hd * 2;
    const size_t pb = static_cast<size_t>(n_crops) * P * 768 * 2;
    if (pb > hb) hb = pb;
    ws += align_up(hb, 256);
  }
  float* stats = reinterpret_cast<float*>(ws);
  // LayerNorm folding (EpiTmaResidualStats -> EpiTmaBf16LN): a residual GEMM whose successor is a plain LayerNorm + Linear
  // also writes bf16(x) into xn and the row statistics; the successor runs on the folded weights. Not folded: norm1 of
  // block 0 (x comes from the patch embedding) and the norm1 passes that also emit a feature tap.
  static const int fold_env = [] { const char* e = getenv("VFM_LN_FOLD"); return e ? atoi(e) : 3; }();   // bit 0: norm1 (fc2 -> qkv), bit 1: norm2 (proj -> fc1)
  const bool fold_ok1 = (fold_env & 1) && (C % 256) == 0, fold_ok2 = (fold_env & 2) && (C % 256) == 0;
  int rc;
  // tokens: patch gather -> patch-embed GEMM (+bias +pos) ; cls rows
  if ((rc = vfm_patch_gather(img, is_u8, nrm, img_h, img_w, crops, n_crops, gh, gw, hid, stream))) return rc;
  if ((rc = vfm_gemm_patch_embed(hid, 768, p->patch_w, 768, p->patch_b, p->pos_embed, x, P, n_crops * P, C, 768, stream))) return rc;
  if ((rc = vfm_cls_rows(x, p->cls_token, p->pos_embed, n_crops, T, C, stream))) return rc;
  // Feature taps (residual stream after block tap_blocks[i]) are emitted by the NEXT LayerNorm pass, which reads x
  // anyway, so all 48 residual GEMMs use the TMA reduce-add epilogue; a tap after the last block gets a tap-only pass.
  int next_tap = 0;
  for (int l = 0; l < p->depth; ++l) {
    const VfmBlockParams& b = p->blocks[l];
    void* tap = nullptr;
    int tap_col0 = 0;
    if (l > 0 && next_tap < p->n_taps && p->tap_blocks[next_tap] == l - 1) {
      tap = taps;
      tap_col0 = next_tap * C;
      ++next_tap;
    }
    // norm1 + qkv
    const bool fold1 = fold_ok1 && l > 0 && tap == nullptr && b.qkv_wf && b.qkv_bf && b.qkv_cs;
    if (fold1) {
      if ((rc = vfm_gemm_lnfold_bf16(xn, C, b.qkv_wf, C, b.qkv_bf, b.qkv_cs, stats, p->ln_eps, 0, qkv, 3 * C, M, 3 * C, C, stream))) return rc;
    } else {
      if ((rc = vfm_layernorm_tap(x, b.ln1_w, b.ln1_b, xn, M, C, p->ln_eps, tap, p->n_taps * C, tap_col0, T, stream))) return rc;
      if ((rc = vfm_gemm_bias_bf16(xn, C, b.qkv_w, C, b.qkv_b, qkv, 3 * C, M, 3 * C, C, stream))) return rc;
    }
    if ((rc = vfm_attention_fwd(qkv, att, n_crops, T, p->heads, stream))) return rc;
    // proj (+ residual) -> norm2 + fc1 (+ GELU)
    const bool fold2 = fold_ok2 && b.fc1_wf && b.fc1_bf && b.fc1_cs;
    if (fold2) {
      if ((rc = vfm_gemm_bias_ls_residual_stats(att, C, b.proj_w, C, b.proj_b, b.ls1, x, C, xn, C, stats, M, C, C, stream))) return rc;
      if ((rc = vfm_gemm_lnfold_bf16(xn, C, b.fc1_wf, C, b.fc1_bf, b.fc1_cs, stats, p->ln_eps, 1, hid, Hd, M, Hd, C, stream))) return rc;
    } else {
      if ((rc = vfm_gemm_bias_ls_residual(att, C, b.proj_w, C, b.proj_b, b.ls1, x, C, nullptr, 0, 0, T, M, C, C, stream))) return rc;
      if ((rc = vfm_layernorm(x, b.ln2_w, b.ln2_b, xn, M, C, p->ln_eps, stream))) return rc;
      if ((rc = vfm_gemm_bias_gelu_bf16(xn, C, b.fc1_w, C, b.fc1_b, hid, Hd, M, Hd, C, stream))) return rc;
    }
    // fc2 (+ residual); it prepares the next block's norm1 when that one is folded
    bool next_fold = false;
    if (fold_ok1 && l + 1 < p->depth) {
      const VfmBlockParams& nb = p->blocks[l + 1];
      const bool next_tap_here = next_tap < p->n_taps && p->tap_blocks[next_tap] == l;
      next_fold = !next_tap_here && nb.qkv_wf && nb.qkv_bf && nb.qkv_cs;
    }
    if (next_fold) {
      if ((rc = vfm_gemm_bias_ls_residual_stats(hid, Hd, b.fc2_w, Hd, b.fc2_b, b.ls2, x, C, xn, C, stats, M, C, Hd, stream))) return rc;
    } else {
      if ((rc = vfm_gemm_bias_ls_residual(hid, Hd, b.fc2_w, Hd, b.fc2_b, b.ls2, x, C, nullptr, 0, 0, T, M, C, Hd, stream))) return rc;
    }
  }
  if (next_tap < p->n_taps && p->tap_blocks[next_tap] == p->depth - 1) {
    if ((rc = vfm_layernorm_tap(x, nullptr, nullptr, nullptr, M, C, p->ln_eps, taps, p->n_taps * C, next_tap * C, T, stream))) return rc;
    ++next_tap;
  }
  if (next_tap != p->n_taps) return fail(VFM_ERR_INVALID, "vit_forward: tap_blocks must be ascending and < depth");
  return VFM_OK;
}

// EVA02 workspace: x fp32 [M, C], xn bf16 [M, C], att bf16 [M, C], qkv bf16 [M, 3C], h12 bf16 [M, 2 Hp] (also the gathered
// patches [n P, 768]), u bf16 [M, Hp], LayerNorm statistics; M = n_crops * (grid^2 + 1).
size_t vfm_eva_workspace_bytes(const VfmEvaParams* p, int n_crops) {
  if (!p) return 0;
  const size_t P = static_cast<size_t>(p->grid) * p->grid, M = static_cast<size_t>(n_crops) * (P + 1);
  const size_t C = p->embed_dim, Hp = p->hidden_pad;
  size_t h12 = M * 2 * Hp * 2;
  const size_t patches = static_cast<size_t>(n_crops) * P * 768 * 2;
  if (patches > h12) h12 = patches;
  return align_up(M * C * 4, 256) + 2 * align_up(M * C * 2, 256) + align_up(M * 3 * C * 2, 256) + align_up(h12, 256) +
         align_up(M * Hp * 2, 256) + align_up(M * (C / 128 + 1) * 8, 256);
}

int vfm_eva_forward(const VfmEvaParams* p, const void* img, int is_u8, const VfmPixelNorm* nrm, int img_h, int img_w,
                    const int* crops, int n_crops, void* taps, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !p->blocks || !img || !crops || !taps || !workspace) return fail(VFM_ERR_INVALID, "eva_forward: null argument");
  const int C = p->embed_dim, H = p->hidden, Hp = p->hidden_pad, g = p->grid;
  if (C != p->heads * 64) return fail(VFM_ERR_INVALID, "eva_forward: head_dim must be 64 (embed_dim=%d heads=%d)", C, p->heads);
  if (p->n_taps <= 0 || p->n_taps > 8) return fail(VFM_ERR_INVALID, "eva_forward: n_taps out of range");
  if (Hp < H || (Hp % 8)) return fail(VFM_ERR_INVALID, "eva_forward: hidden_pad must be a multiple of 8 and >= hidden");
  if (n_crops <= 0 || g <= 0) return fail(VFM_ERR_INVALID, "eva_forward: bad n_crops / grid");
  const size_t need = vfm_eva_workspace_bytes(p, n_crops);
  if (workspace_bytes < need) return fail(VFM_ERR_WORKSPACE, "eva_forward: workspace %zu < %zu", workspace_bytes, need);
  const int P = g * g, T = P + 1;
  const int M = n_crops * T;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws);            ws += align_up(static_cast<size_t>(M) * C * 4, 256);
  void* xn = ws;                                      ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* att = ws;                                     ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* qkv = ws;                                     ws += align_up(static_cast<size_t>(M) * 3 * C * 2, 256);
  void* h12 = ws;
  {
    size_t hb = static_cast<size_t>(M) * 2 * Hp * 2;
    const size_t pb = static_cast<size_t>(n_crops) * P * 768 * 2;
    if (pb > hb) hb = pb;
    ws += align_up(hb, 256);
  }
  void* u = ws;                                       ws += align_up(static_cast<size_t>(M) * Hp * 2, 256);
  float* stats = reinterpret_cast<float*>(ws);
  // VFM_LN_FOLD as in vfm_vit_forward; read per call so that a test can compare the folded and the plain sequence in one process
  int fold_env = 3;
  if (const char* e = getenv("VFM_LN_FOLD")) fold_env = atoi(e);
  const bool have_fold = (C % 256) == 0 && p->blocks[0].qkv_wf != nullptr;
  const bool fold_ok1 = have_fold && (fold_env & 1), fold_ok2 = have_fold && (fold_env & 2);
  auto is_tap = [&](int block) {
    for (int t = 0; t < p->n_taps; ++t)
      if (p->tap_blocks[t] == block) return t;
    return -1;
  };
  int rc;
  if ((rc = vfm_patch_gather(img, is_u8, nrm, img_h, img_w, crops, n_crops, g, g, h12, stream))) return rc;
  if ((rc = vfm_gemm_patch_embed(h12, 768, p->patch_w, 768, p->patch_b, p->pos_embed, x, P, n_crops * P, C, 768, stream))) return rc;
  if ((rc = vfm_cls_rows(x, p->cls_token, p->pos_embed, n_crops, T, C, stream))) return rc;
  bool pre = false;   // xn = bf16(x) and stats are current (written by the previous block's w3 GEMM)
  for (int l = 0; l < p->depth; ++l) {
    const VfmEvaBlockParams& b = p->blocks[l];
    const int tap_i = l > 0 ? is_tap(l - 1) : -1;   // tap of the previous block's output, emitted by this block's norm1 pass
    if (pre) {
      if ((rc = vfm_gemm_lnfold_rope_bf16(xn, C, b.qkv_wf, C, b.qkv_bf, b.qkv_cs, stats, p->ln_eps, qkv, 3 * C, M, 3 * C, C,
                                          p->rope_cos, p->rope_sin, 2 * C, T, stream))) return rc;
    } else {
      if ((rc = vfm_layernorm_tap(x, b.ln1_w, b.ln1_b, xn, M, C, p->ln_eps, tap_i >= 0 ? taps : nullptr, tap_i >= 0 ? p->n_taps * C : 0,
                                  tap_i >= 0 ? tap_i * C : 0, T, stream))) return rc;
      if ((rc = vfm_gemm_bias_rope_bf16(xn, C, b.qkv_w, C, b.qkv_b, qkv, 3 * C, M, 3 * C, C, p->rope_cos, p->rope_sin, 2 * C, T, stream))) return rc;
    }
    if ((rc = vfm_attention_fwd_ex(qkv, att, n_crops, T, p->heads, 0, stream))) return rc;
    if (fold_ok2 && b.w12_wf) {
      if ((rc = vfm_gemm_bias_ls_residual_stats(att, C, b.proj_w, C, b.proj_b, p->ones, x, C, xn, C, stats, M, C, C, stream))) return rc;
      if ((rc = vfm_gemm_lnfold_bf16(xn, C, b.w12_wf, C, b.w12_bf, b.w12_cs, stats, p->ln_eps, 0, h12, 2 * Hp, M, 2 * Hp, C, stream))) return rc;
    } else {
      if ((rc = vfm_gemm_bias_ls_residual(att, C, b.proj_w, C, b.proj_b, p->ones, x, C, nullptr, 0, 0, 1, M, C, C, stream))) return rc;
      if ((rc = vfm_layernorm(x, b.ln2_w, b.ln2_b, xn, M, C, p->ln_eps, stream))) return rc;
      if ((rc = vfm_gemm_bias_bf16(xn, C, b.w12, C, b.b12, h12, 2 * Hp, M, 2 * Hp, C, stream))) return rc;
    }
    if ((rc = vfm_swiglu_layernorm(h12, u, b.ffn_ln_w, b.ffn_ln_b, M, H, Hp, p->ln_eps, stream))) return rc;
    pre = fold_ok1 && l + 1 < p->depth && is_tap(l) < 0 && p->blocks[l + 1].qkv_wf != nullptr;
    if (pre) {
      if ((rc = vfm_gemm_bias_ls_residual_stats(u, Hp, b.w3, Hp, b.b3, p->ones, x, C, xn, C, stats, M, C, Hp, stream))) return rc;
    } else {
      if ((rc = vfm_gemm_bias_ls_residual(u, Hp, b.w3, Hp, b.b3, p->ones, x, C, nullptr, 0, 0, 1, M, C, Hp, stream))) return rc;
    }
  }
  const int last = is_tap(p->depth - 1);
  if (last >= 0) {
    if ((rc = vfm_layernorm_tap(x, nullptr, nullptr, nullptr, M, C, p->ln_eps, taps, p->n_taps * C, last * C, T, stream))) return rc;
  }
  return VFM_OK;
}

// SAM ViT workspace: x fp32 [M, C], xn bf16 [M, C], att bf16 [M, C], att_w bf16 [win_rows, C] (windowed blocks off the tcgen05
// path only), qkv bf16 [max(M, win_rows), max qkv_n], hid bf16 [M, hidden] (also the gathered patches), statistics; M = n P.
static size_t sam_qkv_elems(const VfmSamParams* p, size_t M) {
  size_t best = 0;
  for (int l = 0; l < p->depth; ++l) {
    const size_t rows = p->blocks[l].window ? static_cast<size_t>(p->win_rows) : M;
    const size_t e = rows * static_cast<size_t>(p->blocks[l].qkv_n);
    if (e > best) best = e;
  }
  return best;
}
size_t vfm_sam_workspace_bytes(const VfmSamParams* p, int n_crops) {
  if (!p || !p->blocks) return 0;
  const size_t P = static_cast<size_t>(p->grid) * p->grid, M = static_cast<size_t>(n_crops) * P, C = p->embed_dim;
  size_t hid = M * p->hidden * 2;
  if (M * 768 * 2 > hid) hid = M * 768 * 2;
  return align_up(M * C * 4, 256) + 2 * align_up(M * C * 2, 256) + align_up(static_cast<size_t>(p->win_rows) * C * 2, 256) +
         align_up(sam_qkv_elems(p, M) * 2, 256) + align_up(hid, 256) + align_up(M * (C / 128 + 1) * 8, 256);
}

int vfm_sam_forward(const VfmSamParams* p, const void* img, int is_u8, const VfmPixelNorm* nrm, int img_h, int img_w,
                    const int* crops, int n_crops, void* taps, void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !p->blocks || !img || !crops || !taps || !workspace) return fail(VFM_ERR_INVALID, "sam_forward: null argument");
  const int C = p->embed_dim, Hd = p->hidden, g = p->grid, H = p->heads, d = p->head_dim;
  if (C != H * d || (d != 64 && d != 80)) return fail(VFM_ERR_INVALID, "sam_forward: head_dim must be 64 or 80 (embed_dim=%d heads=%d)", C, H);
  if (p->n_taps <= 0 || p->n_taps > 8) return fail(VFM_ERR_INVALID, "sam_forward: n_taps out of range");
  if (n_crops <= 0 || g <= 0) return fail(VFM_ERR_INVALID, "sam_forward: bad n_crops / grid");
  const size_t need = vfm_sam_workspace_bytes(p, n_crops);
  if (workspace_bytes < need) return fail(VFM_ERR_WORKSPACE, "sam_forward: workspace %zu < %zu", workspace_bytes, need);
  const int P = g * g, M = n_crops * P;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  float* x = reinterpret_cast<float*>(ws);            ws += align_up(static_cast<size_t>(M) * C * 4, 256);
  void* xn = ws;                                      ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* att = ws;                                     ws += align_up(static_cast<size_t>(M) * C * 2, 256);
  void* att_w = ws;                                   ws += align_up(static_cast<size_t>(p->win_rows) * C * 2, 256);
  void* qkv = ws;                                     ws += align_up(sam_qkv_elems(p, M) * 2, 256);
  void* hid = ws;
  {
    size_t hb = static_cast<size_t>(M) * Hd * 2;
    if (static_cast<size_t>(M) * 768 * 2 > hb) hb = static_cast<size_t>(M) * 768 * 2;
    ws += align_up(hb, 256);
  }
  float* stats = reinterpret_cast<float*>(ws);
  int fold_env = 3;
  if (const char* e = getenv("VFM_LN_FOLD")) fold_env = atoi(e);
  const float scale = static_cast<float>(pow(static_cast<double>(d), -0.5));   // head_dim ** -0.5 as the Python driver computed it
  const int g0 = p->use_rel_pos ? 3 * C : -1;   // first table-term column of the extended qkv rows
  auto is_tap = [&](int block) {
    for (int t = 0; t < p->n_taps; ++t)
      if (p->tap_blocks[t] == block) return t;
    return -1;
  };
  int rc;
  if ((rc = vfm_patch_gather(img, is_u8, nrm, img_h, img_w, crops, n_crops, g, g, hid, stream))) return rc;
  if ((rc = vfm_gemm_patch_embed_ex(hid, 768, p->patch_w, 768, p->patch_b, p->pos_embed, x, P, 0, M, C, 768, stream))) return rc;
  for (int l = 0; l < p->depth; ++l) {
    const VfmSamBlockParams& b = p->blocks[l];
    const int tap_i = l > 0 ? is_tap(l - 1) : -1;   // tap of the previous block's output, emitted by this block's norm1 pass
    void* tap = tap_i >= 0 ? taps : nullptr;
    const int tap_ld = tap_i >= 0 ? p->n_taps * C : 0, tap_col0 = tap_i >= 0 ? tap_i * C : 0;
    const int wsz = b.window, Nq = b.qkv_n;
    if (wsz) {
      // window_partition folded into the LayerNorm store, window_unpartition into the attention kernel's store
      if (!p->part || !p->unpart || !p->win_buf || p->win_rows % (wsz * wsz))
        return fail(VFM_ERR_INVALID, "sam_forward: windowed block %d needs part / unpart / win_buf (win_rows %% window^2 == 0)", l);
      const int n_win = p->win_rows / (wsz * wsz);
      if ((rc = vfm_layernorm_tap_map(x, b.ln1_w, b.ln1_b, p->win_buf, M, C, p->ln_eps, tap, tap_ld, tap_col0, 1, 0, p->unpart, stream))) return rc;
      if ((rc = vfm_gemm_bias_bf16(p->win_buf, C, b.qkv_w, C, b.qkv_b, qkv, Nq, p->win_rows, Nq, C, stream))) return rc;
      if (d == 80 && wsz * wsz <= 208 && wsz <= 16) {   // whole window in one tcgen05 score tile
        if ((rc = vfm_attention_window_tc_map(qkv, Nq, g0, att, p->part, n_win, wsz * wsz, H, d, wsz, wsz, scale, stream))) return rc;
      } else {
        if ((rc = vfm_attention_relpos_ex(qkv, Nq, g0, nullptr, att_w, n_win, wsz * wsz, H, d, wsz, wsz, scale, stream))) return rc;
        if ((rc = vfm_rows_gather(att_w, att, p->unpart, M, C, stream))) return rc;
      }
    } else {
      if ((rc = vfm_layernorm_tap_ex(x, b.ln1_w, b.ln1_b, xn, M, C, p->ln_eps, tap, tap_ld, tap_col0, 1, 0, stream))) return rc;
      if ((rc = vfm_gemm_bias_bf16(xn, C, b.qkv_w, C, b.qkv_b, qkv, Nq, M, Nq, C, stream))) return rc;
      const int bias_cols = (g + 15) / 16 * 16 * 2;
      if (d == 80 && g0 >= 0 && bias_cols <= 128 && p->onehot) {   // key-tile loop on tcgen05
        if ((rc = vfm_attention_global_tc(qkv, Nq, g0, p->onehot, p->onehot_rows, att, n_crops, P, H, d, g, g, scale, stream))) return rc;
      } else {
        if ((rc = vfm_attention_relpos_ex(qkv, Nq, g0, nullptr, att, n_crops, P, H, d, g, g, scale, stream))) return rc;
      }
    }
    if ((fold_env & 2) && b.lin1_wf && (C % 256) == 0) {
      // the proj GEMM also emits bf16(x) and the row statistics; lin1 applies norm2 in its epilogue (no LayerNorm pass)
      if ((rc = vfm_gemm_bias_ls_residual_stats(att, C, b.proj_w, C, b.proj_b, p->ones, x, C, xn, C, stats, M, C, C, stream))) return rc;
      if ((rc = vfm_gemm_lnfold_bf16(xn, C, b.lin1_wf, C, b.lin1_bf, b.lin1_cs, stats, p->ln_eps, 1, hid, Hd, M, Hd, C, stream))) return rc;
    } else {
      if ((rc = vfm_gemm_bias_ls_residual(att, C, b.proj_w, C, b.proj_b, p->ones, x, C, nullptr, 0, 0, 1, M, C, C, stream))) return rc;
      if ((rc = vfm_layernorm(x, b.ln2_w, b.ln2_b, xn, M, C, p->ln_eps, stream))) return rc;
      if ((rc = vfm_gemm_bias_gelu_bf16(xn, C, b.lin1_w, C, b.lin1_b, hid, Hd, M, Hd, C, stream))) return rc;
    }
    if ((rc = vfm_gemm_bias_ls_residual(hid, Hd, b.lin2_w, Hd, b.lin2_b, p->ones, x, C, nullptr, 0, 0, 1, M, C, Hd, stream))) return rc;
  }
  const int last = is_tap(p->depth - 1);
  if (last >= 0) {
    if ((rc = vfm_layernorm_tap_ex(x, nullptr, nullptr, nullptr, M, C, p->ln_eps, taps, p->n_taps * C, last * C, 1, 0, stream))) return rc;
  }
  return VFM_OK;
}

// LinearHead workspace: f0 bf16 [R, mid] (fusion out), f1 bf16 [R, mid] (GN+ReLU), u1 bf16 [4R, mid/2],
// u2 bf16 [16R, mid/4]; R = n_crops * gh * gw.
size_t vfm_linear_head_workspace_bytes(const VfmLinearHeadParams* p, int n_crops, int gh, int gw) {
  if (!p) return 0;
  const size_t R = static_cast<size_t>(n_crops) * gh * gw, mid = p->mid_channels;
  return 2 * align_up(R * mid * 2, 256) + align_up(4 * R * (mid / 2) * 2, 256) + align_up(16 * R * (mid / 4) * 2, 256);
}

int vfm_linear_head_forward(const VfmLinearHeadParams* p, const void* taps, int n_crops, int gh, int gw, float* lowres,
                            void* workspace, size_t workspace_bytes, void* stream) {
  if (!p || !taps || !lowres || !workspace) return fail(VFM_ERR_INVALID, "linear_head_forward: null argument");
  const int mid = p->mid_channels;
  if (mid % 128) return fail(VFM_ERR_INVALID, "linear_head_forward: mid_channels must be a multiple of 128");
  const size_t need = vfm_linear_head_workspace_bytes(p, n_crops, gh, gw);
  if (workspace_bytes < need) return fail(VFM_ERR_WORKSPACE, "linear_head_forward: workspace %zu < %zu", workspace_bytes, need);
  const int P = gh * gw;
  const int R = n_crops * P;
  uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
  void* f0 = ws; ws += align_up(static_cast<size_t>(R) * mid * 2, 256);
  void* f1 = ws; ws += align_up(static_cast<size_t>(R) * mid * 2, 256);
  void* u1 = ws; ws += align_up(static_cast<size_t>(4) * R * (mid / 2) * 2, 256);
  void* u2 = ws;
  int rc;
  if ((rc = vfm_gemm_bias_bf16(taps, p->in_channels, p->fusion_w, p->in_channels, nullptr, f0, mid, R, mid, p->in_channels, stream))) return rc;
  if ((mid / p->groups) % 8 == 0) {
    if ((rc = vfm_groupnorm_relu(f0, f1, p->gn_w, p->gn_b, n_crops, P, mid, p->groups, p->gn_eps, 1, stream))) return rc;
  } else {   // any channels-per-group (e.g. 640 / 32 = 20)
    if ((rc = vfm_groupnorm_act(f0, f1, 0, p->gn_w, p->gn_b, n_crops, P, mid, p->groups, p->gn_eps, 1, stream))) return rc;
  }
  if ((rc = vfm_gemm_convt2x2_gelu(f1, mid, p->up1_w, mid, p->up1_b, u1, mid / 2, gh, gw, R, mid, stream))) return rc;
  if ((rc = vfm_gemm_convt2x2_gelu(u1, mid / 2, p->up2_w, mid / 2, p->up2_b, u2, mid / 4, 2 * gh, 2 * gw, 4 * R, mid / 2, stream))) return rc;
  if ((rc = vfm_gemm_cls_nchw(u2, mid / 4, p->cls_w, mid / 4, p->cls_b, lowres, p->num_classes, 16 * P, 16 * R, mid / 4, stream))) return rc;
  return VFM_OK;
}

}  // extern "C"
