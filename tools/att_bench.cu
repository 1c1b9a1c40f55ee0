// Stand-alone correctness + timing harness for the attention entry point of libvfmseg_b200.so (no torch: starts in
// milliseconds on a fresh GPU box).  Build:  nvcc -O2 -o tools/bin/att_bench tools/att_bench.cu -ldl
// usage: att_bench <lib.so> <n_seq> <seq_len> <heads> <mode,mode,...> [iters] [scale] [check_pairs]
// For every mode: max |err| against an fp32 CPU softmax attention on `check_pairs` (sequence, head) pairs (all rows),
// median kernel time over `iters` launches with an L2 flush in between, TFLOP/s on 4 * heads * seq^2 * 64 per sequence.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

typedef int (*att_fn)(const void*, void*, int, int, int, int, void*);
typedef const char* (*err_fn)(void);

static float bf2f(__nv_bfloat16 v) { return __bfloat162float(v); }

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: %s lib n_seq seq_len heads modes [iters] [scale] [check_pairs]\n", argv[0]); return 2; }
  void* h = dlopen(argv[1], RTLD_NOW);
  if (!h) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  att_fn att = (att_fn)dlsym(h, "vfm_attention_fwd_ex");
  err_fn last_err = (err_fn)dlsym(h, "vfm_last_error");
  if (!att || !last_err) { fprintf(stderr, "missing symbols\n"); return 2; }
  const int n_seq = atoi(argv[2]), S = atoi(argv[3]), heads = atoi(argv[4]);
  std::vector<int> modes;
  for (char* tok = strtok(argv[5], ","); tok; tok = strtok(nullptr, ",")) modes.push_back(atoi(tok));
  const int iters = argc > 6 ? atoi(argv[6]) : 10;
  const float scale = argc > 7 ? atof(argv[7]) : 0.7f;
  const int check_pairs = argc > 8 ? atoi(argv[8]) : 2;
  const int C = heads * 64;
  const size_t rows = (size_t)n_seq * S, n_in = rows * 3 * C, n_out = rows * C;

  std::vector<__nv_bfloat16> hq(n_in);
  uint64_t st = 0x9E3779B97F4A7C15ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (size_t i = 0; i < n_in; i += 2) {   // Box-Muller
    const double u = std::max(rnd(), 1e-12), v = rnd();
    const double r = std::sqrt(-2.0 * std::log(u));
    hq[i] = __float2bfloat16((float)(r * std::cos(6.283185307179586 * v)) * scale);
    if (i + 1 < n_in) hq[i + 1] = __float2bfloat16((float)(r * std::sin(6.283185307179586 * v)) * scale);
  }
  // the q third carries the 1/sqrt(d) scale, as the library expects (folded into W_q on the host)
  for (size_t r = 0; r < rows; ++r)
    for (int c = 0; c < C; ++c) hq[r * 3 * C + c] = __float2bfloat16(bf2f(hq[r * 3 * C + c]) * 0.125f * 4.f);

  // CPU reference on the first / last (sequence, head) pairs
  struct Pair { int seq, head; };
  std::vector<Pair> pairs;
  for (int i = 0; i < check_pairs; ++i) {
    Pair p{(i % 2 == 0) ? (i / 2) % n_seq : n_seq - 1 - (i / 2) % n_seq, (i % 2 == 0) ? (i / 2) % heads : heads - 1 - (i / 2) % heads};
    pairs.push_back(p);
  }
  std::vector<std::vector<float>> refs;
  for (auto& pr : pairs) {
    std::vector<float> ref((size_t)S * 64), q((size_t)S * 64), k((size_t)S * 64), v((size_t)S * 64), sc(S);
    for (int t = 0; t < S; ++t)
      for (int d = 0; d < 64; ++d) {
        const size_t base = ((size_t)pr.seq * S + t) * 3 * C + pr.head * 64 + d;
        q[t * 64 + d] = bf2f(hq[base]); k[t * 64 + d] = bf2f(hq[base + C]); v[t * 64 + d] = bf2f(hq[base + 2 * C]);
      }
    for (int i = 0; i < S; ++i) {
      float m = -INFINITY;
      for (int j = 0; j < S; ++j) {
        float a = 0.f;
        for (int d = 0; d < 64; ++d) a += q[i * 64 + d] * k[j * 64 + d];
        sc[j] = a; m = std::max(m, a);
      }
      double l = 0.0;
      for (int j = 0; j < S; ++j) { sc[j] = std::exp(sc[j] - m); l += sc[j]; }
      for (int d = 0; d < 64; ++d) {
        double o = 0.0;
        for (int j = 0; j < S; ++j) o += (double)sc[j] * v[j * 64 + d];
        ref[i * 64 + d] = (float)(o / l);
      }
    }
    refs.push_back(std::move(ref));
  }

  __nv_bfloat16 *dq, *dout;
  uint8_t* flush;
  const size_t flush_bytes = 256u << 20;
  cudaMalloc(&dq, n_in * 2); cudaMalloc(&dout, n_out * 2); cudaMalloc(&flush, flush_bytes);
  cudaMemcpy(dq, hq.data(), n_in * 2, cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<__nv_bfloat16> ho(n_out);
  const double flop = 4.0 * n_seq * heads * (double)S * S * 64;
  int bad = 0;
  for (int mode : modes) {
    cudaMemset(dout, 0xff, n_out * 2);   // NaN pattern: rows that are never written show up
    int rc = att(dq, dout, n_seq, S, heads, mode, nullptr);
    cudaError_t ce = cudaDeviceSynchronize();
    if (rc != 0 || ce != cudaSuccess) {
      printf("mode %d: FAILED rc=%d (%s) cuda=%s\n", mode, rc, last_err(), cudaGetErrorString(ce));
      if (ce != cudaSuccess) return 1;
      bad = 1;
      continue;
    }
    cudaMemcpy(ho.data(), dout, n_out * 2, cudaMemcpyDeviceToHost);
    double max_err = 0.0, max_ref = 0.0; long nan_ct = 0;
    for (size_t i = 0; i < n_out; ++i) if (std::isnan(bf2f(ho[i]))) ++nan_ct;
    int worst_row = -1;
    for (size_t pi = 0; pi < pairs.size(); ++pi)
      for (int t = 0; t < S; ++t)
        for (int d = 0; d < 64; ++d) {
          const float g = bf2f(ho[((size_t)pairs[pi].seq * S + t) * C + pairs[pi].head * 64 + d]);
          const float r = refs[pi][t * 64 + d];
          const double e = std::fabs((double)g - r);
          if (!(e <= max_err)) { max_err = e; worst_row = t; }
          max_ref = std::max(max_ref, (double)std::fabs(r));
        }
    std::vector<float> ts;
    for (int it = 0; it < iters + 2; ++it) {
      cudaMemsetAsync(flush, it, flush_bytes, nullptr);
      cudaEventRecord(e0, nullptr);
      att(dq, dout, n_seq, S, heads, mode, nullptr);
      cudaEventRecord(e1, nullptr);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it >= 2) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    const float med = ts.empty() ? 0.f : ts[ts.size() / 2];
    const bool ok = nan_ct == 0 && max_err <= 2e-2 * std::max(max_ref, 1e-3) + 2e-3;
    if (!ok) bad = 1;
    printf("mode %d: %s max_err %.3e (ref max %.3f, worst row %d) nan %ld | median %.4f ms min %.4f  %.1f TFLOP/s\n", mode,
           ok ? "OK " : "BAD", max_err, max_ref, worst_row, nan_ct, med, ts.empty() ? 0.f : ts[0], med > 0 ? flop / med / 1e9 : 0.0);
    fflush(stdout);
  }
  // debug builds (-DVFM_APP_TRACE): dump the event timeline of CTA 0 of the last ping-pong launch
  typedef int (*trace_fn)(long long*);
  trace_fn tr = (trace_fn)dlsym(h, "vfm_debug_app_trace");
  if (tr) {
    std::vector<long long> buf(2052);
    tr(buf.data());
    const long long t0 = buf[2048];
    printf("trace: kernel clk %lld ns %lld -> %.3f GHz\n", buf[2050] - buf[2048], buf[2051] - buf[2049],
           (double)(buf[2050] - buf[2048]) / (double)std::max(1LL, buf[2051] - buf[2049]));
    const char* names[4] = {"softmaxA", "softmaxB", "S-issuer", "PV-issuer"};
    for (int w = 0; w < 4; ++w) {
      printf("%s\n", names[w]);
      for (int t = 0; t < 40; ++t) {
        printf("  t%2d", t);
        for (int e = 0; e < 8; ++e) printf(" %7lld", buf[(w * 64 + t) * 8 + e] ? buf[(w * 64 + t) * 8 + e] - t0 : -1LL);
        printf("\n");
      }
    }
  }
  return bad;
}
