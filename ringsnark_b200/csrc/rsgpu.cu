// librsgpu.so -- C ABI (include/rsgpu.h) over the sm_100a kernels.  Host side: context (NTT tables, Barrett
// constants, batch-encoder index map), device arenas, launch logic.  No CPU fallback anywhere: every entry point
// that computes runs CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rsgpu.h"
#include "host_poly.hpp"
#include "kernels.cuh"
#include "witness.cuh"
#include "witness_fast.cuh"
#include "instance.cuh"
#include "decode.cuh"
#include "ringops.cuh"
#include "prover_fast.cuh"
#include "encode.cuh"
#include "p2p.cuh"

using namespace rsg;
typedef unsigned __int128 u128;

// ------------------------------------------------------------------------------------------------------------
// errors
static thread_local std::string g_err;
static int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return fail(RSG_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));            \
  } while (0)

extern "C" const char *rsg_last_error(void) { return g_err.c_str(); }
extern "C" int rsg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

// ------------------------------------------------------------------------------------------------------------
// host number theory (context creation only)
static uint64_t h_mulmod(uint64_t a, uint64_t b, uint64_t p) { return (uint64_t)(((u128)a * b) % p); }
static uint64_t h_powmod(uint64_t a, uint64_t e, uint64_t p) {
  uint64_t r = 1 % p;
  a %= p;
  while (e) {
    if (e & 1) r = h_mulmod(r, a, p);
    a = h_mulmod(a, a, p);
    e >>= 1;
  }
  return r;
}
static uint64_t h_inv(uint64_t a, uint64_t p) { return h_powmod(a, p - 2, p); }
static bool h_is_prime(uint64_t p) {
  if (p < 2) return false;
  for (uint64_t s : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    if (p % s == 0) return p == s;
  }
  uint64_t d = p - 1;
  int r = 0;
  while ((d & 1) == 0) { d >>= 1; r++; }
  for (uint64_t a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
    uint64_t x = h_powmod(a, d, p);
    if (x == 1 || x == p - 1) continue;
    bool comp = true;
    for (int i = 1; i < r; i++) {
      x = h_mulmod(x, x, p);
      if (x == p - 1) { comp = false; break; }
    }
    if (comp) return false;
  }
  return true;
}
// SEAL's choice of psi: the smallest primitive `degree`-th root of unity (util/numth.cpp:386-412).
static uint64_t h_minimal_primitive_root(uint64_t degree, uint64_t p) {
  uint64_t root = 0;
  for (uint64_t g = 2; g < p; g++) {
    uint64_t r = h_powmod(g, (p - 1) / degree, p);
    if (h_powmod(r, degree / 2, p) == p - 1) { root = r; break; }
  }
  uint64_t sq = h_mulmod(root, root, p), cur = root, best = root;
  for (uint64_t i = 0; i < degree / 2; i++) {
    if (cur < best) best = cur;
    cur = h_mulmod(cur, sq, p);
  }
  return best;
}
static uint32_t h_bitrev(uint32_t x, int bits) {
  uint32_t r = 0;
  for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
}
static Twiddle h_twiddle(uint64_t w, uint64_t p) {
  Twiddle t;
  t.w = w;
  t.wq = (uint64_t)(((u128)w << 64) / p);
  return t;
}
static ModConst h_modconst(uint64_t p) {
  ModConst m;
  m.p = p;
  // floor(2^128 / p) = floor((2^128 - 1) / p) for odd p > 1
  u128 all = ~(u128)0;
  u128 ratio = all / p;
  m.ratio0 = (uint64_t)ratio;
  m.ratio1 = (uint64_t)(ratio >> 64);
  m.r128 = (uint64_t)((all % p + 1) % p);
  return m;
}
// fwd[(1<<s)+g] = psi^bitrev((1<<s)+g, logn) -- SEAL's root_powers_ (util/ntt.cpp:268-276); inv = element-wise inverse.
static void h_tables(int logn, uint64_t p, std::vector<Twiddle> &fwd, std::vector<Twiddle> &inv) {
  const size_t n = (size_t)1 << logn;
  const uint64_t psi = h_minimal_primitive_root(2 * n, p), ipsi = h_inv(psi, p);
  fwd.assign(n, h_twiddle(1, p));
  inv.assign(n, h_twiddle(1, p));
  uint64_t pw = 1, ipw = 1;
  for (size_t i = 1; i < n; i++) {
    pw = h_mulmod(pw, psi, p);
    ipw = h_mulmod(ipw, ipsi, p);
    const uint32_t k = h_bitrev((uint32_t)i, logn);
    fwd[k] = h_twiddle(pw, p);
    inv[k] = h_twiddle(ipw, p);
  }
}

// ---- host-side API trace (RSG_TRACE=1): wall time and call count per C entry point, process-wide; rsg_trace_report() ----
#include <chrono>
namespace {
struct TraceRow { uint64_t calls = 0; double ms = 0; };
std::mutex g_trace_mu;
std::map<std::string, TraceRow> g_trace;
inline bool trace_on() {
  static const bool on = getenv("RSG_TRACE") && atoi(getenv("RSG_TRACE")) != 0;
  return on;
}
struct ApiTrace {
  const char *name;
  std::chrono::steady_clock::time_point t0;
  explicit ApiTrace(const char *n) : name(n) { if (trace_on()) t0 = std::chrono::steady_clock::now(); }
  ~ApiTrace() {
    if (!trace_on()) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> g(g_trace_mu);
    TraceRow &r = g_trace[name];
    r.calls++;
    r.ms += ms;
  }
};
}  // namespace
#define RSG_TRACE_CALL() ApiTrace rsg_api_trace_(__func__)
extern "C" size_t rsg_trace_report(char *buf, size_t cap, int reset) {
  std::lock_guard<std::mutex> g(g_trace_mu);
  std::string out;
  for (auto &kv : g_trace) {
    char line[160];
    snprintf(line, sizeof line, "%-28s %8llu calls %10.3f ms\n", kv.first.c_str(), (unsigned long long)kv.second.calls, kv.second.ms);
    out += line;
  }
  if (buf && cap) {
    const size_t k = std::min(cap - 1, out.size());
    memcpy(buf, out.data(), k);
    buf[k] = 0;
  }
  if (reset) g_trace.clear();
  return out.size() + 1;
}

// ------------------------------------------------------------------------------------------------------------
struct TimingRec {   // pooled: the events of a record are created once and reused by later timed passes
  const char *name;
  cudaEvent_t a, b;
};

struct WitnessTables {   // per constraint count n
  uint64_t *d_Vinv = nullptr;   // [L_R][n][n]            dense path only (built on first use)
  uint64_t *d_T = nullptr;      // [L_R][n-1][n-1] upper-triangular Toeplitz of rev(Z)^-1 (dense path only)
  uint64_t *d_Z = nullptr;      // [L_R][n+1]
  std::vector<uint64_t> h_Z;    // [L_R][n+1]
  std::vector<uint64_t> h_u;    // [L_R][max(n-1,1)]  rev(Z)^-1 mod x^(n-1)
  Twiddle *d_lagw = nullptr;    // [L_R][n] 1 / prod_{i != j} (j - i): Lagrange denominators (instance.cuh)
  uint64_t *d_Zvec = nullptr;   // [n+1][L_R][N_R]: coefficients_for_Z as ring elements (every slot = the constant), rinocchio
  bool fast_ready = false;      // quasi-linear path (witness_fast.cuh)
  FastTables ft;
};

struct rsg_context {
  int device = 0;
  size_t N_R = 0, L_R = 0, N_E = 0, L_E = 0;
  int logN = 0;
  std::vector<uint64_t> q, Q;
  DevParams hp;                 // host copy
  DevParams *d_params = nullptr;
  ModConst *d_modq = nullptr, *d_modQ = nullptr;
  Twiddle *d_invN_q = nullptr, *d_invNw_q = nullptr, *d_invN_Q = nullptr, *d_invNw_Q = nullptr;
  std::vector<void *> owned;    // device allocations freed at destroy
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::mutex mu;
  uint64_t launches = 0;
  bool timing = false;
  std::vector<TimingRec> recs;      // event pool; the first n_recs entries belong to the current timed pass
  size_t n_recs = 0;
  std::map<size_t, WitnessTables> wit;
  std::vector<std::vector<uint64_t>> h_fwdq;   // host copy of the forward twiddles mod q_j (witness_fast tables)
  std::vector<rsg_host::NttTables> h_ntt;       // the same with their inverses, built on first use (host_poly.hpp)
  uint32_t wf_ts = 0;               // RSG_WF_TS: cap on the witness kernels' transform size (tests: blocked mode at small n)
  int witness_mode = 0;             // 0 = auto, 1 = dense (RSG_WITNESS=dense), 2 = quasi-linear wherever it applies (RSG_WITNESS=fast)
  int wf_sl = 0;                    // RSG_WF_SL: slots per CTA of the quasi-linear kernels (0 = auto)
  int lin_threads = 0;
  // RSG_LIN_SPLITS / RSG_LIN_UNROLL / RSG_LIN_THREADS: tuning overrides of k_crs_lincomb's launch shape.  Measured on B200
  // (C4): terms in flight per thread 1 / 2 / 3 / 4 / 8 -> 5836 / 5903 / 5980 / 5572 / 4300 GB/s on one 2063-term inner product
  // (tools/lin_tune.py) and 2 / 3 / 4 -> 5820 / 4754 / 5493 GB/s inside the prover (bench.py): two is the robust choice
  int lin_splits = 0, lin_unroll = 2;
  int wf_threads = 0;               // RSG_WF_THREADS: threads per CTA of the quasi-linear kernels (0 = auto)
  // scratch (grown on demand)
  uint64_t *d_plain = nullptr, *d_pntt = nullptr, *d_partial = nullptr;
  size_t cap_plain = 0, cap_pntt = 0, cap_partial = 0;
  uint32_t *d_term = nullptr, *d_pidx = nullptr, *d_eidx = nullptr;
  size_t cap_term = 0, cap_pidx = 0, cap_eidx = 0;
  uint64_t *d_out_scratch = nullptr;
  size_t cap_out_scratch = 0;
  uint64_t *d_chunk = nullptr;      // per-chunk partial encodings of lincomb_terms
  size_t cap_chunk = 0;
  uint64_t *d_evals = nullptr, *d_wit = nullptr;   // prover scratch: 9n evaluations; [6n coeffs | n+1 H]
  size_t cap_evals = 0, cap_wit = 0;
  size_t pntt_budget_words = (size_t)8 << 27;      // 8 GiB of NTT-domain plaintexts per chunk
  uint8_t *d_flags = nullptr;
  size_t cap_flags = 0;
  uint64_t *d_zk = nullptr;         // d1, d2, d3 of the zero-knowledge witness map
  size_t cap_zk = 0;
  // transparent-ciphertext emulation (kernels.cuh: k_probe): per-term candidate flags [slot][L_R][T], running sums, nz words
  uint8_t *d_probe = nullptr;
  size_t cap_probe = 0;
  uint64_t *d_probe_carry = nullptr;   // [2][MAX_LR]: second row for the second part of a merged lincomb
  uint32_t *d_nz = nullptr;
  uint64_t *d_psi_pow = nullptr;    // psi^i mod Q_0, i < N_E (k_probe_eval)
  uint32_t *d_mrg = nullptr;        // index lists of a merged lincomb
  size_t cap_mrg = 0;
  uint64_t *d_pval = nullptr;       // k_probe_eval output [plain slot][L_R]
  size_t cap_pval = 0;
  int merge_mode = 1;               // RSG_MERGE=0: never merge inner products over the same CRS range
  uint64_t *d_exact = nullptr;      // scratch encodings of the exact (slow) path
  size_t cap_exact = 0;
  uint64_t *d_ip = nullptr;         // the separate inner products of rsg_groth16_prove
  size_t cap_ip = 0;
  uint64_t exact_fallbacks = 0;     // how often a flagged prefix had to be resolved exactly
  DecodeConsts *d_dec = nullptr;    // constants of rsg_decode (decode.cuh), built on first use
  int Q_bits = 0;                   // bit length of Q = prod Q_l
  uint64_t *d_wfB = nullptr;        // witness_fast scratch buffers in global memory (S = 16384)
  size_t cap_wfB = 0;
  uint64_t *d_decode = nullptr;     // rsg_decode scratch
  size_t cap_decode = 0;
  uint64_t *d_encode = nullptr;     // rsg_encode scratch
  size_t cap_encode = 0;
  uint64_t st_wf = 0, st_wd = 0;
  uint64_t st_lin_terms = 0, st_lin_plain = 0, st_lin_launches = 0, st_fwd_polys = 0, st_inv_polys = 0, st_merged = 0;
  uint64_t st_lin_shared = 0;       // CRS elements of paired splits that a neighbouring CTA streams too (counted once in the byte floor)
  // ---- static-plan prover (prover_fast.cuh)
  int fast_mode = 1;                // RSG_FAST=0: always the host-driven exact path
  int lin_mode = 2;                 // static-plan lincomb: 2 = k_crs_lincomb_wide (default), 0 = k_crs_lincomb (RSG_LIN=narrow), 1 = TMA-fed (RSG_LIN=tma)
  int lt_ctas = 2;                  // RSG_LT_CTAS: persistent CTAs per SM of k_crs_lincomb_tma
  int overlap_mode = 0;             // RSG_OVERLAP=1: lincomb of one term group on a second stream under the next group's NTTs
  int fast_splits = 0;              // RSG_FAST_SPLITS: number of term chunks of the one-launch lincomb (0 = auto)
  int ntt_half = 0;                 // RSG_NTT_HALF=1 (experiment, see fast_launch_ntt)
  int enc_rows = 1;                 // N_E = 2^14, N_R <= N_E / 2: compact-row batch encode (k_encode_*_rows); RSG_ENC_ROWS=0 -> encode_body
  int ntt_cluster = 1;              // N_E = 2^14, FP64 path: k_lift_fwd_ntt_f64_cl (4-CTA clusters); RSG_NTT_CLUSTER=0 -> single-CTA kernel
  bool lift_smallq = false;         // LiftIoSmallQ applies (kernels.cuh): t < 2^54 and t / min Q_l < 2^11; RSG_LIFT=barrett turns it off
  int overlap_chunks = 4;           // RSG_OVERLAP_CHUNKS: term chunks (<= 128 terms each) per overlap phase
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_phase[8] = {};     // NTT group done (stream -> stream2)
  cudaEvent_t ev_join = nullptr;    // stream2 -> stream
  struct FastPlan *fplan = nullptr; // cached index arrays of the last layout
  uint8_t *d_fp_flags = nullptr;    // [n_elems elem_flag | n_slots slot_skip]
  size_t cap_fp_flags = 0;
  uint64_t *d_fp_parts = nullptr, *d_fp_nttsrc = nullptr;
  size_t cap_fp_parts = 0, cap_fp_nttsrc = 0;
  uint64_t *d_fp_totals = nullptr;  // [6][MAX_LR] probe totals
  uint64_t *fp_block = nullptr;     // rsg_groth16_lincombs_shard: probe block of this call (device), see include/rsgpu.h
  uint32_t fp_pstride = 0;
  uint32_t *d_fp_status = nullptr, *h_fp_status = nullptr;   // FPS_WORDS device words + pinned host mirror
  uint64_t st_fast = 0, st_fast_fallback = 0;
  bool f64_ntt = false;             // every Q_l < 2^49: forward NTTs of the plaintext pipeline run on the FP64 pipe
  int ntt_mode = 0;                 // 0 = auto; RSG_NTT=int forces the integer kernel
  size_t enc_words() const { return L_R * 2 * L_E * N_E; }
  size_t ring_words() const { return L_R * N_R; }
};
struct rsg_crs {
  rsg_context *ctx;
  size_t n;
  uint64_t *d;
};
struct rsg_ringvec {
  rsg_context *ctx;
  size_t n;
  uint64_t *d;
  bool owned = true;
};

struct rsg_r1cs {
  rsg_context *ctx;
  size_t n, n_io, n_aux;
  uint32_t *d_row_ptr = nullptr, *d_col = nullptr;
  uint64_t *d_coeff = nullptr;
  std::vector<uint32_t> h_row_ptr, h_col;   // host copy of the CSR system (the instance map transposes it on first use)
  std::vector<uint64_t> h_coeff;
  uint32_t *d_col_ptr = nullptr, *d_rows = nullptr;   // CSC of A | B | C over variables 0..n_io+n_aux (instance map)
  uint64_t *d_ccoeff = nullptr;
  std::vector<uint64_t> h_const;   // [2][L_R][n]: constant-wire coefficient of A_i, B_i mod q_j
  uint64_t *d_cc = nullptr;        // [2][L_R][n]: its interpolant V^-1 * const, built on first use
};

struct LaunchScope {   // counts launches and optionally brackets them with events
  rsg_context *c;
  const char *name;
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  LaunchScope(rsg_context *ctx, const char *nm, cudaStream_t on = nullptr) : c(ctx), name(nm), st(on ? on : ctx->stream) {
    c->launches++;
    if (c->timing) {
      if (c->n_recs == c->recs.size()) {
        TimingRec r{nm, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        c->recs.push_back(r);
      }
      TimingRec &r = c->recs[c->n_recs];
      r.name = nm;
      a = r.a;
      b = r.b;
      cudaEventRecord(a, st);
    }
  }
  ~LaunchScope() {
    if (c->timing && a) {
      cudaEventRecord(b, st);
      c->n_recs++;
    }
  }
};

template <typename T>
static int dev_alloc(rsg_context *c, T **p, size_t count, bool track = true) {
  void *v = nullptr;
  CUDA_TRY(cudaMalloc(&v, std::max<size_t>(count, 1) * sizeof(T)));
  *p = (T *)v;
  if (track) c->owned.push_back(v);
  return RSG_OK;
}
template <typename T>
static int ensure(rsg_context *c, T **p, size_t *cap, size_t need) {
  if (*cap >= need) return RSG_OK;
  if (*p) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaFree(*p));
    *p = nullptr;
  }
  void *v = nullptr;
  CUDA_TRY(cudaMalloc(&v, need * sizeof(T)));
  *p = (T *)v;
  *cap = need;
  return RSG_OK;
}
template <typename T>
static int upload_vec(rsg_context *c, const std::vector<T> &h, T **d) {
  int rc = dev_alloc(c, d, h.size());
  if (rc) return rc;
  CUDA_TRY(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
extern "C" int rsg_context_create(rsg_context **out, size_t N_R, size_t L_R, const uint64_t *q, size_t N_E, size_t L_E,
                                  const uint64_t *Q, int device) {
  if (!out || !q || !Q) return fail(RSG_ERR_ARG, "null argument");
  if (N_E < 256 || N_E > 32768 || (N_E & (N_E - 1))) return fail(RSG_ERR_UNSUPPORTED, "N_E must be a power of two in [256, 32768]");
  if (N_R == 0 || N_R > N_E || (N_R & (N_R - 1))) return fail(RSG_ERR_ARG, "N_R must be a power of two <= N_E");
  if (L_R == 0 || L_R > (size_t)MAX_LR || L_E == 0 || L_E > (size_t)MAX_LE) return fail(RSG_ERR_ARG, "limb count out of range");
  for (size_t i = 0; i < L_R + L_E; i++) {
    uint64_t p = i < L_R ? q[i] : Q[i - L_R];
    if (p >= (1ull << 61) || p < 2 * N_E || (p - 1) % (2 * N_E) != 0 || !h_is_prime(p))
      return fail(RSG_ERR_ARG, "every modulus must be a prime < 2^61 with p = 1 mod 2*N_E");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(RSG_ERR_CUDA, "no CUDA device: librsgpu has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(RSG_ERR_ARG, "bad device index");
  CUDA_TRY(cudaSetDevice(device));
  rsg_context *c = new rsg_context();
  struct Guard {   // an error on the way out releases what was allocated so far
    rsg_context *c;
    ~Guard() { if (c) rsg_context_destroy(c); }
  } guard{c};
  c->device = device;
  c->N_R = N_R; c->L_R = L_R; c->N_E = N_E; c->L_E = L_E;
  while (((size_t)1 << c->logN) < N_E) c->logN++;
  c->q.assign(q, q + L_R);
  c->Q.assign(Q, Q + L_E);
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  if (const char *m = getenv("RSG_NTT")) c->ntt_mode = !strcmp(m, "int") ? 1 : (!strcmp(m, "f64s") ? 2 : 0);

  DevParams &hp = c->hp;
  memset(&hp, 0, sizeof(hp));
  hp.N_R = (uint32_t)N_R; hp.L_R = (uint32_t)L_R; hp.N_E = (uint32_t)N_E; hp.L_E = (uint32_t)L_E; hp.logN_E = (uint32_t)c->logN;
  std::vector<Twiddle> fwd, inv;
  int rc;
  for (size_t l = 0; l < L_E; l++) {
    hp.Q[l] = h_modconst(Q[l]);
    h_tables(c->logN, Q[l], fwd, inv);
    Twiddle *d;
    if ((rc = upload_vec(c, fwd, &d))) return rc;
    hp.fwdQ[l] = d;
    if ((rc = upload_vec(c, inv, &d))) return rc;
    hp.invQ[l] = d;
    hp.invN_Q[l] = h_twiddle(h_inv(N_E % Q[l], Q[l]), Q[l]);
    hp.invNw_Q[l] = h_twiddle(h_mulmod(hp.invN_Q[l].w, inv[1].w, Q[l]), Q[l]);
  }
  // FP64-pipe forward tables (ntt_f64.cuh) when every data-level prime is below 2^49
  c->f64_ntt = true;
  for (size_t l = 0; l < L_E; l++) c->f64_ntt = c->f64_ntt && Q[l] < (1ull << 49);
  if (c->f64_ntt)
    for (size_t l = 0; l < L_E; l++) {
      h_tables(c->logN, Q[l], fwd, inv);
      std::vector<double> tf(fwd.size());
      for (size_t i = 0; i < fwd.size(); i++) tf[i] = (double)fwd[i].w;   // exact: w < 2^49
      double *d;
      if ((rc = upload_vec(c, tf, &d))) return rc;
      hp.fwdQ_f64[l] = d;
      hp.Qinv_f64[l] = (double)(1.0L / (long double)Q[l]);
    }
  for (size_t j = 0; j < L_R; j++) {
    hp.q[j] = h_modconst(q[j]);
    h_tables(c->logN, q[j], fwd, inv);
    c->h_fwdq.emplace_back(fwd.size());
    for (size_t i = 0; i < fwd.size(); i++) c->h_fwdq.back()[i] = fwd[i].w;
    Twiddle *d;
    if ((rc = upload_vec(c, fwd, &d))) return rc;
    hp.fwdq[j] = d;
    if ((rc = upload_vec(c, inv, &d))) return rc;
    hp.invq[j] = d;
    hp.invN_q[j] = h_twiddle(h_inv(N_E % q[j], q[j]), q[j]);
    hp.invNw_q[j] = h_twiddle(h_mulmod(hp.invN_q[j].w, inv[1].w, q[j]), q[j]);
    hp.thr[j] = (q[j] + 1) >> 1;
    for (size_t l = 0; l < L_E; l++) hp.tmodQ[j][l] = q[j] % Q[l];
  }
  {  // batchencoder.cpp:64-88
    std::vector<uint32_t> map(N_E);
    const size_t row = N_E >> 1, m = N_E << 1;
    uint64_t pos = 1;
    for (size_t i = 0; i < row; i++) {
      map[i] = h_bitrev((uint32_t)((pos - 1) >> 1), c->logN);
      map[row | i] = h_bitrev((uint32_t)((m - pos - 1) >> 1), c->logN);
      pos = (pos * 3) & (m - 1);
    }
    uint32_t *d;
    if ((rc = upload_vec(c, map, &d))) return rc;
    hp.index_map = d;
  }
  if ((rc = dev_alloc(c, &c->d_params, 1))) return rc;
  CUDA_TRY(cudaMemcpy(c->d_params, &hp, sizeof(hp), cudaMemcpyHostToDevice));
  {
    std::vector<ModConst> mq(hp.q, hp.q + L_R), mQ(hp.Q, hp.Q + L_E);
    if ((rc = upload_vec(c, mq, &c->d_modq))) return rc;
    if ((rc = upload_vec(c, mQ, &c->d_modQ))) return rc;
    std::vector<Twiddle> a(hp.invN_q, hp.invN_q + L_R), b(hp.invNw_q, hp.invNw_q + L_R), a2(hp.invN_Q, hp.invN_Q + L_E),
        b2(hp.invNw_Q, hp.invNw_Q + L_E);
    if ((rc = upload_vec(c, a, &c->d_invN_q)) || (rc = upload_vec(c, b, &c->d_invNw_q)) || (rc = upload_vec(c, a2, &c->d_invN_Q)) ||
        (rc = upload_vec(c, b2, &c->d_invNw_Q)))
      return rc;
  }
  if ((rc = dev_alloc(c, &c->d_probe_carry, 2 * MAX_LR, false))) return rc;
  {
    std::vector<uint64_t> pw(N_E);
    const uint64_t psi = h_minimal_primitive_root(2 * N_E, Q[0]);
    pw[0] = 1;
    for (size_t i = 1; i < N_E; i++) pw[i] = h_mulmod(pw[i - 1], psi, Q[0]);
    if ((rc = upload_vec(c, pw, &c->d_psi_pow))) return rc;
  }
  if (const char *m = getenv("RSG_MERGE")) c->merge_mode = atoi(m);
  if (const char *m = getenv("RSG_FAST")) c->fast_mode = atoi(m);
  if (const char *m = getenv("RSG_LIN")) c->lin_mode = !strcmp(m, "tma") ? 1 : (!strcmp(m, "narrow") ? 0 : 2);
  if (const char *m = getenv("RSG_LT_CTAS")) c->lt_ctas = std::max(1, atoi(m));
  if (const char *m = getenv("RSG_OVERLAP")) c->overlap_mode = atoi(m);
  if (const char *m = getenv("RSG_FAST_SPLITS")) c->fast_splits = atoi(m);
  if (const char *m = getenv("RSG_NTT_HALF")) c->ntt_half = atoi(m);
  if (const char *m = getenv("RSG_NTT_CLUSTER")) c->ntt_cluster = atoi(m);
  if (const char *m = getenv("RSG_ENC_ROWS")) c->enc_rows = atoi(m);
  {
    uint64_t tmax = 0, Qmin = ~0ull;
    for (uint64_t p : c->q) tmax = std::max(tmax, p);
    for (uint64_t p : c->Q) Qmin = std::min(Qmin, p);
    const char *m = getenv("RSG_LIFT");
    c->lift_smallq = c->f64_ntt && tmax < (1ull << 54) && tmax / Qmin < 2048 && !(m && !strcmp(m, "barrett"));
  }
  if (const char *m = getenv("RSG_OVERLAP_CHUNKS")) c->overlap_chunks = std::max(1, atoi(m));
  if (const char *m = getenv("RSG_WITNESS")) c->witness_mode = !strcmp(m, "dense") ? 1 : (!strcmp(m, "fast") ? 2 : 0);
  if (const char *m = getenv("RSG_WF_SL")) c->wf_sl = atoi(m);
  if (const char *m = getenv("RSG_WF_TS")) c->wf_ts = (uint32_t)atoi(m);
  if (const char *m = getenv("RSG_LIN_SPLITS")) c->lin_splits = atoi(m);
  if (const char *m = getenv("RSG_LIN_THREADS")) c->lin_threads = atoi(m);
  if (const char *m = getenv("RSG_LIN_UNROLL")) c->lin_unroll = atoi(m);
  if (const char *m = getenv("RSG_WF_THREADS")) c->wf_threads = atoi(m);
  if (const char *m = getenv("RSG_PNTT_BUDGET_WORDS")) c->pntt_budget_words = std::max<size_t>(1, strtoull(m, nullptr, 10));   // tests: force chunking
  if ((rc = dev_alloc(c, &c->d_nz, MAX_LR, false))) return rc;
  guard.c = nullptr;
  *out = c;
  return RSG_OK;
}

static void fast_release(rsg_context *c);
extern "C" void rsg_context_destroy(rsg_context *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (void *p : c->owned) cudaFree(p);
  for (auto &kv : c->wit) { cudaFree(kv.second.d_Vinv); cudaFree(kv.second.d_T); cudaFree(kv.second.d_Z); cudaFree(kv.second.d_Zvec); }   // d_lagw / fast tables: c->owned
  cudaFree(c->d_plain); cudaFree(c->d_pntt); cudaFree(c->d_partial);
  cudaFree(c->d_term); cudaFree(c->d_pidx); cudaFree(c->d_eidx); cudaFree(c->d_flags); cudaFree(c->d_out_scratch);
  cudaFree(c->d_chunk); cudaFree(c->d_evals); cudaFree(c->d_wit); cudaFree(c->d_zk);
  cudaFree(c->d_decode); cudaFree(c->d_wfB); cudaFree(c->d_encode);
  cudaFree(c->d_probe); cudaFree(c->d_probe_carry); cudaFree(c->d_nz); cudaFree(c->d_exact); cudaFree(c->d_ip);
  for (auto &r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  fast_release(c);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}
extern "C" int rsg_context_sync(rsg_context *c) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
extern "C" int rsg_context_set_stream(rsg_context *c, void *s) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->own_stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return RSG_OK;
}
extern "C" uint64_t rsg_context_launch_count(const rsg_context *c) { return c ? c->launches : 0; }
extern "C" uint64_t rsg_context_stat(const rsg_context *c, const char *name) {
  if (!c || !name) return 0;
  const std::string n(name);
  if (n == "lincomb_terms") return c->st_lin_terms;       // CRS elements streamed by k_crs_lincomb
  if (n == "lincomb_plain_terms") return c->st_lin_plain; // of which multiplied by an NTT-domain plaintext
  if (n == "lincomb_shared_terms") return c->st_lin_shared;   // CRS elements read by both splits of a pair (second read: L2)
  if (n == "lincomb_launches") return c->st_lin_launches;
  if (n == "ntt_forward_polys") return c->st_fwd_polys;   // N_E-point forward transforms (k_lift_fwd_ntt)
  if (n == "ntt_inverse_polys") return c->st_inv_polys;   // N_E-point inverse transforms (k_encode_intt)
  if (n == "merged_lincombs") return c->st_merged;
  if (n == "exact_fallbacks") return c->exact_fallbacks;
  if (n == "witness_fast_launches") return c->st_wf;      // k_interp_fast / k_quotient_fast launches
  if (n == "witness_dense_launches") return c->st_wd;     // k_modmat* / k_conv_top launches
  if (n == "fast_proofs") return c->st_fast;              // lincomb phases run as the static launch sequence (prover_fast.cuh)
  if (n == "fast_fallbacks") return c->st_fast_fallback;  // ... of which a probe candidate sent to the exact path
  return 0;
}
extern "C" int rsg_context_enable_timing(rsg_context *c, int on) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  std::lock_guard<std::mutex> g(c->mu);
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->stream2) cudaStreamSynchronize(c->stream2);
  c->n_recs = 0;   // the pooled events are reused
  c->timing = on != 0;
  return RSG_OK;
}
extern "C" int rsg_context_last_timing(rsg_context *c, const char *kernel, float *ms, uint64_t *launches) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (c->stream2) CUDA_TRY(cudaStreamSynchronize(c->stream2));
  float total = 0;
  uint64_t n = 0;
  for (size_t i = 0; i < c->n_recs; i++) {
    const TimingRec &r = c->recs[i];
    if (!kernel || !strcmp(r.name, kernel)) {
      float t = 0;
      CUDA_TRY(cudaEventElapsedTime(&t, r.a, r.b));
      total += t;
      n++;
    }
  }
  if (ms) *ms = total;
  if (launches) *launches = n;
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// arenas
extern "C" int rsg_crs_create(rsg_context *c, size_t n, rsg_crs **out) {
  RSG_TRACE_CALL();
  if (!c || !out) return fail(RSG_ERR_STATE, "context not set");
  CUDA_TRY(cudaSetDevice(c->device));
  rsg_crs *r = new rsg_crs{c, n, nullptr};
  void *v = nullptr;
  cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * c->enc_words() * 8);
  if (e != cudaSuccess) { delete r; return fail(RSG_ERR_CUDA, std::string("cudaMalloc CRS: ") + cudaGetErrorString(e)); }
  r->d = (uint64_t *)v;
  *out = r;
  return RSG_OK;
}
extern "C" int rsg_crs_upload(rsg_crs *r, size_t first, size_t count, const uint64_t *h) {
  RSG_TRACE_CALL();
  if (!r || first + count > r->n) return fail(RSG_ERR_ARG, "CRS range");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(r->d + first * c->enc_words(), h, count * c->enc_words() * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
extern "C" int rsg_crs_download(const rsg_crs *r, size_t first, size_t count, uint64_t *h) {
  RSG_TRACE_CALL();
  if (!r || first + count > r->n) return fail(RSG_ERR_ARG, "CRS range");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(h, r->d + first * c->enc_words(), count * c->enc_words() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
static int fill_uniform(rsg_context *c, uint64_t *d, size_t words, uint32_t row_words, const ModConst *mods, uint32_t n_mods,
                        uint64_t seed, uint64_t w_base = 0) {
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  LaunchScope ls(c, "k_fill_uniform");
  k_fill_uniform<<<148 * 8, 256, 0, c->stream>>>(d, words, row_words, mods, n_mods, 1, seed, w_base);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_crs_fill_uniform(rsg_crs *r, uint64_t seed) {
  if (!r) return fail(RSG_ERR_ARG, "null CRS");
  rsg_context *c = r->ctx;
  return fill_uniform(c, r->d, r->n * c->enc_words(), (uint32_t)c->N_E, c->d_modQ, (uint32_t)c->L_E, seed);
}
extern "C" int rsg_crs_fill_uniform_at(rsg_crs *r, size_t first, size_t count, uint64_t virtual_first, uint64_t seed) {
  if (!r || first + count > r->n) return fail(RSG_ERR_ARG, "CRS range");
  if (!count) return RSG_OK;
  rsg_context *c = r->ctx;
  return fill_uniform(c, r->d + first * c->enc_words(), count * c->enc_words(), (uint32_t)c->N_E, c->d_modQ, (uint32_t)c->L_E, seed,
                      virtual_first * c->enc_words());
}
extern "C" uint64_t *rsg_crs_device_ptr(rsg_crs *r) { return r ? r->d : nullptr; }
extern "C" void rsg_crs_destroy(rsg_crs *r) {
  RSG_TRACE_CALL();
  if (!r) return;
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->d);
  delete r;
}

extern "C" int rsg_ringvec_create(rsg_context *c, size_t n, rsg_ringvec **out) {
  RSG_TRACE_CALL();
  if (!c || !out) return fail(RSG_ERR_STATE, "context not set");
  CUDA_TRY(cudaSetDevice(c->device));
  rsg_ringvec *r = new rsg_ringvec{c, n, nullptr};
  void *v = nullptr;
  cudaError_t e = cudaMalloc(&v, std::max<size_t>(n, 1) * c->ring_words() * 8);
  if (e != cudaSuccess) { delete r; return fail(RSG_ERR_CUDA, std::string("cudaMalloc ringvec: ") + cudaGetErrorString(e)); }
  r->d = (uint64_t *)v;
  cudaMemsetAsync(r->d, 0, std::max<size_t>(n, 1) * c->ring_words() * 8, c->stream);
  *out = r;
  return RSG_OK;
}
extern "C" int rsg_ringvec_upload(rsg_ringvec *r, size_t first, size_t count, const uint64_t *h) {
  RSG_TRACE_CALL();
  if (!r || first + count > r->n) return fail(RSG_ERR_ARG, "ringvec range");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(r->d + first * c->ring_words(), h, count * c->ring_words() * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
extern "C" int rsg_ringvec_download(const rsg_ringvec *r, size_t first, size_t count, uint64_t *h) {
  RSG_TRACE_CALL();
  if (!r || first + count > r->n) return fail(RSG_ERR_ARG, "ringvec range");
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(h, r->d + first * c->ring_words(), count * c->ring_words() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
extern "C" int rsg_ringvec_fill_uniform(rsg_ringvec *r, uint64_t seed) {
  if (!r) return fail(RSG_ERR_ARG, "null ringvec");
  rsg_context *c = r->ctx;
  return fill_uniform(c, r->d, r->n * c->ring_words(), (uint32_t)c->N_R, c->d_modq, (uint32_t)c->L_R, seed);
}
extern "C" uint64_t *rsg_ringvec_device_ptr(rsg_ringvec *r) { return r ? r->d : nullptr; }
extern "C" size_t rsg_ringvec_size(const rsg_ringvec *r) { return r ? r->n : 0; }
extern "C" void rsg_ringvec_destroy(rsg_ringvec *r) {
  RSG_TRACE_CALL();
  if (!r) return;
  cudaStreamSynchronize(r->ctx->stream);
  if (r->owned) cudaFree(r->d);
  delete r;
}
extern "C" int rsg_ringvec_wrap(rsg_context *c, uint64_t *d_words, size_t n, rsg_ringvec **out) {
  if (!c || !d_words || !out) return fail(RSG_ERR_ARG, "null argument");
  rsg_ringvec *r = new rsg_ringvec{c, n, d_words};
  r->owned = false;
  *out = r;
  return RSG_OK;
}

extern "C" int rsg_ringvec_is_zero_prefix(const rsg_ringvec *r, size_t first, size_t count, uint8_t *h_flags) {
  RSG_TRACE_CALL();
  if (!r || first + count > r->n || !h_flags) return fail(RSG_ERR_ARG, "ringvec range");
  if (!count) return RSG_OK;
  rsg_context *c = r->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure(c, &c->d_flags, &c->cap_flags, count);
  if (rc) return rc;
  {
    LaunchScope ls(c, "k_is_zero_prefix");
    k_is_zero_prefix<<<(unsigned)count, 256, 0, c->stream>>>(r->d + first * c->ring_words(), (uint32_t)c->ring_words(), c->d_flags);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(h_flags, c->d_flags, count, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// element-wise RingElem operators
static int ring_range(const rsg_ringvec *v, size_t first, size_t count) { return v && first + count <= v->n; }
static dim3 ring_grid(const rsg_context *c, size_t count) {
  return dim3((unsigned)((c->N_R + 255) / 256), (unsigned)c->L_R, (unsigned)count);
}
extern "C" int rsg_ring_binop(rsg_context *c, int op, const rsg_ringvec *a, size_t a_first, const rsg_ringvec *b, size_t b_first,
                              rsg_ringvec *out, size_t out_first, size_t count) {
  RSG_TRACE_CALL();
  if (!c || op < 0 || op > 2) return fail(RSG_ERR_ARG, "bad operator");
  if (!ring_range(a, a_first, count) || !ring_range(b, b_first, count) || !ring_range(out, out_first, count)) return fail(RSG_ERR_ARG, "ringvec range");
  if (!count) return RSG_OK;
  if (count > 65535) return fail(RSG_ERR_UNSUPPORTED, "at most 65535 elements per call");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t W = c->ring_words();
  LaunchScope ls(c, "k_ring_binop");
  k_ring_binop<<<ring_grid(c, count), 256, 0, c->stream>>>(c->d_modq, op, a->d + a_first * W, b->d + b_first * W, out->d + out_first * W,
                                                           (uint32_t)c->N_R, (uint32_t)c->L_R);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_ring_scalar_op(rsg_context *c, int op, const rsg_ringvec *a, size_t a_first, uint64_t scalar, rsg_ringvec *out,
                                  size_t out_first, size_t count) {
  if (!c || op < 0 || op > 2) return fail(RSG_ERR_ARG, "bad operator");
  if (!ring_range(a, a_first, count) || !ring_range(out, out_first, count)) return fail(RSG_ERR_ARG, "ringvec range");
  if (!count) return RSG_OK;
  if (count > 65535) return fail(RSG_ERR_UNSUPPORTED, "at most 65535 elements per call");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t W = c->ring_words();
  LaunchScope ls(c, "k_ring_scalar");
  k_ring_scalar<<<ring_grid(c, count), 256, 0, c->stream>>>(c->d_modq, op, a->d + a_first * W, scalar, out->d + out_first * W,
                                                            (uint32_t)c->N_R, (uint32_t)c->L_R);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_ring_negate(rsg_context *c, const rsg_ringvec *a, size_t a_first, rsg_ringvec *out, size_t out_first, size_t count) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  if (!ring_range(a, a_first, count) || !ring_range(out, out_first, count)) return fail(RSG_ERR_ARG, "ringvec range");
  if (!count) return RSG_OK;
  if (count > 65535) return fail(RSG_ERR_UNSUPPORTED, "at most 65535 elements per call");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t W = c->ring_words();
  LaunchScope ls(c, "k_ring_negate");
  k_ring_negate<<<ring_grid(c, count), 256, 0, c->stream>>>(c->d_modq, a->d + a_first * W, out->d + out_first * W, (uint32_t)c->N_R,
                                                            (uint32_t)c->L_R);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_ring_invert(rsg_context *c, const rsg_ringvec *a, size_t a_first, rsg_ringvec *out, size_t out_first, size_t count,
                               uint8_t *h_ok) {
  RSG_TRACE_CALL();
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  if (!ring_range(a, a_first, count) || !ring_range(out, out_first, count)) return fail(RSG_ERR_ARG, "ringvec range");
  if (!count) return RSG_OK;
  if (count > 65535) return fail(RSG_ERR_UNSUPPORTED, "at most 65535 elements per call");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t W = c->ring_words();
  uint32_t *d_bad = nullptr;
  void *v = nullptr;
  CUDA_TRY(cudaMalloc(&v, count * 4));
  d_bad = (uint32_t *)v;
  CUDA_TRY(cudaMemsetAsync(d_bad, 0, count * 4, c->stream));
  {
    LaunchScope ls(c, "k_ring_invert");
    k_ring_invert<<<ring_grid(c, count), 256, 0, c->stream>>>(c->d_modq, a->d + a_first * W, out->d + out_first * W, (uint32_t)c->N_R,
                                                              (uint32_t)c->L_R, d_bad);
  }
  std::vector<uint32_t> bad(count);
  cudaError_t e = cudaMemcpyAsync(bad.data(), d_bad, count * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d_bad);
  if (e != cudaSuccess) return fail(RSG_ERR_CUDA, cudaGetErrorString(e));
  bool any = false;
  for (size_t i = 0; i < count; i++) {
    if (h_ok) h_ok[i] = bad[i] ? 0 : 1;
    any |= bad[i] != 0;
  }
  return any ? fail(RSG_ERR_NOTINV, "element is not invertible in ring") : RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// NTT launch helpers.  N_E = 2^15 runs as two 2^14-point halves per polynomial (LV = 1, see kernels.cuh).
static int local_logn(int logN) { return logN > 14 ? 14 : logN; }
static unsigned ntt_threads(int logN) { return (unsigned)std::max(32, std::min(512, (1 << local_logn(logN)) / 16)); }
static size_t ntt_smem(int logN) { return (size_t)padded_words(1u << local_logn(logN)) * 8; }

template <int LOGN, int LV>
static int set_smem_attrs() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev]) return RSG_OK;
  const int bytes = (int)padded_words(1u << LOGN) * 8;
  CUDA_TRY(cudaFuncSetAttribute(k_encode_intt<LOGN, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt<LOGN, LV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt<LOGN, LV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<LOGN, LV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<LOGN, LV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_r96<LOGN, LV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<LOGN, LV, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64<LOGN, LV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_lift_fwd_ntt_f64_r96<LOGN, LV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_ntt<LOGN, LV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  CUDA_TRY(cudaFuncSetAttribute(k_ntt<LOGN, LV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[dev] = true;
  return RSG_OK;
}

#define DISPATCH_LOGN(logn, ...)                  \
  switch (logn) {                                 \
    case 8: { constexpr int LG = 8, LV = 0; __VA_ARGS__; } break;   \
    case 9: { constexpr int LG = 9, LV = 0; __VA_ARGS__; } break;   \
    case 10: { constexpr int LG = 10, LV = 0; __VA_ARGS__; } break; \
    case 11: { constexpr int LG = 11, LV = 0; __VA_ARGS__; } break; \
    case 12: { constexpr int LG = 12, LV = 0; __VA_ARGS__; } break; \
    case 13: { constexpr int LG = 13, LV = 0; __VA_ARGS__; } break; \
    case 14: { constexpr int LG = 14, LV = 0; __VA_ARGS__; } break; \
    case 15: { constexpr int LG = 14, LV = 1; __VA_ARGS__; } break; \
    default: return fail(RSG_ERR_UNSUPPORTED, "unsupported N_E"); \
  }

// split inverse transforms leave the last level to this kernel; fixed_mod = 0xFFFFFFFF: modulus = polynomial index % n_mod
static int launch_intt_finish(rsg_context *c, uint64_t *d, size_t polys, const ModConst *mods, const Twiddle *invn,
                              const Twiddle *invnw, uint32_t n_mod, uint32_t fixed_mod, uint32_t centre = 0) {
  const uint32_t half = (uint32_t)(c->N_E / 2);
  LaunchScope ls(c, "k_intt_finish");
  k_intt_finish<<<(unsigned)((polys * half + 255) / 256), 256, 0, c->stream>>>(d, half, polys, mods, invn, invnw, n_mod, fixed_mod, centre);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

static int launch_encode(rsg_context *c, const uint64_t *d_ring, const uint32_t *d_eidx, size_t count, uint64_t *d_plain) {
  if (!count) return RSG_OK;
  c->st_inv_polys += count * c->L_R;
  const unsigned th = ntt_threads(c->logN);
  const size_t sm = ntt_smem(c->logN);
  const unsigned split = c->logN > 14 ? 2 : 1;
  dim3 grid((unsigned)count * split, (unsigned)c->L_R);
  if (c->enc_rows && c->logN == 14 && 2 * c->N_R <= c->N_E) {
    LaunchScope ls(c, "k_encode_intt");
    static bool attr_done[64] = {};
    if (!attr_done[c->device]) {
      CUDA_TRY(cudaFuncSetAttribute(k_encode_intt_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_ROWS_SMEM));
      attr_done[c->device] = true;
    }
    k_encode_intt_rows<<<dim3((unsigned)count, (unsigned)c->L_R), 256, ENC_ROWS_SMEM, c->stream>>>(c->d_params, d_ring, d_eidx, d_plain);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  {
    LaunchScope ls(c, "k_encode_intt");
    DISPATCH_LOGN(c->logN, { int rc = set_smem_attrs<LG, LV>(); if (rc) return rc;
                             k_encode_intt<LG, LV><<<grid, th, sm, c->stream>>>(c->d_params, d_ring, d_eidx, d_plain); });
  }
  CUDA_TRY(cudaGetLastError());
  if (split > 1)   // plaintexts are [count][L_R] polynomials: modulus index = polynomial index % L_R
    return launch_intt_finish(c, d_plain, count * c->L_R, c->d_modq, c->d_invN_q, c->d_invNw_q, (uint32_t)c->L_R, 0xFFFFFFFFu);
  return RSG_OK;
}
// is_signed: d_plain holds int64 sums of centred plaintexts (k_centre_add) instead of residues mod t
static int launch_lift_ntt(rsg_context *c, const uint64_t *d_plain, size_t count, uint64_t *d_pntt, bool is_signed = false) {
  if (!count) return RSG_OK;
  const unsigned th = ntt_threads(c->logN);
  const size_t sm = ntt_smem(c->logN);
  const unsigned split = c->logN > 14 ? 2 : 1;
  c->st_fwd_polys += count * c->L_R * c->L_E;
  // grid.x = (term * L_E + limb) * split: up to 2^31-1
  dim3 grid((unsigned)(count * c->L_E * split), (unsigned)c->L_R);
  bool lazy = true;   // correction-free butterflies need (4 * log2 N + 1) * Q_l < 2^64
  for (uint64_t p : c->Q) lazy = lazy && p < (1ull << 58);
  LaunchScope ls(c, "k_lift_fwd_ntt");
  if (c->f64_ntt && c->ntt_mode != 1 && c->logN == 14 && c->ntt_cluster) {
    const dim3 cgrid((unsigned)(count * c->L_E * NTT_CL), (unsigned)c->L_R);
    if (c->lift_smallq) {
      if (is_signed) k_lift_fwd_ntt_f64_cl<true, true><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
      else k_lift_fwd_ntt_f64_cl<false, true><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
    } else if (is_signed) k_lift_fwd_ntt_f64_cl<true, false><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
    else k_lift_fwd_ntt_f64_cl<false, false><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  if (c->f64_ntt && c->ntt_mode != 1) {
    DISPATCH_LOGN(c->logN, { int rc = set_smem_attrs<LG, LV>(); if (rc) return rc;
                             if (c->lift_smallq) {
                               if (is_signed) k_lift_fwd_ntt_f64<LG, LV, true, true><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, d_pntt);
                               else k_lift_fwd_ntt_f64<LG, LV, false, true><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, d_pntt);
                             } else if (is_signed) k_lift_fwd_ntt_f64<LG, LV, true><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, d_pntt);
                             else k_lift_fwd_ntt_f64<LG, LV, false><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, d_pntt); });
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  if (lazy && c->ntt_cluster && (c->logN == 15 || c->logN == 14)) {   // integer transform on clusters (kernels.cuh)
    const size_t smc = 4 * (1024 + 64) * 8;
    const unsigned cs = c->logN == 15 ? 8 : 4;
    const dim3 cgrid((unsigned)(count * c->L_E * cs), (unsigned)c->L_R);
    if (c->logN == 15) {
      if (is_signed) k_lift_fwd_ntt_int_cl<5, true><<<cgrid, 128, smc, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
      else k_lift_fwd_ntt_int_cl<5, false><<<cgrid, 128, smc, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
    } else {
      if (is_signed) k_lift_fwd_ntt_int_cl<4, true><<<cgrid, 128, smc, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
      else k_lift_fwd_ntt_int_cl<4, false><<<cgrid, 128, smc, c->stream>>>(c->d_params, d_plain, d_pntt, nullptr);
    }
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  DISPATCH_LOGN(c->logN, { int rc = set_smem_attrs<LG, LV>(); if (rc) return rc;
                           if (lazy) k_lift_fwd_ntt<LG, LV, true><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, is_signed, d_pntt);
                           else k_lift_fwd_ntt<LG, LV, false><<<grid, th, sm, c->stream>>>(c->d_params, d_plain, is_signed, d_pntt); });
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

extern "C" int rsg_batch_encode(rsg_context *c, const uint64_t *d_ring, size_t count, uint64_t *d_plain) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return launch_encode(c, d_ring, nullptr, count, d_plain);
}
extern "C" int rsg_plain_to_ntt(rsg_context *c, const uint64_t *d_plain, size_t count, uint64_t *d_pntt) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return launch_lift_ntt(c, d_plain, count, d_pntt);
}
static int ntt_dev(rsg_context *c, uint64_t *d, size_t batch, int which, size_t idx, int inverse);
extern "C" int rsg_ntt(rsg_context *c, uint64_t *d, size_t batch, int which, size_t idx, int inverse) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  if ((which == 0 && idx >= c->L_E) || (which == 1 && idx >= c->L_R) || which < 0 || which > 1) return fail(RSG_ERR_ARG, "bad modulus index");
  if (!batch) return RSG_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return ntt_dev(c, d, batch, which, idx, inverse);
}
// `batch` contiguous N_E-point polynomials, in place, canonical in and out; which = 0: mod Q_idx, 1: mod q_idx.  Caller holds c->mu.
static int ntt_dev(rsg_context *c, uint64_t *d, size_t batch, int which, size_t idx, int inverse) {
  const uint64_t p = which == 0 ? c->Q[idx] : c->q[idx];
  const Twiddle *tab = which == 0 ? (inverse ? c->hp.invQ[idx] : c->hp.fwdQ[idx]) : (inverse ? c->hp.invq[idx] : c->hp.fwdq[idx]);
  const Twiddle invn = which == 0 ? c->hp.invN_Q[idx] : c->hp.invN_q[idx];
  const unsigned th = ntt_threads(c->logN);
  const size_t sm = ntt_smem(c->logN);
  const unsigned split = c->logN > 14 ? 2 : 1;
  const uint64_t *src = d;
  if (split > 1 && !inverse) {   // the split forward transform is out of place: stage the input
    int rc = ensure(c, &c->d_plain, &c->cap_plain, batch * c->N_E);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->d_plain, d, batch * c->N_E * 8, cudaMemcpyDeviceToDevice, c->stream));
    src = c->d_plain;
  }
  {
    LaunchScope ls(c, inverse ? "k_ntt_inv" : "k_ntt_fwd");
    DISPATCH_LOGN(c->logN, {
      int rc = set_smem_attrs<LG, LV>();
      if (rc) return rc;
      if (inverse) k_ntt<LG, LV, true><<<(unsigned)batch * split, th, sm, c->stream>>>(src, d, tab, p, invn);
      else k_ntt<LG, LV, false><<<(unsigned)batch * split, th, sm, c->stream>>>(src, d, tab, p, invn);
    });
  }
  CUDA_TRY(cudaGetLastError());
  if (split > 1 && inverse)
    return launch_intt_finish(c, d, batch, which == 0 ? c->d_modQ : c->d_modq, which == 0 ? c->d_invN_Q : c->d_invN_q,
                              which == 0 ? c->d_invNw_Q : c->d_invNw_q, 1, (uint32_t)idx);
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// lincomb
// d_term / d_pidx: device arrays of n_terms entries (already uploaded).
static int launch_lincomb(rsg_context *c, const uint64_t *d_crs, const uint32_t *d_term, const uint32_t *d_pidx, size_t n_terms,
                          const uint64_t *d_pntt, uint64_t *d_out) {
  int rc;
  const unsigned th = (unsigned)std::min<size_t>(c->lin_threads > 0 ? c->lin_threads : 256, c->N_E / 2);
  const unsigned gx = (unsigned)(c->N_E / 2 / th), gy = (unsigned)(c->L_R * c->L_E);
  // split the term range so that the grid has >= ~8 blocks per SM; each split streams >= 8 terms
  const unsigned base_blocks = gx * gy;
  unsigned splits = std::max(1u, (148u * 8 + base_blocks - 1) / base_blocks);
  if (c->lin_splits > 0) splits = (unsigned)c->lin_splits;
  splits = (unsigned)std::min<size_t>(splits, (n_terms + 7) / 8);
  splits = std::max(1u, splits);
  const unsigned tps = (unsigned)((n_terms + splits - 1) / splits);
  splits = (unsigned)((n_terms + tps - 1) / tps);
  uint64_t *d_partial = d_out;
  if (splits > 1) {
    if ((rc = ensure(c, &c->d_partial, &c->cap_partial, (size_t)splits * c->enc_words()))) return rc;
    d_partial = c->d_partial;
  }
  c->st_lin_terms += n_terms;
  c->st_lin_launches++;
  {
    LaunchScope ls(c, "k_crs_lincomb");
    if (c->lin_unroll == 8)
      k_crs_lincomb<8><<<dim3(gx, gy, splits), th, 0, c->stream>>>(c->d_params, d_crs, d_term, d_pidx, (uint32_t)n_terms, tps, d_pntt, d_partial);
    else if (c->lin_unroll == 1)
      k_crs_lincomb<1><<<dim3(gx, gy, splits), th, 0, c->stream>>>(c->d_params, d_crs, d_term, d_pidx, (uint32_t)n_terms, tps, d_pntt, d_partial);
    else if (c->lin_unroll == 4)
      k_crs_lincomb<4><<<dim3(gx, gy, splits), th, 0, c->stream>>>(c->d_params, d_crs, d_term, d_pidx, (uint32_t)n_terms, tps, d_pntt, d_partial);
    else if (c->lin_unroll == 3)
      k_crs_lincomb<3><<<dim3(gx, gy, splits), th, 0, c->stream>>>(c->d_params, d_crs, d_term, d_pidx, (uint32_t)n_terms, tps, d_pntt, d_partial);
    else
      k_crs_lincomb<2><<<dim3(gx, gy, splits), th, 0, c->stream>>>(c->d_params, d_crs, d_term, d_pidx, (uint32_t)n_terms, tps, d_pntt, d_partial);
  }
  CUDA_TRY(cudaGetLastError());
  if (splits > 1) {
    LaunchScope ls(c, "k_enc_sum");
    const size_t pairs = c->enc_words() / 2;
    k_enc_sum<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, d_partial, splits, 1, d_out);
    CUDA_TRY(cudaGetLastError());
  }
  return RSG_OK;
}

extern "C" int rsg_crs_lincomb(rsg_context *c, const uint64_t *d_crs, const uint32_t *h_term, const uint32_t *h_pidx, size_t n_terms,
                               const uint64_t *d_pntt, uint64_t *d_out) {
  if (!c) return fail(RSG_ERR_STATE, "context not set");
  if (!n_terms) return fail(RSG_ERR_ARG, "empty term list");
  std::lock_guard<std::mutex> g(c->mu);
  int rc;
  if ((rc = ensure(c, &c->d_term, &c->cap_term, std::max<size_t>(n_terms, 1024)))) return rc;
  if ((rc = ensure(c, &c->d_pidx, &c->cap_pidx, std::max<size_t>(n_terms, 1024)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(c->d_term, h_term, n_terms * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->d_pidx, h_pidx, n_terms * 4, cudaMemcpyHostToDevice, c->stream));
  return launch_lincomb(c, d_crs, c->d_term, c->d_pidx, n_terms, d_pntt, d_out);
}

// One linear combination over an arbitrary term list: the engine behind rsg_inner_product and rsg_groth16_prove.
struct TermSpec {
  uint32_t crs_idx;            // which encoding of the arena
  const uint64_t *ring_base;   // vector the coefficient lives in (nullptr for RSG_TERM_ONE)
  uint32_t elem;               // element index inside ring_base
};
// d_probe_flags (nullable): receives the k_probe candidate flags, layout [L_R][terms.size()].
static int lincomb_terms(rsg_context *c, const uint64_t *d_crs, const std::vector<TermSpec> &terms, uint64_t *d_out,
                         uint8_t *d_probe_flags = nullptr) {
  int rc;
  if (d_probe_flags && !terms.empty()) CUDA_TRY(cudaMemsetAsync(c->d_probe_carry, 0, MAX_LR * 8, c->stream));
  if (terms.empty()) {
    CUDA_TRY(cudaMemsetAsync(d_out, 0, c->enc_words() * 8, c->stream));
    return RSG_OK;
  }
  const size_t per_general = c->L_R * c->L_E * c->N_E;
  const size_t max_general = std::max<size_t>(1, c->pntt_budget_words / per_general);
  // cut into chunks holding <= max_general GENERAL terms each
  struct Chunk { size_t t0, t1, g0, g1; };
  std::vector<Chunk> chunks;
  std::vector<uint32_t> term(terms.size()), pidx(terms.size()), eidx;
  {
    size_t t0 = 0, g0 = 0, g = 0;
    for (size_t t = 0; t < terms.size(); t++) {
      term[t] = terms[t].crs_idx;
      if (terms[t].ring_base) {
        if (g - g0 == max_general) { chunks.push_back({t0, t, g0, g}); t0 = t; g0 = g; }
        pidx[t] = (uint32_t)(g - g0);
        eidx.push_back(terms[t].elem);
        g++;
      } else {
        pidx[t] = 0xFFFFFFFFu;
      }
    }
    chunks.push_back({t0, terms.size(), g0, g});
  }
  const size_t G = eidx.size();
  c->st_lin_plain += G;
  if ((rc = ensure(c, &c->d_term, &c->cap_term, std::max<size_t>(terms.size(), 1024)))) return rc;
  if ((rc = ensure(c, &c->d_pidx, &c->cap_pidx, std::max<size_t>(terms.size(), 1024)))) return rc;
  if ((rc = ensure(c, &c->d_eidx, &c->cap_eidx, std::max<size_t>(G, 1024)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(c->d_term, term.data(), term.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(c->d_pidx, pidx.data(), pidx.size() * 4, cudaMemcpyHostToDevice, c->stream));
  if (G) CUDA_TRY(cudaMemcpyAsync(c->d_eidx, eidx.data(), G * 4, cudaMemcpyHostToDevice, c->stream));
  const size_t max_g = std::min(G, max_general);
  if (max_g) {
    if ((rc = ensure(c, &c->d_plain, &c->cap_plain, max_g * c->L_R * c->N_E))) return rc;
    if ((rc = ensure(c, &c->d_pntt, &c->cap_pntt, max_g * per_general))) return rc;
  }
  uint64_t *chunk_out = d_out;
  if (chunks.size() > 1) {
    if ((rc = ensure(c, &c->d_chunk, &c->cap_chunk, chunks.size() * c->enc_words()))) return rc;
    chunk_out = c->d_chunk;
  }
  // pageable host staging above is consumed synchronously by cudaMemcpyAsync, so the vectors may die with this frame
  for (size_t k = 0; k < chunks.size(); k++) {
    const Chunk &ch = chunks[k];
    // encode runs of consecutive general terms that share a source vector
    size_t g = ch.g0;
    size_t t = ch.t0;
    while (g < ch.g1) {
      while (!terms[t].ring_base) t++;
      const uint64_t *base = terms[t].ring_base;
      size_t run = 0, tt = t;
      while (tt < ch.t1 && g + run < ch.g1) {
        if (terms[tt].ring_base) {
          if (terms[tt].ring_base != base) break;
          run++;
        }
        tt++;
      }
      if ((rc = launch_encode(c, base, c->d_eidx + g, run, c->d_plain + (g - ch.g0) * c->L_R * c->N_E))) return rc;
      g += run;
      t = tt;
    }
    if ((rc = launch_lift_ntt(c, c->d_plain, ch.g1 - ch.g0, c->d_pntt))) return rc;
    if ((rc = launch_lincomb(c, d_crs, c->d_term + ch.t0, c->d_pidx + ch.t0, ch.t1 - ch.t0, c->d_pntt,
                             chunk_out + (chunks.size() > 1 ? k * c->enc_words() : 0))))
      return rc;
    if (d_probe_flags) {
      LaunchScope ls(c, "k_probe");
      k_probe<<<(unsigned)c->L_R, 256, 0, c->stream>>>(c->d_params, d_crs, c->d_term + ch.t0, c->d_pidx + ch.t0, (uint32_t)(ch.t1 - ch.t0),
                                                        c->d_pntt, c->d_probe_carry, d_probe_flags + ch.t0, (uint32_t)terms.size(), 0u);
      CUDA_TRY(cudaGetLastError());
    }
  }
  if (chunks.size() > 1) {
    LaunchScope ls(c, "k_enc_sum");
    const size_t pairs = c->enc_words() / 2;
    k_enc_sum<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, chunk_out, (uint32_t)chunks.size(), 1, d_out);
    CUDA_TRY(cudaGetLastError());
  }
  return RSG_OK;
}

// Two inner products over the SAME CRS vector in one pass (groth16.tcc:89-103: <s_pows, a_io> + <s_pows, a_mid>):
//   sum_i ct_i (.) NTT(lift(x_i)) + sum_i ct_i (.) NTT(lift(y_i)) = sum_i ct_i (.) NTT(lift(x_i) + lift(y_i))   (mod Q_l, exactly),
// so each CRS element is streamed once and transformed once instead of twice.  The two batch encodings stay separate
// (the centred lift is not additive).  d_out receives X + Y.  The transparent-prefix probes of the two parts are kept apart:
// their NTT-domain probe words come from k_probe_eval, the flags go to d_flagsX / d_flagsY ([L_R][|X|] and [L_R][|Y|]).
// Returns 1 (nothing launched) when the lists do not qualify: a non-general term, mixed source vectors, unsorted CRS indices.
static int lincomb_merged(rsg_context *c, const uint64_t *d_crs, const std::vector<TermSpec> &X, const std::vector<TermSpec> &Y,
                          uint64_t *d_out, uint8_t *d_flagsX, uint8_t *d_flagsY) {
  int rc;
  if (X.empty() || Y.empty()) return 1;
  for (const std::vector<TermSpec> *v : {&X, &Y})
    for (size_t i = 0; i < v->size(); i++) {
      if (!(*v)[i].ring_base || (*v)[i].ring_base != (*v)[0].ring_base) return 1;
      if (i && (*v)[i].crs_idx <= (*v)[i - 1].crs_idx) return 1;
    }
  struct M { uint32_t crs, x, y; };
  const uint32_t NONE = 0xFFFFFFFFu;
  std::vector<M> mt;
  mt.reserve(X.size() + Y.size());
  for (size_t a = 0, b = 0; a < X.size() || b < Y.size();) {
    if (b >= Y.size() || (a < X.size() && X[a].crs_idx < Y[b].crs_idx)) { mt.push_back({X[a].crs_idx, (uint32_t)a, NONE}); a++; }
    else if (a >= X.size() || Y[b].crs_idx < X[a].crs_idx) { mt.push_back({Y[b].crs_idx, NONE, (uint32_t)b}); b++; }
    else { mt.push_back({X[a].crs_idx, (uint32_t)a, (uint32_t)b}); a++; b++; }
  }
  const size_t per_general = c->L_R * c->L_E * c->N_E, poly = c->L_R * c->N_E;
  const size_t max_m = std::max<size_t>(1, c->pntt_budget_words / per_general);
  const size_t n_chunks = (mt.size() + max_m - 1) / max_m, cm = std::min(mt.size(), max_m);
  if ((rc = ensure(c, &c->d_plain, &c->cap_plain, 3 * cm * poly))) return rc;   // [X parts | Y parts | centred sums]
  if ((rc = ensure(c, &c->d_pntt, &c->cap_pntt, cm * per_general))) return rc;
  if ((rc = ensure(c, &c->d_pval, &c->cap_pval, std::max<size_t>(2 * cm * c->L_R, 1024)))) return rc;
  // one upload: per chunk [term M | pidx M | pair 2M | eidxX | eidxY | crsX | slotX | crsY | slotY]
  std::vector<uint32_t> buf;
  struct Off { size_t term, pidx, pair, ex, ey, cx, sx, cy, sy, nx, ny, m, x0, y0; };
  std::vector<Off> offs;
  for (size_t k = 0; k < n_chunks; k++) {
    const size_t m0 = k * max_m, m1 = std::min(mt.size(), m0 + max_m), Mk = m1 - m0;
    std::vector<uint32_t> ex, ey, cx, cy, sx, sy, pair(2 * Mk);
    size_t x0 = X.size(), y0 = Y.size();
    for (size_t m = m0; m < m1; m++) {
      if (mt[m].x != NONE) { x0 = std::min<size_t>(x0, mt[m].x); ex.push_back(X[mt[m].x].elem); cx.push_back(mt[m].crs); }
      if (mt[m].y != NONE) { y0 = std::min<size_t>(y0, mt[m].y); ey.push_back(Y[mt[m].y].elem); cy.push_back(mt[m].crs); }
    }
    size_t ax = 0, ay = 0;
    for (size_t m = m0; m < m1; m++) {
      uint32_t s0 = NONE, s1 = NONE;
      if (mt[m].x != NONE) { s0 = (uint32_t)ax; sx.push_back(s0); ax++; }
      if (mt[m].y != NONE) { s1 = (uint32_t)(ex.size() + ay); sy.push_back(s1); ay++; }
      if (s0 == NONE) std::swap(s0, s1);
      pair[2 * (m - m0)] = s0;
      pair[2 * (m - m0) + 1] = s1;
    }
    Off o;
    o.m = Mk; o.nx = ex.size(); o.ny = ey.size(); o.x0 = x0; o.y0 = y0;
    o.term = buf.size(); for (size_t m = m0; m < m1; m++) buf.push_back(mt[m].crs);
    o.pidx = buf.size(); for (size_t m = 0; m < Mk; m++) buf.push_back((uint32_t)m);
    o.pair = buf.size(); buf.insert(buf.end(), pair.begin(), pair.end());
    o.ex = buf.size(); buf.insert(buf.end(), ex.begin(), ex.end());
    o.ey = buf.size(); buf.insert(buf.end(), ey.begin(), ey.end());
    o.cx = buf.size(); buf.insert(buf.end(), cx.begin(), cx.end());
    o.sx = buf.size(); buf.insert(buf.end(), sx.begin(), sx.end());
    o.cy = buf.size(); buf.insert(buf.end(), cy.begin(), cy.end());
    o.sy = buf.size(); buf.insert(buf.end(), sy.begin(), sy.end());
    offs.push_back(o);
  }
  c->st_merged++;
  c->st_lin_plain += mt.size();
  if ((rc = ensure(c, &c->d_mrg, &c->cap_mrg, std::max<size_t>(buf.size(), 4096)))) return rc;
  CUDA_TRY(cudaMemcpyAsync(c->d_mrg, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(c->d_probe_carry, 0, 2 * MAX_LR * 8, c->stream));
  uint64_t *chunk_out = d_out;
  if (n_chunks > 1) {
    if ((rc = ensure(c, &c->d_chunk, &c->cap_chunk, n_chunks * c->enc_words()))) return rc;
    chunk_out = c->d_chunk;
  }
  for (size_t k = 0; k < n_chunks; k++) {
    const Off &o = offs[k];
    const uint32_t *d = c->d_mrg;
    if ((rc = launch_encode(c, X[0].ring_base, d + o.ex, o.nx, c->d_plain))) return rc;
    if ((rc = launch_encode(c, Y[0].ring_base, d + o.ey, o.ny, c->d_plain + o.nx * poly))) return rc;
    uint64_t *comb = c->d_plain + (o.nx + o.ny) * poly;
    {
      LaunchScope ls(c, "k_centre_add");
      k_centre_add<<<dim3((unsigned)o.m, (unsigned)c->L_R), 256, 0, c->stream>>>(c->d_params, c->d_plain, d + o.pair, comb);
    }
    if ((rc = launch_lift_ntt(c, comb, o.m, c->d_pntt, true))) return rc;
    if ((rc = launch_lincomb(c, d_crs, d + o.term, d + o.pidx, o.m, c->d_pntt, chunk_out + (n_chunks > 1 ? k * c->enc_words() : 0))))
      return rc;
    {
      LaunchScope ls(c, "k_probe_eval");
      k_probe_eval<<<dim3((unsigned)(o.nx + o.ny), (unsigned)c->L_R), 256, 0, c->stream>>>(c->d_params, c->d_plain, c->d_psi_pow, c->d_pval);
    }
    if (o.nx) {
      LaunchScope ls(c, "k_probe");
      k_probe<<<(unsigned)c->L_R, 256, 0, c->stream>>>(c->d_params, d_crs, d + o.cx, d + o.sx, (uint32_t)o.nx, c->d_pval, c->d_probe_carry,
                                                        d_flagsX + o.x0, (uint32_t)X.size(), (uint32_t)c->L_R);
    }
    if (o.ny) {
      LaunchScope ls(c, "k_probe");
      k_probe<<<(unsigned)c->L_R, 256, 0, c->stream>>>(c->d_params, d_crs, d + o.cy, d + o.sy, (uint32_t)o.ny, c->d_pval,
                                                        c->d_probe_carry + MAX_LR, d_flagsY + o.y0, (uint32_t)Y.size(), (uint32_t)c->L_R);
    }
    CUDA_TRY(cudaGetLastError());
  }
  if (n_chunks > 1) {
    LaunchScope ls(c, "k_enc_sum");
    const size_t pairs = c->enc_words() / 2;
    k_enc_sum<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, chunk_out, (uint32_t)n_chunks, 1, d_out);
    CUDA_TRY(cudaGetLastError());
  }
  return RSG_OK;
}

// acc += other with the reference's transparent-result rule (seal_ring.tcc:493-504): a ring limb whose c1 sums to zero
// becomes the empty zero ciphertext (all-zero words).  Entirely on the device, no host round trip.
static int transparent_fix(rsg_context *c, uint64_t *d_acc);
static int enc_add_fix(rsg_context *c, uint64_t *d_acc, const uint64_t *d_other) {
  const size_t pairs = c->enc_words() / 2;
  {
    LaunchScope ls(c, "k_enc_add");
    k_enc_add<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, d_acc, d_other, d_acc);
  }
  return transparent_fix(c, d_acc);
}
// the rule alone, applied to a sum that is already formed
static int transparent_fix(rsg_context *c, uint64_t *d_acc) {
  CUDA_TRY(cudaMemsetAsync(c->d_nz, 0, MAX_LR * 4, c->stream));
  const dim3 grid((unsigned)std::min<size_t>(64, (c->L_E * c->N_E + 255) / 256), (unsigned)c->L_R);
  {
    LaunchScope ls(c, "k_c1_nonzero");
    k_c1_nonzero<<<grid, 256, 0, c->stream>>>(c->d_params, d_acc, c->d_nz);
  }
  {
    LaunchScope ls(c, "k_zero_transparent");
    k_zero_transparent<<<grid, 256, 0, c->stream>>>(c->d_params, d_acc, c->d_nz);
  }
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

// Exact resolution of flagged prefixes for ONE inner product (rare: needs a structured CRS, e.g. the seeded test CRS in
// which every ciphertext shares its uniform polynomial).  flags: host copy of the probe flags [L_R][T].
// Reference sequence per ring limb (seal_ring.tcc:415-431,479-507): res = tmp_0; then res += tmp_t and, if the sum's c1 is
// identically zero, res = empty zero ciphertext, from which the next += restarts.  The probe's running sum is zero at a
// restart, so the candidate set stays valid after it.
static int inner_product_exact(rsg_context *c, const uint64_t *d_crs, const std::vector<TermSpec> &terms, const uint8_t *flags,
                               uint64_t *d_out) {
  int rc;
  const size_t T = terms.size(), L_R = c->L_R, ct_words = 2 * c->L_E * c->N_E;
  c->exact_fallbacks++;
  if ((rc = ensure(c, &c->d_exact, &c->cap_exact, c->enc_words()))) return rc;
  std::vector<size_t> start(L_R, 0);
  const dim3 grid((unsigned)std::min<size_t>(64, (c->L_E * c->N_E + 255) / 256), (unsigned)L_R);
  for (size_t j = 0; j < L_R; j++)
    for (size_t t = 0; t < T; t++) {
      if (!flags[j * T + t] || t < start[j]) continue;
      std::vector<TermSpec> sub(terms.begin() + start[j], terms.begin() + t + 1);
      if ((rc = lincomb_terms(c, d_crs, sub, c->d_exact))) return rc;
      CUDA_TRY(cudaMemsetAsync(c->d_nz, 0, MAX_LR * 4, c->stream));
      k_c1_nonzero<<<grid, 256, 0, c->stream>>>(c->d_params, c->d_exact, c->d_nz);
      c->launches++;
      uint32_t nz[MAX_LR];
      CUDA_TRY(cudaMemcpyAsync(nz, c->d_nz, L_R * 4, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
      if (!nz[j]) start[j] = t + 1;   // transparent prefix: the reference drops it
    }
  for (size_t j = 0; j < L_R; j++) {
    if (start[j] == 0) continue;      // untouched limb: d_out already holds the plain sum
    uint64_t *dst = d_out + j * ct_words;
    if (start[j] >= T) {
      CUDA_TRY(cudaMemsetAsync(dst, 0, ct_words * 8, c->stream));
    } else {
      std::vector<TermSpec> sub(terms.begin() + start[j], terms.end());
      if ((rc = lincomb_terms(c, d_crs, sub, c->d_exact))) return rc;
      CUDA_TRY(cudaMemcpyAsync(dst, c->d_exact + j * ct_words, ct_words * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
  }
  return RSG_OK;
}

// One EncodingElem::inner_product with the reference's exact semantics; synchronises the stream.
static int inner_product_terms(rsg_context *c, const uint64_t *d_crs, const std::vector<TermSpec> &terms, uint64_t *d_out) {
  int rc;
  const size_t T = terms.size(), L_R = c->L_R;
  if (T) {
    if ((rc = ensure(c, &c->d_probe, &c->cap_probe, std::max<size_t>(L_R * T, 4096)))) return rc;
  }
  if ((rc = lincomb_terms(c, d_crs, terms, d_out, T ? c->d_probe : nullptr))) return rc;
  if (!T) return RSG_OK;
  std::vector<uint8_t> flags(L_R * T);
  CUDA_TRY(cudaMemcpyAsync(flags.data(), c->d_probe, L_R * T, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  bool any = false;
  for (uint8_t f : flags) any |= f != 0;
  if (any && (rc = inner_product_exact(c, d_crs, terms, flags.data(), d_out))) return rc;
  return RSG_OK;
}

static int inner_product_impl(rsg_context *c, const rsg_crs *crs, size_t crs_first, const uint32_t *h_crs_idx,
                              const rsg_ringvec *coeffs, size_t coeff_first, const uint32_t *h_coeff_idx, size_t count,
                              const uint8_t *h_tags, uint64_t *h_out, uint64_t *d_out, size_t *n_used) {
  if (!c || !crs || !coeffs || !h_tags) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  std::vector<TermSpec> terms;
  for (size_t i = 0; i < count; i++) {
    if (h_tags[i] == RSG_TERM_SKIP) continue;
    const size_t ci = h_crs_idx ? h_crs_idx[i] : crs_first + i, ri = h_coeff_idx ? h_coeff_idx[i] : coeff_first + i;
    if (ci >= crs->n || (h_tags[i] != RSG_TERM_ONE && ri >= coeffs->n)) return fail(RSG_ERR_ARG, "term index out of range");
    terms.push_back({(uint32_t)ci, h_tags[i] == RSG_TERM_ONE ? nullptr : coeffs->d, (uint32_t)ri});
  }
  if (n_used) *n_used = terms.size();
  int rc;
  uint64_t *out = d_out;
  if (!out) {
    if ((rc = ensure(c, &c->d_out_scratch, &c->cap_out_scratch, 3 * c->enc_words()))) return rc;
    out = c->d_out_scratch;
  }
  if ((rc = inner_product_terms(c, crs->d, terms, out))) return rc;
  if (h_out) CUDA_TRY(cudaMemcpyAsync(h_out, out, c->enc_words() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}
extern "C" int rsg_inner_product(rsg_context *c, const rsg_crs *crs, size_t crs_first, const rsg_ringvec *coeffs, size_t coeff_first,
                                 size_t count, const uint8_t *h_tags, uint64_t *h_out, uint64_t *d_out, size_t *n_used) {
  RSG_TRACE_CALL();
  if (!crs || !coeffs) return fail(RSG_ERR_ARG, "null argument");
  if (crs_first + count > crs->n || coeff_first + count > coeffs->n) return fail(RSG_ERR_ARG, "range");
  return inner_product_impl(c, crs, crs_first, nullptr, coeffs, coeff_first, nullptr, count, h_tags, h_out, d_out, n_used);
}
extern "C" int rsg_inner_product_idx(rsg_context *c, const rsg_crs *crs, const uint32_t *h_crs_idx, const rsg_ringvec *coeffs,
                                     const uint32_t *h_coeff_idx, size_t count, const uint8_t *h_tags, uint64_t *h_out,
                                     uint64_t *d_out, size_t *n_used) {
  RSG_TRACE_CALL();
  if (!h_crs_idx || !h_coeff_idx) return fail(RSG_ERR_ARG, "null index list");
  return inner_product_impl(c, crs, 0, h_crs_idx, coeffs, 0, h_coeff_idx, count, h_tags, h_out, d_out, n_used);
}

extern "C" int rsg_enc_sum_strided(rsg_context *c, const uint64_t *d_parts, size_t parts, size_t n_enc, size_t part_stride_words,
                                   uint64_t *d_out) {
  if (!c || !d_parts || !d_out || !parts || !n_enc) return fail(RSG_ERR_ARG, "null argument");
  if (part_stride_words && (part_stride_words < n_enc * c->enc_words() || (part_stride_words & 1))) return fail(RSG_ERR_ARG, "part stride");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  LaunchScope ls(c, "k_enc_sum");
  const size_t pairs = n_enc * c->enc_words() / 2;
  k_enc_sum<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, d_parts, (uint32_t)parts, (uint32_t)n_enc, d_out,
                                                                     part_stride_words);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_enc_sum(rsg_context *c, const uint64_t *d_parts, size_t parts, size_t n_enc, uint64_t *d_out) {
  return rsg_enc_sum_strided(c, d_parts, parts, n_enc, 0, d_out);
}
extern "C" int rsg_enc_add(rsg_context *c, uint64_t *d_acc, const uint64_t *d_other) {
  RSG_TRACE_CALL();
  if (!c || !d_acc || !d_other) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return enc_add_fix(c, d_acc, d_other);
}
extern "C" int rsg_crs_copy(rsg_crs *dst, size_t dst_first, const rsg_crs *src, size_t src_first, size_t count) {
  RSG_TRACE_CALL();
  if (!dst || !src || dst->ctx != src->ctx) return fail(RSG_ERR_ARG, "null or foreign arena");
  if (dst_first + count > dst->n || src_first + count > src->n) return fail(RSG_ERR_ARG, "CRS range");
  rsg_context *c = dst->ctx;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  CUDA_TRY(cudaMemcpyAsync(dst->d + dst_first * c->enc_words(), src->d + src_first * c->enc_words(), count * c->enc_words() * 8,
                           cudaMemcpyDeviceToDevice, c->stream));
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// witness map
// Per-prime constants for the domain {0..n-1}: Z, V^-1 (column j = coefficients of the j-th Lagrange basis
// polynomial, i.e. what interpolate() accumulates for y = e_j, polynomials.tcc:26-41) and the Toeplitz matrix of
// rev(Z)^-1 mod x^(n-1) that turns "top half of the dividend" into the quotient of the long division by Z.
static const rsg_host::NttTables &host_ntt(rsg_context *c, size_t j) {
  if (c->h_ntt.size() != c->L_R) c->h_ntt.assign(c->L_R, rsg_host::NttTables());
  if (!c->h_ntt[j].p) c->h_ntt[j].set(c->q[j], c->h_fwdq[j]);
  return c->h_ntt[j];
}
static int get_witness_tables(rsg_context *c, size_t n, WitnessTables **out) {
  auto it = c->wit.find(n);
  if (it != c->wit.end()) { *out = &it->second; return RSG_OK; }
  if (n < 1) return fail(RSG_ERR_ARG, "n must be >= 1");
  for (uint64_t p : c->q)
    if (p <= 2 * n) return fail(RSG_ERR_ARG, "domain {0..n-1} is not an exceptional set for this modulus");
  WitnessTables wt;
  const size_t L_R = c->L_R, m = n > 0 ? n - 1 : 0;
  wt.h_Z.assign(L_R * (n + 1), 0);
  wt.h_u.assign(L_R * std::max<size_t>(m, 1), 0);
  for (size_t j = 0; j < L_R; j++) {
    const uint64_t p = c->q[j];
    uint64_t *Z = wt.h_Z.data() + j * (n + 1);
    if (n >= 2048) {   // divide and conquer + Newton iteration over transform products: the same residues in O(n log^2 n)
      const rsg_host::NttTables &t = host_ntt(c, j);
      const rsg_host::Poly z = rsg_host::node_product(0, n, t);
      std::copy(z.begin(), z.end(), Z);
      if (m) {
        rsg_host::Poly rz(n + 1);
        for (size_t i = 0; i <= n; i++) rz[i] = z[n - i];
        const rsg_host::Poly u = rsg_host::series_inverse(rz, m, t);
        std::copy(u.begin(), u.begin() + m, wt.h_u.data() + j * m);
      }
      continue;
    }
    Z[0] = 1;
    for (size_t i = 0; i < n; i++) {   // multiply by (x - i)
      const uint64_t neg = (p - i % p) % p;
      for (size_t k = i + 2; k-- > 0;) {
        const uint64_t lower = k ? Z[k - 1] : 0;
        const uint64_t same = k <= i ? h_mulmod(Z[k], neg, p) : 0;
        Z[k] = (same + lower) % p;
      }
    }
    if (m) {
      // u = rev(Z)^-1 mod x^m, rev(Z)_i = Z[n - i] (rev(Z)_0 = 1)
      uint64_t *u = wt.h_u.data() + j * m;
      u[0] = 1;
      for (size_t i = 1; i < m; i++) {
        u128 acc = 0;
        for (size_t t = 1; t <= i; t++) acc += (u128)h_mulmod(Z[n - t], u[i - t], p);
        u[i] = (p - (uint64_t)(acc % p)) % p;
      }
    }
  }
  void *v = nullptr;
  CUDA_TRY(cudaMalloc(&v, wt.h_Z.size() * 8));
  wt.d_Z = (uint64_t *)v;
  CUDA_TRY(cudaMemcpy(wt.d_Z, wt.h_Z.data(), wt.h_Z.size() * 8, cudaMemcpyHostToDevice));
  auto ins = c->wit.emplace(n, std::move(wt));
  *out = &ins.first->second;
  return RSG_OK;
}

// Dense-path constants: V^-1 and the Toeplitz matrix of rev(Z)^-1 (O(n^2) words each; small n, or RSG_WITNESS=dense).
static int ensure_dense_tables(rsg_context *c, size_t n, WitnessTables *wt) {
  if (wt->d_Vinv) return RSG_OK;
  const size_t L_R = c->L_R, m = n - 1;
  std::vector<uint64_t> Vinv(L_R * n * n), T(L_R * std::max<size_t>(m * m, 1), 0);
  for (size_t j = 0; j < L_R; j++) {
    const uint64_t p = c->q[j];
    const uint64_t *Z = wt->h_Z.data() + j * (n + 1);
    // phi_x = Z'(x) = prod_{i != x} (x - i) = x! (n-1-x)! (-1)^(n-1-x)
    std::vector<uint64_t> fact(n + 1, 1);
    for (size_t i = 1; i <= n; i++) fact[i] = h_mulmod(fact[i - 1], i % p, p);
    uint64_t *V = Vinv.data() + j * n * n;
    std::vector<uint64_t> b(n);
    for (size_t x = 0; x < n; x++) {
      uint64_t phi = h_mulmod(fact[x], fact[n - 1 - x], p);
      if ((n - 1 - x) & 1) phi = (p - phi) % p;
      const uint64_t iphi = h_inv(phi, p);
      // synthetic division Z(x) / (x - x0): b[n-1] = 1, b[k-1] = Z[k] + x0 * b[k]
      b[n - 1] = 1;
      for (size_t k = n - 1; k > 0; k--) b[k - 1] = (Z[k] + h_mulmod(x % p, b[k], p)) % p;
      for (size_t k = 0; k < n; k++) V[k * n + x] = h_mulmod(b[k], iphi, p);
    }
    if (m) {
      const uint64_t *u = wt->h_u.data() + j * m;
      uint64_t *Tm = T.data() + j * m * m;
      for (size_t k = 0; k < m; k++)
        for (size_t col = k; col < m; col++) Tm[k * m + col] = u[col - k];
    }
  }
  void *v = nullptr;
  CUDA_TRY(cudaMalloc(&v, T.size() * 8));
  wt->d_T = (uint64_t *)v;
  CUDA_TRY(cudaMemcpy(wt->d_T, T.data(), T.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&v, Vinv.size() * 8));
  CUDA_TRY(cudaMemcpy(v, Vinv.data(), Vinv.size() * 8, cudaMemcpyHostToDevice));
  wt->d_Vinv = (uint64_t *)v;
  return RSG_OK;
}

// ---- quasi-linear path (witness_fast.cuh) ---------------------------------------------------------------------
// Transform size for n nodes: the power of two S >= n, doubled when more than WF_WC_MAX coefficients of a length-(2n-1)
// product would wrap around modulo x^S + 1.
static void wf_shape(size_t n, uint32_t *S, uint32_t *logS, uint32_t *wc) {
  uint32_t lg = 5;
  while (((size_t)1 << lg) < n) lg++;
  size_t s = (size_t)1 << lg;
  if (2 * n - 1 > s + WF_WC_MAX) { lg++; s <<= 1; }
  *S = (uint32_t)s;
  *logS = lg;
  *wc = 2 * n - 1 > s ? (uint32_t)(2 * n - 1 - s) : 0;
}
// the largest transform the witness kernels take: no 2S-th root of unity is guaranteed beyond N_E, buffers are laid out to 32768
static uint32_t wf_max_ts(const rsg_context *c) {
  uint32_t ts = (uint32_t)std::min<size_t>(c->N_E, 32768);
  if (c->wf_ts >= 64 && !(c->wf_ts & (c->wf_ts - 1))) ts = std::min(ts, c->wf_ts);
  return ts;
}
// blocked mode (witness_fast.cuh, k_interp_big / k_quotient_big): products assembled from blocks of TS/2 coefficients, n <= 2*TS
static bool wf_big(const rsg_context *c, size_t n) {
  uint32_t S, logS, wc;
  wf_shape(n, &S, &logS, &wc);
  const uint32_t ts = wf_max_ts(c);
  return S > ts && n <= 2 * (size_t)ts && n >= ts / 2 + WF_B;
}
static bool wf_supported(const rsg_context *c, size_t n) {
  if (n < 2) return false;
  uint32_t S, logS, wc;
  wf_shape(n, &S, &logS, &wc);
  if (S > wf_max_ts(c) && !wf_big(c, n)) return false;
  for (uint64_t p : c->q)
    if (p >= (1ull << 61)) return false;
  return true;
}
static bool wf_use(const rsg_context *c, size_t n) {
  if (c->witness_mode == 1 || !wf_supported(c, n)) return false;
  // measured on B200 (profiles/r1d_c5_sweep.json, N_R = 2^15): n = 256 dense 12.9 ms / quasi-linear 18.4 ms, n = 1024 185 / 50 ms
  return c->witness_mode == 2 || n >= 320;
}
// forward negacyclic transform on the host, the device's table order (h_tables): natural in, bit-reversed out
static void h_ntt_fwd(std::vector<uint64_t> &a, int lg, uint64_t p, const std::vector<uint64_t> &tw) {
  const size_t n = (size_t)1 << lg;
  for (int s = 0; s < lg; s++) {
    const size_t gap = n >> (s + 1);
    for (size_t blk = 0; blk < ((size_t)1 << s); blk++) {
      const uint64_t w = tw[((size_t)1 << s) + blk];
      for (size_t o = 0; o < gap; o++) {
        const size_t i = blk * 2 * gap + o;
        const uint64_t x = a[i], y = h_mulmod(a[i + gap], w, p);
        a[i] = (x + y) % p;
        a[i + gap] = (x + p - y) % p;
      }
    }
  }
}
static std::vector<uint64_t> h_polymul(const std::vector<uint64_t> &a, const std::vector<uint64_t> &b, uint64_t p) {
  std::vector<uint64_t> r(a.size() + b.size() - 1, 0);
  for (size_t i = 0; i < a.size(); i++) {
    if (!a[i]) continue;
    for (size_t k = 0; k < b.size(); k++) r[i + k] = (uint64_t)(((u128)a[i] * b[k] + r[i + k]) % p);
  }
  return r;
}
static int ensure_big_tables(rsg_context *c, size_t n, WitnessTables *wt);
static int ensure_fast_tables(rsg_context *c, size_t n, WitnessTables *wt) {
  if (wt->fast_ready) return RSG_OK;
  if (wf_big(c, n)) return ensure_big_tables(c, n, wt);
  FastTables &ft = wt->ft;
  memset(&ft, 0, sizeof(ft));
  uint32_t S, logS, wc;
  wf_shape(n, &S, &logS, &wc);
  ft.n = (uint32_t)n; ft.S = S; ft.logS = logS; ft.wc = wc;
  uint32_t levels = 0;
  for (size_t m = WF_B; m < n; m <<= 1) levels++;
  ft.levels = std::max<uint32_t>(levels, 1);
  const size_t L_R = c->L_R, npad = (n + WF_B - 1) / WF_B * WF_B, lu = n - 1;
  std::vector<Twiddle> invfact(L_R * n), pts(L_R * npad), Ghat(L_R * S), Phat(L_R * ft.levels * S, Twiddle{0, 0}), Vhat(L_R * S);
  std::vector<uint64_t> g_nat(L_R * n), Pnat(L_R * ft.levels * (S / 2 + 1), 0), v_nat(L_R * n, 0);
  for (size_t j = 0; j < L_R; j++) {
    const uint64_t p = c->q[j];
    const std::vector<uint64_t> &tw = c->h_fwdq[j];
    std::vector<uint64_t> fact(n + 1, 1), ifact(n + 1);
    for (size_t i = 1; i <= n; i++) fact[i] = h_mulmod(fact[i - 1], i % p, p);
    ifact[n] = h_inv(fact[n], p);
    for (size_t i = n; i > 0; i--) ifact[i - 1] = h_mulmod(ifact[i], i % p, p);
    for (size_t i = 0; i < n; i++) {
      invfact[j * n + i] = h_twiddle(ifact[i], p);
      g_nat[j * n + i] = (i & 1) ? (p - ifact[i]) % p : ifact[i];
    }
    for (size_t i = 0; i < npad; i++) pts[j * npad + i] = h_twiddle(i % p, p);
    const uint64_t invS = h_inv(S % p, p);
    ft.invS[j] = h_twiddle(invS, p);
    {
      std::vector<uint64_t> a(S, 0);
      for (size_t i = 0; i < n; i++) a[i] = g_nat[j * n + i];
      h_ntt_fwd(a, (int)logS, p, tw);
      for (size_t i = 0; i < S; i++) Ghat[j * S + i] = h_twiddle(h_mulmod(a[i], invS, p), p);
    }
    {
      std::vector<uint64_t> a(S, 0);
      for (size_t i = 0; i < lu; i++) a[i] = v_nat[j * n + i] = wt->h_u[j * std::max<size_t>(lu, 1) + i];
      h_ntt_fwd(a, (int)logS, p, tw);
      for (size_t i = 0; i < S; i++) Vhat[j * S + i] = h_twiddle(h_mulmod(a[i], invS, p), p);
    }
    // subproduct tree over aligned node ranges: tree[r] = prod_{x in [r*size, (r+1)*size)} (X - x)
    std::vector<std::vector<uint64_t>> tree(npad / WF_B);
    for (size_t r = 0; r < tree.size(); r++) {
      std::vector<uint64_t> f{1};
      for (size_t x = r * WF_B; x < (r + 1) * WF_B; x++) f = h_polymul(f, {(p - x % p) % p, 1}, p);
      tree[r] = std::move(f);
    }
    size_t lvl = 0;
    for (size_t m = WF_B; m < n; m <<= 1, lvl++) {
      // tree holds the products over ranges of m nodes; block b of this level needs tree[2b]
      const size_t two_m = 2 * m, nb_active = (n - m + two_m - 1) / two_m;
      int lg = 0;
      while (((size_t)1 << lg) < two_m) lg++;
      const uint64_t inv2m = h_inv(two_m % p, p);
      for (size_t b = 0; b < nb_active; b++) {
        std::vector<uint64_t> a(two_m, 0);
        const std::vector<uint64_t> &P = tree[2 * b];
        for (size_t i = 0; i <= m; i++) a[i] = P[i];
        h_ntt_fwd(a, lg, p, tw);
        Twiddle *dst = Phat.data() + (j * ft.levels + lvl) * S + b * two_m;
        for (size_t i = 0; i < two_m; i++) dst[i] = h_twiddle(h_mulmod(a[i], inv2m, p), p);
      }
      const std::vector<uint64_t> &Pl = tree[2 * (nb_active - 1)];
      std::copy(Pl.begin(), Pl.end(), Pnat.begin() + (j * ft.levels + lvl) * (S / 2 + 1));
      std::vector<std::vector<uint64_t>> next;
      for (size_t r = 0; 2 * r + 1 < tree.size(); r++) next.push_back(h_polymul(tree[2 * r], tree[2 * r + 1], p));
      tree.swap(next);
    }
  }
  int rc;
  Twiddle *dt;
  uint64_t *du;
  if ((rc = upload_vec(c, invfact, &dt))) return rc;
  ft.invfact = dt;
  if ((rc = upload_vec(c, pts, &dt))) return rc;
  ft.pts = dt;
  if ((rc = upload_vec(c, Ghat, &dt))) return rc;
  ft.Ghat = dt;
  if ((rc = upload_vec(c, Phat, &dt))) return rc;
  ft.Phat = dt;
  if ((rc = upload_vec(c, Vhat, &dt))) return rc;
  ft.Vhat = dt;
  if ((rc = upload_vec(c, g_nat, &du))) return rc;
  ft.g_nat = du;
  if ((rc = upload_vec(c, Pnat, &du))) return rc;
  ft.Pnat = du;
  if ((rc = upload_vec(c, v_nat, &du))) return rc;
  ft.v_nat = du;
  wt->fast_ready = true;
  return RSG_OK;
}
// Blocked mode: the same constants cut into blocks of h = TS/2 coefficients, each transformed at size TS and scaled by 1/TS;
// the tree levels with 2m <= TS keep the layout above (S = the coefficient-buffer size), the level m = TS is P - x^TS as two blocks.
static int ensure_big_tables(rsg_context *c, size_t n, WitnessTables *wt) {
  FastTables &ft = wt->ft;
  memset(&ft, 0, sizeof(ft));
  const uint32_t TS = wf_max_ts(c), h = TS / 2;
  uint32_t lgT = 0;
  while ((1u << lgT) < TS) lgT++;
  const uint32_t Sb = n <= 2 * (size_t)h ? 2 * h : 4 * h, nx = (uint32_t)((n + h - 1) / h);
  ft.n = (uint32_t)n; ft.S = Sb; ft.wc = 0; ft.big = 1; ft.TS = TS; ft.logTS = lgT; ft.nx = nx;
  while ((1u << ft.logS) < Sb) ft.logS++;
  uint32_t levels = 0;
  for (size_t m = WF_B; m < n && 2 * m <= TS; m <<= 1) levels++;
  ft.levels = std::max<uint32_t>(levels, 1);
  const size_t L_R = c->L_R, npad = (n + WF_B - 1) / WF_B * WF_B, lu = n - 1;
  std::vector<Twiddle> invfact(L_R * n), pts(L_R * npad), Phat(L_R * ft.levels * Sb, Twiddle{0, 0});
  std::vector<uint64_t> Pnat(L_R * ft.levels * (Sb / 2 + 1), 0), Gblk(L_R * nx * TS, 0), Vblk(L_R * nx * TS, 0), Ptop(L_R * 2 * TS, 0);
  for (size_t j = 0; j < L_R; j++) {
    const uint64_t p = c->q[j];
    const rsg_host::NttTables &t = host_ntt(c, j);
    const uint64_t invT = h_inv(TS % p, p);
    ft.invTS[j] = h_twiddle(invT, p);
    ft.invS[j] = ft.invTS[j];
    std::vector<uint64_t> fact(n + 1, 1), ifact(n + 1), g(n);
    for (size_t i = 1; i <= n; i++) fact[i] = h_mulmod(fact[i - 1], i % p, p);
    ifact[n] = h_inv(fact[n], p);
    for (size_t i = n; i > 0; i--) ifact[i - 1] = h_mulmod(ifact[i], i % p, p);
    for (size_t i = 0; i < n; i++) {
      invfact[j * n + i] = h_twiddle(ifact[i], p);
      g[i] = (i & 1) ? (p - ifact[i]) % p : ifact[i];
    }
    for (size_t i = 0; i < npad; i++) pts[j * npad + i] = h_twiddle(i % p, p);
    auto put_blocks = [&](const std::vector<uint64_t> &src, size_t len, uint64_t *dst) {   // transformed, scaled blocks of src[0..len)
      const rsg_host::Poly a(src.begin(), src.begin() + len);
      const std::vector<rsg_host::Poly> blk = rsg_host::blocks_fwd(a, h, (int)lgT, t);
      for (size_t b = 0; b < blk.size(); b++)
        for (size_t i = 0; i < TS; i++) dst[b * TS + i] = h_mulmod(blk[b][i], invT, p);
    };
    put_blocks(g, n, Gblk.data() + j * nx * TS);
    {
      std::vector<uint64_t> u(wt->h_u.begin() + j * std::max<size_t>(lu, 1), wt->h_u.begin() + j * std::max<size_t>(lu, 1) + lu);
      put_blocks(u, lu, Vblk.data() + j * nx * TS);
    }
    std::vector<rsg_host::Poly> tree(npad / WF_B);
    for (size_t r = 0; r < tree.size(); r++) tree[r] = rsg_host::node_product(r * WF_B, (r + 1) * WF_B, t);
    size_t lvl = 0;
    for (size_t m = WF_B; m < n; m <<= 1, lvl++) {
      const size_t two_m = 2 * m;
      if (two_m <= TS) {
        const size_t nb_active = (n - m + two_m - 1) / two_m;
        int lg = 0;
        while (((size_t)1 << lg) < two_m) lg++;
        const uint64_t inv2m = h_inv(two_m % p, p);
        for (size_t b = 0; b < nb_active; b++) {
          rsg_host::Poly a(two_m, 0);
          const rsg_host::Poly &Pb = tree[2 * b];
          for (size_t i = 0; i <= m; i++) a[i] = Pb[i];
          rsg_host::ntt_fwd(a, lg, t);
          Twiddle *dst = Phat.data() + (j * ft.levels + lvl) * Sb + b * two_m;
          for (size_t i = 0; i < two_m; i++) dst[i] = h_twiddle(h_mulmod(a[i], inv2m, p), p);
        }
        const rsg_host::Poly &Pl = tree[2 * (nb_active - 1)];
        std::copy(Pl.begin(), Pl.end(), Pnat.begin() + (j * ft.levels + lvl) * (Sb / 2 + 1));
      } else {   // m = TS: tree[0] = prod_{x < TS} (X - x), monic of degree TS
        put_blocks(tree[0], m, Ptop.data() + j * 2 * TS);
      }
      std::vector<rsg_host::Poly> next;
      if (2 * m < n)   // the root itself is Z: not needed again
        for (size_t r = 0; 2 * r + 1 < tree.size(); r++) next.push_back(rsg_host::polymul(tree[2 * r], tree[2 * r + 1], t));
      tree.swap(next);
    }
  }
  int rc;
  Twiddle *dt;
  uint64_t *du;
  if ((rc = upload_vec(c, invfact, &dt))) return rc;
  ft.invfact = dt;
  if ((rc = upload_vec(c, pts, &dt))) return rc;
  ft.pts = dt;
  if ((rc = upload_vec(c, Phat, &dt))) return rc;
  ft.Phat = dt;
  if ((rc = upload_vec(c, Pnat, &du))) return rc;
  ft.Pnat = du;
  if ((rc = upload_vec(c, Gblk, &du))) return rc;
  ft.Gblk = du;
  if ((rc = upload_vec(c, Vblk, &du))) return rc;
  ft.Vblk = du;
  if ((rc = upload_vec(c, Ptop, &du))) return rc;
  ft.Ptop = du;
  wt->fast_ready = true;
  return RSG_OK;
}
// slots per CTA: a power of two dividing the slot count; both polynomial buffers of a CTA stay below ~100 KiB (2 CTAs/SM)
// S = 16384: two S-word buffers per slot exceed an SM's shared memory; the scratch buffer then lives in global memory (L2)
static int wf_n_global(uint32_t S) { return wf_smem_bytes(S, 1, 0) <= 227 * 1024 ? 0 : (wf_smem_bytes(S, 1, 1) <= 227 * 1024 ? 1 : 2); }
static bool wf_b_global(uint32_t S) { return wf_n_global(S) > 0; }
static uint32_t wf_pick_sl(const rsg_context *c, uint32_t S, size_t nslots, size_t vectors) {
  if (wf_b_global(S)) return 1;
  // measured on B200 at C4 (n = 1031, S = 2048): 2 slots per CTA, two to three CTAs per SM, beats 4 x 512 threads by 12 %
  uint32_t sl = 2;
  if (c->wf_sl > 0) sl = (uint32_t)c->wf_sl;
  while (sl > 1 && (nslots % sl || wf_smem_bytes(S, sl) > (c->wf_sl > 0 ? 227 : 100) * 1024)) sl >>= 1;
  // (one slot per CTA when a multi-GPU slot shard leaves less than two waves of CTAs was tried: 0.69 vs 0.65 ms at 4 GPUs)
  (void)vectors;
  return sl;
}
// correction-free butterflies (witness_fast.cuh) need every ring prime below 2^57; RSG_WF_LAZY=0 forces the corrected ones
static bool wf_lazy(const rsg_context *c) {
  if (const char *m = getenv("RSG_WF_LAZY")) if (!atoi(m)) return false;
  for (uint64_t p : c->q)
    if (p >= (1ull << 57)) return false;
  return true;
}
static unsigned wf_threads(const rsg_context *c, int sl, uint32_t S) {
  if (c->wf_threads >= 32 && c->wf_threads <= 512) return (unsigned)c->wf_threads & ~31u;
  // a full-size pass has sl * S / 16 radix-16 items: one thread per item, at most 256 threads.  Measured on B200
  // (profiles/r1d_wf_tune.log, r1d/r1e_c5_sweep.json; ms for 8 vectors): S = 512: 64 thr 11.5 / 128 thr 17.2; S = 2048 x 2 slots:
  // 128 thr 51.5 / 256 thr 46.7 (N_R = 2^15) and 3.32 / 3.49 (N_R = 2^11); S = 4096: 128 thr 186 / 256 thr 113; S = 8192: 256 thr
  // 302 / 512 thr 542 -- idle threads lengthen the barriers, too few warps expose the latencies
  return (unsigned)std::min<size_t>(256, std::max<size_t>(64, (size_t)sl * S / 16));
}
template <int SL>
static int wf_launch_interp(rsg_context *c, const FastTables &ft, const uint64_t *Y, uint64_t *C, size_t batch, size_t nslots,
                            size_t coef_stride, size_t limb_stride, size_t vec_stride) {
  const int n_global = wf_n_global(ft.S);
  const size_t smem = wf_smem_bytes(ft.S, SL, n_global);
  const dim3 grid((unsigned)(nslots / SL), (unsigned)(batch * c->L_R));
  uint64_t *gB = nullptr, *gA = nullptr;
  if (n_global) {
    const size_t per = (size_t)grid.x * grid.y * SL * wf_slot_stride(ft.S);
    int rc = ensure(c, &c->d_wfB, &c->cap_wfB, per * n_global);
    if (rc) return rc;
    gB = c->d_wfB;
    if (n_global == 2) gA = c->d_wfB + per;
  }
  auto kern = wf_lazy(c) ? k_interp_fast<SL, true> : k_interp_fast<SL, false>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, wf_threads(c, SL, ft.S), smem, c->stream>>>(c->d_params, ft, Y, C, coef_stride, limb_stride, vec_stride, gB, gA);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
// blocked mode: a persistent grid, a few CTAs per SM, each with a private range of global scratch
static int wf_big_grid(rsg_context *c, const FastTables &ft, size_t items, unsigned *grid, uint64_t **scratch) {
  int per_sm = 4;
  if (const char *m = getenv("RSG_WF_BIG_CTAS")) per_sm = std::max(1, std::min(8, atoi(m)));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  *grid = (unsigned)std::min<size_t>(items, (size_t)sms * per_sm);
  int rc = ensure(c, &c->d_wfB, &c->cap_wfB, (size_t)*grid * wf_big_words(ft.S, ft.TS));
  if (rc) return rc;
  *scratch = c->d_wfB;
  return RSG_OK;
}
static int launch_interp_fast(rsg_context *c, WitnessTables *wt, const uint64_t *Y, uint64_t *C, size_t batch, size_t nslots,
                              size_t coef_stride, size_t limb_stride, size_t vec_stride, const char *name = "k_interp_fast") {
  if (!batch) return RSG_OK;
  LaunchScope ls(c, name);
  c->st_wf++;
  if (wt->ft.big) {
    const size_t items = nslots * batch * c->L_R;
    unsigned grid;
    uint64_t *scratch;
    int rc = wf_big_grid(c, wt->ft, items, &grid, &scratch);
    if (rc) return rc;
    auto kern = wf_lazy(c) ? k_interp_big<true> : k_interp_big<false>;
    kern<<<grid, 256, 0, c->stream>>>(c->d_params, wt->ft, Y, C, coef_stride, limb_stride, vec_stride, (uint32_t)nslots, (uint32_t)items, scratch);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  switch (wf_pick_sl(c, wt->ft.S, nslots, batch * c->L_R)) {
    case 8: return wf_launch_interp<8>(c, wt->ft, Y, C, batch, nslots, coef_stride, limb_stride, vec_stride);
    case 4: return wf_launch_interp<4>(c, wt->ft, Y, C, batch, nslots, coef_stride, limb_stride, vec_stride);
    case 2: return wf_launch_interp<2>(c, wt->ft, Y, C, batch, nslots, coef_stride, limb_stride, vec_stride);
    default: return wf_launch_interp<1>(c, wt->ft, Y, C, batch, nslots, coef_stride, limb_stride, vec_stride);
  }
}
template <int SL>
static int wf_launch_quotient(rsg_context *c, const FastTables &ft, const uint64_t *A, const uint64_t *B, uint64_t *H) {
  const int n_global = wf_n_global(ft.S);
  const size_t smem = wf_smem_bytes(ft.S, SL, n_global);
  const dim3 grid((unsigned)(c->N_R / SL), (unsigned)c->L_R);
  uint64_t *gB = nullptr, *gA = nullptr;
  if (n_global) {
    const size_t per = (size_t)grid.x * grid.y * SL * wf_slot_stride(ft.S);
    int rc = ensure(c, &c->d_wfB, &c->cap_wfB, per * n_global);
    if (rc) return rc;
    gB = c->d_wfB;
    if (n_global == 2) gA = c->d_wfB + per;
  }
  auto kern = wf_lazy(c) ? k_quotient_fast<SL, true> : k_quotient_fast<SL, false>;
  CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, wf_threads(c, SL, ft.S), smem, c->stream>>>(c->d_params, ft, A, B, H, gB, gA);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
static int launch_quotient_fast(rsg_context *c, WitnessTables *wt, const uint64_t *A, const uint64_t *B, uint64_t *H) {
  LaunchScope ls(c, "k_quotient_fast");
  c->st_wf++;
  if (wt->ft.big) {
    const size_t items = c->N_R * c->L_R;
    unsigned grid;
    uint64_t *scratch;
    int rc = wf_big_grid(c, wt->ft, items, &grid, &scratch);
    if (rc) return rc;
    auto kern = wf_lazy(c) ? k_quotient_big<true> : k_quotient_big<false>;
    kern<<<grid, 256, 0, c->stream>>>(c->d_params, wt->ft, A, B, H, (uint32_t)items, scratch);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  switch (wf_pick_sl(c, wt->ft.S, c->N_R, c->L_R)) {
    case 8: return wf_launch_quotient<8>(c, wt->ft, A, B, H);
    case 4: return wf_launch_quotient<4>(c, wt->ft, A, B, H);
    case 2: return wf_launch_quotient<2>(c, wt->ft, A, B, H);
    default: return wf_launch_quotient<1>(c, wt->ft, A, B, H);
  }
}

static int launch_modmat(rsg_context *c, const uint64_t *d_M, size_t rows, size_t K, const uint64_t *d_Y, uint64_t *d_C, size_t batch,
                         bool upper, const char *name) {
  if (!rows || !batch) return RSG_OK;
  LaunchScope ls(c, name);
  c->st_wd++;
  bool small = true;   // every ring prime < 2^54: the FP64-pipe kernel (exact; see witness.cuh)
  for (uint64_t p : c->q) small = small && p < (1ull << 54);
  dim3 grid((unsigned)((rows + MM_ROWS - 1) / MM_ROWS), (unsigned)((c->N_R + MM_THREADS - 1) / MM_THREADS), (unsigned)(batch * c->L_R));
  if (small) {
    // 8 rows x 1 slot per thread, 2 columns of Y prefetched, 3 CTAs/SM: best of the shapes measured on B200
    dim3 gridf((unsigned)((rows + 7) / 8), (unsigned)((c->N_R + MM_THREADS - 1) / MM_THREADS), grid.z);
    k_modmat_f64<8, 1, 2, 3><<<gridf, MM_THREADS, 0, c->stream>>>(c->d_modq, d_M, (uint32_t)rows, (uint32_t)K, d_Y, d_C, (uint32_t)c->N_R,
                                                                  (uint32_t)c->L_R, upper ? 1u : 0u);
  }
  else
    k_modmat<<<grid, MM_THREADS, 0, c->stream>>>(c->d_modq, d_M, (uint32_t)rows, (uint32_t)K, d_Y, d_C, (uint32_t)c->N_R, (uint32_t)c->L_R,
                                                 upper ? 1u : 0u, (uint32_t)K);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

// `batch` vectors of n ring elements: evaluations on {0..n-1} -> monomial coefficients, by whichever path applies
static int interpolate_dev(rsg_context *c, size_t n, WitnessTables *wt, const uint64_t *d_Y, uint64_t *d_C, size_t batch) {
  int rc;
  const size_t W = c->ring_words();
  if (wf_use(c, n)) {
    if ((rc = ensure_fast_tables(c, n, wt))) return rc;
    return launch_interp_fast(c, wt, d_Y, d_C, batch, c->N_R, W, c->N_R, n * W);
  }
  if ((rc = ensure_dense_tables(c, n, wt))) return rc;
  return launch_modmat(c, wt->d_Vinv, n, n, d_Y, d_C, batch, false, "k_modmat_interp");
}

extern "C" int rsg_interpolate(rsg_context *c, size_t n, size_t batch, const rsg_ringvec *y, size_t y_first, rsg_ringvec *out,
                               size_t out_first) {
  RSG_TRACE_CALL();
  if (!c || !y || !out) return fail(RSG_ERR_ARG, "null argument");
  if (y_first + batch * n > y->n || out_first + batch * n > out->n) return fail(RSG_ERR_ARG, "range");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  WitnessTables *wt;
  int rc = get_witness_tables(c, n, &wt);
  if (rc) return rc;
  return interpolate_dev(c, n, wt, y->d + y_first * c->ring_words(), out->d + out_first * c->ring_words(), batch);
}

extern "C" int rsg_vanishing(rsg_context *c, size_t n, uint64_t *h_Z) {
  RSG_TRACE_CALL();
  if (!c || !h_Z) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  WitnessTables *wt;
  int rc = get_witness_tables(c, n, &wt);
  if (rc) return rc;
  memcpy(h_Z, wt->h_Z.data(), wt->h_Z.size() * 8);
  return RSG_OK;
}

// r1cs (nullable): the evaluations come from rsg_r1cs_evaluate on this system, so full = mid + io - constant wire and the
// two interpolants of the full assignment follow by linearity (6 interpolations per proof instead of 8).
static int witness_map_dev(rsg_context *c, size_t n, const uint64_t *d_evals, uint64_t *d_coeffs, uint64_t *d_H,
                           const uint64_t *d_zk = nullptr, rsg_r1cs *r1cs = nullptr, bool need_C = true) {
  WitnessTables *wt;
  int rc = get_witness_tables(c, n, &wt);
  if (rc) return rc;
  const bool fast = wf_use(c, n);
  if (fast ? (rc = ensure_fast_tables(c, n, wt)) : (rc = ensure_dense_tables(c, n, wt))) return rc;
  const size_t W = c->ring_words();
  // scratch: aA, aB (n each) + Ptop (n-1): reuse d_plain (words)
  if ((rc = ensure(c, &c->d_plain, &c->cap_plain, (3 * n) * W))) return rc;
  uint64_t *aA = c->d_plain, *aB = aA + n * W, *Ptop = aB + n * W;
  // evals order: A_mid,B_mid,C_mid,A_io,B_io,C_io,A_full,B_full,C_full ; coeffs order: A_io,B_io,C_io,A_mid,B_mid,C_mid
  // need_C = false (ringGroth16, groth16.tcc:89-112): C_io / C_mid are never read by the prover and C does not reach the
  // quotient (deg C < n), so only A and B are interpolated
  const size_t nb = need_C ? 3 : 2;
  if ((rc = interpolate_dev(c, n, wt, d_evals + 3 * n * W, d_coeffs, nb))) return rc;
  if ((rc = interpolate_dev(c, n, wt, d_evals, d_coeffs + 3 * n * W, nb))) return rc;
  if (r1cs && r1cs->n == n) {
    if (!r1cs->d_cc) {
      uint64_t *d_const = nullptr;
      void *v = nullptr;
      CUDA_TRY(cudaMalloc(&v, r1cs->h_const.size() * 8));
      d_const = (uint64_t *)v;
      struct DevFree {   // staging buffer: released on every way out
        void *p;
        ~DevFree() { cudaFree(p); }
      } staging{d_const};
      CUDA_TRY(cudaMalloc(&v, r1cs->h_const.size() * 8));
      r1cs->d_cc = (uint64_t *)v;
      CUDA_TRY(cudaMemcpyAsync(d_const, r1cs->h_const.data(), r1cs->h_const.size() * 8, cudaMemcpyHostToDevice, c->stream));
      if (fast) {   // the constants are [vector][L_R][n] words: one "slot" per limb
        if ((rc = launch_interp_fast(c, wt, d_const, r1cs->d_cc, 2, 1, 1, n, c->L_R * n, "k_interp_fast_const"))) return rc;
      } else {
        for (int m = 0; m < 2; m++) {
          LaunchScope ls(c, "k_matvec");
          k_matvec<<<dim3((unsigned)((n + 127) / 128), (unsigned)c->L_R), 128, 0, c->stream>>>(
              c->d_modq, wt->d_Vinv, (uint32_t)n, d_const + m * c->L_R * n, r1cs->d_cc + m * c->L_R * n);
        }
      }
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    LaunchScope ls(c, "k_full_from_parts");
    k_full_from_parts<<<dim3((unsigned)n, (unsigned)((W + 255) / 256), 2), 256, 0, c->stream>>>(c->d_modq, d_coeffs, r1cs->d_cc, aA, (uint32_t)n,
                                                                                               (uint32_t)c->N_R, (uint32_t)c->L_R);
    CUDA_TRY(cudaGetLastError());
  } else if ((rc = interpolate_dev(c, n, wt, d_evals + 6 * n * W, aA, 2))) {
    return rc;
  }
  CUDA_TRY(cudaMemsetAsync(d_H, 0, (n + 1) * W * 8, c->stream));
  if (n >= 2) {
    if (fast) {
      if ((rc = launch_quotient_fast(c, wt, aA, aB, d_H))) return rc;
    } else {
      {
        dim3 grid((unsigned)((n - 1 + MM_ROWS - 1) / MM_ROWS), (unsigned)((c->N_R + MM_THREADS - 1) / MM_THREADS), (unsigned)c->L_R);
        LaunchScope ls(c, "k_conv_top");
        k_conv_top<<<grid, MM_THREADS, 0, c->stream>>>(c->d_modq, aA, aB, (uint32_t)n, (uint32_t)n, (uint32_t)n, Ptop, (uint32_t)c->N_R,
                                                       (uint32_t)c->L_R);
        CUDA_TRY(cudaGetLastError());
      }
      if ((rc = launch_modmat(c, wt->d_T, n - 1, n - 1, Ptop, d_H, 1, true, "k_modmat_divZ"))) return rc;
    }
  }
  if (d_zk) {
    LaunchScope ls(c, "k_h_patch");
    k_h_patch<<<dim3((unsigned)(n + 1), (unsigned)((W + 255) / 256)), 256, 0, c->stream>>>(c->d_modq, d_H, aA, aB, d_zk, wt->d_Z, (uint32_t)n,
                                                                                         (uint32_t)c->N_R, (uint32_t)c->L_R);
    CUDA_TRY(cudaGetLastError());
  }
  return RSG_OK;
}

extern "C" int rsg_witness_map_zk(rsg_context *c, size_t n, const rsg_ringvec *evals, const uint64_t *h_d, rsg_ringvec *coeffs,
                                  rsg_ringvec *H) {
  RSG_TRACE_CALL();
  if (!c || !evals || !coeffs || !H) return fail(RSG_ERR_ARG, "null argument");
  if (evals->n < 9 * n || coeffs->n < 6 * n || H->n < n + 1) return fail(RSG_ERR_ARG, "witness-map vector sizes");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t *d_zk = nullptr;
  if (h_d) {
    int rc = ensure(c, &c->d_zk, &c->cap_zk, 3 * c->ring_words());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->d_zk, h_d, 3 * c->ring_words() * 8, cudaMemcpyHostToDevice, c->stream));
    d_zk = c->d_zk;
  }
  return witness_map_dev(c, n, evals->d, coeffs->d, H->d, d_zk);
}
extern "C" int rsg_witness_map_r1cs(rsg_context *c, rsg_r1cs *r1cs, const rsg_ringvec *evals, const uint64_t *h_d, rsg_ringvec *coeffs,
                                    rsg_ringvec *H) {
  RSG_TRACE_CALL();
  if (!c || !r1cs || !evals || !coeffs || !H) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r1cs->n;
  if (evals->n < 9 * n || coeffs->n < 6 * n || H->n < n + 1) return fail(RSG_ERR_ARG, "witness-map vector sizes");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const uint64_t *d_zk = nullptr;
  if (h_d) {
    int rc = ensure(c, &c->d_zk, &c->cap_zk, 3 * c->ring_words());
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->d_zk, h_d, 3 * c->ring_words() * 8, cudaMemcpyHostToDevice, c->stream));
    d_zk = c->d_zk;
  }
  return witness_map_dev(c, n, evals->d, coeffs->d, H->d, d_zk, r1cs);
}
extern "C" int rsg_witness_map_groth16(rsg_context *c, rsg_r1cs *r1cs, const rsg_ringvec *evals, rsg_ringvec *coeffs, rsg_ringvec *H) {
  RSG_TRACE_CALL();
  if (!c || !r1cs || !evals || !coeffs || !H) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r1cs->n;
  if (evals->n < 9 * n || coeffs->n < 6 * n || H->n < n + 1) return fail(RSG_ERR_ARG, "witness-map vector sizes");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return witness_map_dev(c, n, evals->d, coeffs->d, H->d, nullptr, r1cs, /*need_C=*/false);
}
extern "C" int rsg_witness_map(rsg_context *c, size_t n, const rsg_ringvec *evals, rsg_ringvec *coeffs, rsg_ringvec *H) {
  return rsg_witness_map_zk(c, n, evals, nullptr, coeffs, H);
}

// ------------------------------------------------------------------------------------------------------------
// R1CS evaluation (the step before the hot path)

extern "C" int rsg_r1cs_create(rsg_context *c, size_t n, size_t n_io, size_t n_aux, const uint32_t *h_row_ptr, const uint32_t *h_col,
                               const uint64_t *h_coeff, rsg_r1cs **out) {
  RSG_TRACE_CALL();
  if (!c || !h_row_ptr || !out || !n) return fail(RSG_ERR_ARG, "null argument");
  const size_t nnz = h_row_ptr[3 * n];
  for (size_t t = 0; t < nnz; t++)
    if (h_col[t] > n_io + n_aux) return fail(RSG_ERR_ARG, "variable index out of range");
  CUDA_TRY(cudaSetDevice(c->device));
  rsg_r1cs *r = new rsg_r1cs{c, n, n_io, n_aux};
  r->h_row_ptr.assign(h_row_ptr, h_row_ptr + 3 * n + 1);
  r->h_col.assign(h_col, h_col + nnz);
  r->h_coeff.assign(h_coeff, h_coeff + nnz);
  r->h_const.assign(2 * c->L_R * n, 0);
  for (size_t m = 0; m < 2; m++)
    for (size_t i = 0; i < n; i++)
      for (size_t t = h_row_ptr[m * n + i]; t < h_row_ptr[m * n + i + 1]; t++)
        if (h_col[t] == 0)
          for (size_t j = 0; j < c->L_R; j++) {
            uint64_t &acc = r->h_const[(m * c->L_R + j) * n + i];
            acc = (uint64_t)(((u128)acc + h_coeff[t] % c->q[j]) % c->q[j]);
          }
  void *v;
  CUDA_TRY(cudaMalloc(&v, (3 * n + 1) * 4)); r->d_row_ptr = (uint32_t *)v;
  CUDA_TRY(cudaMalloc(&v, std::max<size_t>(nnz, 1) * 4)); r->d_col = (uint32_t *)v;
  CUDA_TRY(cudaMalloc(&v, std::max<size_t>(nnz, 1) * 8)); r->d_coeff = (uint64_t *)v;
  CUDA_TRY(cudaMemcpy(r->d_row_ptr, h_row_ptr, (3 * n + 1) * 4, cudaMemcpyHostToDevice));
  if (nnz) {
    CUDA_TRY(cudaMemcpy(r->d_col, h_col, nnz * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->d_coeff, h_coeff, nnz * 8, cudaMemcpyHostToDevice));
  }
  *out = r;
  return RSG_OK;
}
extern "C" void rsg_r1cs_destroy(rsg_r1cs *r) {
  RSG_TRACE_CALL();
  if (!r) return;
  cudaStreamSynchronize(r->ctx->stream);
  cudaFree(r->d_row_ptr); cudaFree(r->d_col); cudaFree(r->d_coeff); cudaFree(r->d_cc);
  cudaFree(r->d_col_ptr); cudaFree(r->d_rows); cudaFree(r->d_ccoeff);
  delete r;
}
// ------------------------------------------------------------------------------------------------------------
// EncodingElem::decode on the device (decode.cuh)
static int get_decode_consts(rsg_context *c) {
  if (c->d_dec) return RSG_OK;
  DecodeConsts K;
  memset(&K, 0, sizeof(K));
  const size_t L_E = c->L_E, L_R = c->L_R;
  for (size_t l = 0; l < L_E; l++) {
    uint64_t prod = 1;   // Q / Q_l mod Q_l
    for (size_t k = 0; k < L_E; k++)
      if (k != l) prod = h_mulmod(prod, c->Q[k] % c->Q[l], c->Q[l]);
    K.inv_punct[l] = h_inv(prod, c->Q[l]);
    for (size_t j = 0; j < L_R; j++) {
      uint64_t pm = 1;
      for (size_t k = 0; k < L_E; k++)
        if (k != l) pm = h_mulmod(pm, c->Q[k] % c->q[j], c->q[j]);
      K.punct_mod_t[j][l] = pm;
    }
    for (size_t k = 0; k < l; k++) K.garner_inv[l][k] = h_inv(c->Q[k] % c->Q[l], c->Q[l]);
  }
  for (size_t j = 0; j < L_R; j++) {
    uint64_t qm = 1;
    for (size_t k = 0; k < L_E; k++) qm = h_mulmod(qm, c->Q[k] % c->q[j], c->q[j]);
    K.Q_mod_t[j] = qm;
  }
  K.Qw[0] = 1;   // Q = prod Q_l, little-endian words
  for (size_t l = 0; l < L_E; l++) {
    uint64_t carry = 0;
    for (size_t w = 0; w < DEC_MAXW; w++) {
      const u128 t = (u128)K.Qw[w] * c->Q[l] + carry;
      K.Qw[w] = (uint64_t)t;
      carry = (uint64_t)(t >> 64);
    }
  }
  {   // (Q + 1) >> 1
    uint64_t tmp[DEC_MAXW + 1] = {0}, carry = 1;
    for (size_t w = 0; w < DEC_MAXW; w++) {
      tmp[w] = K.Qw[w] + carry;
      carry = tmp[w] < carry ? 1 : 0;
    }
    tmp[DEC_MAXW] = carry;
    for (size_t w = 0; w < DEC_MAXW; w++) K.halfw[w] = (tmp[w] >> 1) | (tmp[w + 1] << 63);
    c->Q_bits = 0;
    for (int w = DEC_MAXW - 1; w >= 0; w--)
      if (K.Qw[w]) { c->Q_bits = w * 64 + (64 - __builtin_clzll(K.Qw[w])); break; }
  }
  int rc = dev_alloc(c, &c->d_dec, 1);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpy(c->d_dec, &K, sizeof(K), cudaMemcpyHostToDevice));
  return RSG_OK;
}

extern "C" int rsg_decode(rsg_context *c, const uint64_t *h_sk, const uint64_t *d_enc, const uint64_t *h_enc, size_t count,
                          uint64_t *h_ring, int32_t *h_budget) {
  RSG_TRACE_CALL();
  if (!c || !h_sk || (!d_enc == !h_enc) || !h_ring) return fail(RSG_ERR_ARG, "null argument (exactly one of d_enc / h_enc)");
  if (!count) return RSG_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = get_decode_consts(c);
  if (rc) return rc;
  const size_t E = c->enc_words(), N = c->N_E, L_E = c->L_E, L_R = c->L_R, W = c->ring_words();
  // scratch: [enc (if from host) | sk | phase | plain | ring | bits]
  const size_t need = (h_enc ? count * E : 0) + L_R * L_E * N + L_E * count * L_R * N + L_R * count * N + count * W + count * L_R;
  if ((rc = ensure(c, &c->d_decode, &c->cap_decode, need))) return rc;
  uint64_t *p = c->d_decode;
  const uint64_t *enc = d_enc;
  if (h_enc) {
    CUDA_TRY(cudaMemcpyAsync(p, h_enc, count * E * 8, cudaMemcpyHostToDevice, c->stream));
    enc = p;
    p += count * E;
  }
  uint64_t *sk = p; p += L_R * L_E * N;
  uint64_t *phase = p; p += L_E * count * L_R * N;
  uint64_t *plain = p; p += L_R * count * N;
  uint64_t *ring = p; p += count * W;
  int *bits = (int *)p;
  CUDA_TRY(cudaMemcpyAsync(sk, h_sk, L_R * L_E * N * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemsetAsync(bits, 0, count * L_R * sizeof(int), c->stream));
  {
    LaunchScope ls(c, "k_dec_phase");
    k_dec_phase<<<dim3((unsigned)((N + 255) / 256), (unsigned)(L_E * L_R), (unsigned)count), 256, 0, c->stream>>>(c->d_params, enc, sk,
                                                                                                              (uint32_t)count, phase);
    CUDA_TRY(cudaGetLastError());
  }
  for (size_t l = 0; l < L_E; l++)
    if ((rc = ntt_dev(c, phase + l * count * L_R * N, count * L_R, 0, l, 1))) return rc;
  {
    LaunchScope ls(c, "k_dec_modt");
    k_dec_modt<<<dim3((unsigned)((N + 127) / 128), (unsigned)L_R, (unsigned)count), 128, 0, c->stream>>>(c->d_params, c->d_dec, phase,
                                                                                                      (uint32_t)count, plain, bits);
    CUDA_TRY(cudaGetLastError());
  }
  for (size_t j = 0; j < L_R; j++)
    if ((rc = ntt_dev(c, plain + j * count * N, count, 1, j, 0))) return rc;
  {
    LaunchScope ls(c, "k_dec_gather");
    k_dec_gather<<<dim3((unsigned)((c->N_R + 255) / 256), (unsigned)L_R, (unsigned)count), 256, 0, c->stream>>>(c->d_params, plain,
                                                                                                             (uint32_t)count, ring);
    CUDA_TRY(cudaGetLastError());
  }
  std::vector<int> hb(count * L_R);
  CUDA_TRY(cudaMemcpyAsync(h_ring, ring, count * W * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(hb.data(), bits, hb.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  bool exhausted = false;
  for (size_t k = 0; k < hb.size(); k++) {
    const int budget = std::max(0, c->Q_bits - hb[k] - 1);   // decryptor.cpp:457-459
    if (h_budget) h_budget[k] = budget;
    exhausted = exhausted || budget <= 0;
  }
  // seal_ring.tcc:445-453: decoding_error when a ciphertext has no noise budget left
  if (exhausted) return fail(RSG_ERR_NOISE, "a ciphertext has remaining noise budget <= 0");
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// EncodingElem::encode on the device (encode.cuh)
extern "C" int rsg_encode(rsg_context *c, const uint64_t *h_sk, const rsg_ringvec *elems, size_t first, size_t count,
                          const uint64_t *h_seeds, rsg_crs *out, size_t out_first) {
  RSG_TRACE_CALL();
  if (!c || !h_sk || !elems || !h_seeds || !out) return fail(RSG_ERR_ARG, "null argument");
  if (first + count > elems->n || out_first + count > out->n) return fail(RSG_ERR_ARG, "range");
  if (!count) return RSG_OK;
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc;
  const size_t L_R = c->L_R, L_E = c->L_E, N_E = c->N_E, poly = L_R * N_E, per_general = poly * L_E, W = c->ring_words();
  const size_t chunk = std::min<size_t>(count, std::max<size_t>(1, std::min<size_t>(256, c->pntt_budget_words / per_general)));
  const uint32_t nb_boot = (uint32_t)((64 + 6 * N_E + 4095) / 4096), nb_ct = (uint32_t)((L_E * N_E + 511) / 512);
  const bool aside = (L_E * N_E) % 512 != 0;   // the bulk of sample_poly_uniform is not a whole number of PRNG buffers
  const size_t boot_stride = (size_t)nb_boot * 512, bulk_stride = (size_t)nb_ct * 512, S = chunk * L_R;
  // scratch: [sk | seeds | boot | roots | noise | bulk (if aside) | err]
  const size_t n_sk = L_R * L_E * N_E, n_seeds = S * 8, n_boot = S * boot_stride, n_roots = S * std::max(nb_boot, nb_ct) * 8,
               n_noise = S * L_E * N_E, n_bulk = aside ? S * bulk_stride : 0;
  if ((rc = ensure(c, &c->d_encode, &c->cap_encode, n_sk + n_seeds + n_boot + n_roots + n_noise + n_bulk + 8))) return rc;
  if ((rc = ensure(c, &c->d_plain, &c->cap_plain, chunk * poly))) return rc;
  if ((rc = ensure(c, &c->d_pntt, &c->cap_pntt, chunk * per_general))) return rc;
  uint64_t *d_sk = c->d_encode, *d_seeds = d_sk + n_sk, *d_boot = d_seeds + n_seeds, *d_roots = d_boot + n_boot, *d_noise = d_roots + n_roots;
  uint64_t *d_bulk = d_noise + n_noise;
  uint32_t *d_err = (uint32_t *)(d_bulk + n_bulk);
  cudaStream_t st = c->stream;
  CUDA_TRY(cudaMemcpyAsync(d_sk, h_sk, n_sk * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(d_err, 0, 8, st));
  for (size_t e0 = 0; e0 < count; e0 += chunk) {
    const size_t ne = std::min(chunk, count - e0), ns = ne * L_R;
    const uint32_t n_streams = (uint32_t)ns;
    // BatchEncoder::encode + the centred lift and forward NTTs of encryptor.cpp:260-308 (the prover's own kernels)
    if ((rc = launch_encode(c, elems->d + (first + e0) * W, nullptr, ne, c->d_plain))) return rc;
    if ((rc = launch_lift_ntt(c, c->d_plain, ne, c->d_pntt))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_seeds, h_seeds + e0 * L_R * 8, ns * 8 * 8, cudaMemcpyHostToDevice, st));
    {   // bootstrap streams: public seed + noise bytes
      LaunchScope ls(c, "k_b2x");
      const size_t nr = ns * nb_boot;
      k_b2x_roots<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(d_seeds, 8, n_streams, nb_boot, d_roots);
      k_b2x_blocks<<<(unsigned)((nr * 64 + 127) / 128), 128, 0, st>>>(d_roots, n_streams, nb_boot, d_boot, boot_stride);
    }
    uint64_t *arena = out->d;
    {   // ciphertext streams, keyed by the public seeds: the bulk of sample_poly_uniform lands in c1
      LaunchScope ls(c, "k_b2x");
      const size_t nr = ns * nb_ct;
      k_b2x_roots<<<(unsigned)((nr + 127) / 128), 128, 0, st>>>(d_boot, boot_stride, n_streams, nb_ct, d_roots);
      if (aside) k_b2x_blocks<<<(unsigned)((nr * 64 + 127) / 128), 128, 0, st>>>(d_roots, n_streams, nb_ct, d_bulk, bulk_stride);
      else k_b2x_blocks<<<(unsigned)((nr * 64 + 127) / 128), 128, 0, st>>>(d_roots, n_streams, nb_ct,
                                                                          arena + (out_first + e0) * c->enc_words() + L_E * N_E, 2 * L_E * N_E);
    }
    {
      LaunchScope ls(c, "k_enc_uniform_fix");
      k_enc_uniform_fix<<<n_streams, 256, 0, st>>>(c->d_params, arena, out_first + e0, d_boot, boot_stride, d_err, aside ? d_bulk : nullptr,
                                                   bulk_stride);
    }
    {
      LaunchScope ls(c, "k_enc_noise");
      k_enc_noise<<<dim3((unsigned)((N_E + 255) / 256), n_streams), 256, 0, st>>>(c->d_params, d_boot, boot_stride, n_streams, d_noise);
    }
    CUDA_TRY(cudaGetLastError());
    for (size_t l = 0; l < L_E; l++)
      if ((rc = ntt_dev(c, d_noise + l * ns * N_E, ns, 0, l, 0))) return rc;
    {
      LaunchScope ls(c, "k_enc_finish");
      k_enc_finish<<<dim3((unsigned)((N_E + 255) / 256), (unsigned)L_E, n_streams), 256, 0, st>>>(c->d_params, arena, out_first + e0, d_sk, d_noise,
                                                                                                 c->d_pntt, n_streams);
    }
    CUDA_TRY(cudaGetLastError());
  }
  uint32_t err = 0;
  CUDA_TRY(cudaMemcpyAsync(&err, d_err, 4, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (err) return fail(RSG_ERR_UNSUPPORTED, "uniform sampling: more redraws than the device list holds");
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// instance map with evaluation (the O(m^2) step of setup and of every verification; instance.cuh)
extern "C" int rsg_instance_map(rsg_context *c, rsg_r1cs *r, const rsg_ringvec *t, size_t t_first, rsg_ringvec *ABCt, rsg_ringvec *Ht,
                                rsg_ringvec *Zt) {
  RSG_TRACE_CALL();
  if (!c || !r || !t || !ABCt || !Ht || !Zt) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r->n, nv1 = r->n_io + r->n_aux + 1, W = c->ring_words();
  if (t_first >= t->n || ABCt->n < 3 * nv1 || Ht->n < n + 1 || Zt->n < 1) return fail(RSG_ERR_ARG, "instance-map vector sizes");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  WitnessTables *wt;
  int rc = get_witness_tables(c, n, &wt);
  if (rc) return rc;
  if (!wt->d_lagw) {   // 1 / prod_{i != j} (j - i) = (-1)^(n-1-j) / (j! (n-1-j)!)
    std::vector<Twiddle> w(c->L_R * n);
    for (size_t j = 0; j < c->L_R; j++) {
      const uint64_t p = c->q[j];
      std::vector<uint64_t> fact(n + 1, 1);
      for (size_t i = 1; i <= n; i++) fact[i] = h_mulmod(fact[i - 1], i % p, p);
      for (size_t x = 0; x < n; x++) {
        uint64_t d = h_mulmod(fact[x], fact[n - 1 - x], p);
        if ((n - 1 - x) & 1) d = (p - d) % p;
        w[j * n + x] = h_twiddle(h_inv(d, p), p);
      }
    }
    if ((rc = upload_vec(c, w, &wt->d_lagw))) return rc;
  }
  if (!r->d_col_ptr) {   // transpose the CSR rows (matrix, constraint) -> CSC columns (matrix, variable)
    const size_t nnz = r->h_col.size();
    std::vector<uint32_t> col_ptr(3 * nv1 + 1, 0), rows(std::max<size_t>(nnz, 1));
    std::vector<uint64_t> cc(std::max<size_t>(nnz, 1));
    for (size_t m = 0; m < 3; m++)
      for (size_t i = 0; i < n; i++)
        for (size_t e = r->h_row_ptr[m * n + i]; e < r->h_row_ptr[m * n + i + 1]; e++) col_ptr[m * nv1 + r->h_col[e] + 1]++;
    for (size_t k = 0; k < 3 * nv1; k++) col_ptr[k + 1] += col_ptr[k];
    std::vector<uint32_t> fill(col_ptr.begin(), col_ptr.end() - 1);
    for (size_t m = 0; m < 3; m++)
      for (size_t i = 0; i < n; i++)
        for (size_t e = r->h_row_ptr[m * n + i]; e < r->h_row_ptr[m * n + i + 1]; e++) {
          const uint32_t at = fill[m * nv1 + r->h_col[e]]++;
          rows[at] = (uint32_t)i;
          cc[at] = r->h_coeff[e];
        }
    void *v;
    CUDA_TRY(cudaMalloc(&v, col_ptr.size() * 4)); r->d_col_ptr = (uint32_t *)v;
    CUDA_TRY(cudaMalloc(&v, rows.size() * 4)); r->d_rows = (uint32_t *)v;
    CUDA_TRY(cudaMalloc(&v, cc.size() * 8)); r->d_ccoeff = (uint64_t *)v;
    CUDA_TRY(cudaMemcpy(r->d_col_ptr, col_ptr.data(), col_ptr.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(r->d_ccoeff, cc.data(), cc.size() * 8, cudaMemcpyHostToDevice));
  }
  if ((rc = ensure(c, &c->d_plain, &c->cap_plain, n * W))) return rc;   // u[0..n): the Lagrange basis at t
  uint64_t *d_u = c->d_plain;
  {
    LaunchScope ls(c, "k_lagrange_at");
    k_lagrange_at<<<dim3((unsigned)((c->N_R + 127) / 128), (unsigned)c->L_R), 128, 0, c->stream>>>(
        c->d_modq, t->d + t_first * W, wt->d_lagw, (uint32_t)n, d_u, Ht->d, Zt->d, (uint32_t)c->N_R, (uint32_t)c->L_R);
    CUDA_TRY(cudaGetLastError());
  }
  {
    LaunchScope ls(c, "k_instance_accum");
    const unsigned sblocks = (unsigned)((c->N_R + 127) / 128);
    k_instance_accum<<<dim3((unsigned)nv1, 3, (unsigned)(c->L_R * sblocks)), 128, 0, c->stream>>>(
        c->d_modq, r->d_col_ptr, r->d_rows, r->d_ccoeff, (uint32_t)nv1, d_u, ABCt->d, (uint32_t)c->N_R, (uint32_t)c->L_R);
    CUDA_TRY(cudaGetLastError());
  }
  return RSG_OK;
}

static int r1cs_eval_dev(rsg_context *c, const rsg_r1cs *r, const uint64_t *d_assign, uint64_t *d_evals) {
  const unsigned sblocks = (unsigned)((c->N_R + MM_THREADS - 1) / MM_THREADS);
  dim3 grid((unsigned)r->n, 3, (unsigned)(c->L_R * sblocks));
  LaunchScope ls(c, "k_r1cs_eval");
  k_r1cs_eval<<<grid, MM_THREADS, 0, c->stream>>>(c->d_modq, r->d_row_ptr, r->d_col, r->d_coeff, (uint32_t)r->n, (uint32_t)r->n_io, d_assign,
                                                  d_evals, (uint32_t)c->N_R, (uint32_t)c->L_R);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_r1cs_evaluate(rsg_context *c, const rsg_r1cs *r, const rsg_ringvec *assignment, rsg_ringvec *evals) {
  RSG_TRACE_CALL();
  if (!c || !r || !assignment || !evals) return fail(RSG_ERR_ARG, "null argument");
  if (assignment->n < r->n_io + r->n_aux || evals->n < 9 * r->n) return fail(RSG_ERR_ARG, "vector sizes");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  return r1cs_eval_dev(c, r, assignment->d, evals->d);
}


// ------------------------------------------------------------------------------------------------------------
// The lincomb phase as a static launch sequence (prover_fast.cuh).  Index arrays depend only on the layout, the circuit
// shape and the caller's scalar tags, so they are built once and cached.
struct FastPlan {
  std::vector<uint8_t> key;    // what the plan was built from (layout, shape, tags): cache check
  FastTable T;                 // vec[].base filled per call
  uint32_t n_terms = 0, Z = 0, n_out = 0;
  std::vector<uint32_t> zs0, zs1;   // NTT slots [zs0[z], zs1[z]) feed the terms of split z
  uint32_t n_paired = 0;            // splits placed next to a split over the same CRS encodings (0: d_idx has no zorder)
  uint32_t shared_terms = 0;        // terms whose CRS encoding a neighbouring split reads too
  uint32_t *d_idx = nullptr;   // [(unused) n_terms | pidx n_terms | zoff Z+1 | zr n_out+1 | zorder Z]
  const uint64_t **d_tptr = nullptr;   // device pointer of every term's encoding
  uint8_t *d_kind = nullptr;   // device copies of the per-vector kind arrays
};
static void fast_free_plan(FastPlan *fp) {
  if (!fp) return;
  cudaFree(fp->d_idx);
  cudaFree((void *)fp->d_tptr);
  cudaFree(fp->d_kind);
  delete fp;
}
static void fast_release(rsg_context *c) {
  fast_free_plan(c->fplan);
  c->fplan = nullptr;
  cudaFree(c->d_fp_flags); cudaFree(c->d_fp_parts); cudaFree(c->d_fp_nttsrc); cudaFree(c->d_fp_totals); cudaFree(c->d_fp_status);
  if (c->h_fp_status) cudaFreeHost(c->h_fp_status);
  for (auto &e : c->ev_phase) if (e) cudaEventDestroy(e);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->stream2) cudaStreamDestroy(c->stream2);
}

// What a plan is built from: the coefficient vectors' ranges and scalar tags, and per output the list of (inner product,
// optional bare CRS element appended as a scalar-1 term).
struct FastSpec {
  uint32_t n_vec = 0, nS = 0;
  uint32_t lo[FP_MAXV] = {0}, count[FP_MAXV] = {0};
  const uint8_t *kind[FP_MAXV] = {nullptr};   // host, ABSOLUTE index, nullable; kind_len = elements of the whole vector
  size_t kind_len[FP_MAXV] = {0};
  uint32_t n_ip = 0, ip_vec[FP_MAXIP] = {0};
  const uint64_t *ip_base[FP_MAXIP] = {nullptr};   // device address of the CRS encoding multiplying element 0 of the vector's RANGE
  // term groups in output order: group g sums inner products grp_ip[g][0..1] (second may be 0xFFFFFFFF) + an optional bare CRS
  // encoding; merged = the two inner products are the io / mid halves of a merged pair (one slot per term)
  uint32_t n_grp = 0, grp_ip[FP_MAXIP][2], grp_out[FP_MAXIP];
  const uint64_t *grp_extra[FP_MAXIP] = {nullptr};
  bool grp_merged[FP_MAXIP] = {false};
  uint32_t n_out = 0;
  const uint64_t *alpha = nullptr, *beta = nullptr;
};

static int fast_build_plan(rsg_context *c, const FastSpec &sp, const std::vector<uint8_t> &key, FastPlan **out) {
  FastPlan *fp = c->fplan;
  if (fp && fp->key == key) { *out = fp; return RSG_OK; }
  if (fp) {
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    fast_free_plan(fp);
    c->fplan = nullptr;
  }
  fp = new FastPlan();
  fp->key = key;
  FastTable &T = fp->T;
  memset(&T, 0, sizeof(T));
  T.n_vec = sp.n_vec; T.nS = sp.nS;
  T.n_parts = 4 * sp.nS;
  uint32_t eid = 0;
  size_t kind_bytes = 0;
  for (uint32_t k = 0; k < sp.n_vec; k++) {
    T.vec[k].lo = sp.lo[k]; T.vec[k].count = sp.count[k]; T.vec[k].eid0 = eid;
    eid += sp.count[k];
    if (sp.kind[k]) kind_bytes += sp.kind_len[k];
  }
  T.n_elems = eid;
  T.n_slots = eid - T.n_parts + 2 * sp.nS;
  T.n_ip = sp.n_ip;
  for (uint32_t p = 0; p < sp.n_ip; p++) { T.ip_vec[p] = sp.ip_vec[p]; T.ip_base[p] = sp.ip_base[p]; }
  T.alpha = sp.alpha; T.beta = sp.beta;
  const size_t E = c->enc_words();
  auto slot_of = [&](uint32_t k, uint32_t i, bool merged, uint32_t pair) -> uint32_t {   // i = index inside the range
    if (merged) return pair * sp.nS + i;                        // merged pair 0 (A) / 1 (B)
    return 2 * sp.nS + (T.vec[k].eid0 + i - T.n_parts);
  };
  std::vector<const uint64_t *> term;
  std::vector<uint32_t> pidx, grp_t{0};
  for (uint32_t g = 0; g < sp.n_grp; g++) {
    const uint32_t p0 = sp.grp_ip[g][0], p1 = sp.grp_ip[g][1];
    if (sp.grp_merged[g]) {   // one term per element of the pair
      const uint32_t k = sp.ip_vec[p0];
      for (uint32_t i = 0; i < sp.count[k]; i++) { term.push_back(sp.ip_base[p0] + (size_t)i * E); pidx.push_back(slot_of(k, i, true, k / 2)); }
    } else {
      for (uint32_t p : {p0, p1}) {
        if (p == 0xFFFFFFFFu) continue;
        const uint32_t k = sp.ip_vec[p];
        for (uint32_t i = 0; i < sp.count[k]; i++) {
          const uint8_t kd = sp.kind[k] ? sp.kind[k][sp.lo[k] + i] : (uint8_t)RSG_AUX_POLY;
          term.push_back(sp.ip_base[p] + (size_t)i * E);
          pidx.push_back(kd == RSG_TERM_ONE ? 0xFFFFFFFFu : slot_of(k, i, false, 0));   // a SKIP tag is honoured through slot_skip
        }
      }
    }
    if (sp.grp_extra[g]) { term.push_back(sp.grp_extra[g]); pidx.push_back(0xFFFFFFFFu); }
    grp_t.push_back((uint32_t)term.size());
  }
  fp->n_terms = (uint32_t)term.size();
  fp->n_out = sp.n_out;
  // near-equal term chunks that never straddle a group: ~ 148 x 8 x 4 CTAs of k_crs_lincomb in one launch
  const uint32_t base_blocks = (uint32_t)std::max<size_t>(1, (c->N_E / 512) * c->L_R * c->L_E);
  uint32_t want = c->fast_splits > 0 ? (uint32_t)c->fast_splits : std::max(4u, (148u * 8 * 4 + base_blocks - 1) / base_blocks);
  uint32_t chunk = std::max(8u, (fp->n_terms + want - 1) / std::max(1u, want));
  if (c->overlap_mode && c->fast_splits <= 0) chunk = std::min(chunk, 128u);   // phases are built from whole chunks
  std::vector<uint32_t> zoff{0}, zr(sp.n_out + 1, 0), grp_z(sp.n_grp + 1, 0);
  {
    uint32_t g = 0;
    for (uint32_t o = 0; o < sp.n_out; o++) {
      zr[o] = (uint32_t)zoff.size() - 1;
      for (; g < sp.n_grp && sp.grp_out[g] == o; g++) {
        const uint32_t t0 = grp_t[g], t1 = grp_t[g + 1], len = t1 - t0;
        grp_z[g] = (uint32_t)zoff.size() - 1;
        if (!len) continue;
        const uint32_t pieces = (len + chunk - 1) / chunk, per = (len + pieces - 1) / pieces;
        for (uint32_t a = t0; a < t1; a += per) zoff.push_back(std::min(t1, a + per));
      }
    }
    zr[sp.n_out] = (uint32_t)zoff.size() - 1;
    for (; g <= sp.n_grp; g++) grp_z[g] = (uint32_t)zoff.size() - 1;
  }
  fp->Z = (uint32_t)zoff.size() - 1;
  // Launch order of the splits: two groups of equal length and equal cuts whose terms point at the same CRS encodings (the A and B
  // inner products over s_pows; a trailing bare element may differ) are interleaved split by split -- their CTAs then run side
  // by side and the second read of an encoding hits the L2 (k_crs_lincomb_wide, zorder).  RSG_LIN_PAIR=0: plan order.
  std::vector<uint32_t> zorder;
  {
    bool allow = c->lin_mode == 2 && c->N_E % 1024 == 0 && !c->overlap_mode;
    if (const char *m = getenv("RSG_LIN_PAIR")) allow = allow && atoi(m) != 0;
    std::vector<uint32_t> partner(sp.n_grp, 0xFFFFFFFFu);
    for (uint32_t g = 0; allow && g < sp.n_grp; g++) {
      if (partner[g] != 0xFFFFFFFFu || grp_z[g + 1] == grp_z[g]) continue;
      const uint32_t len = grp_t[g + 1] - grp_t[g], nz = grp_z[g + 1] - grp_z[g];
      for (uint32_t h = g + 1; h < sp.n_grp; h++) {
        if (partner[h] != 0xFFFFFFFFu || grp_t[h + 1] - grp_t[h] != len || grp_z[h + 1] - grp_z[h] != nz) continue;
        uint32_t shared = 0;
        for (uint32_t i = 0; i < len; i++) shared += term[grp_t[g] + i] == term[grp_t[h] + i];
        bool same_cuts = true;
        for (uint32_t k = 0; k < nz; k++)
          same_cuts = same_cuts && zoff[grp_z[g] + k + 1] - zoff[grp_z[g] + k] == zoff[grp_z[h] + k + 1] - zoff[grp_z[h] + k];
        if (!same_cuts || 2 * shared < len) continue;
        partner[g] = h;
        partner[h] = g;
        fp->shared_terms += shared;
        fp->n_paired += 2 * nz;
        break;
      }
    }
    if (fp->n_paired) {   // the pairs first (first, second, first, second, ..), then every other split in plan order
      for (uint32_t g = 0; g < sp.n_grp; g++) {
        const uint32_t h = partner[g];
        if (h == 0xFFFFFFFFu || h < g) continue;
        for (uint32_t k = 0; k < grp_z[g + 1] - grp_z[g]; k++) {
          zorder.push_back(grp_z[g] + k);
          zorder.push_back(grp_z[h] + k);
        }
      }
      for (uint32_t g = 0; g < sp.n_grp; g++)
        if (partner[g] == 0xFFFFFFFFu)
          for (uint32_t k = grp_z[g]; k < grp_z[g + 1]; k++) zorder.push_back(k);
    }
  }
  for (uint32_t z = 0; z < fp->Z; z++) {
    uint32_t a = 0xFFFFFFFFu, b = 0;
    for (uint32_t t = zoff[z]; t < zoff[z + 1]; t++)
      if (pidx[t] != 0xFFFFFFFFu) { a = std::min(a, pidx[t]); b = std::max(b, pidx[t] + 1); }
    fp->zs0.push_back(a == 0xFFFFFFFFu ? 0 : a);
    fp->zs1.push_back(a == 0xFFFFFFFFu ? 0 : b);
  }
  std::vector<uint32_t> idx(fp->n_terms, 0);
  idx.insert(idx.end(), pidx.begin(), pidx.end());
  idx.insert(idx.end(), zoff.begin(), zoff.end());
  idx.insert(idx.end(), zr.begin(), zr.end());
  idx.insert(idx.end(), zorder.begin(), zorder.end());
  void *v = nullptr;
  cudaError_t e = cudaMalloc(&v, idx.size() * 4);
  if (e != cudaSuccess) { delete fp; return fail(RSG_ERR_CUDA, cudaGetErrorString(e)); }
  fp->d_idx = (uint32_t *)v;
  e = cudaMemcpy(fp->d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMalloc(&v, std::max<size_t>(1, term.size()) * 8);
  if (e == cudaSuccess) {
    fp->d_tptr = (const uint64_t **)v;
    e = cudaMemcpy((void *)fp->d_tptr, term.data(), term.size() * 8, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && kind_bytes) {
    e = cudaMalloc(&v, kind_bytes);
    if (e == cudaSuccess) {
      fp->d_kind = (uint8_t *)v;
      size_t at = 0;
      for (uint32_t k = 0; k < sp.n_vec && e == cudaSuccess; k++)
        if (sp.kind[k]) {
          e = cudaMemcpy(fp->d_kind + at, sp.kind[k], sp.kind_len[k], cudaMemcpyHostToDevice);
          T.kind[k] = fp->d_kind + at;
          at += sp.kind_len[k];
        }
    }
  }
  if (e != cudaSuccess) { fast_free_plan(fp); return fail(RSG_ERR_CUDA, cudaGetErrorString(e)); }
  c->fplan = fp;
  *out = fp;
  return RSG_OK;
}
static void key_put(std::vector<uint8_t> &key, const void *p, size_t n) { key.insert(key.end(), (const uint8_t *)p, (const uint8_t *)p + n); }

// CRS references of a ringGroth16 key: contiguous encodings s_pows / delta_ts / delta_mid (element `lo` of each range first) and
// the two bare encodings (nullable) -- one arena with a layout (rsg_groth16_prove) or one arena per vector (rsg_groth16_prove_refs)
struct G16Ptrs { const uint64_t *s_pows, *delta_ts, *delta_mid, *alpha, *beta; };
static G16Ptrs g16_ptrs(const rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L) {
  const size_t NONE = (size_t)-1, E = c->enc_words();
  return G16Ptrs{crs->d + L->s_pows_off * E, crs->d + L->delta_ts_off * E, crs->d + L->delta_mid_off * E,
                 L->alpha_idx == NONE ? nullptr : crs->d + L->alpha_idx * E, L->beta_idx == NONE ? nullptr : crs->d + L->beta_idx * E};
}
static int fast_get_plan(rsg_context *c, const G16Ptrs &G, const rsg_groth16_layout *L, size_t n, size_t n_aux, const uint8_t *h_aux_kind,
                         FastPlan **out) {
  const size_t s_lo = L->s_pows_lo, s_hi = std::min(L->s_pows_hi, n), t_lo = L->delta_ts_lo, t_hi = std::min(L->delta_ts_hi, n + 1);
  const size_t m_lo = L->delta_mid_lo, m_hi = std::min(L->delta_mid_hi, n_aux);
  const uint32_t nS = (uint32_t)(s_hi > s_lo ? s_hi - s_lo : 0), nH = (uint32_t)(t_hi > t_lo ? t_hi - t_lo : 0),
                 nM = (uint32_t)(m_hi > m_lo ? m_hi - m_lo : 0);
  std::vector<uint8_t> key;
  const uint32_t tagk = 0x67313600u | (uint32_t)c->overlap_mode;
  key_put(key, &tagk, 4); key_put(key, L, sizeof(*L)); key_put(key, &n, sizeof(n)); key_put(key, &n_aux, sizeof(n_aux));
  key_put(key, &G, sizeof(G));
  if (h_aux_kind) key_put(key, h_aux_kind + m_lo, nM);
  if (c->fplan && c->fplan->key == key) { *out = c->fplan; return RSG_OK; }
  FastSpec sp;
  sp.n_vec = 6; sp.nS = nS;
  const uint32_t lo[6] = {(uint32_t)s_lo, (uint32_t)s_lo, (uint32_t)s_lo, (uint32_t)s_lo, (uint32_t)t_lo, (uint32_t)m_lo};
  const uint32_t cnt[6] = {nS, nS, nS, nS, nH, nM};
  const uint64_t *base[6] = {G.s_pows, G.s_pows, G.s_pows, G.s_pows, G.delta_ts, G.delta_mid};
  sp.n_ip = 6;
  for (int k = 0; k < 6; k++) { sp.lo[k] = lo[k]; sp.count[k] = cnt[k]; sp.ip_vec[k] = k; sp.ip_base[k] = base[k]; }
  std::vector<uint8_t> all;
  if (h_aux_kind && n_aux) {   // indexed by ABSOLUTE auxiliary index like the host array
    all.assign(n_aux, (uint8_t)RSG_AUX_POLY);
    memcpy(all.data() + m_lo, h_aux_kind + m_lo, nM);
    sp.kind[5] = all.data();
    sp.kind_len[5] = n_aux;
  }
  sp.alpha = G.alpha;
  sp.beta = G.beta;
  // groups A (merged io + mid, + alpha) | B (merged, + beta) | H | aux; outputs A, B, C = H + aux
  sp.n_grp = 4; sp.n_out = 3;
  sp.grp_ip[0][0] = 0; sp.grp_ip[0][1] = 1; sp.grp_merged[0] = true; sp.grp_extra[0] = sp.alpha; sp.grp_out[0] = 0;
  sp.grp_ip[1][0] = 2; sp.grp_ip[1][1] = 3; sp.grp_merged[1] = true; sp.grp_extra[1] = sp.beta; sp.grp_out[1] = 1;
  sp.grp_ip[2][0] = 4; sp.grp_ip[2][1] = 0xFFFFFFFFu; sp.grp_out[2] = 2;
  sp.grp_ip[3][0] = 5; sp.grp_ip[3][1] = 0xFFFFFFFFu; sp.grp_out[3] = 2;
  return fast_build_plan(c, sp, key, out);
}

static bool fast_applies(const rsg_context *c, const rsg_groth16_layout *L, size_t n, size_t n_aux) {
  if (!c->fast_mode || !c->merge_mode) return false;
  const size_t s_lo = L->s_pows_lo, s_hi = std::min(L->s_pows_hi, n), t_lo = L->delta_ts_lo, t_hi = std::min(L->delta_ts_hi, n + 1);
  const size_t m_lo = L->delta_mid_lo, m_hi = std::min(L->delta_mid_hi, n_aux);
  const size_t nS = s_hi > s_lo ? s_hi - s_lo : 0, nH = t_hi > t_lo ? t_hi - t_lo : 0, nM = m_hi > m_lo ? m_hi - m_lo : 0;
  const size_t slots = 2 * nS + nH + nM;
  if (!slots || slots > (1u << 24)) return false;
  return slots * c->L_R * c->L_E * c->N_E <= c->pntt_budget_words;   // larger term sets go through the chunked path
}

template <int LG, int LV>
static int fast_set_attr() {
  static bool done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev]) return RSG_OK;
  CUDA_TRY(cudaFuncSetAttribute(k_encode_fast<LG, LV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)padded_words(1u << LG) * 8));
  done[dev] = true;
  return RSG_OK;
}

// NTT slots [s0, s0 + count) of the fast path
static int fast_launch_ntt(rsg_context *c, const uint64_t *nttsrc, uint32_t s0, uint32_t count, const uint8_t *slot_skip, bool lowreg = false) {
  if (!count) return RSG_OK;
  const size_t poly = c->L_R * c->N_E, per_general = poly * c->L_E;
  const unsigned th = ntt_threads(c->logN);
  const size_t sm = ntt_smem(c->logN);
  const unsigned split = c->logN > 14 ? 2 : 1;
  c->st_fwd_polys += (uint64_t)count * c->L_R * c->L_E;
  dim3 grid((unsigned)(count * c->L_E * split), (unsigned)c->L_R);
  bool lazy = true;
  for (uint64_t p : c->Q) lazy = lazy && p < (1ull << 58);
  const uint64_t *src = nttsrc + (size_t)s0 * poly;
  uint64_t *dst = c->d_pntt + (size_t)s0 * per_general;
  const uint8_t *sk = slot_skip + s0;
  LaunchScope ls(c, "k_lift_fwd_ntt");
  if (c->ntt_half && c->logN == 14 && c->f64_ntt && c->ntt_mode != 1) {
    // experiment (RSG_NTT_HALF=1): two 2^13-point CTAs of 256 threads per polynomial (first level fused into the load, as at
    // N_E = 2^15) -- two independent CTAs per SM instead of one, so one CTA's barriers and loads hide behind the other's math
    int rc = set_smem_attrs<13, 1>();
    if (rc) return rc;
    k_lift_fwd_ntt_f64<13, 1, true><<<dim3((unsigned)(count * c->L_E * 2), (unsigned)c->L_R), 256, (size_t)padded_words(1u << 13) * 8, c->stream>>>(
        c->d_params, src, dst, sk);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  if (c->f64_ntt && c->ntt_mode != 1 && c->logN == 14 && c->ntt_cluster && !lowreg) {
    const dim3 cgrid((unsigned)(count * c->L_E * NTT_CL), (unsigned)c->L_R);
    if (c->lift_smallq) k_lift_fwd_ntt_f64_cl<true, true><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, src, dst, sk);
    else k_lift_fwd_ntt_f64_cl<true, false><<<cgrid, NTT_CL_THREADS, NTT_CL_SMEM, c->stream>>>(c->d_params, src, dst, sk);
    CUDA_TRY(cudaGetLastError());
    return RSG_OK;
  }
  if (c->f64_ntt && c->ntt_mode != 1) {
    DISPATCH_LOGN(c->logN, { int rc = set_smem_attrs<LG, LV>(); if (rc) return rc;
                             if (c->lift_smallq) {
                               if (lowreg) k_lift_fwd_ntt_f64_r96<LG, LV, true, true><<<grid, th, sm, c->stream>>>(c->d_params, src, dst, sk);
                               else k_lift_fwd_ntt_f64<LG, LV, true, true><<<grid, th, sm, c->stream>>>(c->d_params, src, dst, sk);
                             } else if (lowreg) k_lift_fwd_ntt_f64_r96<LG, LV, true><<<grid, th, sm, c->stream>>>(c->d_params, src, dst, sk);
                             else k_lift_fwd_ntt_f64<LG, LV, true><<<grid, th, sm, c->stream>>>(c->d_params, src, dst, sk); });
  } else if (lazy && c->ntt_cluster && !lowreg && (c->logN == 15 || c->logN == 14)) {
    const size_t smc = 4 * (1024 + 64) * 8;
    const dim3 cgrid((unsigned)(count * c->L_E * (c->logN == 15 ? 8 : 4)), (unsigned)c->L_R);
    if (c->logN == 15) k_lift_fwd_ntt_int_cl<5, true><<<cgrid, 128, smc, c->stream>>>(c->d_params, src, dst, sk);
    else k_lift_fwd_ntt_int_cl<4, true><<<cgrid, 128, smc, c->stream>>>(c->d_params, src, dst, sk);
  } else {
    DISPATCH_LOGN(c->logN, { int rc = set_smem_attrs<LG, LV>(); if (rc) return rc;
                             if (lazy) k_lift_fwd_ntt<LG, LV, true><<<grid, th, sm, c->stream>>>(c->d_params, src, 1u, dst, sk);
                             else k_lift_fwd_ntt<LG, LV, false><<<grid, th, sm, c->stream>>>(c->d_params, src, 1u, dst, sk); });
  }
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

// splits [z0, z1) of the plan's term chunks -> partial[z]
static int fast_launch_lincomb(rsg_context *c, const FastPlan *fp, uint32_t z0, uint32_t z1, const uint8_t *slot_skip,
                               cudaStream_t st, bool lowreg = false) {
  const uint64_t *d_crs = nullptr;   // every term carries its own pointer
  if (z1 <= z0) return RSG_OK;
  const uint32_t *d_term = fp->d_idx, *d_pidx = fp->d_idx + fp->n_terms, *d_zoff = fp->d_idx + 2 * fp->n_terms;
  uint64_t *partial = c->d_partial + (size_t)z0 * c->enc_words();
  c->st_lin_launches++;
  LaunchScope ls(c, "k_crs_lincomb", st);
  if (c->lin_mode == 1 && c->N_E % LT_XC == 0) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev]) {
      CUDA_TRY(cudaFuncSetAttribute(k_crs_lincomb_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM));
      CUDA_TRY(cudaFuncSetAttribute(k_crs_lincomb_tma, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      attr_done[dev] = true;
    }
    const size_t items = (size_t)(z1 - z0) * c->L_R * c->L_E * (c->N_E / LT_XC);
    const unsigned grid = (unsigned)std::min<size_t>(items, (size_t)148 * c->lt_ctas);
    k_crs_lincomb_tma<<<grid, LT_THREADS, LT_SMEM, st>>>(c->d_params, fp->d_tptr, d_pidx, d_zoff + z0, z1 - z0, slot_skip, c->d_pntt, partial);
  } else {
    if (c->overlap_mode) {
      // the transform kernel needs the maximum shared-memory carve-out; a kernel that prefers another split of the L1/shared
      // array cannot become resident on the same SM at the same time -- ask for the same carve-out
      static bool carve_done[64] = {};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!carve_done[dev]) {
        CUDA_TRY(cudaFuncSetAttribute(k_crs_lincomb_r64, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CUDA_TRY(cudaFuncSetAttribute(k_crs_lincomb<2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        carve_done[dev] = true;
      }
    }
    const unsigned th = (unsigned)std::min<size_t>(c->lin_threads > 0 ? c->lin_threads : 256, c->N_E / 2);
    const dim3 grid((unsigned)(c->N_E / 2 / th), (unsigned)(c->L_R * c->L_E), z1 - z0);
    if (c->lin_mode == 2 && c->N_E % 1024 == 0) {   // four x per thread, 256-bit loads (default)
      const dim3 gw((unsigned)(c->N_E / 4 / 256), (unsigned)(c->L_R * c->L_E), z1 - z0);
      if (c->lin_unroll == 1) k_crs_lincomb_wide<1><<<gw, 256, 0, st>>>(c->d_params, d_pidx, c->d_pntt, partial, d_zoff + z0, slot_skip, fp->d_tptr);
      else if (c->lin_unroll == 3) k_crs_lincomb_wide<3><<<gw, 256, 0, st>>>(c->d_params, d_pidx, c->d_pntt, partial, d_zoff + z0, slot_skip, fp->d_tptr);
      else if (fp->n_paired && z0 == 0 && z1 == fp->Z)
        // (launching the pairs as clusters of two, co-scheduled by construction, measured the same: 2.125 vs 2.116 ms)
        k_crs_lincomb_wide<2><<<gw, 256, 0, st>>>(c->d_params, d_pidx, c->d_pntt, partial, d_zoff, slot_skip, fp->d_tptr,
                                                  d_zoff + (fp->Z + 1) + (fp->n_out + 1), fp->n_paired);
      else k_crs_lincomb_wide<2><<<gw, 256, 0, st>>>(c->d_params, d_pidx, c->d_pntt, partial, d_zoff + z0, slot_skip, fp->d_tptr);
    } else if (lowreg) k_crs_lincomb_r64<<<grid, th, 0, st>>>(c->d_params, d_crs, d_term, d_pidx, fp->n_terms, 0u, c->d_pntt, partial, d_zoff + z0, slot_skip,
                                                       fp->d_tptr);
    else k_crs_lincomb<2><<<grid, th, 0, st>>>(c->d_params, d_crs, d_term, d_pidx, fp->n_terms, 0u, c->d_pntt, partial, d_zoff + z0, slot_skip,
                                               fp->d_tptr);
  }
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

// Everything up to the probes for one plan: flags -> batch encode -> (merged pairs) -> forward NTTs -> the one-launch lincomb ->
// per-output sums into `out` (fp->n_out encodings) -> probe totals.  No synchronisation.
static int fast_run(rsg_context *c, FastPlan *fp, const FastTable &T, uint64_t *out) {
  int rc;
  const size_t poly = c->L_R * c->N_E, per_general = poly * c->L_E, E = c->enc_words();
  if ((rc = ensure(c, &c->d_fp_flags, &c->cap_fp_flags, (size_t)T.n_elems + T.n_slots + 64))) return rc;
  if ((rc = ensure(c, &c->d_fp_parts, &c->cap_fp_parts, std::max<size_t>(1, (size_t)T.n_parts * poly)))) return rc;
  if ((rc = ensure(c, &c->d_fp_nttsrc, &c->cap_fp_nttsrc, (size_t)T.n_slots * poly))) return rc;
  if ((rc = ensure(c, &c->d_pntt, &c->cap_pntt, (size_t)T.n_slots * per_general))) return rc;
  if ((rc = ensure(c, &c->d_pval, &c->cap_pval, std::max<size_t>((size_t)T.n_parts * c->L_R, 1024)))) return rc;
  if ((rc = ensure(c, &c->d_partial, &c->cap_partial, (size_t)fp->Z * E))) return rc;
  if (!c->d_fp_status) {
    void *v = nullptr;
    CUDA_TRY(cudaMalloc(&v, FPS_WORDS * 4));
    c->d_fp_status = (uint32_t *)v;
    CUDA_TRY(cudaMalloc(&v, FP_MAXIP * MAX_LR * 8));
    c->d_fp_totals = (uint64_t *)v;
    CUDA_TRY(cudaHostAlloc(&v, FPS_WORDS * 4, cudaHostAllocDefault));
    c->h_fp_status = (uint32_t *)v;
  }
  uint8_t *elem_flag = c->d_fp_flags, *slot_skip = c->d_fp_flags + T.n_elems;
  cudaStream_t st = c->stream;
  CUDA_TRY(cudaMemsetAsync(c->d_fp_status, 0, FPS_WORDS * 4, st));
  {
    LaunchScope ls(c, "k_is_zero_prefix");
    k_term_flags<<<T.n_elems, 256, 0, st>>>(c->d_params, T, elem_flag, slot_skip, c->d_fp_status);
  }
  const unsigned th = ntt_threads(c->logN);
  const size_t sm = ntt_smem(c->logN);
  const unsigned split = c->logN > 14 ? 2 : 1;
  c->st_inv_polys += (uint64_t)T.n_elems * c->L_R;
  if (c->enc_rows && c->logN == 14 && 2 * c->N_R <= c->N_E) {
    LaunchScope ls(c, "k_encode_intt");
    static bool attr_done[64] = {};
    if (!attr_done[c->device]) {
      CUDA_TRY(cudaFuncSetAttribute(k_encode_fast_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ENC_ROWS_SMEM));
      attr_done[c->device] = true;
    }
    k_encode_fast_rows<<<dim3(T.n_elems, (unsigned)c->L_R), 256, ENC_ROWS_SMEM, st>>>(c->d_params, T, elem_flag, c->d_fp_parts, c->d_fp_nttsrc);
  } else {
    LaunchScope ls(c, "k_encode_intt");
    DISPATCH_LOGN(c->logN, { if ((rc = fast_set_attr<LG, LV>())) return rc;
                             k_encode_fast<LG, LV><<<dim3(T.n_elems * split, (unsigned)c->L_R), th, sm, st>>>(c->d_params, T, elem_flag, c->d_fp_parts,
                                                                                                          c->d_fp_nttsrc); });
  }
  CUDA_TRY(cudaGetLastError());
  if (split > 1) {   // last inverse level + N^-1 + centring of the two plaintext regions
    if (T.n_parts && (rc = launch_intt_finish(c, c->d_fp_parts, (size_t)T.n_parts * c->L_R, c->d_modq, c->d_invN_q, c->d_invNw_q, (uint32_t)c->L_R,
                                             0xFFFFFFFFu, 1u)))
      return rc;
    const uint32_t plain_slots = T.n_slots - 2 * T.nS;
    if (plain_slots && (rc = launch_intt_finish(c, c->d_fp_nttsrc + (size_t)2 * T.nS * poly, (size_t)plain_slots * c->L_R, c->d_modq, c->d_invN_q,
                                               c->d_invNw_q, (uint32_t)c->L_R, 0xFFFFFFFFu, 1u)))
      return rc;
  }
  if (T.nS) {
    LaunchScope ls(c, "k_centre_add");
    // the probe sums fit a signed 128-bit accumulator when bits(t) + bits(Q_0) + log2 N_E <= 126
    uint64_t tmax = 0;
    for (uint64_t p : c->q) tmax = std::max(tmax, p);
    const int tb = 64 - __builtin_clzll(tmax), qb = 64 - __builtin_clzll(c->Q[0]);
    const uint32_t s128 = tb + qb + c->logN <= 126 ? 1u : 0u;
    k_centre_add_fast<<<dim3(2 * T.nS, (unsigned)c->L_R), 256, 0, st>>>(c->d_params, T, elem_flag, c->d_fp_parts, c->d_fp_nttsrc, slot_skip,
                                                                        c->d_psi_pow, c->d_pval, s128);
    CUDA_TRY(cudaGetLastError());
  }
  c->st_lin_terms += fp->n_terms;
  c->st_lin_shared += fp->shared_terms;
  c->st_lin_plain += T.n_slots;
  c->st_merged += T.nS ? 2 : 0;
  if (c->overlap_mode) {
    // Phases of whole term chunks: the transforms of phase p+1 run on the main stream while the HBM-bound lincomb of phase p
    // streams on the second one.  RSG_OVERLAP=1 uses the register-capped kernel pair (96 x 512 + 64 x 256 = 64 Ki registers:
    // one lincomb CTA fits next to the transform CTA on an SM), RSG_OVERLAP=2 the full-register kernels.  Measured on B200
    // (C4): neither shortens the proof -- kept as an experiment, not the default.
    if (!c->stream2) {
      CUDA_TRY(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
      for (auto &e : c->ev_phase) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CUDA_TRY(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    }
    const bool lowreg = c->overlap_mode == 1;
    const uint32_t per_phase = (uint32_t)std::max(1, c->overlap_chunks);
    uint32_t ntt_done = 0;   // slots [0, ntt_done) have been launched
    int pe = 0;
    for (uint32_t z0 = 0; z0 < fp->Z; z0 += per_phase, pe = (pe + 1) % 8) {
      const uint32_t z1 = std::min(fp->Z, z0 + per_phase);
      uint32_t s_end = ntt_done;
      for (uint32_t z = z0; z < z1; z++) s_end = std::max(s_end, fp->zs1[z]);
      if (z1 == fp->Z) s_end = T.n_slots;
      if (s_end > ntt_done && (rc = fast_launch_ntt(c, c->d_fp_nttsrc, ntt_done, s_end - ntt_done, slot_skip, lowreg))) return rc;
      ntt_done = s_end;
      CUDA_TRY(cudaEventRecord(c->ev_phase[pe], st));
      CUDA_TRY(cudaStreamWaitEvent(c->stream2, c->ev_phase[pe], 0));
      if ((rc = fast_launch_lincomb(c, fp, z0, z1, slot_skip, c->stream2, lowreg))) return rc;
    }
    CUDA_TRY(cudaEventRecord(c->ev_join, c->stream2));
    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_join, 0));
  } else {
    if ((rc = fast_launch_ntt(c, c->d_fp_nttsrc, 0, T.n_slots, slot_skip))) return rc;
    if ((rc = fast_launch_lincomb(c, fp, 0, fp->Z, slot_skip, st))) return rc;
  }
  {
    LaunchScope ls(c, "k_enc_sum");
    const size_t pairs = E / 2;
    k_enc_sum_ranges<<<dim3((unsigned)((pairs + 255) / 256), fp->n_out), 256, 0, st>>>(c->d_params, c->d_partial,
                                                                                      fp->d_idx + 2 * fp->n_terms + fp->Z + 1, out);
  }
  {
    LaunchScope ls(c, "k_probe");
    k_probe_fast<<<dim3(T.n_ip, (unsigned)c->L_R), 256, 0, st>>>(c->d_params, T, elem_flag, c->d_pval, c->d_pntt, c->d_fp_totals,
                                                                 c->fp_block ? c->fp_block + RSG_SHARD_BLOCK_HEADER + 8 * c->L_R : nullptr, c->fp_pstride,
                                                                 c->d_fp_status);
  }
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

// Returns RSG_OK with *candidate = 1 when a probe sum vanished (the caller then runs the exact path).  Synchronises once.
static int groth16_lincombs_fast(rsg_context *c, const G16Ptrs &G, const rsg_groth16_layout *L, size_t n, size_t n_aux,
                                 const uint64_t *const vec[6], const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *out,
                                 size_t *n_used, int *candidate) {
  int rc;
  FastPlan *fp;
  if ((rc = fast_get_plan(c, G, L, n, n_aux, h_aux_kind, &fp))) return rc;
  FastTable T = fp->T;
  for (int k = 0; k < 6; k++) T.vec[k].base = vec[k];
  if ((rc = fast_run(c, fp, T, out))) return rc;
  cudaStream_t st = c->stream;
  {
    LaunchScope ls(c, "k_probe");
    k_probe_chain<<<1, 32, 0, st>>>(c->d_params, T, c->d_fp_totals, c->d_fp_status);
  }
  if (c->fp_block) {
    LaunchScope ls(c, "k_probe");
    k_probe_block_header<<<1, 32, 0, st>>>(c->d_params, T, c->d_fp_totals, c->d_fp_status, c->fp_pstride, c->fp_block);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(c->h_fp_status, c->d_fp_status, FPS_WORDS * 4, cudaMemcpyDeviceToHost, st));
  if (h_proof) CUDA_TRY(cudaMemcpyAsync(h_proof, out, 3 * c->enc_words() * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint32_t *hs = c->h_fp_status;
  *candidate = hs[FPS_CANDIDATE] != 0;
  if (n_used) {
    n_used[0] = (size_t)hs[FPS_COUNT0 + 0] + hs[FPS_COUNT0 + 1] + (T.alpha != nullptr);
    n_used[1] = (size_t)hs[FPS_COUNT0 + 2] + hs[FPS_COUNT0 + 3] + (T.beta != nullptr);
    n_used[2] = (size_t)hs[FPS_COUNT0 + 4] + hs[FPS_COUNT0 + 5];
  }
  c->st_fast++;
  return RSG_OK;
}

// The six term lists of groth16.tcc:89-112 over this shard's ranges: which coefficients SealPoly::is_zero skips (one flag
// kernel per vector range, one copy; synchronises), scalar auxiliary inputs by kind.
static int g16_term_lists(rsg_context *c, const rsg_groth16_layout *L, size_t n, size_t n_aux, const uint64_t *const vec[6],
                          const uint8_t *h_aux_kind, std::vector<TermSpec> (&ip)[6]) {
  int rc;
  const size_t W = c->ring_words();
  struct Range { size_t lo, hi; };
  const Range rg[6] = {{L->s_pows_lo, std::min(L->s_pows_hi, n)}, {L->s_pows_lo, std::min(L->s_pows_hi, n)},
                       {L->s_pows_lo, std::min(L->s_pows_hi, n)}, {L->s_pows_lo, std::min(L->s_pows_hi, n)},
                       {L->delta_ts_lo, std::min(L->delta_ts_hi, n + 1)}, {L->delta_mid_lo, std::min(L->delta_mid_hi, n_aux)}};
  size_t foff[7] = {0};
  for (int k = 0; k < 6; k++) foff[k + 1] = foff[k] + (rg[k].hi > rg[k].lo ? rg[k].hi - rg[k].lo : 0);
  if ((rc = ensure(c, &c->d_flags, &c->cap_flags, std::max<size_t>(foff[6], 1)))) return rc;
  for (int k = 0; k < 6; k++)
    if (foff[k + 1] > foff[k]) {
      LaunchScope ls(c, "k_is_zero_prefix");
      k_is_zero_prefix<<<(unsigned)(foff[k + 1] - foff[k]), 256, 0, c->stream>>>(vec[k] + rg[k].lo * W, (uint32_t)W, c->d_flags + foff[k]);
    }
  CUDA_TRY(cudaGetLastError());
  std::vector<uint8_t> flags(std::max<size_t>(foff[6], 1));
  if (foff[6]) CUDA_TRY(cudaMemcpyAsync(flags.data(), c->d_flags, foff[6], cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  const size_t crs_off[6] = {L->s_pows_off, L->s_pows_off, L->s_pows_off, L->s_pows_off, L->delta_ts_off, L->delta_mid_off};
  for (int k = 0; k < 5; k++)
    for (size_t i = rg[k].lo; i < rg[k].hi; i++)
      if (!flags[foff[k] + i - rg[k].lo]) ip[k].push_back({(uint32_t)(crs_off[k] + i - rg[k].lo), vec[k], (uint32_t)i});
  for (size_t i = rg[5].lo; i < rg[5].hi; i++) {
    const uint8_t kind = h_aux_kind ? h_aux_kind[i] : (uint8_t)RSG_AUX_POLY;
    const uint32_t ci = (uint32_t)(crs_off[5] + i - rg[5].lo);
    if (kind == RSG_AUX_POLY) {
      if (!flags[foff[5] + i - rg[5].lo]) ip[5].push_back({ci, vec[5], (uint32_t)i});
    } else if (kind == RSG_TERM_ONE) {
      ip[5].push_back({ci, nullptr, 0});
    } else if (kind == RSG_TERM_GENERAL) {
      ip[5].push_back({ci, vec[5], (uint32_t)i});
    }
  }
  return RSG_OK;
}

// ------------------------------------------------------------------------------------------------------------
// groth16::prover (groth16.tcc:69-115): the witness map, then the reference's own sequence of six inner products combined
// with operator+= (kept separate, not fused into three, because the transparent-ciphertext rule is order-dependent).
// The six inner products + operator+= chain of groth16.tcc:89-112 over this shard's term ranges.  vec[k] (k = A_io, A_mid,
// B_io, B_mid, H, aux) is the address element 0 of that coefficient vector WOULD have: only the elements of the shard's
// ranges ([s_lo, min(s_hi, n)) for the first four, [t_lo, min(t_hi, n+1)) for H, [m_lo, min(m_hi, n_aux)) for aux) are read.
static int groth16_lincombs_exact(rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L, size_t n, size_t n_aux,
                                  const uint64_t *const vec[6], const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                                  size_t *n_used) {
  int rc;
  const size_t NONE = (size_t)-1;
  std::vector<TermSpec> ip[6];
  if ((rc = g16_term_lists(c, L, n, n_aux, vec, h_aux_kind, ip))) return rc;
  uint64_t *out = d_proof;
  if (!out) {
    if ((rc = ensure(c, &c->d_out_scratch, &c->cap_out_scratch, 3 * c->enc_words()))) return rc;
    out = c->d_out_scratch;
  }
  const size_t E = c->enc_words();
  if (n_used) {
    n_used[0] = ip[0].size() + ip[1].size() + (L->alpha_idx != NONE);
    n_used[1] = ip[2].size() + ip[3].size() + (L->beta_idx != NONE);
    n_used[2] = ip[4].size() + ip[5].size();
  }
  size_t probe_off[7] = {0};
  for (int k = 0; k < 6; k++) probe_off[k + 1] = probe_off[k] + c->L_R * ip[k].size();
  if ((rc = ensure(c, &c->d_ip, &c->cap_ip, 6 * E))) return rc;
  if ((rc = ensure(c, &c->d_probe, &c->cap_probe, std::max<size_t>(probe_off[6], 4096)))) return rc;
  // A and B: <s_pows, x_io> + <s_pows, x_mid> share their CRS range -> one merged pass each (lincomb_merged)
  bool merged[3] = {false, false, false};
  for (int e = 0; e < 2 && c->merge_mode; e++) {
    rc = lincomb_merged(c, crs->d, ip[2 * e], ip[2 * e + 1], c->d_ip + 2 * e * E, c->d_probe + probe_off[2 * e],
                        c->d_probe + probe_off[2 * e + 1]);
    if (rc < 0) return rc;
    merged[e] = rc == 0;
  }
  for (int k = 0; k < 6; k++)
    if (!(k < 4 && merged[k / 2]) && !ip[k].empty() &&
        (rc = lincomb_terms(c, crs->d, ip[k], c->d_ip + k * E, c->d_probe + probe_off[k])))
      return rc;
  // EncodingElem::operator+= chain of one proof element: copy the first non-empty operand, then add with the
  // transparent-result rule (an empty inner product is the additive identity, seal_ring.tcc:482-488)
  auto combine = [&](int e) -> int {
    const uint64_t *src[3] = {nullptr, nullptr, nullptr};
    const int a = e == 2 ? 4 : 2 * e, b = a + 1;
    if (!ip[a].empty()) src[0] = c->d_ip + a * E;
    if (!ip[b].empty() && !merged[e]) src[1] = c->d_ip + b * E;   // merged: d_ip[a] already holds ip[a] + ip[b]
    const size_t extra = e == 0 ? L->alpha_idx : (e == 1 ? L->beta_idx : NONE);
    if (extra != NONE) src[2] = crs->d + extra * E;
    uint64_t *acc = out + e * E;
    bool have = false;
    for (int k = 0; k < 3; k++) {
      if (!src[k]) continue;
      if (!have) {
        CUDA_TRY(cudaMemcpyAsync(acc, src[k], E * 8, cudaMemcpyDeviceToDevice, c->stream));
        have = true;
        if (k == 0 && merged[e]) {
          int r = transparent_fix(c, acc);
          if (r) return r;
        }
      } else {
        int r = enc_add_fix(c, acc, src[k]);
        if (r) return r;
      }
    }
    if (!have) CUDA_TRY(cudaMemsetAsync(acc, 0, E * 8, c->stream));
    return RSG_OK;
  };
  for (int e = 0; e < 3; e++)
    if ((rc = combine(e))) return rc;
  if (probe_off[6]) {
    std::vector<uint8_t> pf(probe_off[6]);
    CUDA_TRY(cudaMemcpyAsync(pf.data(), c->d_probe, probe_off[6], cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    bool redo[3] = {false, false, false};
    // a flagged prefix inside a merged pair: form the two inner products separately after all, then resolve as usual
    for (int e = 0; e < 2; e++) {
      if (!merged[e]) continue;
      bool any = false;
      for (size_t i = probe_off[2 * e]; i < probe_off[2 * e + 2]; i++) any |= pf[i] != 0;
      if (!any) continue;
      merged[e] = false;
      redo[e] = true;
      for (int k = 2 * e; k < 2 * e + 2; k++)
        if ((rc = lincomb_terms(c, crs->d, ip[k], c->d_ip + k * E, c->d_probe + probe_off[k]))) return rc;
      CUDA_TRY(cudaMemcpyAsync(pf.data() + probe_off[2 * e], c->d_probe + probe_off[2 * e], probe_off[2 * e + 2] - probe_off[2 * e],
                               cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(cudaStreamSynchronize(c->stream));
    }
    for (int k = 0; k < 6; k++) {
      bool any = false;
      for (size_t i = probe_off[k]; i < probe_off[k + 1]; i++) any |= pf[i] != 0;
      if (!any) continue;
      if ((rc = inner_product_exact(c, crs->d, ip[k], pf.data() + probe_off[k], c->d_ip + k * E))) return rc;
      redo[k == 4 || k == 5 ? 2 : k / 2] = true;
    }
    for (int e = 0; e < 3; e++)
      if (redo[e] && (rc = combine(e))) return rc;
  }
  if (h_proof) CUDA_TRY(cudaMemcpyAsync(h_proof, out, 3 * c->enc_words() * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}

// The static launch sequence first; the host-driven exact path when it does not apply (RSG_FAST=0, term set above the
// NTT-plaintext budget) or when a probe sum vanished (transparent-ciphertext candidate, seal_ring.tcc:493-504).
static int groth16_lincombs_dev(rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L, size_t n, size_t n_aux,
                                const uint64_t *const vec[6], const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                                size_t *n_used) {
  if (fast_applies(c, L, n, n_aux)) {
    int rc;
    uint64_t *out = d_proof;
    if (!out) {
      if ((rc = ensure(c, &c->d_out_scratch, &c->cap_out_scratch, 3 * c->enc_words()))) return rc;
      out = c->d_out_scratch;
    }
    int candidate = 0;
    if ((rc = groth16_lincombs_fast(c, g16_ptrs(c, crs, L), L, n, n_aux, vec, h_aux_kind, h_proof, out, n_used, &candidate))) return rc;
    if (!candidate) return RSG_OK;
    c->st_fast_fallback++;
  }
  return groth16_lincombs_exact(c, crs, L, n, n_aux, vec, h_aux_kind, h_proof, d_proof, n_used);
}

extern "C" int rsg_groth16_lincombs(rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L, size_t n, size_t n_aux,
                                    const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                                    size_t *n_used) {
  RSG_TRACE_CALL();
  if (!c || !crs || !L || !d_vec) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  // shard pointers address the FIRST element of each range; rebase them to element 0 (never dereferenced below the range)
  const size_t W = c->ring_words();
  const uint64_t *vec[6];
  const size_t lo[6] = {L->s_pows_lo, L->s_pows_lo, L->s_pows_lo, L->s_pows_lo, L->delta_ts_lo, L->delta_mid_lo};
  for (int k = 0; k < 6; k++) vec[k] = d_vec[k] ? d_vec[k] - lo[k] * W : nullptr;
  return groth16_lincombs_dev(c, crs, L, n, n_aux, vec, h_aux_kind, h_proof, d_proof, n_used);
}

// ---- term shards with the order-dependent transparent-ciphertext rule kept exact (include/rsgpu.h) ----------------------
extern "C" size_t rsg_groth16_shard_block_words(size_t L_R, size_t pstride) { return RSG_SHARD_BLOCK_HEADER + 8 * L_R + 6 * L_R * pstride; }

static void g16_rebase(const rsg_context *c, const rsg_groth16_layout *L, const uint64_t *const d_vec[6], const uint64_t *vec[6]) {
  const size_t W = c->ring_words();
  const size_t lo[6] = {L->s_pows_lo, L->s_pows_lo, L->s_pows_lo, L->s_pows_lo, L->delta_ts_lo, L->delta_mid_lo};
  for (int k = 0; k < 6; k++) vec[k] = d_vec[k] ? d_vec[k] - lo[k] * W : nullptr;
}

extern "C" int rsg_groth16_lincombs_shard(rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L, size_t n, size_t n_aux,
                                          const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *d_part, size_t pstride,
                                          size_t *n_used) {
  RSG_TRACE_CALL();
  if (!c || !crs || !L || !d_vec || !d_part) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t cnt[3] = {std::min(L->s_pows_hi, n) > L->s_pows_lo ? std::min(L->s_pows_hi, n) - L->s_pows_lo : 0,
                         std::min(L->delta_ts_hi, n + 1) > L->delta_ts_lo ? std::min(L->delta_ts_hi, n + 1) - L->delta_ts_lo : 0,
                         std::min(L->delta_mid_hi, n_aux) > L->delta_mid_lo ? std::min(L->delta_mid_hi, n_aux) - L->delta_mid_lo : 0};
  if (cnt[0] > pstride || cnt[1] > pstride || cnt[2] > pstride || pstride > 0xFFFFFFFFull) return fail(RSG_ERR_ARG, "pstride below a term range");
  const uint64_t *vec[6];
  g16_rebase(c, L, d_vec, vec);
  const size_t E = c->enc_words(), bw = rsg_groth16_shard_block_words(c->L_R, pstride);
  uint64_t *block = d_part + 3 * E;
  CUDA_TRY(cudaMemsetAsync(block, 0xFF, bw * 8, c->stream));   // ~0: no term at this position
  if (n_used) n_used[0] = n_used[1] = n_used[2] = 0;
  const bool no_terms = !cnt[0] && !cnt[1] && !cnt[2] && L->alpha_idx == (size_t)-1 && L->beta_idx == (size_t)-1;
  if (no_terms || !fast_applies(c, L, n, n_aux)) {
    // a rank without terms (more ranks than terms) contributes nothing; without a static plan (RSG_FAST=0, term set above the
    // NTT-plaintext budget) the block asks for the chain
    const uint64_t hdr[RSG_SHARD_BLOCK_HEADER] = {no_terms ? 0ull : 1ull, 0, 0, 0, 0, 0, 0, (uint64_t)pstride};
    CUDA_TRY(cudaMemsetAsync(d_part, 0, 3 * E * 8, c->stream));
    CUDA_TRY(cudaMemcpyAsync(block, hdr, sizeof(hdr), cudaMemcpyHostToDevice, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return RSG_OK;
  }
  c->fp_block = block;
  c->fp_pstride = (uint32_t)pstride;
  int candidate = 0;   // local only: the global check (rsg_groth16_shard_check) decides
  const int rc = groth16_lincombs_fast(c, g16_ptrs(c, crs, L), L, n, n_aux, vec, h_aux_kind, nullptr, d_part, n_used, &candidate);
  c->fp_block = nullptr;
  c->fp_pstride = 0;
  return rc;
}

extern "C" int rsg_groth16_shard_check(const uint64_t *h_blocks, size_t world, size_t L_R, size_t pstride, uint64_t Q0, int *verdict) {
  if (!h_blocks || !verdict || !world || !L_R || L_R > (size_t)MAX_LR || Q0 < 2) return fail(RSG_ERR_ARG, "bad argument");
  const size_t bw = rsg_groth16_shard_block_words(L_R, pstride);
  bool cand = false;
  uint64_t gtot[6][MAX_LR] = {}, gcount[6] = {};
  for (size_t r = 0; r < world; r++) {
    const uint64_t *b = h_blocks + r * bw;
    if (b[7] != pstride) return fail(RSG_ERR_ARG, "probe block of another pstride");
    cand = cand || (b[0] & 1);
    for (int k = 0; k < 6; k++) gcount[k] += b[1 + k];
  }
  // global prefix sums: rank r's running sums shifted by the totals of ranks 0..r-1 (term ranges are in rank order)
  for (int ip = 0; ip < 6 && !cand; ip++)
    for (size_t j = 0; j < L_R; j++) {
      uint64_t off = 0;
      for (size_t r = 0; r < world; r++) {
        const uint64_t *b = h_blocks + r * bw;
        const uint64_t *pre = b + RSG_SHARD_BLOCK_HEADER + 8 * L_R + ((size_t)ip * L_R + j) * pstride;
        for (size_t t = 0; t < pstride; t++)
          if (pre[t] != ~0ull && (pre[t] + off) % Q0 == 0) cand = true;
        if (b[1 + ip]) off = (off + b[RSG_SHARD_BLOCK_HEADER + ip * L_R + j]) % Q0;
      }
      gtot[ip][j] = off;
    }
  // the operator+= chains of groth16.tcc:89-112 at the probe slot, on the global totals (k_probe_chain on one GPU)
  for (int e = 0; e < 3 && !cand; e++) {
    const int a = 2 * e, b2 = a + 1;
    if (gcount[a] + gcount[b2] == 0) continue;
    for (size_t j = 0; j < L_R; j++) {
      uint64_t sum = (gtot[a][j] + gtot[b2][j]) % Q0;
      cand = cand || sum == 0;
      if (e < 2)
        for (size_t r = 0; r < world; r++) {
          const uint64_t x = h_blocks[r * bw + RSG_SHARD_BLOCK_HEADER + 6 * L_R + (size_t)e * L_R + j];
          if (x == ~0ull) continue;
          sum = (sum + x) % Q0;
          cand = cand || sum == 0;
        }
    }
  }
  *verdict = cand ? 1 : 0;
  return RSG_OK;
}

extern "C" int rsg_groth16_lincombs_chain(rsg_context *c, rsg_crs *crs, size_t carry_first, const rsg_groth16_layout *L, size_t n,
                                          size_t n_aux, const uint64_t *const d_vec[6], const uint8_t *h_aux_kind, uint64_t *d_carry,
                                          uint8_t *h_present) {
  RSG_TRACE_CALL();
  if (!c || !crs || !L || !d_vec || !d_carry || !h_present) return fail(RSG_ERR_ARG, "null argument");
  if (carry_first + 6 > crs->n) return fail(RSG_ERR_ARG, "the arena needs six spare encodings from carry_first on");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc;
  const uint64_t *vec[6];
  g16_rebase(c, L, d_vec, vec);
  std::vector<TermSpec> ip[6];
  if ((rc = g16_term_lists(c, L, n, n_aux, vec, h_aux_kind, ip))) return rc;
  const size_t E = c->enc_words();
  if ((rc = ensure(c, &c->d_ip, &c->cap_ip, 6 * E))) return rc;
  for (int k = 0; k < 6; k++) {
    std::vector<TermSpec> terms;
    if (h_present[k]) {   // everything before this rank's range enters as ONE term with coefficient 1 (seal_ring.tcc:525-528)
      CUDA_TRY(cudaMemcpyAsync(crs->d + (carry_first + k) * E, d_carry + k * E, E * 8, cudaMemcpyDeviceToDevice, c->stream));
      terms.push_back({(uint32_t)(carry_first + k), nullptr, 0});
    }
    terms.insert(terms.end(), ip[k].begin(), ip[k].end());
    if (terms.empty()) continue;   // still the empty EncodingElem
    if ((rc = inner_product_terms(c, crs->d, terms, c->d_ip + k * E))) return rc;
    CUDA_TRY(cudaMemcpyAsync(d_carry + k * E, c->d_ip + k * E, E * 8, cudaMemcpyDeviceToDevice, c->stream));
    h_present[k] = 1;
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}

extern "C" int rsg_groth16_chain_finish(rsg_context *c, const rsg_crs *crs, const rsg_groth16_layout *L, const uint64_t *d_carry,
                                        const uint8_t *h_present, uint64_t *d_proof) {
  RSG_TRACE_CALL();
  if (!c || !crs || !L || !d_carry || !h_present || !d_proof) return fail(RSG_ERR_ARG, "null argument");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  const size_t E = c->enc_words(), NONE = (size_t)-1;
  for (int e = 0; e < 3; e++) {   // EncodingElem::operator+= chain of groth16.tcc:89-112
    const int a = 2 * e, b = a + 1;
    const size_t extra = e == 0 ? L->alpha_idx : (e == 1 ? L->beta_idx : NONE);
    const uint64_t *src[3] = {h_present[a] ? d_carry + a * E : nullptr, h_present[b] ? d_carry + b * E : nullptr,
                              extra != NONE ? crs->d + extra * E : nullptr};
    uint64_t *acc = d_proof + e * E;
    bool have = false;
    for (int k = 0; k < 3; k++) {
      if (!src[k]) continue;
      if (!have) {
        CUDA_TRY(cudaMemcpyAsync(acc, src[k], E * 8, cudaMemcpyDeviceToDevice, c->stream));
        have = true;
      } else {
        const int r = enc_add_fix(c, acc, src[k]);
        if (r) return r;
      }
    }
    if (!have) CUDA_TRY(cudaMemsetAsync(acc, 0, E * 8, c->stream));
  }
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return RSG_OK;
}

// ---- the two exchange steps over NVLink peer memory (p2p.cuh) -----------------------------------------------------------
extern "C" int rsg_exchange_p2p(rsg_context *c, const uint64_t *d_wit, size_t n, size_t world, size_t rank, size_t per,
                                uint64_t *const *h_peer_full) {
  RSG_TRACE_CALL();
  if (!c || !d_wit || !h_peer_full || !world || world > (size_t)P2P_MAX || rank >= world || !per) return fail(RSG_ERR_ARG, "bad argument");
  if (c->N_R & 1) return fail(RSG_ERR_ARG, "the slot block of a rank must hold an even number of slots");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  PeerPtrs pp;
  for (size_t d = 0; d < world; d++) pp.p[d] = h_peer_full[d];
  LaunchScope ls(c, "k_exchange_p2p");
  // c is the WITNESS context of the rank: its N_R is the slot block S
  k_exchange_p2p<<<dim3((unsigned)(5 * per), (unsigned)world, (unsigned)c->L_R), 128, 0, c->stream>>>(d_wit, pp, (uint32_t)n, (uint32_t)per, (uint32_t)c->N_R,
                                                                                                  (uint32_t)c->L_R, (uint32_t)world, (uint32_t)rank);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}
extern "C" int rsg_enc_sum_p2p(rsg_context *c, uint64_t *const *h_peer_parts, uint64_t *const *h_peer_final, size_t world, size_t rank,
                               size_t n_enc, size_t block_words, uint64_t *d_blocks) {
  RSG_TRACE_CALL();
  if (!c || !h_peer_parts || !h_peer_final || !world || world > (size_t)P2P_MAX || rank >= world || !n_enc) return fail(RSG_ERR_ARG, "bad argument");
  if (block_words && !d_blocks) return fail(RSG_ERR_ARG, "null block buffer");
  const size_t words3 = n_enc * c->enc_words();
  if (words3 % (2 * world)) return fail(RSG_ERR_ARG, "the proof does not split into `world` even slices");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  PeerPtrs pa, pf;
  for (size_t d = 0; d < world; d++) { pa.p[d] = h_peer_parts[d]; pf.p[d] = h_peer_final[d]; }
  LaunchScope ls(c, "k_enc_sum");
  const size_t pairs = words3 / world / 2;
  k_enc_sum_p2p<<<(unsigned)((pairs + 255) / 256), 256, 0, c->stream>>>(c->d_params, pa, pf, (uint32_t)world, (uint32_t)rank, words3,
                                                                       (uint32_t)block_words, d_blocks);
  CUDA_TRY(cudaGetLastError());
  return RSG_OK;
}

extern "C" int rsg_groth16_prove(rsg_context *c, const rsg_r1cs *r1cs, const rsg_crs *crs, const rsg_groth16_layout *L,
                                 rsg_ringvec *assignment, const uint64_t *h_assignment, const uint8_t *h_aux_kind, uint64_t *h_proof,
                                 uint64_t *d_proof, size_t *n_used) {
  RSG_TRACE_CALL();
  if (!c || !r1cs || !crs || !L || !assignment) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r1cs->n, n_io = r1cs->n_io, n_aux = r1cs->n_aux, W = c->ring_words();
  if (assignment->n < n_io + n_aux) return fail(RSG_ERR_ARG, "assignment too short");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc;
  if (h_assignment)
    CUDA_TRY(cudaMemcpyAsync(assignment->d, h_assignment, (n_io + n_aux) * W * 8, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ensure(c, &c->d_evals, &c->cap_evals, 9 * n * W))) return rc;
  if ((rc = ensure(c, &c->d_wit, &c->cap_wit, (7 * n + 1) * W))) return rc;
  uint64_t *coeffs = c->d_wit, *H = c->d_wit + 6 * n * W;
  if ((rc = r1cs_eval_dev(c, r1cs, assignment->d, c->d_evals))) return rc;
  if ((rc = witness_map_dev(c, n, c->d_evals, coeffs, H, nullptr, const_cast<rsg_r1cs *>(r1cs), /*need_C=*/false))) return rc;
  // coeffs order in HBM: A_io, B_io, C_io, A_mid, B_mid, C_mid
  const uint64_t *vec[6] = {coeffs, coeffs + 3 * n * W, coeffs + n * W, coeffs + 4 * n * W, H, assignment->d + n_io * W};
  return groth16_lincombs_dev(c, crs, L, n, n_aux, vec, h_aux_kind, h_proof, d_proof, n_used);
}
// ------------------------------------------------------------------------------------------------------------
// Provers over CRS vectors that live in DIFFERENT arenas (one per EncodingElem::encode call of the generator templates).
static int ref_ptr(const rsg_context *c, const rsg_crs_ref &r, size_t count, const uint64_t **out) {
  if (!r.crs) { *out = nullptr; return count ? fail(RSG_ERR_ARG, "null CRS reference") : RSG_OK; }
  if (r.crs->ctx != c || r.first + count > r.crs->n) return fail(RSG_ERR_ARG, "CRS reference out of range or of another context");
  *out = r.crs->d + r.first * c->enc_words();
  return RSG_OK;
}

extern "C" int rsg_groth16_prove_refs(rsg_context *c, const rsg_r1cs *r1cs, const rsg_crs_ref refs[5], rsg_ringvec *assignment,
                                      const uint64_t *h_assignment, const uint8_t *h_aux_kind, uint64_t *h_proof, uint64_t *d_proof,
                                      size_t *n_used) {
  RSG_TRACE_CALL();
  if (!c || !r1cs || !refs || !assignment) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r1cs->n, n_io = r1cs->n_io, n_aux = r1cs->n_aux, W = c->ring_words();
  if (assignment->n < n_io + n_aux) return fail(RSG_ERR_ARG, "assignment too short");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc;
  G16Ptrs G;
  if ((rc = ref_ptr(c, refs[0], n + 1, &G.s_pows)) || (rc = ref_ptr(c, refs[1], n + 1, &G.delta_ts)) ||
      (rc = ref_ptr(c, refs[2], n_aux, &G.delta_mid)) || (rc = ref_ptr(c, refs[3], 1, &G.alpha)) || (rc = ref_ptr(c, refs[4], 1, &G.beta)))
    return rc;
  rsg_groth16_layout L;
  memset(&L, 0, sizeof(L));
  L.s_pows_hi = n + 1; L.delta_ts_hi = n + 1; L.delta_mid_hi = n_aux;   // whole vectors; the offsets are carried by G
  if (!fast_applies(c, &L, n, n_aux)) return fail(RSG_ERR_UNSUPPORTED, "term set above the NTT-plaintext budget: use the per-inner-product entry points");
  if (h_assignment)
    CUDA_TRY(cudaMemcpyAsync(assignment->d, h_assignment, (n_io + n_aux) * W * 8, cudaMemcpyHostToDevice, c->stream));
  if ((rc = ensure(c, &c->d_evals, &c->cap_evals, 9 * n * W))) return rc;
  if ((rc = ensure(c, &c->d_wit, &c->cap_wit, (7 * n + 1) * W))) return rc;
  uint64_t *coeffs = c->d_wit, *H = c->d_wit + 6 * n * W;
  if ((rc = r1cs_eval_dev(c, r1cs, assignment->d, c->d_evals))) return rc;
  if ((rc = witness_map_dev(c, n, c->d_evals, coeffs, H, nullptr, const_cast<rsg_r1cs *>(r1cs), /*need_C=*/false))) return rc;
  const uint64_t *vec[6] = {coeffs, coeffs + 3 * n * W, coeffs + n * W, coeffs + 4 * n * W, H, assignment->d + n_io * W};
  uint64_t *out = d_proof;
  if (!out) {
    if ((rc = ensure(c, &c->d_out_scratch, &c->cap_out_scratch, 3 * c->enc_words()))) return rc;
    out = c->d_out_scratch;
  }
  int candidate = 0;
  if ((rc = groth16_lincombs_fast(c, G, &L, n, n_aux, vec, h_aux_kind, h_proof, out, n_used, &candidate))) return rc;
  if (candidate) {
    c->st_fast_fallback++;
    return fail(RSG_ERR_TRANSPARENT, "transparent-ciphertext candidate: resolve with the per-inner-product entry points");
  }
  return RSG_OK;
}

// rinocchio::prover (zk_proof_systems/rinocchio/rinocchio.tcc:74-190) in one call: zero-knowledge witness map, the eleven inner
// products with ONE batch encoding and ONE set of forward NTTs per coefficient (the reference, and the template path, transform
// a_mid / b_mid / c_mid / h / z twice: once for s_pows, once for alpha_s_pows), and the nine shifts.
extern "C" int rsg_rinocchio_prove(rsg_context *c, const rsg_r1cs *r1cs, const rsg_crs_ref refs[6], rsg_ringvec *assignment,
                                   const uint64_t *h_assignment, const uint8_t *h_aux_kind, const uint64_t *h_d, uint64_t *h_proof,
                                   uint64_t *d_proof, size_t *n_used) {
  RSG_TRACE_CALL();
  if (!c || !r1cs || !refs || !assignment) return fail(RSG_ERR_ARG, "null argument");
  const size_t n = r1cs->n, n_io = r1cs->n_io, n_aux = r1cs->n_aux, W = c->ring_words(), E = c->enc_words();
  if (assignment->n < n_io + n_aux) return fail(RSG_ERR_ARG, "assignment too short");
  if (!c->fast_mode) return fail(RSG_ERR_UNSUPPORTED, "RSG_FAST=0: use the per-inner-product entry points");
  std::lock_guard<std::mutex> g(c->mu);
  CUDA_TRY(cudaSetDevice(c->device));
  int rc;
  const bool zk = h_d != nullptr;
  const uint64_t *p_s, *p_as, *p_bp, *p_bt[3] = {nullptr, nullptr, nullptr};
  if ((rc = ref_ptr(c, refs[0], n + 1, &p_s)) || (rc = ref_ptr(c, refs[1], n + 1, &p_as)) || (rc = ref_ptr(c, refs[2], n_aux, &p_bp))) return rc;
  if (zk && n_aux)
    for (int k = 0; k < 3; k++)
      if ((rc = ref_ptr(c, refs[3 + k], 1, &p_bt[k]))) return rc;
  const size_t slots = 3 * n + 2 * (n + 1) + n_aux + (zk ? 3 : 0);
  if (slots * c->L_R * c->L_E * c->N_E > c->pntt_budget_words)
    return fail(RSG_ERR_UNSUPPORTED, "term set above the NTT-plaintext budget: use the per-inner-product entry points");
  if (h_assignment)
    CUDA_TRY(cudaMemcpyAsync(assignment->d, h_assignment, (n_io + n_aux) * W * 8, cudaMemcpyHostToDevice, c->stream));
  const uint64_t *d_zk = nullptr;
  if (zk) {
    if ((rc = ensure(c, &c->d_zk, &c->cap_zk, 3 * W))) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->d_zk, h_d, 3 * W * 8, cudaMemcpyHostToDevice, c->stream));
    d_zk = c->d_zk;
  }
  if ((rc = ensure(c, &c->d_evals, &c->cap_evals, 9 * n * W))) return rc;
  if ((rc = ensure(c, &c->d_wit, &c->cap_wit, (7 * n + 1) * W))) return rc;
  uint64_t *coeffs = c->d_wit, *H = c->d_wit + 6 * n * W;
  if ((rc = r1cs_eval_dev(c, r1cs, assignment->d, c->d_evals))) return rc;
  if ((rc = witness_map_dev(c, n, c->d_evals, coeffs, H, d_zk, const_cast<rsg_r1cs *>(r1cs)))) return rc;
  // coefficients_for_Z as ring elements: every coefficient below the leading one has been through a negation in the reference
  // (evaluation_domain.tcc:53-60) and is a polynomial with the constant in every slot; the leading one is the scalar 1
  WitnessTables *wt;
  if ((rc = get_witness_tables(c, n, &wt))) return rc;
  if (!wt->d_Zvec) {
    std::vector<uint64_t> zv((n + 1) * W);
    for (size_t k = 0; k <= n; k++)
      for (size_t j = 0; j < c->L_R; j++) std::fill(zv.begin() + k * W + j * c->N_R, zv.begin() + k * W + (j + 1) * c->N_R, wt->h_Z[j * (n + 1) + k]);
    void *v = nullptr;
    CUDA_TRY(cudaMalloc(&v, zv.size() * 8));
    wt->d_Zvec = (uint64_t *)v;
    CUDA_TRY(cudaMemcpy(wt->d_Zvec, zv.data(), zv.size() * 8, cudaMemcpyHostToDevice));
  }
  // plan: vectors A_mid, B_mid, C_mid, H, Z, aux, D; inner products in the proof's order + z, alpha_z, f
  std::vector<uint8_t> key, zkind(n + 1, (uint8_t)RSG_AUX_POLY), all;
  zkind[n] = RSG_TERM_ONE;
  const uint32_t tagk = 0x72696E6Fu;
  key_put(key, &tagk, 4); key_put(key, &n, sizeof(n)); key_put(key, &n_aux, sizeof(n_aux)); key_put(key, &zk, sizeof(zk));
  key_put(key, &p_s, 8); key_put(key, &p_as, 8); key_put(key, &p_bp, 8); key_put(key, p_bt, sizeof(p_bt));
  if (h_aux_kind) key_put(key, h_aux_kind, n_aux);
  FastPlan *fp;
  {
    FastSpec sp;
    sp.n_vec = zk ? 7 : 6; sp.nS = 0;
    const uint32_t cnt[7] = {(uint32_t)n, (uint32_t)n, (uint32_t)n, (uint32_t)(n + 1), (uint32_t)(n + 1), (uint32_t)n_aux, 3};
    for (uint32_t k = 0; k < sp.n_vec; k++) { sp.lo[k] = 0; sp.count[k] = cnt[k]; }
    sp.kind[4] = zkind.data(); sp.kind_len[4] = n + 1;
    if (h_aux_kind && n_aux) { sp.kind[5] = h_aux_kind; sp.kind_len[5] = n_aux; }
    sp.n_ip = n_aux ? 11 : 10;
    const uint32_t ipv[11] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5};
    for (uint32_t p = 0; p < sp.n_ip; p++) { sp.ip_vec[p] = ipv[p]; sp.ip_base[p] = p == 10 ? p_bp : ((p & 1) ? p_as : p_s); }
    sp.n_grp = sp.n_out = sp.n_ip;
    for (uint32_t p = 0; p < sp.n_ip; p++) { sp.grp_ip[p][0] = p; sp.grp_ip[p][1] = 0xFFFFFFFFu; sp.grp_out[p] = p; }
    // the D vector has no inner product of its own: its slots are only transformed (the shift plaintexts)
    if ((rc = fast_build_plan(c, sp, key, &fp))) return rc;
  }
  FastTable T = fp->T;
  const uint64_t *bases[7] = {coeffs + 3 * n * W, coeffs + 4 * n * W, coeffs + 5 * n * W, H, wt->d_Zvec, assignment->d + n_io * W, c->d_zk};
  for (uint32_t k = 0; k < T.n_vec; k++) T.vec[k].base = bases[k];
  if ((rc = ensure(c, &c->d_ip, &c->cap_ip, 11 * E))) return rc;
  CUDA_TRY(cudaMemsetAsync(c->d_ip + 10 * E, 0, E * 8, c->stream));   // f stays the empty encoding without auxiliary inputs
  if ((rc = fast_run(c, fp, T, c->d_ip))) return rc;
  uint64_t *out = d_proof;
  if (!out) {
    if ((rc = ensure(c, &c->d_out_scratch, &c->cap_out_scratch, 9 * E))) return rc;
    out = c->d_out_scratch;
  }
  cudaStream_t st = c->stream;
  RinoShift R;
  for (int k = 0; k < 3; k++) R.beta_ts[k] = p_bt[k];
  const uint32_t d_slot0 = (uint32_t)(T.vec[6].eid0);   // nS = 0: slot = element id
  const uint32_t shifts = zk ? (n_aux ? 2u : 1u) : 0u;     // 2: also the three beta_r?_ts shifts of f
  if (zk) {
    LaunchScope ls(c, "k_probe");
    k_probe_chain_rino<<<1, 32, 0, st>>>(c->d_params, R, c->d_pntt, d_slot0, c->d_fp_totals, c->d_fp_status, shifts);
  }
  {
    LaunchScope ls(c, "k_rino_combine");
    k_rino_combine<<<dim3((unsigned)((E / 2 + 255) / 256), 9), 256, 0, st>>>(c->d_params, c->d_ip, R, c->d_pntt, d_slot0, shifts, out);
  }
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(c->h_fp_status, c->d_fp_status, FPS_WORDS * 4, cudaMemcpyDeviceToHost, st));
  if (h_proof) CUDA_TRY(cudaMemcpyAsync(h_proof, out, 9 * E * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  const uint32_t *hs = c->h_fp_status;
  if (n_used) {
    const int v[9] = {0, 0, 1, 1, 2, 2, 3, 3, 5};
    for (int e = 0; e < 9; e++) n_used[e] = hs[FPS_COUNT0 + v[e]] + ((zk && e != 6 && e != 7 && (e < 8 || n_aux)) ? 1 : 0);
    if (!n_aux) n_used[8] = 0;
  }
  c->st_fast++;
  if (hs[FPS_CANDIDATE]) {
    c->st_fast_fallback++;
    return fail(RSG_ERR_TRANSPARENT, "transparent-ciphertext candidate: resolve with the per-inner-product entry points");
  }
  return RSG_OK;
}
#include "serialize.inl"
