// Dense tensor contraction on FP64 / complex-FP64 DMMA tiles with the index permutation
// fused into the tile loads (GETT-style: no permuted copy of any operand is ever made).
//
// Replaces `_contract!` -> CUTENSOR.contraction! (/root/reference/src/tensor/cudense.jl:238-331),
// `_gemm_contract!` (cudense.jl:170-236) and the scalar/outer branches of `contract!!`
// (cudense.jl:83-110).
//
// Design (sm_100a): tcgen05/UMMA has no f64 kind, so FP64 tensor math is warp-level
// `mma.sync.m8n8k4.f64` (SASS DMMA.8x8x4).  One CTA owns a BM x BN tile of the
// "matricised" C; modes are grouped into M (A&C), N (B&C), K (A&B), each group linearised
// in mixed radix, so element (m,k) of A lives at offM_A(m) + offK_A(k).  Tiles are staged
// global->shared with a 4-deep cp.async (LDGSTS) ring, each operand along whichever of its
// two directions is contiguous in HBM (16-byte copies when parity allows), into a padded
// layout that makes the 64-bit fragment loads bank-conflict free.
#include "tnb_internal.h"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

namespace tnb {

// kernel families live in contract_f64.cu / contract_c128.cu (device code: contract_kernel.cuh)
int launch_tiles_f64(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, bool small, cudaStream_t st);
int launch_tiles_c128(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, bool small, cudaStream_t st);
int launch_smallk_f64(Handle* h, GemmParams& p, cudaStream_t st);
int launch_smallk_c128(Handle* h, GemmParams& p, cudaStream_t st);
// TMA-staged kernel for operand pairs that are both K-major (contract_tma.cu)
bool tma_eligible(const GemmParams& p, int dtype, bool small);
bool tma_shape_ok(const GemmParams& p, int dtype, bool small);
int launch_tma(Handle* h, int dtype, GemmParams& p, bool small, cudaStream_t st);
bool tma_would_split(Handle* h, const GemmParams& p, int dtype, bool small);
uint64_t tma_split_launches();

// ------------------------------------------------------------------------------------
// Plan cache (the analogue of the reference's `ContractionPlans` dictionary + cuTENSOR autotune,
// /root/reference/src/ITensorsGPU.jl:54-55 and src/tensor/cudense.jl:285-326, which keys a string of (mode, extent)
// pairs).  Key: dtype, flags, every (mode label, extent, stride) of the three operands and the 16-byte alignment of
// the base pointers.  Value: the grouped / merged GEMM parameters and the kernel variant (staging directions, copy
// widths, tile configuration).  A hit skips the whole planning step (mode matching, grouping, merging -- what
// dominates the host time of the small contractions of C1); a miss plans once and, when autotuning is on and the
// call is idempotent (beta = 0), times the candidate tile configurations on the caller's own operands and keeps the
// fastest.  All candidates accumulate every output element in the same k order, so the choice never changes a bit of
// the result (ranks of a sharded sweep stay bit-identical whatever each one measured).
// ------------------------------------------------------------------------------------
struct Variant {
  bool smallk = false;
  bool ak = true, bk = true;
  int va = 1, vb = 1;
  bool small = false;
  bool tma = false;      // both operands K-major and expressible as tensor maps: contract_tma.cu
};
struct Plan {
  GemmParams p;
  Variant v;
  bool tuned = false;
};
struct PlanCache {
  std::unordered_map<std::string, Plan> map;
  uint64_t hits = 0, misses = 0, tuned = 0;
  int autotune = 1;                 // 0: heuristic only, 1: time the tile configurations of new large shapes
};
static std::unordered_map<Handle*, PlanCache>& caches() {
  static std::unordered_map<Handle*, PlanCache> c;
  return c;
}
static std::mutex& cache_mutex() { static std::mutex m; return m; }

void plan_cache_drop(Handle* h) {
  std::lock_guard<std::mutex> g(cache_mutex());
  caches().erase(h);
}
void plan_cache_stats(Handle* h, uint64_t* entries, uint64_t* hits, uint64_t* misses, uint64_t* tuned) {
  std::lock_guard<std::mutex> g(cache_mutex());
  PlanCache& c = caches()[h];
  if (entries) *entries = c.map.size();
  if (hits) *hits = c.hits;
  if (misses) *misses = c.misses;
  if (tuned) *tuned = c.tuned;
}
void plan_cache_set_autotune(Handle* h, int mode) {
  std::lock_guard<std::mutex> g(cache_mutex());
  caches()[h].autotune = mode;
}
void plan_cache_clear(Handle* h) {
  std::lock_guard<std::mutex> g(cache_mutex());
  PlanCache& c = caches()[h];
  c.map.clear();
  c.hits = c.misses = c.tuned = 0;
}

static bool all_even(const long long* s, int from, int n) {
  for (int i = from; i < n; ++i)
    if (s[i] & 1) return false;
  return true;
}

static void set_scalars(GemmParams& p, int dtype, const void* alpha, const void* beta);

// Decide per-operand staging direction and copy width and pick a tile config (heuristic).
static Variant choose_variant(Handle* h, int dtype, const GemmParams& p) {
  Variant v;
  const bool cplx = dtype == TNB_C128;
  {  // small-K / small-N streaming form
    static const bool off = getenv("TNB_SMALLK") && !strcmp(getenv("TNB_SMALLK"), "off");
    if (!off && p.K <= 32 && p.N <= 32 && p.M >= 16384 && p.batch <= 1 && !p.boffA && !p.boffB && !p.boffC && !p.splitN &&
        p.npeer == 0 && !p.lowerOnly) {
      v.smallk = true;
      return v;
    }
  }
  // A: contiguous along K if the first K mode has unit stride in A; along M if the first M mode has.
  int va = 1, vb = 1;
  const bool a_k1 = p.gk.sX[0] == 1 && p.K > 1, a_m1 = p.gm.sX[0] == 1 && p.M > 1;
  const bool b_k1 = p.gk.sY[0] == 1 && p.K > 1, b_n1 = p.gn.sX[0] == 1 && p.N > 1;
  v.ak = a_k1 || !a_m1;
  v.bk = b_k1 || !b_n1;
  if (!cplx) {
    if (a_k1) {
      if ((p.gk.ext[0] % 2 == 0) && all_even(p.gk.sX, 1, p.gk.n) && all_even(p.gm.sX, 0, p.gm.n) &&
          ((uintptr_t)p.A % 16 == 0)) va = 2;
    } else if (a_m1) {
      if ((p.gm.ext[0] % 2 == 0) && all_even(p.gm.sX, 1, p.gm.n) && all_even(p.gk.sX, 0, p.gk.n) &&
          ((uintptr_t)p.A % 16 == 0)) va = 2;
    }
    if (b_k1) {
      if ((p.gk.ext[0] % 2 == 0) && all_even(p.gk.sY, 1, p.gk.n) && all_even(p.gn.sX, 0, p.gn.n) &&
          ((uintptr_t)p.B % 16 == 0)) vb = 2;
    } else if (b_n1) {
      if ((p.gn.ext[0] % 2 == 0) && all_even(p.gn.sX, 1, p.gn.n) && all_even(p.gk.sY, 0, p.gk.n) &&
          ((uintptr_t)p.B % 16 == 0)) vb = 2;
    }
  }
  if (!p.boffA && (p.bstrideA & 1)) va = 1;   // batched: 16-byte copies need even element offsets
  if (!p.boffB && (p.bstrideB & 1)) vb = 1;   // (offset tables must hold even offsets for f64)
  v.va = va; v.vb = vb;
  // tile config: big tiles unless they cannot fill the machine (tile shapes: contract_kernel.cuh -- real 64x128
  // / 64x64, complex 64x64 / 64x32)
  auto ntiles = [&](int bm, int bn) { return ((long long)(p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn) * std::max(p.batch, 1); };
  if (!cplx) v.small = ntiles(64, 128) < h->num_sms || p.M <= 64 || p.N <= 64;
  else v.small = ntiles(64, 64) < h->num_sms || p.M <= 64 || p.N <= 32;
  v.tma = a_k1 && b_k1 && tma_eligible(p, dtype, v.small);
  return v;
}

// calls per kernel family since load: [0] LDGSTS tile kernel, [1] small-K streaming kernel, [2] TMA kernel,
// [3] TMA calls that split K over clusters
static uint64_t g_family[3] = {0, 0, 0};
void kernel_family_counts(uint64_t out[4]) { for (int i = 0; i < 3; ++i) out[i] = g_family[i]; out[3] = tma_split_launches(); }

static int launch_variant(Handle* h, int dtype, GemmParams& p, const Variant& v, cudaStream_t st) {
  const bool cplx = dtype == TNB_C128;
  g_family[v.smallk ? 1 : (v.tma ? 2 : 0)]++;
  if (v.smallk) return cplx ? launch_smallk_c128(h, p, st) : launch_smallk_f64(h, p, st);
  if (v.tma) return launch_tma(h, dtype, p, v.small, st);
  if (!cplx) return launch_tiles_f64(h, p, v.ak, v.bk, v.va, v.vb, v.small, st);
  return launch_tiles_c128(h, p, v.ak, v.bk, v.va, v.vb, v.small, st);
}

static int launch_planned(Handle* h, int dtype, GemmParams& p, cudaStream_t st) {
  const Variant v = choose_variant(h, dtype, p);
  return launch_variant(h, dtype, p, v, st);
}

// One-shot autotune of a new shape: time the candidate variants on the caller's operands (beta = 0 only: the call is
// then idempotent) and keep the fastest.  Candidates never differ in the bits they produce:
//   * TMA-eligible shape that does not split K over clusters: TMA-staged kernel vs LDGSTS kernel, same tile (the two families
//     share the k order; measured at chi = 4096: step 1 is 2.5% faster through TMA, step 4 2.8% faster through LDGSTS);
//   * TMA-eligible shape that DOES split K (sub-wave problems): never tuned (the split changes the summation order and is a function of
//     the shape alone);
//   * other shapes: large vs small tile of the LDGSTS kernel.
// Returns with C holding the result.
static int autotune_variant(Handle* h, int dtype, GemmParams& p, Variant& v, bool* tuned, cudaStream_t st) {
  *tuned = false;
  const double flop = 2.0 * p.M * (double)p.N * p.K * std::max(p.batch, 1) * (dtype == TNB_C128 ? 4 : 1);
  const bool idempotent = p.beta_re == 0.0 && p.beta_im == 0.0;
  if (v.smallk || !idempotent || flop < 2e9 || p.M <= 64 || p.N <= 64) return launch_variant(h, dtype, p, v, st);
  Variant cands[2] = {v, v};
  if (v.tma) {
    if (tma_would_split(h, p, dtype, v.small)) return launch_variant(h, dtype, p, v, st);
    cands[1].tma = false;
  } else {
    // tile shapes differ only when both are sensible: enough tiles for the small one to matter, and not so many that
    // quantisation is irrelevant (> 8 waves of the big tile: keep the big tile)
    const long long big_tiles = ((long long)(p.M + 63) / 64) * ((p.N + (dtype == TNB_C128 ? 63 : 127)) / (dtype == TNB_C128 ? 64 : 128)) * std::max(p.batch, 1);
    if (big_tiles > 8LL * 2 * h->num_sms || p.npeer > 0) return launch_variant(h, dtype, p, v, st);
    cands[0].small = false;
    cands[1].small = true;
  }
  cudaEvent_t e[2];
  for (auto& x : e) TNB_CUDA(h, cudaEventCreate(&x));
  float best = 1e30f;
  int rc = TNB_OK;
  Variant bestv = v;
  for (int cand = 0; cand < 2 && !rc; ++cand) {
    const Variant& c = cands[cand];
    rc = launch_variant(h, dtype, p, c, st);                 // warm (instruction cache, L2)
    if (rc) break;
    cudaEventRecord(e[0], st);
    rc = launch_variant(h, dtype, p, c, st);
    cudaEventRecord(e[1], st);
    if (rc) break;
    cudaEventSynchronize(e[1]);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e[0], e[1]);
    if (ms < best) { best = ms; bestv = c; }
  }
  for (auto& x : e) cudaEventDestroy(x);
  if (rc) return rc;
  v = bestv;
  *tuned = true;
  return TNB_OK;      // the last candidate run left a complete, correct C (every candidate computes the same bits)
}

// ------------------------------------------------------------------------------------
// host side: planning from mode labels
// ------------------------------------------------------------------------------------
struct ModeRec {
  int label;
  long long ext, sA, sB, sC;
  int posA, posB, posC;
};

static void finish_group(Group& g, std::vector<ModeRec>& ms, int which /*0=M,1=N,2=K*/) {
  // drop extent-1 modes, merge modes adjacent in both carrying tensors
  std::vector<std::array<long long, 3>> v;  // ext, sX, sY
  for (auto& m : ms) {
    if (m.ext == 1) continue;
    long long sx = which == 1 ? m.sB : m.sA;
    long long sy = which == 2 ? m.sB : m.sC;
    if (!v.empty() && v.back()[1] * v.back()[0] == sx && v.back()[2] * v.back()[0] == sy &&
        v.back()[0] * m.ext < 2147483647LL) {
      v.back()[0] *= m.ext;
    } else {
      v.push_back({m.ext, sx, sy});
    }
  }
  if (v.empty()) v.push_back({1, 0, 0});
  g.n = (int)v.size();
  for (int i = 0; i < g.n; ++i) {
    g.ext[i] = (int)v[i][0];
    g.sX[i] = v[i][1];
    g.sY[i] = v[i][2];
  }
}

static void set_scalars(GemmParams& p, int dtype, const void* alpha, const void* beta) {
  p.alpha_re = 1.0; p.alpha_im = 0.0; p.beta_re = 0.0; p.beta_im = 0.0;
  if (alpha) { p.alpha_re = ((const double*)alpha)[0]; if (dtype == TNB_C128) p.alpha_im = ((const double*)alpha)[1]; }
  if (beta) { p.beta_re = ((const double*)beta)[0]; if (dtype == TNB_C128) p.beta_im = ((const double*)beta)[1]; }
}

// Mode labels -> grouped GEMM description (everything of GemmParams except the operand pointers, the scalars and the
// peer list).  Pure host logic: no CUDA call, `h` only receives the error text (tnb_plan_describe runs it without a GPU).
static int plan_from_modes(Handle* h, int dtype, int nA, const int64_t* extA, const int32_t* modeA, int nB,
                           const int64_t* extB, const int32_t* modeB, int nC, const int64_t* extC, const int32_t* modeC,
                           int flags, const int64_t* strideA, const int64_t* strideB, const int64_t* strideC,
                           GemmParams& p) {
  std::vector<ModeRec> recs;
  auto find = [&](int label) -> ModeRec* {
    for (auto& r : recs) if (r.label == label) return &r;
    return nullptr;
  };
  long long s = 1;
  for (int i = 0; i < nA; ++i) {
    if (find(modeA[i])) return set_err(h, TNB_ERR_BAD_ARG, "contract: repeated mode %d in A", modeA[i]);
    if (extA[i] < 1) return set_err(h, TNB_ERR_BAD_ARG, "contract: extent < 1 in A");
    recs.push_back({modeA[i], extA[i], strideA ? strideA[i] : s, 0, 0, i, -1, -1});
    s *= extA[i];
  }
  s = 1;
  for (int i = 0; i < nB; ++i) {
    for (int j = 0; j < i; ++j) if (modeB[j] == modeB[i]) return set_err(h, TNB_ERR_BAD_ARG, "contract: repeated mode %d in B", modeB[i]);
    if (extB[i] < 1) return set_err(h, TNB_ERR_BAD_ARG, "contract: extent < 1 in B");
    ModeRec* r = find(modeB[i]);
    if (r) {
      if (r->ext != extB[i]) return set_err(h, TNB_ERR_DIM_MISMATCH, "contract: mode %d has extent %lld in A, %lld in B", modeB[i], r->ext, (long long)extB[i]);
      r->sB = strideB ? strideB[i] : s; r->posB = i;
    } else {
      recs.push_back({modeB[i], extB[i], 0, strideB ? strideB[i] : s, 0, -1, i, -1});
    }
    s *= extB[i];
  }
  s = 1;
  for (int i = 0; i < nC; ++i) {
    for (int j = 0; j < i; ++j) if (modeC[j] == modeC[i]) return set_err(h, TNB_ERR_BAD_ARG, "contract: repeated mode %d in C", modeC[i]);
    ModeRec* r = find(modeC[i]);
    if (!r) return set_err(h, TNB_ERR_BAD_ARG, "contract: output mode %d is in neither input", modeC[i]);
    if (r->ext != extC[i]) return set_err(h, TNB_ERR_DIM_MISMATCH, "contract: mode %d has extent %lld in C", modeC[i], (long long)extC[i]);
    r->sC = strideC ? strideC[i] : s; r->posC = i;
    s *= extC[i];
  }
  std::vector<ModeRec> gm, gn, gk;
  for (auto& r : recs) {
    const bool a = r.posA >= 0, b = r.posB >= 0, c = r.posC >= 0;
    if (a && b && c) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: batch mode %d (in A, B and C)", r.label);
    if (a && b) gk.push_back(r);
    else if (a && c) gm.push_back(r);
    else if (b && c) gn.push_back(r);
    else return set_err(h, TNB_ERR_BAD_ARG, "contract: mode %d appears in only one tensor", r.label);
  }
  // orders: M follows A (= C for the NDTensors output order), N follows B, K follows the
  // operand whose unit-stride mode is contracted (A first).
  std::sort(gm.begin(), gm.end(), [](const ModeRec& x, const ModeRec& y) { return x.posA < y.posA; });
  std::sort(gn.begin(), gn.end(), [](const ModeRec& x, const ModeRec& y) { return x.posB < y.posB; });
  bool k_by_b = false;
  if (!gk.empty()) {
    bool a_first_in_k = false, b_first_in_k = false;
    for (auto& r : gk) { if (r.sA == 1 && r.ext > 1) a_first_in_k = true; if (r.sB == 1 && r.ext > 1) b_first_in_k = true; }
    k_by_b = !a_first_in_k && b_first_in_k;
  }
  if (k_by_b) std::sort(gk.begin(), gk.end(), [](const ModeRec& x, const ModeRec& y) { return x.posB < y.posB; });
  else std::sort(gk.begin(), gk.end(), [](const ModeRec& x, const ModeRec& y) { return x.posA < y.posA; });

  memset(&p, 0, sizeof(p));
  finish_group(p.gm, gm, 0);
  finish_group(p.gn, gn, 1);
  finish_group(p.gk, gk, 2);
  if (p.gm.n > MAXG || p.gn.n > MAXG || p.gk.n > MAXG) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: more than %d unmergeable modes in a group", MAXG);
  auto total = [&](const Group& g, long long& out) { out = 1; for (int i = 0; i < g.n; ++i) out *= g.ext[i]; return out <= 2147483647LL; };
  long long M, N, K;
  if (!total(p.gm, M) || !total(p.gn, N) || !total(p.gk, K)) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: grouped extent exceeds 2^31-1");
  p.M = (int)M; p.N = (int)N; p.K = (int)K;
  p.conjA = (flags & TNB_CONJ_A) ? 1 : 0;
  p.conjB = (flags & TNB_CONJ_B) ? 1 : 0;
  if (flags & TNB_HERM_UPPER) {
    if (p.M != p.N) return set_err(h, TNB_ERR_BAD_ARG, "contract: TNB_HERM_UPPER needs a square (M = N) result");
    // The kernel's (m, n) follow the OPERANDS' mode orders.  "m <= n" is C's upper triangle only if both
    // linearisations are C's own (row offset ascending with m, column offset ascending with n, both dense):
    // otherwise the hint is ignored and the full matrix is computed.
    auto c_ascending = [](const Group& g, long long first) {
      long long expect = first;
      for (int i = 0; i < g.n; ++i) { if (g.sY[i] != expect) return false; expect *= g.ext[i]; }
      return true;
    };
    if (c_ascending(p.gm, 1) && c_ascending(p.gn, (long long)p.M)) p.lowerOnly = 2;
  }
  return TNB_OK;
}

int contract_impl(Handle* h, int dtype, int nA, const int64_t* extA, const int32_t* modeA,
                  const void* A, int nB, const int64_t* extB, const int32_t* modeB,
                  const void* B, int nC, const int64_t* extC, const int32_t* modeC, void* C,
                  const void* alpha, const void* beta, int flags, cudaStream_t st) {
  return contract_impl_ex(h, dtype, nA, extA, modeA, A, nB, extB, modeB, B, nC, extC, modeC, C, alpha, beta, flags, st,
                          nullptr, nullptr, 0);
}

// strideA / strideB / strideC (optional): element stride of every mode (the operand is then a strided window
// of a larger tensor);
// peerC/npeer (optional): the epilogue stores every output element to ALL npeer base pointers (peer-mapped
// buffers of the other GPUs included) instead of C -- the all-gather of a sharded result fused into the GEMM.
int contract_impl_ex(Handle* h, int dtype, int nA, const int64_t* extA, const int32_t* modeA,
                     const void* A, int nB, const int64_t* extB, const int32_t* modeB,
                     const void* B, int nC, const int64_t* extC, const int32_t* modeC, void* C,
                     const void* alpha, const void* beta, int flags, cudaStream_t st,
                     const int64_t* strideC, void* const* peerC, int npeer, const int64_t* strideA,
                     const int64_t* strideB) {
  if (npeer < 0 || npeer > TNB_MAX_PEERS) return set_err(h, TNB_ERR_BAD_ARG, "contract: npeer %d", npeer);
  if (npeer > 0 && beta) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: beta with peer stores");
  if (dtype != TNB_F64 && dtype != TNB_C128) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: dtype %d", dtype);
  if (nA < 0 || nB < 0 || nC < 0 || nA > 64 || nB > 64 || nC > 64)
    return set_err(h, TNB_ERR_BAD_ARG, "contract: bad rank");
  if (!A || !B || !C) return set_err(h, TNB_ERR_BAD_ARG, "contract: null tensor pointer");
  // ---- plan cache lookup
  std::string key;
  key.reserve(64 + 20 * (nA + nB + nC));
  auto put = [&](long long x) { key.append((const char*)&x, sizeof(x)); };
  put(dtype); put(flags & (TNB_CONJ_A | TNB_CONJ_B | TNB_HERM_UPPER)); put(nA); put(nB); put(nC); put(npeer);
  put(((uintptr_t)A % 16 == 0) | (((uintptr_t)B % 16 == 0) << 1));
  for (int i = 0; i < nA; ++i) { put(modeA[i]); put(extA[i]); put(strideA ? strideA[i] : -1); }
  for (int i = 0; i < nB; ++i) { put(modeB[i]); put(extB[i]); put(strideB ? strideB[i] : -1); }
  for (int i = 0; i < nC; ++i) { put(modeC[i]); put(extC[i]); put(strideC ? strideC[i] : -1); }
  PlanCache* cache;
  {
    std::lock_guard<std::mutex> g(cache_mutex());
    cache = &caches()[h];
  }
  {
    auto it = cache->map.find(key);
    if (it != cache->map.end()) {
      cache->hits++;
      GemmParams p = it->second.p;
      p.A = A; p.B = B; p.C = C;
      set_scalars(p, dtype, alpha, beta);
      for (int g = 0; g < npeer; ++g) p.peerC[g] = peerC[g];
      return launch_variant(h, dtype, p, it->second.v, st);
    }
  }
  GemmParams p;
  TNB_TRY(plan_from_modes(h, dtype, nA, extA, modeA, nB, extB, modeB, nC, extC, modeC, flags, strideA, strideB, strideC, p));
  p.A = A; p.B = B; p.C = C;
  set_scalars(p, dtype, alpha, beta);
  p.npeer = npeer;
  for (int g = 0; g < npeer; ++g) p.peerC[g] = peerC[g];
  cache->misses++;
  Plan plan;
  plan.v = choose_variant(h, dtype, p);
  int rc;
  if (cache->autotune) rc = autotune_variant(h, dtype, p, plan.v, &plan.tuned, st);
  else rc = launch_variant(h, dtype, p, plan.v, st);
  if (rc) return rc;
  if (plan.tuned) cache->tuned++;
  plan.p = p;
  plan.p.A = plan.p.B = nullptr; plan.p.C = nullptr;
  if (cache->map.size() > 65536) cache->map.clear();        // unbounded shape churn: start over
  cache->map.emplace(std::move(key), plan);
  return TNB_OK;
}

int gemm_impl(Handle* h, int dtype, char opA, char opB, int64_t m, int64_t n, int64_t k,
              const void* alpha, const void* A, int64_t lda, const void* B, int64_t ldb,
              const void* beta, void* C, int64_t ldc, cudaStream_t st, int lower_only) {
  if (m == 0 || n == 0) return TNB_OK;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.lowerOnly = lower_only;
  p.gm.n = p.gn.n = p.gk.n = 1;
  p.gm.ext[0] = (int)m; p.gn.ext[0] = (int)n; p.gk.ext[0] = (int)std::max<int64_t>(k, 1);
  // op: 'N' none, 'T' transpose, 'C' conjugate transpose, 'J' conjugate (no transpose)
  const bool na = (opA == 'N' || opA == 'J'), nb = (opB == 'N' || opB == 'J');
  p.gm.sX[0] = na ? 1 : lda;  p.gk.sX[0] = na ? lda : 1;
  p.gn.sX[0] = nb ? ldb : 1;  p.gk.sY[0] = nb ? 1 : ldb;
  p.gm.sY[0] = 1; p.gn.sY[0] = ldc;
  p.M = (int)m; p.N = (int)n; p.K = (int)std::max<int64_t>(k, 1);
  p.A = A; p.B = B; p.C = C;
  set_scalars(p, dtype, alpha, beta);
  if (k == 0) { p.alpha_re = 0; p.alpha_im = 0; p.gk.sX[0] = 0; p.gk.sY[0] = 0; }
  p.conjA = (opA == 'C' || opA == 'J'); p.conjB = (opB == 'C' || opB == 'J');
  return launch_planned(h, dtype, p, st);
}

int gemm_batched_impl(Handle* h, int dtype, char opA, char opB, int64_t m, int64_t n, int64_t k,
                      const void* alpha, const void* A, int64_t lda, const long long* offA, long long strideA,
                      const void* B, int64_t ldb, const long long* offB, long long strideB, const void* beta,
                      void* C, int64_t ldc, const long long* offC, long long strideC, int batch,
                      cudaStream_t st, int splitN, const long long* offC2) {
  if (m == 0 || n == 0 || batch == 0) return TNB_OK;
  if (splitN && (!offC || !offC2)) return set_err(h, TNB_ERR_BAD_ARG, "gemm_batched: split output needs both offset tables");
  if (k < 1) return set_err(h, TNB_ERR_BAD_ARG, "gemm_batched: k < 1");
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.gm.n = p.gn.n = p.gk.n = 1;
  p.gm.ext[0] = (int)m; p.gn.ext[0] = (int)n; p.gk.ext[0] = (int)k;
  p.gm.sX[0] = (opA == 'N') ? 1 : lda;  p.gk.sX[0] = (opA == 'N') ? lda : 1;
  p.gn.sX[0] = (opB == 'N') ? ldb : 1;  p.gk.sY[0] = (opB == 'N') ? 1 : ldb;
  p.gm.sY[0] = 1; p.gn.sY[0] = ldc;
  p.M = (int)m; p.N = (int)n; p.K = (int)k;
  p.A = A; p.B = B; p.C = C;
  set_scalars(p, dtype, alpha, beta);
  p.conjA = opA == 'C'; p.conjB = opB == 'C';
  p.batch = batch;
  p.boffA = offA; p.boffB = offB; p.boffC = offC;
  p.bstrideA = strideA; p.bstrideB = strideB; p.bstrideC = strideC;
  p.splitN = splitN; p.boffC2 = offC2;
  return launch_planned(h, dtype, p, st);
}

}  // namespace tnb

// ---- planner dry run (host only)
using namespace tnb;
extern "C" int tnb_plan_describe_strided(int dtype, int nA, const int64_t* extA, const int32_t* modeA, const int64_t* strideA,
                                         int nB, const int64_t* extB, const int32_t* modeB, const int64_t* strideB, int nC,
                                         const int64_t* extC, const int32_t* modeC, const int64_t* strideC, int flags,
                                         int num_sms, tnb_plan_desc* out, char* err, size_t errlen);

extern "C" int tnb_plan_describe(int dtype, int nA, const int64_t* extA, const int32_t* modeA, int nB, const int64_t* extB,
                                 const int32_t* modeB, int nC, const int64_t* extC, const int32_t* modeC, int flags,
                                 int num_sms, tnb_plan_desc* out, char* err, size_t errlen) {
  return tnb_plan_describe_strided(dtype, nA, extA, modeA, nullptr, nB, extB, modeB, nullptr, nC, extC, modeC, nullptr, flags,
                                   num_sms, out, err, errlen);
}

// strideX (optional, elements): the operand is a strided window of a larger tensor -- how the chunked / sharded DMRG
// entry points address slabs of L, R, A and the output vector (heff.cu, shardops.cu)
extern "C" int tnb_plan_describe_strided(int dtype, int nA, const int64_t* extA, const int32_t* modeA, const int64_t* strideA,
                                         int nB, const int64_t* extB, const int32_t* modeB, const int64_t* strideB, int nC,
                                         const int64_t* extC, const int32_t* modeC, const int64_t* strideC, int flags,
                                         int num_sms, tnb_plan_desc* out, char* err, size_t errlen) {
  Handle tmp;                      // never touches CUDA: carries num_sms in and the error text out
  tmp.num_sms = num_sms > 0 ? num_sms : 148;
  auto fail = [&](int rc) {
    if (err && errlen) snprintf(err, errlen, "%s", tmp.err.c_str());
    return rc;
  };
  if (!out) return fail(set_err(&tmp, TNB_ERR_BAD_ARG, "plan_describe: null output"));
  if (dtype != TNB_F64 && dtype != TNB_C128) return fail(set_err(&tmp, TNB_ERR_UNSUPPORTED, "contract: dtype %d", dtype));
  if (nA < 0 || nB < 0 || nC < 0 || nA > 64 || nB > 64 || nC > 64) return fail(set_err(&tmp, TNB_ERR_BAD_ARG, "contract: bad rank"));
  if ((nA && (!extA || !modeA)) || (nB && (!extB || !modeB)) || (nC && (!extC || !modeC)))
    return fail(set_err(&tmp, TNB_ERR_BAD_ARG, "plan_describe: null extent / mode array"));
  GemmParams p;
  const int rc = plan_from_modes(&tmp, dtype, nA, extA, modeA, nB, extB, modeB, nC, extC, modeC, flags, strideA, strideB, strideC, p);
  if (rc) return fail(rc);
  p.alpha_re = 1.0;                // base pointers stay null: "16-byte aligned operands"
  const Variant v = choose_variant(&tmp, dtype, p);
  memset(out, 0, sizeof(*out));
  out->M = p.M; out->N = p.N; out->K = p.K;
  out->n_m = p.gm.n; out->n_n = p.gn.n; out->n_k = p.gk.n;
  for (int i = 0; i < p.gm.n; ++i) { out->ext_m[i] = p.gm.ext[i]; out->a_stride_m[i] = p.gm.sX[i]; out->c_stride_m[i] = p.gm.sY[i]; }
  for (int i = 0; i < p.gn.n; ++i) { out->ext_n[i] = p.gn.ext[i]; out->b_stride_n[i] = p.gn.sX[i]; out->c_stride_n[i] = p.gn.sY[i]; }
  for (int i = 0; i < p.gk.n; ++i) { out->ext_k[i] = p.gk.ext[i]; out->a_stride_k[i] = p.gk.sX[i]; out->b_stride_k[i] = p.gk.sY[i]; }
  const bool cplx = dtype == TNB_C128;
  const bool a_k1 = p.gk.sX[0] == 1 && p.K > 1, b_k1 = p.gk.sY[0] == 1 && p.K > 1;
  const bool tma_shape = !v.smallk && a_k1 && b_k1 && tma_shape_ok(p, dtype, v.small);
  out->family = v.smallk ? 1 : (tma_shape ? 2 : 0);
  out->a_k_major = v.ak; out->b_k_major = v.bk;
  out->a_vec = v.va; out->b_vec = v.vb;
  out->herm_upper = p.lowerOnly == 2;
  if (v.smallk) {
    out->tile_m = 256; out->tile_n = p.N; out->tile_k = p.K;
    out->tiles = ((long long)p.M + 255) / 256;
  } else {
    out->tile_m = 64;
    out->tile_n = cplx ? (v.small ? 32 : 64) : (v.small ? 64 : 128);
    out->tile_k = cplx ? 8 : 16;           // both tile families: one 128-byte row of K per tile row
    out->tiles = (((long long)p.M + out->tile_m - 1) / out->tile_m) * (((long long)p.N + out->tile_n - 1) / out->tile_n);
  }
  out->waves = (double)out->tiles / (2.0 * tmp.num_sms);
  return TNB_OK;
}
