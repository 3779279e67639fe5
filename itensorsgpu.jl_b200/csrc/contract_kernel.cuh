// Device code of the contraction engine: tile loader, DMMA kernel, small-K streaming kernel and their launch
// templates.  Included by contract_f64.cu and contract_c128.cu, which instantiate the Float64 and ComplexF64
// kernel families in separate translation units (they compile in parallel; one unit took 6 minutes);
// contract.cu holds the planner and calls launch_tiles_f64 / launch_tiles_c128.
#pragma once
#include "tnb_internal.h"

#include <algorithm>
#include <array>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace tnb {


// ------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------
template <bool USE_Y>
__device__ __forceinline__ long long decode(const Group& g, int idx) {
  long long off = 0;
  const int n1 = g.n - 1;
#pragma unroll 1
  for (int i = 0; i < n1; ++i) {
    const int e = g.ext[i];
    const int q = idx / e;
    const int r = idx - q * e;
    off += (long long)r * (USE_Y ? g.sY[i] : g.sX[i]);
    idx = q;
  }
  off += (long long)idx * (USE_Y ? g.sY[n1] : g.sX[n1]);
  return off;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* g, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(g), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor op.  Lane l holds A[l/4][l%4],
// B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int BM_, int BN_, int BK_, int WM_, int WN_, int STAGES_, int MINB_>
struct Cfg {
  static constexpr int BM = BM_, BN = BN_, BK = BK_, WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8, NI = WN / 8;
};

// Shared-memory geometry of one operand tile (ROWS free-dim entries x BK), in elements.
// K-major  : [row][k], pitch PK  (contiguous direction in HBM is k)
// free-major: [k][row], pitch PR (contiguous direction in HBM is the free index)
// Pitches chosen so a half-warp's 64-bit (quarter-warp's 128-bit) fragment loads hit
// distinct banks: real  PK%16==4, PR%16==4 ; complex PK%8==4, PR%8==2.
template <bool CPLX, int ROWS, int BK>
struct TileGeom {
  static constexpr int PK = BK + 4;
  static constexpr int PR = ROWS + (CPLX ? 2 : 4);
  static constexpr int ELEMS_K = ROWS * PK;
  static constexpr int ELEMS_R = BK * PR;
  static constexpr int ELEMS = ELEMS_K > ELEMS_R ? ELEMS_K : ELEMS_R;
};

// cp.async staging of one operand tile.  Everything a copy needs is kept as running state so that the inner
// loop pays ONE 64-bit add + one LDGSTS per 16-byte copy (ncu, round 1: with offsets re-derived per copy the
// kernel executed as many IMADs as DMMAs and the resulting fixed-latency stalls held the DMMA pipe at 82-89%):
//   rowptr[]  global byte address of each tile row this thread copies (decoded once per CTA)
//   kb[]      byte offset of this thread's k (per k-row for the free-major form); advanced by BK * stride of the
//             leading K mode per tile, re-decoded only when the leading mode wraps (once per ext0/BK tiles)
//   dst[]     shared-memory byte offset inside a stage (constant)
// The kernel issues the passes inside the LAST k-step of the tile being computed: LDGSTS share the MIO queue
// with the fragment LDS, and a copy burst placed before the tile's fragment loads delays them.
template <bool CPLX, bool CONTIG_K, int VEC, int ROWS, int BK, int NT, bool KY>
struct Loader {
  using G = TileGeom<CPLX, ROWS, BK>;
  static constexpr int EB = CPLX ? 16 : 8;
  static constexpr int VPR = BK / VEC;                       // CONTIG_K: vectors per row
  static constexpr int RPP = NT / VPR;                       //           rows per pass
  static constexpr int VPK = ROWS / VEC;                     // CONTIG_R: vectors per k-row
  static constexpr int KPP = NT / VPK;                       //           k-rows per pass
  static constexpr int NPASS = CONTIG_K ? (ROWS / RPP) : (BK / KPP);
  static constexpr int NROW = CONTIG_K ? NPASS : 1;
  static constexpr int NKO = CONTIG_K ? 1 : NPASS;
  static_assert(NT % (CONTIG_K ? VPR : VPK) == 0, "thread mapping");
  static_assert(NPASS >= 1, "tile too small for the CTA");

  const char* rowptr[NROW];
  bool rowok[NROW];
  long long kb[NKO];     // byte offset of k
  int kabs[NKO];         // absolute k of this thread (pass)
  int kpos[NKO];         // position inside the leading K mode
  bool kok[NKO];
  unsigned dst[NPASS];
  long long kstep;       // BK * stride(leading K mode) * EB
  int kext0;

  __device__ __forceinline__ void init(const Group& g, const Group& gk, const char* base, int row0, int R, int K, int tid) {
    if (CONTIG_K) {
#pragma unroll
      for (int i = 0; i < NPASS; ++i) {
        const int rl = tid / VPR + i * RPP;
        const int r = row0 + rl;
        rowok[i] = r < R;
        rowptr[i] = base + (rowok[i] ? decode<false>(g, r) * EB : 0);
        dst[i] = (unsigned)((rl * G::PK + (tid % VPR) * VEC) * EB);
      }
      kabs[0] = (tid % VPR) * VEC;
    } else {
      const int r = row0 + (tid % VPK) * VEC;
      rowok[0] = r < R;
      rowptr[0] = base + (rowok[0] ? decode<false>(g, r) * EB : 0);
#pragma unroll
      for (int i = 0; i < NPASS; ++i) {
        const int kl = tid / VPK + i * KPP;
        kabs[i] = kl;
        dst[i] = (unsigned)((kl * G::PR + (tid % VPK) * VEC) * EB);
      }
    }
    kext0 = gk.ext[0];
    kstep = (long long)BK * (KY ? gk.sY[0] : gk.sX[0]) * EB;
#pragma unroll
    for (int i = 0; i < NKO; ++i) {
      kok[i] = kabs[i] < K;
      kb[i] = kok[i] ? decode<KY>(gk, kabs[i]) * EB : 0;
      kpos[i] = kabs[i] % kext0;
    }
  }

  // move this thread's k forward by one tile (called once per tile, after the tile's copies were issued)
  __device__ __forceinline__ void advance(const Group& gk, int K) {
#pragma unroll
    for (int i = 0; i < NKO; ++i) {
      kabs[i] += BK;
      kpos[i] += BK;
      kok[i] = kabs[i] < K;
      if (kpos[i] >= kext0) {                   // leading K mode wrapped: re-decode (rare)
        kb[i] = kok[i] ? decode<KY>(gk, kabs[i]) * EB : 0;
        kpos[i] = kabs[i] % kext0;
      } else {
        kb[i] += kstep;
      }
    }
  }

  __device__ __forceinline__ void issue(int pass, unsigned stage_smem) const {
    const int ri = CONTIG_K ? pass : 0, ki = CONTIG_K ? 0 : pass;
    const bool ok = rowok[ri] && kok[ki];
    const char* src = rowptr[ri] + (ok ? kb[ki] : 0);
    const unsigned d = stage_smem + dst[pass];
    const int sz = ok ? VEC * EB : 0;
    if (VEC * EB == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(src), "r"(sz));
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(src), "r"(sz));
  }
};

// ------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------
template <bool CPLX, bool AK, bool BKM, int VA, int VB, class CFG>
__global__ void __launch_bounds__(CFG::NT, CFG::MINB) contract_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BM = CFG::BM, BN = CFG::BN, BK = CFG::BK, NT = CFG::NT, ST = CFG::STAGES;
  constexpr int MI = CFG::MI, NI = CFG::NI;
  constexpr int EB = CPLX ? 16 : 8;
  using GA = TileGeom<CPLX, BM, BK>;
  using GB = TileGeom<CPLX, BN, BK>;
  constexpr int A_BYTES = GA::ELEMS * EB, B_BYTES = GB::ELEMS * EB;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

  extern __shared__ __align__(16) unsigned char smem[];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % CFG::WARPS_M) * CFG::WM;
  const int wn0 = (warp / CFG::WARPS_M) * CFG::WN;
  const int lr = lane >> 2, lc = lane & 3;

  // grouped rasterisation: GROUP_M consecutive row-tiles share each B panel in L2
  int tm, tn;
  {
    const int t = blockIdx.x;
    const int per_group = p.groupM * p.tilesN;
    const int g = t / per_group;
    const int first = g * p.groupM;
    const int gsz = min(p.tilesM - first, p.groupM);
    const int w = t - g * per_group;
    tm = first + w % gsz;
    tn = w / gsz;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  // Hermitian result: only one triangle is needed (1: lower, rank-k update of the tridiagonalisation;
  // 2: upper, Gram matrices feeding the 'U'-convention eigensolver) -- tiles strictly on the other side exit
  if (p.lowerOnly == 1 && m0 + BM <= n0) return;
  if (p.lowerOnly == 2 && n0 + BN <= m0) return;

  using LA = Loader<CPLX, AK, VA, BM, BK, NT, false>;
  using LB = Loader<CPLX, BKM, VB, BN, BK, NT, true>;
  LA la;
  LB lb;

  const char* Ab = (const char*)p.A;
  const char* Bb = (const char*)p.B;
  char* Cb = (char*)p.C;
  if (p.batch > 1 || p.boffA || p.boffB || p.boffC) {
    const int bz = blockIdx.y;
    Ab += (p.boffA ? p.boffA[bz] : bz * p.bstrideA) * EB;
    Bb += (p.boffB ? p.boffB[bz] : bz * p.bstrideB) * EB;
    Cb += (p.boffC ? p.boffC[bz] : bz * p.bstrideC) * EB;
  }
  const int KT = (p.K + BK - 1) / BK;
  la.init(p.gm, p.gk, Ab, m0, p.M, p.K, tid);
  lb.init(p.gn, p.gk, Bb, n0, p.N, p.K, tid);
  const unsigned smem_u32 = (unsigned)__cvta_generic_to_shared(smem);

  double acc[MI][NI][CPLX ? 4 : 2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int e = 0; e < (CPLX ? 4 : 2); ++e) acc[i][j][e] = 0.0;

  // prologue
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) {
    if (s < KT) {
#pragma unroll
      for (int c = 0; c < LA::NPASS; ++c) la.issue(c, smem_u32 + s * STAGE_BYTES);
#pragma unroll
      for (int c = 0; c < LB::NPASS; ++c) lb.issue(c, smem_u32 + s * STAGE_BYTES + A_BYTES);
      la.advance(p.gk, p.K);
      lb.advance(p.gk, p.K);
    }
    cp_async_commit();
  }

  const double sa = p.conjA ? -1.0 : 1.0, sb = p.conjB ? -1.0 : 1.0;
  constexpr int KK = BK / 4;

  constexpr int NCOPY = LA::NPASS + LB::NPASS;
  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<ST - 2>();
    __syncthreads();
    const int nk = kt + ST - 1;
    const bool more = nk < KT;
    const unsigned nsA = smem_u32 + (nk % ST) * STAGE_BYTES;
    const unsigned nsB = nsA + A_BYTES;
    const unsigned char* sA = smem + (kt % ST) * STAGE_BYTES;
    const unsigned char* sB = sA + A_BYTES;
    // copy pass c of the tile ST-1 ahead (A passes first, then B)
    auto copy_pass = [&](int c) {
      if (c < LA::NPASS) la.issue(c, nsA);
      else lb.issue(c - LA::NPASS, nsB);
    };
    if (!CPLX) {
      const double* As = (const double*)sA;
      const double* Bs = (const double*)sB;
      double a[2][MI], b[2][NI];
      auto ldfrag = [&](int kk, double* fa, double* fb) {
        const int k = kk * 4 + lc;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r = wm0 + i * 8 + lr;
          fa[i] = AK ? As[r * GA::PK + k] : As[k * GA::PR + r];
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int r = wn0 + j * 8 + lr;
          fb[j] = BKM ? Bs[r * GB::PK + k] : Bs[k * GB::PR + r];
        }
      };
      constexpr int TOT = MI * NI;   // DMMAs of one k-step per warp
      ldfrag(0, a[0], b[0]);
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        if (kk + 1 < KK) ldfrag(kk + 1, a[(kk + 1) & 1], b[(kk + 1) & 1]);
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            dmma(acc[i][j][0], acc[i][j][1], a[kk & 1][i], b[kk & 1][j]);
            if (kk == KK - 1) {   // copies ride in the last k-step, after all fragment loads of this tile
              const int q = i * NI + j;
              const int c0 = q * NCOPY / TOT, c1 = (q + 1) * NCOPY / TOT;
              if (more && c1 > c0) {
#pragma unroll
                for (int c = c0; c < c1; ++c) copy_pass(c);
              }
            }
          }
      }
    } else {
      const double2* As = (const double2*)sA;
      const double2* Bs = (const double2*)sB;
      double2 a[2][MI], b[2][NI];
      auto ldfrag = [&](int kk, double2* fa, double2* fb) {
        const int k = kk * 4 + lc;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const int r = wm0 + i * 8 + lr;
          fa[i] = AK ? As[r * GA::PK + k] : As[k * GA::PR + r];
          fa[i].y *= sa;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          const int r = wn0 + j * 8 + lr;
          fb[j] = BKM ? Bs[r * GB::PK + k] : Bs[k * GB::PR + r];
          fb[j].y *= sb;
        }
      };
      constexpr int TOT = MI * NI;   // complex DMMA quads of one k-step per warp
      ldfrag(0, a[0], b[0]);
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        if (kk + 1 < KK) ldfrag(kk + 1, a[(kk + 1) & 1], b[(kk + 1) & 1]);
#pragma unroll
        for (int i = 0; i < MI; ++i) {
          const double2 av = a[kk & 1][i];
          const double nai = -av.y;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const double2 bv = b[kk & 1][j];
            dmma(acc[i][j][0], acc[i][j][1], av.x, bv.x);
            dmma(acc[i][j][0], acc[i][j][1], nai, bv.y);
            dmma(acc[i][j][2], acc[i][j][3], av.x, bv.y);
            dmma(acc[i][j][2], acc[i][j][3], av.y, bv.x);
            if (kk == KK - 1) {
              const int q = i * NI + j;
              const int c0 = q * NCOPY / TOT, c1 = (q + 1) * NCOPY / TOT;
              if (more && c1 > c0) {
#pragma unroll
                for (int c = c0; c < c1; ++c) copy_pass(c);
              }
            }
          }
        }
      }
    }
    if (more) { la.advance(p.gk, p.K); lb.advance(p.gk, p.K); }
    cp_async_commit();
  }
  cp_async_wait<0>();

  // epilogue: direct stores from the accumulator fragments (8 consecutive m per quad-column
  // -> full 32-byte sectors when C is M-major, which is the NDTensors output order).
  long long offm[MI];
  bool okm[MI];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int m = m0 + wm0 + i * 8 + lr;
    okm[i] = m < p.M;
    offm[i] = okm[i] ? decode<true>(p.gm, m) : 0;
  }
  const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
  if (has_beta && p.npeer == 0) {
    // beta != 0: read-modify-write.  Loads of C cannot be hoisted over stores to C by the compiler, and one
    // dependent global round trip per element (64 per thread) made a K=128 update epilogue-bound (9 TFLOP/s):
    // fetch the 2*MI old values of a column slice first, then combine and store.
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      long long offn[2];
      bool okn[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int n = n0 + wn0 + j * 8 + lc * 2 + e;
        okn[e] = n < p.N;
        offn[e] = 0;
        if (okn[e]) {
          if (p.splitN && n >= p.splitN) offn[e] = decode<true>(p.gn, n - p.splitN) + (p.boffC2[blockIdx.y] - p.boffC[blockIdx.y]);
          else offn[e] = decode<true>(p.gn, n);
        }
      }
      if (!CPLX) {
        double old[2][MI];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < MI; ++i) old[e][i] = (okn[e] && okm[i]) ? ((const double*)Cb)[offm[i] + offn[e]] : 0.0;
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < MI; ++i)
            if (okn[e] && okm[i]) ((double*)Cb)[offm[i] + offn[e]] = p.alpha_re * acc[i][j][e] + p.beta_re * old[e][i];
      } else {
        double2 old[2][MI];
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < MI; ++i) old[e][i] = (okn[e] && okm[i]) ? ((const double2*)Cb)[offm[i] + offn[e]] : make_double2(0.0, 0.0);
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int i = 0; i < MI; ++i)
            if (okn[e] && okm[i]) {
              const double xr = acc[i][j][e], xi = acc[i][j][2 + e];
              const double2 o = old[e][i];
              double2 v;
              v.x = p.alpha_re * xr - p.alpha_im * xi + p.beta_re * o.x - p.beta_im * o.y;
              v.y = p.alpha_re * xi + p.alpha_im * xr + p.beta_re * o.y + p.beta_im * o.x;
              ((double2*)Cb)[offm[i] + offn[e]] = v;
            }
      }
    }
    return;
  }
#pragma unroll
  for (int j = 0; j < NI; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int n = n0 + wn0 + j * 8 + lc * 2 + e;
      if (n >= p.N) continue;
      long long offn;
      if (p.splitN && n >= p.splitN) offn = decode<true>(p.gn, n - p.splitN) + (p.boffC2[blockIdx.y] - p.boffC[blockIdx.y]);
      else offn = decode<true>(p.gn, n);
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        if (!okm[i]) continue;
        if (!CPLX) {
          const double v = p.alpha_re * acc[i][j][e];
          if (p.npeer > 0) {      // fused all-gather: same element to every GPU's buffer (NVLink peer stores)
            for (int g = 0; g < p.npeer; ++g) ((double*)p.peerC[g])[offm[i] + offn] = v;
          } else {
            ((double*)Cb)[offm[i] + offn] = v;
          }
        } else {
          const double xr = acc[i][j][e], xi = acc[i][j][2 + e];
          double2 v;
          v.x = p.alpha_re * xr - p.alpha_im * xi;
          v.y = p.alpha_re * xi + p.alpha_im * xr;
          if (p.npeer > 0) {
            for (int g = 0; g < p.npeer; ++g) ((double2*)p.peerC[g])[offm[i] + offn] = v;
          } else {
            ((double2*)Cb)[offm[i] + offn] = v;
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// small-K / small-N contractions (K, N <= 32, M huge): H_eff steps 2 and 3 issued as separate contractions
// (K = w d = 10), the middle step of the environment updates, gate application (K = N = d^2).  These are pure
// HBM streaming (read A once, write C once), and the DMMA tile kernel wastes most of its 64x128x16 tile on them
// (1.0 TB/s).  Here one thread owns one m: it loads its K values of A (registers), B sits zero-padded in shared
// memory as [n][k] so that a k-pair is one broadcast LDS.128, and every n is one coalesced store across the warp.
// ------------------------------------------------------------------------------------
template <bool CPLX, int KB>
__global__ void __launch_bounds__(256) smallk_kernel(const __grid_constant__ GemmParams p) {
  using T = typename std::conditional<CPLX, double2, double>::type;
  __shared__ __align__(16) T Bs[32 * KB];
  __shared__ long long koffA[KB], noffC[32];
  const int tid = threadIdx.x;
  for (int e = tid; e < 32 * KB; e += 256) {
    const int n = e / KB, k = e % KB;
    T v;
    if constexpr (CPLX) v = make_double2(0.0, 0.0); else v = 0.0;
    if (n < p.N && k < p.K) {
      v = ((const T*)p.B)[decode<true>(p.gk, k) + decode<false>(p.gn, n)];
      if constexpr (CPLX) { if (p.conjB) v.y = -v.y; }
    }
    Bs[e] = v;
  }
  if (tid < KB) koffA[tid] = tid < p.K ? decode<false>(p.gk, tid) : -1;
  if (tid >= 32 && tid < 64) noffC[tid - 32] = (tid - 32) < p.N ? decode<true>(p.gn, tid - 32) : 0;
  __syncthreads();
  const long long m = (long long)blockIdx.x * 256 + tid;
  if (m >= p.M) return;
  const T* Ap = (const T*)p.A + decode<false>(p.gm, (int)m);
  T* Cp = (T*)p.C + decode<true>(p.gm, (int)m);
  T a[KB];
#pragma unroll
  for (int k = 0; k < KB; ++k) {
    if constexpr (CPLX) a[k] = make_double2(0.0, 0.0); else a[k] = 0.0;
    if (koffA[k] >= 0) {
      a[k] = Ap[koffA[k]];
      if constexpr (CPLX) { if (p.conjA) a[k].y = -a[k].y; }
    }
  }
  const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
  for (int n = 0; n < p.N; ++n) {
    const T* bn = Bs + n * KB;
    if constexpr (!CPLX) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < KB; k += 2) {
        const double2 b2 = *reinterpret_cast<const double2*>(bn + k);
        acc = fma(a[k], b2.x, acc);
        acc = fma(a[k + 1], b2.y, acc);
      }
      double v = p.alpha_re * acc;
      if (has_beta) v += p.beta_re * Cp[noffC[n]];
      Cp[noffC[n]] = v;
    } else {
      double xr = 0.0, xi = 0.0;
#pragma unroll
      for (int k = 0; k < KB; ++k) {
        const double2 b = bn[k];
        xr += a[k].x * b.x - a[k].y * b.y;
        xi += a[k].x * b.y + a[k].y * b.x;
      }
      double2 v = make_double2(p.alpha_re * xr - p.alpha_im * xi, p.alpha_re * xi + p.alpha_im * xr);
      if (has_beta) {
        const double2 o = Cp[noffC[n]];
        v.x += p.beta_re * o.x - p.beta_im * o.y;
        v.y += p.beta_re * o.y + p.beta_im * o.x;
      }
      Cp[noffC[n]] = v;
    }
  }
}

template <bool CPLX>
static int launch_smallk(Handle* h, GemmParams& p, cudaStream_t st) {
  const unsigned grid = (unsigned)(((long long)p.M + 255) / 256);
  if (p.K <= 4) smallk_kernel<CPLX, 4><<<grid, 256, 0, st>>>(p);
  else if (p.K <= 8) smallk_kernel<CPLX, 8><<<grid, 256, 0, st>>>(p);
  else if (p.K <= 16) smallk_kernel<CPLX, 16><<<grid, 256, 0, st>>>(p);
  else smallk_kernel<CPLX, 32><<<grid, 256, 0, st>>>(p);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "smallk_kernel launch");
}

// ------------------------------------------------------------------------------------
// host side: launch
// ------------------------------------------------------------------------------------
// Two independent 4-warp CTAs per SM: their barrier / copy phases drift apart, so one CTA's
// DMMA stream covers the other's bubbles (a single 8-warp CTA measured 83% DMMA-pipe
// utilisation; see profiles/).
using CfgR = Cfg<64, 128, 16, 32, 64, 3, 2>;    // real:    4 warps, warp tile 32x64, 2 CTAs/SM
using CfgRS = Cfg<64, 64, 16, 32, 32, 3, 2>;    // real, small problems: warp tile 32x32
using CfgC = Cfg<64, 64, 8, 32, 32, 3, 2>;      // complex: 4 warps, warp tile 32x32, 2 CTAs/SM
using CfgCS = Cfg<64, 32, 8, 32, 16, 3, 2>;     // complex small

template <bool CPLX, class CFG>
constexpr int smem_bytes() {
  return CFG::STAGES * (TileGeom<CPLX, CFG::BM, CFG::BK>::ELEMS + TileGeom<CPLX, CFG::BN, CFG::BK>::ELEMS) *
             (CPLX ? 16 : 8);
}

template <bool CPLX, bool AK, bool BKM, int VA, int VB, class CFG>
static int launch_one(Handle* h, GemmParams& p, cudaStream_t st) {
  auto kern = contract_kernel<CPLX, AK, BKM, VA, VB, CFG>;
  constexpr int SM = smem_bytes<CPLX, CFG>();
  TNB_ONCE_PER_DEVICE(h, TNB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)));
  p.tilesM = (p.M + CFG::BM - 1) / CFG::BM;
  p.tilesN = (p.N + CFG::BN - 1) / CFG::BN;
  p.groupM = 16;
  const long long tiles = (long long)p.tilesM * p.tilesN;
  if (tiles > 2147483647LL) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: too many tiles");
  if (p.batch < 1) p.batch = 1;
  if (p.batch > 65535) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: batch > 65535");
  kern<<<dim3((unsigned)tiles, (unsigned)p.batch), CFG::NT, SM, st>>>(p);
  h->launches++;
  return check_cuda(h, cudaGetLastError(), "contract_kernel launch");
}

template <bool CPLX, class CFG>
static int launch_cfg(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, cudaStream_t st) {
  if (CPLX) { va = 1; vb = 1; }
#define TNB_L(AKv, BKv, VAv, VBv) return launch_one<CPLX, AKv, BKv, CPLX ? 1 : VAv, CPLX ? 1 : VBv, CFG>(h, p, st)
  if (ak) {
    if (bk) {
      if (va == 2) { if (vb == 2) TNB_L(true, true, 2, 2); else TNB_L(true, true, 2, 1); }
      else         { if (vb == 2) TNB_L(true, true, 1, 2); else TNB_L(true, true, 1, 1); }
    } else {
      if (va == 2) { if (vb == 2) TNB_L(true, false, 2, 2); else TNB_L(true, false, 2, 1); }
      else         { if (vb == 2) TNB_L(true, false, 1, 2); else TNB_L(true, false, 1, 1); }
    }
  } else {
    if (bk) {
      if (va == 2) { if (vb == 2) TNB_L(false, true, 2, 2); else TNB_L(false, true, 2, 1); }
      else         { if (vb == 2) TNB_L(false, true, 1, 2); else TNB_L(false, true, 1, 1); }
    } else {
      if (va == 2) { if (vb == 2) TNB_L(false, false, 2, 2); else TNB_L(false, false, 2, 1); }
      else         { if (vb == 2) TNB_L(false, false, 1, 2); else TNB_L(false, false, 1, 1); }
    }
  }
#undef TNB_L
}


// Pick the tile configuration and launch (one family per translation unit).
template <bool CPLX>
static int launch_tiles(Handle* h, GemmParams& p, bool ak, bool bk, int va, int vb, bool small, cudaStream_t st) {
  if constexpr (!CPLX) {
    if (small) return launch_cfg<false, CfgRS>(h, p, ak, bk, va, vb, st);
    return launch_cfg<false, CfgR>(h, p, ak, bk, va, vb, st);
  } else {
    if (small) return launch_cfg<true, CfgCS>(h, p, ak, bk, va, vb, st);
    return launch_cfg<true, CfgC>(h, p, ak, bk, va, vb, st);
  }
}

}  // namespace tnb
