// TMA-staged form of the contraction kernel: both operands K-major ("TN": H_eff steps 1 and 4, phi = A1 A2^T-type
// products, the projection GEMMs of factorize), the index permutation expressed in the TENSOR MAP instead of in
// per-thread address arithmetic.
//
// Replaces CUTENSOR.contraction! (/root/reference/src/tensor/cudense.jl:328) for those shapes; the LDGSTS kernel of
// contract_kernel.cuh stays for everything a tensor map cannot express (free-major operands, odd strides, more than
// five unmergeable modes per operand, tiles that straddle a mode boundary).
//
//   * One elected thread per CTA issues two `cp.async.bulk.tensor` (SASS: UTMALDG) per k-tile -- A box [BK x BM], B box
//     [BK x BN] of up-to-5-D maps whose dimensions are the merged K modes followed by the merged M (resp. N) modes of
//     the operand, each with its own byte stride: mode order is free, so no permuted copy and no address math in the
//     consumer warps, which issue only LDS + DMMA (round 1's LDGSTS loader spent 12 copies + 12 adds per thread per
//     k-tile and shared the MIO queue with the fragment loads).
//   * 128-byte swizzled shared tiles (row = one M / N index, 16 doubles or 8 complex of K): the 64-bit fragment loads
//     of a half-warp hit 32 distinct banks once the 8 rows of a DMMA fragment are taken in the order 0,2,4,6,1,3,5,7
//     (complex: 0,4,1,5,2,6,3,7 for the 128-bit loads of a quarter-warp) -- a relabelling of rows/columns inside the
//     8x8 DMMA tile that the epilogue undoes.  No padding: 24 KB per stage, 4 stages, 2 CTAs per SM.
//   * mbarrier hand-off both ways: the producer arms full[s] with the byte count and consumers wait on its phase; each
//     warp releases a stage on empty[s] once its DMMAs have consumed the last fragment, and the producer refills the
//     slot two k-tiles ahead.  There is no CTA-wide barrier in the main loop.
//   * Split-K over thread-block clusters for problems smaller than one wave of the 296 CTA slots (chi = 256-512 and edge
//     bonds): the tiles are launched as clusters of 2 or 4 CTAs that each take a K range and reduce through distributed
//     shared memory in a fixed order (deterministic; decided from the shape alone, so every rank of a sharded sweep
//     sums in the same order).  Splitting the partial LAST wave of a large problem was measured and rejected (see
//     tail_split below).
//
// tcgen05 / TMEM has no FP64 kind: the math stays warp-level mma.sync.m8n8k4.f64 = DMMA.8x8x4.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "tnb_internal.h"

namespace tnb {

struct TmaExtra {
  int tileOffset;     // first tile (rasterised index) of this launch
  int nTiles;         // tiles of this launch
  int kSplit;         // CTAs per cluster, each owning a K range
  int nkm, nmm, nnm;  // modes in the K / M / N groups (after merging)
};

template <bool USE_Y>
__device__ __forceinline__ long long tdecode(const Group& g, int idx) {
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

__device__ __forceinline__ void tdmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// one box of a rank-R tensor map -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_load(unsigned dst, const CUtensorMap* map, unsigned bar, int rank, const int* c) {
  const unsigned long long m = (unsigned long long)map;
  switch (rank) {
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1])
                   : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2])
                   : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n" ::"r"(dst),
                   "l"(m), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                   : "memory");
      break;
  }
}

template <int BM_, int BN_, int WM_, int WN_>
struct TCfg {
  static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_;
  static constexpr int WARPS_M = BM / WM, WARPS_N = BN / WN;
  static constexpr int NT = WARPS_M * WARPS_N * 32;
  static constexpr int MI = WM / 8, NI = WN / 8;
  static constexpr int ST = 4;
  static constexpr int STAGE_BYTES = (BM + BN) * 128;
  static constexpr int SMEM = ST * STAGE_BYTES + 1024;      // + slack for the 1024-byte alignment of the swizzle atom
};

template <bool CPLX, class CFG, int PD_>
__global__ void __launch_bounds__(CFG::NT, 2) contract_tma_kernel(const __grid_constant__ CUtensorMap mapA,
                                                                   const __grid_constant__ CUtensorMap mapB,
                                                                   const __grid_constant__ GemmParams p,
                                                                   const __grid_constant__ TmaExtra x) {
  constexpr int BM = CFG::BM, BN = CFG::BN, NT = CFG::NT, ST = CFG::ST, MI = CFG::MI, NI = CFG::NI;
  constexpr int BK = CPLX ? 8 : 16;                     // one 128-byte row of K per tile row
  constexpr int A_BYTES = BM * 128, STAGE_BYTES = CFG::STAGE_BYTES;
  constexpr int NACC = MI * NI * (CPLX ? 4 : 2);
  static_assert(NACC * NT * 8 <= ST * STAGE_BYTES, "split-K reduction buffer must fit the stage ring");

  extern __shared__ unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_bar[2 * ST];      // full[0..ST), empty[ST..2ST)
  const unsigned smem_u32 = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (smem_u32 - (unsigned)__cvta_generic_to_shared(smem_raw));
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(full_bar);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp % CFG::WARPS_M) * CFG::WM;
  const int wn0 = (warp / CFG::WARPS_M) * CFG::WN;
  const int lr = lane >> 2, lc = lane & 3;
  // physical row inside an 8-row group taken by DMMA row/column `lr` (bank-conflict-free swizzled fragment loads)
  const int fr = CPLX ? ((lr >> 1) + 4 * (lr & 1)) : (2 * (lr & 3) + (lr >> 2));

  unsigned crank = 0;
  if (x.kSplit > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(crank));

  // grouped rasterisation: GROUP_M consecutive row-tiles share each B panel in L2
  int tm, tn;
  {
    const int t = x.tileOffset + (int)(blockIdx.x / (unsigned)x.kSplit);
    const int per_group = p.groupM * p.tilesN;
    const int g = t / per_group;
    const int first = g * p.groupM;
    const int gsz = min(p.tilesM - first, p.groupM);
    const int w = t - g * per_group;
    tm = first + w % gsz;
    tn = w / gsz;
  }
  const int m0 = tm * BM, n0 = tn * BN;
  if (p.lowerOnly == 1 && m0 + BM <= n0) return;      // never combined with kSplit > 1 (host)
  if (p.lowerOnly == 2 && n0 + BN <= m0) return;

  const int KT = p.K / BK;                              // K is a multiple of BK (host check)
  int kt0 = 0, kt1 = KT;
  if (x.kSplit > 1) {
    const int per = (KT + x.kSplit - 1) / x.kSplit;
    kt0 = min(KT, (int)crank * per);
    kt1 = min(KT, kt0 + per);
  }
  const int nIt = kt1 - kt0;

  constexpr int PD = PD_;                               // prefetch distance in k-tiles (ST - 2: refill never waits on the tile just finished)
  constexpr int NWARP = NT / 32;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < ST; ++s) {
      mbar_init(bar0 + 8 * s, 1);                       // full[s]:  the producer's arrive + the bytes of both boxes
      mbar_init(bar0 + 8 * (ST + s), NWARP);            // empty[s]: one arrive per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  // ---- producer state (thread 0): coordinates of the A and B boxes
  int ca[5], cb[5];
  const int ra = x.nkm + x.nmm, rb = x.nkm + x.nnm;
  auto set_k = [&](int kt) {
    int idx = kt * BK;
#pragma unroll 1
    for (int i = 0; i < x.nkm; ++i) {
      const int e = p.gk.ext[i];
      const int q = (i + 1 < x.nkm) ? idx / e : 0;
      const int r = (i + 1 < x.nkm) ? idx - q * e : idx;
      ca[i] = cb[i] = (i == 0 && CPLX) ? 2 * r : r;     // complex: the map's inner dimension counts doubles
      idx = q;
    }
  };
  if (tid == 0) {
    int idx = m0;
#pragma unroll 1
    for (int i = 0; i < x.nmm; ++i) {
      const int e = p.gm.ext[i];
      const int q = (i + 1 < x.nmm) ? idx / e : 0;
      ca[x.nkm + i] = (i + 1 < x.nmm) ? idx - q * e : idx;
      idx = q;
    }
    idx = n0;
#pragma unroll 1
    for (int i = 0; i < x.nnm; ++i) {
      const int e = p.gn.ext[i];
      const int q = (i + 1 < x.nnm) ? idx / e : 0;
      cb[x.nkm + i] = (i + 1 < x.nnm) ? idx - q * e : idx;
      idx = q;
    }
  }
  auto issue = [&](int it) {          // thread 0 only: k-tile kt0 + it into slot it % ST
    const int s = it % ST;
    const unsigned bar = bar0 + 8 * s;
    // the slot's previous occupant (k-tile it - ST) must have been read by every warp
    if (it >= ST) mbar_wait(bar0 + 8 * (ST + s), (unsigned)(((it / ST) - 1) & 1));
    set_k(kt0 + it);
    mbar_expect_tx(bar, STAGE_BYTES);
    tma_load(smem_u32 + s * STAGE_BYTES, &mapA, bar, ra, ca);
    tma_load(smem_u32 + s * STAGE_BYTES + A_BYTES, &mapB, bar, rb, cb);
  };
  if (tid == 0) {
#pragma unroll 1
    for (int it = 0; it < PD && it < nIt; ++it) issue(it);
  }

  double acc[MI][NI][CPLX ? 4 : 2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j)
#pragma unroll
      for (int e = 0; e < (CPLX ? 4 : 2); ++e) acc[i][j][e] = 0.0;

  const double sa = p.conjA ? -1.0 : 1.0, sb = p.conjB ? -1.0 : 1.0;
  constexpr int KK = BK / 4;
  static_assert(KK % 2 == 0, "fragment double buffer assumes an even number of k-steps per tile");
  // byte offsets inside a 128-byte row of this thread's k for every k-step (swizzle: 16-byte chunk index ^ (row & 7))
  int koff[KK];
#pragma unroll
  for (int kk = 0; kk < KK; ++kk) {
    const int k = kk * 4 + lc;
    koff[kk] = CPLX ? ((k ^ fr) * 16) : ((((k >> 1) ^ fr) * 16) + (k & 1) * 8);
  }
  const int arow = (wm0 + fr) * 128, brow = (wn0 + fr) * 128;

  // No CTA-wide barrier in the main loop: warps run ahead of one another as far as the stage ring allows, so their
  // tile-boundary bubbles (barrier wait + first fragment loads) do not line up across the 2 warps of an SMSP -- with a
  // __syncthreads per k-tile the DMMA pipe idled 7.7% (ncu, same as round 1's LDGSTS kernel).  The fragments of the
  // first k-step of tile it+1 are loaded during the last k-step of tile it.
  using Frag = typename std::conditional<CPLX, double2, double>::type;
  Frag a[2][MI], b[2][NI];
  auto ldfrag = [&](const unsigned char* sA, int kk, Frag* fa, Frag* fb) {
    const unsigned char* sB = sA + A_BYTES;
#pragma unroll
    for (int i = 0; i < MI; ++i) {
      fa[i] = *reinterpret_cast<const Frag*>(sA + arow + i * 1024 + koff[kk]);
      if constexpr (CPLX) fa[i].y *= sa;
    }
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      fb[j] = *reinterpret_cast<const Frag*>(sB + brow + j * 1024 + koff[kk]);
      if constexpr (CPLX) fb[j].y *= sb;
    }
  };
  if (nIt > 0) {
    mbar_wait(bar0, 0u);
    ldfrag(smem, 0, a[0], b[0]);
  }
#pragma unroll 1
  for (int it = 0; it < nIt; ++it) {
    if (tid == 0 && it + PD < nIt) issue(it + PD);
    const unsigned char* sA = smem + (it % ST) * STAGE_BYTES;
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      if (kk + 1 < KK) {
        ldfrag(sA, kk + 1, a[(kk + 1) & 1], b[(kk + 1) & 1]);
      } else if (it + 1 < nIt) {
        mbar_wait(bar0 + 8 * ((it + 1) % ST), (unsigned)(((it + 1) / ST) & 1));
        ldfrag(smem + ((it + 1) % ST) * STAGE_BYTES, 0, a[0], b[0]);
      }
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        if constexpr (!CPLX) {
#pragma unroll
          for (int j = 0; j < NI; ++j) tdmma(acc[i][j][0], acc[i][j][1], a[kk & 1][i], b[kk & 1][j]);
        } else {
          const double2 av = a[kk & 1][i];
          const double nai = -av.y;
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            const double2 bv = b[kk & 1][j];
            tdmma(acc[i][j][0], acc[i][j][1], av.x, bv.x);
            tdmma(acc[i][j][0], acc[i][j][1], nai, bv.y);
            tdmma(acc[i][j][2], acc[i][j][3], av.x, bv.y);
            tdmma(acc[i][j][2], acc[i][j][3], av.y, bv.x);
          }
        }
      }
    }
    // every fragment of this stage has been consumed by a DMMA (its loads have completed): release the slot
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar0 + 8 * (ST + it % ST)) : "memory");
  }

  // ---- split-K: ranks 1.. of the cluster hand their accumulators to rank 0 through distributed shared memory
  if (x.kSplit > 1) {
    __syncthreads();                                    // the stage ring is free
    double* red = reinterpret_cast<double*>(smem);
    if (crank != 0) {
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j)
#pragma unroll
          for (int e = 0; e < (CPLX ? 4 : 2); ++e) red[((i * NI + j) * (CPLX ? 4 : 2) + e) * NT + tid] = acc[i][j][e];
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (crank == 0) {
      for (int r = 1; r < x.kSplit; ++r) {               // fixed order: rank 1, 2, 3
        unsigned remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(remote) : "r"(smem_u32), "r"(r));
#pragma unroll
        for (int i = 0; i < MI; ++i)
#pragma unroll
          for (int j = 0; j < NI; ++j)
#pragma unroll
            for (int e = 0; e < (CPLX ? 4 : 2); ++e) {
              double v;
              asm volatile("ld.shared::cluster.f64 %0, [%1];\n"
                           : "=d"(v)
                           : "r"(remote + (unsigned)((((i * NI + j) * (CPLX ? 4 : 2) + e) * NT + tid) * 8))
                           : "memory");
              acc[i][j][e] += v;
            }
      }
    }
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    if (crank != 0) return;
  }

  // ---- epilogue: direct stores from the accumulator fragments (rows / columns in the permuted order `fr`)
  char* Cb = (char*)p.C;
  long long offm[MI];
  bool okm[MI];
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int m = m0 + wm0 + i * 8 + fr;
    okm[i] = m < p.M;
    offm[i] = okm[i] ? tdecode<true>(p.gm, m) : 0;
  }
  const bool has_beta = (p.beta_re != 0.0) || (p.beta_im != 0.0);
#pragma unroll
  for (int j = 0; j < NI; ++j) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * lc + e;                                           // DMMA column of this accumulator element
      const int fc = CPLX ? ((c >> 1) + 4 * (c & 1)) : (2 * (c & 3) + (c >> 2));
      const int n = n0 + wn0 + j * 8 + fc;
      if (n >= p.N) continue;
      const long long offn = tdecode<true>(p.gn, n);
      if (has_beta && p.npeer == 0) {
        if (!CPLX) {
          double old[MI];
#pragma unroll
          for (int i = 0; i < MI; ++i) old[i] = okm[i] ? ((const double*)Cb)[offm[i] + offn] : 0.0;
#pragma unroll
          for (int i = 0; i < MI; ++i)
            if (okm[i]) ((double*)Cb)[offm[i] + offn] = p.alpha_re * acc[i][j][e] + p.beta_re * old[i];
        } else {
          double2 old[MI];
#pragma unroll
          for (int i = 0; i < MI; ++i) old[i] = okm[i] ? ((const double2*)Cb)[offm[i] + offn] : make_double2(0.0, 0.0);
#pragma unroll
          for (int i = 0; i < MI; ++i)
            if (okm[i]) {
              const double xr = acc[i][j][e], xi = acc[i][j][2 + e];
              double2 v;
              v.x = p.alpha_re * xr - p.alpha_im * xi + p.beta_re * old[i].x - p.beta_im * old[i].y;
              v.y = p.alpha_re * xi + p.alpha_im * xr + p.beta_re * old[i].y + p.beta_im * old[i].x;
              ((double2*)Cb)[offm[i] + offn] = v;
            }
        }
        continue;
      }
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
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
    else
      cudaGetLastError();
  }
  return fn;
}

using TBig = TCfg<64, 128, 32, 64>;      // 4 warps, warp tile 32x64
using TSmall = TCfg<64, 64, 32, 32>;     // 4 warps, warp tile 32x32
using TCBig = TCfg<64, 64, 32, 32>;      // complex: four real DMMAs per product, warp tile 32x32
using TCSmall = TCfg<64, 32, 32, 16>;

// Can the operand pair (K group + free group) be described by one tensor map of rank <= 5 with a [BK x rows] box?
static bool map_ok(const Group& gk, const Group& gf, bool kIsY, int rows, int bk, size_t es, const void* base) {
  if (gk.n + gf.n > 5 || gk.n + gf.n < 2) return false;
  if ((kIsY ? gk.sY[0] : gk.sX[0]) != 1) return false;
  if (gk.ext[0] % bk) return false;
  if (gf.n > 1 && gf.ext[0] % rows) return false;
  if ((uintptr_t)base % 16) return false;
  for (int i = 1; i < gk.n; ++i) { const long long s = (kIsY ? gk.sY[i] : gk.sX[i]); if (s <= 0 || (s * (long long)es) % 16 || s * (long long)es >= (1LL << 40)) return false; }
  for (int i = 0; i < gf.n; ++i) { const long long s = gf.sX[i]; if (s <= 0 || (s * (long long)es) % 16 || s * (long long)es >= (1LL << 40)) return false; }
  return true;
}

bool tma_shape_ok(const GemmParams& p, int dtype, bool small);

bool tma_eligible(const GemmParams& p, int dtype, bool small) {
  static const bool off = getenv("TNB_TMA") && !strcmp(getenv("TNB_TMA"), "off");
  if (off || !encode_fn()) return false;
  return tma_shape_ok(p, dtype, small);
}

// the shape half of the eligibility test (pure host logic; tnb_plan_describe reports it without a driver)
bool tma_shape_ok(const GemmParams& p, int dtype, bool small) {
  const bool cplx = dtype == TNB_C128;
  const size_t es = cplx ? 16 : 8;
  const int bk = cplx ? 8 : 16;
  const int bm = 64, bn = cplx ? (small ? 32 : 64) : (small ? 64 : 128);
  if (p.batch > 1 || p.boffA || p.boffB || p.boffC || p.splitN) return false;
  if (p.K < 4 * bk || p.K % bk || p.M < bm || p.N < bn) return false;
  if (!map_ok(p.gk, p.gm, false, bm, bk, es, p.A)) return false;
  if (!map_ok(p.gk, p.gn, true, bn, bk, es, p.B)) return false;
  return true;
}

static int make_map(Handle* h, CUtensorMap* map, const Group& gk, const Group& gf, bool kIsY, int rows, int bk, bool cplx,
                    const void* base) {
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  cuuint32_t box[5], estr[5];
  const size_t es = cplx ? 16 : 8;
  int r = 0;
  for (int i = 0; i < gk.n; ++i, ++r) {
    dims[r] = (cuuint64_t)gk.ext[i] * ((i == 0 && cplx) ? 2 : 1);
    box[r] = i == 0 ? (cuuint32_t)(bk * (cplx ? 2 : 1)) : 1;
    if (r > 0) strides[r - 1] = (cuuint64_t)((kIsY ? gk.sY[i] : gk.sX[i]) * (long long)es);
    estr[r] = 1;
  }
  for (int i = 0; i < gf.n; ++i, ++r) {
    dims[r] = (cuuint64_t)gf.ext[i];
    box[r] = i == 0 ? (cuuint32_t)std::min<long long>(rows, 256) : 1;
    strides[r - 1] = (cuuint64_t)(gf.sX[i] * (long long)es);
    estr[r] = 1;
  }
  const CUresult rc = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, (cuuint32_t)r, const_cast<void*>(base), dims, strides, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return set_err(h, TNB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d), rank %d", (int)rc, r);
  return TNB_OK;
}

// Split-K over clusters for problems SMALLER THAN ONE WAVE of CTA slots (chi = 256-512 bonds, edge bonds): their tiles
// run as clusters of 2 or 4 CTAs that split K (a function of the shape only: every rank / every run sums in the same
// order).  Larger problems are never split: measured on the step-4 GEMM of an 8-way shard (1024 tiles on 296 slots =
// 3.46 waves) the un-split launch is 1.5% FASTER (10.55 vs 10.71 ms, profiles/r02_ab_tail_split_step4_n8_shape.jsonl)
// -- the CTAs of a partial last wave have their SM's DMMA pipe to themselves and finish in about half the time, so
// wave quantisation costs far less than the tile count suggests.  TNB_SPLITK=tail restores the tail-wave split.
static uint64_t g_split_launches = 0;
uint64_t tma_split_launches() { return g_split_launches; }

static int tail_split(Handle* h, const GemmParams& p, long long tiles, int KT) {
  static const bool nosplit = getenv("TNB_SPLITK") && !strcmp(getenv("TNB_SPLITK"), "off");
  static const bool tailsplit = getenv("TNB_SPLITK") && !strcmp(getenv("TNB_SPLITK"), "tail");
  const long long slots = 2LL * h->num_sms;
  const long long fullw = tiles / slots, rem = tiles - fullw * slots;
  if (nosplit || rem == 0 || p.lowerOnly) return 1;
  if (fullw > 0 && !tailsplit) return 1;
  const double eff = (double)tiles / (double)((fullw + 1) * slots);
  if (eff >= 0.9) return 1;
  if (rem * 4 <= slots && KT >= 32) return 4;
  if (rem * 2 <= slots && KT >= 16) return 2;
  return 1;
}

bool tma_would_split(Handle* h, const GemmParams& p, int dtype, bool small) {
  const bool cplx = dtype == TNB_C128;
  const int bm = 64, bn = cplx ? (small ? 32 : 64) : (small ? 64 : 128), bk = cplx ? 8 : 16;
  const long long tiles = ((long long)(p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn);
  return tail_split(h, p, tiles, p.K / bk) > 1;
}

template <bool CPLX, class CFG>
static int launch_tma_cfg(Handle* h, GemmParams& p, cudaStream_t st) {
  static const int pd = getenv("TNB_TMA_PD") ? atoi(getenv("TNB_TMA_PD")) : 2;
  auto kern = pd == 3 ? contract_tma_kernel<CPLX, CFG, 3> : contract_tma_kernel<CPLX, CFG, 2>;
  TNB_ONCE_PER_DEVICE(h, TNB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CFG::SMEM)));
  constexpr int BK = CPLX ? 8 : 16;
  p.tilesM = (p.M + CFG::BM - 1) / CFG::BM;
  p.tilesN = (p.N + CFG::BN - 1) / CFG::BN;
  p.groupM = 16;
  const long long tiles = (long long)p.tilesM * p.tilesN;
  if (tiles > 1000000000LL) return set_err(h, TNB_ERR_UNSUPPORTED, "contract: too many tiles");
  CUtensorMap mapA, mapB;
  TNB_TRY(make_map(h, &mapA, p.gk, p.gm, false, CFG::BM, BK, CPLX, p.A));
  TNB_TRY(make_map(h, &mapB, p.gk, p.gn, true, CFG::BN, BK, CPLX, p.B));
  TmaExtra x;
  x.nkm = p.gk.n; x.nmm = p.gm.n; x.nnm = p.gn.n;
  const long long slots = 2LL * h->num_sms;
  const long long fullw = tiles / slots, rem = tiles - fullw * slots;
  const int ks = tail_split(h, p, tiles, p.K / BK);
  auto launch = [&](long long first, long long count, int ksplit) -> int {
    if (count <= 0) return TNB_OK;
    x.tileOffset = (int)first; x.nTiles = (int)count; x.kSplit = ksplit;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(count * ksplit));
    cfg.blockDim = dim3(CFG::NT);
    cfg.dynamicSmemBytes = CFG::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = ksplit; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = ksplit > 1 ? 1 : 0;
    TNB_CUDA(h, cudaLaunchKernelEx(&cfg, kern, mapA, mapB, p, x));
    h->launches++;
    return TNB_OK;
  };
  if (ks == 1) return launch(0, tiles, 1);
  g_split_launches++;
  TNB_TRY(launch(0, fullw * slots, 1));
  return launch(fullw * slots, rem, ks);
}

int launch_tma(Handle* h, int dtype, GemmParams& p, bool small, cudaStream_t st) {
  if (dtype == TNB_C128) return small ? launch_tma_cfg<true, TCSmall>(h, p, st) : launch_tma_cfg<true, TCBig>(h, p, st);
  return small ? launch_tma_cfg<false, TSmall>(h, p, st) : launch_tma_cfg<false, TBig>(h, p, st);
}

}  // namespace tnb
