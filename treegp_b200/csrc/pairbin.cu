// Two-point correlation function by brute-force pair binning.
//
// Replaces the TreeCorr call of /root/reference/treegp/two_pcf.py:297-305 (KKCorrelation,
// bin_type="TwoD", bin_slop=0) and :330-338 (default Log binning, meanr), and -- launched over a batch
// of resampled catalogues -- the bootstrap loop of two_pcf.py:342-362.
//
// TreeCorr itself is not vendored in the reference; the semantics implemented here are the ones
// restated in oracle/pairbin_oracle.c (documented TreeCorr >= 4.2 behaviour):
//   TwoD: keep a pair iff r2 != 0, r2 >= min_sep^2 and max(|dx|,|dy|) < max_sep; column
//         floor((dx+max_sep)/bin_size), row floor((dy+max_sep)/bin_size); each unordered pair is entered
//         at (dx,dy) and at (-dx,-dy).
//   Log : keep iff min_sep^2 <= r2 < max_sep^2; bin floor((0.5 ln r2 - ln min_sep)/bin_size); once per
//         unordered pair; also accumulates sum w r for meanr.
// Bin decisions are made by comparing against host-supplied *thresholds* (the smallest double that
// the formula above sends to bin >= k), so counts are bit-exact with the oracle without FP64
// division or logarithms on the device; r2 is formed with un-fused mul/add for the same reason.
//
// Kernel shape: persistent CTAs walk equal-sized runs of (row-tile, column-tile) pairs of the upper
// triangle of the pair matrix.  Each thread owns one row point in registers; column tiles are staged
// in shared memory and read as broadcasts.  Every pair is evaluated individually.
//
// Accumulation has two paths, chosen per (warp = 32 row points, chunk = 32 column points):
//   * REGISTER path.  When the points are spatially sorted (the host sorts along a Hilbert curve,
//     tgp_hilbert_keys) the displacements of a 32 x 32 block span at most 2 x 2 bins.  The block's
//     bounding boxes give that window exactly (FP subtraction is monotone), one threshold per axis
//     decides the bin, and each lane keeps the four bin sums of the forward (dx,dy) and of the
//     mirrored (-dx,-dy) entry in registers: no atomics in the inner loop.  Counts use
//     inclusion-exclusion on three integer counters (exact).  Blocks entirely inside the range test
//     skip it; blocks entirely outside are skipped.  Registers are flushed (warp shuffle reduction,
//     then one shared-memory add per bin) only when the window moves.
//   * GENERIC path (unsorted input, wide windows, Log bins, the diagonal block): per-pair bin search
//     and shared-memory atomics on warp-privatised histograms.
// Shared histograms are flushed to global memory with red.global once per catalogue.
// Multi-GPU: runs are dealt round-robin to ranks; the caller all-reduces the bin sums.
#include <float.h>
#include <math.h>
#include "tgp_common.cuh"

constexpr int PB_T = 256;       // points per tile (rows per CTA = threads per CTA)
constexpr int PB_WARPS = PB_T / 32;
constexpr int PB_FLUSH_TILEPAIRS = 100000;  // keeps the 32-bit private counters from overflowing

struct PBParams {
  const double *px, *py, *pk, *pw;
  const int64_t* cat_off;
  const double* edges;
  int64_t *npairs;
  double *sumw, *sumwkk, *sumwr;
  double lo2;       // pairs need r2 >= lo2 (= max(min_sep^2, DBL_TRUE_MIN) for TwoD; min_sep^2 for Log)
  double hi;        // TwoD: max_sep (|dx|,|dy| < hi);  Log: max_sep^2 (r2 < hi)
  double inv_bin;   // TwoD: nbins / (2 max_sep)
  int64_t items_per_cat, run;  // run = tile pairs per work item
  int64_t my_items;            // number of work items of this rank
  int32_t ncat, nbins, nb, ncopy, rank, nranks;
};

__device__ __forceinline__ int pb_bin_twod(double d, double hi, double inv_bin, int nbins,
                                           const double* __restrict__ ed) {
  int i = (int)((d + hi) * inv_bin);
  i = max(0, min(i, nbins - 1));
  // exact fix-up: ed[0] = -inf, ed[nbins] = +inf, ed[k] = smallest d that belongs to bin >= k
  if (d < ed[i]) --i;
  else if (d >= ed[i + 1]) ++i;
  return i;
}

// Robust version for window end points (the estimate may be off by more than one there).
__device__ __forceinline__ int pb_bin_twod_search(double d, int nbins, const double* __restrict__ ed) {
  int lo = 0, hi = nbins - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (d >= ed[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ int pb_bin_log(double r2, int nbins, const double* __restrict__ ed) {
  // number of interior thresholds ed[1..nbins-1] that are <= r2 (binary search)
  int lo = 0, hi = nbins - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (r2 >= ed[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-lane register accumulators of the 2 x 2 window, forward (dx,dy) and mirrored (-dx,-dy) entry.
// With px = "column bit" and py = "row bit" of a pair inside the window, the lane keeps
//   tot = sum kk,  sx = sum kk [px],  sy = sum kk [py],  sxy = sum kk [px & py]
// (masks applied as a multiplication by 0.0 / 1.0 inside one FMA, so there is no branch and no select
// on the 64-bit data path) and the same three counters as integers.  The four window bins follow by
// inclusion-exclusion at flush time: exact for the counts, and of the size of ordinary summation
// rounding for the FP64 sums.
template <bool WEIGHTED>
struct RegAcc {
  double tot, fsx, fsy, fsxy, rsx, rsy, rsxy;
  double wtot, fwx, fwy, fwxy, rwx, rwy, rwxy;   // weights (WEIGHTED only)
  unsigned fcx, fcy, fcxy, rcx, rcy, rcxy, nin;
  int fx0, fy0, rx0, ry0;    // window origins (bins); fx0 == -1: no open window
  long long ownerI;          // row tile the per-lane thresholds were derived for
  // Per-lane thresholds on the COLUMN point's coordinates, equivalent to the bin thresholds on the
  // displacement because rounding is monotone:  fl(xj - xi) >= t  <=>  xj >= Tx(xi, t).
  double Tx, Ty, RTx, RTy;   // forward: px = xj >= Tx ; mirrored: px = xj <= RTx
  __device__ __forceinline__ void zero() {
    tot = fsx = fsy = fsxy = rsx = rsy = rsxy = 0.0;
    wtot = fwx = fwy = fwxy = rwx = rwy = rwxy = 0.0;
    fcx = fcy = fcxy = rcx = rcy = rcxy = nin = 0u;
  }
};

// next representable double above / below (finite inputs; NaN and the matching infinity are returned unchanged)
__device__ __forceinline__ double pb_next_up(double v) {
  if (!(v < INFINITY)) return v;
  if (v == 0.0) return __longlong_as_double(1ll);
  const long long b = __double_as_longlong(v);
  return __longlong_as_double(v > 0.0 ? b + 1 : b - 1);
}
__device__ __forceinline__ double pb_next_down(double v) {
  if (!(v > -INFINITY)) return v;
  if (v == 0.0) return __longlong_as_double((long long)0x8000000000000001ull);
  const long long b = __double_as_longlong(v);
  return __longlong_as_double(v > 0.0 ? b - 1 : b + 1);
}

// Smallest X with fl(X - xi) >= t.  `ok` is cleared if the short search did not settle (the caller
// then uses the generic path).  NaN in -> NaN out (dead lanes: every comparison false).
__device__ __forceinline__ double pb_coord_ge(double xi, double t, bool& ok) {
  if (isinf(t) || isnan(xi)) return isnan(xi) ? xi : t;
  double c = t + xi;
  int it = 0;
  while ((c - xi) < t && it < 16) { c = pb_next_up(c); ++it; }
  while ((pb_next_down(c) - xi) >= t && it < 32) { c = pb_next_down(c); ++it; }
  if (!((c - xi) >= t) || ((pb_next_down(c) - xi) >= t)) ok = false;
  return c;
}
// Largest X with fl(X - xi) <= t.
__device__ __forceinline__ double pb_coord_le(double xi, double t, bool& ok) {
  if (isinf(t) || isnan(xi)) return isnan(xi) ? xi : t;
  double c = t + xi;
  int it = 0;
  while ((c - xi) > t && it < 16) { c = pb_next_down(c); ++it; }
  while ((pb_next_up(c) - xi) <= t && it < 32) { c = pb_next_up(c); ++it; }
  if (!((c - xi) <= t) || ((pb_next_up(c) - xi) <= t)) ok = false;
  return c;
}

// One pair into the window registers.  Written in PTX: four compares give the window bits of the
// forward and of the mirrored entry, each masked sum is one FMA with a 0.0 / 1.0 mask, and each counter
// one predicated integer add (nvcc's code for the equivalent C++ needs ~2x the instructions).
#define PB_ONE "0d3FF0000000000000"
#define PB_ZERO "0d0000000000000000"
#define PB_PAIR(A, XJ, YJ, KK)                                                                      \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, qx, qy, pxy, qxy;\n\t"                                                    \
      ".reg .f64 m0, m1, m2, m3, m4, m5;\n\t"                                                       \
      "setp.ge.f64 px, %14, %16;\n\t"                                                               \
      "setp.ge.f64 py, %15, %17;\n\t"                                                               \
      "setp.le.f64 qx, %14, %18;\n\t"                                                               \
      "setp.le.f64 qy, %15, %19;\n\t"                                                               \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "and.pred qxy, qx, qy;\n\t"                                                                   \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      "selp.f64 m3, " PB_ONE ", " PB_ZERO ", qx;\n\t"                                               \
      "selp.f64 m4, " PB_ONE ", " PB_ZERO ", qy;\n\t"                                               \
      "selp.f64 m5, " PB_ONE ", " PB_ZERO ", qxy;\n\t"                                              \
      "add.f64 %0, %0, %13;\n\t"                                                                    \
      "fma.rn.f64 %1, %13, m0, %1;\n\t"                                                             \
      "fma.rn.f64 %2, %13, m1, %2;\n\t"                                                             \
      "fma.rn.f64 %3, %13, m2, %3;\n\t"                                                             \
      "fma.rn.f64 %4, %13, m3, %4;\n\t"                                                             \
      "fma.rn.f64 %5, %13, m4, %5;\n\t"                                                             \
      "fma.rn.f64 %6, %13, m5, %6;\n\t"                                                             \
      "@px add.u32 %7, %7, 1;\n\t"                                                                  \
      "@py add.u32 %8, %8, 1;\n\t"                                                                  \
      "@pxy add.u32 %9, %9, 1;\n\t"                                                                 \
      "@qx add.u32 %10, %10, 1;\n\t"                                                                \
      "@qy add.u32 %11, %11, 1;\n\t"                                                                \
      "@qxy add.u32 %12, %12, 1;\n\t"                                                               \
      "}\n"                                                                                         \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+d"(A.rsx), "+d"(A.rsy), "+d"(A.rsxy), \
        "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy), "+r"(A.rcx), "+r"(A.rcy), "+r"(A.rcxy)              \
      : "d"(KK), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty), "d"(A.RTx), "d"(A.RTy))

#define PB_PAIR_W(A, XJ, YJ, KK, WW)                                                                \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, qx, qy, pxy, qxy;\n\t"                                                    \
      ".reg .f64 m0, m1, m2, m3, m4, m5;\n\t"                                                       \
      "setp.ge.f64 px, %22, %24;\n\t"                                                               \
      "setp.ge.f64 py, %23, %25;\n\t"                                                               \
      "setp.le.f64 qx, %22, %26;\n\t"                                                               \
      "setp.le.f64 qy, %23, %27;\n\t"                                                               \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "and.pred qxy, qx, qy;\n\t"                                                                   \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      "selp.f64 m3, " PB_ONE ", " PB_ZERO ", qx;\n\t"                                               \
      "selp.f64 m4, " PB_ONE ", " PB_ZERO ", qy;\n\t"                                               \
      "selp.f64 m5, " PB_ONE ", " PB_ZERO ", qxy;\n\t"                                              \
      "add.f64 %0, %0, %20;\n\t"                                                                    \
      "fma.rn.f64 %1, %20, m0, %1;\n\t"                                                             \
      "fma.rn.f64 %2, %20, m1, %2;\n\t"                                                             \
      "fma.rn.f64 %3, %20, m2, %3;\n\t"                                                             \
      "fma.rn.f64 %4, %20, m3, %4;\n\t"                                                             \
      "fma.rn.f64 %5, %20, m4, %5;\n\t"                                                             \
      "fma.rn.f64 %6, %20, m5, %6;\n\t"                                                             \
      "add.f64 %7, %7, %21;\n\t"                                                                    \
      "fma.rn.f64 %8, %21, m0, %8;\n\t"                                                             \
      "fma.rn.f64 %9, %21, m1, %9;\n\t"                                                             \
      "fma.rn.f64 %10, %21, m2, %10;\n\t"                                                           \
      "fma.rn.f64 %11, %21, m3, %11;\n\t"                                                           \
      "fma.rn.f64 %12, %21, m4, %12;\n\t"                                                           \
      "fma.rn.f64 %13, %21, m5, %13;\n\t"                                                           \
      "@px add.u32 %14, %14, 1;\n\t"                                                                \
      "@py add.u32 %15, %15, 1;\n\t"                                                                \
      "@pxy add.u32 %16, %16, 1;\n\t"                                                               \
      "@qx add.u32 %17, %17, 1;\n\t"                                                                \
      "@qy add.u32 %18, %18, 1;\n\t"                                                                \
      "@qxy add.u32 %19, %19, 1;\n\t"                                                               \
      "}\n"                                                                                         \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+d"(A.rsx), "+d"(A.rsy), "+d"(A.rsxy), \
        "+d"(A.wtot), "+d"(A.fwx), "+d"(A.fwy), "+d"(A.fwxy), "+d"(A.rwx), "+d"(A.rwy), "+d"(A.rwxy), \
        "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy), "+r"(A.rcx), "+r"(A.rcy), "+r"(A.rcxy)              \
      : "d"(KK), "d"(WW), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty), "d"(A.RTx), "d"(A.RTy))

enum { PB_OUT = 0, PB_REG_FULL = 1, PB_REG_CHECK = 2, PB_GENERIC = 3 };

template <int BT, bool WEIGHTED>
__global__ void __launch_bounds__(PB_T)
pairbin_kernel(PBParams P) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  const int nb = P.nb, ncopy = P.ncopy, nbins = P.nbins;
  double* ed = reinterpret_cast<double*>(pb_smem);                 // nbins + 1 (padded to even)
  double2* txy = reinterpret_cast<double2*>(ed + ((nbins + 2) & ~1));  // PB_T column points (16-byte aligned)
  double* tk = reinterpret_cast<double*>(txy + PB_T);              // PB_T
  double* tw = tk + PB_T;                                          // PB_T
  double* cbox = tw + PB_T;                                        // PB_WARPS * 4: minx, maxx, miny, maxy per chunk
  double* hs = cbox + PB_WARPS * 4;                                // ncopy * nb   sum wk wk
  double* hw = hs + (size_t)ncopy * nb;                            // ncopy * nb   sum w w      (WEIGHTED)
  double* hr = hw + (WEIGHTED ? (size_t)ncopy * nb : 0);           // ncopy * nb   sum w w r    (LOG)
  unsigned int* hc = reinterpret_cast<unsigned int*>(hr + (BT == TGP_BIN_LOG ? (size_t)ncopy * nb : 0));  // counts

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int copy = warp % ncopy;
  double* my_s = hs + (size_t)copy * nb;
  double* my_w = hw + (size_t)copy * nb;
  double* my_r = hr + (size_t)copy * nb;
  unsigned int* my_c = hc + (size_t)copy * nb;

  for (int i = tid; i <= nbins; i += PB_T) ed[i] = P.edges[i];
  auto clear_hist = [&]() {
    for (int i = tid; i < ncopy * nb; i += PB_T) {
      hs[i] = 0.0;
      hc[i] = 0u;
      if constexpr (WEIGHTED) hw[i] = 0.0;
      if (BT == TGP_BIN_LOG) hr[i] = 0.0;
    }
  };

  RegAcc<WEIGHTED> A;
  A.zero();
  A.fx0 = -1;
  // Move the lane registers of the open window into the warp's shared histogram.
  auto flush_regs = [&]() {
    if (BT != TGP_BIN_TWOD) return;
    if (A.fx0 >= 0) {
      const unsigned n_in = warp_sum_u(A.nin);
      const unsigned fcx = warp_sum_u(A.fcx), fcy = warp_sum_u(A.fcy), fcxy = warp_sum_u(A.fcxy);
      const unsigned rcx = warp_sum_u(A.rcx), rcy = warp_sum_u(A.rcy), rcxy = warp_sum_u(A.rcxy);
      const double tot = warp_sum(A.tot);
      const double fsx = warp_sum(A.fsx), fsy = warp_sum(A.fsy), fsxy = warp_sum(A.fsxy);
      const double rsx = warp_sum(A.rsx), rsy = warp_sum(A.rsy), rsxy = warp_sum(A.rsxy);
      double wtot = 0, fwx = 0, fwy = 0, fwxy = 0, rwx = 0, rwy = 0, rwxy = 0;
      if constexpr (WEIGHTED) {
        wtot = warp_sum(A.wtot);
        fwx = warp_sum(A.fwx); fwy = warp_sum(A.fwy); fwxy = warp_sum(A.fwxy);
        rwx = warp_sum(A.rwx); rwy = warp_sum(A.rwy); rwxy = warp_sum(A.rwxy);
      }
      if (lane == 0 && n_in) {
        // inclusion-exclusion per window bin (index = bx + 2*by)
        const unsigned fc[4] = {n_in - fcx - fcy + fcxy, fcx - fcxy, fcy - fcxy, fcxy};
        const unsigned rc[4] = {n_in - rcx - rcy + rcxy, rcx - rcxy, rcy - rcxy, rcxy};
        const double fs[4] = {(tot - fsx) - (fsy - fsxy), fsx - fsxy, fsy - fsxy, fsxy};
        const double rs[4] = {(tot - rsx) - (rsy - rsxy), rsx - rsxy, rsy - rsxy, rsxy};
        const double fw[4] = {(wtot - fwx) - (fwy - fwxy), fwx - fwxy, fwy - fwxy, fwxy};
        const double rw[4] = {(wtot - rwx) - (rwy - rwxy), rwx - rwxy, rwy - rwxy, rwxy};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (fc[b]) {  // a non-empty bin is always inside the grid
            const int o = (A.fy0 + (b >> 1)) * nbins + A.fx0 + (b & 1);
            atomicAdd(my_c + o, fc[b]);
            atomicAdd(my_s + o, fs[b]);
            if constexpr (WEIGHTED) atomicAdd(my_w + o, fw[b]);
          }
          if (rc[b]) {
            const int o = (A.ry0 + (b >> 1)) * nbins + A.rx0 + (b & 1);
            atomicAdd(my_c + o, rc[b]);
            atomicAdd(my_s + o, rs[b]);
            if constexpr (WEIGHTED) atomicAdd(my_w + o, rw[b]);
          }
        }
      }
      A.zero();
      A.fx0 = -1;
    }
  };
  auto flush_hist = [&](int cat) {
    flush_regs();
    __syncthreads();
    if (cat >= 0) {
      for (int b = tid; b < nb; b += PB_T) {
        unsigned long long c = 0;
        double s = 0.0, w = 0.0, r = 0.0;
        for (int k = 0; k < ncopy; ++k) {
          c += hc[k * nb + b];
          s += hs[k * nb + b];
          if constexpr (WEIGHTED) w += hw[k * nb + b];
          if (BT == TGP_BIN_LOG) r += hr[k * nb + b];
        }
        if (c) {
          const size_t o = (size_t)cat * nb + b;
          atomicAdd(reinterpret_cast<unsigned long long*>(P.npairs) + o, c);
          atomicAdd(P.sumwkk + o, s);
          atomicAdd(P.sumw + o, WEIGHTED ? w : (double)c);
          if (BT == TGP_BIN_LOG && P.sumwr) atomicAdd(P.sumwr + o, r);
        }
      }
    }
    __syncthreads();
    clear_hist();
    __syncthreads();
  };
  clear_hist();
  __syncthreads();

  const double M = P.hi, lo2 = P.lo2;
  int cur_cat = -1;
  int since_flush = 0;
  for (int64_t q = blockIdx.x; q < P.my_items; q += gridDim.x) {
    const int64_t item = q * P.nranks + P.rank;
    const int cat = (int)(item / P.items_per_cat);
    if (cat >= P.ncat) break;
    const int64_t local = item % P.items_per_cat;
    const int64_t off = P.cat_off[cat];
    const int64_t n = P.cat_off[cat + 1] - off;
    const int64_t nt = (n + PB_T - 1) / PB_T;
    const int64_t npair_tiles = nt * (nt + 1) / 2;
    int64_t p = local * P.run;
    if (p >= npair_tiles) continue;
    const int64_t p_end = (p + P.run < npair_tiles) ? (p + P.run) : npair_tiles;
    if (cat != cur_cat || since_flush >= PB_FLUSH_TILEPAIRS) {
      flush_hist(cur_cat);
      cur_cat = cat;
      since_flush = 0;
    }
    // p -> (I, J): row-major upper triangle, offset(I) = I*nt - I*(I-1)/2
    const double tn = 2.0 * (double)nt + 1.0;
    int64_t I = (int64_t)((tn - sqrt(tn * tn - 8.0 * (double)p)) * 0.5);
    if (I < 0) I = 0;
    if (I >= nt) I = nt - 1;
    while (I > 0 && I * nt - I * (I - 1) / 2 > p) --I;
    while ((I + 1) * nt - (I + 1) * I / 2 <= p) ++I;
    int64_t J = I + (p - (I * nt - I * (I - 1) / 2));

    int64_t loadedI = -1;
    double xi = 0.0, yi = 0.0, ki = 0.0, wi = 0.0;
    double ib_minx = 0.0, ib_maxx = 0.0, ib_miny = 0.0, ib_maxy = 0.0;  // bounding box of this warp's rows
    bool live = false;
    for (; p < p_end; ++p) {
      if (I != loadedI) {
        const int64_t ig = I * PB_T + tid;
        live = ig < n;
        // dead lanes: NaN coordinates make every comparison false, zero field/weight adds nothing
        xi = live ? P.px[off + ig] : __longlong_as_double(0x7ff8000000000000ll);
        yi = live ? P.py[off + ig] : __longlong_as_double(0x7ff8000000000000ll);
        wi = live ? (WEIGHTED ? P.pw[off + ig] : 1.0) : 0.0;
        ki = live ? P.pk[off + ig] * wi : 0.0;
        if (BT == TGP_BIN_TWOD) {
          ib_minx = warp_min(live ? xi : INFINITY);
          ib_maxx = warp_max(live ? xi : -INFINITY);
          ib_miny = warp_min(live ? yi : INFINITY);
          ib_maxy = warp_max(live ? yi : -INFINITY);
        }
        loadedI = I;
      }
      __syncthreads();  // previous column tile fully consumed
      {
        const int64_t jg = J * PB_T + tid;
        const bool ok = jg < n;
        const double x = ok ? P.px[off + jg] : 0.0, y = ok ? P.py[off + jg] : 0.0;
        txy[tid] = make_double2(x, y);
        const double w = (ok && WEIGHTED) ? P.pw[off + jg] : 1.0;
        tk[tid] = ok ? P.pk[off + jg] * w : 0.0;
        if constexpr (WEIGHTED) tw[tid] = ok ? w : 0.0;
        if (BT == TGP_BIN_TWOD) {  // chunk `warp` of the column tile: bounding box
          const double a = warp_min(ok ? x : INFINITY), b = warp_max(ok ? x : -INFINITY);
          const double c = warp_min(ok ? y : INFINITY), d = warp_max(ok ? y : -INFINITY);
          if (lane == 0) { cbox[warp * 4 + 0] = a; cbox[warp * 4 + 1] = b; cbox[warp * 4 + 2] = c; cbox[warp * 4 + 3] = d; }
        }
      }
      __syncthreads();
      const int jcount = (int)((n - J * PB_T < PB_T) ? (n - J * PB_T) : PB_T);
      const bool diag = (I == J);

      if (BT == TGP_BIN_TWOD) {
        // ---- lanes 0..7 classify chunks 0..7 of this column tile against this warp's rows ----
        int cls = PB_OUT, wx = 0, wy = 0, wrx = 0, wry = 0;
        if (lane < PB_WARPS && lane * 32 < jcount && ib_minx <= ib_maxx) {
          const double cminx = cbox[lane * 4 + 0], cmaxx = cbox[lane * 4 + 1];
          const double cminy = cbox[lane * 4 + 2], cmaxy = cbox[lane * 4 + 3];
          // every dx = x_j - x_i of the block lies in [dx0, dx1] (rounding is monotone)
          const double dx0 = cminx - ib_maxx, dx1 = cmaxx - ib_minx;
          const double dy0 = cminy - ib_maxy, dy1 = cmaxy - ib_miny;
          const double ax = fmax(fabs(dx0), fabs(dx1)), ay = fmax(fabs(dy0), fabs(dy1));   // max |dx|, |dy|
          const double nx = dx0 > 0.0 ? dx0 : (dx1 < 0.0 ? -dx1 : 0.0);                    // min |dx|
          const double ny = dy0 > 0.0 ? dy0 : (dy1 < 0.0 ? -dy1 : 0.0);
          const double r2max = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
          const double r2min = __dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny));
          if (nx >= M || ny >= M || r2max < lo2) {
            cls = PB_OUT;
          } else if (diag && lane <= warp) {
            // diagonal tile: chunks before this warp's rows hold only j < i; the warp's own chunk needs j > i
            cls = (lane == warp) ? PB_GENERIC : PB_OUT;
          } else {
            const int x0 = pb_bin_twod_search(dx0, nbins, ed), x1 = pb_bin_twod_search(dx1, nbins, ed);
            const int y0 = pb_bin_twod_search(dy0, nbins, ed), y1 = pb_bin_twod_search(dy1, nbins, ed);
            const int rx0 = pb_bin_twod_search(-dx1, nbins, ed), rx1 = pb_bin_twod_search(-dx0, nbins, ed);
            const int ry0 = pb_bin_twod_search(-dy1, nbins, ed), ry1 = pb_bin_twod_search(-dy0, nbins, ed);
            if (x1 - x0 > 1 || y1 - y0 > 1 || rx1 - rx0 > 1 || ry1 - ry0 > 1) {
              cls = PB_GENERIC;
            } else {
              cls = (ax < M && ay < M && r2min >= lo2) ? PB_REG_FULL : PB_REG_CHECK;
              wx = x0 | ((x1 - x0) << 16); wy = y0 | ((y1 - y0) << 16);
              wrx = rx0 | ((rx1 - rx0) << 16); wry = ry0 | ((ry1 - ry0) << 16);
            }
          }
        }
        const int nchunk = (jcount + 31) >> 5;
        for (int c = 0; c < nchunk; ++c) {
          const int ccls = __shfl_sync(0xffffffffu, cls, c);
          if (ccls == PB_OUT) continue;
          const int j0 = c * 32;
          const int jn = (jcount - j0 < 32) ? (jcount - j0) : 32;
          if (ccls == PB_GENERIC) {
            const int jstart = (diag && c == warp) ? lane + 1 : 0;
            if (live) {
              for (int jj = j0 + jstart; jj < j0 + jn; ++jj) {
                const double2 pj = txy[jj];
                const double dx = pj.x - xi, dy = pj.y - yi;
                const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (r2 >= lo2 && fabs(dx) < M && fabs(dy) < M) {
                  const int b1 = pb_bin_twod(dy, M, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(dx, M, P.inv_bin, nbins, ed);
                  const int b2 = pb_bin_twod(-dy, M, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(-dx, M, P.inv_bin, nbins, ed);
                  const double kk = ki * tk[jj];
                  atomicAdd(my_c + b1, 1u);
                  atomicAdd(my_c + b2, 1u);
                  atomicAdd(my_s + b1, kk);
                  atomicAdd(my_s + b2, kk);
                  if constexpr (WEIGHTED) {
                    const double ww = wi * tw[jj];
                    atomicAdd(my_w + b1, ww);
                    atomicAdd(my_w + b2, ww);
                  }
                }
              }
            }
            __syncwarp();
            continue;
          }
          // ---- register path: make sure the open window covers this block ----
          const int cwx = __shfl_sync(0xffffffffu, wx, c), cwy = __shfl_sync(0xffffffffu, wy, c);
          const int cwrx = __shfl_sync(0xffffffffu, wrx, c), cwry = __shfl_sync(0xffffffffu, wry, c);
          const int x0 = cwx & 0xffff, x1 = x0 + (cwx >> 16), y0 = cwy & 0xffff, y1 = y0 + (cwy >> 16);
          const int rx0 = cwrx & 0xffff, rx1 = rx0 + (cwrx >> 16), ry0 = cwry & 0xffff, ry1 = ry0 + (cwry >> 16);
          const bool fits = A.fx0 >= 0 && x0 >= A.fx0 && x1 <= A.fx0 + 1 && y0 >= A.fy0 && y1 <= A.fy0 + 1 &&
                            rx0 >= A.rx0 && rx1 <= A.rx0 + 1 && ry0 >= A.ry0 && ry1 <= A.ry0 + 1;
          bool reg_ok = true;
          if (!fits || A.ownerI != loadedI) {
            flush_regs();
            // a window [b0, b0+1] must stay inside the grid unless nbins == 1
            A.fx0 = min(x0, max(nbins - 2, 0)); A.fy0 = min(y0, max(nbins - 2, 0));
            A.rx0 = min(rx0, max(nbins - 2, 0)); A.ry0 = min(ry0, max(nbins - 2, 0));
            A.ownerI = loadedI;
            // bit = (bin >= b0 + 1): forward dx >= ed[b0+1]; mirrored -dx >= ed[b0+1] <=> dx <= -ed[b0+1];
            // turned into thresholds on the column coordinate for this lane's row point
            const double tx = (nbins > 1) ? ed[A.fx0 + 1] : INFINITY, ty = (nbins > 1) ? ed[A.fy0 + 1] : INFINITY;
            const double ntx = (nbins > 1) ? -ed[A.rx0 + 1] : -INFINITY, nty = (nbins > 1) ? -ed[A.ry0 + 1] : -INFINITY;
            bool okl = true;
            A.Tx = pb_coord_ge(xi, tx, okl);
            A.Ty = pb_coord_ge(yi, ty, okl);
            A.RTx = pb_coord_le(xi, ntx, okl);
            A.RTy = pb_coord_le(yi, nty, okl);
            reg_ok = __all_sync(0xffffffffu, okl);
            if (!reg_ok) A.fx0 = -1;  // nothing accumulated yet: simply close the window again
          }
          if (!reg_ok) {
            // (never seen in practice) per-lane thresholds did not settle: generic path for this block
            if (live) {
              for (int jj = j0; jj < j0 + jn; ++jj) {
                const double2 pj = txy[jj];
                const double dx = pj.x - xi, dy = pj.y - yi;
                const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (r2 >= lo2 && fabs(dx) < M && fabs(dy) < M) {
                  const int b1 = pb_bin_twod(dy, M, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(dx, M, P.inv_bin, nbins, ed);
                  const int b2 = pb_bin_twod(-dy, M, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(-dx, M, P.inv_bin, nbins, ed);
                  const double kk = ki * tk[jj];
                  atomicAdd(my_c + b1, 1u); atomicAdd(my_c + b2, 1u);
                  atomicAdd(my_s + b1, kk); atomicAdd(my_s + b2, kk);
                  if constexpr (WEIGHTED) { const double ww = wi * tw[jj]; atomicAdd(my_w + b1, ww); atomicAdd(my_w + b2, ww); }
                }
              }
            }
            __syncwarp();
            continue;
          }
          if (ccls == PB_REG_FULL) {
#pragma unroll 4
            for (int jj = j0; jj < j0 + jn; ++jj) {
              const double2 pj = txy[jj];
              const double kk = ki * tk[jj];
              if constexpr (WEIGHTED) {
                const double ww = wi * tw[jj];
                PB_PAIR_W(A, pj.x, pj.y, kk, ww);
              } else {
                PB_PAIR(A, pj.x, pj.y, kk);
              }
            }
            A.nin += live ? (unsigned)jn : 0u;
          } else {
#pragma unroll 2
            for (int jj = j0; jj < j0 + jn; ++jj) {
              const double2 pj = txy[jj];
              const double dx = pj.x - xi, dy = pj.y - yi;
              const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
              const bool ok = r2 >= lo2 && fabs(dx) < M && fabs(dy) < M;  // false for dead lanes (NaN)
              if (ok) {
                const double kk = ki * tk[jj];
                if constexpr (WEIGHTED) {
                  const double ww = wi * tw[jj];
                  PB_PAIR_W(A, pj.x, pj.y, kk, ww);
                } else {
                  PB_PAIR(A, pj.x, pj.y, kk);
                }
                A.nin += 1u;
              }
            }
          }
        }
      } else {
        // ---- Log bins: generic path ----
        const int jstart = diag ? tid + 1 : 0;  // diagonal tile: j > i only
        if (live) {
          for (int jj = jstart; jj < jcount; ++jj) {
            const double2 pj = txy[jj];
            const double dx = pj.x - xi, dy = pj.y - yi;
            const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // no FMA: matches the oracle bit for bit
            if (r2 >= lo2 && r2 < M) {
              const int b = pb_bin_log(r2, nbins, ed);
              const double ww = WEIGHTED ? wi * tw[jj] : 1.0;
              atomicAdd(my_c + b, 1u);
              atomicAdd(my_s + b, ki * tk[jj]);
              if constexpr (WEIGHTED) atomicAdd(my_w + b, ww);
              atomicAdd(my_r + b, ww * sqrt(r2));
            }
          }
        }
      }
      ++since_flush;
      if (++J == nt) { ++I; J = I; }
    }
    // the per-lane 32-bit counters are bounded by the flush policy below: flush the registers at
    // the end of every work item (cheap: once per `run` tile pairs)
    flush_regs();
  }
  flush_hist(cur_cat);
}

extern "C" int tgp_pairbin_tile(void) { return PB_T; }

extern "C" int tgp_pairbin(const double* px, const double* py, const double* pk, const double* pw,
                           const int64_t* cat_off, int32_t ncat, int64_t max_cat_len, int32_t bin_type,
                           const double* edges, int32_t nbins, double min_sep2, double max_sep,
                           int32_t tile_rank, int32_t tile_nranks, int64_t* npairs, double* sumw,
                           double* sumwkk, double* sumwr, void* stream) {
  TGP_CHECK_ARG(bin_type == TGP_BIN_TWOD || bin_type == TGP_BIN_LOG, "bin_type");
  TGP_CHECK_ARG(ncat >= 0 && max_cat_len >= 0 && nbins >= 1, "ncat/max_cat_len/nbins");
  TGP_CHECK_ARG(tile_nranks >= 1 && tile_rank >= 0 && tile_rank < tile_nranks, "rank");
  TGP_CHECK_ARG(max_sep > 0.0 && min_sep2 >= 0.0, "separations");
  if (ncat == 0 || max_cat_len < 2) return TGP_OK;
  TGP_CHECK_ARG(px && py && pk && cat_off && edges && npairs && sumw && sumwkk, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;

  PBParams P;
  P.px = px; P.py = py; P.pk = pk; P.pw = pw; P.cat_off = cat_off; P.edges = edges;
  P.npairs = npairs; P.sumw = sumw; P.sumwkk = sumwkk; P.sumwr = sumwr;
  P.ncat = ncat; P.nbins = nbins; P.rank = tile_rank; P.nranks = tile_nranks;
  const bool twod = bin_type == TGP_BIN_TWOD;
  TGP_CHECK_ARG(!twod || nbins <= 4096, "nbins too large");
  P.nb = twod ? nbins * nbins : nbins;
  if (twod) {
    P.lo2 = min_sep2 > 0.0 ? min_sep2 : 4.9406564584124654e-324;  // r2 != 0
    P.hi = max_sep;
    P.inv_bin = (double)nbins / (2.0 * max_sep);
  } else {
    P.lo2 = min_sep2;
    P.hi = max_sep * max_sep;
    P.inv_bin = 0.0;
  }
  const bool weighted = pw != nullptr;

  // shared-memory budget -> number of private histogram copies
  const size_t per_copy = (size_t)P.nb * (8 + 4 + (weighted ? 8 : 0) + (twod ? 0 : 8));
  const size_t fixed = (size_t)((nbins + 2) & ~1) * 8 + 4 * PB_T * 8 + PB_WARPS * 4 * 8;
  const size_t budget = 200 * 1024;
  TGP_CHECK_ARG(fixed + per_copy <= budget, "too many bins for the shared-memory histogram");
  int ncopy = (int)((budget - fixed) / per_copy);
  if (ncopy > PB_WARPS) ncopy = PB_WARPS;
  // keep at least 2 CTAs per SM when that costs no privatisation below 4 copies
  while (ncopy > 4 && fixed + ncopy * per_copy > 100 * 1024) --ncopy;
  P.ncopy = ncopy;
  const size_t smem = fixed + (size_t)ncopy * per_copy;

  // work decomposition
  const int64_t nt = tgp_cdiv(max_cat_len, PB_T);
  const int64_t tile_pairs = nt * (nt + 1) / 2;
  const int sms = tgp_num_sms();
  const int ctas_per_sm = (smem <= 100 * 1024) ? 2 : 1;
  const int64_t grid_target = (int64_t)sms * ctas_per_sm;
  // aim for >= 16 work items per CTA over the whole batch, runs of at most 64 tile pairs
  int64_t run = (tile_pairs * ncat) / (grid_target * 16 * tile_nranks);
  if (run < 1) run = 1;
  if (run > 64) run = 64;
  P.run = run;
  P.items_per_cat = tgp_cdiv(tile_pairs, run);
  const int64_t total_items = P.items_per_cat * ncat;
  P.my_items = (total_items - tile_rank + tile_nranks - 1) / tile_nranks;
  if (P.my_items <= 0) return TGP_OK;
  const int64_t grid = P.my_items < grid_target ? P.my_items : grid_target;

#define TGP_PB_LAUNCH(BT, W)                                                                          \
  do {                                                                                                \
    TGP_CUDA(cudaFuncSetAttribute(pairbin_kernel<BT, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                  (int)budget));                                                      \
    pairbin_kernel<BT, W><<<(unsigned)grid, PB_T, smem, st>>>(P);                                     \
  } while (0)
  if (twod) {
    if (weighted) TGP_PB_LAUNCH(TGP_BIN_TWOD, true); else TGP_PB_LAUNCH(TGP_BIN_TWOD, false);
  } else {
    if (weighted) TGP_PB_LAUNCH(TGP_BIN_LOG, true); else TGP_PB_LAUNCH(TGP_BIN_LOG, false);
  }
#undef TGP_PB_LAUNCH
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

// ============================================================================================
// Hilbert-curve keys: sorting the points by this key makes consecutive points spatial neighbours,
// which is what lets the pair-binning kernel keep whole 32 x 32 blocks inside a 2 x 2 bin window.
// ============================================================================================
__global__ void hilbert_keys_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                    double x0, double y0, double inv_cell, int order, int64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned side = 1u << order;
  long long cx = (long long)floor((x[i] - x0) * inv_cell), cy = (long long)floor((y[i] - y0) * inv_cell);
  unsigned ux = (unsigned)(cx < 0 ? 0 : (cx >= (long long)side ? side - 1 : cx));
  unsigned uy = (unsigned)(cy < 0 ? 0 : (cy >= (long long)side ? side - 1 : cy));
  // classic xy -> d conversion
  unsigned long long d = 0;
  for (unsigned s = side >> 1; s > 0; s >>= 1) {
    const unsigned rx = (ux & s) ? 1u : 0u, ry = (uy & s) ? 1u : 0u;
    d += (unsigned long long)s * s * ((3u * rx) ^ ry);
    if (ry == 0) {
      if (rx == 1) { ux = side - 1 - ux; uy = side - 1 - uy; }
      const unsigned t = ux; ux = uy; uy = t;
    }
  }
  keys[i] = (int64_t)d;
}

extern "C" int tgp_hilbert_keys(const double* x, const double* y, int64_t n, double xmin, double ymin,
                                double extent, int32_t order, int64_t* keys, void* stream) {
  TGP_CHECK_ARG(n >= 0 && order >= 1 && order <= 30 && extent > 0.0, "n/order/extent");
  if (n == 0) return TGP_OK;
  TGP_CHECK_ARG(x && y && keys, "null pointer");
  const double inv_cell = (double)(1u << order) / extent;
  hilbert_keys_kernel<<<(unsigned)tgp_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, n, xmin, ymin, inv_cell,
                                                                                  order, keys);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}
