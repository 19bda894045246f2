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
// Kernel shape (warp-centric, no CTA barriers in the main loop).  The pair matrix is cut into
// blocks of 32 row points x 32 column points.  A work item is one row block times a run of column
// chunks; warps of persistent CTAs pull items from a global counter (dynamic load balance: blocks
// that are provably out of range are skipped, so items differ in cost).  A lane owns one row point in
// registers; the warp stages one column chunk at a time in its private slice of shared memory and
// reads it back as broadcasts.
//
// Two instantiations per bin type (template parameter BS):
//   BS = false  the pair-by-pair kernel: every pair of every in-range block is evaluated individually
//               (tgp_set_option("pairbin_block_sums", 0)); the 10-FP64-ops-per-pair roofline is stated for it.
//               Blocks that provably sit in one bin per axis whose mirrored window is the mirror image skip the
//               per-pair mirrored bits (the forward bits are still evaluated for every pair).
//   BS = true   (default) block forms on top of the same classification, all exact:
//               * TwoD, block in ONE forward bin whose mirrored window is its mirror image: booked whole by the
//                 classifying lane from the chunk sums of the pre-pass (points never loaded);
//               * TwoD, one varying axis: rank query (4-ary search, 8 probes per row point) on the chunk's copy
//                 sorted along that axis with suffix sums (pre-pass); the exact mirrored bits need only the two
//                 neighbours of the split, a failed check hands the block back to the pair-by-pair loop.  Blocks
//                 that the open window already covers take a short cut in front of the general dispatch;
//               * TwoD, two bins along both axes: column-bit and row-bit sums from two rank queries on the x- and
//                 y-sorted copies, only the "both bits" quadrant summed per pair;
//               * Log, block in one radial bin or two adjacent ones: counts / k-sums from row and chunk sums,
//                 one square root per pair for sum w r, no atomics.
//
// For 32 chunks at a time each lane derives the bounding box of "its" chunk and classifies the
// block against the warp's row bounding box (FP subtraction is monotone, so the box gives exact bounds
// on dx, dy):
//   OUT        every pair fails the range test                          -> skipped, not even loaded
//   REG_FULL   every pair passes it and the displacements span <= 2 x 2 bins (both entries)
//   REG_CHECK  same window, range test needed per pair
//   GENERIC    wider window (unsorted input), Log bins, the diagonal block
// REG blocks accumulate in registers (see RegAcc): no atomics in the inner loop; the registers are
// flushed to the warp's private shared-memory histogram only when the window moves.  When the points
// are sorted along a Hilbert curve (tgp_hilbert_keys) nearly all blocks are REG; a GENERIC block is
// first retried as four 8-column sub-blocks.  The GENERIC path does a per-pair bin search and
// shared-memory atomics.  Histograms go to global memory with red.global when the warp changes
// catalogue.  Multi-GPU: items are dealt round-robin to ranks; the caller all-reduces the bin sums.
#include <float.h>
#include <math.h>
#include <atomic>
#include "tgp_common.cuh"

constexpr int PB_CHUNK = 32;          // column points per block (= row points per warp)
constexpr int PB_MAX_WARPS = 8;       // warps per CTA (each owns a private histogram)
constexpr int PB_FLUSH_ITEMS = 2048;  // keeps the 32-bit private counters from overflowing
constexpr int PB_SLOT = 8;            // doubles per chunk slot of the pre-pass: box + sums
constexpr int PB_SORTED = 6 * PB_CHUNK;  // + {sorted x, suffix k w, suffix w, sorted y, suffix k w, suffix w}
constexpr int PB_STRIDE = PB_SLOT + PB_SORTED;   // the two records of a chunk are stored back to back

struct PBParams {
  const double *px, *py, *pk, *pw;
  const int64_t* cat_off;
  const double* edges;
  int64_t* npairs;
  double *sumw, *sumwkk, *sumwr;
  unsigned long long* counter;  // dynamic work counter (zeroed before the launch)
  const double* boxes;          // optional precomputed chunk boxes + sums (8 doubles per slot) or NULL
  const double* sorted;         // optional per-chunk sorted coordinates + suffix sums (PB_SORTED doubles per slot)
  double lo2;       // pairs need r2 >= lo2 (= max(min_sep^2, DBL_TRUE_MIN) for TwoD; min_sep^2 for Log)
  double hi;        // TwoD: max_sep (|dx|,|dy| < hi);  Log: max_sep^2 (r2 < hi)
  double inv_bin;   // TwoD: nbins / (2 max_sep)
  int64_t items_per_cat;  // upper bound from max_cat_len
  int64_t my_items;       // number of work items of this rank
  int32_t run;            // column chunks per work item (multiple of 32)
  int32_t ncat, nbins, nb, warps, rank, nranks;
  int32_t block_sums;     // 1: blocks whose pairs provably share a window bit use the block forms (default)
  int32_t fast_paths;     // bit 0: short-cut dispatch of one-axis blocks that fit the open window; bit 1: 2 x 2-window
                          // blocks with marginal sums from rank queries; bit 2: pair-by-pair kernel, no mirrored bits
                          // per pair in blocks that provably sit in one bin and its mirror image (default: all)
};

// Which path the pairs of all launches since the last reset took (in pairs): [0] closed form (block in one bin),
// [1] one varying axis, pair by pair, [2] pair by pair (2 x 2 window with or without the per-pair range test,
// generic sub-blocks), [3] one varying axis answered by a rank query on the sorted chunk, [4] 2 x 2 window with the
// marginal sums from rank queries (one quadrant pair by pair).  Diagnostics.
__device__ unsigned long long g_pb_stats[8];

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

// The same search out of line (rarely reached), and the bin of d given a guess g in [0, nbins-1]: two compares
// when the guess is right (ed[0] = -inf, ed[nbins] = +inf, thresholds ascending: the bin is unique).
__device__ __noinline__ int pb_bin_twod_search_cold(double d, int nbins, const double* ed) {
  return pb_bin_twod_search(d, nbins, ed);
}
__device__ __forceinline__ int pb_bin_twod_guess(double d, int g, int nbins, const double* __restrict__ ed) {
  if (d >= ed[g] && d < ed[g + 1]) return g;
  return pb_bin_twod_search_cold(d, nbins, ed);
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

// Per-lane register accumulators of the 2 x 2 window.
// With px = "column bit" and py = "row bit" of a pair inside the window, the lane keeps
//   tot = sum kk,  sx = sum kk [px],  sy = sum kk [py],  sxy = sum kk [px & py]
// (masks applied as a multiplication by 0.0 / 1.0 inside one FMA: no branch, no 64-bit select) and the same
// three counters as integers; the four window bins follow by inclusion-exclusion at flush time (exact for
// the counts, of the size of ordinary summation rounding for the FP64 sums).
//
// Mirrored entry.  Every pair is also entered at (-dx,-dy).  The grid is point symmetric, so the mirrored
// bin of a pair is "almost always" the mirror image (nbins-1-ix, nbins-1-iy) of its forward bin; it can
// differ only when a displacement sits within rounding of a bin edge, because the thresholds of the
// TreeCorr formula are not exactly symmetric in floating point.  The window of the mirrored entry is
// therefore fixed to the mirror image of the forward window, the exact mirrored bits qx, qy are still
// evaluated for every pair, and `mmc` counts the pairs for which (qx, qy) != (!px, !py).  The forward
// registers are flushed into both the forward bins and their mirror images; blocks with mmc != 0 are
// rescanned and the (rare) inconsistent pairs moved from the assumed to the exact mirrored bin.
template <bool WEIGHTED>
struct RegAcc {
  double tot, fsx, fsy, fsxy;
  double wtot, fwx, fwy, fwxy;   // weights (WEIGHTED only)
  unsigned fcx, fcy, fcxy, nin, mmc;
  int fx0, fy0;              // forward window origin (bins); fx0 == -1: no open window.  Mirrored window
                             // origin: (nbins-2-fx0, nbins-2-fy0)
  long long ownerI;          // row block the per-lane thresholds were derived for
  double rki, rwi;           // this lane's row-point factors k_i w_i and w_i, applied to the sums at flush time
  // Per-lane thresholds on the COLUMN point's coordinates, equivalent to the bin thresholds on the
  // displacement because rounding is monotone:  fl(xj - xi) >= t  <=>  xj >= Tx(xi, t).
  double Tx, Ty, RTx, RTy;   // forward: px = xj >= Tx ; mirrored: qx = xj <= RTx
  __device__ __forceinline__ void zero() {
    tot = fsx = fsy = fsxy = 0.0;
    wtot = fwx = fwy = fwxy = 0.0;
    fcx = fcy = fcxy = nin = mmc = 0u;
  }
};

// Monotone map between doubles and signed integers (order-preserving; -0.0 and +0.0 share key 0).
__device__ __forceinline__ long long pb_key(double d) {
  const long long b = __double_as_longlong(d);
  return b >= 0 ? b : -(b & 0x7fffffffffffffffll);
}
__device__ __forceinline__ double pb_unkey(long long k) {
  return __longlong_as_double(k >= 0 ? k : (long long)(0x8000000000000000ull | (unsigned long long)(-k)));
}

// Smallest X with fl(X - xi) >= t.  X -> fl(X - xi) is monotone, so the answer is found by bisection over
// the ordered doubles inside a bracket of a few ulp(t) + ulp(t + xi) around t + xi.  `ok` is cleared if no
// bracket was found (the caller then uses the generic path).  NaN xi -> NaN (dead lanes: every comparison
// false); infinite t is returned unchanged (nbins == 1: the bit is constant).
__device__ __noinline__ double pb_coord_ge(double xi, double t, bool& ok) {
  if (isinf(t) || isnan(xi)) return isnan(xi) ? xi : t;
  const double c = t + xi;
  double e = (fabs(t) + fabs(c)) * 0x1p-50 + 0x1p-1060;
  double lo = c - e, hi = c + e;
  for (int it = 0; it < 8 && (!((hi - xi) >= t) || ((lo - xi) >= t)); ++it) { e *= 8.0; lo = c - e; hi = c + e; }
  if (!((hi - xi) >= t) || ((lo - xi) >= t)) { ok = false; return c; }
  long long klo = pb_key(lo), khi = pb_key(hi);  // pred(lo) false, pred(hi) true
  while (khi - klo > 1) {
    const long long mid = klo + ((khi - klo) >> 1);
    if ((pb_unkey(mid) - xi) >= t) khi = mid; else klo = mid;
  }
  return pb_unkey(khi);
}
// Largest X with fl(X - xi) <= t.
__device__ __noinline__ double pb_coord_le(double xi, double t, bool& ok) {
  if (isinf(t) || isnan(xi)) return isnan(xi) ? xi : t;
  const double c = t + xi;
  double e = (fabs(t) + fabs(c)) * 0x1p-50 + 0x1p-1060;
  double lo = c - e, hi = c + e;
  for (int it = 0; it < 8 && (!((lo - xi) <= t) || ((hi - xi) <= t)); ++it) { e *= 8.0; lo = c - e; hi = c + e; }
  if (!((lo - xi) <= t) || ((hi - xi) <= t)) { ok = false; return c; }
  long long klo = pb_key(lo), khi = pb_key(hi);  // pred(lo) true, pred(hi) false
  while (khi - klo > 1) {
    const long long mid = klo + ((khi - klo) >> 1);
    if ((pb_unkey(mid) - xi) <= t) klo = mid; else khi = mid;
  }
  return pb_unkey(klo);
}

// One pair into the window registers.  Written in PTX: four compares give the window bits of the
// forward entry and the exact bits of the mirrored entry, each masked sum is one FMA with a 0.0 / 1.0
// mask, each counter one predicated integer add (nvcc's code for the equivalent C++ needs ~2x the
// instructions; ptxas turns a predicated FP64 add into add + two selects, so the mask form is the shortest).
// What is summed is the COLUMN point's value k_j w_j (and w_j): the row point's factor k_i w_i (w_i) is
// constant for a lane while a window is open and is applied once, when the registers are flushed.
// TOT = "add.f64 tot, tot, kj" for blocks that need the range test per pair; blocks that are entirely in
// range add the chunk sum once instead.
#define PB_ONE "0d3FF0000000000000"
#define PB_ZERO "0d0000000000000000"
#define PB_PAIR_BODY(TOT)                                                                           \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, qx, qy, pxy, cx, cy, mm;\n\t"                                             \
      ".reg .f64 m0, m1, m2;\n\t"                                                                   \
      "setp.ge.f64 px, %9, %11;\n\t"                                                                \
      "setp.ge.f64 py, %10, %12;\n\t"                                                               \
      "setp.le.f64 qx, %9, %13;\n\t"                                                                \
      "setp.le.f64 qy, %10, %14;\n\t"                                                               \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "xor.pred cx, px, qx;\n\t"                                                                    \
      "xor.pred cy, py, qy;\n\t"                                                                    \
      "and.pred mm, cx, cy;\n\t"                                                                    \
      "not.pred mm, mm;\n\t"                                                                        \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      TOT                                                                                           \
      "fma.rn.f64 %1, %8, m0, %1;\n\t"                                                              \
      "fma.rn.f64 %2, %8, m1, %2;\n\t"                                                              \
      "fma.rn.f64 %3, %8, m2, %3;\n\t"                                                              \
      "@px add.u32 %4, %4, 1;\n\t"                                                                  \
      "@py add.u32 %5, %5, 1;\n\t"                                                                  \
      "@pxy add.u32 %6, %6, 1;\n\t"                                                                 \
      "@mm add.u32 %7, %7, 1;\n\t"                                                                  \
      "}\n"
#define PB_PAIR_OPS(A, XJ, YJ, KJ)                                                                  \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy), \
        "+r"(A.mmc)                                                                                 \
      : "d"(KJ), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty), "d"(A.RTx), "d"(A.RTy)
#define PB_PAIR(A, XJ, YJ, KJ) asm volatile(PB_PAIR_BODY("add.f64 %0, %0, %8;\n\t") PB_PAIR_OPS(A, XJ, YJ, KJ))
#define PB_PAIR_NT(A, XJ, YJ, KJ) asm volatile(PB_PAIR_BODY("") PB_PAIR_OPS(A, XJ, YJ, KJ))

#define PB_PAIR_W_BODY(TOT)                                                                         \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, qx, qy, pxy, cx, cy, mm;\n\t"                                             \
      ".reg .f64 m0, m1, m2;\n\t"                                                                   \
      "setp.ge.f64 px, %14, %16;\n\t"                                                               \
      "setp.ge.f64 py, %15, %17;\n\t"                                                               \
      "setp.le.f64 qx, %14, %18;\n\t"                                                               \
      "setp.le.f64 qy, %15, %19;\n\t"                                                               \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "xor.pred cx, px, qx;\n\t"                                                                    \
      "xor.pred cy, py, qy;\n\t"                                                                    \
      "and.pred mm, cx, cy;\n\t"                                                                    \
      "not.pred mm, mm;\n\t"                                                                        \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      TOT                                                                                           \
      "fma.rn.f64 %1, %12, m0, %1;\n\t"                                                             \
      "fma.rn.f64 %2, %12, m1, %2;\n\t"                                                             \
      "fma.rn.f64 %3, %12, m2, %3;\n\t"                                                             \
      "fma.rn.f64 %5, %13, m0, %5;\n\t"                                                             \
      "fma.rn.f64 %6, %13, m1, %6;\n\t"                                                             \
      "fma.rn.f64 %7, %13, m2, %7;\n\t"                                                             \
      "@px add.u32 %8, %8, 1;\n\t"                                                                  \
      "@py add.u32 %9, %9, 1;\n\t"                                                                  \
      "@pxy add.u32 %10, %10, 1;\n\t"                                                               \
      "@mm add.u32 %11, %11, 1;\n\t"                                                                \
      "}\n"
#define PB_PAIR_W_OPS(A, XJ, YJ, KJ, WJ)                                                            \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+d"(A.wtot), "+d"(A.fwx), "+d"(A.fwy), \
        "+d"(A.fwxy), "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy), "+r"(A.mmc)                           \
      : "d"(KJ), "d"(WJ), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty), "d"(A.RTx), "d"(A.RTy)
#define PB_PAIR_W(A, XJ, YJ, KJ, WJ)                                                                \
  asm volatile(PB_PAIR_W_BODY("add.f64 %0, %0, %12;\n\tadd.f64 %4, %4, %13;\n\t") PB_PAIR_W_OPS(A, XJ, YJ, KJ, WJ))
#define PB_PAIR_W_NT(A, XJ, YJ, KJ, WJ) asm volatile(PB_PAIR_W_BODY("") PB_PAIR_W_OPS(A, XJ, YJ, KJ, WJ))

// The same pair without the mirrored bits (no mismatch counter): for blocks whose bounding boxes prove that the
// mirrored bin of every pair is the mirror image of its forward bin.
#define PB_PAIR_NM_BODY(TOT)                                                                        \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, pxy;\n\t"                                                                 \
      ".reg .f64 m0, m1, m2;\n\t"                                                                   \
      "setp.ge.f64 px, %8, %10;\n\t"                                                                \
      "setp.ge.f64 py, %9, %11;\n\t"                                                                \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      TOT                                                                                           \
      "fma.rn.f64 %1, %7, m0, %1;\n\t"                                                              \
      "fma.rn.f64 %2, %7, m1, %2;\n\t"                                                              \
      "fma.rn.f64 %3, %7, m2, %3;\n\t"                                                              \
      "@px add.u32 %4, %4, 1;\n\t"                                                                  \
      "@py add.u32 %5, %5, 1;\n\t"                                                                  \
      "@pxy add.u32 %6, %6, 1;\n\t"                                                                 \
      "}\n"
#define PB_PAIR_NM_OPS(A, XJ, YJ, KJ)                                                               \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy) \
      : "d"(KJ), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty)
#define PB_PAIR_NM(A, XJ, YJ, KJ) asm volatile(PB_PAIR_NM_BODY("add.f64 %0, %0, %7;\n\t") PB_PAIR_NM_OPS(A, XJ, YJ, KJ))
#define PB_PAIR_NM_NT(A, XJ, YJ, KJ) asm volatile(PB_PAIR_NM_BODY("") PB_PAIR_NM_OPS(A, XJ, YJ, KJ))
#define PB_PAIR_NM_W_BODY(TOT)                                                                      \
      "{\n\t"                                                                                       \
      ".reg .pred px, py, pxy;\n\t"                                                                 \
      ".reg .f64 m0, m1, m2;\n\t"                                                                   \
      "setp.ge.f64 px, %13, %15;\n\t"                                                               \
      "setp.ge.f64 py, %14, %16;\n\t"                                                               \
      "and.pred pxy, px, py;\n\t"                                                                   \
      "selp.f64 m0, " PB_ONE ", " PB_ZERO ", px;\n\t"                                               \
      "selp.f64 m1, " PB_ONE ", " PB_ZERO ", py;\n\t"                                               \
      "selp.f64 m2, " PB_ONE ", " PB_ZERO ", pxy;\n\t"                                              \
      TOT                                                                                           \
      "fma.rn.f64 %1, %11, m0, %1;\n\t"                                                             \
      "fma.rn.f64 %2, %11, m1, %2;\n\t"                                                             \
      "fma.rn.f64 %3, %11, m2, %3;\n\t"                                                             \
      "fma.rn.f64 %5, %12, m0, %5;\n\t"                                                             \
      "fma.rn.f64 %6, %12, m1, %6;\n\t"                                                             \
      "fma.rn.f64 %7, %12, m2, %7;\n\t"                                                             \
      "@px add.u32 %8, %8, 1;\n\t"                                                                  \
      "@py add.u32 %9, %9, 1;\n\t"                                                                  \
      "@pxy add.u32 %10, %10, 1;\n\t"                                                               \
      "}\n"
#define PB_PAIR_NM_W_OPS(A, XJ, YJ, KJ, WJ)                                                         \
      : "+d"(A.tot), "+d"(A.fsx), "+d"(A.fsy), "+d"(A.fsxy), "+d"(A.wtot), "+d"(A.fwx), "+d"(A.fwy), \
        "+d"(A.fwxy), "+r"(A.fcx), "+r"(A.fcy), "+r"(A.fcxy)                                        \
      : "d"(KJ), "d"(WJ), "d"(XJ), "d"(YJ), "d"(A.Tx), "d"(A.Ty)
#define PB_PAIR_NM_W(A, XJ, YJ, KJ, WJ)                                                             \
  asm volatile(PB_PAIR_NM_W_BODY("add.f64 %0, %0, %11;\n\tadd.f64 %4, %4, %12;\n\t") PB_PAIR_NM_W_OPS(A, XJ, YJ, KJ, WJ))
#define PB_PAIR_NM_W_NT(A, XJ, YJ, KJ, WJ) asm volatile(PB_PAIR_NM_W_BODY("") PB_PAIR_NM_W_OPS(A, XJ, YJ, KJ, WJ))

// One pair of a block whose displacements span two bins along ONE axis only (the other window bit is the same
// for every pair of the block): CJ = the column point's coordinate on the varying axis, T / RT the lane's
// forward / mirrored thresholds on it.  bs / bw / bc = block-local masked sums and count of the "upper bin" pairs.
#define PB_PAIR_1D(BS, BC, MMC, CJ, KJ, T, RT)                                                      \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred p, c;\n\t"                                                                        \
      ".reg .f64 m;\n\t"                                                                            \
      "setp.ge.f64 p, %3, %5;\n\t"                                                                  \
      "setp.le.f64 c, %3, %6;\n\t"                                                                  \
      "xor.pred c, c, p;\n\t"                                                                       \
      "selp.f64 m, " PB_ONE ", " PB_ZERO ", p;\n\t"                                                 \
      "fma.rn.f64 %0, %4, m, %0;\n\t"                                                               \
      "@p add.u32 %1, %1, 1;\n\t"                                                                   \
      "@!c add.u32 %2, %2, 1;\n\t"                                                                  \
      "}\n"                                                                                         \
      : "+d"(BS), "+r"(BC), "+r"(MMC)                                                               \
      : "d"(CJ), "d"(KJ), "d"(T), "d"(RT))
#define PB_PAIR_1D_W(BS, BW, BC, MMC, CJ, KJ, WJ, T, RT)                                            \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred p, c;\n\t"                                                                        \
      ".reg .f64 m;\n\t"                                                                            \
      "setp.ge.f64 p, %4, %7;\n\t"                                                                  \
      "setp.le.f64 c, %4, %8;\n\t"                                                                  \
      "xor.pred c, c, p;\n\t"                                                                       \
      "selp.f64 m, " PB_ONE ", " PB_ZERO ", p;\n\t"                                                 \
      "fma.rn.f64 %0, %5, m, %0;\n\t"                                                               \
      "fma.rn.f64 %1, %6, m, %1;\n\t"                                                               \
      "@p add.u32 %2, %2, 1;\n\t"                                                                   \
      "@!c add.u32 %3, %3, 1;\n\t"                                                                  \
      "}\n"                                                                                         \
      : "+d"(BS), "+d"(BW), "+r"(BC), "+r"(MMC)                                                     \
      : "d"(CJ), "d"(KJ), "d"(WJ), "d"(T), "d"(RT))

// One pair of a 2 x 2-window block whose marginal sums (column bit, row bit) come from rank queries on the sorted
// copies of the chunk: only the "both bits set" quadrant is accumulated per pair.
#define PB_PAIR_Q(BS, BC, XJ, YJ, KJ, TX, TY)                                                       \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred p, q;\n\t"                                                                        \
      ".reg .f64 m;\n\t"                                                                            \
      "setp.ge.f64 q, %2, %5;\n\t"                                                                  \
      "setp.ge.and.f64 p, %3, %6, q;\n\t"                                                           \
      "selp.f64 m, " PB_ONE ", " PB_ZERO ", p;\n\t"                                                 \
      "fma.rn.f64 %0, %4, m, %0;\n\t"                                                               \
      "@p add.u32 %1, %1, 1;\n\t"                                                                   \
      "}\n"                                                                                         \
      : "+d"(BS), "+r"(BC)                                                                          \
      : "d"(XJ), "d"(YJ), "d"(KJ), "d"(TX), "d"(TY))
#define PB_PAIR_Q_W(BS, BW, BC, XJ, YJ, KJ, WJ, TX, TY)                                             \
  asm volatile(                                                                                     \
      "{\n\t"                                                                                       \
      ".reg .pred p, q;\n\t"                                                                        \
      ".reg .f64 m;\n\t"                                                                            \
      "setp.ge.f64 q, %3, %7;\n\t"                                                                  \
      "setp.ge.and.f64 p, %4, %8, q;\n\t"                                                           \
      "selp.f64 m, " PB_ONE ", " PB_ZERO ", p;\n\t"                                                 \
      "fma.rn.f64 %0, %5, m, %0;\n\t"                                                               \
      "fma.rn.f64 %1, %6, m, %1;\n\t"                                                               \
      "@p add.u32 %2, %2, 1;\n\t"                                                                   \
      "}\n"                                                                                         \
      : "+d"(BS), "+d"(BW), "+r"(BC)                                                                \
      : "d"(XJ), "d"(YJ), "d"(KJ), "d"(WJ), "d"(TX), "d"(TY))

// Rank query in a chunk's sorted copy read straight from global memory (L1-resident after the first-level probes
// p7, p15, p23 = xs[7], xs[15], xs[23], which the caller issues early): pos = number of columns with c_j < T;
// returns whether the exact mirrored bits agree for this lane (see rank_query in the kernel).
__device__ __forceinline__ bool pb_rank_query_g(const double* __restrict__ xs, double p7, double p15, double p23,
                                                double T, double RT, bool live, int& pos) {
  const int b = 8 * ((p7 < T ? 1 : 0) + (p15 < T ? 1 : 0) + (p23 < T ? 1 : 0));
  const int sp = b + 2 * ((xs[b + 1] < T ? 1 : 0) + (xs[b + 3] < T ? 1 : 0) + (xs[b + 5] < T ? 1 : 0));
  pos = sp + (xs[sp] < T ? 1 : 0) + (xs[sp + 1] < T ? 1 : 0);
  return !live || ((pos == 0 || xs[pos - 1] <= RT) && (pos == PB_CHUNK || xs[pos] > RT));
}
__device__ __forceinline__ void pb_prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
}

enum { PB_OUT = 0, PB_REG_FULL = 1, PB_REG_CHECK = 2, PB_GENERIC = 3 };

// Classification of one (row block, column block) pair from the two bounding boxes.
// Returns the class; for REG classes w[0..3] receive the packed windows (origin | extent << 16) of the
// forward x, forward y, mirrored x and mirrored y bins.
__device__ __forceinline__ int pb_classify_twod(double iminx, double imaxx, double iminy, double imaxy,
                                                double cminx, double cmaxx, double cminy, double cmaxy,
                                                double M, double lo2, int nbins, const double* __restrict__ ed,
                                                int (&w)[4]) {
  // every dx = x_j - x_i of the block lies in [dx0, dx1] (rounding is monotone)
  const double dx0 = cminx - imaxx, dx1 = cmaxx - iminx;
  const double dy0 = cminy - imaxy, dy1 = cmaxy - iminy;
  const double ax = fmax(fabs(dx0), fabs(dx1)), ay = fmax(fabs(dy0), fabs(dy1));   // max |dx|, |dy|
  const double nx = dx0 > 0.0 ? dx0 : (dx1 < 0.0 ? -dx1 : 0.0);                    // min |dx|
  const double ny = dy0 > 0.0 ? dy0 : (dy1 < 0.0 ? -dy1 : 0.0);
  const double r2max = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
  const double r2min = __dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny));
  if (nx >= M || ny >= M || r2max < lo2) return PB_OUT;
  const int x0 = pb_bin_twod_search(dx0, nbins, ed), x1 = pb_bin_twod_search(dx1, nbins, ed);
  const int y0 = pb_bin_twod_search(dy0, nbins, ed), y1 = pb_bin_twod_search(dy1, nbins, ed);
  // mirrored ends: the grid is point symmetric up to rounding of the thresholds, so bin(-d) is nbins-1-bin(d)
  // unless d sits within a few ulp of an edge; the guess is verified against the thresholds (exact either way)
  const int rx0 = pb_bin_twod_guess(-dx1, nbins - 1 - x1, nbins, ed), rx1 = pb_bin_twod_guess(-dx0, nbins - 1 - x0, nbins, ed);
  const int ry0 = pb_bin_twod_guess(-dy1, nbins - 1 - y1, nbins, ed), ry1 = pb_bin_twod_guess(-dy0, nbins - 1 - y0, nbins, ed);
  if (x1 - x0 > 1 || y1 - y0 > 1 || rx1 - rx0 > 1 || ry1 - ry0 > 1) return PB_GENERIC;
  w[0] = x0 | ((x1 - x0) << 16);
  w[1] = y0 | ((y1 - y0) << 16);
  w[2] = rx0 | ((rx1 - rx0) << 16);
  w[3] = ry0 | ((ry1 - ry0) << 16);
  return (ax < M && ay < M && r2min >= lo2) ? PB_REG_FULL : PB_REG_CHECK;
}

#ifndef PB_MIN_CTAS
#define PB_MIN_CTAS 2
#endif
// BS: block forms on (closed-form / one-axis blocks).  BS = false is the pair-by-pair kernel: every pair of every
// in-range block goes through the compare / masked-FMA loop (TwoD) -- the kernel the 10-FP64-ops-per-pair
// roofline is stated for; it is compiled without any of the block-form code.
template <int BT, bool WEIGHTED, bool BS>
__global__ void __launch_bounds__(PB_MAX_WARPS * 32, PB_MIN_CTAS)
pairbin_kernel(PBParams P) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  const int nb = P.nb, nbins = P.nbins, nwarps = P.warps;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- shared memory: thresholds (CTA), then per warp: column chunk buffer + private histogram ----
  double* ed = reinterpret_cast<double*>(pb_smem);                          // nbins + 1 (padded to even)
  double2* cxy_all = reinterpret_cast<double2*>(ed + ((nbins + 2) & ~1));   // [warps][32]
  double* ck_all = reinterpret_cast<double*>(cxy_all + nwarps * PB_CHUNK);  // [warps][32]
  double* cw_all = ck_all + nwarps * PB_CHUNK;                              // [warps][32]
  double* hs_all = cw_all + nwarps * PB_CHUNK;                              // [warps][nb]  sum wk wk
  double* hw_all = hs_all + (size_t)nwarps * nb;                            // [warps][nb]  sum w w     (WEIGHTED)
  double* hr_all = hw_all + (WEIGHTED ? (size_t)nwarps * nb : 0);           // [warps][nb]  sum w w r   (LOG)
  unsigned* hc_all = reinterpret_cast<unsigned*>(hr_all + (BT == TGP_BIN_LOG ? (size_t)nwarps * nb : 0));
  double2* cxy = cxy_all + warp * PB_CHUNK;
  double* ck = ck_all + warp * PB_CHUNK;
  double* cw = cw_all + warp * PB_CHUNK;
  double* my_s = hs_all + (size_t)warp * nb;
  double* my_w = hw_all + (size_t)warp * nb;
  double* my_r = hr_all + (size_t)warp * nb;
  unsigned* my_c = hc_all + (size_t)warp * nb;

  for (int i = tid; i <= nbins; i += blockDim.x) ed[i] = P.edges[i];
  for (int i = lane; i < nb; i += 32) {
    my_s[i] = 0.0;
    my_c[i] = 0u;
    if constexpr (WEIGHTED) my_w[i] = 0.0;
    if (BT == TGP_BIN_LOG) my_r[i] = 0.0;
  }
  __syncthreads();  // the only CTA-wide barrier

  RegAcc<WEIGHTED> A;
  A.zero();
  A.fx0 = -1;
  A.ownerI = -1;
  // TwoD: the warp histogram holds the FORWARD entries only; every booking is point symmetric, so flush_hist adds
  // the mirror image (bin nb-1-b) when it writes a bin out.  The rare pair whose exact mirrored bin is not the
  // mirror image of its forward bin is corrected directly in global memory: -1 at the assumed bin, +1 at the exact.
  int cur_cat = -1;
  auto move_mirrored = [&](int assumed, int exact, double kk, double ww) {
    const size_t g = (size_t)cur_cat * nb;
    atomicAdd(reinterpret_cast<unsigned long long*>(P.npairs) + g + assumed, ~0ull);   // -1 (mod 2^64)
    atomicAdd(reinterpret_cast<unsigned long long*>(P.npairs) + g + exact, 1ull);
    atomicAdd(P.sumwkk + g + assumed, -kk);
    atomicAdd(P.sumwkk + g + exact, kk);
    atomicAdd(P.sumw + g + assumed, -ww);
    atomicAdd(P.sumw + g + exact, ww);
  };
  // Move the lane registers of the open window into the warp's shared histogram.
  auto flush_regs = [&]() {
    if (BT != TGP_BIN_TWOD) return;
    if (A.fx0 >= 0) {
      const unsigned n_in = warp_sum_u(A.nin);
      const unsigned fcx = warp_sum_u(A.fcx), fcy = warp_sum_u(A.fcy), fcxy = warp_sum_u(A.fcxy);
      // the lane sums hold sum_j k_j w_j [mask]: times the row point's k_i w_i they are the pair sums
      const double tot = warp_sum(A.tot * A.rki);
      const double fsx = warp_sum(A.fsx * A.rki), fsy = warp_sum(A.fsy * A.rki), fsxy = warp_sum(A.fsxy * A.rki);
      double wtot = 0, fwx = 0, fwy = 0, fwxy = 0;
      if constexpr (WEIGHTED) {
        wtot = warp_sum(A.wtot * A.rwi);
        fwx = warp_sum(A.fwx * A.rwi); fwy = warp_sum(A.fwy * A.rwi); fwxy = warp_sum(A.fwxy * A.rwi);
      }
      if (lane == 0 && n_in) {
        // inclusion-exclusion per window bin (index = bx + 2*by)
        const unsigned fc[4] = {n_in - fcx - fcy + fcxy, fcx - fcxy, fcy - fcxy, fcxy};
        const double fs[4] = {(tot - fsx) - (fsy - fsxy), fsx - fsxy, fsy - fsxy, fsxy};
        const double fw[4] = {(wtot - fwx) - (fwy - fwxy), fwx - fwxy, fwy - fwxy, fwxy};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          if (fc[b]) {  // a non-empty bin is always inside the grid; the histogram is private to this warp
            const int bx = A.fx0 + (b & 1), by = A.fy0 + (b >> 1);
            const int o = by * nbins + bx;       // forward entry; flush_hist adds the mirror image
            my_c[o] += fc[b];
            my_s[o] += fs[b];
            if constexpr (WEIGHTED) my_w[o] += fw[b];
          }
        }
      }
      __syncwarp();
      A.zero();
      A.fx0 = -1;
    }
  };
  // Rare: some pair of the block just processed has a mirrored bin that is not the mirror image of its
  // forward bin.  Rescan the block and move those pairs from the assumed to the exact mirrored bin.
  auto fix_mirror = [&](int j0, int jn, bool check, double xi, double yi, double ki, double wi) {
    if (!__any_sync(0xffffffffu, A.mmc != 0u)) return;
    if (A.mmc) {
      const int rx0 = nbins - 2 - A.fx0, ry0 = nbins - 2 - A.fy0;
      for (int jj = j0; jj < j0 + jn; ++jj) {
        const double2 pj = cxy[jj];
        if (check) {
          const double dx = pj.x - xi, dy = pj.y - yi;
          const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (!(r2 >= P.lo2 && fabs(dx) < P.hi && fabs(dy) < P.hi)) continue;
        }
        const bool px = pj.x >= A.Tx, py = pj.y >= A.Ty, qx = pj.x <= A.RTx, qy = pj.y <= A.RTy;
        if ((px != qx) && (py != qy)) continue;  // consistent
        const int assumed = (nbins - 1 - (A.fy0 + (py ? 1 : 0))) * nbins + (nbins - 1 - (A.fx0 + (px ? 1 : 0)));
        const int exact = (ry0 + (qy ? 1 : 0)) * nbins + (rx0 + (qx ? 1 : 0));
        double ww = 1.0;
        if constexpr (WEIGHTED) ww = wi * cw[jj];
        move_mirrored(assumed, exact, ki * ck[jj], ww);
      }
      A.mmc = 0u;
    }
    __syncwarp();
  };
  // Warp-private shared histogram -> global (red.global), then clear.
  auto flush_hist = [&](int cat) {   // the window registers were flushed at the end of the last item
    __syncwarp();
    if (cat >= 0) {
      if (BT == TGP_BIN_TWOD) {
        for (int b = lane; b < nb; b += 32) {       // forward entries of bin b + those of its mirror image
          const int bm = nb - 1 - b;
          const unsigned long long c = (unsigned long long)my_c[b] + (unsigned long long)my_c[bm];
          if (c) {
            const size_t o = (size_t)cat * nb + b;
            atomicAdd(reinterpret_cast<unsigned long long*>(P.npairs) + o, c);
            atomicAdd(P.sumwkk + o, my_s[b] + my_s[bm]);
            atomicAdd(P.sumw + o, WEIGHTED ? my_w[b] + my_w[bm] : (double)c);
          }
        }
        __syncwarp();   // every bin is read twice: clear after all reads
        for (int b = lane; b < nb; b += 32) {
          my_c[b] = 0u;
          my_s[b] = 0.0;
          if constexpr (WEIGHTED) my_w[b] = 0.0;
        }
      } else {
        for (int b = lane; b < nb; b += 32) {
          const unsigned c = my_c[b];
          if (c) {
            const size_t o = (size_t)cat * nb + b;
            atomicAdd(reinterpret_cast<unsigned long long*>(P.npairs) + o, (unsigned long long)c);
            atomicAdd(P.sumwkk + o, my_s[b]);
            atomicAdd(P.sumw + o, WEIGHTED ? my_w[b] : (double)c);
            if (P.sumwr) atomicAdd(P.sumwr + o, my_r[b]);
            my_c[b] = 0u;
            my_s[b] = 0.0;
            if constexpr (WEIGHTED) my_w[b] = 0.0;
            my_r[b] = 0.0;
          }
        }
      }
    }
    __syncwarp();
  };

  // per-pair generic accumulation of columns [j0, j0+jn) of the staged chunk; `jfirst` = first column this
  // lane may pair with (diagonal block: j > i)
  auto generic_block = [&](int j0, int jn, int jfirst, double xi, double yi, double ki, double wi, bool live) {
    if (live) {
      for (int jj = max(j0, jfirst); jj < j0 + jn; ++jj) {
        const double2 pj = cxy[jj];
        const double dx = pj.x - xi, dy = pj.y - yi;
        const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // no FMA: matches the oracle bit for bit
        if (BT == TGP_BIN_TWOD) {
          if (r2 >= P.lo2 && fabs(dx) < P.hi && fabs(dy) < P.hi) {
            const int b1 = pb_bin_twod(dy, P.hi, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(dx, P.hi, P.inv_bin, nbins, ed);
            const int b2 = pb_bin_twod(-dy, P.hi, P.inv_bin, nbins, ed) * nbins + pb_bin_twod(-dx, P.hi, P.inv_bin, nbins, ed);
            const double kk = ki * ck[jj];
            double ww = 1.0;
            if constexpr (WEIGHTED) ww = wi * cw[jj];
            atomicAdd(my_c + b1, 1u);
            atomicAdd(my_s + b1, kk);
            if constexpr (WEIGHTED) atomicAdd(my_w + b1, ww);
            if (b2 != nb - 1 - b1) move_mirrored(nb - 1 - b1, b2, kk, ww);   // displacement on a bin edge
          }
        } else {
          if (r2 >= P.lo2 && r2 < P.hi) {
            const int b = pb_bin_log(r2, nbins, ed);
            double ww = 1.0;
            if constexpr (WEIGHTED) ww = wi * cw[jj];
            atomicAdd(my_c + b, 1u);
            atomicAdd(my_s + b, ki * ck[jj]);
            if constexpr (WEIGHTED) atomicAdd(my_w + b, ww);
            atomicAdd(my_r + b, ww * sqrt(r2));
          }
        }
      }
    }
    __syncwarp();
  };

  // Rank query of this lane's threshold T in the staged sorted copy of a chunk (cxy[].x = the coordinates along the
  // varying axis in ascending order, cxy[].y = suffix sums of k w, cw[] = suffix sums of w): pos = number of columns
  // with c_j < T, found by a 4-ary search -- 3 + 3 + 2 probes in three dependent steps instead of six dependent
  // probes (the kernel is latency-bound here).  Returns whether the exact mirrored bits agree for this lane: columns
  // below the split need c_j <= RT, columns from it on need c_j > RT, and in ascending order only the two
  // neighbours of the split can fail.
  auto rank_query = [&](double T, double RT, bool live, int& pos) -> bool {
    const int b = 8 * ((cxy[7].x < T ? 1 : 0) + (cxy[15].x < T ? 1 : 0) + (cxy[23].x < T ? 1 : 0));
    const int sp = b + 2 * ((cxy[b + 1].x < T ? 1 : 0) + (cxy[b + 3].x < T ? 1 : 0) + (cxy[b + 5].x < T ? 1 : 0));
    pos = sp + (cxy[sp].x < T ? 1 : 0) + (cxy[sp + 1].x < T ? 1 : 0);          // 0 .. 32
    return !live || ((pos == 0 || cxy[pos - 1].x <= RT) && (pos == PB_CHUNK || cxy[pos].x > RT));
  };

  // per-lane open bin of the blocks booked whole at classification time (forward bin cf_bin and its mirror image)
  int cf_bin = -1;
  unsigned cf_cnt = 0u;
  double cf_s = 0.0, cf_w = 0.0;
  auto cf_spill = [&]() {
    if (cf_bin >= 0 && cf_cnt) {
      atomicAdd(my_c + cf_bin, cf_cnt);    // forward entry; the mirror image is added by flush_hist
      atomicAdd(my_s + cf_bin, cf_s);
      if constexpr (WEIGHTED) atomicAdd(my_w + cf_bin, cf_w);
    }
    cf_cnt = 0u;
    cf_s = 0.0;
    cf_w = 0.0;
  };
  const double M = P.hi, lo2 = P.lo2;
  const int R = P.run;
  unsigned st_closed = 0, st_1d = 0, st_pw = 0, st_sorted = 0, st_quad = 0;   // column counts (x 32 rows = pairs) per path
  int since_flush = 0;
  while (true) {
    // ---- next work item for this warp ----
    unsigned long long q = 0;
    if (lane == 0) q = atomicAdd(P.counter, 1ull);
    q = __shfl_sync(0xffffffffu, q, 0);
    if ((int64_t)q >= P.my_items) break;
    const int64_t item = (int64_t)q * P.nranks + P.rank;
    const int cat = (int)(item / P.items_per_cat);
    if (cat >= P.ncat) break;
    const int64_t lq = item % P.items_per_cat;
    const int64_t off = P.cat_off[cat];
    const int64_t n = P.cat_off[cat + 1] - off;
    const int64_t nblk = (n + PB_CHUNK - 1) / PB_CHUNK;
    if (nblk == 0) continue;
    const int64_t nruns = (nblk + R - 1) / R;
    // decode lq -> (row block ib, run r): rows of group g = ib / R pair with runs g .. nruns-1;
    // items before group g: R * (g*nruns - g*(g-1)/2)
    const double tn = 2.0 * (double)nruns + 1.0;
    const double disc = tn * tn - 8.0 * (double)lq / (double)R;
    int64_t g = (int64_t)((tn - sqrt(disc > 0.0 ? disc : 0.0)) * 0.5);
    if (g < 0) g = 0;
    if (g >= nruns) g = nruns - 1;
    while (g > 0 && (int64_t)R * (g * nruns - g * (g - 1) / 2) > lq) --g;
    while (g + 1 < nruns && (int64_t)R * ((g + 1) * nruns - (g + 1) * g / 2) <= lq) ++g;
    const int64_t rem = lq - (int64_t)R * (g * nruns - g * (g - 1) / 2);
    const int64_t per_row = nruns - g;
    const int64_t row_in_g = rem / per_row;
    const int64_t ib = g * R + row_in_g;
    const int64_t r = g + rem % per_row;
    if (row_in_g >= R || ib >= nblk) continue;  // past the end of this (shorter) catalogue
    const int64_t c_lo = (r * R > ib) ? r * R : ib;
    const int64_t c_hi = ((r + 1) * R < nblk) ? (r + 1) * R : nblk;
    if (c_lo >= c_hi) continue;

    if (cat != cur_cat || since_flush >= PB_FLUSH_ITEMS) {
      flush_hist(cur_cat);
      cur_cat = cat;
      since_flush = 0;
    }
    ++since_flush;

    // ---- this warp's row points ----
    const int64_t ig = ib * PB_CHUNK + lane;
    const bool live = ig < n;
    // dead lanes: NaN coordinates make every comparison false, zero field/weight adds nothing
    const double xi = live ? P.px[off + ig] : __longlong_as_double(0x7ff8000000000000ll);
    const double yi = live ? P.py[off + ig] : __longlong_as_double(0x7ff8000000000000ll);
    double wi = live ? 1.0 : 0.0;
    if constexpr (WEIGHTED) wi = live ? P.pw[off + ig] : 0.0;
    const double ki = live ? P.pk[off + ig] * wi : 0.0;
    const double iminx = warp_min(live ? xi : INFINITY), imaxx = warp_max(live ? xi : -INFINITY);
    const double iminy = warp_min(live ? yi : INFINITY), imaxy = warp_max(live ? yi : -INFINITY);
    const long long owner = (long long)(off + ib * PB_CHUNK);
    // pre-pass record of chunk c of this catalogue: slot (off + 32 c) / 32 + cat = off / 32 + c + cat
    const double* rec0 = P.boxes ? P.boxes + (size_t)PB_STRIDE * (size_t)(off / PB_CHUNK + cat) : nullptr;
    // row-block aggregates for the blocks that are booked whole
    const double rowK = warp_sum(ki);
    double rowW = 0.0;
    if constexpr (WEIGHTED) rowW = warp_sum(wi);
    const int nlive = __popc(__ballot_sync(0xffffffffu, live));

    for (int64_t sc = c_lo; sc < c_hi; sc += 32) {
      // ---- lane l classifies column chunk sc + l ----
      const int64_t mychunk = sc + lane;
      int cls = PB_OUT;
      int kind = 0;   // 0: raw points; 1 / 2: one-axis block answered from the x- / y-sorted copy of the chunk
      // kind != 0 and the varying axis spans exactly two forward bins [v0, v0+1] whose mirrored window is the mirror
      // image (the usual case): 1 | kind << 1 | x0 << 3 | y0 << 15, everything the warp needs to see whether the
      // open window covers the block; 0 otherwise
      int fastw = 0;
      int win[4] = {0, 0, 0, 0};
      if (mychunk < c_hi) {
        const int64_t j0g = mychunk * PB_CHUNK;
        const int cnt = (int)((n - j0g < PB_CHUNK) ? (n - j0g) : PB_CHUNK);
        double cminx = INFINITY, cmaxx = -INFINITY, cminy = INFINITY, cmaxy = -INFINITY;
        if (P.boxes) {
          const double4 bb = *reinterpret_cast<const double4*>(rec0 + (size_t)PB_STRIDE * (size_t)mychunk);
          cminx = bb.x; cmaxx = bb.y; cminy = bb.z; cmaxy = bb.w;
        } else {
          const double* xs = P.px + off + j0g;
          const double* ys = P.py + off + j0g;
#pragma unroll 1
          for (int t = 0; t < PB_CHUNK; ++t) {   // only without a pre-pass (work == NULL): kept small
            if (t < cnt) {
              const double x = xs[t], y = ys[t];
              cminx = fmin(cminx, x); cmaxx = fmax(cmaxx, x);
              cminy = fmin(cminy, y); cmaxy = fmax(cmaxy, y);
            }
          }
        }
        if (BT == TGP_BIN_TWOD) {
          if (mychunk == ib) {
            cls = PB_GENERIC;  // diagonal block: needs j > i
          } else {
            cls = pb_classify_twod(iminx, imaxx, iminy, imaxy, cminx, cmaxx, cminy, cmaxy, M, lo2, nbins, ed, win);
          }
        } else {
          const double dx0 = cminx - imaxx, dx1 = cmaxx - iminx, dy0 = cminy - imaxy, dy1 = cmaxy - iminy;
          const double ax = fmax(fabs(dx0), fabs(dx1)), ay = fmax(fabs(dy0), fabs(dy1));
          const double nx = dx0 > 0.0 ? dx0 : (dx1 < 0.0 ? -dx1 : 0.0), ny = dy0 > 0.0 ? dy0 : (dy1 < 0.0 ? -dy1 : 0.0);
          const double r2max = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
          const double r2min = __dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny));
          cls = (r2max < lo2 || r2min >= M) ? PB_OUT : PB_GENERIC;  // M = max_sep^2 for Log
          if (BS && cls == PB_GENERIC && mychunk != ib && r2min >= lo2 && r2max < M) {
            // every pair in range; if the smallest and the largest r^2 of the block share a bin, all pairs do
            const int k0 = pb_bin_log(r2min, nbins, ed), k1 = pb_bin_log(r2max, nbins, ed);
            if (k0 == k1) { cls = PB_REG_FULL; win[0] = k0; }
            else if (k1 == k0 + 1) { cls = PB_REG_CHECK; win[0] = k0; }     // two adjacent bins: one compare per pair
          }
          if (mychunk == ib) cls = PB_GENERIC;
        }
        if (!(iminx <= imaxx)) cls = PB_OUT;  // no live row in this warp
        if (BS && BT == TGP_BIN_TWOD && cls == PB_REG_FULL && P.boxes) {
          // Whole block in ONE forward bin and in its mirror image: the classifying lane books the block itself
          // from the chunk sums of the pre-pass (count = rows x columns, sum = (sum of the rows' k w) x (sum of the
          // columns' k w)); the points of the chunk are never loaded.  Consecutive chunks of a lane mostly hit
          // the same bin: the lane keeps one open bin in registers and spills to the warp histogram on a change.
          const int x0 = win[0] & 0xffff, y0 = win[1] & 0xffff, rx0 = win[2] & 0xffff, ry0 = win[3] & 0xffff;
          const bool one_x = (win[0] >> 16) == 0 && (win[2] >> 16) == 0 && rx0 == nbins - 1 - x0;
          const bool one_y = (win[1] >> 16) == 0 && (win[3] >> 16) == 0 && ry0 == nbins - 1 - y0;
          // exactly one varying axis: the block is answered from the chunk's points sorted along that axis
          if (P.sorted && cnt == PB_CHUNK) kind = (one_y && !one_x) ? 1 : ((one_x && !one_y) ? 2 : 0);
          if (P.fast_paths & 1) {
            const bool clean_x = (win[0] >> 16) == 1 && (win[2] >> 16) == 1 && rx0 == nbins - 2 - x0;
            const bool clean_y = (win[1] >> 16) == 1 && (win[3] >> 16) == 1 && ry0 == nbins - 2 - y0;
            if ((kind == 1 && clean_x) || (kind == 2 && clean_y)) fastw = 1 | (kind << 1) | (x0 << 3) | (y0 << 15);
          }
          // two bins along BOTH axes: the raw points are staged as usual (kind stays 0); field value 3
          if ((P.fast_paths & 2) && P.sorted && cnt == PB_CHUNK && (win[0] >> 16) == 1 && (win[2] >> 16) == 1 &&
              rx0 == nbins - 2 - x0 && (win[1] >> 16) == 1 && (win[3] >> 16) == 1 && ry0 == nbins - 2 - y0)
            fastw = 1 | (3 << 1) | (x0 << 3) | (y0 << 15);
          if (one_x && one_y) {
            const double2 sums = *reinterpret_cast<const double2*>(rec0 + (size_t)PB_STRIDE * (size_t)mychunk + 4);
            const int o = y0 * nbins + x0;
            if (o != cf_bin) { cf_spill(); cf_bin = o; }
            cf_cnt += (unsigned)(nlive * cnt);
            cf_s = fma(rowK, sums.x, cf_s);
            if constexpr (WEIGHTED) cf_w = fma(rowW, sums.y, cf_w);
            st_closed += (unsigned)cnt;
            cls = PB_OUT;
          }
        }
      }
      const unsigned todo = __ballot_sync(0xffffffffu, cls != PB_OUT);
      // ---- process the non-OUT chunks; the next chunk's points are prefetched into registers ----
      unsigned rest = todo;
      double nx_ = 0.0, ny_ = 0.0, nk_ = 0.0, nw_ = 1.0;
      double2 nsum_ = make_double2(0.0, 0.0);
      int nkind_ = 0;
      auto fetch_raw = [&](int c) {
        const int64_t jg = (sc + c) * PB_CHUNK + lane;
        const bool ok = jg < n;
        nx_ = ok ? P.px[off + jg] : 0.0;
        ny_ = ok ? P.py[off + jg] : 0.0;
        if constexpr (WEIGHTED) nw_ = ok ? P.pw[off + jg] : 0.0;
        nk_ = ok ? P.pk[off + jg] * nw_ : 0.0;
      };
      auto fetch = [&](int c) {
        const double* rec = rec0 + (size_t)PB_STRIDE * (size_t)(sc + c);
        if (P.boxes) nsum_ = *reinterpret_cast<const double2*>(rec + 4);   // chunk sums of k w and w (pre-pass)
        if constexpr (!BS) {
          fetch_raw(c);
          return;
        }
        nkind_ = __shfl_sync(0xffffffffu, kind, c);
        if (nkind_ == 0) {
          fetch_raw(c);
        } else {
          // sorted copy: coordinate along the varying axis, suffix sums of k w (and w) in that order
          const double* sp = rec + PB_SLOT + (nkind_ == 1 ? 0 : 3 * PB_CHUNK) + lane;
          nx_ = sp[0];
          ny_ = sp[PB_CHUNK];
          if constexpr (WEIGHTED) nw_ = sp[2 * PB_CHUNK];
        }
      };
      if (rest) fetch(__ffs(rest) - 1);
      while (rest) {
        const int c = __ffs(rest) - 1;
        rest &= rest - 1;
        __syncwarp();  // previous chunk fully consumed
        cxy[lane] = make_double2(nx_, ny_);
        ck[lane] = nk_;
        if constexpr (WEIGHTED) cw[lane] = nw_;
        int ckind = nkind_;
        const double2 csum = nsum_;
        __syncwarp();
        if (rest) fetch(__ffs(rest) - 1);
        // a block staged from the sorted copy that ends up on a pair-by-pair path needs the raw points after all
        auto ensure_raw = [&]() {
          if (ckind == 0) return;
          __syncwarp();
          fetch_raw(c);
          cxy[lane] = make_double2(nx_, ny_);
          ck[lane] = nk_;
          if constexpr (WEIGHTED) cw[lane] = nw_;
          __syncwarp();
          if (rest) fetch(__ffs(rest) - 1);   // fetch_raw clobbered the prefetch registers
          ckind = 0;
        };
        if constexpr (BS && BT == TGP_BIN_TWOD) {
          // Short cut for the bulk of the one-axis blocks: a full chunk, every pair in range, two bins along the
          // varying axis, and the OPEN window already covers it (one shuffle and four integer compares instead of
          // the general dispatch below).  Anything else -- no open window, a window that has to move, a column
          // within rounding of a bin edge -- falls through to the general path, which decides again from scratch.
          const int fw = __shfl_sync(0xffffffffu, fastw, c);
          if (fw != 0 && A.fx0 >= 0 && A.ownerI == owner) {
            const bool vx = ((fw >> 1) & 3) == 1;
            const int bx0 = (fw >> 3) & 0xfff, by0 = (fw >> 15) & 0xfff;
            if (((fw >> 1) & 3) == 3) {
              // 2 x 2 window, every pair in range, and the open window is exactly the block's: the column-bit and
              // row-bit sums are rank queries on the x- / y-sorted copies (read from global memory: their lines
              // are pulled in while the pair loop runs); only the "both bits" quadrant is summed pair by pair.
              if (bx0 == A.fx0 && by0 == A.fy0) {
                const double* srt = rec0 + (size_t)PB_STRIDE * (size_t)(sc + c) + PB_SLOT;
                const double* srty = srt + 3 * PB_CHUNK;
                const double x7 = srt[7], x15 = srt[15], x23 = srt[23];
                const double y7 = srty[7], y15 = srty[15], y23 = srty[23];
                pb_prefetch_l1(srt + PB_CHUNK + 8); pb_prefetch_l1(srt + PB_CHUNK + 24);
                pb_prefetch_l1(srty + PB_CHUNK + 8); pb_prefetch_l1(srty + PB_CHUNK + 24);
                if constexpr (WEIGHTED) {
                  pb_prefetch_l1(srt + 2 * PB_CHUNK + 8); pb_prefetch_l1(srt + 2 * PB_CHUNK + 24);
                  pb_prefetch_l1(srty + 2 * PB_CHUNK + 8); pb_prefetch_l1(srty + 2 * PB_CHUNK + 24);
                }
                double bsq = 0.0, bwq = 0.0;
                unsigned bcq = 0u;
#pragma unroll 4
                for (int jj = 0; jj < PB_CHUNK; ++jj) {
                  const double2 pj = cxy[jj];
                  const double kj = ck[jj];
                  if constexpr (WEIGHTED) { const double wj = cw[jj]; PB_PAIR_Q_W(bsq, bwq, bcq, pj.x, pj.y, kj, wj, A.Tx, A.Ty); }
                  else PB_PAIR_Q(bsq, bcq, pj.x, pj.y, kj, A.Tx, A.Ty);
                }
                int posx, posy;
                const bool okx = pb_rank_query_g(srt, x7, x15, x23, A.Tx, A.RTx, live, posx);
                const bool oky = pb_rank_query_g(srty, y7, y15, y23, A.Ty, A.RTy, live, posy);
                if (__all_sync(0xffffffffu, okx && oky)) {
                  const unsigned n_add = live ? (unsigned)PB_CHUNK : 0u;
                  A.tot += csum.x;
                  A.nin += n_add;
                  A.fsx += (posx < PB_CHUNK) ? srt[PB_CHUNK + posx] : 0.0;
                  A.fsy += (posy < PB_CHUNK) ? srty[PB_CHUNK + posy] : 0.0;
                  A.fcx += live ? (unsigned)(PB_CHUNK - posx) : 0u;
                  A.fcy += live ? (unsigned)(PB_CHUNK - posy) : 0u;
                  A.fsxy += bsq;
                  A.fcxy += bcq;          // dead lanes: NaN thresholds, no compare is true
                  if constexpr (WEIGHTED) {
                    A.wtot += csum.y;
                    A.fwx += (posx < PB_CHUNK) ? srt[2 * PB_CHUNK + posx] : 0.0;
                    A.fwy += (posy < PB_CHUNK) ? srty[2 * PB_CHUNK + posy] : 0.0;
                    A.fwxy += bwq;
                  }
                  st_quad += (unsigned)PB_CHUNK;
                  continue;
                }
              }
            } else {
            const int dvar = vx ? bx0 - A.fx0 : by0 - A.fy0;      // varying axis: the window must start at its lower bin
            const int dcon = vx ? by0 - A.fy0 : bx0 - A.fx0;      // constant axis: either bin of the window
            if (dvar == 0 && (unsigned)dcon <= 1u) {
              int pos;
              const bool okl = rank_query(vx ? A.Tx : A.Ty, vx ? A.RTx : A.RTy, live, pos);
              if (__all_sync(0xffffffffu, okl)) {
                const unsigned n_add = live ? (unsigned)PB_CHUNK : 0u;
                const unsigned bc = live ? (unsigned)(PB_CHUNK - pos) : 0u;
                const double bs = (pos < PB_CHUNK) ? cxy[pos].y : 0.0;
                double bw_ = 0.0;
                if constexpr (WEIGHTED) bw_ = (pos < PB_CHUNK) ? cw[pos] : 0.0;
                const bool other = dcon != 0;
                A.tot += csum.x;
                A.nin += n_add;
                if constexpr (WEIGHTED) A.wtot += csum.y;
                if (vx) {
                  A.fsx += bs; A.fcx += bc;
                  if constexpr (WEIGHTED) A.fwx += bw_;
                  if (other) { A.fsy += csum.x; A.fcy += n_add; if constexpr (WEIGHTED) A.fwy += csum.y; }
                } else {
                  A.fsy += bs; A.fcy += bc;
                  if constexpr (WEIGHTED) A.fwy += bw_;
                  if (other) { A.fsx += csum.x; A.fcx += n_add; if constexpr (WEIGHTED) A.fwx += csum.y; }
                }
                if (other) { A.fsxy += bs; A.fcxy += bc; if constexpr (WEIGHTED) A.fwxy += bw_; }
                st_sorted += (unsigned)PB_CHUNK;
                continue;
              }
            }
            }
          }
        }
        const int64_t j0g = (sc + c) * PB_CHUNK;
        const int jcount = (int)((n - j0g < PB_CHUNK) ? (n - j0g) : PB_CHUNK);
        int ccls = __shfl_sync(0xffffffffu, cls, c);
        // Log bins and the diagonal block (needs j > i) go pair by pair through the generic path -- which has ONE
        // call site, at the top of the sub-block loop below, so that its code exists once in the kernel
        const bool whole_generic = (BT != TGP_BIN_TWOD && ccls == PB_GENERIC) || (sc + c) == ib;
        const int gfirst = ((sc + c) == ib) ? lane + 1 : 0;
        int cw4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) cw4[k] = __shfl_sync(0xffffffffu, win[k], c);
        if (BT == TGP_BIN_LOG && !whole_generic) {
          // Log bins, every pair of the block in range and in ONE radial bin (or in two adjacent ones): counts and
          // the k-weighted sums of the block are products of the row and chunk sums; sum w_i w_j r needs the pairs
          // (one square root each); with two bins one compare per pair splits off the upper bin's share.  No
          // per-pair bin search, no atomics.
          const int kb = cw4[0];
          const bool two = (ccls == PB_REG_CHECK);
          double accr = 0.0, hr = 0.0, hk = 0.0, hw = 0.0;
          unsigned hcn = 0u;
          if (!two) {
#pragma unroll 4
            for (int jj = 0; jj < jcount; ++jj) {
              const double2 pj = cxy[jj];
              const double dx = pj.x - xi, dy = pj.y - yi;
              const double r = sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
              if constexpr (WEIGHTED) accr = fma(cw[jj], r, accr);
              else accr += r;
            }
          } else {
            const double tsplit = ed[kb + 1];          // r2 >= tsplit <=> bin kb + 1
#pragma unroll 2
            for (int jj = 0; jj < jcount; ++jj) {
              const double2 pj = cxy[jj];
              const double dx = pj.x - xi, dy = pj.y - yi;
              const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
              double wr = sqrt(r2);
              if constexpr (WEIGHTED) wr *= cw[jj];
              accr += wr;
              const double m = (r2 >= tsplit) ? 1.0 : 0.0;
              hr = fma(wr, m, hr);
              hk = fma(ck[jj], m, hk);
              if constexpr (WEIGHTED) hw = fma(cw[jj], m, hw);
              hcn += (r2 >= tsplit) ? 1u : 0u;
            }
          }
          const double tot_r = warp_sum(live ? accr * wi : 0.0);
          double S, Sw = 0.0;
          if (P.boxes) {
            S = csum.x;
            Sw = csum.y;
          } else {
            S = warp_sum(ck[lane]);
            if constexpr (WEIGHTED) Sw = warp_sum(cw[lane]);
          }
          double up_r = 0.0, up_k = 0.0, up_w = 0.0;
          unsigned up_c = 0u;
          if (two) {
            up_r = warp_sum(live ? hr * wi : 0.0);
            up_k = warp_sum(live ? hk * ki : 0.0);
            if constexpr (WEIGHTED) up_w = warp_sum(live ? hw * wi : 0.0);
            up_c = warp_sum_u(live ? hcn : 0u);
          }
          if (lane == 0) {       // the histogram is private to this warp
            my_c[kb] += (unsigned)(nlive * jcount) - up_c;
            my_s[kb] += rowK * S - up_k;
            if constexpr (WEIGHTED) my_w[kb] += rowW * Sw - up_w;
            my_r[kb] += tot_r - up_r;
            if (two) {
              my_c[kb + 1] += up_c;
              my_s[kb + 1] += up_k;
              if constexpr (WEIGHTED) my_w[kb + 1] += up_w;
              my_r[kb + 1] += up_r;
              st_1d += (unsigned)jcount;
            } else {
              st_closed += (unsigned)jcount;
            }
          }
          __syncwarp();
          continue;
        }
        // a block whose window is too wide is retried as four 8-column sub-blocks
        const int nsub = (ccls == PB_GENERIC && !whole_generic) ? 4 : 1;
        int scls = ccls;
        int sw[4] = {cw4[0], cw4[1], cw4[2], cw4[3]};
        if (nsub == 4) {
          scls = PB_OUT;
          if (lane < 4 && lane * 8 < jcount) {
            double cminx = INFINITY, cmaxx = -INFINITY, cminy = INFINITY, cmaxy = -INFINITY;
#pragma unroll 1
            for (int t = lane * 8; t < min(lane * 8 + 8, jcount); ++t) {
              const double2 pj = cxy[t];
              cminx = fmin(cminx, pj.x); cmaxx = fmax(cmaxx, pj.x);
              cminy = fmin(cminy, pj.y); cmaxy = fmax(cmaxy, pj.y);
            }
            scls = pb_classify_twod(iminx, imaxx, iminy, imaxy, cminx, cmaxx, cminy, cmaxy, M, lo2, nbins, ed, sw);
          }
        }
        for (int sb = 0; sb < nsub; ++sb) {
          int bcls = scls, bw[4] = {sw[0], sw[1], sw[2], sw[3]};
          int j0 = 0, jn = jcount;
          if (nsub == 4) {
            bcls = __shfl_sync(0xffffffffu, scls, sb);
#pragma unroll
            for (int k = 0; k < 4; ++k) bw[k] = __shfl_sync(0xffffffffu, sw[k], sb);
            j0 = sb * 8;
            jn = min(8, jcount - j0);
            if (jn <= 0) bcls = PB_OUT;
          }
          if (bcls == PB_OUT) continue;
          bool gen = (bcls == PB_GENERIC);
          // ---- register path: make sure the open window covers this block ----
          const int x0 = bw[0] & 0xffff, x1 = x0 + (bw[0] >> 16), y0 = bw[1] & 0xffff, y1 = y0 + (bw[1] >> 16);
          const int rx0 = bw[2] & 0xffff, rx1 = rx0 + (bw[2] >> 16), ry0 = bw[3] & 0xffff, ry1 = ry0 + (bw[3] >> 16);
          // the mirrored window is the mirror image [n-2-f0, n-1-f0] of the forward window [f0, f0+1]
          auto covers = [&](int fx0, int fy0) {
            return fx0 >= 0 && fy0 >= 0 && x0 >= fx0 && x1 <= fx0 + 1 && y0 >= fy0 && y1 <= fy0 + 1 &&
                   rx0 >= nbins - 2 - fx0 && rx1 <= nbins - 1 - fx0 && ry0 >= nbins - 2 - fy0 && ry1 <= nbins - 1 - fy0;
          };
          const bool fits = A.fx0 >= 0 && A.ownerI == owner && covers(A.fx0, A.fy0);
          if (!fits && !gen) {
            flush_regs();
            // candidate origins: the block's lowest bin, or one below it (a window must stay inside the grid
            // unless nbins == 1)
            const int hi0 = max(nbins - 2, 0);
            int fx0 = min(x0, hi0), fy0 = min(y0, hi0);
            if (!covers(fx0, fy0)) {
              const int ax = (x1 == x0 && x0 >= 1) ? x0 - 1 : fx0, ay = (y1 == y0 && y0 >= 1) ? y0 - 1 : fy0;
              if (covers(ax, fy0)) fx0 = ax;
              else if (covers(fx0, ay)) fy0 = ay;
              else if (covers(ax, ay)) { fx0 = ax; fy0 = ay; }
              else gen = true;   // edge asymmetry: exact path
            }
            if (!gen) {
            A.fx0 = fx0; A.fy0 = fy0;
            A.ownerI = owner;
            A.rki = ki;
            A.rwi = wi;
            // bit = (bin >= b0 + 1): forward dx >= ed[b0+1]; mirrored -dx >= ed[r0+1] <=> dx <= -ed[r0+1] with
            // r0 = nbins-2-b0; turned into thresholds on the column coordinate for this lane's row point
            const double tx = (nbins > 1) ? ed[fx0 + 1] : INFINITY, ty = (nbins > 1) ? ed[fy0 + 1] : INFINITY;
            const double ntx = -ed[nbins - 1 - fx0], nty = -ed[nbins - 1 - fy0];   // ed[0] = -inf when nbins == 1
            bool okl = true;
            A.Tx = pb_coord_ge(xi, tx, okl);
            A.Ty = pb_coord_ge(yi, ty, okl);
            A.RTx = pb_coord_le(xi, ntx, okl);
            A.RTy = pb_coord_le(yi, nty, okl);
            if (!__all_sync(0xffffffffu, okl)) {
              // (never seen in practice) the per-lane thresholds did not settle: generic path for this block
              A.fx0 = -1;
              gen = true;
            }
            }
          }
          if (gen) {
            ensure_raw();
            generic_block(j0, jn, gfirst, xi, yi, ki, wi, live);
            if (!whole_generic) st_pw += (unsigned)jn;
            continue;
          }
          if (bcls == PB_REG_FULL) {
            // Every pair of the block is in range.  How many bins do its displacements span?  one_x: all dx in ONE
            // forward bin, all -dx in ONE mirrored bin, and that bin is the mirror image (the usual case: bins are
            // much wider than a 32-point chunk of a sorted catalogue).  Such an axis needs no per-pair decision:
            // the window bit is the same for the whole block.
            const bool whole = BS && (nsub == 1);
            const bool one_x = whole && x1 == x0 && rx1 == rx0 && rx0 == nbins - 1 - x0;
            const bool one_y = whole && y1 == y0 && ry1 == ry0 && ry0 == nbins - 1 - y0;
            // Pair-by-pair kernel: when the bounding boxes put the whole block into ONE forward bin per axis and its
            // mirrored window is that bin's mirror image, the mirrored bits of every pair are the complement of the
            // forward bits by construction; the pair loop then evaluates the forward bits only (every pair is still
            // binned individually).
            const bool nm = !BS && (P.fast_paths & 4) && x1 == x0 && rx1 == rx0 && rx0 == nbins - 1 - x0 &&
                            y1 == y0 && ry1 == ry0 && ry0 == nbins - 1 - y0;
            const unsigned n_add = live ? (unsigned)jn : 0u;
            if (one_x || one_y) {
              // chunk sums of the column values (dead columns were staged as zero)
              double S, Sw = 0.0;
              if (P.boxes) {
                S = csum.x;
                Sw = csum.y;
              } else {
                S = warp_sum(ck[lane]);
                if constexpr (WEIGHTED) Sw = warp_sum(cw[lane]);
              }
              const bool bx = (x0 - A.fx0) != 0, by = (y0 - A.fy0) != 0;   // window bits of the constant axes
              double bs = 0.0, bw_ = 0.0;     // block-local sums / count of the pairs in the upper bin of the varying axis
              unsigned bc = 0u;
              bool done = false;
              if (ckind != 0) {
                // Staged: the chunk's coordinates along the varying axis in ascending order (cxy[].x) with the
                // suffix sums of k w (cxy[].y) and w (cw[]).  The lane's bit "c_j >= T" is a rank query: lower
                // bound by bisection (6 probes instead of 32 compares), count and sum read off the suffix arrays.
                const double T = (ckind == 1) ? A.Tx : A.Ty, RT = (ckind == 1) ? A.RTx : A.RTy;
                int pos;
                const bool okl = rank_query(T, RT, live, pos);
                if (__all_sync(0xffffffffu, okl)) {
                  bc = live ? (unsigned)(PB_CHUNK - pos) : 0u;
                  bs = (pos < PB_CHUNK) ? cxy[pos].y : 0.0;
                  if constexpr (WEIGHTED) bw_ = (pos < PB_CHUNK) ? cw[pos] : 0.0;
                  done = true;
                } else {
                  // (rare) a column within rounding of a bin edge: redo the block pair by pair from the raw points
                  ensure_raw();
                }
              }
              const bool vary_x = one_y;                 // (one_x && one_y): treated as "x varies" with a constant bit
              if (done) {
              } else if (one_x && one_y) {
                // all 32 x jn pairs in one bin (and its mirror image): closed form, no per-pair work at all
                bs = bx ? S : 0.0;
                bw_ = bx ? Sw : 0.0;
                bc = bx ? n_add : 0u;
              } else if (one_y) {
#pragma unroll 2
                for (int jj = 0; jj < jn; ++jj) {
                  const double cj = cxy[jj].x, kj = ck[jj];
                  if constexpr (WEIGHTED) { const double wj = cw[jj]; PB_PAIR_1D_W(bs, bw_, bc, A.mmc, cj, kj, wj, A.Tx, A.RTx); }
                  else PB_PAIR_1D(bs, bc, A.mmc, cj, kj, A.Tx, A.RTx);
                }
              } else {
#pragma unroll 2
                for (int jj = 0; jj < jn; ++jj) {
                  const double cj = cxy[jj].y, kj = ck[jj];
                  if constexpr (WEIGHTED) { const double wj = cw[jj]; PB_PAIR_1D_W(bs, bw_, bc, A.mmc, cj, kj, wj, A.Ty, A.RTy); }
                  else PB_PAIR_1D(bs, bc, A.mmc, cj, kj, A.Ty, A.RTy);
                }
              }
              // fold into the window registers.  v = the varying axis (x unless only x is constant):
              // sum[v bit] += bs; sum[other bit] += S if that bit is set; sum[both] += bs if the other bit is set.
              const bool other = vary_x ? by : bx;
              A.tot += S;
              if constexpr (WEIGHTED) A.wtot += Sw;
              A.nin += n_add;
              if (vary_x) {
                A.fsx += bs; A.fcx += bc;
                if constexpr (WEIGHTED) A.fwx += bw_;
                if (other) { A.fsy += S; A.fcy += n_add; if constexpr (WEIGHTED) A.fwy += Sw; }
              } else {
                A.fsy += bs; A.fcy += bc;
                if constexpr (WEIGHTED) A.fwy += bw_;
                if (other) { A.fsx += S; A.fcx += n_add; if constexpr (WEIGHTED) A.fwx += Sw; }
              }
              if (other) { A.fsxy += bs; A.fcxy += bc; if constexpr (WEIGHTED) A.fwxy += bw_; }
              if (!live) A.mmc = 0u;
              if (one_x && one_y) { if (lane == 0) st_closed += (unsigned)jn; }
              else if (done) st_sorted += (unsigned)jn;
              else st_1d += (unsigned)jn;
              if ((one_x && one_y) || done) continue;   // nothing can be inconsistent
            } else {
            // all pairs of the block are in range: the unmasked sums (sum of k w, of w) are the chunk sums of the
            // pre-pass when the whole chunk is processed, so only the masked sums are accumulated per pair
            const bool tot_from_sums = P.boxes && nsub == 1;
            if (!BS && tot_from_sums && nm) {
#pragma unroll 4
              for (int jj = j0; jj < j0 + jn; ++jj) {
                const double2 pj = cxy[jj];
                const double kj = ck[jj];
                if constexpr (WEIGHTED) {
                  const double wj = cw[jj];
                  PB_PAIR_NM_W_NT(A, pj.x, pj.y, kj, wj);
                } else {
                  PB_PAIR_NM_NT(A, pj.x, pj.y, kj);
                }
              }
              A.tot += csum.x;
              if constexpr (WEIGHTED) A.wtot += csum.y;
            } else if (tot_from_sums) {
#pragma unroll 4
              for (int jj = j0; jj < j0 + jn; ++jj) {
                const double2 pj = cxy[jj];
                const double kj = ck[jj];
                if constexpr (WEIGHTED) {
                  const double wj = cw[jj];
                  PB_PAIR_W_NT(A, pj.x, pj.y, kj, wj);
                } else {
                  PB_PAIR_NT(A, pj.x, pj.y, kj);
                }
              }
              A.tot += csum.x;
              if constexpr (WEIGHTED) A.wtot += csum.y;
            } else {
#pragma unroll 4
              for (int jj = j0; jj < j0 + jn; ++jj) {
                const double2 pj = cxy[jj];
                const double kj = ck[jj];
                if constexpr (WEIGHTED) {
                  const double wj = cw[jj];
                  PB_PAIR_W(A, pj.x, pj.y, kj, wj);
                } else {
                  PB_PAIR(A, pj.x, pj.y, kj);
                }
              }
            }
            A.nin += n_add;
            if (!live) A.mmc = 0u;  // dead lanes (NaN coordinates) compare false everywhere: not a mismatch
            st_pw += (unsigned)jn;
            }
          } else {
            st_pw += (unsigned)jn;
#pragma unroll 2
            for (int jj = j0; jj < j0 + jn; ++jj) {
              const double2 pj = cxy[jj];
              const double dx = pj.x - xi, dy = pj.y - yi;
              const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
              const bool ok = r2 >= lo2 && fabs(dx) < M && fabs(dy) < M;  // false for dead lanes (NaN)
              if (ok) {
                const double kj = ck[jj];
                if constexpr (WEIGHTED) {
                  const double wj = cw[jj];
                  PB_PAIR_W(A, pj.x, pj.y, kj, wj);
                } else {
                  PB_PAIR(A, pj.x, pj.y, kj);
                }
                A.nin += 1u;
              }
            }
          }
          fix_mirror(j0, jn, bcls != PB_REG_FULL, xi, yi, ki, wi);
        }
      }
    }
    // bound the per-lane 32-bit counters: registers go to shared memory at the end of every item
    cf_spill();
    cf_bin = -1;
    __syncwarp();
    flush_regs();
  }
  flush_hist(cur_cat);
  {  // blocks booked at classification time are tallied by the classifying lane
    if (st_closed) atomicAdd(&g_pb_stats[0], 32ull * st_closed);
  }
  if (lane == 0) {
    if (st_1d) atomicAdd(&g_pb_stats[1], 32ull * st_1d);
    if (st_pw) atomicAdd(&g_pb_stats[2], 32ull * st_pw);
    if (st_sorted) atomicAdd(&g_pb_stats[3], 32ull * st_sorted);
    if (st_quad) atomicAdd(&g_pb_stats[4], 32ull * st_quad);
  }
}

// Pre-pass: bounding box and sums of every 32-point chunk of every catalogue (one warp per chunk).
// Slot of chunk c of catalogue `cat`: (cat_off[cat] + 32 c) / 32 + cat  (distinct and monotone); PB_SLOT doubles
// per slot: {xmin, xmax, ymin, ymax, sum k w, sum w, -, -}.
__global__ void __launch_bounds__(256)
pairbin_boxes_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pk,
                     const double* __restrict__ pw, const int64_t* __restrict__ cat_off,
                     int32_t ncat, int64_t chunks_per_cat, double* __restrict__ boxes, double* __restrict__ sorted) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t cat = w / chunks_per_cat, c = w % chunks_per_cat;
  if (cat >= ncat) return;
  const int64_t off = cat_off[cat], n = cat_off[cat + 1] - off;
  const int64_t j = c * PB_CHUNK + lane;
  if (c * PB_CHUNK >= n) return;
  const bool ok = j < n;
  const double x = ok ? px[off + j] : 0.0, y = ok ? py[off + j] : 0.0;
  const double wj = ok ? (pw ? pw[off + j] : 1.0) : 0.0;
  const double kj = ok ? pk[off + j] * wj : 0.0;
  const double a = warp_min(ok ? x : INFINITY), b = warp_max(ok ? x : -INFINITY);
  const double cc = warp_min(ok ? y : INFINITY), d = warp_max(ok ? y : -INFINITY);
  const double sk = warp_sum(kj), sw = warp_sum(wj);
  const int64_t slot = (off + c * PB_CHUNK) / PB_CHUNK + cat;
  if (lane == 0) {
    double* o = boxes + PB_STRIDE * slot;
    o[0] = a; o[1] = b; o[2] = cc; o[3] = d;
    o[4] = sk; o[5] = sw; o[6] = 0.0; o[7] = 0.0;
  }
  if (sorted) {
    // the chunk sorted by x and by y (bitonic network over the 32 lanes; absent points sort last as +inf with
    // zero payload), each with the suffix sums of k w and w in that order
#pragma unroll
    for (int axis = 0; axis < 2; ++axis) {
      double key = ok ? (axis == 0 ? x : y) : INFINITY, vk = kj, vw = wj;
#pragma unroll
      for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
        for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
          const double okey = __shfl_xor_sync(0xffffffffu, key, j2);
          const double ovk = __shfl_xor_sync(0xffffffffu, vk, j2);
          const double ovw = __shfl_xor_sync(0xffffffffu, vw, j2);
          const bool take_min = ((lane & j2) == 0) == ((lane & k2) == 0);
          const bool swap = take_min ? (okey < key) : (okey > key);
          if (swap) { key = okey; vk = ovk; vw = ovw; }
        }
      }
#pragma unroll
      for (int o2 = 1; o2 < 32; o2 <<= 1) {
        const double tk = __shfl_down_sync(0xffffffffu, vk, o2), tw = __shfl_down_sync(0xffffffffu, vw, o2);
        if (lane + o2 < 32) { vk += tk; vw += tw; }
      }
      double* o = sorted + PB_STRIDE * slot + axis * 3 * PB_CHUNK + lane;
      o[0] = key;
      o[PB_CHUNK] = vk;
      o[2 * PB_CHUNK] = vw;
    }
  }
}

extern "C" int tgp_pairbin_tile(void) { return PB_CHUNK; }

extern "C" int64_t tgp_pairbin_work_doubles(int64_t total_points, int32_t ncat) {
  return (int64_t)PB_STRIDE * (total_points / PB_CHUNK + (int64_t)ncat + 2);
}

static int g_pb_fast_paths = 7;   // tgp_set_option("pairbin_fast_paths", bits): see PBParams::fast_paths
extern "C" int tgp_pairbin_set_fast_paths(int bits) { g_pb_fast_paths = bits; return TGP_OK; }
static int g_pb_block_sums = 1;   // tgp_set_option("pairbin_block_sums", 0): every pair evaluated individually
extern "C" int tgp_pairbin_set_block_sums(int on) { g_pb_block_sums = on ? 1 : 0; return TGP_OK; }

extern "C" int tgp_pairbin_stats(unsigned long long* host8, int reset) {
  TGP_CHECK_ARG(host8 != nullptr, "host8");
  TGP_CUDA(cudaMemcpyFromSymbol(host8, g_pb_stats, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    TGP_CUDA(cudaMemcpyToSymbol(g_pb_stats, z, sizeof(z)));
  }
  return TGP_OK;
}

// ring of work counters so that launches on different streams do not share one
constexpr int PB_COUNTER_SLOTS = 64;
__device__ unsigned long long g_pb_counters[PB_COUNTER_SLOTS];

extern "C" int tgp_pairbin(const double* px, const double* py, const double* pk, const double* pw,
                           const int64_t* cat_off, int32_t ncat, int64_t max_cat_len, int32_t bin_type,
                           const double* edges, int32_t nbins, double min_sep2, double max_sep,
                           int32_t tile_rank, int32_t tile_nranks, int64_t* npairs, double* sumw,
                           double* sumwkk, double* sumwr, double* work, void* stream) {
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
  P.block_sums = g_pb_block_sums;
  P.fast_paths = g_pb_fast_paths;
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

  // shared memory: every warp owns a private histogram; the budget decides the warps per CTA
  const size_t per_warp = (size_t)P.nb * (8 + 4 + (weighted ? 8 : 0) + (twod ? 0 : 8)) + PB_CHUNK * (16 + 8 + 8) + 16;
  const size_t fixed = (size_t)((nbins + 2) & ~1) * 8;
  const size_t budget = 200 * 1024;
  TGP_CHECK_ARG(fixed + per_warp <= budget, "too many bins for the shared-memory histogram");
  int warps = (int)((100 * 1024 - fixed) / per_warp);  // aim at two CTAs per SM
  if (warps < 1) warps = (int)((budget - fixed) / per_warp);
  if (warps > PB_MAX_WARPS) warps = PB_MAX_WARPS;
  if (warps < 1) warps = 1;
  P.warps = warps;
  const size_t smem = fixed + (size_t)warps * per_warp + 64;
  const int sms = tgp_num_sms();
  const int ctas_per_sm = (2 * smem <= 220 * 1024) ? 2 : 1;
  const int64_t grid_target = (int64_t)sms * ctas_per_sm;

  // work decomposition: item = 32 row points x a run of `run` column chunks
  const int64_t nblk = tgp_cdiv(max_cat_len, PB_CHUNK);
  const double chunk_pairs = 0.5 * (double)nblk * (double)nblk * (double)ncat;
  const double slots = (double)grid_target * warps * tile_nranks;
  int64_t run = (int64_t)(chunk_pairs / (slots * 32.0));  // >= ~32 items per warp slot
  run = (run / 32) * 32;
  if (run < 32) run = 32;
  if (run > 256) run = 256;
  P.run = (int32_t)run;
  const int64_t nruns = tgp_cdiv(nblk, run);
  // items of one catalogue: groups g of `run` rows pair with runs g..nruns-1
  int64_t items = 0;
  for (int64_t g = 0; g < nruns; ++g) {
    const int64_t rows = (nblk - g * run < run) ? (nblk - g * run) : run;
    items += rows * (nruns - g);
  }
  // the decode assumes full groups; the upper bound below covers the (shorter) last group too
  P.items_per_cat = run * (nruns * nruns - nruns * (nruns - 1) / 2);
  (void)items;
  const int64_t total_items = P.items_per_cat * ncat;
  P.my_items = (total_items - tile_rank + tile_nranks - 1) / tile_nranks;
  if (P.my_items <= 0) return TGP_OK;
  int64_t grid = tgp_cdiv(P.my_items, warps);
  if (grid > grid_target) grid = grid_target;

  P.boxes = nullptr;
  P.sorted = nullptr;
  if (work) {
    TGP_CHECK_ARG(((uintptr_t)work % 32) == 0, "work must be 32-byte aligned");
    const int64_t warps_needed = nblk * ncat;
    pairbin_boxes_kernel<<<(unsigned)tgp_cdiv(warps_needed * 32, 256), 256, 0, st>>>(px, py, pk, pw, cat_off, ncat, nblk,
                                                                                       work, work + PB_SLOT);
    TGP_LAUNCH_CHECK();
    P.boxes = work;
    P.sorted = work + PB_SLOT;
  }

  static std::atomic<unsigned> launch_seq{0};   // host threads launching on different streams get different slots
  static void* cbase_dev[TGP_MAX_DEVICES] = {};
  void*& cbase = cbase_dev[tgp_current_device()];
  if (!cbase) TGP_CUDA(cudaGetSymbolAddress(&cbase, g_pb_counters));
  P.counter = reinterpret_cast<unsigned long long*>(cbase) + (launch_seq.fetch_add(1u) % PB_COUNTER_SLOTS);
  TGP_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(unsigned long long), st));

#define TGP_PB_LAUNCH(BT, W, BS)                                                                          \
  do {                                                                                                    \
    static TgpPerDeviceOnce once_;                                                                        \
    if (tgp_first_use_on_device(once_))                                                                   \
      TGP_CUDA(cudaFuncSetAttribute(pairbin_kernel<BT, W, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)budget));                                                        \
    pairbin_kernel<BT, W, BS><<<(unsigned)grid, warps * 32, smem, st>>>(P);                               \
  } while (0)
  if (twod && P.block_sums) {
    if (weighted) TGP_PB_LAUNCH(TGP_BIN_TWOD, true, true); else TGP_PB_LAUNCH(TGP_BIN_TWOD, false, true);
  } else if (twod) {
    if (weighted) TGP_PB_LAUNCH(TGP_BIN_TWOD, true, false); else TGP_PB_LAUNCH(TGP_BIN_TWOD, false, false);
  } else {
    if (P.block_sums) {
      if (weighted) TGP_PB_LAUNCH(TGP_BIN_LOG, true, true); else TGP_PB_LAUNCH(TGP_BIN_LOG, false, true);
    } else {
      if (weighted) TGP_PB_LAUNCH(TGP_BIN_LOG, true, false); else TGP_PB_LAUNCH(TGP_BIN_LOG, false, false);
    }
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

// ---- the same keys with the bounding square found on the device: no host round trip in the middle of a call ----
__device__ __forceinline__ unsigned long long hb_key(double v) {   // order-preserving map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double hb_unkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}
__global__ void hilbert_bounds_init_kernel(unsigned long long* b4) {
  if (threadIdx.x < 4) b4[threadIdx.x] = (threadIdx.x & 1) ? 0ull : ~0ull;   // {min x, max x, min y, max y}
}
__global__ void __launch_bounds__(256)
hilbert_bounds_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, unsigned long long* b4) {
  double xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double a = x[i], b = y[i];
    xmin = fmin(xmin, a); xmax = fmax(xmax, a);
    ymin = fmin(ymin, b); ymax = fmax(ymax, b);
  }
  xmin = warp_min(xmin); xmax = warp_max(xmax);
  ymin = warp_min(ymin); ymax = warp_max(ymax);
  if ((threadIdx.x & 31) == 0 && xmin <= xmax) {
    atomicMin(b4 + 0, hb_key(xmin)); atomicMax(b4 + 1, hb_key(xmax));
    atomicMin(b4 + 2, hb_key(ymin)); atomicMax(b4 + 3, hb_key(ymax));
  }
}
__global__ void hilbert_keys_auto_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n,
                                         const unsigned long long* __restrict__ b4, int order,
                                         int64_t* __restrict__ keys) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x0 = hb_unkey(b4[0]), x1 = hb_unkey(b4[1]), y0 = hb_unkey(b4[2]), y1 = hb_unkey(b4[3]);
  // the formula of treegp_b200.backend.hilbert_order (same IEEE operations, so the same keys as tgp_hilbert_keys)
  double extent = fmax(__dsub_rn(x1, x0), __dsub_rn(y1, y0));
  extent = extent > 0.0 ? __dmul_rn(extent, 1.0 + 1e-9) : 1.0;
  const double inv_cell = __ddiv_rn((double)(1u << order), extent);
  const unsigned side = 1u << order;
  long long cx = (long long)floor(__dmul_rn(__dsub_rn(x[i], x0), inv_cell));
  long long cy = (long long)floor(__dmul_rn(__dsub_rn(y[i], y0), inv_cell));
  unsigned ux = (unsigned)(cx < 0 ? 0 : (cx >= (long long)side ? side - 1 : cx));
  unsigned uy = (unsigned)(cy < 0 ? 0 : (cy >= (long long)side ? side - 1 : cy));
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

extern "C" int tgp_hilbert_keys_auto(const double* x, const double* y, int64_t n, int32_t order, void* scratch32,
                                     int64_t* keys, void* stream) {
  TGP_CHECK_ARG(n >= 0 && order >= 1 && order <= 30, "n/order");
  if (n == 0) return TGP_OK;
  TGP_CHECK_ARG(x && y && keys && scratch32 && ((uintptr_t)scratch32 % 8) == 0, "null / unaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long* b4 = reinterpret_cast<unsigned long long*>(scratch32);
  hilbert_bounds_init_kernel<<<1, 32, 0, st>>>(b4);
  int64_t blocks = tgp_cdiv(n, 256 * 8);
  if (blocks > 4 * tgp_num_sms()) blocks = 4 * tgp_num_sms();
  hilbert_bounds_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, n, b4);
  hilbert_keys_auto_kernel<<<(unsigned)tgp_cdiv(n, 256), 256, 0, st>>>(x, y, n, b4, order, keys);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
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
