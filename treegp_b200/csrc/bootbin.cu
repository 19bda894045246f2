// Bootstrap batch of the anisotropic (TwoD) two-point correlation function with the pair GEOMETRY SHARED by all
// resamples.
//
// Replaces the loop of /root/reference/treegp/two_pcf.py:342-362 (comp_xi_covariance: n_bootstrap times
// resample_bootstrap :269-281 + comp_2pcf :283-328, i.e. one TreeCorr KKCorrelation.process per resample).  A
// resample is the base catalogue with integer multiplicities m_b[i] (how often point i was drawn); duplicates sit at
// distance zero and are skipped by TreeCorr, so with a_b[i] = m_b[i] w_i (w_i = 1 / y_err_i^2, or 1)
//     sumw_b[bin]   = sum over pairs (i < j) of the BASE catalogue in `bin` of  a_b[i] a_b[j]
//     sumwkk_b[bin] = ... of  a_b[i] a_b[j] (y_i - ybar_b)(y_j - ybar_b),          xi_b = sumwkk_b / sumw_b.
// Which bin a pair falls into does not depend on b.  With z = y - ybar0 (any fixed centring), delta_b = mean_b(z) and
// c_b[i] = a_b[i] z_i the three sums
//     S0 = sum a_i a_j,   S1 = sum (c_i a_j + a_i c_j),   S2 = sum c_i c_j          (per bin and resample)
// give sumw = S0 and sumwkk = S2 - delta_b S1 + delta_b^2 S0, so the mean of the resample is applied AFTER the
// reduction.  tgp_pairbin run on B independent weighted catalogues (the round-1 batch) repeats the whole geometry
// -- bounding-box classification, per-pair compares, rank queries -- B times; here it is done once per 32 resamples:
//
//   * a warp owns 32 row points of the base catalogue (Hilbert order) and ONE group of 32 resamples; in the
//     geometry phases its lanes are points (rows or columns), in the accumulation phases its lanes are RESAMPLES;
//   * (row block, column chunk) pairs are classified from bounding boxes exactly like tgp_pairbin (same thresholds,
//     same exactness argument: FP subtraction is monotone);
//   * a block that falls into ONE bin is booked from the per-resample chunk sums of the pre-pass (4 FMAs per lane);
//   * a block with two bins along an axis takes that axis' "upper bin" sums from a sweep: a column (coordinate s,
//     columns in ascending order along the axis) pairs in the upper bin with the rows whose coordinate is small
//     enough, fl(s - r) >= t, i.e. with the first nle rows in ascending order; nle is found by bisection in the
//     column's lane, and the column's share is (a, c) times the prefix sums of the rows' (a, c) in ascending order --
//     tables built once per work item in shared memory (lanes = resamples), so a sweep is 32 independent steps of
//     6 FMAs instead of 1024 pair steps; the exact mirrored bits are checked at the two neighbours of the split;
//   * what is left (the "both bits" quadrant of 2 x 2-window blocks, blocks that need the range test per pair, the
//     diagonal block) goes pair by pair: per-row 32-bit masks from ballots in the geometry phase, then per resample
//     lane masked FMAs over the chunk's 32 columns held in registers (the 0.0 / 1.0 factors of four columns come
//     from a 16-entry table indexed by a nibble of the mask; bytes of the mask without a bit are skipped: columns
//     are kept in ascending x, so the zeros of a quadrant mask are contiguous);
//   * multiplicities travel as bytes (three copies: row-major for rows, chunk-packed in ascending-x and ascending-y
//     column order) and become doubles by the 2^52 trick inside one FMA: a = fma(2^52 + m, w, -2^52 w) = m w exactly.
//
// Window registers hold the sums in inclusion-exclusion form {all, x-bit, y-bit, both} for an open 2 x 2 bin window
// and are flushed with red.global.add.f64 into a FORWARD-ONLY histogram [3][bins][resamples]; the mirrored entry
// (-dx, -dy) of every pair is the mirror image bin except for displacements within rounding of a bin edge -- blocks
// where the exact mirrored bits disagree take a per-pair path that books the correction into a second histogram.
// Parity: tests/test_gpu_bootbin.py compares with tgp_pairbin on the B compacted weighted catalogues (itself pinned
// to the oracle) and with the sequential oracle.
#include <float.h>
#include <math.h>
#include <atomic>
#include "tgp_common.cuh"

constexpr int BB_CHUNK = 32;
constexpr int BB_WARPS = 4;   // warps per CTA
constexpr int BB_GEO = 80;   // doubles per chunk record: box[4], xs[32], ys[32], permx[32] u8, permy[32] u8, pad

struct BBParams {
  const double *px, *py;
  const double4* pt;          // per point {w, -2^52 w, w z, -2^52 w z}, padded to whole chunks with zeros
  const double* geo;          // per chunk record (BB_GEO doubles)
  const uint8_t *m_row;       // [nblk*32][bpad]
  const uint8_t *m_sx, *m_sy;   // [nblk][bpad][32]: column multiplicities in ascending-x / ascending-y column order
  const double* csum;         // [nblk][3][bpad]: per chunk and resample sum a, sum c, sum m z
  const double* edges;
  double* hist;               // [2][3][nb][bpad]: forward-only sums, then corrections
  unsigned long long* counter;
  double lo2, hi, inv_bin;
  int64_t n, nblk, items_per_group, my_items;
  int32_t nbins, nb, bpad, ngroups, run, rank, nranks, paths;
};

// diagnostics: [0] blocks booked whole, [1] sweeps, [2] pair-by-pair blocks, [3] exact per-pair (slow) blocks,
// [4] window flushes, [5] sweeps handed back (mirrored split differs), [6] wide-window blocks summed bin by bin
__device__ unsigned long long g_bb_stats[8];

enum { BB_OUT = 0, BB_REG_FULL = 1, BB_REG_CHECK = 2, BB_GENERIC = 3 };

__device__ __forceinline__ int bb_bin_twod(double d, double hi, double inv_bin, int nbins, const double* __restrict__ ed) {
  int i = (int)((d + hi) * inv_bin);
  i = max(0, min(i, nbins - 1));
  if (d < ed[i]) --i;
  else if (d >= ed[i + 1]) ++i;
  return i;
}
__device__ __forceinline__ int bb_bin_search(double d, int nbins, const double* __restrict__ ed) {
  int lo = 0, hi = nbins - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (d >= ed[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}
__device__ __forceinline__ double bb_warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double bb_warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// m as a double with the exponent of 2^52: fma(bb_magic(m), w, -2^52 w) = m w, correctly rounded
__device__ __forceinline__ double bb_magic(unsigned m) { return __hiloint2double(0x43300000, (int)m); }

// Classification of one (row block, column chunk) pair from the two bounding boxes (the rules of tgp_pairbin):
// returns 0 for a block with no pair in range, else  cls | ex << 2 | ey << 3 | x0 << 8 | y0 << 20  where
// [x0, x0+ex] x [y0, y0+ey] is the forward bin window of all displacements of the block.  REG classes guarantee
// ex, ey <= 1 and that the mirrored window is the mirror image of the forward one.
__device__ __forceinline__ int bb_classify(double iminx, double imaxx, double iminy, double imaxy, double cminx,
                                           double cmaxx, double cminy, double cmaxy, double M, double lo2, int nbins,
                                           const double* __restrict__ ed) {
  const double dx0 = cminx - imaxx, dx1 = cmaxx - iminx;
  const double dy0 = cminy - imaxy, dy1 = cmaxy - iminy;
  const double ax = fmax(fabs(dx0), fabs(dx1)), ay = fmax(fabs(dy0), fabs(dy1));
  const double nx = dx0 > 0.0 ? dx0 : (dx1 < 0.0 ? -dx1 : 0.0);
  const double ny = dy0 > 0.0 ? dy0 : (dy1 < 0.0 ? -dy1 : 0.0);
  const double r2max = __dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay));
  const double r2min = __dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny));
  if (!(nx < M && ny < M && r2max >= lo2)) return BB_OUT;   // also rejects NaN boxes (empty chunks)
  const int x0 = bb_bin_search(dx0, nbins, ed), x1 = bb_bin_search(dx1, nbins, ed);
  const int y0 = bb_bin_search(dy0, nbins, ed), y1 = bb_bin_search(dy1, nbins, ed);
  const int rx0 = bb_bin_search(-dx1, nbins, ed), rx1 = bb_bin_search(-dx0, nbins, ed);
  const int ry0 = bb_bin_search(-dy1, nbins, ed), ry1 = bb_bin_search(-dy0, nbins, ed);
  const int ex = x1 - x0, ey = y1 - y0;
  int cls;
  if (ex > 1 || ey > 1 || rx0 != nbins - 1 - x1 || rx1 != nbins - 1 - x0 || ry0 != nbins - 1 - y1 ||
      ry1 != nbins - 1 - y0)
    cls = BB_GENERIC;
  else
    cls = (ax < M && ay < M && r2min >= lo2) ? BB_REG_FULL : BB_REG_CHECK;
  return cls | ((ex & 1) << 2) | ((ey & 1) << 3) | (x0 << 8) | (y0 << 20);
}

__device__ __forceinline__ void bb_prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(__cvta_generic_to_global(p)));
}

// CTAs per SM: 2 (248 registers, no spills: 251 ms for 100 resamples at N = 200k) measured faster than 3 (168
// registers, 300 bytes of spills: 268 ms)
#ifndef BB_MIN_CTAS
#define BB_MIN_CTAS 2
#endif
constexpr int BB_WARP_SMEM = 2 * 1024 + 512;   // per warp: scx, scy, msk

// A CTA (4 warps) owns one work item = (group of 32 resamples, row block of 32 points) x all column chunks from the
// row block on; what depends only on the item -- the row constants and multiplicities, the two row-prefix tables --
// is built once in shared memory, and the warps pull rounds of 32 column chunks from a shared counter.
__global__ void __launch_bounds__(BB_WARPS * 32, BB_MIN_CTAS)
bootbin_kernel(BBParams P) {
  extern __shared__ __align__(16) unsigned char bb_smem[];
  const int nbins = P.nbins, nb = P.nb, bpad = P.bpad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  double* ed = reinterpret_cast<double*>(bb_smem);              // nbins + 1 thresholds, padded to even
  double* mtab = ed + ((nbins + 2) & ~1);                       // [16][4]: the bits of a nibble as 0.0 / 1.0
  double2* tab = reinterpret_cast<double2*>(mtab + 64);         // [2 axes][33][32 lanes] row-prefix sums {a, c}
  double4* rowc = reinterpret_cast<double4*>(tab + 2 * 33 * 32);       // [32] constants of the row points
  unsigned char* rowm = reinterpret_cast<unsigned char*>(rowc + 32);   // [32 rows][32 resample lanes]
  double* rs = reinterpret_cast<double*>(rowm + 1024);          // [2][32] row coordinates in ascending order
  double2* rxy = reinterpret_cast<double2*>(rs + 64);           // [32] row coordinates, storage order
  long long* ctl = reinterpret_cast<long long*>(rxy + 32);      // [0] item of the CTA, [1] round counter
  unsigned char* wbase = reinterpret_cast<unsigned char*>(ctl + 2) + (size_t)warp * BB_WARP_SMEM;
  double4* scx = reinterpret_cast<double4*>(wbase);             // [32] constants of the chunk's points, x-sorted
  double4* scy = scx + 32;                                      // [32] the same, y-sorted
  unsigned* msk = reinterpret_cast<unsigned*>(scy + 32);        // [32 rows][4]: in range, & x bit, & y bit, & both

  for (int i = tid; i <= nbins; i += blockDim.x) ed[i] = P.edges[i];
  for (int e = tid; e < 64; e += blockDim.x) mtab[e] = (((e >> 2) >> (e & 3)) & 1) ? 1.0 : 0.0;

  const double M = P.hi, lo2 = P.lo2;
  const double NaN = __longlong_as_double(0x7ff8000000000000ll);
  const int G = P.ngroups;
  const int64_t n = P.n, nblk = P.nblk;
  unsigned st_closed = 0, st_sweep = 0, st_pair = 0, st_slow = 0, st_flush = 0, st_back = 0, st_gen = 0;

  // window registers: element u = ux + 2 uy holds the sums over pairs with (x bit >= ux) and (y bit >= uy)
  double wacc[4][3];
#pragma unroll
  for (int u = 0; u < 4; ++u) wacc[u][0] = wacc[u][1] = wacc[u][2] = 0.0;
  int fx0 = -1, fy0 = -1;
  unsigned touched = 0u;   // window bins that received something

  while (true) {
    __syncthreads();   // everybody is done with the previous item's shared data
    if (tid == 0) {
      ctl[0] = (long long)atomicAdd(P.counter, 1ull);
      ctl[1] = 0;
    }
    __syncthreads();
    const int64_t q = ctl[0];
    if (q >= P.my_items) break;
    const int64_t item = q * P.nranks + P.rank;   // largest items (small ib) first
    const int g = (int)(item % G);
    const int64_t ib = item / G;
    if (ib >= nblk) break;
    const int gb = g * 32 + lane;              // this lane's resample
    const int64_t c_lo = ib, c_hi = nblk;

    // ---- the item's row points (lanes = rows) and their per-resample values (lanes = resamples) ----
    const int64_t ig = ib * BB_CHUNK + lane;
    const bool live = ig < n;
    const double xi = live ? P.px[ig] : NaN, yi = live ? P.py[ig] : NaN;
    const double* rrec = P.geo + (size_t)ib * BB_GEO;
    if (warp == 0) {
      rowc[lane] = P.pt[ig];
      rxy[lane] = make_double2(xi, yi);
      rs[lane] = rrec[4 + lane];          // ascending x of the row block (+inf for absent points)
      rs[32 + lane] = rrec[36 + lane];    // ascending y
    }
    {
      const uint8_t* mr = P.m_row + (size_t)ib * BB_CHUNK * bpad + gb;
#pragma unroll
      for (int i = warp * 8; i < warp * 8 + 8; ++i) rowm[i * 32 + lane] = mr[(size_t)i * bpad];
    }
    __syncthreads();
    auto row_vals = [&](int i, double& ai, double& ci) {
      const double Mi = bb_magic(rowm[i * 32 + lane]);
      const double4 rc = rowc[i];
      ai = fma(Mi, rc.x, rc.y);
      ci = fma(Mi, rc.z, rc.w);
    };
    // prefix sums of a, c over the rows in ascending x (warp 0) / y (warp 1): tab[axis][e] = sum of the first e rows
    if (warp < 2) {
      const int axis = warp;
      const int prk = reinterpret_cast<const unsigned char*>(rrec + 68)[32 * axis + lane];
      double2* tb = tab + axis * 33 * 32 + lane;
      double pa = 0.0, pc = 0.0;
      tb[0] = make_double2(0.0, 0.0);
#pragma unroll 4
      for (int k = 0; k < 32; ++k) {
        const int i = __shfl_sync(0xffffffffu, prk, k);
        double ai, ci;
        row_vals(i, ai, ci);
        pa += ai;
        pc += ci;
        tb[(k + 1) * 32] = make_double2(pa, pc);
      }
    }
    const double iminx = bb_warp_min(live ? xi : INFINITY), imaxx = bb_warp_max(live ? xi : -INFINITY);
    const double iminy = bb_warp_min(live ? yi : INFINITY), imaxy = bb_warp_max(live ? yi : -INFINITY);
    const int nlive = __popc(__ballot_sync(0xffffffffu, live));
    __syncthreads();
    const double RA = tab[32 * 32 + lane].x, RC = tab[32 * 32 + lane].y;   // sums over all rows

    // ---- helpers -------------------------------------------------------------------------------------
    auto flush = [&]() {
      if (fx0 >= 0 && touched) {
        ++st_flush;
        double v[4][3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          v[0][k] = (wacc[0][k] - wacc[1][k]) - (wacc[2][k] - wacc[3][k]);
          v[1][k] = wacc[1][k] - wacc[3][k];
          v[2][k] = wacc[2][k] - wacc[3][k];
          v[3][k] = wacc[3][k];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int bx = fx0 + (u & 1), by = fy0 + (u >> 1);
          if ((touched >> u) & 1u) {   // a touched bin is inside the grid
            double* h = P.hist + ((size_t)(by * nbins + bx)) * bpad + gb;
#pragma unroll
            for (int k = 0; k < 3; ++k) atomicAdd(h + (size_t)k * nb * bpad, v[u][k]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) wacc[u][0] = wacc[u][1] = wacc[u][2] = 0.0;
      touched = 0u;
      fx0 = -1;
    };
    // exact per-pair path (lanes = columns in storage order): forward and mirrored bin of every pair from the
    // thresholds, straight into the histograms
    auto slow_block = [&](int64_t c, bool diag) {
      ++st_slow;
      const int64_t jg = c * BB_CHUNK + lane;
      const bool clive = jg < n;
      const double xj = clive ? P.px[jg] : NaN, yj = clive ? P.py[jg] : NaN;
      double* hf = P.hist;
      double* hc = P.hist + (size_t)3 * nb * bpad;
      for (int i = 0; i < nlive; ++i) {
        const double2 ri = rxy[i];
        const double dx = xj - ri.x, dy = yj - ri.y;
        const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        const bool inr = clive && r2 >= lo2 && fabs(dx) < M && fabs(dy) < M && (!diag || lane > i);
        int b1 = -1, b2 = -1;
        if (inr) {
          b1 = bb_bin_twod(dy, M, P.inv_bin, nbins, ed) * nbins + bb_bin_twod(dx, M, P.inv_bin, nbins, ed);
          b2 = bb_bin_twod(-dy, M, P.inv_bin, nbins, ed) * nbins + bb_bin_twod(-dx, M, P.inv_bin, nbins, ed);
        }
        unsigned mask = __ballot_sync(0xffffffffu, inr);
        if (!mask) continue;
        double ai, ci;
        row_vals(i, ai, ci);
        while (mask) {
          const int j = __ffs(mask) - 1;
          mask &= mask - 1;
          const int bb1 = __shfl_sync(0xffffffffu, b1, j), bb2 = __shfl_sync(0xffffffffu, b2, j);
          const double Mj = bb_magic(P.m_row[((size_t)c * BB_CHUNK + j) * bpad + gb]);
          const double4 k4 = P.pt[c * BB_CHUNK + j];
          const double aj = fma(Mj, k4.x, k4.y), cj = fma(Mj, k4.z, k4.w);
          const double v0 = ai * aj, v1 = ci * aj + ai * cj, v2 = ci * cj;
          const size_t s = (size_t)nb * bpad;
          double* h = hf + (size_t)bb1 * bpad + gb;
          atomicAdd(h, v0); atomicAdd(h + s, v1); atomicAdd(h + 2 * s, v2);
          if (bb2 != nb - 1 - bb1) {   // displacement within rounding of a bin edge
            double* h1 = hc + (size_t)(nb - 1 - bb1) * bpad + gb;
            double* h2 = hc + (size_t)bb2 * bpad + gb;
            atomicAdd(h1, -v0); atomicAdd(h1 + s, -v1); atomicAdd(h1 + 2 * s, -v2);
            atomicAdd(h2, v0); atomicAdd(h2 + s, v1); atomicAdd(h2 + 2 * s, v2);
          }
        }
      }
    };
    // Blocks whose displacements span more than 2 x 2 bins (sparse catalogues, small bins): the local bin of every
    // pair is found once (lanes = columns, exact thresholds); then, bin by bin of the block's window, the pairs of
    // that bin are summed per resample lane (masked FMAs over the chunk's 32 columns in registers, nibbles without
    // a bit skipped) into registers -- three red.global per BIN of the block instead of three per pair.
    // false: not handled (window of more than 64 bins, or a pair whose mirrored bin is not the mirror image).
    auto generic_block = [&](int64_t c, bool diag, int x0, int y0) -> bool {
      const double* grec = P.geo + (size_t)c * BB_GEO;
      const double4 bb = *reinterpret_cast<const double4*>(grec);
      const int x1 = bb_bin_search(bb.y - iminx, nbins, ed), y1 = bb_bin_search(bb.w - iminy, nbins, ed);
      const int wxn = x1 - x0 + 1, wyn = y1 - y0 + 1;
      if (wxn < 1 || wyn < 1 || wxn * wyn > 64) return false;
      const unsigned char* cperm = reinterpret_cast<const unsigned char*>(grec + 68);
      const int cnt = (int)((n - c * BB_CHUNK < BB_CHUNK) ? (n - c * BB_CHUNK) : BB_CHUNK);
      const bool clive = lane < cnt;
      const double sx = grec[4 + lane];
      const int pcx = cperm[lane];
      const double yjx = clive ? P.py[c * BB_CHUNK + pcx] : NaN;
      const uint4* mpx = reinterpret_cast<const uint4*>(P.m_sx + ((size_t)c * bpad + gb) * 32);
      const uint4 ux0 = mpx[0], ux1 = mpx[1];
      const unsigned wbx[8] = {ux0.x, ux0.y, ux0.z, ux0.w, ux1.x, ux1.y, ux1.z, ux1.w};
      __syncwarp();
      scx[lane] = P.pt[c * BB_CHUNK + pcx];
      __syncwarp();
      // local bin of every pair of this lane's column (255: not in range), four rows per word
      unsigned lbw[8];
      unsigned anymism = 0u;
#pragma unroll
      for (int k = 0; k < 8; ++k) lbw[k] = 0u;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const double2 ri = rxy[i];
        const double dx = sx - ri.x, dy = yjx - ri.y;
        const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        const bool inr = clive && i < nlive && r2 >= lo2 && fabs(dx) < M && fabs(dy) < M && (!diag || pcx > i);
        unsigned lb = 255u;
        if (inr) {
          const int bx = bb_bin_twod(dx, M, P.inv_bin, nbins, ed), by = bb_bin_twod(dy, M, P.inv_bin, nbins, ed);
          const int mx = bb_bin_twod(-dx, M, P.inv_bin, nbins, ed), my = bb_bin_twod(-dy, M, P.inv_bin, nbins, ed);
          if (mx != nbins - 1 - bx || my != nbins - 1 - by) anymism = 1u;
          lb = (unsigned)((by - y0) * wxn + (bx - x0));
        }
        lbw[i >> 2] |= lb << (8 * (i & 3));
      }
      if (__any_sync(0xffffffffu, anymism != 0u)) return false;
      double a[32], cc[32];
#pragma unroll
      for (int p = 0; p < 32; ++p) {
        const double Mj = bb_magic((wbx[p >> 2] >> (8 * (p & 3))) & 0xffu);
        const double4 k4 = scx[p];
        a[p] = fma(Mj, k4.x, k4.y);
        cc[p] = fma(Mj, k4.z, k4.w);
      }
      const double4* mt4 = reinterpret_cast<const double4*>(mtab);
      const size_t s3 = (size_t)nb * bpad;
#pragma unroll 1
      for (int lb = 0; lb < wxn * wyn; ++lb) {
        unsigned mymask = 0u;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const unsigned m = __ballot_sync(0xffffffffu, ((lbw[i >> 2] >> (8 * (i & 3))) & 0xffu) == (unsigned)lb);
          if (lane == i) mymask = m;
        }
        unsigned rows = __ballot_sync(0xffffffffu, mymask != 0u);
        if (!rows) continue;
        double v0 = 0.0, v1 = 0.0, v2 = 0.0;
        while (rows) {
          const int i = __ffs(rows) - 1;
          rows &= rows - 1;
          const unsigned W = __shfl_sync(0xffffffffu, mymask, i);
          double ai, ci;
          row_vals(i, ai, ci);
          double sa0 = 0.0, sa1 = 0.0, sc0 = 0.0, sc1 = 0.0;
#pragma unroll
          for (int q8 = 0; q8 < 8; ++q8) {
            const unsigned nib = (W >> (4 * q8)) & 15u;
            if (nib) {
              const double4 m = mt4[nib];
              sa0 = fma(a[4 * q8], m.x, sa0);     sc0 = fma(cc[4 * q8], m.x, sc0);
              sa1 = fma(a[4 * q8 + 1], m.y, sa1); sc1 = fma(cc[4 * q8 + 1], m.y, sc1);
              sa0 = fma(a[4 * q8 + 2], m.z, sa0); sc0 = fma(cc[4 * q8 + 2], m.z, sc0);
              sa1 = fma(a[4 * q8 + 3], m.w, sa1); sc1 = fma(cc[4 * q8 + 3], m.w, sc1);
            }
          }
          const double sa = sa0 + sa1, scv = sc0 + sc1;
          v0 = fma(ai, sa, v0);
          v1 = fma(ci, sa, fma(ai, scv, v1));
          v2 = fma(ci, scv, v2);
        }
        double* h = P.hist + (size_t)((y0 + lb / wxn) * nbins + x0 + lb % wxn) * bpad + gb;
        atomicAdd(h, v0); atomicAdd(h + s3, v1); atomicAdd(h + 2 * s3, v2);
      }
      return true;
    };
    // Sweep along one axis: the sums over the pairs whose bit along that axis is set (upper bin), for the 32
    // resamples of this group.  Column p (ascending along the axis, coordinate s) pairs with the rows whose
    // coordinate is small enough, fl(s - r_k) >= t: the first nle rows in ascending order, so its share is its own
    // (a, c) times the row-prefix sums tab[axis][nle] -- found by bisection in the column's lane, no dependence
    // between columns.  false: the exact mirrored split differs somewhere (the caller goes pair by pair).
    auto sweep = [&](int axis, int v0, double s, int cnt, const unsigned (&wb)[8], const double4* sc, double& h0,
                     double& h1, double& h2) -> bool {
      const double t = ed[v0 + 1], rt = ed[nbins - 1 - v0];
      const double* rsa = rs + 32 * axis;
      int lo = 0, hi = 32;   // nle = number of rows with fl(s - r_k) >= t (true for the small r_k)
#pragma unroll
      for (int it = 0; it < 6; ++it) {
        if (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if ((s - rsa[mid]) >= t) lo = mid + 1; else hi = mid;
        }
      }
      const int nle = lo;
      // exact mirrored bits: rows in the upper forward bin must sit in the lower mirrored bin, fl(r_k - s) < rt,
      // the others in the upper one; in ascending order only the two neighbours of the split can fail
      const bool ok = lane >= cnt || ((nle == 0 || (rsa[nle - 1] - s) < rt) &&
                                      (nle >= nlive || (rsa[nle] - s) >= rt));
      if (!__all_sync(0xffffffffu, ok)) return false;
      const double2* tb = tab + axis * 33 * 32 + lane;
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
      for (int p = 0; p < 32; p += 2) {
        {
          const int e = __shfl_sync(0xffffffffu, nle, p);
          const double Mj = bb_magic((wb[p >> 2] >> (8 * (p & 3))) & 0xffu);
          const double4 k4 = sc[p];
          const double2 T = tb[e * 32];
          const double A = fma(Mj, k4.x, k4.y), C = fma(Mj, k4.z, k4.w);
          a0 = fma(A, T.x, a0);
          a1 = fma(C, T.x, fma(A, T.y, a1));
          a2 = fma(C, T.y, a2);
        }
        {
          const int e = __shfl_sync(0xffffffffu, nle, p + 1);
          const double Mj = bb_magic((wb[(p + 1) >> 2] >> (8 * ((p + 1) & 3))) & 0xffu);
          const double4 k4 = sc[p + 1];
          const double2 T = tb[e * 32];
          const double A = fma(Mj, k4.x, k4.y), C = fma(Mj, k4.z, k4.w);
          b0 = fma(A, T.x, b0);
          b1 = fma(C, T.x, fma(A, T.y, b1));
          b2 = fma(C, T.y, b2);
        }
      }
      h0 = a0 + b0; h1 = a1 + b1; h2 = a2 + b2;
      return true;
    };
    // lines the next block will read, into L1 (the loads of a block are otherwise one exposed L2 round trip each)
    auto prefetch_block = [&](int64_t c, int d) {
      const bool closed = (d & 3) == BB_REG_FULL && !(d & 12) && (P.paths & 1);
      const char* cs = reinterpret_cast<const char*>(P.csum + (size_t)c * 3 * bpad + g * 32);
      if (closed || (d & 3) == BB_GENERIC) {
        if (lane < 4) bb_prefetch_l1(cs + (lane >> 1) * (size_t)bpad * 8 + (lane & 1) * 128);
        return;
      }
      const char* a;
      const size_t mo = ((size_t)c * bpad + g * 32) * 32;
      if (lane < 8) a = reinterpret_cast<const char*>(P.m_sx) + mo + lane * 128;
      else if (lane < 16) a = reinterpret_cast<const char*>(P.m_sy) + mo + (lane - 8) * 128;
      else if (lane < 24) a = reinterpret_cast<const char*>(P.pt + c * BB_CHUNK) + (lane - 16) * 128;
      else if (lane < 26) a = reinterpret_cast<const char*>(P.py + (c * BB_CHUNK + 16 * (lane - 24) < n ? c * BB_CHUNK : 0)) + (lane - 24) * 128;
      else if (lane < 31) a = reinterpret_cast<const char*>(P.geo + (size_t)c * BB_GEO) + (lane - 26) * 128;
      else a = cs;
      bb_prefetch_l1(a);
      if (lane < 3) bb_prefetch_l1(cs + (lane == 0 ? 128 : (size_t)bpad * 8 + (lane - 1) * 128));
    };

    while (true) {
      long long rd = 0;
      if (lane == 0) rd = (long long)atomicAdd(reinterpret_cast<unsigned long long*>(ctl + 1), 1ull);
      rd = __shfl_sync(0xffffffffu, rd, 0);
      const int64_t sc0 = c_lo + rd * 32;
      if (sc0 >= c_hi) break;
      // ---- lane l classifies column chunk sc0 + l ----
      const int64_t mychunk = sc0 + lane;
      int desc = 0;
      if (mychunk < c_hi) {
        const double4 bb = *reinterpret_cast<const double4*>(P.geo + (size_t)mychunk * BB_GEO);
        desc = bb_classify(iminx, imaxx, iminy, imaxy, bb.x, bb.y, bb.z, bb.w, M, lo2, nbins, ed);
        if (desc != 0 && mychunk == ib) {   // diagonal block: j > i, range test (r2 != 0) per pair
          if ((desc & 3) == BB_REG_FULL) desc = (desc & ~3) | BB_REG_CHECK;
          desc |= 1 << 4;
        }
      }
      unsigned todo = __ballot_sync(0xffffffffu, desc != 0);
      if (todo) prefetch_block(sc0 + __ffs(todo) - 1, __shfl_sync(0xffffffffu, desc, __ffs(todo) - 1));
      while (todo) {
        const int tl = __ffs(todo) - 1;
        todo &= todo - 1;
        const int d = __shfl_sync(0xffffffffu, desc, tl);
        const int64_t c = sc0 + tl;
        if (todo) prefetch_block(sc0 + __ffs(todo) - 1, __shfl_sync(0xffffffffu, desc, __ffs(todo) - 1));
        const int cls = d & 3, ex = (d >> 2) & 1, ey = (d >> 3) & 1, x0 = (d >> 8) & 0xfff, y0 = (d >> 20) & 0xfff;
        const bool diag = (d >> 4) & 1;
        if (cls == BB_GENERIC || !(P.paths & 2)) {
          if (cls == BB_GENERIC && (P.paths & 2) && (P.paths & 8)) {
            flush();
            if (generic_block(c, diag, x0, y0)) { ++st_gen; continue; }
          }
          slow_block(c, diag);
          continue;
        }
        const bool full = cls == BB_REG_FULL;
        // ---- window ----
        {
          const bool fitx = ex ? (x0 == fx0) : (x0 == fx0 || x0 == fx0 + 1);
          const bool fity = ey ? (y0 == fy0) : (y0 == fy0 || y0 == fy0 + 1);
          if (fx0 < 0 || !fitx || !fity) {
            flush();
            fx0 = x0;
            fy0 = y0;
          }
        }
        const int cx = x0 - fx0, cy = y0 - fy0;   // constant bits of the axes that do not vary
        double L[4][3];                           // block sums: all, x bit, y bit, both
#pragma unroll
        for (int u = 0; u < 4; ++u) L[u][0] = L[u][1] = L[u][2] = 0.0;
        bool booked = false;
        if (full) {   // every pair in range: the "all" sums are products of row-block and chunk sums
          const double* cs = P.csum + (size_t)c * 3 * bpad + gb;
          const double CA = cs[0], CC = cs[bpad];
          L[0][0] = RA * CA;
          L[0][1] = fma(RC, CA, RA * CC);
          L[0][2] = RC * CC;
          if (!ex && !ey && (P.paths & 1)) { ++st_closed; booked = true; }
        }
        if (!booked) {
          // ---- the chunk in ascending x (and y) order: coordinates in the lanes, constants staged ----
          const double* grec = P.geo + (size_t)c * BB_GEO;
          const unsigned char* cperm = reinterpret_cast<const unsigned char*>(grec + 68);
          const int cnt = (int)((n - c * BB_CHUNK < BB_CHUNK) ? (n - c * BB_CHUNK) : BB_CHUNK);
          const bool clive = lane < cnt;           // absent points sort last (+inf) and carry zero multiplicities
          const double sx = grec[4 + lane], sy = grec[36 + lane];
          const int pcx = cperm[lane], pcy = cperm[32 + lane];
          const uint4* mpx = reinterpret_cast<const uint4*>(P.m_sx + ((size_t)c * bpad + gb) * 32);
          const uint4 ux0 = mpx[0], ux1 = mpx[1];
          const unsigned wbx[8] = {ux0.x, ux0.y, ux0.z, ux0.w, ux1.x, ux1.y, ux1.z, ux1.w};
          const double yjx = clive ? P.py[c * BB_CHUNK + pcx] : NaN;   // y of the x-sorted columns
          __syncwarp();
          scx[lane] = P.pt[c * BB_CHUNK + pcx];
          scy[lane] = P.pt[c * BB_CHUNK + pcy];
          __syncwarp();
          bool done_x = false, done_y = false;
          if (full && (P.paths & 4)) {
            if (ex) {
              done_x = sweep(0, x0, sx, cnt, wbx, scx, L[1][0], L[1][1], L[1][2]);
              if (done_x) ++st_sweep; else ++st_back;
            }
            if (ey) {
              const uint4* mpy = reinterpret_cast<const uint4*>(P.m_sy + ((size_t)c * bpad + gb) * 32);
              const uint4 uy0 = mpy[0], uy1 = mpy[1];
              const unsigned wby[8] = {uy0.x, uy0.y, uy0.z, uy0.w, uy1.x, uy1.y, uy1.z, uy1.w};
              done_y = sweep(1, y0, sy, cnt, wby, scy, L[2][0], L[2][1], L[2][2]);
              if (done_y) ++st_sweep; else ++st_back;
            }
          }
          const bool need_t = !full || !(P.paths & 1 || ex || ey);
          const bool need_x = ex && !done_x, need_y = ey && !done_y, need_xy = ex && ey;
          if (need_t || need_x || need_y || need_xy) {
            ++st_pair;
            // ---- geometry (lanes = columns in ascending x): per-row masks of the pairs in range and their bits ----
            const double tx = ed[x0 + 1], rtx = ed[nbins - 1 - x0], ty = ed[y0 + 1], rty = ed[nbins - 1 - y0];
            uint2* msk2 = reinterpret_cast<uint2*>(msk);   // [half of the columns][row]: {t | x << 16, y | xy << 16}
            unsigned bad = 0u;
            if (!need_t && !need_x && !need_y) {
              // both axes came from sweeps (which also settled the mirrored bits): only the "both bits" quadrant is left
#pragma unroll 4
              for (int i = 0; i < 32; ++i) {
                const double2 ri = rxy[i];
                const unsigned wxy = __ballot_sync(0xffffffffu, (sx - ri.x) >= tx && (yjx - ri.y) >= ty);
                if (lane < 2) msk2[lane * 32 + i] = make_uint2(0u, (lane ? wxy >> 16 : wxy & 0xffffu) << 16);
              }
            } else {
              const unsigned kt = need_t ? 0xffffu : 0u, kx = need_x ? 0xffffu : 0u, ky = need_y ? 0xffffu : 0u;
              const unsigned kxy = need_xy ? 0xffffu : 0u;
#pragma unroll 2
              for (int i = 0; i < 32; ++i) {
                const double2 ri = rxy[i];
                const double dx = sx - ri.x, dy = yjx - ri.y;
                bool inr = true;
                if (!full) {
                  const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                  inr = r2 >= lo2 && fabs(dx) < M && fabs(dy) < M && (!diag || pcx > i);
                }
                const bool bx = ex && dx >= tx, by = ey && dy >= ty;
                const bool qx = (-dx) >= rtx, qy = (-dy) >= rty;
                const bool wrong = inr && clive && i < nlive && ((ex && bx == qx) || (ey && by == qy));
                const unsigned wt = __ballot_sync(0xffffffffu, inr), wx = __ballot_sync(0xffffffffu, inr && bx);
                const unsigned wy = __ballot_sync(0xffffffffu, inr && by);
                const unsigned wxy = __ballot_sync(0xffffffffu, inr && bx && by);
                bad |= __ballot_sync(0xffffffffu, wrong);
                if (lane < 2) {
                  const int sh = 16 * lane;
                  msk2[lane * 32 + i] = make_uint2(((wt >> sh) & kt) | (((wx >> sh) & kx) << 16),
                                                   ((wy >> sh) & ky) | (((wxy >> sh) & kxy) << 16));
                }
              }
            }
            __syncwarp();
            if (bad) {   // some displacement sits on a bin edge: exact per-pair path for the whole block
              --st_pair;
              slow_block(c, diag);
              continue;
            }
            // ---- accumulation (lanes = resamples): 16 columns at a time in registers ----
            if (need_t) L[0][0] = L[0][1] = L[0][2] = 0.0;
            const double4* mt4 = reinterpret_cast<const double4*>(mtab);
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
              const unsigned w4[4] = {h ? wbx[4] : wbx[0], h ? wbx[5] : wbx[1], h ? wbx[6] : wbx[2], h ? wbx[7] : wbx[3]};
              double a[16], cc[16];
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) {
                const double Mj = bb_magic((w4[jj >> 2] >> (8 * (jj & 3))) & 0xffu);
                const double4 k4 = scx[h * 16 + jj];
                a[jj] = fma(Mj, k4.x, k4.y);
                cc[jj] = fma(Mj, k4.z, k4.w);
              }
              // sum of a[j], c[j] over the set bits of the 16-bit W: each nibble fetches four 0.0 / 1.0 factors from
              // the table (a predicated FP64 add costs ptxas an add + two selects; the factor inside an FMA nothing);
              // four independent chains of eight FMAs (skipping the nibbles without a bit measured slower: 306 vs 267 ms)
              auto masked = [&](unsigned W, double& sa, double& sc) {
                const double4 m0 = mt4[W & 15u], m1 = mt4[(W >> 4) & 15u], m2 = mt4[(W >> 8) & 15u], m3 = mt4[W >> 12];
                double sa0 = a[0] * m0.x, sa1 = a[1] * m0.y, sc0 = cc[0] * m0.x, sc1 = cc[1] * m0.y;
                sa0 = fma(a[2], m0.z, sa0);  sa1 = fma(a[3], m0.w, sa1);  sc0 = fma(cc[2], m0.z, sc0);  sc1 = fma(cc[3], m0.w, sc1);
                sa0 = fma(a[4], m1.x, sa0);  sa1 = fma(a[5], m1.y, sa1);  sc0 = fma(cc[4], m1.x, sc0);  sc1 = fma(cc[5], m1.y, sc1);
                sa0 = fma(a[6], m1.z, sa0);  sa1 = fma(a[7], m1.w, sa1);  sc0 = fma(cc[6], m1.z, sc0);  sc1 = fma(cc[7], m1.w, sc1);
                sa0 = fma(a[8], m2.x, sa0);  sa1 = fma(a[9], m2.y, sa1);  sc0 = fma(cc[8], m2.x, sc0);  sc1 = fma(cc[9], m2.y, sc1);
                sa0 = fma(a[10], m2.z, sa0); sa1 = fma(a[11], m2.w, sa1); sc0 = fma(cc[10], m2.z, sc0); sc1 = fma(cc[11], m2.w, sc1);
                sa0 = fma(a[12], m3.x, sa0); sa1 = fma(a[13], m3.y, sa1); sc0 = fma(cc[12], m3.x, sc0); sc1 = fma(cc[13], m3.y, sc1);
                sa0 = fma(a[14], m3.z, sa0); sa1 = fma(a[15], m3.w, sa1); sc0 = fma(cc[14], m3.z, sc0); sc1 = fma(cc[15], m3.w, sc1);
                sa = sa0 + sa1;
                sc = sc0 + sc1;
              };
              // the row loop is software-pipelined: row i+1's mask words and values are fetched while row i is summed
              uint2 mk_n = msk2[h * 32];
              unsigned mb_n = rowm[lane];
              double4 rc_n = rowc[0];
#pragma unroll 1
              for (int i = 0; i < 32; ++i) {
                const uint2 mk = mk_n;
                const unsigned mb = mb_n;
                const double4 rc = rc_n;
                const int i1 = (i + 1) & 31;
                mk_n = msk2[h * 32 + i1];
                mb_n = rowm[i1 * 32 + lane];
                rc_n = rowc[i1];
                if (!(mk.x | mk.y)) continue;
                const unsigned Wt = mk.x & 0xffffu, Wx = mk.x >> 16, Wy = mk.y & 0xffffu, Wxy = mk.y >> 16;
                const double Mi = bb_magic(mb);
                const double ai = fma(Mi, rc.x, rc.y), ci = fma(Mi, rc.z, rc.w);
                double sa, scv;
                if (Wt) {
                  masked(Wt, sa, scv);
                  L[0][0] = fma(ai, sa, L[0][0]); L[0][1] = fma(ci, sa, fma(ai, scv, L[0][1])); L[0][2] = fma(ci, scv, L[0][2]);
                }
                if (Wx) {
                  masked(Wx, sa, scv);
                  L[1][0] = fma(ai, sa, L[1][0]); L[1][1] = fma(ci, sa, fma(ai, scv, L[1][1])); L[1][2] = fma(ci, scv, L[1][2]);
                }
                if (Wy) {
                  masked(Wy, sa, scv);
                  L[2][0] = fma(ai, sa, L[2][0]); L[2][1] = fma(ci, sa, fma(ai, scv, L[2][1])); L[2][2] = fma(ci, scv, L[2][2]);
                }
                if (Wxy) {
                  masked(Wxy, sa, scv);
                  L[3][0] = fma(ai, sa, L[3][0]); L[3][1] = fma(ci, sa, fma(ai, scv, L[3][1])); L[3][2] = fma(ci, scv, L[3][2]);
                }
              }
            }
          }
        }
        // ---- block sums into the window registers ----
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ux = u & 1, uy = u >> 1;
          const bool allowed = (!ux || ex || cx) && (!uy || ey || cy);
          if (allowed) {
            const int src = (ux && ex ? 1 : 0) + (uy && ey ? 2 : 0);
#pragma unroll
            for (int k = 0; k < 3; ++k)
              wacc[u][k] += (src == 0 ? L[0][k] : (src == 1 ? L[1][k] : (src == 2 ? L[2][k] : L[3][k])));
          }
        }
        // bins of the window this block can have touched: {cx .. cx+ex} x {cy .. cy+ey}
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ux = u & 1, uy = u >> 1;
          const bool inx = ex ? true : (ux == cx), iny = ey ? true : (uy == cy);
          if (inx && iny) touched |= 1u << u;
        }
      }
    }
    flush();   // the next item may belong to another resample group
  }
  if (lane == 0) {
    if (st_closed) atomicAdd(&g_bb_stats[0], (unsigned long long)st_closed);
    if (st_sweep) atomicAdd(&g_bb_stats[1], (unsigned long long)st_sweep);
    if (st_pair) atomicAdd(&g_bb_stats[2], (unsigned long long)st_pair);
    if (st_slow) atomicAdd(&g_bb_stats[3], (unsigned long long)st_slow);
    if (st_flush) atomicAdd(&g_bb_stats[4], (unsigned long long)st_flush);
    if (st_back) atomicAdd(&g_bb_stats[5], (unsigned long long)st_back);
    if (st_gen) atomicAdd(&g_bb_stats[6], (unsigned long long)st_gen);
  }
}

// ---- pre-pass 1 (independent of the resamples): chunk boxes, sorted copies with permutations, point constants ----
__global__ void __launch_bounds__(256)
bootbin_geo_kernel(const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pz,
                   const double* __restrict__ pw, int64_t n, int64_t nblk, double* __restrict__ geo,
                   double4* __restrict__ pt) {
  const int lane = threadIdx.x & 31;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= nblk) return;
  const int64_t j = c * BB_CHUNK + lane;
  const bool ok = j < n;
  const double x = ok ? px[j] : 0.0, y = ok ? py[j] : 0.0;
  const double w = ok ? (pw ? pw[j] : 1.0) : 0.0;
  const double wz = ok ? w * pz[j] : 0.0;
  pt[j] = make_double4(w, -0x1p52 * w, wz, -0x1p52 * wz);
  const double a = bb_warp_min(ok ? x : INFINITY), b = bb_warp_max(ok ? x : -INFINITY);
  const double cc = bb_warp_min(ok ? y : INFINITY), d = bb_warp_max(ok ? y : -INFINITY);
  double* o = geo + (size_t)c * BB_GEO;
  if (lane == 0) { o[0] = a; o[1] = b; o[2] = cc; o[3] = d; }
  unsigned char* perm = reinterpret_cast<unsigned char*>(o + 68);
#pragma unroll
  for (int axis = 0; axis < 2; ++axis) {
    double key = ok ? (axis == 0 ? x : y) : INFINITY;
    int idx = lane;
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
      for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
        const double okey = __shfl_xor_sync(0xffffffffu, key, j2);
        const int oidx = __shfl_xor_sync(0xffffffffu, idx, j2);
        const bool take_min = ((lane & j2) == 0) == ((lane & k2) == 0);
        // ties broken by the index: a total order, so the network is a permutation
        const bool less = okey < key || (okey == key && oidx < idx);
        const bool greater = okey > key || (okey == key && oidx > idx);
        if (take_min ? less : greater) { key = okey; idx = oidx; }
      }
    }
    o[4 + 32 * axis + lane] = key;
    perm[32 * axis + lane] = (unsigned char)idx;
  }
}

// ---- pre-pass 2: the multiplicities of one (chunk, group of 32 resamples) in the four layouts + chunk sums ----
__global__ void __launch_bounds__(256)
bootbin_mult_kernel(const uint8_t* __restrict__ mult, int64_t n, int64_t nblk, int32_t nboot, int32_t bpad,
                    const double* __restrict__ geo, const double4* __restrict__ pt, const double* __restrict__ pz,
                    uint8_t* __restrict__ m_row, uint8_t* __restrict__ m_sx, uint8_t* __restrict__ m_sy,
                    double* __restrict__ csum) {
  __shared__ unsigned tile_all[8][32][9];   // [warp][resample][column bytes, 36 B pitch]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int G = bpad / 32;
  const int64_t wid = (int64_t)blockIdx.x * 8 + warp;
  if (wid >= nblk * G) return;
  const int64_t c = wid / G;
  const int g = (int)(wid % G);
  unsigned char* tile = reinterpret_cast<unsigned char*>(&tile_all[warp][0][0]);
  const int64_t j = c * BB_CHUNK + lane;
  for (int bb = 0; bb < 32; ++bb) {   // lane = column: coalesced 32-byte reads of the b-major input
    const int b = g * 32 + bb;
    tile[bb * 36 + lane] = (b < nboot && j < n) ? mult[(size_t)b * n + j] : (unsigned char)0;
  }
  __syncwarp();
  const unsigned char* perm = reinterpret_cast<const unsigned char*>(geo + (size_t)c * BB_GEO + 68);
  const int gb = g * 32 + lane;
  // lane = resample
  unsigned sx[8], sy[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sx[k] = 0u; sy[k] = 0u; }
  double CA = 0.0, CC = 0.0, CZ = 0.0;
  for (int p = 0; p < 32; ++p) {
    const int jx = perm[p], jy = perm[32 + p];
    const unsigned bx = tile[lane * 36 + jx], by = tile[lane * 36 + jy];
    const unsigned m = tile[lane * 36 + p];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (k == (p >> 2)) { sx[k] |= bx << (8 * (p & 3)); sy[k] |= by << (8 * (p & 3)); }
    }
    const int64_t jp = c * BB_CHUNK + p;
    const double4 k4 = pt[jp];
    const double Mj = bb_magic(m);
    CA += fma(Mj, k4.x, k4.y);
    CC += fma(Mj, k4.z, k4.w);
    CZ += (jp < n) ? (double)m * pz[jp] : 0.0;
    m_row[(size_t)jp * bpad + gb] = (unsigned char)m;
  }
  uint4* o;
  o = reinterpret_cast<uint4*>(m_sx + ((size_t)c * bpad + gb) * 32);
  o[0] = make_uint4(sx[0], sx[1], sx[2], sx[3]); o[1] = make_uint4(sx[4], sx[5], sx[6], sx[7]);
  o = reinterpret_cast<uint4*>(m_sy + ((size_t)c * bpad + gb) * 32);
  o[0] = make_uint4(sy[0], sy[1], sy[2], sy[3]); o[1] = make_uint4(sy[4], sy[5], sy[6], sy[7]);
  double* cs = csum + (size_t)c * 3 * bpad + gb;
  cs[0] = CA; cs[bpad] = CC; cs[2 * (size_t)bpad] = CZ;
}

// ---- pre-pass 3: delta_b = mean of the resampled (centred) values, fixed summation order ----
__global__ void __launch_bounds__(256)
bootbin_delta_kernel(const double* __restrict__ csum, int64_t nblk, int32_t bpad, int64_t n, double* __restrict__ delta) {
  __shared__ double part[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gb = blockIdx.x * 32 + lane;
  double s = 0.0;
  for (int64_t c = warp; c < nblk; c += 8) s += csum[((size_t)c * 3 + 2) * bpad + gb];
  part[warp][lane] = s;
  __syncthreads();
  if (warp == 0) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += part[k][lane];
    delta[gb] = t / (double)n;
  }
}

// ---- xi of every resample from the (all-reduced) sums ----
__global__ void bootbin_xi_kernel(const double* __restrict__ hist, const double* __restrict__ delta, int32_t nb,
                                  int32_t bpad, int32_t nboot, double* __restrict__ xi, double* __restrict__ sumw) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)nb * bpad) return;
  const int b = (int)(t % bpad), bin = (int)(t / bpad);
  if (b >= nboot) return;
  const size_t s = (size_t)nb * bpad;
  const double* hf = hist;
  const double* hc = hist + 3 * s;
  const size_t o = (size_t)bin * bpad + b, om = (size_t)(nb - 1 - bin) * bpad + b;
  const double s0 = hf[o] + hf[om] + hc[o];
  const double s1 = hf[s + o] + hf[s + om] + hc[s + o];
  const double s2 = hf[2 * s + o] + hf[2 * s + om] + hc[2 * s + o];
  const double d = delta[b];
  const double kk = (s2 - d * s1) + (d * d) * s0;
  xi[(size_t)b * nb + bin] = s0 != 0.0 ? kk / s0 : 0.0;
  if (sumw) sumw[(size_t)b * nb + bin] = s0;
}

static int g_bb_paths = 15;  // tgp_set_option("bootbin_paths", bits): 1 whole-block bookings, 2 window paths (else
                             // every block exact pair by pair), 4 sweeps, 8 bin-by-bin form of wide-window blocks
extern "C" int tgp_bootbin_set_paths(int bits) { g_bb_paths = bits; return TGP_OK; }

extern "C" int tgp_bootbin_stats(unsigned long long* host8, int reset) {
  TGP_CHECK_ARG(host8 != nullptr, "host8");
  TGP_CUDA(cudaMemcpyFromSymbol(host8, g_bb_stats, sizeof(unsigned long long) * 8));
  if (reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    TGP_CUDA(cudaMemcpyToSymbol(g_bb_stats, z, sizeof(z)));
  }
  return TGP_OK;
}

static inline int64_t bb_align(int64_t v) { return (v + 255) & ~(int64_t)255; }
struct BBLayout {
  int64_t geo, pt, m_row, m_sx, m_sy, csum, total;
};
static BBLayout bb_layout(int64_t n, int32_t bpad) {
  const int64_t nblk = tgp_cdiv(n, BB_CHUNK);
  BBLayout L;
  int64_t o = 0;
  L.geo = o; o = bb_align(o + nblk * BB_GEO * 8);
  L.pt = o; o = bb_align(o + nblk * BB_CHUNK * 32);
  L.m_row = o; o = bb_align(o + nblk * BB_CHUNK * (int64_t)bpad);
  L.m_sx = o; o = bb_align(o + nblk * BB_CHUNK * (int64_t)bpad);
  L.m_sy = o; o = bb_align(o + nblk * BB_CHUNK * (int64_t)bpad);
  L.csum = o; o = bb_align(o + nblk * 3 * (int64_t)bpad * 8);
  L.total = o;
  return L;
}

extern "C" int64_t tgp_bootbin_work_bytes(int64_t n, int32_t nboot) {
  if (n < 0 || nboot < 0) return 0;
  return bb_layout(n, (int32_t)(tgp_cdiv(nboot, 32) * 32)).total;
}

extern "C" int64_t tgp_bootbin_sums_doubles(int32_t nbins, int32_t nboot) {
  if (nbins < 1 || nboot < 0) return 0;
  return (int64_t)6 * nbins * nbins * (tgp_cdiv(nboot, 32) * 32);
}

__device__ unsigned long long g_bb_counters[64];

extern "C" int tgp_bootbin_twod(const double* px, const double* py, const double* pz, const double* pw, int64_t n,
                                const uint8_t* mult, int32_t nboot, const double* edges, int32_t nbins,
                                double min_sep2, double max_sep, int32_t tile_rank, int32_t tile_nranks,
                                double* sums, double* delta, void* work, void* stream) {
  TGP_CHECK_ARG(n >= 0 && nboot >= 0 && nbins >= 1 && nbins <= 2048, "n/nboot/nbins");
  TGP_CHECK_ARG(tile_nranks >= 1 && tile_rank >= 0 && tile_rank < tile_nranks, "rank");
  TGP_CHECK_ARG(max_sep > 0.0 && min_sep2 >= 0.0, "separations");
  if (nboot == 0) return TGP_OK;
  TGP_CHECK_ARG(px && py && pz && mult && edges && sums && delta && work, "null pointer");
  TGP_CHECK_ARG(((uintptr_t)work % 256) == 0, "work must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t bpad = (int32_t)(tgp_cdiv(nboot, 32) * 32);
  const int32_t nb = nbins * nbins;
  const int64_t nblk = tgp_cdiv(n, BB_CHUNK);
  TGP_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 6 * (size_t)nb * bpad, st));
  TGP_CUDA(cudaMemsetAsync(delta, 0, sizeof(double) * bpad, st));
  if (n < 2) return TGP_OK;

  const BBLayout L = bb_layout(n, bpad);
  unsigned char* wk = reinterpret_cast<unsigned char*>(work);
  BBParams P;
  P.px = px; P.py = py;
  P.pt = reinterpret_cast<const double4*>(wk + L.pt);
  P.geo = reinterpret_cast<const double*>(wk + L.geo);
  P.m_row = wk + L.m_row; P.m_sx = wk + L.m_sx; P.m_sy = wk + L.m_sy;
  P.csum = reinterpret_cast<const double*>(wk + L.csum);
  P.edges = edges;
  P.hist = sums;
  P.lo2 = min_sep2 > 0.0 ? min_sep2 : 4.9406564584124654e-324;   // r2 != 0
  P.hi = max_sep;
  P.inv_bin = (double)nbins / (2.0 * max_sep);
  P.n = n; P.nblk = nblk; P.nbins = nbins; P.nb = nb; P.bpad = bpad; P.ngroups = bpad / 32;
  P.rank = tile_rank; P.nranks = tile_nranks; P.paths = g_bb_paths;

  bootbin_geo_kernel<<<(unsigned)tgp_cdiv(nblk * 32, 256), 256, 0, st>>>(
      px, py, pz, pw, n, nblk, reinterpret_cast<double*>(wk + L.geo), reinterpret_cast<double4*>(wk + L.pt));
  TGP_LAUNCH_CHECK();
  bootbin_mult_kernel<<<(unsigned)tgp_cdiv(nblk * P.ngroups, 8), 256, 0, st>>>(
      mult, n, nblk, nboot, bpad, P.geo, P.pt, pz, wk + L.m_row, wk + L.m_sx, wk + L.m_sy,
      reinterpret_cast<double*>(wk + L.csum));
  TGP_LAUNCH_CHECK();
  bootbin_delta_kernel<<<(unsigned)P.ngroups, 256, 0, st>>>(P.csum, nblk, bpad, n, delta);
  TGP_LAUNCH_CHECK();

  // shared memory of a CTA: thresholds, nibble masks, two row-prefix tables, row data, per-warp staging
  const size_t smem = (size_t)((nbins + 2) & ~1) * 8 + 512 + 2 * 33 * 32 * 16 + 1024 + 1024 + 512 + 512 + 16 +
                      (size_t)BB_WARPS * BB_WARP_SMEM;
  TGP_CHECK_ARG(smem <= 200 * 1024, "too many bins");
  const int sms = tgp_num_sms();
  const int64_t grid_target = (int64_t)sms * BB_MIN_CTAS;
  // work decomposition: item = (resample group, row block) x all column chunks from the row block on
  P.run = 0;
  P.items_per_group = nblk;
  const int64_t total_items = nblk * P.ngroups;
  P.my_items = (total_items - tile_rank + tile_nranks - 1) / tile_nranks;
  if (P.my_items <= 0) return TGP_OK;
  int64_t grid = P.my_items;
  if (grid > grid_target) grid = grid_target;

  static std::atomic<unsigned> launch_seq{0};
  static void* cbase_dev[TGP_MAX_DEVICES] = {};
  void*& cbase = cbase_dev[tgp_current_device()];
  if (!cbase) TGP_CUDA(cudaGetSymbolAddress(&cbase, g_bb_counters));
  P.counter = reinterpret_cast<unsigned long long*>(cbase) + (launch_seq.fetch_add(1u) % 64);
  TGP_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(unsigned long long), st));
  static TgpPerDeviceOnce once_;
  if (tgp_first_use_on_device(once_))
    TGP_CUDA(cudaFuncSetAttribute(bootbin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  bootbin_kernel<<<(unsigned)grid, BB_WARPS * 32, smem, st>>>(P);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

extern "C" int tgp_bootbin_xi(const double* sums, const double* delta, int32_t nbins, int32_t nboot, double* xi,
                              double* sumw, void* stream) {
  TGP_CHECK_ARG(nbins >= 1 && nboot >= 0, "nbins/nboot");
  if (nboot == 0) return TGP_OK;
  TGP_CHECK_ARG(sums && delta && xi, "null pointer");
  const int32_t bpad = (int32_t)(tgp_cdiv(nboot, 32) * 32);
  const int32_t nb = nbins * nbins;
  bootbin_xi_kernel<<<(unsigned)tgp_cdiv((int64_t)nb * bpad, 256), 256, 0, (cudaStream_t)stream>>>(
      sums, delta, nb, bpad, nboot, xi, sumw);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}
