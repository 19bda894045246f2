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
// triangle.  Each thread owns one row point in registers; column tiles are staged in shared memory
// and read as broadcasts.  Histograms are privatised per warp in shared memory (32-bit native
// atomics for counts, FP64 CAS adds for sums) and flushed to global memory with red.global once per
// catalogue.  Multi-GPU: runs are dealt round-robin to ranks; the caller all-reduces the bin sums.
#include <float.h>
#include <math.h>
#include "tgp_common.cuh"

constexpr int PB_T = 256;       // points per tile (rows per CTA = threads per CTA)
constexpr int PB_WARPS = PB_T / 32;
constexpr int PB_FLUSH_TILEPAIRS = 200000;  // keeps the 32-bit private counters from overflowing

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

__device__ __forceinline__ int pb_bin_log(double r2, int nbins, const double* __restrict__ ed) {
  // number of interior thresholds ed[1..nbins-1] that are <= r2 (binary search)
  int lo = 0, hi = nbins - 1;  // answer in [lo, hi]
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (r2 >= ed[mid]) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <int BT, bool WEIGHTED>
__global__ void __launch_bounds__(PB_T)
pairbin_kernel(PBParams P) {
  extern __shared__ __align__(16) unsigned char pb_smem[];
  const int nb = P.nb, ncopy = P.ncopy;
  double* ed = reinterpret_cast<double*>(pb_smem);                 // nbins + 1
  double* hs = ed + ((P.nbins + 2) & ~1);                          // ncopy * nb   sum wk wk
  double* hw = hs + (size_t)ncopy * nb;                            // ncopy * nb   sum w w      (WEIGHTED)
  double* hr = hw + (WEIGHTED ? (size_t)ncopy * nb : 0);           // ncopy * nb   sum w w r    (LOG)
  double* tx = hr + (BT == TGP_BIN_LOG ? (size_t)ncopy * nb : 0);  // PB_T each
  double* ty = tx + PB_T;
  double* tk = ty + PB_T;
  double* tw = tk + PB_T;
  unsigned int* hc = reinterpret_cast<unsigned int*>(tw + PB_T);   // ncopy * nb   counts

  const int tid = threadIdx.x, warp = tid >> 5;
  const int copy = warp % ncopy;
  double* my_s = hs + (size_t)copy * nb;
  double* my_w = hw + (size_t)copy * nb;
  double* my_r = hr + (size_t)copy * nb;
  unsigned int* my_c = hc + (size_t)copy * nb;

  for (int i = tid; i <= P.nbins; i += PB_T) ed[i] = P.edges[i];
  auto clear_hist = [&]() {
    for (int i = tid; i < ncopy * nb; i += PB_T) {
      hs[i] = 0.0;
      hc[i] = 0u;
      if (WEIGHTED) hw[i] = 0.0;
      if (BT == TGP_BIN_LOG) hr[i] = 0.0;
    }
  };
  auto flush_hist = [&](int cat) {
    __syncthreads();
    if (cat >= 0) {
      for (int b = tid; b < nb; b += PB_T) {
        unsigned long long c = 0;
        double s = 0.0, w = 0.0, r = 0.0;
        for (int k = 0; k < ncopy; ++k) {
          c += hc[k * nb + b];
          s += hs[k * nb + b];
          if (WEIGHTED) w += hw[k * nb + b];
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
    bool live = false;
    for (; p < p_end; ++p) {
      if (I != loadedI) {
        const int64_t ig = I * PB_T + tid;
        live = ig < n;
        if (live) {
          xi = P.px[off + ig];
          yi = P.py[off + ig];
          wi = WEIGHTED ? P.pw[off + ig] : 1.0;
          ki = P.pk[off + ig] * wi;
        }
        loadedI = I;
      }
      __syncthreads();  // previous column tile fully consumed
      {
        const int64_t jg = J * PB_T + tid;
        const bool ok = jg < n;
        tx[tid] = ok ? P.px[off + jg] : 0.0;
        ty[tid] = ok ? P.py[off + jg] : 0.0;
        const double w = (ok && WEIGHTED) ? P.pw[off + jg] : 1.0;
        tk[tid] = ok ? P.pk[off + jg] * w : 0.0;
        if (WEIGHTED) tw[tid] = w;
      }
      __syncthreads();
      const int jcount = (int)((n - J * PB_T < PB_T) ? (n - J * PB_T) : PB_T);
      const int jstart = (I == J) ? tid + 1 : 0;  // diagonal tile: j > i only
      if (live) {
        for (int jj = jstart; jj < jcount; ++jj) {
          const double dx = tx[jj] - xi, dy = ty[jj] - yi;
          const double r2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));  // no FMA: matches the oracle bit for bit
          if (BT == TGP_BIN_TWOD) {
            if (r2 >= P.lo2 && fabs(dx) < P.hi && fabs(dy) < P.hi) {
              const int b1 = pb_bin_twod(dy, P.hi, P.inv_bin, P.nbins, ed) * P.nbins +
                             pb_bin_twod(dx, P.hi, P.inv_bin, P.nbins, ed);
              const int b2 = pb_bin_twod(-dy, P.hi, P.inv_bin, P.nbins, ed) * P.nbins +
                             pb_bin_twod(-dx, P.hi, P.inv_bin, P.nbins, ed);
              const double kk = ki * tk[jj];
              atomicAdd(my_c + b1, 1u);
              atomicAdd(my_c + b2, 1u);
              atomicAdd(my_s + b1, kk);
              atomicAdd(my_s + b2, kk);
              if (WEIGHTED) {
                const double ww = wi * tw[jj];
                atomicAdd(my_w + b1, ww);
                atomicAdd(my_w + b2, ww);
              }
            }
          } else {
            if (r2 >= P.lo2 && r2 < P.hi) {
              const int b = pb_bin_log(r2, P.nbins, ed);
              const double ww = WEIGHTED ? wi * tw[jj] : 1.0;
              atomicAdd(my_c + b, 1u);
              atomicAdd(my_s + b, ki * tk[jj]);
              if (WEIGHTED) atomicAdd(my_w + b, ww);
              atomicAdd(my_r + b, ww * sqrt(r2));
            }
          }
        }
      }
      ++since_flush;
      if (++J == nt) { ++I; J = I; }
    }
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
  const size_t fixed = (size_t)((nbins + 2) & ~1) * 8 + 4 * PB_T * 8;
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
