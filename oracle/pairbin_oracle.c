/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the product path
 * (treegp_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * CPU restatement of the pair binning that /root/reference/treegp/two_pcf.py:297-305 (TwoD,
 * bin_slop=0) and :330-338 (Log) delegate to TreeCorr's KKCorrelation.process.
 *
 * PARITY UNPINNED: TreeCorr (`treecorr>=4.2` setup.py:46, `>=5.0` requirements.txt:6) is an
 * un-vendored third-party dependency that is neither installed nor installable here, and the reference
 * ships no golden xi/npairs vectors (its tests only check self-consistency,
 * tests/test_hyp_search.py:47).  This file therefore restates TreeCorr's published brute-force
 * semantics (bin_slop=0 makes TreeCorr's tree traversal equal to brute force up to summation order):
 *
 *  TwoD (BinTypeHelper<TwoD>): bin_size = 2*max_sep/nbins.  A pair with d = p2 - p1 is used iff
 *      rsq != 0, rsq >= min_sep^2 and max(|dx|,|dy|) < max_sep.  i = int((dx+max_sep)/bin_size),
 *      j = int((dy+max_sep)/bin_size), each clamped from nbins to nbins-1, index j*nbins+i.  For an
 *      auto-correlation every unordered pair is entered at d and at -d (point-symmetric grid, the
 *      property two_pcf.py:306-321 relies on).
 *  Log: bin_size = ln(max_sep/min_sep)/nbins; used iff min_sep^2 <= rsq < max_sep^2;
 *      k = int((0.5*ln(rsq) - ln(min_sep))/bin_size) clamped to [0, nbins-1]; each unordered pair once;
 *      meanr = sum(w r)/sum(w).
 *  xi = sum(w_i k_i w_j k_j) / sum(w_i w_j); weight = sum(w_i w_j); npairs = count.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -ffp-contract=off: no FMA contraction, so rsq and
 * the bin indices are plain IEEE double arithmetic).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int twod_index(double dx, double dy, double max_sep, double bin_size, int nbins) {
  int i = (int)((dx + max_sep) / bin_size);
  int j = (int)((dy + max_sep) / bin_size);
  if (i == nbins) --i;
  if (j == nbins) --j;
  return j * nbins + i;
}

/* Rows [i_begin, i_end) of the pair matrix (i < j); whole job = [0, n).  Outputs are accumulated. */
void oracle_pairbin_twod(const double* x, const double* y, const double* k, const double* w, int64_t n,
                         int64_t i_begin, int64_t i_end, double min_sep, double max_sep, int nbins,
                         int64_t* npairs, double* sumw, double* sumwkk) {
  const double bin_size = 2.0 * max_sep / nbins;
  const double minsepsq = min_sep * min_sep;
  const int nb = nbins * nbins;
#pragma omp parallel
  {
    int64_t* c = (int64_t*)calloc(nb, sizeof(int64_t));
    double* sw = (double*)calloc(nb, sizeof(double));
    double* sk = (double*)calloc(nb, sizeof(double));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = i_begin; i < i_end; ++i) {
      const double xi = x[i], yi = y[i];
      const double wi = w ? w[i] : 1.0;
      const double wki = wi * k[i];
      for (int64_t j = i + 1; j < n; ++j) {
        const double dx = x[j] - xi, dy = y[j] - yi;
        const double rsq = dx * dx + dy * dy;
        if (rsq == 0.0 || rsq < minsepsq) continue;
        if (!(fmax(fabs(dx), fabs(dy)) < max_sep)) continue;
        const int b1 = twod_index(dx, dy, max_sep, bin_size, nbins);
        const int b2 = twod_index(-dx, -dy, max_sep, bin_size, nbins);
        const double wj = w ? w[j] : 1.0;
        const double ww = wi * wj;
        const double kk = wki * (wj * k[j]);
        c[b1] += 1; sw[b1] += ww; sk[b1] += kk;
        c[b2] += 1; sw[b2] += ww; sk[b2] += kk;
      }
    }
#pragma omp critical
    for (int b = 0; b < nb; ++b) { npairs[b] += c[b]; sumw[b] += sw[b]; sumwkk[b] += sk[b]; }
    free(c); free(sw); free(sk);
  }
}

void oracle_pairbin_log(const double* x, const double* y, const double* k, const double* w, int64_t n,
                        int64_t i_begin, int64_t i_end, double min_sep, double max_sep, int nbins,
                        int64_t* npairs, double* sumw, double* sumwkk, double* sumwr) {
  const double bin_size = log(max_sep / min_sep) / nbins;
  const double logminsep = log(min_sep);
  const double minsepsq = min_sep * min_sep, maxsepsq = max_sep * max_sep;
#pragma omp parallel
  {
    int64_t* c = (int64_t*)calloc(nbins, sizeof(int64_t));
    double* sw = (double*)calloc(nbins, sizeof(double));
    double* sk = (double*)calloc(nbins, sizeof(double));
    double* sr = (double*)calloc(nbins, sizeof(double));
#pragma omp for schedule(dynamic, 16)
    for (int64_t i = i_begin; i < i_end; ++i) {
      const double xi = x[i], yi = y[i];
      const double wi = w ? w[i] : 1.0;
      const double wki = wi * k[i];
      for (int64_t j = i + 1; j < n; ++j) {
        const double dx = x[j] - xi, dy = y[j] - yi;
        const double rsq = dx * dx + dy * dy;
        if (!(rsq >= minsepsq && rsq < maxsepsq)) continue;
        int b = (int)((0.5 * log(rsq) - logminsep) / bin_size);
        if (b < 0) b = 0;
        if (b >= nbins) b = nbins - 1;
        const double wj = w ? w[j] : 1.0;
        const double ww = wi * wj;
        c[b] += 1; sw[b] += ww; sk[b] += wki * (wj * k[j]); sr[b] += ww * sqrt(rsq);
      }
    }
#pragma omp critical
    for (int b = 0; b < nbins; ++b) { npairs[b] += c[b]; sumw[b] += sw[b]; sumwkk[b] += sk[b]; sumwr[b] += sr[b]; }
    free(c); free(sw); free(sk); free(sr);
  }
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
