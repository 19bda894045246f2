// Batched chi-square of the robust anisotropic 2PCF fit.
//
// Replaces B sequential evaluations of robust_2dfit.chi2 (/root/reference/treegp/two_pcf.py:115-148, which calls
// _model_skl :96-113 and get_correlation_length_matrix :12-31) -- the objective iminuit's MIGRAD probes 2 n + 1 times
// per gradient and n (n + 1) / 2 + n times per Hessian.  One CTA per parameter set (size, g1, g2):
//   L      = R^T diag(size^2, (size q)^2) R,  q = (1 - e) / (1 + e),  e = |(g1, g2)|,  angle atan2(g2, g1) / 2
//   model  = profile(coord^T L^-1 coord)                 at the P masked bin lags (amplitude 1)
//   alpha  = (F^T W F)^-1 F^T W y,  F = [model, 1];  alpha_0 <- |alpha_0|
//   chi2   = r^T W r,  r = y - alpha_0 model - alpha_1;  +inf for |g| > 1 or non-finite parameters
// P <= 1024 masked pixels (221 for the default 21 x 21 map); W is read from global memory (L2-resident, 390 KB).
#include <math.h>
#include "tgp_common.cuh"

constexpr int RF_T = 256;
constexpr int RF_MAXP = 1024;

__device__ __forceinline__ double rf_block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < RF_T / 32; ++w) s += red[w];   // fixed order: deterministic
  return s;
}

template <int FAM>
__global__ void __launch_bounds__(RF_T)
robust_chi2_kernel(const double* __restrict__ coord, const double* __restrict__ y, const double* __restrict__ W, int P,
                   const double* __restrict__ params, double* __restrict__ out, const double* __restrict__ phi_g) {
  __shared__ double model[RF_MAXP], wm[RF_MAXP], r[RF_MAXP];
  __shared__ double red[RF_T / 32];
  __shared__ double phi_s[FAM == TGP_FAM_VONKARMAN ? TGP_VK_PHI_SIZE : 1];
  const int tid = threadIdx.x;
  const double size = params[3 * blockIdx.x], g1 = params[3 * blockIdx.x + 1], g2 = params[3 * blockIdx.x + 2];
  double* o = out + 4 * blockIdx.x;
  if (!isfinite(size + g1 + g2) || fabs(g1) > 1.0 || fabs(g2) > 1.0) {   // two_pcf.py:127-133, :105-106
    if (tid == 0) { o[0] = INFINITY; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0; }
    return;
  }
  if (FAM == TGP_FAM_VONKARMAN) tgp_stage_phi(phi_s, phi_g);
  // correlation-length matrix and its inverse (two_pcf.py:23-30, :108-109)
  const double e = sqrt(g1 * g1 + g2 * g2), q = (1.0 - e) / (1.0 + e);
  const double ang = 0.5 * atan2(g2, g1), c = cos(ang), s = sin(ang);
  const double a = size * size, b = (size * q) * (size * q);
  const double L00 = c * c * a + s * s * b, L01 = c * s * a - s * c * b, L11 = s * s * a + c * c * b;
  const double det = L00 * L11 - L01 * L01;
  KDesc kd;
  kd.amp = 1.0; kd.m00 = L11 / det; kd.m01x2 = 2.0 * (-L01 / det); kd.m11 = L00 / det; kd.family = FAM; kd.ndim = 2;
  __syncthreads();
  for (int p = tid; p < P; p += RF_T) model[p] = tgp_profile<FAM>(tgp_qform(kd, coord[2 * p], coord[2 * p + 1]), phi_s);
  __syncthreads();
  // F^T W F and F^T W y:  W is symmetric up to rounding; follow the reference's order F^T W (row vector) . F
  double fwf00 = 0.0, fwf01 = 0.0, fwf10 = 0.0, fwf11 = 0.0, fy0 = 0.0, fy1 = 0.0;
  for (int p = tid; p < P; p += RF_T) {
    double mw = 0.0, ow = 0.0;                       // (model^T W)_p, (1^T W)_p : column p of W
    for (int k = 0; k < P; ++k) {
      const double w = W[(size_t)k * P + p];
      mw = fma(model[k], w, mw);
      ow += w;
    }
    fwf00 = fma(mw, model[p], fwf00); fwf01 += mw;
    fwf10 = fma(ow, model[p], fwf10); fwf11 += ow;
    fy0 = fma(mw, y[p], fy0); fy1 = fma(ow, y[p], fy1);
  }
  fwf00 = rf_block_sum(fwf00, red); fwf01 = rf_block_sum(fwf01, red);
  fwf10 = rf_block_sum(fwf10, red); fwf11 = rf_block_sum(fwf11, red);
  fy0 = rf_block_sum(fy0, red); fy1 = rf_block_sum(fy1, red);
  const double d2 = fwf00 * fwf11 - fwf01 * fwf10;
  double a0 = (fwf11 * fy0 - fwf01 * fy1) / d2, a1 = (-fwf10 * fy0 + fwf00 * fy1) / d2;
  a0 = fabs(a0);                                      // two_pcf.py:141
  for (int p = tid; p < P; p += RF_T) r[p] = y[p] - (a0 * model[p] + a1);
  __syncthreads();
  double chi = 0.0;
  for (int p = tid; p < P; p += RF_T) {
    double rw = 0.0;                                  // (r^T W)_p
    for (int k = 0; k < P; ++k) rw = fma(r[k], W[(size_t)k * P + p], rw);
    chi = fma(rw, r[p], chi);
  }
  chi = rf_block_sum(chi, red);
  (void)wm;
  if (tid == 0) { o[0] = chi; o[1] = a0; o[2] = a1; o[3] = 1.0; }
}

extern "C" int tgp_robust_chi2_batch(const double* coord, const double* y, const double* W, int32_t P, int32_t family,
                                     const double* params, int32_t nsets, double* out, void* stream) {
  TGP_CHECK_ARG(P >= 2 && P <= RF_MAXP && nsets >= 0, "P (2..1024) / nsets");
  TGP_CHECK_ARG(family == TGP_FAM_RBF || family == TGP_FAM_VONKARMAN, "family must be RBF or von Karman");
  if (nsets == 0) return TGP_OK;
  TGP_CHECK_ARG(coord && y && W && params && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const double* phi = tgp_phi_device();
  if (family == TGP_FAM_RBF)
    robust_chi2_kernel<TGP_FAM_RBF><<<(unsigned)nsets, RF_T, 0, st>>>(coord, y, W, P, params, out, phi);
  else
    robust_chi2_kernel<TGP_FAM_VONKARMAN><<<(unsigned)nsets, RF_T, 0, st>>>(coord, y, W, P, params, out, phi);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}
