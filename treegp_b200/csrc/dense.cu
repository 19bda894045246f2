// FP64 dense linear algebra for the GP solve: blocked right-looking Cholesky, triangular solves,
// log-determinant and the marginal log-likelihood.
//
// Replaces scipy.linalg.cholesky / cho_solve / np.dot at
//   /root/reference/treegp/gp_interp.py:181-191 and /root/reference/treegp/log_likelihood.py:29-37.
//
// Layout: row-major, lower triangle (numpy's default layout for K, `lower=True` factor).
//
// Everything O(N^3) funnels into ONE contraction kernel, gemm_nt_sub (C -= A * B^T with both
// operands K-contiguous), which runs on the FP64 tensor pipe via mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4;
// there is no tcgen05/TMEM path for FP64 -- larger PTX shapes lower to the same DMMA.8x8x4).
//   * Cholesky trailing update      A22 -= L21 L21^T            (lower tiles only)
//   * panel / multi-RHS solves      B2  -= B1  L21^T
//   * predictive covariance         C   -= V   V^T
// Operand tiles are brought in by a 4-stage cp.async ring into XOR-swizzled shared memory so the
// fragment loads are conflict-free 16-byte LDS; the K order inside a 16-wide stage is permuted
// (even k's then odd k's) so one LDS.128 feeds two DMMAs.
//
// The O(N^2 nb) parts (64x64 potf2, 64-wide triangular panel solve) are FP64-ALU kernels.
#include <float.h>
#include <math.h>
#include <string.h>
#include <atomic>
#include "tgp_common.cuh"

// ============================================================================================
// gemm_nt_sub
// ============================================================================================
constexpr int BM = 128, BK = 16, STAGES = 4, GEMM_THREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Load one BK-wide stage of a (ROWS x BK) operand tile: 8 chunks of 16 B per row.
// Chunk c of row r is stored at physical chunk c ^ ((r & 1) << 2).
template <int ROWS>
__device__ __forceinline__ void load_operand_stage(double* smem, const double* __restrict__ G,
                                                   int64_t ldg, int64_t row0, int64_t nrows,
                                                   int64_t k0, int64_t Kd, int tid) {
  const int c = tid & 7;
  const int rbase = tid >> 3;  // 0..31
  const int64_t k = k0 + 2 * c;
  int kbytes = (int)((Kd - k) * 8);
  kbytes = kbytes < 0 ? 0 : (kbytes > 16 ? 16 : kbytes);
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    const int r = rbase + 32 * i;
    const int64_t gr = row0 + r;
    const bool ok = gr < nrows;
    const double* src = G + (ok ? gr : 0) * ldg + (kbytes ? k : 0);
    const int pc = c ^ ((r & 1) << 2);
    cp_async16(smem + (r * 8 + pc) * 2, src, ok ? kbytes : 0);
  }
}

// C (M x Nc) -= A (M x Kd) * B^T (Nc x Kd).  lower_only: skip tiles above the diagonal and mask n > m.
// CTA tile BM x BN_, 8 warps arranged WARPS_M x WARPS_N.  Two configurations are instantiated:
//   <128, 2, 4, 1>: 128 x 128 tile, warp tile 64 x 32, one CTA per SM (largest operand reuse);
//   < 64, 4, 2, 2>: 128 x  64 tile, warp tile 32 x 32, two CTAs per SM so that the epilogue (C read-modify-
//                   write) and the pipeline fill of one CTA hide under the DMMA work of the other.
template <int BN_, int WARPS_M, int WARPS_N, int MIN_CTAS>
__global__ void __launch_bounds__(GEMM_THREADS, MIN_CTAS)
gemm_nt_sub_kernel(double* __restrict__ C, int64_t M, int64_t Nc, int64_t ldc,
                   const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
                   int64_t Kd, int lower_only) {
  constexpr int WM = BM / WARPS_M, WN = BN_ / WARPS_N;   // warp tile
  constexpr int MB = WM / 8, NBK = WN / 8;                // 8x8 blocks per warp tile
  constexpr int STAGE_D = (BM + BN_) * BK;
  static_assert(WARPS_M * WARPS_N == 8, "8 warps");
  extern __shared__ __align__(16) double gsm[];
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN_;
  if (lower_only && n0 > m0 + BM - 1) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * WN;
  const int g = lane >> 2, t = lane & 3;

  double acc[MB][NBK][2];
#pragma unroll
  for (int i = 0; i < MB; ++i)
#pragma unroll
    for (int j = 0; j < NBK; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int KT_ = (int)((Kd + BK - 1) / BK);
  auto issue = [&](int kt) {
    if (kt < KT_) {
      double* sa = gsm + (kt % STAGES) * STAGE_D;
      double* sb = sa + BM * BK;
      load_operand_stage<BM>(sa, A, lda, m0, M, (int64_t)kt * BK, Kd, tid);
      load_operand_stage<BN_>(sb, B, ldb, n0, Nc, (int64_t)kt * BK, Kd, tid);
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) issue(s);

  const int swz = (g & 1) << 2;
  for (int kt = 0; kt < KT_; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    issue(kt + STAGES - 1);
    const double* sa = gsm + (kt % STAGES) * STAGE_D;
    const double* sb = sa + BM * BK;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int pc = (h * 4 + t) ^ swz;
      double2 af[MB], bf[NBK];
#pragma unroll
      for (int i = 0; i < MB; ++i)
        af[i] = *reinterpret_cast<const double2*>(sa + ((wm0 + i * 8 + g) * 8 + pc) * 2);
#pragma unroll
      for (int j = 0; j < NBK; ++j)
        bf[j] = *reinterpret_cast<const double2*>(sb + ((wn0 + j * 8 + g) * 8 + pc) * 2);
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NBK; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
      for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int j = 0; j < NBK; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
    }
  }
  cp_async_wait<0>();

  const bool vec = ((ldc & 1) == 0) && ((((uintptr_t)C) & 15) == 0);
#pragma unroll
  for (int i = 0; i < MB; ++i) {
    const int64_t r = m0 + wm0 + i * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < NBK; ++j) {
      const int64_t c = n0 + wn0 + j * 8 + 2 * t;
      bool ok0 = c < Nc, ok1 = c + 1 < Nc;
      if (lower_only) {
        ok0 = ok0 && (c <= r);
        ok1 = ok1 && (c + 1 <= r);
      }
      double* p = C + r * ldc + c;
      if (vec && ok0 && ok1) {
        double2 v = *reinterpret_cast<double2*>(p);
        v.x -= acc[i][j][0];
        v.y -= acc[i][j][1];
        *reinterpret_cast<double2*>(p) = v;
      } else {
        if (ok0) p[0] -= acc[i][j][0];
        if (ok1) p[1] -= acc[i][j][1];
      }
    }
  }
}

static int g_lookahead = 1;  // 0 disables the two-stream look-ahead Cholesky driver
static int g_ob_large = 0;   // outer block of the look-ahead driver: 0 = automatic, else a multiple of 512 (option "potrf_ob")
static int g_fused_panel = 1;   // option "potrf_fused": 0 = the potf2 / trsm_panel / gemm chain per 64 columns
static int g_env_lookahead = 1;   // option "potrf_env_lookahead": 0 = the sequential envelope driver, 2 = see there
static int g_gemm_config = -1;  // -1: pick by shape; 0: 128x128; 1: 128x64 (tgp_set_option for experiments)

template <int BN_, int WARPS_M, int WARPS_N, int MIN_CTAS>
static int gemm_launch_cfg(double* C, int64_t M, int64_t Nc, int64_t ldc, const double* A, int64_t lda,
                           const double* B, int64_t ldb, int64_t Kd, int lower_only, cudaStream_t st) {
  constexpr int SMEM = STAGES * (BM + BN_) * BK * 8;
  static TgpPerDeviceOnce attr_once;
  if (tgp_first_use_on_device(attr_once)) {
    TGP_CUDA(cudaFuncSetAttribute(gemm_nt_sub_kernel<BN_, WARPS_M, WARPS_N, MIN_CTAS>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    // same shared-memory carve-out for every kernel of the factorisation: no L1/shared reconfiguration
    // between the back-to-back small launches of a panel
    TGP_CUDA(cudaFuncSetAttribute(gemm_nt_sub_kernel<BN_, WARPS_M, WARPS_N, MIN_CTAS>,
                                  cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  dim3 grid((unsigned)tgp_cdiv(Nc, BN_), (unsigned)tgp_cdiv(M, BM));
  TGP_CHECK_ARG(grid.y <= 65535u, "M too large for one launch");
  gemm_nt_sub_kernel<BN_, WARPS_M, WARPS_N, MIN_CTAS><<<grid, GEMM_THREADS, SMEM, st>>>(C, M, Nc, ldc, A, lda, B, ldb,
                                                                                       Kd, lower_only);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

static int gemm_nt_sub_launch(double* C, int64_t M, int64_t Nc, int64_t ldc, const double* A, int64_t lda,
                              const double* B, int64_t ldb, int64_t Kd, int lower_only, cudaStream_t st) {
  if (M <= 0 || Nc <= 0 || Kd <= 0) return TGP_OK;
  TGP_CHECK_ARG((lda % 2 == 0) && (ldb % 2 == 0), "leading dimensions must be even (16-byte rows)");
  TGP_CHECK_ARG(((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0), "operands must be 16-byte aligned");
  const int cfg = g_gemm_config >= 0 ? g_gemm_config : 1;
  if (cfg == 0) return gemm_launch_cfg<128, 2, 4, 1>(C, M, Nc, ldc, A, lda, B, ldb, Kd, lower_only, st);
  return gemm_launch_cfg<64, 4, 2, 2>(C, M, Nc, ldc, A, lda, B, ldb, Kd, lower_only, st);
}

extern "C" int tgp_pairbin_set_block_sums(int on);
extern "C" int tgp_pairbin_set_fast_paths(int bits);
extern "C" int tgp_trsv_set_cluster(int v);
extern "C" int tgp_bootbin_set_paths(int bits);
extern "C" int tgp_set_option(const char* name, int value) {
  if (name && !strcmp(name, "trsv_cluster")) return tgp_trsv_set_cluster(value);
  if (name && !strcmp(name, "gemm_config")) { g_gemm_config = value; return TGP_OK; }
  if (name && !strcmp(name, "potrf_lookahead")) { g_lookahead = value; return TGP_OK; }
  if (name && !strcmp(name, "potrf_fused")) { g_fused_panel = value; return TGP_OK; }
  if (name && !strcmp(name, "potrf_env_lookahead")) { g_env_lookahead = value; return TGP_OK; }
  if (name && !strcmp(name, "pairbin_block_sums")) return tgp_pairbin_set_block_sums(value);
  if (name && !strcmp(name, "pairbin_fast_paths")) return tgp_pairbin_set_fast_paths(value);
  if (name && !strcmp(name, "bootbin_paths")) return tgp_bootbin_set_paths(value);
  if (name && !strcmp(name, "potrf_ob") && value >= 0 && value % 512 == 0) { g_ob_large = value; return TGP_OK; }
  tgp_set_error("tgp_set_option: unknown option");
  return TGP_ERR_INVALID;
}

// ============================================================================================
// potf2: unblocked Cholesky of an n x n (n <= 64) diagonal block, one CTA.
// ============================================================================================
constexpr int NB = 64;           // inner block
constexpr int NB_PITCH = NB + 1;

// 256 threads, the block in shared memory, factorised in four 16-column steps so that only 4 x 3
// CTA barriers separate the phases instead of 64 x 2:
//   (i)   warp 0 factorises the 16 x 16 diagonal sub-block in registers (lane = row, shuffles),
//   (ii)  one thread per row below solves its 16 entries against that sub-block,
//   (iii) all threads apply the rank-16 update to the trailing lower triangle.
constexpr int PF_B = 16;
#ifdef TGP_PANEL_TIMING
__device__ unsigned long long g_pf_cycles[4];   // accumulated cycles of potf2 phases (i), (ii), (iii)
#define PF_T(k) if (threadIdx.x == 0) { const long long t_ = clock64(); g_pf_cycles[k] += t_ - pf_t0; pf_t0 = t_; }
#else
#define PF_T(k)
#endif
// Factorise the NB x NB tile S (pitch NB_PITCH, identity-padded beyond n) in shared memory; all 256 threads
// of the CTA call this.  rdiag receives 1 / L[j][j].  Ends with a CTA barrier.
__device__ __forceinline__ void potf2_smem(double* __restrict__ S, double* __restrict__ rdiag, int n,
                                           int32_t* __restrict__ info, int64_t global_off) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef TGP_PANEL_TIMING
  long long pf_t0 = clock64();
  if (tid == 0) g_pf_cycles[0] = g_pf_cycles[1] = g_pf_cycles[2] = 0;
#endif
  for (int k0 = 0; k0 < NB; k0 += PF_B) {
    // (i) 16 x 16 diagonal sub-block, lanes 0..15 hold one row each.  Pivots in groups of four: inside a group
    // every pivot updates only the group's own columns (<= 3 shuffle + DFMA pairs on the critical path); the
    // columns to the right take the group's rank-4 update afterwards, as independent shuffle / DFMA pairs.
    if (warp == 0) {
      double a[PF_B];
      const int l15 = lane & 15;
      const int r = k0 + l15;
      unsigned badmask = 0u;       // the same in every lane (d is a broadcast)
#pragma unroll
      for (int c = 0; c < PF_B; ++c) a[c] = S[r * NB_PITCH + k0 + c];
#pragma unroll
      for (int grp = 0; grp < PF_B / 4; ++grp) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const int j = 4 * grp + jj;
          const double d = __shfl_sync(0xffffffffu, a[j], j);           // pivot a_jj from lane j
          // not positive definite / not finite: remembered branch-free (bit j), reported once after the block
          badmask |= (k0 + j < n && (!(d > 0.0) || !isfinite(d))) ? (1u << j) : 0u;
          // one reciprocal square root per pivot: l_jj = d * rsqrt(d), l_ij = a_ij * rsqrt(d) (no FP64 divide or
          // square-root call on the critical path; results agree with sqrt/divide to ~1 ulp)
          const double rinv = rsqrt(d);
          a[j] = (l15 == j) ? d * rinv : ((l15 > j) ? a[j] * rinv : a[j]);
          if (lane == j) rdiag[k0 + j] = rinv;
#pragma unroll
          for (int c = 0; c < PF_B; ++c) {   // constant bounds + static predicate: keeps a[] in registers
            if (c > j && c < 4 * grp + 4) {
              const double lcj = __shfl_sync(0xffffffffu, a[j], c);     // l_cj from lane c
              a[c] = (l15 >= c) ? fma(-a[j], lcj, a[c]) : a[c];
            }
          }
        }
#pragma unroll
        for (int c = 0; c < PF_B; ++c) {
          if (c >= 4 * grp + 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const double lck = __shfl_sync(0xffffffffu, a[4 * grp + k], c);
              a[c] = (l15 >= c) ? fma(-a[4 * grp + k], lck, a[c]) : a[c];
            }
          }
        }
      }
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < PF_B; ++c) S[r * NB_PITCH + k0 + c] = a[c];
      }
      if (badmask && lane == 0)   // keep the first failure (LAPACK's info)
        atomicCAS(info, 0, (int32_t)(global_off + k0 + __ffs(badmask)));
    }
    __syncthreads();
    PF_T(0);
    // (ii) rows below: x <- x * L11^-T (thread per row)
    const int below = NB - k0 - PF_B;
    if (tid < below) {
      const int r = k0 + PF_B + tid;
      double x[PF_B];
#pragma unroll
      for (int c = 0; c < PF_B; ++c) x[c] = S[r * NB_PITCH + k0 + c];
#pragma unroll
      for (int j = 0; j < PF_B; ++j) {
        x[j] = x[j] * rdiag[k0 + j];
#pragma unroll
        for (int c = j + 1; c < PF_B; ++c) x[c] = fma(-x[j], S[(k0 + c) * NB_PITCH + k0 + j], x[c]);
      }
#pragma unroll
      for (int c = 0; c < PF_B; ++c) S[r * NB_PITCH + k0 + c] = x[c];
    }
    __syncthreads();
    PF_T(1);
    // (iii) trailing update S[i][c] -= sum_k S[i][k0+k] S[c][k0+k], k0+16 <= c <= i < 64, on the DMMA pipe:
    // the lower 8 x 8 blocks of the trailing square are dealt to the warps, K = 16 = four m8n8k4 steps each
    if (below > 0) {
      const int m0 = k0 + PF_B, nblk = below / 8, g8 = lane >> 2, t4 = lane & 3;
      for (int p = warp; p < nblk * (nblk + 1) / 2; p += 8) {
        int bi = 0, bj = p;
        while (bj > bi) { bj -= bi + 1; ++bi; }
        const double* Ar = S + (m0 + 8 * bi + g8) * NB_PITCH + k0 + t4;
        const double* Br = S + (m0 + 8 * bj + g8) * NB_PITCH + k0 + t4;
        double c0 = 0.0, c1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < PF_B / 4; ++ks) dmma884(c0, c1, Ar[4 * ks], Br[4 * ks]);
        double* Cr = S + (m0 + 8 * bi + g8) * NB_PITCH + m0 + 8 * bj + 2 * t4;
        Cr[0] -= c0;
        Cr[1] -= c1;
      }
    }
    __syncthreads();
    PF_T(2);
  }
}

// Load / store the lower triangle of an n x n (n <= NB) global block into the identity-padded tile S.
__device__ __forceinline__ void potf2_load(double* __restrict__ S, const double* __restrict__ A, int n, int64_t ld) {
#pragma unroll 4
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
    const int r = idx / NB, c = idx % NB;  // coalesced along c
    S[r * NB_PITCH + c] = (r < n && c <= r) ? A[(int64_t)r * ld + c] : (r == c ? 1.0 : 0.0);
  }
}
__device__ __forceinline__ void potf2_store(const double* __restrict__ S, double* __restrict__ A, int n, int64_t ld) {
#pragma unroll 4
  for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
    const int r = idx / NB, c = idx % NB;
    if (r < n && c <= r) A[(int64_t)r * ld + c] = S[r * NB_PITCH + c];
  }
}

__global__ void __launch_bounds__(256)
potf2_kernel(double* __restrict__ A, int n, int64_t ld, int32_t* __restrict__ info, int64_t global_off) {
  __shared__ double S[NB * NB_PITCH];
  __shared__ double rdiag[NB];   // 1 / L[j][j]
  potf2_load(S, A, n, ld);
  __syncthreads();
  potf2_smem(S, rdiag, n, info, global_off);
  potf2_store(S, A, n, ld);
}

// ============================================================================================
// trsm panel: B (M x nb) <- B * L^-T for an nb x nb (nb <= 64) lower-triangular L.  Thread per row,
// right-looking over columns so the FMAs of one step are independent.
// ============================================================================================
constexpr int TRSM_ROWS = 64;   // rows per CTA: small CTAs spread short panels over many SMs
constexpr int TRSM_SMEM = (NB * NB + NB + TRSM_ROWS * NB_PITCH) * 8;

// FULL: nb == 64 and 16-byte aligned rows: every thread streams its own row with 32 independent 16-byte
// loads (no staging, nothing to wait for but the L tile), solves in registers and streams it back.
// Otherwise (edge panels): rows are staged through shared memory element by element.
template <bool FULL>
__global__ void __launch_bounds__(TRSM_ROWS)
trsm_panel_kernel(const double* __restrict__ L, int nb, int64_t ldl, double* __restrict__ B, int64_t M,
                  int64_t ldb) {
  extern __shared__ __align__(16) double tsm[];
  double* LsT = tsm;                 // NB x NB, LsT[j*NB + i] = L[i][j] (column j contiguous), zero padded
  double* dinv = tsm + NB * NB;      // 1 / L[j][j]
  double* Bs = dinv + NB;            // TRSM_ROWS x NB_PITCH (edge path only)
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * TRSM_ROWS;
  const int64_t gr = r0 + tid;
  double x[NB];
  if (FULL) {
    if (gr < M) {
      const double2* row = reinterpret_cast<const double2*>(B + gr * ldb);
#pragma unroll
      for (int j = 0; j < NB / 2; ++j) { const double2 v = row[j]; x[2 * j] = v.x; x[2 * j + 1] = v.y; }
    } else {
#pragma unroll
      for (int j = 0; j < NB; ++j) x[j] = 0.0;
    }
  }
  if (FULL && ((ldl & 1) == 0) && (((uintptr_t)L & 15) == 0)) {
    // one batch of 32 independent 16-byte loads per thread: element pairs (i, 2c), (i, 2c+1) of the L tile
    double2 lv[NB * NB / 2 / TRSM_ROWS];
#pragma unroll
    for (int q = 0; q < NB * NB / 2 / TRSM_ROWS; ++q) {
      const int idx = tid + q * TRSM_ROWS;       // pair index: row i = idx / 32, column pair c = idx % 32
      lv[q] = *reinterpret_cast<const double2*>(L + (int64_t)(idx >> 5) * ldl + 2 * (idx & 31));
    }
#pragma unroll
    for (int q = 0; q < NB * NB / 2 / TRSM_ROWS; ++q) {
      const int idx = tid + q * TRSM_ROWS;
      const int i = idx >> 5, j = 2 * (idx & 31);
      LsT[j * NB + i] = (j < i) ? lv[q].x : 0.0;
      LsT[(j + 1) * NB + i] = (j + 1 < i) ? lv[q].y : 0.0;
    }
  } else {
#pragma unroll 16
    for (int idx = tid; idx < NB * NB; idx += TRSM_ROWS) {
      const int i = idx / NB, j = idx % NB;  // coalesced along j
      double v = 0.0;
      if (i < nb && j < i) v = L[(int64_t)i * ldl + j];
      LsT[j * NB + i] = v;
    }
  }
  if (tid < NB) dinv[tid] = (tid < nb) ? 1.0 / L[(int64_t)tid * ldl + tid] : 1.0;
  if (!FULL) {
    for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
      const int r = idx / NB, j = idx % NB;
      const int64_t g2 = r0 + r;
      Bs[r * NB_PITCH + j] = (g2 < M && j < nb) ? B[g2 * ldb + j] : 0.0;
    }
  }
  __syncthreads();
  if (!FULL) {
#pragma unroll
    for (int j = 0; j < NB; ++j) x[j] = Bs[tid * NB_PITCH + j];
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    x[j] *= dinv[j];
    const double nx = -x[j];
#pragma unroll
    for (int ip = ((j + 1) & ~1); ip < NB; ip += 2) {  // pairs (ip, ip+1); entries with i <= j are zero/skipped
      const double2 l = *reinterpret_cast<const double2*>(LsT + j * NB + ip);
      if (ip > j) x[ip] = fma(nx, l.x, x[ip]);
      x[ip + 1] = fma(nx, l.y, x[ip + 1]);
    }
  }
  if (FULL) {
    if (gr < M) {
      double2* row = reinterpret_cast<double2*>(B + gr * ldb);
#pragma unroll
      for (int j = 0; j < NB / 2; ++j) row[j] = make_double2(x[2 * j], x[2 * j + 1]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NB; ++j) Bs[tid * NB_PITCH + j] = x[j];
    __syncthreads();
    for (int idx = tid; idx < TRSM_ROWS * NB; idx += TRSM_ROWS) {
      const int r = idx / NB, j = idx % NB;
      const int64_t g2 = r0 + r;
      if (g2 < M && j < nb) B[g2 * ldb + j] = Bs[r * NB_PITCH + j];
    }
  }
}

static int trsm_panel_launch(const double* L, int nb, int64_t ldl, double* B, int64_t M, int64_t ldb,
                             cudaStream_t st) {
  if (M <= 0 || nb <= 0) return TGP_OK;
  static TgpPerDeviceOnce attr_once;
  if (tgp_first_use_on_device(attr_once)) {
    TGP_CUDA(cudaFuncSetAttribute(trsm_panel_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
    TGP_CUDA(cudaFuncSetAttribute(trsm_panel_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_SMEM));
    TGP_CUDA(cudaFuncSetAttribute(trsm_panel_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    TGP_CUDA(cudaFuncSetAttribute(trsm_panel_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    TGP_CUDA(cudaFuncSetAttribute(potf2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  }
  const bool full = (nb == NB) && ((ldb & 1) == 0) && (((uintptr_t)B & 15) == 0);
  const unsigned grid = (unsigned)tgp_cdiv(M, TRSM_ROWS);
  if (full) trsm_panel_kernel<true><<<grid, TRSM_ROWS, TRSM_SMEM, st>>>(L, nb, ldl, B, M, ldb);
  else trsm_panel_kernel<false><<<grid, TRSM_ROWS, TRSM_SMEM, st>>>(L, nb, ldl, B, M, ldb);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

// ============================================================================================
// Fused panel step.  A w-wide (w <= OB = 512) panel = the w x w diagonal block D at `Akk` plus `below` rows
// stored under it.  ONE launch per 64-wide block column j factorises D's diagonal block (j,j) and solves every
// row block under it, instead of the potf2 / trsm_panel / K=64 gemm chain (3 launches per block column inside D
// plus ~15 for the rows below).  Inside the panel the off-diagonal blocks are LEFT-looking and the diagonal
// blocks of D RIGHT-looking:
//   CTA 0 (row block j)      : load block (j,j) -- already fully updated --, potf2 in shared memory, store,
//                              publish a flag in global memory.
//   CTA i > 0 (row block r)  : T = A[r, j] - A[r, 0:j] A[j, 0:j]^T  on the DMMA pipe (K = 64 j <= 448; this
//                              overlaps with CTA 0's potf2), wait for the flag, X = T L_jj^-T by substitution
//                              (4 threads per row), store X; if r is a row block of D also apply its own rank-64
//                              update A[r, r] -= X X^T, so that block (j+1, j+1) is final when launch j+1 starts.
// Waiting CTAs spin on the flag; CTA 0 is always dispatched first, so the wait cannot deadlock; the spin is
// bounded anyway (info = -1 if it ever trips).
// ============================================================================================
constexpr int PL_THREADS = 256;
constexpr int PL_P = NB + 2;                                 // pitch of the T / L tiles (even: 16-byte LDS)
constexpr int PL_RING_D = STAGES * (NB + NB) * BK;            // operand ring, doubles
constexpr int PL_MP = 9;                                      // pitch of the 8 x 8 inverse blocks
constexpr int PL_TAIL_D = 2 * NB * PL_P + NB + 8 * 8 * PL_MP;  // T tile + L tile + 1/diag + 8 inverses (aliases the ring)
constexpr int PL_SMEM = (PL_RING_D > PL_TAIL_D ? PL_RING_D : PL_TAIL_D) * 8;
constexpr int PL_NFLAGS = 4096;
__device__ unsigned g_panel_flags[PL_NFLAGS];
// every flag-synchronised launch takes a fresh (slot, epoch) pair: a stale value in a reused slot never matches
static unsigned next_flag_epoch() {
  static std::atomic<unsigned> counter{0};
  unsigned e = counter.fetch_add(1u) + 1u;
  if (e == 0) e = counter.fetch_add(1u) + 1u;   // 0 is the value of a never-used flag
  return e;
}

#ifdef TGP_PANEL_TIMING   // phase timestamps of one launch (tools/panel_timing.py builds a private copy of the library)
__device__ unsigned long long g_pl_times[32];   // [0,16): globaltimer ns, [16,32): clock64 of the same stamps
__device__ __forceinline__ void pl_stamp(int i) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pl_times[i] = t;
    g_pl_times[16 + i] = (unsigned long long)clock64();
  }
}
#define PL_T(i) pl_stamp(i)
extern "C" int tgp_debug_panel_times(unsigned long long* host32) {
  return cudaMemcpyFromSymbol(host32, g_pl_times, sizeof(unsigned long long) * 32) == cudaSuccess ? 0 : -2;
}
extern "C" int tgp_debug_potf2_cycles(unsigned long long* host4) {
  return cudaMemcpyFromSymbol(host4, g_pf_cycles, sizeof(unsigned long long) * 4) == cudaSuccess ? 0 : -2;
}
#else
#define PL_T(i)
#endif

__global__ void __launch_bounds__(PL_THREADS, 2)
panel_left_kernel(double* __restrict__ Akk, int64_t ld, int w, int64_t below, int j, int32_t* __restrict__ info,
                  int64_t goff, unsigned slot, unsigned epoch) {
  extern __shared__ __align__(16) double psm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ndiag = (w + NB - 1) / NB;
  const int c0 = j * NB;
  const int nbj = (w - c0 < NB) ? (w - c0) : NB;
  unsigned* flag = &g_panel_flags[slot];

  if (blockIdx.x == 0) {
    double* S = psm;                       // NB x NB_PITCH
    double* rdiag = psm + NB * NB_PITCH;
    double* D = Akk + (int64_t)c0 * ld + c0;
    PL_T(0);
    potf2_load(S, D, nbj, ld);
    __syncthreads();
    PL_T(1);
    potf2_smem(S, rdiag, nbj, info, goff + c0);
    PL_T(2);
    potf2_store(S, D, nbj, ld);
    __threadfence();
    __syncthreads();
    if (tid == 0) atomicExch(flag, epoch);
    PL_T(3);
    return;
  }

  const int rb = j + blockIdx.x;           // row block: < ndiag inside D, else among the rows below
  const bool in_diag = rb < ndiag;
  int64_t r0;
  int nrows;
  if (in_diag) {
    r0 = (int64_t)rb * NB;
    nrows = (w - rb * NB < NB) ? (w - rb * NB) : NB;
  } else {
    r0 = w + (int64_t)(rb - ndiag) * NB;
    const int64_t left = w + below - r0;
    nrows = left < NB ? (int)left : NB;
  }
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp >> 2) * 32, wn0 = (warp & 3) * 16;   // warp tile 32 x 16 of the 64 x 64 block
#ifdef TGP_PANEL_TIMING
  const bool stamp = blockIdx.x == 1;
#undef PL_T
#define PL_T(i) if (stamp) pl_stamp(i)
#endif
  PL_T(8);

  // ---- T = A[r, j] (prefetched) - A[r, 0:c0] A[j, 0:c0]^T ------------------------------------
  double cv[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = wm0 + i * 8 + g;
#pragma unroll
    for (int jn = 0; jn < 2; ++jn) {
      const int c = wn0 + jn * 8 + 2 * t;
      const double* p = Akk + (r0 + r) * ld + c0 + c;
      double v0 = 0.0, v1 = 0.0;
      if (r < nrows) {
        if (c + 1 < nbj) {
          const double2 v = *reinterpret_cast<const double2*>(p);
          v0 = v.x;
          v1 = v.y;
        } else if (c < nbj) {
          v0 = p[0];
        }
      }
      cv[i][jn][0] = v0;
      cv[i][jn][1] = v1;
    }
  }
  const int KT_ = c0 / BK;
  if (KT_ > 0) {
    double acc[4][2][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jn = 0; jn < 2; ++jn) acc[i][jn][0] = acc[i][jn][1] = 0.0;
    constexpr int STAGE_D = (NB + NB) * BK;
    auto issue = [&](int kt) {
      if (kt < KT_) {
        double* sa = psm + (kt % STAGES) * STAGE_D;
        double* sb = sa + NB * BK;
        load_operand_stage<NB>(sa, Akk, ld, r0, r0 + nrows, (int64_t)kt * BK, c0, tid);
        load_operand_stage<NB>(sb, Akk, ld, c0, w, (int64_t)kt * BK, c0, tid);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) issue(s);
    const int swz = (g & 1) << 2;
    for (int kt = 0; kt < KT_; ++kt) {
      cp_async_wait<STAGES - 2>();
      __syncthreads();
      issue(kt + STAGES - 1);
      const double* sa = psm + (kt % STAGES) * STAGE_D;
      const double* sb = sa + NB * BK;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pc = (h * 4 + t) ^ swz;
        double2 af[4], bf[2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          af[i] = *reinterpret_cast<const double2*>(sa + ((wm0 + i * 8 + g) * 8 + pc) * 2);
#pragma unroll
        for (int jn = 0; jn < 2; ++jn)
          bf[jn] = *reinterpret_cast<const double2*>(sb + ((wn0 + jn * 8 + g) * 8 + pc) * 2);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jn = 0; jn < 2; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], af[i].x, bf[jn].x);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jn = 0; jn < 2; ++jn) dmma884(acc[i][jn][0], acc[i][jn][1], af[i].y, bf[jn].y);
      }
    }
    cp_async_wait<0>();
    __syncthreads();   // the ring is dead from here on; its storage becomes the T / L tiles
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int jn = 0; jn < 2; ++jn) {
        cv[i][jn][0] -= acc[i][jn][0];
        cv[i][jn][1] -= acc[i][jn][1];
      }
  }
  double* Ts = psm;                        // NB x PL_P
  double* Ls = psm + NB * PL_P;            // NB x PL_P, Ls[i][c] = L_jj[i][c]
  double* dinv = Ls + NB * PL_P;           // 1 / L_jj[c][c]
  double* Minv = dinv + NB;                // inverses of the 8 x 8 diagonal blocks of L_jj, pitch PL_MP
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jn = 0; jn < 2; ++jn) {
      const int r = wm0 + i * 8 + g, c = wn0 + jn * 8 + 2 * t;
      *reinterpret_cast<double2*>(Ts + r * PL_P + c) = make_double2(cv[i][jn][0], cv[i][jn][1]);
    }

  PL_T(9);
  // ---- wait for the diagonal block ---------------------------------------------------------
  if (tid == 0) {
    unsigned spins = 0;
    while (*reinterpret_cast<volatile unsigned*>(flag) != epoch) {
      __nanosleep(40);
      if (++spins > (1u << 26)) {          // seconds: something is badly wrong, do not hang the GPU
        atomicExch(info, -1);
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
  PL_T(10);
  const double* Lg = Akk + (int64_t)c0 * ld + c0;
  {
    // Ls[i][c] = L[i][c] (strictly lower part), dinv[c] = 1 / L[c][c].  Every thread touches column tid % 64.
    const int c = tid & (NB - 1);
#pragma unroll 4
    for (int idx = tid; idx < NB * NB; idx += PL_THREADS) {
      const int i = idx / NB;
      Ls[i * PL_P + c] = (i < nbj && c < i) ? __ldcg(Lg + (int64_t)i * ld + c) : 0.0;
    }
    if (tid < NB) dinv[tid] = (tid < nbj) ? 1.0 / __ldcg(Lg + (int64_t)tid * ld + tid) : 1.0;
  }
  __syncthreads();
  // inverses of the eight 8 x 8 diagonal blocks of L (warp b, lane c < 8: column c of block b by forward
  // substitution); everything else of the solve is then DMMA work without a scalar dependency chain
  if (lane < 8) {
    const int o = warp * 8, c = lane;
    double z[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (k < i) sacc = fma(-Ls[(o + i) * PL_P + o + k], (k >= c) ? z[k] : 0.0, sacc);
      z[i] = (i >= c) ? sacc * dinv[o + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Minv[warp * PL_MP * 8 + i * PL_MP + c] = z[i];
  }
  __syncthreads();
  PL_T(11);

  // ---- X = T L^-T by 8-column blocks: X_b = (T_b - sum_{b' < b} X_b' L_bb'^T) L_bb^-T.  Warp w owns rows
  // 8 w .. 8 w + 7 of the tile, so every dependency stays inside the warp (syncwarp only); right-looking: as
  // soon as X_b exists it is subtracted from all later column blocks (independent accumulators) -------------
  {
    double* Tw = Ts + (warp * 8 + g) * PL_P;      // this lane's row of the tile
    double acc[8][2];
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      const double2 v = *reinterpret_cast<const double2*>(Tw + 8 * bb + 2 * t);
      acc[bb][0] = v.x;
      acc[bb][1] = v.y;
    }
#pragma unroll
    for (int bb = 0; bb < 8; ++bb) {
      // R_b from the accumulator layout (row g, columns 2t, 2t+1) to the A-operand layout (row g, k = t)
      __syncwarp();
      *reinterpret_cast<double2*>(Tw + 8 * bb + 2 * t) = make_double2(acc[bb][0], acc[bb][1]);
      __syncwarp();
      const double a0 = Tw[8 * bb + t], a1 = Tw[8 * bb + 4 + t];
      const double* Mi = Minv + bb * PL_MP * 8 + g * PL_MP;
      double x0 = 0.0, x1 = 0.0;
      dmma884(x0, x1, a0, Mi[t]);
      dmma884(x0, x1, a1, Mi[4 + t]);
      __syncwarp();
      *reinterpret_cast<double2*>(Tw + 8 * bb + 2 * t) = make_double2(x0, x1);     // X_b, final
      if (bb < 7) {
        __syncwarp();
        const double na0 = -Tw[8 * bb + t], na1 = -Tw[8 * bb + 4 + t];
#pragma unroll
        for (int b2 = bb + 1; b2 < 8; ++b2) {
          const double* Lr = Ls + (8 * b2 + g) * PL_P + 8 * bb;
          dmma884(acc[b2][0], acc[b2][1], na0, Lr[t]);
          dmma884(acc[b2][0], acc[b2][1], na1, Lr[4 + t]);
        }
      }
    }
  }
  __syncthreads();
  PL_T(12);

  // ---- own rank-64 update of the diagonal block (r, r) of D (prefetch, DMMA, write back) -----------------
  // warps 0..3 take block rows (warp, 7 - warp) of the 8 x 8 grid of 8x8 blocks: 9 lower blocks each.
  if (in_diag && warp < 4) {
    double* Dg = Akk + r0 * ld + r0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int bi = pass ? 7 - warp : warp;
      const int r = bi * 8 + g;
      double d[8][2], acc[8][2];
#pragma unroll
      for (int bj = 0; bj < 8; ++bj) {
        acc[bj][0] = acc[bj][1] = 0.0;
        d[bj][0] = d[bj][1] = 0.0;
        if (bj <= bi && r < nrows) {
          const int c = bj * 8 + 2 * t;
          const double* p = Dg + (int64_t)r * ld + c;
          if (c + 1 <= r) {
            const double2 v = *reinterpret_cast<const double2*>(p);
            d[bj][0] = v.x;
            d[bj][1] = v.y;
          } else if (c <= r) {
            d[bj][0] = p[0];
          }
        }
      }
#pragma unroll 4
      for (int k4 = 0; k4 < NB / 4; ++k4) {
        const double a = Ts[r * PL_P + k4 * 4 + t];
#pragma unroll
        for (int bj = 0; bj < 8; ++bj) {
          if (bj <= bi) {
            const double b = Ts[(bj * 8 + g) * PL_P + k4 * 4 + t];
            dmma884(acc[bj][0], acc[bj][1], a, b);
          }
        }
      }
#pragma unroll
      for (int bj = 0; bj < 8; ++bj) {
        if (bj <= bi && r < nrows) {
          const int c = bj * 8 + 2 * t;
          double* p = Dg + (int64_t)r * ld + c;
          if (c + 1 <= r) *reinterpret_cast<double2*>(p) = make_double2(d[bj][0] - acc[bj][0], d[bj][1] - acc[bj][1]);
          else if (c <= r) p[0] = d[bj][0] - acc[bj][0];
        }
      }
    }
  }

  PL_T(13);
  // ---- store X -----------------------------------------------------------------------------
#pragma unroll 4
  for (int idx = tid; idx < NB * NB; idx += PL_THREADS) {
    const int r = idx / NB, c = idx % NB;
    if (r < nrows && c < nbj) Akk[(r0 + r) * ld + c0 + c] = Ts[r * PL_P + c];
  }
  PL_T(14);
}


// Factorise the w x w (w <= OB) diagonal block at Akk and solve the `below` rows under it (L21 = A21 L11^-T).
static int panel_factor(double* Akk, int64_t w, int64_t ld, int64_t below, int32_t* info, int64_t goff,
                        cudaStream_t st) {
  static TgpPerDeviceOnce attr_once;
  if (tgp_first_use_on_device(attr_once)) {
    TGP_CUDA(cudaFuncSetAttribute(panel_left_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PL_SMEM));
    TGP_CUDA(cudaFuncSetAttribute(panel_left_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                  cudaSharedmemCarveoutMaxShared));
  }
  const int ndiag = (int)tgp_cdiv(w, NB);
  const int64_t nbelow = tgp_cdiv(below, NB);
  for (int j = 0; j < ndiag; ++j) {
    const unsigned epoch = next_flag_epoch();
    const unsigned grid = (unsigned)((ndiag - j) + nbelow);
    panel_left_kernel<<<grid, PL_THREADS, PL_SMEM, st>>>(Akk, ld, (int)w, below, j, info, goff,
                                                         epoch % PL_NFLAGS, epoch);
    TGP_LAUNCH_CHECK();
  }
  return TGP_OK;
}

// ============================================================================================
// Blocked drivers (two levels: OB-wide outer blocks whose updates run on DMMA with Kd = OB,
// NB-wide inner blocks handled by the FP64-ALU kernels above).
// ============================================================================================
constexpr int OB = 512;

// B (M x n) <- B * L^-T for an n x n (n <= OB) lower-triangular L, by recursive halving of the columns:
//   [B1 B2] <- [B1 L11^-T,  (B2 - B1 L21^T) L22^-T]
// so the updates are GEMMs with K = n/2, n/4, ... (large K, each C column block touched once per level)
// and only the 64-wide leaves run the FP64-ALU substitution kernel.
static int trsm_rows_halving(const double* L, int64_t n, int64_t ldl, double* B, int64_t M, int64_t ldb,
                             cudaStream_t st) {
  if (n <= NB) return trsm_panel_launch(L, (int)n, ldl, B, M, ldb, st);
  const int64_t h = ((n / 2 + NB - 1) / NB) * NB;  // split at a multiple of 64
  int rc = trsm_rows_halving(L, h, ldl, B, M, ldb, st);
  if (rc) return rc;
  // B2 -= B1 * L21^T, L21 = L[h:n, 0:h]
  rc = gemm_nt_sub_launch(B + h, M, n - h, ldb, B, ldb, L + h * ldl, ldl, h, 0, st);
  if (rc) return rc;
  return trsm_rows_halving(L + h * ldl + h, n - h, ldl, B + h, M, ldb, st);
}

// B (M x n) <- B * L^-T, L n x n lower.  bs = block size of this level (OB: right-looking over OB-wide
// column blocks with DMMA updates of everything to the right; the OB-wide diagonal solves use halving).
static int trsm_rows_rec(const double* L, int64_t n, int64_t ldl, double* B, int64_t M, int64_t ldb, int bs,
                         cudaStream_t st) {
  if (bs == NB) return trsm_rows_halving(L, n, ldl, B, M, ldb, st);
  for (int64_t k = 0; k < n; k += bs) {
    const int64_t w = (n - k < bs) ? (n - k) : bs;
    const double* Lkk = L + k * ldl + k;
    double* Bk = B + k;
    int rc = trsm_rows_halving(Lkk, w, ldl, Bk, M, ldb, st);
    if (rc) return rc;
    const int64_t rest = n - k - w;
    if (rest > 0) {
      // B[:, k+w:] -= B[:, k:k+w] * L[k+w:, k:k+w]^T
      rc = gemm_nt_sub_launch(B + k + w, M, rest, ldb, Bk, ldb, L + (k + w) * ldl + k, ldl, w, 0, st);
      if (rc) return rc;
    }
  }
  return TGP_OK;
}

// `extra` rows stored below the n x n matrix (right-hand sides as ROWS) are carried through the panel solves and
// trailing updates, so that on exit they hold Y L^-T (= L^-1 y per row): the forward substitution at DMMA speed.
static int potrf_rec(double* A, int64_t n, int64_t ld, int bs, int32_t* info, int64_t goff, cudaStream_t st,
                     int64_t extra = 0) {
  if (g_fused_panel) bs = OB;   // the fused panel kernel handles everything below the OB level
  for (int64_t k = 0; k < n; k += bs) {
    const int64_t w = (n - k < bs) ? (n - k) : bs;
    double* Akk = A + k * ld + k;
    const int64_t rest = n - k - w;
    int rc;
    if (g_fused_panel) {
      rc = panel_factor(Akk, w, ld, rest + extra, info, goff + k, st);
      if (rc) return rc;
    } else {
      if (bs == NB) {
        potf2_kernel<<<1, 256, 0, st>>>(Akk, (int)w, ld, info, goff + k);
        TGP_LAUNCH_CHECK();
        rc = TGP_OK;
      } else {
        rc = potrf_rec(Akk, w, ld, NB, info, goff + k, st);
      }
      if (rc) return rc;
      if (rest + extra > 0) {
        rc = trsm_rows_rec(Akk, w, ld, A + (k + w) * ld + k, rest + extra, ld, NB, st);
        if (rc) return rc;
      }
    }
    if (rest > 0) {
      double* Ark = A + (k + w) * ld + k;  // rows below the diagonal block (+ the extra right-hand-side rows)
      rc = gemm_nt_sub_launch(A + (k + w) * ld + (k + w), rest + extra, rest, ld, Ark, ld, Ark, ld, w, 1, st);
      if (rc) return rc;
    }
  }
  return TGP_OK;
}

// --------------------------------------------------------------------------------------------
// Look-ahead driver for large N.  The OB-wide panel (diagonal factorisation + solve of the rows
// below) is a chain of ~40 small, latency-bound launches; run on its own high-priority stream it
// hides under the previous step's big trailing update:
//   panel stream P : panel(b) .. wait U2(b-1) .. U1(b) = update of the NEXT panel's column block only
//   caller stream U: wait panel(b) .. U2(b) = update of everything right of the next panel
// U1(b) and U2(b) write disjoint blocks; panel(b+1) needs U1(b) and U2(b-1) only, so U2(b) overlaps
// with panel(b+1).  The caller's stream is joined with P before returning.
// --------------------------------------------------------------------------------------------
#include <vector>
// Helper stream + events of the look-ahead schedule: one set per host thread AND device (streams and events
// belong to the device that was current when they were created).
struct LookaheadCtx {
  cudaStream_t panel = nullptr;
  std::vector<cudaEvent_t> ev;
  bool ensure(size_t n) {
    while (ev.size() < n) {
      cudaEvent_t e;
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return false;
      ev.push_back(e);
    }
    return true;
  }
  cudaEvent_t get(size_t i) { return ev[i]; }
};
static LookaheadCtx* lookahead_ctx() {
  static thread_local LookaheadCtx c[TGP_MAX_DEVICES];
  LookaheadCtx& L = c[tgp_current_device()];
  if (!L.panel) {
    int lo = 0, hi = 0;
    if (cudaDeviceGetStreamPriorityRange(&lo, &hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&L.panel, cudaStreamNonBlocking, hi) != cudaSuccess) {
      L.panel = nullptr;
      return nullptr;
    }
  }
  return &L;
}
static int potrf_lookahead(double* A, int64_t n, int64_t ld, int32_t* info, cudaStream_t U, int64_t extra = 0) {
  LookaheadCtx* Lp = lookahead_ctx();
  if (!Lp) {
    tgp_set_error("potrf: could not create the look-ahead stream: %s", cudaGetErrorString(cudaGetLastError()));
    return TGP_ERR_CUDA;
  }
  LookaheadCtx& L = *Lp;
  cudaStream_t P = L.panel;
  // K = 1024 trailing updates amortise the C read-modify-write better once the updates dominate (measured: K = 1024 from N ~ 14k, 2048 from ~ 28k)
  const int64_t OBL = (g_ob_large > 0) ? g_ob_large : (n >= 28000 ? 2048 : (n >= 14000 ? 1024 : 512));
  const int64_t nblk = tgp_cdiv(n, OBL);
  if (!L.ensure((size_t)(3 + 2 * nblk))) {
    tgp_set_error("potrf: cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    return TGP_ERR_CUDA;
  }
  cudaEvent_t e_start = L.get(0);
  TGP_CUDA(cudaEventRecord(e_start, U));
  TGP_CUDA(cudaStreamWaitEvent(P, e_start, 0));
  // event slots: 1 + 2*b = panel(b) done, 2 + 2*b = U2(b) done
  for (int64_t b = 0; b < nblk; ++b) {
    const int64_t k = b * OBL;
    const int64_t w = (n - k < OBL) ? (n - k) : OBL;
    double* Akk = A + k * ld + k;
    int rc = TGP_OK;
    const int64_t rest = n - k - w;
    if (g_fused_panel) {
      // OB-wide sub-panels, each solved against ALL rows below it; the columns of the later sub-panels of this
      // outer block are brought up to date by one K = OB update in between
      for (int64_t s = 0; s < w; s += OB) {
        const int64_t ws = (w - s < OB) ? (w - s) : OB;
        double* Ass = A + (k + s) * ld + (k + s);
        const int64_t below_s = (n - (k + s) - ws) + extra;
        rc = panel_factor(Ass, ws, ld, below_s, info, k + s, P);
        if (rc) return rc;
        const int64_t rem = w - s - ws;
        if (rem > 0) {
          double* Aop = A + (k + s + ws) * ld + (k + s);
          rc = gemm_nt_sub_launch(A + (k + s + ws) * ld + (k + s + ws), below_s, rem, ld, Aop, ld, Aop, ld, ws, 1, P);
          if (rc) return rc;
        }
      }
    } else {
      rc = potrf_rec(Akk, w, ld, w > OB ? OB : NB, info, k, P);
      if (rc) return rc;
      if (rest + extra > 0) {
        double* Ark = A + (k + w) * ld + k;
        rc = trsm_rows_rec(Akk, w, ld, Ark, rest + extra, ld, w > OB ? OB : NB, P);
        if (rc) return rc;
      }
    }
    cudaEvent_t e_panel = L.get(1 + 2 * b);
    TGP_CUDA(cudaEventRecord(e_panel, P));
    if (rest <= 0) {
      TGP_CUDA(cudaStreamWaitEvent(U, e_panel, 0));
      break;
    }
    double* Ark = A + (k + w) * ld + k;
    const int64_t w2 = (rest < OBL) ? rest : OBL;
    TGP_CUDA(cudaStreamWaitEvent(U, e_panel, 0));
    if (b >= 1) TGP_CUDA(cudaStreamWaitEvent(P, L.get(2 + 2 * (b - 1)), 0));
    // U1(b): next panel's column block, all rows below the current panel (+ extra rows)
    rc = gemm_nt_sub_launch(A + (k + w) * ld + (k + w), rest + extra, w2, ld, Ark, ld, Ark, ld, w, 1, P);
    if (rc) return rc;
    // U2(b): everything to the right of the next panel
    const int64_t rest2 = rest - w2;
    if (rest2 > 0) {
      rc = gemm_nt_sub_launch(A + (k + w + w2) * ld + (k + w + w2), rest2 + extra, rest2, ld, Ark + w2 * ld, ld,
                              Ark + w2 * ld, ld, w, 1, U);
      if (rc) return rc;
    }
    TGP_CUDA(cudaEventRecord(L.get(2 + 2 * b), U));
  }
  return TGP_OK;
}

static int check_mat(const double* A, int64_t N, int64_t ld) {
  if (N < 0 || ld < N) return 1;
  if (N > 0 && (!A || (ld & 1) || ((uintptr_t)A & 15))) return 1;
  return 0;
}

static int potrf_with_rows(double* A, int64_t N, int64_t ld, int64_t extra, int32_t* info, cudaStream_t st) {
  TGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  if (N == 0) return TGP_OK;
  if (g_lookahead && N >= 3 * OB) return potrf_lookahead(A, N, ld, info, st, extra);
  return potrf_rec(A, N, ld, N > OB ? OB : NB, info, 0, st, extra);
}

extern "C" int tgp_potrf(double* A, int64_t N, int64_t ld, int32_t* info, void* stream) {
  TGP_CHECK_ARG(!check_mat(A, N, ld), "A must be 16-byte aligned with even ld >= N");
  TGP_CHECK_ARG(info != nullptr, "info");
  return potrf_with_rows(A, N, ld, 0, info, (cudaStream_t)stream);
}

extern "C" int tgp_potrf_rows(double* A, int64_t N, int64_t ld, int64_t nrows, int32_t* info, void* stream) {
  TGP_CHECK_ARG(!check_mat(A, N, ld), "A must be 16-byte aligned with even ld >= N");
  TGP_CHECK_ARG(info != nullptr && nrows >= 0, "info/nrows");
  return potrf_with_rows(A, N, ld, nrows, info, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------------------------
// Envelope (variable-band) Cholesky.  With the points sorted along one axis, a kernel that is below 1e-40 of its
// amplitude beyond a coordinate difference d_cut gives a matrix whose row i starts (to 1e-40) at the first point
// within d_cut of point i; a Cholesky factor keeps the envelope of its matrix.  Per OB-wide block column b the caller
// states row_end[b]: the rows [row_end[b], N) of that block column are (numerically) zero in K and therefore in L, so
// the panel solve and the trailing update stop there: N bw^2 flop instead of N^3 / 3 for a band of bw rows.  Rows
// outside the envelope are neither read nor written (they keep whatever the K build left there: entries below
// 1e-40 amp).  `extra` right-hand-side rows below the matrix are carried through as in potrf_rec.
// --------------------------------------------------------------------------------------------
static int potrf_envelope(double* A, int64_t n, int64_t ld, const int64_t* row_end, int32_t* info, cudaStream_t st,
                          int64_t extra) {
  int64_t prev_end = 0;
  for (int64_t k = 0, b = 0; k < n; k += OB, ++b) {
    const int64_t w = (n - k < OB) ? (n - k) : OB;
    const int64_t c1 = k + w;
    int64_t re = row_end[b];
    if (re < prev_end) re = prev_end;   // an envelope never shrinks from one block column to the next
    if (re < c1) re = c1;
    if (re > n) re = n;
    prev_end = re;
    const int64_t below = re - c1;
    double* Akk = A + k * ld + k;
    double* Ark = A + c1 * ld + k;      // the rows of this block column inside the envelope
    double* Aek = A + n * ld + k;       // the extra rows
    int rc;
    if (g_fused_panel) {
      rc = panel_factor(Akk, w, ld, below, info, k, st);
      if (rc) return rc;
    } else {
      rc = potrf_rec(Akk, w, ld, NB, info, k, st);
      if (rc) return rc;
      if (below > 0) {
        rc = trsm_rows_rec(Akk, w, ld, Ark, below, ld, NB, st);
        if (rc) return rc;
      }
    }
    if (extra > 0) {
      rc = trsm_rows_rec(Akk, w, ld, Aek, extra, ld, NB, st);
      if (rc) return rc;
    }
    if (below > 0) {
      rc = gemm_nt_sub_launch(A + c1 * ld + c1, below, below, ld, Ark, ld, Ark, ld, w, 1, st);
      if (rc) return rc;
      if (extra > 0) {
        rc = gemm_nt_sub_launch(A + n * ld + c1, extra, below, ld, Aek, ld, Ark, ld, w, 0, st);
        if (rc) return rc;
      }
    }
  }
  return TGP_OK;
}

// The same factorisation on two streams, as potrf_lookahead does for the dense matrix: the panel of block column b+1
// (a chain of 8 latency-bound launches) runs on the high-priority stream P while the caller's stream U still applies
// block column b's update to everything right of that panel.
//   P: panel(b) .. wait U2(b-1) .. U1(b) = update of the next panel's columns (rows inside the envelope) .. panel(b+1)
//   U: wait panel(b) .. U2(b) = update of the columns right of the next panel
// U1(b) and U2(b) write disjoint column ranges; U1(b+1) is ordered after U2(b) (both touch the columns right of panel
// b+2's left edge).  No extra right-hand-side rows here (tgp_loglike_env uses the sweeps).
static int potrf_envelope_lookahead(double* A, int64_t n, int64_t ld, const int64_t* row_end, int32_t* info,
                                    cudaStream_t U) {
  LookaheadCtx* Lp = lookahead_ctx();
  if (!Lp) {
    tgp_set_error("potrf_env: could not create the look-ahead stream: %s", cudaGetErrorString(cudaGetLastError()));
    return TGP_ERR_CUDA;
  }
  LookaheadCtx& L = *Lp;
  cudaStream_t P = L.panel;
  const int64_t nblk = tgp_cdiv(n, (int64_t)OB);
  if (!L.ensure((size_t)(4 + 2 * nblk))) {
    tgp_set_error("potrf_env: cudaEventCreate failed: %s", cudaGetErrorString(cudaGetLastError()));
    return TGP_ERR_CUDA;
  }
  TGP_CUDA(cudaEventRecord(L.get(0), U));
  TGP_CUDA(cudaStreamWaitEvent(P, L.get(0), 0));
  int64_t prev_end = 0;
  // event slots: 1 + 2 b = panel(b) done (on P), 2 + 2 b = U2(b) done (on U)
  for (int64_t b = 0; b < nblk; ++b) {
    const int64_t k = b * OB;
    const int64_t w = (n - k < OB) ? (n - k) : OB;
    const int64_t c1 = k + w;
    int64_t re = row_end[b];
    if (re < prev_end) re = prev_end;
    if (re < c1) re = c1;
    if (re > n) re = n;
    prev_end = re;
    const int64_t below = re - c1;
    double* Akk = A + k * ld + k;
    double* Ark = A + c1 * ld + k;
    int rc = panel_factor(Akk, w, ld, below, info, k, P);
    if (rc) return rc;
    cudaEvent_t e_panel = L.get(1 + 2 * b);
    // U2(b) is released right after panel(b), beside U1(b).  Option value 2 releases it after U1(b) instead (U1 then
    // has the SMs to itself and the next panel starts earlier): measured 33.1 ms against 31.8 ms at N = 40k -- the
    // panel CTAs (134 KB of shared memory) cannot become resident beside U2's CTAs anyway, so a later U2 only ends later.
    const bool u2_beside_u1 = (g_env_lookahead != 2);
    if (u2_beside_u1) {
      TGP_CUDA(cudaEventRecord(e_panel, P));
      TGP_CUDA(cudaStreamWaitEvent(U, e_panel, 0));
    }
    // everything the next panel reads must be final: the U2 updates up to block b-1 (stream order on U covers the
    // earlier ones)
    if (b >= 1) TGP_CUDA(cudaStreamWaitEvent(P, L.get(2 + 2 * (b - 1)), 0));
    if (below > 0) {
      const int64_t w2 = (below < OB) ? below : OB;
      rc = gemm_nt_sub_launch(A + c1 * ld + c1, below, w2, ld, Ark, ld, Ark, ld, w, 1, P);            // U1(b)
      if (rc) return rc;
    }
    if (!u2_beside_u1) {
      TGP_CUDA(cudaEventRecord(e_panel, P));
      TGP_CUDA(cudaStreamWaitEvent(U, e_panel, 0));
    }
    if (below > 0) {
      const int64_t w2 = (below < OB) ? below : OB;
      const int64_t rest2 = below - w2;
      if (rest2 > 0) {
        rc = gemm_nt_sub_launch(A + (c1 + w2) * ld + (c1 + w2), rest2, rest2, ld, Ark + w2 * ld, ld, Ark + w2 * ld,
                                ld, w, 1, U);                                                         // U2(b)
        if (rc) return rc;
      }
    }
    TGP_CUDA(cudaEventRecord(L.get(2 + 2 * b), U));
  }
  // join: the caller's stream continues after the last panel-stream work
  TGP_CUDA(cudaEventRecord(L.get(3 + 2 * nblk), P));
  TGP_CUDA(cudaStreamWaitEvent(U, L.get(3 + 2 * nblk), 0));
  return TGP_OK;
}

static int check_envelope(const int64_t* row_end, int64_t nblocks, int64_t N) {
  return row_end == nullptr || nblocks != tgp_cdiv(N, (int64_t)OB);
}

extern "C" int tgp_envelope_block(void) { return OB; }

extern "C" int tgp_potrf_env(double* A, int64_t N, int64_t ld, const int64_t* row_end, int64_t nblocks,
                             int64_t nrows, int32_t* info, void* stream) {
  TGP_CHECK_ARG(!check_mat(A, N, ld), "A must be 16-byte aligned with even ld >= N");
  TGP_CHECK_ARG(info != nullptr && nrows >= 0, "info/nrows");
  TGP_CHECK_ARG(!check_envelope(row_end, nblocks, N), "row_end must hold one entry per tgp_envelope_block() columns");
  cudaStream_t st = (cudaStream_t)stream;
  TGP_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  if (N == 0) return TGP_OK;
  if (g_env_lookahead && g_fused_panel && nrows == 0 && N >= 3 * OB)
    return potrf_envelope_lookahead(A, N, ld, row_end, info, st);
  return potrf_envelope(A, N, ld, row_end, info, st, nrows);
}

// B <- B L^-T for a factor with the envelope row_end: after the block of unknowns [k, c1) is solved it only enters
// the unknowns [c1, row_end[b]) -- 2 M N bw flop instead of M N^2.
static int trsm_rows_envelope(const double* L, int64_t n, int64_t ld, const int64_t* row_end, double* B, int64_t M,
                              int64_t ldb, cudaStream_t st) {
  int64_t prev_end = 0;
  for (int64_t k = 0, b = 0; k < n; k += OB, ++b) {
    const int64_t w = (n - k < OB) ? (n - k) : OB;
    const int64_t c1 = k + w;
    int64_t re = row_end[b];
    if (re < prev_end) re = prev_end;
    if (re < c1) re = c1;
    if (re > n) re = n;
    prev_end = re;
    int rc = trsm_rows_rec(L + k * ld + k, w, ld, B + k, M, ldb, NB, st);
    if (rc) return rc;
    if (re > c1) {
      rc = gemm_nt_sub_launch(B + c1, M, re - c1, ldb, B + k, ldb, L + c1 * ld + k, ld, w, 0, st);
      if (rc) return rc;
    }
  }
  return TGP_OK;
}

extern "C" int tgp_trsm_rows_env(const double* L, int64_t N, int64_t ld, const int64_t* row_end, int64_t nblocks,
                                 double* B, int64_t M, int64_t ldb, void* stream) {
  TGP_CHECK_ARG(!check_mat(L, N, ld), "L must be 16-byte aligned with even ld >= N");
  TGP_CHECK_ARG(M >= 0 && ldb >= N, "M/ldb");
  TGP_CHECK_ARG(!check_envelope(row_end, nblocks, N), "row_end must hold one entry per tgp_envelope_block() columns");
  if (M == 0 || N == 0) return TGP_OK;
  TGP_CHECK_ARG(B && (ldb % 2 == 0) && ((uintptr_t)B % 16 == 0), "B must be 16-byte aligned with even ldb");
  return trsm_rows_envelope(L, N, ld, row_end, B, M, ldb, (cudaStream_t)stream);
}

extern "C" int tgp_trsm_rows(const double* L, int64_t N, int64_t ld, double* B, int64_t M, int64_t ldb,
                             void* stream) {
  TGP_CHECK_ARG(!check_mat(L, N, ld), "L must be 16-byte aligned with even ld >= N");
  TGP_CHECK_ARG(M >= 0 && ldb >= N, "M/ldb");
  if (M == 0 || N == 0) return TGP_OK;
  TGP_CHECK_ARG(B && (ldb % 2 == 0) && ((uintptr_t)B % 16 == 0), "B must be 16-byte aligned with even ldb");
  return trsm_rows_rec(L, N, ld, B, M, ldb, N > OB ? OB : NB, (cudaStream_t)stream);
}

extern "C" int tgp_gemm_nt_sub(double* C, int64_t M, int64_t Nc, int64_t ldc, const double* A, int64_t lda,
                               const double* B, int64_t ldb, int64_t Kd, int lower_only, void* stream) {
  TGP_CHECK_ARG(M >= 0 && Nc >= 0 && Kd >= 0 && ldc >= Nc && lda >= Kd && ldb >= Kd, "shape");
  if (M == 0 || Nc == 0 || Kd == 0) return TGP_OK;
  TGP_CHECK_ARG(C && A && B, "null pointer");
  return gemm_nt_sub_launch(C, M, Nc, ldc, A, lda, B, ldb, Kd, lower_only, (cudaStream_t)stream);
}

// ============================================================================================
// Single right-hand-side solves  L w = b  (forward) and  L^T x = w  (backward): persistent sweep kernels in trsv.cu.
// ============================================================================================
int tgp_trsv_sweeps(const double* L, int64_t N, int64_t ld, double* b, int which, cudaStream_t st);
static int trsv_backward(const double* L, int64_t N, int64_t ld, double* b, cudaStream_t st) {
  return tgp_trsv_sweeps(L, N, ld, b, 2, st);
}

extern "C" int tgp_potrs_vec(const double* L, int64_t N, int64_t ld, double* b, void* stream) {
  TGP_CHECK_ARG(N >= 0 && ld >= N, "N/ld");
  if (N == 0) return TGP_OK;
  TGP_CHECK_ARG(L && b, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  return tgp_trsv_sweeps(L, N, ld, b, 3, st);   // forward, then backward (one workspace, one set of inverses)
}

// ============================================================================================
// Reductions: log-determinant, chi2 = y . alpha, final log-likelihood.
// ============================================================================================
// out[0] = sum 2 log L_ii ; out[1] = y . alpha (if given).  Single CTA, deterministic order.
__global__ void __launch_bounds__(1024)
logdet_chi2_kernel(const double* __restrict__ L, int64_t N, int64_t ld, const double* __restrict__ y,
                   const double* __restrict__ alpha, double* __restrict__ out) {
  __shared__ double s0[32], s1[32];
  double a = 0.0, c = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
    a += 2.0 * log(L[i * ld + i]);
    if (y) c = fma(y[i], alpha[i], c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s0[warp] = a; s1[warp] = c; }
  __syncthreads();
  if (warp == 0) {
    a = s0[lane];
    c = s1[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) {
      out[0] = a;
      if (y) out[1] = c;
    }
  }
}

extern "C" int tgp_logdet_chi2(const double* L, int64_t N, int64_t ld, const double* y, const double* alpha,
                               double* out, void* stream) {
  TGP_CHECK_ARG(N >= 0 && ld >= N && out, "N/ld/out");
  TGP_CHECK_ARG((y == nullptr) == (alpha == nullptr), "y and alpha go together");
  logdet_chi2_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(L, N, ld, y, alpha, out);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

// out = {logL, chi2, logdet}; tmp = {logdet, chi2}.
__global__ void loglike_finish_kernel(const double* tmp, int64_t N, const int32_t* __restrict__ info, double* out) {
  const double logdet = tmp[0], chi2 = tmp[1];
  double ll = -0.5 * chi2 - 0.5 * (double)N * log(2.0 * TGP_PI) - 0.5 * logdet;
  // info > 0: leading minor not positive definite -> -inf (log_likelihood.py:38-39).  info < 0: an internal
  // flag wait timed out -- NOT a property of the matrix: NaN here, and the host raises when it sees info < 0.
  if (*info > 0 || isnan(ll)) ll = -INFINITY;
  if (*info < 0) ll = NAN;
  out[0] = ll;
  out[1] = chi2;
  out[2] = logdet;
}

// chi2 = ||w||^2 with w = L^-1 y (forward sweep only) -- y^T K^-1 y without the backward solve.
__global__ void __launch_bounds__(1024)
sumsq_kernel(const double* __restrict__ w, int64_t N, double* __restrict__ out1) {
  __shared__ double s0[32];
  double a = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += blockDim.x) a = fma(w[i], w[i], a);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s0[warp] = a;
  __syncthreads();
  if (warp == 0) {
    a = s0[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) *out1 = a;
  }
}

static int loglike_impl(const double* X, const double* y, const double* yerr2, int64_t N,
                        const tgp_kernel* k, double* work, int64_t ld, double* alpha, int want_alpha,
                        double* out, int32_t* info, const int64_t* row_end, int64_t nblocks, void* stream) {
  TGP_CHECK_ARG(N > 0 && X && y && work && alpha && out && info, "null pointer / N");
  TGP_CHECK_ARG(row_end == nullptr || !check_envelope(row_end, nblocks, N),
                "row_end must hold one entry per tgp_envelope_block() columns");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = tgp_kmat_sym(X, N, k, yerr2, work, ld, /*lower_only=*/1, stream);
  if (rc) return rc;
  // out[1], out[2] double as scratch {logdet, chi2} until the finishing kernel reorders them
  double* tmp = out + 1;
  if (row_end) {
    // Envelope: a row riding through the factorisation would cost two more (one-row) launches chains per block column;
    // the persistent sweep kernels read the factor once instead (w = L^-1 y, and alpha = L^-T w in the same call).
    rc = tgp_potrf_env(work, N, ld, row_end, nblocks, 0, info, stream);
    if (rc) return rc;
    TGP_CUDA(cudaMemcpyAsync(alpha, y, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    rc = tgp_trsv_sweeps(work, N, ld, alpha, want_alpha ? 3 : 1, st);
    if (rc) return rc;
  } else {
    // y rides along as row N of the workspace: the factorisation's panel solves turn it into w = L^-1 y
    double* wrow = work + N * ld;
    TGP_CUDA(cudaMemcpyAsync(wrow, y, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    rc = tgp_potrf_rows(work, N, ld, 1, info, stream);
    if (rc) return rc;
    TGP_CUDA(cudaMemcpyAsync(alpha, wrow, N * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (want_alpha) {
      rc = trsv_backward(work, N, ld, alpha, st);   // alpha = L^-T w
      if (rc) return rc;
    }
  }
  if (want_alpha) {
    logdet_chi2_kernel<<<1, 1024, 0, st>>>(work, N, ld, y, alpha, tmp);  // chi2 = y . alpha
    TGP_LAUNCH_CHECK();
  } else {
    // y^T K^-1 y = ||L^-1 y||^2: the backward sweep is not needed for the likelihood alone
    logdet_chi2_kernel<<<1, 1024, 0, st>>>(work, N, ld, nullptr, nullptr, tmp);
    TGP_LAUNCH_CHECK();
    sumsq_kernel<<<1, 1024, 0, st>>>(alpha, N, tmp + 1);
    TGP_LAUNCH_CHECK();
  }
  loglike_finish_kernel<<<1, 1, 0, st>>>(tmp, N, info, out);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}

extern "C" int tgp_loglike(const double* X, const double* y, const double* yerr2, int64_t N,
                           const tgp_kernel* k, double* work, int64_t ld, double* alpha, int want_alpha,
                           double* out, int32_t* info, void* stream) {
  return loglike_impl(X, y, yerr2, N, k, work, ld, alpha, want_alpha, out, info, nullptr, 0, stream);
}

extern "C" int tgp_loglike_env(const double* X, const double* y, const double* yerr2, int64_t N,
                               const tgp_kernel* k, double* work, int64_t ld, double* alpha, int want_alpha,
                               double* out, int32_t* info, const int64_t* row_end, int64_t nblocks, void* stream) {
  TGP_CHECK_ARG(row_end != nullptr, "row_end");
  return loglike_impl(X, y, yerr2, N, k, work, ld, alpha, want_alpha, out, info, row_end, nblocks, stream);
}
