// Single right-hand-side triangular solves  L w = b  (forward) and  L^T x = w  (backward): the alpha solve of
// /root/reference/treegp/gp_interp.py:182 and log_likelihood.py:31 (scipy cho_solve -> LAPACK dpotrs).
//
// The operation is HBM-bound: each sweep reads the 4 N^2-byte triangle once.  ONE
// persistent kernel per sweep, one CTA per SM, launched cooperatively (all CTAs co-resident, so waiting on
// another CTA's flag always makes progress -- no assumption about dispatch order):
//
//  * the triangle is cut into 64 x 64 tiles; block row k ("slab") is owned by CTA k mod G, which keeps the slab's
//    64 partial sums in shared memory and accumulates them in a FIXED order (column blocks ascending): the
//    result is deterministic, bit-identical from run to run;
//  * a CTA walks the column blocks j = 0, 1, ... and, for each, its slabs k > j.  The tiles of this sequence are
//    streamed into a 5-stage shared-memory ring by bulk async copies (cp.async.bulk, one 512-byte row segment per
//    thread, completion on an mbarrier) issued 4 tiles ahead: the tile data do not depend on the unknowns, so HBM
//    keeps streaming while a CTA waits for a dependency;
//  * the only serial chain: when x_j is published, the owner of slab j + 1 applies its last tile and multiplies by
//    the INVERSE of the 64 x 64 diagonal block (computed beforehand for all blocks by trsv_diag_inv_kernel,
//    prefetched into shared memory) -- two 64 x 64 mat-vecs per 64 unknowns instead of a 64-step substitution.
//    Publication is the payload itself: the 64 unknowns are stored into a buffer pre-filled with an all-ones
//    pattern and every waiting thread polls its own word, i.e. ONE L2 round trip per chain step (no flag + fence
//    + second load).
//
// The backward sweep is the same kernel on the mirrored index set (block i' = nb-1-i) with transposed tile
// products.  Workspace (inverse diagonal blocks + publication buffers) comes from the stream-ordered allocator and is freed
// on the stream: nothing persists.  A wait that times out (cannot happen with co-resident CTAs; guards
// against a lost launch) sets a sticky device error word that tgp_device_error() reports.
#include <cuda_runtime.h>
#include <stdint.h>
#include "tgp_common.cuh"

constexpr int TS = 64;                 // tile edge = diagonal block = unknowns per chain step
constexpr int TS_P = TS + 2;           // shared-memory row pitch in doubles (528 B: 16-byte aligned rows)
constexpr int TS_STAGES = 5;
constexpr int TS_THREADS = 256;
constexpr int TS_TILE_D = TS * TS_P;   // doubles per staged tile

__device__ int g_tgp_device_error = 0;

extern "C" int tgp_device_error(int reset) {
  int v = 0;
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  if (cudaMemcpyFromSymbol(&v, g_tgp_device_error, sizeof(int)) != cudaSuccess) return -1;
  if (reset && v) {
    int z = 0;
    cudaMemcpyToSymbol(g_tgp_device_error, &z, sizeof(int));
  }
  return v;
}

// ---- mbarrier / bulk-copy primitives (PTX) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_addr(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
__device__ __forceinline__ double ld_relaxed_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// ---- inverse of every 64 x 64 diagonal block ---------------------------------------------------
// One CTA (64 threads) per block; thread c solves L z = e_c by forward substitution from a shared-memory copy.
// Output block: dense 64 x 64, row-major, zero above the diagonal, identity in the padding of a short last block.
__global__ void __launch_bounds__(TS)
trsv_diag_inv_kernel(const double* __restrict__ L, int64_t ld, int64_t N, double* __restrict__ dinv) {
  extern __shared__ __align__(16) double dism[];
  double (*Ls)[TS + 1] = reinterpret_cast<double (*)[TS + 1]>(dism);
  double (*Zs)[TS + 1] = reinterpret_cast<double (*)[TS + 1]>(dism + TS * (TS + 1));
  const int64_t k0 = (int64_t)blockIdx.x * TS;
  const int w = (int)((N - k0 < TS) ? (N - k0) : TS);
  const int c = threadIdx.x;
  for (int r = 0; r < TS; ++r) {
    double v = (r == c) ? 1.0 : 0.0;
    if (r < w && c <= r) v = L[(k0 + r) * ld + k0 + c];
    Ls[r][c] = v;
  }
  __syncthreads();
  // column c of the inverse: z_r = (delta_rc - sum_{m<r} L_rm z_m) / L_rr
  for (int r = 0; r < TS; ++r) {
    double s0 = (r == c) ? 1.0 : 0.0, s1 = 0.0;
    int m = c;
    for (; m + 1 < r; m += 2) {
      s0 = fma(-Ls[r][m], Zs[m][c], s0);
      s1 = fma(-Ls[r][m + 1], Zs[m + 1][c], s1);
    }
    if (m < r) s0 = fma(-Ls[r][m], Zs[m][c], s0);
    Zs[r][c] = (r >= c) ? (s0 + s1) / Ls[r][r] : 0.0;
  }
  __syncthreads();
  double* out = dinv + (int64_t)blockIdx.x * TS * TS;
  for (int r = 0; r < TS; ++r) out[r * TS + c] = Zs[r][c];
}

// ---- the sweep ---------------------------------------------------------------------------------
struct TrsvParams {
  const double* L;
  int64_t ld, N;
  double* b;             // right-hand side in, unknowns out
  const double* dinv;    // nb blocks of 64 x 64
  double* xpub;          // nb * 64 words, all-ones on entry: block kp's unknowns are published at xpub + 64 kp
  int nb, G, aligned;
};

// position in a CTA's tile sequence: column block j (primed index), slab index m among the CTA's slabs
struct TileIter {
  int j, m, m0;          // m0 = first slab of this CTA with k > j
  bool done;
};

template <bool FWD>
__global__ void __launch_bounds__(TS_THREADS, 1)
trsv_sweep_kernel(TrsvParams P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  double* ring = reinterpret_cast<double*>(smraw);                 // TS_STAGES tiles
  double* dtile = ring + TS_STAGES * TS_TILE_D;                    // inverse diagonal block of the next solve
  double* xs = dtile + TS_TILE_D;                                  // unknowns of the current column block
  double* vs = xs + TS;                                            // right-hand side of the current solve
  double* bs = vs + TS;                                            // b of the next solve (prefetched)
  double* red = bs + TS;                                           // 4 x 64 partials (backward)
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 4 * TS);      // TS_STAGES + 1 barriers
  double* acc = reinterpret_cast<double*>(full + TS_STAGES + 2);   // [slabs of this CTA][64]

  const int tid = threadIdx.x;
  const int c = blockIdx.x, G = P.G, nb = P.nb;
  const int M = (nb - c + G - 1) / G;                              // slabs owned: k = c + m G (primed indices)
  const int64_t ld = P.ld, N = P.N;

  // primed block index -> first row / column of the block in the matrix, and its height
  auto blk0 = [&](int kp) -> int64_t { return (int64_t)(FWD ? kp : nb - 1 - kp) * TS; };
  auto blkw = [&](int kp) -> int { const int64_t r0 = blk0(kp); return (int)((N - r0 < TS) ? (N - r0) : TS); };
  // tile of slab kp against column block jp (jp < kp): forward L[blk kp, blk jp]; backward L[blk jp', blk kp']
  auto tile_src = [&](int kp, int jp) -> const double* {
    return FWD ? P.L + blk0(kp) * ld + blk0(jp) : P.L + blk0(jp) * ld + blk0(kp);
  };
  auto tile_rows = [&](int kp, int jp) -> int { return FWD ? blkw(kp) : blkw(jp); };

  if (tid == 0) {
    for (int s = 0; s <= TS_STAGES; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < M * TS; i += TS_THREADS) acc[i] = 0.0;
  __syncthreads();

  auto advance = [&](TileIter& it) {
    if (it.done) return;
    ++it.m;
    if (it.m >= M) {
      ++it.j;
      while (it.m0 < M && c + it.m0 * G <= it.j) ++it.m0;
      it.m = it.m0;
      if (it.m0 >= M || it.j >= nb - 1) it.done = true;
    }
  };
  auto start = [&]() {
    TileIter it;
    it.j = 0;
    it.m0 = (c == 0) ? 1 : 0;       // slab 0 has no tiles
    it.m = it.m0;
    it.done = (it.m0 >= M) || nb < 2;
    return it;
  };
  // issue the bulk copies of one tile into ring stage `s` (threads 0..63: one row each)
  auto issue = [&](const TileIter& it, int s) {
    if (!P.aligned) return;
    const int kp = c + it.m * G;
    const int rows = tile_rows(kp, it.j);
    if (tid == 0) mbar_expect_tx(&full[s], (uint32_t)rows * TS * 8);
    if (tid < rows) bulk_g2s(ring + s * TS_TILE_D + tid * TS_P, tile_src(kp, it.j) + (int64_t)tid * ld, TS * 8, &full[s]);
  };
  // prefetch what the next solve (slab index ms) needs: inverse diagonal block (bulk) and its right-hand side
  auto prefetch_solve = [&](int ms) {
    if (ms >= M) return;
    const int kp = c + ms * G;
    const int64_t r0 = blk0(kp);
    const int w = blkw(kp);
    const double* src = P.dinv + (int64_t)(FWD ? kp : nb - 1 - kp) * TS * TS;
    if (tid == 0) mbar_expect_tx(&full[TS_STAGES], TS * TS * 8);
    if (tid < TS) {
      bulk_g2s(dtile + tid * TS_P, src + tid * TS, TS * 8, &full[TS_STAGES]);
      bs[tid] = (tid < w) ? P.b[r0 + tid] : 0.0;
    }
  };

  // Product of the staged 64 x 64 tile `T` (pitch TS_P) with `xvec`: forward sum_c T[r][c] x[c] for row r, backward
  // sum_r T[r][c] x[r] for column c.  The result is returned in the register of the element's OWNER thread
  // (forward: the q == 0 lane of row r = tid / 4; backward: thread tid < 64 for column tid); `owner`/`idx` say which.
  const bool owner = FWD ? ((tid & 3) == 0) : (tid < TS);
  const int idx = FWD ? (tid >> 2) : tid;
  auto matvec = [&](const double* T, const double* xvec, int rows_valid) -> double {
    if (FWD) {
      const int r = tid >> 2, q = tid & 3;
      const double* row = T + r * TS_P + q * 16;
      const double* xv = xvec + q * 16;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const int c0 = (i + (q & 1)) & 15, c1 = (i + 1 + (q & 1)) & 15;   // rotation: conflict-free LDS.64
        const int c2 = (i + 2 + (q & 1)) & 15, c3 = (i + 3 + (q & 1)) & 15;
        s0 = fma(row[c0], xv[c0], s0);
        s1 = fma(row[c1], xv[c1], s1);
        s2 = fma(row[c2], xv[c2], s2);
        s3 = fma(row[c3], xv[c3], s3);
      }
      double sv = (s0 + s1) + (s2 + s3);
      sv += __shfl_xor_sync(0xffffffffu, sv, 1);
      sv += __shfl_xor_sync(0xffffffffu, sv, 2);
      return sv;
    } else {
      const int cc = tid & 63, q = tid >> 6;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const int r0 = q * 16 + i;
        if (r0 < rows_valid) s0 = fma(T[r0 * TS_P + cc], xvec[r0], s0);
        if (r0 + 1 < rows_valid) s1 = fma(T[(r0 + 1) * TS_P + cc], xvec[r0 + 1], s1);
        if (r0 + 2 < rows_valid) s2 = fma(T[(r0 + 2) * TS_P + cc], xvec[r0 + 2], s2);
        if (r0 + 3 < rows_valid) s3 = fma(T[(r0 + 3) * TS_P + cc], xvec[r0 + 3], s3);
      }
      red[q * TS + cc] = (s0 + s1) + (s2 + s3);
      __syncthreads();
      return (tid < TS) ? (red[tid] + red[TS + tid]) + (red[2 * TS + tid] + red[3 * TS + tid]) : 0.0;
    }
  };

  uint32_t dphase = 0;
  // Solve slab index ms given the finished partial sums `accv` (owner threads): x = Dinv (b - acc).  The unknowns
  // are PUBLISHED as the payload itself: every waiting thread polls its own word of xpub until it differs from
  // the all-ones pattern the workspace was filled with -- one L2 round trip per chain step, no flag, no fence.
  auto solve = [&](int ms, double accv) {
    const int kp = c + ms * G;
    const int64_t r0 = blk0(kp);
    const int w = blkw(kp);
    if (owner) vs[idx] = (idx < w) ? bs[idx] - accv : 0.0;
    mbar_wait(&full[TS_STAGES], dphase);
    dphase ^= 1;
    __syncthreads();
    double xv = matvec(dtile, vs, TS);
    if (owner) {
      if (__double_as_longlong(xv) == -1ll) xv = __longlong_as_double(0x7ff8000000000000ll);  // never publish the sentinel
      st_relaxed_f64(P.xpub + (int64_t)kp * TS + idx, xv);
      if (idx < w) P.b[r0 + idx] = xv;
    }
    __syncthreads();                               // dtile, bs, vs are free again
    prefetch_solve(ms + 1);
  };

  // ---- prologue: prefetch the first solve and the first tiles ------------------------------------
  prefetch_solve(0);
  TileIter pf = start();
  int pf_stage = 0;
  for (int s = 0; s < TS_STAGES - 1; ++s) {
    if (!pf.done) {
      issue(pf, pf_stage);
      advance(pf);
      pf_stage = (pf_stage + 1) % TS_STAGES;
    }
  }
  int ms_next = 0;                                // next slab (index) this CTA has to solve
  if (c == 0) {
    __syncthreads();
    solve(0, 0.0);
    ms_next = 1;
  }

  TileIter it = start();
  int stage = 0;
  uint32_t phase = 0;                              // parity of ring stage 0's current fill
  int cur_j = -1;
  while (!it.done) {
    if (it.j != cur_j) {
      // ---- new column block: wait until its unknowns are published, bring them into shared memory ----
      cur_j = it.j;
      __syncthreads();                             // everyone is done with the previous xs
      // Only the CTA that owns slab j + 1 is on the serial chain: its 64 threads poll their own words back to
      // back.  Every other CTA has streaming work queued behind this column and polls lazily with ONE thread
      // (thousands of threads hammering the same 512 bytes would delay the very store they wait for).
      const bool urgent = (c + it.m * G == cur_j + 1);
      const double* src = P.xpub + (int64_t)cur_j * TS;
      if (!urgent) {
        if (tid == 0) {
          unsigned spins = 0;
          while (__double_as_longlong(ld_relaxed_f64(src)) == -1ll) {
            __nanosleep(spins < 8u ? 100 : 400);
            if (++spins > (1u << 23)) { atomicExch(&g_tgp_device_error, 1); break; }
          }
        }
        __syncthreads();
      }
      if (tid < TS) {
        double v = ld_relaxed_f64(src + tid);
        unsigned spins = 0;
        while (__double_as_longlong(v) == -1ll) {
          if (++spins > (1u << 25)) { atomicExch(&g_tgp_device_error, 1); break; }
          v = ld_relaxed_f64(src + tid);
        }
        xs[tid] = v;
      }
      __syncthreads();
    }
    const int kp = c + it.m * G;
    const int rows = tile_rows(kp, it.j);
    const double* T = ring + stage * TS_TILE_D;
    if (P.aligned) {
      mbar_wait(&full[stage], phase);
    } else {
      // unaligned matrix (odd ld / base): plain loads, no prefetch
      const double* src = tile_src(kp, it.j);
      for (int i2 = tid; i2 < TS * TS; i2 += TS_THREADS) {
        const int r = i2 >> 6, cc = i2 & 63;
        ring[stage * TS_TILE_D + r * TS_P + cc] = (r < rows) ? src[(int64_t)r * ld + cc] : 0.0;
      }
      __syncthreads();
    }
    const double sv = matvec(T, xs, rows);
    const double accv = owner ? acc[it.m * TS + idx] + sv : 0.0;
    const bool solve_now = (kp == it.j + 1) && (it.m == ms_next);   // that was the slab's last tile
    if (!solve_now) {
      if (owner) acc[it.m * TS + idx] = accv;
      __syncthreads();                             // the ring stage is free again
    }
    if (solve_now) {
      solve(ms_next, accv);                        // (its first __syncthreads also frees the ring stage)
      ++ms_next;
    }
    if (!pf.done) {                                // refill the ring, TS_STAGES - 1 tiles ahead
      issue(pf, pf_stage);
      advance(pf);
      pf_stage = (pf_stage + 1) % TS_STAGES;
    }
    stage = (stage + 1) % TS_STAGES;
    if (stage == 0) phase ^= 1;
    advance(it);
  }
}

static size_t trsv_smem_bytes(int slabs_per_cta) {
  return (size_t)(TS_STAGES + 1) * TS_TILE_D * 8 + (size_t)(3 * TS + 4 * TS) * 8 + (TS_STAGES + 2) * 8 +
         (size_t)slabs_per_cta * TS * 8;
}

// Workspace layout: [nb * 64 * 64 doubles: inverse diagonal blocks][2 * nb * 64 doubles: published unknowns of the
// forward and of the backward sweep, filled with 0xFF bytes = "not yet published"]
template <bool FWD>
static int trsv_launch(const TrsvParams& P0, cudaStream_t st) {
  TrsvParams P = P0;
  const size_t smem = trsv_smem_bytes((P.nb + P.G - 1) / P.G);
  static TgpPerDeviceOnce once;
  if (tgp_first_use_on_device(once))
    TGP_CUDA(cudaFuncSetAttribute(trsv_sweep_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  TGP_CHECK_ARG(smem <= 227 * 1024, "matrix too large for the sweep kernel's per-CTA accumulators");
  void* args[] = {(void*)&P};
  TGP_CUDA(cudaLaunchCooperativeKernel((void*)trsv_sweep_kernel<FWD>, dim3((unsigned)P.G), dim3(TS_THREADS), args, smem, st));
  return TGP_OK;
}

// which = 1: forward only, 2: backward only, 3: both (forward then backward)
int tgp_trsv_sweeps(const double* L, int64_t N, int64_t ld, double* b, int which, cudaStream_t st) {
  if (N <= 0) return TGP_OK;
  const int nb = (int)tgp_cdiv(N, TS);
  const int sms = tgp_num_sms();
  TrsvParams P;
  P.L = L; P.ld = ld; P.N = N; P.b = b; P.nb = nb;
  P.G = nb < sms ? nb : sms;
  P.aligned = ((ld & 1) == 0) && ((((uintptr_t)L) & 15) == 0);
  const size_t dinv_bytes = (size_t)nb * TS * TS * 8;
  const size_t flag_bytes = (size_t)2 * nb * TS * sizeof(double);
  void* ws = nullptr;
  TGP_CUDA(cudaMallocAsync(&ws, dinv_bytes + flag_bytes, st));
  double* dinv = reinterpret_cast<double*>(ws);
  double* xpub = dinv + (size_t)nb * TS * TS;
  int rc = TGP_OK;
  do {
    if (cudaMemsetAsync(xpub, 0xFF, flag_bytes, st) != cudaSuccess) { rc = TGP_ERR_CUDA; break; }
    static TgpPerDeviceOnce inv_once;
    constexpr int inv_smem = 2 * TS * (TS + 1) * 8;
    if (tgp_first_use_on_device(inv_once) &&
        cudaFuncSetAttribute(trsv_diag_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, inv_smem) != cudaSuccess) {
      rc = TGP_ERR_CUDA;
      break;
    }
    trsv_diag_inv_kernel<<<(unsigned)nb, TS, inv_smem, st>>>(L, ld, N, dinv);
    if (cudaGetLastError() != cudaSuccess) { rc = TGP_ERR_CUDA; break; }
    P.dinv = dinv;
    if (which & 1) {
      P.xpub = xpub;
      rc = trsv_launch<true>(P, st);
      if (rc) break;
    }
    if (which & 2) {
      P.xpub = xpub + (size_t)nb * TS;
      rc = trsv_launch<false>(P, st);
      if (rc) break;
    }
  } while (0);
  if (rc == TGP_ERR_CUDA) tgp_set_error("tgp_trsv_sweeps: %s", cudaGetErrorString(cudaGetLastError()));
  cudaFreeAsync(ws, st);
  return rc;
}
