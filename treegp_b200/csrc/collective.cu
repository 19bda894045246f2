// Collective behind the C ABI (tgp_allreduce_bins): the ONE exchange step of the hot path --
// summing the packed bin buffer {npairs, sumw, sumwkk[, sumwr]} of tgp_pairbin over the ranks that were dealt the
// pair tiles (replaces nothing in the reference, which is single-process: /root/reference/treegp/two_pcf.py:283-340
// runs one TreeCorr process call).  NCCL is bound at run time (dlopen of the libnccl the process already carries --
// torch's bundled one -- or the system library): the library has no link-time dependency on it, and a caller that
// never shards never loads it.
//
// Usage from a maintainer's binding (INTEGRATION.md route B): rank 0 calls tgp_comm_unique_id and ships the 128 bytes
// to the other ranks by whatever means the job has (MPI, a file, torch.distributed); every rank calls
// tgp_comm_init_rank, then tgp_pairbin(tile_rank = rank, tile_nranks = nranks) followed by tgp_allreduce_bins on the
// packed buffer, on the same stream.
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>
#include "tgp_common.cuh"

namespace {
typedef struct { char internal[128]; } nccl_unique_id;
typedef void* nccl_comm;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm*, int, nccl_unique_id, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t);
typedef int (*fn_comm_destroy)(nccl_comm);
typedef const char* (*fn_get_error_string)(int);
struct NcclApi {
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_get_error_string get_error_string = nullptr;
  bool ok = false;
};
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t values (stable across NCCL 2.x)

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // already in the process (torch)?
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
  api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
  api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
  api.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
  api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  api.get_error_string = (fn_get_error_string)dlsym(h, "ncclGetErrorString");
  api.ok = api.get_unique_id && api.comm_init_rank && api.all_reduce && api.comm_destroy;
  return api;
}
int nccl_fail(const char* what, int rc) {
  NcclApi& a = nccl();
  tgp_set_error("%s: NCCL error %d (%s)", what, rc, a.get_error_string ? a.get_error_string(rc) : "?");
  return TGP_ERR_CUDA;
}
// counts (int64 words) <-> FP64 values, in place: sums of integer-valued doubles below 2^53 are exact in any order,
// so the reduced counts are bit-identical for any number of ranks
__global__ void counts_to_f64_kernel(double* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (double)reinterpret_cast<const long long*>(p)[i];
}
__global__ void f64_to_counts_kernel(double* p, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) reinterpret_cast<long long*>(p)[i] = (long long)p[i];
}
}  // namespace

extern "C" int tgp_comm_unique_id(void* id128) {
  TGP_CHECK_ARG(id128 != nullptr, "id128");
  NcclApi& a = nccl();
  if (!a.ok) { tgp_set_error("tgp_comm_unique_id: libnccl.so.2 not found"); return TGP_ERR_UNSUPPORTED; }
  nccl_unique_id id;
  const int rc = a.get_unique_id(&id);
  if (rc) return nccl_fail("ncclGetUniqueId", rc);
  memcpy(id128, &id, sizeof(id));
  return TGP_OK;
}

extern "C" int tgp_comm_init_rank(const void* id128, int32_t rank, int32_t nranks, void** comm) {
  TGP_CHECK_ARG(id128 && comm && nranks >= 1 && rank >= 0 && rank < nranks, "id / comm / rank");
  NcclApi& a = nccl();
  if (!a.ok) { tgp_set_error("tgp_comm_init_rank: libnccl.so.2 not found"); return TGP_ERR_UNSUPPORTED; }
  nccl_unique_id id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm c = nullptr;
  const int rc = a.comm_init_rank(&c, nranks, id, rank);
  if (rc) return nccl_fail("ncclCommInitRank", rc);
  *comm = c;
  return TGP_OK;
}

extern "C" int tgp_comm_destroy(void* comm) {
  if (!comm) return TGP_OK;
  NcclApi& a = nccl();
  if (!a.ok) return TGP_ERR_UNSUPPORTED;
  const int rc = a.comm_destroy(comm);
  return rc ? nccl_fail("ncclCommDestroy", rc) : TGP_OK;
}

extern "C" int tgp_allreduce_bins(void* comm, double* packed, int64_t planes, int64_t per_plane, void* stream) {
  TGP_CHECK_ARG(comm && packed && planes >= 1 && per_plane >= 0, "comm / packed / shape");
  if (per_plane == 0) return TGP_OK;
  NcclApi& a = nccl();
  if (!a.ok) { tgp_set_error("tgp_allreduce_bins: libnccl.so.2 not found"); return TGP_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)tgp_cdiv(per_plane, 256);
  counts_to_f64_kernel<<<grid, 256, 0, st>>>(packed, per_plane);          // plane 0: the int64 pair counts
  TGP_LAUNCH_CHECK();
  const int rc = a.all_reduce(packed, packed, (size_t)(planes * per_plane), NCCL_FLOAT64, NCCL_SUM, comm, st);
  if (rc) return nccl_fail("ncclAllReduce", rc);
  f64_to_counts_kernel<<<grid, 256, 0, st>>>(packed, per_plane);
  TGP_LAUNCH_CHECK();
  return TGP_OK;
}
