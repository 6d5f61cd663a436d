// NCCL glue of the column-block-cyclic factorisations (linalg.cu: potrf_flat / chol_rev_flat).
//
// The reference has no multi-device path at all; the one exchange the distributed blocked Cholesky and its
// reverse mode need is "the owner of a finished panel sends it to every other rank" -- ncclBroadcast over NVLink/NVSwitch
// on a dedicated high-priority stream, overlapped with the trailing updates of the previous panel.
//
// NCCL is bound at RUN TIME (dlopen of the libnccl.so.2 the host process already carries -- torch ships one -- or the
// system one): a single-GPU process never touches it and the library has no link-time dependency on it.
#include "comm.cuh"
#include <dlfcn.h>
#include <cstring>

namespace hb {
namespace {

// the handful of NCCL entry points used, declared locally (nccl.h: ncclUniqueId is 128 opaque bytes, ncclFloat = 7,
// ncclUint8 = 1, ncclSum = 0, ncclMax = 2, ncclSuccess = 0)
struct NcclId { char internal[128]; };
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(void** comm, int nranks, NcclId id, int rank);
typedef int (*fn_destroy)(void* comm);
typedef int (*fn_bcast)(const void* send, void* recv, size_t count, int dtype, int root, void* comm, cudaStream_t st);
typedef int (*fn_allreduce)(const void* send, void* recv, size_t count, int dtype, int op, void* comm, cudaStream_t st);
typedef int (*fn_allgather)(const void* send, void* recv, size_t sendcount, int dtype, void* comm, cudaStream_t st);

struct Nccl {
  void* h = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_destroy destroy = nullptr;
  fn_bcast bcast = nullptr;
  fn_allreduce allreduce = nullptr;
  fn_allgather allgather = nullptr;
  bool tried = false;
};
Nccl g_nccl;      // resolved symbols of a shared library: a resource cache, not configuration

bool nccl_load() {
  if (g_nccl.tried) return g_nccl.h != nullptr;
  g_nccl.tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy the process already has (torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return false;
  g_nccl.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
  g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
  g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
  g_nccl.bcast = (fn_bcast)dlsym(h, "ncclBroadcast");
  g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
  g_nccl.allgather = (fn_allgather)dlsym(h, "ncclAllGather");
  if (!g_nccl.get_id || !g_nccl.allgather || !g_nccl.init_rank || !g_nccl.destroy || !g_nccl.bcast || !g_nccl.allreduce) return false;
  g_nccl.h = h;
  return true;
}

}  // namespace

int comm_unique_id(void* out128) {
  if (!out128) return HB_ERR_ARG;
  if (!nccl_load()) return HB_ERR_CUDA;
  NcclId id;
  if (g_nccl.get_id(&id) != 0) return HB_ERR_CUDA;
  std::memcpy(out128, &id, sizeof(id));
  return HB_OK;
}

int comm_create(const void* id128, int rank, int world, void** comm_out) {
  if (!id128 || !comm_out || world < 1 || rank < 0 || rank >= world) return HB_ERR_ARG;
  if (!nccl_load()) return HB_ERR_CUDA;
  NcclId id;
  std::memcpy(&id, id128, sizeof(id));
  void* c = nullptr;
  if (g_nccl.init_rank(&c, world, id, rank) != 0 || !c) return HB_ERR_CUDA;
  *comm_out = c;
  return HB_OK;
}

int comm_destroy(void* comm) {
  if (!comm) return HB_OK;
  if (!nccl_load()) return HB_ERR_CUDA;
  return g_nccl.destroy(comm) == 0 ? HB_OK : HB_ERR_CUDA;
}

int comm_bcast_f32(void* comm, float* buf, size_t count, int root, cudaStream_t st) {
  if (!comm || !g_nccl.h) return HB_ERR_ARG;
  if (count == 0) return HB_OK;
  return g_nccl.bcast(buf, buf, count, /*ncclFloat*/ 7, root, comm, st) == 0 ? HB_OK : HB_ERR_CUDA;
}

// recv [world x count] <- every rank's send [count], in rank order
int comm_allgather_f32(void* comm, const float* send, float* recv, size_t count, cudaStream_t st) {
  if (!comm || !g_nccl.h) return HB_ERR_ARG;
  if (count == 0) return HB_OK;
  return g_nccl.allgather(send, recv, count, /*ncclFloat*/ 7, comm, st) == 0 ? HB_OK : HB_ERR_CUDA;
}

int comm_allreduce_f32(void* comm, float* buf, size_t count, int op_max, cudaStream_t st) {
  if (!comm || !g_nccl.h) return HB_ERR_ARG;
  if (count == 0) return HB_OK;
  return g_nccl.allreduce(buf, buf, count, /*ncclFloat*/ 7, op_max ? /*ncclMax*/ 2 : /*ncclSum*/ 0, comm, st) == 0 ? HB_OK
                                                                                                                   : HB_ERR_CUDA;
}

}  // namespace hb
