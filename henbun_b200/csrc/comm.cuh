// Run-time bound NCCL (comm.cu) + the description of a rank's place in a column-block-cyclic factorisation.
#pragma once
#include "common.cuh"

namespace hb {

int comm_unique_id(void* out128_host);
int comm_create(const void* id128_host, int rank, int world, void** comm_out);
int comm_destroy(void* comm);
int comm_bcast_f32(void* comm, float* buf, size_t count, int root, cudaStream_t st);
int comm_allgather_f32(void* comm, const float* send, float* recv, size_t count, cudaStream_t st);
int comm_allreduce_f32(void* comm, float* buf, size_t count, int op_max, cudaStream_t st);

// Column blocks of `block` columns are dealt round-robin, `turn` at a time: block b belongs to rank (b / turn) % world.  world == 1, comm == nullptr:
// the same right-looking schedule on one GPU (panel chain on a high-priority stream next to the trailing updates).
struct DistEnv {
  void* comm = nullptr;
  int rank = 0, world = 1;
  int block = 2048;
  int shard_samples = 0;   // the GP step: every rank draws its own S samples; Z and R are all-gathered before L-bar is formed
  int block_bwd = 0;   // the GP step: block width of the reverse mode (0 = block)
  int turn = 1;        // consecutive blocks per rank before the next rank's turn (exchange granularity = block, ownership = turn blocks)
  int batch = 1;       // far blocks take the finished panels `batch` at a time, as one product over all their columns
};

}  // namespace hb
