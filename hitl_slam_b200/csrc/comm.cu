// comm.cu — the multi-GPU exchange of the hot path, inside the C ABI.
//
// The correspondence search shards by SOURCE pose with no data-path collective (each rank searches its own contiguous source
// range against all targets on replicated scans + trees) and every STF residual block lives on the rank that found it.  What
// crosses GPUs per Gauss-Newton / LM iteration of the reference's PostHumanOptimization (JointOptimization.cpp:1156-1256):
//   * hitl_normal_eq_allreduce   one ncclAllReduce(sum, f64) over the packed resident buffer [H_diag N x 9 | g N x 3 | cost]
//                                (pose blocks touched by blocks of several ranks, hence a reduction), issued on the context's
//                                stream directly behind the kernels that fill it — one host synchronisation for both;
//   * hitl_gather_stf_blocks     the "Ceres on the host" feed: every rank's STF blocks (pair_i, pair_j, r[2], J[12] = 14 doubles
//                                per block, what SizedCostFunction<2,3,3>::Evaluate hands out) gathered to one root in rank order,
//                                which is the reference's block order because the shards are ascending source ranges.
// NCCL is bound at run time (dlopen of libnccl.so.2: whichever copy the process already holds — e.g. the one torch loaded —
// else the system one), so libhitl_gpu.so itself has no link-time dependency and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <vector>
#include "hitl_internal.h"

using namespace hitl;

namespace {
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  bool ok = false;
};
NcclApi g_nccl;
std::once_flag g_nccl_once;

void load_nccl() {
  void* h = nullptr;
  if (const char* path = getenv("HITL_NCCL_LIBRARY")) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);    // the copy this process already holds (e.g. torch's: ONE copy per process)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) return;
  g_nccl.handle = h;
#define HITL_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name)
  HITL_SYM(GetUniqueId, "ncclGetUniqueId"); HITL_SYM(CommInitRank, "ncclCommInitRank"); HITL_SYM(CommDestroy, "ncclCommDestroy");
  HITL_SYM(AllReduce, "ncclAllReduce"); HITL_SYM(AllGather, "ncclAllGather"); HITL_SYM(Send, "ncclSend"); HITL_SYM(Recv, "ncclRecv");
  HITL_SYM(GroupStart, "ncclGroupStart"); HITL_SYM(GroupEnd, "ncclGroupEnd"); HITL_SYM(GetErrorString, "ncclGetErrorString");
  HITL_SYM(GetVersion, "ncclGetVersion");
#undef HITL_SYM
  g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather && g_nccl.Send && g_nccl.Recv &&
              g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
}
bool nccl_ready() { std::call_once(g_nccl_once, load_nccl); return g_nccl.ok; }

int nccl_fail(hitl_ctx* ctx, ncclResult_t r, const char* where) {
  char buf[256];
  snprintf(buf, sizeof(buf), "%s: NCCL error %d (%s)", where, (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
  return fail(ctx, HITL_ERR_NCCL, buf);
}
#define HITL_NCCL(call)                                       \
  do {                                                        \
    ncclResult_t r__ = (call);                                \
    if (r__ != ncclSuccess) return nccl_fail(ctx, r__, #call); \
  } while (0)
}  // namespace

static_assert(sizeof(ncclUniqueId) == HITL_COMM_ID_BYTES, "hitl_gpu.h: HITL_COMM_ID_BYTES must be sizeof(ncclUniqueId)");

namespace hitl {
static size_t upload_slice(const hitl_ctx* ctx, size_t bytes) {
  const size_t w = (size_t)ctx->comm_world;
  return (((bytes + w - 1) / w) + 15) & ~(size_t)15;             // equal slices, 16-byte aligned
}
size_t replicated_upload_capacity(const hitl_ctx* ctx, size_t bytes) {
  if (!ctx->upload_sharded || !ctx->comm || ctx->comm_world <= 1) return bytes;
  return upload_slice(ctx, bytes) * (size_t)ctx->comm_world;
}
int replicated_upload(hitl_ctx* ctx, void* dev, const void* host, size_t bytes) {
  if (bytes == 0) return HITL_OK;
  if (!ctx->upload_sharded || !ctx->comm || ctx->comm_world <= 1) {
    HITL_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return HITL_OK;
  }
  const size_t slice = upload_slice(ctx, bytes), lo = slice * (size_t)ctx->comm_rank;
  if (lo < bytes)
    HITL_CUDA(cudaMemcpyAsync(static_cast<char*>(dev) + lo, static_cast<const char*>(host) + lo, std::min(slice, bytes - lo), cudaMemcpyHostToDevice, ctx->stream));
  HITL_NCCL(g_nccl.AllGather(static_cast<char*>(dev) + lo, dev, slice, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));   // in place: every rank's slice sits at its own offset
  ctx->launches++;
  return HITL_OK;
}
}  // namespace hitl

extern "C" int hitl_comm_unique_id(void* id_out) {
  if (!id_out || !nccl_ready()) return HITL_ERR_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return HITL_ERR_NCCL;
  memcpy(id_out, &id, sizeof(id));
  return HITL_OK;
}

extern "C" int hitl_comm_init(hitl_ctx* ctx, const void* nccl_unique_id, int rank, int world) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!nccl_unique_id || world < 1 || rank < 0 || rank >= world) return fail(ctx, HITL_ERR_ARG, "hitl_comm_init: bad argument");
  if (!nccl_ready()) return fail(ctx, HITL_ERR_NCCL, "hitl_comm_init: libnccl.so.2 not found (no NCCL in this process and none on the library path)");
  if (ctx->comm) return fail(ctx, HITL_ERR_STATE, "hitl_comm_init: this context already has a communicator (hitl_comm_destroy first)");
  HITL_CUDA(cudaSetDevice(ctx->device));
  ncclUniqueId id;
  memcpy(&id, nccl_unique_id, sizeof(id));
  ncclComm_t comm = nullptr;
  HITL_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
  ctx->comm = comm; ctx->comm_rank = rank; ctx->comm_world = world;
  return HITL_OK;
}

extern "C" int hitl_comm_destroy(hitl_ctx* ctx) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_rank = 0; ctx->comm_world = 1;
  return HITL_OK;
}

extern "C" int hitl_comm_info(const hitl_ctx* ctx, int* rank, int* world, int* nccl_version) {
  if (!ctx) return HITL_ERR_ARG;
  if (rank) *rank = ctx->comm_rank;
  if (world) *world = ctx->comm ? ctx->comm_world : 1;
  if (nccl_version) { *nccl_version = 0; if (nccl_ready() && g_nccl.GetVersion) g_nccl.GetVersion(nccl_version); }
  return HITL_OK;
}

namespace hitl { int normal_eq_launch(hitl_ctx* ctx, const double* pose_array); }

extern "C" int hitl_normal_eq_allreduce(hitl_ctx* ctx, const double* pose_array, double* H_diag, double* g, double* cost, float* ms_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  const size_t n = ctx->n_poses;
  HITL_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
  if (pose_array) {                                   // evaluate this rank's blocks first, same stream, no synchronisation in between
    const int rc = normal_eq_launch(ctx, pose_array);
    if (rc) return rc;
  } else if (!ctx->d_neq.p || !ctx->neq_valid) {
    return fail(ctx, HITL_ERR_STATE, "hitl_normal_eq_allreduce: no normal equations resident (pass pose_array or call hitl_normal_eq first)");
  }
  if (ctx->comm && ctx->comm_world > 1) {
    HITL_KERNEL_BEGIN(HITL_K_ALLREDUCE);
    HITL_NCCL(g_nccl.AllReduce(ctx->d_neq.p, ctx->d_neq.p, 12 * n + 1, ncclDouble, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    HITL_KERNEL_END(HITL_K_ALLREDUCE);
    ctx->launches++;
  }
  HITL_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
  if (H_diag && n) HITL_CUDA(cudaMemcpyAsync(H_diag, ctx->d_neq.p, 72 * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (g && n) HITL_CUDA(cudaMemcpyAsync(g, ctx->d_neq.p + 9 * n, 24 * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (cost) HITL_CUDA(cudaMemcpyAsync(cost, ctx->d_neq.p + 12 * n, 8, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ms_out) HITL_CUDA(cudaEventElapsedTime(ms_out, ctx->ev[0], ctx->ev[1]));
  return HITL_OK;
}

extern "C" int hitl_gather_stf_blocks(hitl_ctx* ctx, int root, uint64_t cap_blocks, uint64_t* n_blocks_per_rank, uint32_t* pair_i, uint32_t* pair_j,
                                      double* r, double* J) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  const int world = ctx->comm ? ctx->comm_world : 1, rank = ctx->comm ? ctx->comm_rank : 0;
  if (root < 0 || root >= world) return fail(ctx, HITL_ERR_ARG, "hitl_gather_stf_blocks: root out of range");
  if (!ctx->eval_valid) return fail(ctx, HITL_ERR_STATE, "hitl_gather_stf_blocks: call hitl_eval (with Jacobians) first: r and J of this rank's blocks must be resident");
  const uint64_t mine = ctx->nb_stf;
  const bool fs = ctx->stf_from_search;
  const uint32_t* d_pi = fs ? ctx->d_pair_i.p : ctx->d_blk_i.p;
  const uint32_t* d_pj = fs ? ctx->d_pair_j.p : ctx->d_blk_j.p;
  // this rank's STF slice of the evaluation buffers (block order of hitl_eval: odometry | human | stf | ...)
  const double* d_r = ctx->d_r.p + 3 * ctx->nb_odo + 3 * ctx->nb_human;
  const double* d_J = ctx->d_J.p + 18 * ctx->nb_odo + 9 * ctx->nb_human;
  std::vector<uint64_t> counts(world, 0);
  counts[rank] = mine;
  if (world > 1) {
    HITL_CUDA(ctx->d_comm_cnt.ensure(world));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_comm_cnt.p + rank, &mine, 8, cudaMemcpyHostToDevice, ctx->stream));
    HITL_NCCL(g_nccl.AllGather(ctx->d_comm_cnt.p + rank, ctx->d_comm_cnt.p, 1, ncclUint64, (ncclComm_t)ctx->comm, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(counts.data(), ctx->d_comm_cnt.p, 8 * (size_t)world, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  if (n_blocks_per_rank) for (int q = 0; q < world; ++q) n_blocks_per_rank[q] = counts[q];
  uint64_t total = 0;
  for (int q = 0; q < world; ++q) total += counts[q];
  if (rank == root) {
    // every sender must be received even when the caller's buffers are too small (a rank that returned early would hang the others)
    HITL_CUDA(ctx->d_g_pi.ensure(total)); HITL_CUDA(ctx->d_g_pj.ensure(total)); HITL_CUDA(ctx->d_g_r.ensure(2 * total)); HITL_CUDA(ctx->d_g_J.ensure(12 * total));
  }
  if (world > 1) {
    HITL_NCCL(g_nccl.GroupStart());
    ncclResult_t first_error = ncclSuccess;           // the group is always closed; the first failure inside it is reported afterwards
    auto note = [&](ncclResult_t r) { if (r != ncclSuccess && first_error == ncclSuccess) first_error = r; };
    if (rank == root) {
      uint64_t o = 0;
      for (int q = 0; q < world; ++q) {
        const uint64_t c = counts[q];
        if (q != root && c) {
          note(g_nccl.Recv(ctx->d_g_pi.p + o, c, ncclUint32, q, (ncclComm_t)ctx->comm, ctx->stream));
          note(g_nccl.Recv(ctx->d_g_pj.p + o, c, ncclUint32, q, (ncclComm_t)ctx->comm, ctx->stream));
          note(g_nccl.Recv(ctx->d_g_r.p + 2 * o, 2 * c, ncclDouble, q, (ncclComm_t)ctx->comm, ctx->stream));
          note(g_nccl.Recv(ctx->d_g_J.p + 12 * o, 12 * c, ncclDouble, q, (ncclComm_t)ctx->comm, ctx->stream));
        }
        o += c;
      }
    } else if (mine) {
      note(g_nccl.Send(d_pi, mine, ncclUint32, root, (ncclComm_t)ctx->comm, ctx->stream));
      note(g_nccl.Send(d_pj, mine, ncclUint32, root, (ncclComm_t)ctx->comm, ctx->stream));
      note(g_nccl.Send(d_r, 2 * mine, ncclDouble, root, (ncclComm_t)ctx->comm, ctx->stream));
      note(g_nccl.Send(d_J, 12 * mine, ncclDouble, root, (ncclComm_t)ctx->comm, ctx->stream));
    }
    HITL_NCCL(g_nccl.GroupEnd());
    if (first_error != ncclSuccess) return nccl_fail(ctx, first_error, "ncclSend / ncclRecv inside hitl_gather_stf_blocks");
    ctx->launches++;
  }
  if (rank == root) {
    uint64_t o = 0;
    for (int q = 0; q < root; ++q) o += counts[q];
    if (mine) {                                        // the root's own blocks take their place in rank order
      HITL_CUDA(cudaMemcpyAsync(ctx->d_g_pi.p + o, d_pi, 4 * mine, cudaMemcpyDeviceToDevice, ctx->stream));
      HITL_CUDA(cudaMemcpyAsync(ctx->d_g_pj.p + o, d_pj, 4 * mine, cudaMemcpyDeviceToDevice, ctx->stream));
      HITL_CUDA(cudaMemcpyAsync(ctx->d_g_r.p + 2 * o, d_r, 16 * mine, cudaMemcpyDeviceToDevice, ctx->stream));
      HITL_CUDA(cudaMemcpyAsync(ctx->d_g_J.p + 12 * o, d_J, 96 * mine, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (total <= cap_blocks && total) {
      if (pair_i) HITL_CUDA(cudaMemcpyAsync(pair_i, ctx->d_g_pi.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
      if (pair_j) HITL_CUDA(cudaMemcpyAsync(pair_j, ctx->d_g_pj.p, 4 * total, cudaMemcpyDeviceToHost, ctx->stream));
      if (r) HITL_CUDA(cudaMemcpyAsync(r, ctx->d_g_r.p, 16 * total, cudaMemcpyDeviceToHost, ctx->stream));
      if (J) HITL_CUDA(cudaMemcpyAsync(J, ctx->d_g_J.p, 96 * total, cudaMemcpyDeviceToHost, ctx->stream));
    }
  }
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  if (rank == root && total > cap_blocks) return fail(ctx, HITL_ERR_OVERFLOW, "hitl_gather_stf_blocks: more blocks than cap_blocks (sizes are in n_blocks_per_rank)");
  return HITL_OK;
}
