// hitl_internal.h — context object and device-side layouts shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/hitl_gpu.h"

namespace hitl {

// Target-side per-pose record written by pose_prep_kernel on every search call (64 B, 16 B
// aligned so a warp-uniform read is four LDG.128 broadcasts).
struct __align__(16) PoseRec {
  float i00, i01, i10, i11;       // inverse of Translation(t) * Rotation(theta): linear part ...
  float itx, ity;                 // ... and translation
  uint32_t off, n;                // scan offset / size
  float gx0, gy0, ginv;           // occupancy grid of the scan: origin and 1 / cell size (cell >= thr)
  uint32_t gdim;                  // nx | ny << 16  (0 = empty scan)
  uint32_t goff;                  // first word of the scan's bitmap
  float theta;                    // (float) pose angle: the direction filter works in float angles with a margin (search.cu)
  float nmax;                     // largest |normal| among the scan's nodes (1 for unit normals): bounds |nb| in the angle gate
  uint32_t foff;                  // first word of the scan's FINE bitmap (kFineCells x kFineCells cells per coarse cell), kNoFine = none
};
constexpr uint32_t kNoFine = 0xFFFFFFFFu;
constexpr int kFineCells = 4;     // fine cell = coarse cell / 4  (power of two: the fine 1/cell is exact)
constexpr int kMipCells = 4;      // mip cell = 4 x 4 coarse cells
// Per-scan occupancy grid descriptor (host-built from the scan AABBs for a given threshold).
struct GridRec { float gx0, gy0, ginv; uint32_t gdim, goff, foff; };
static_assert(sizeof(PoseRec) == 64, "PoseRec must be 64 bytes");

template <typename T> struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;   // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T));
    if (e == cudaSuccess) cap = n ? n : 1;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
// A DevBuf local to one call: freed on every exit path (the HITL_CUDA early returns included).
template <typename T> struct TmpBuf : DevBuf<T> {
  TmpBuf() {}
  ~TmpBuf() { this->release(); }
  TmpBuf(const TmpBuf&) = delete;
  TmpBuf& operator=(const TmpBuf&) = delete;
};

}  // namespace hitl

struct hitl_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evx[4] = {nullptr, nullptr, nullptr, nullptr};   // extra phase marks of hitl_find_stf (HITL_STF_TIMING=1 prints them)
  cudaEvent_t kev[HITL_K_COUNT][2] = {};                       // begin / end of the last launch of the kernels hitl_last_kernel_ms names
  bool kev_set[HITL_K_COUNT] = {};
  std::string err;
  uint64_t launches = 0;

  // ---- scans ----
  uint32_t n_poses = 0;
  uint64_t n_points = 0;
  uint32_t max_scan = 0;
  std::vector<uint32_t> h_off;           // n_poses + 1
  std::vector<float> h_pts, h_nrm;       // kept for the host tree builder
  hitl::DevBuf<uint32_t> d_off;
  hitl::DevBuf<float2> d_pts, d_nrm;
  hitl::DevBuf<float4> d_aabb;           // robot-frame AABB per scan (minx, miny, maxx, maxy)
  // tiles: up to 32 consecutive points of one scan
  uint32_t n_tiles = 0;
  hitl::DevBuf<uint32_t> d_tile_scan, d_tile_k0, d_tile_begin;   // tile -> scan, first point | length << 16; scan -> first tile
  std::vector<uint32_t> h_tile_scan, h_tile_kl, h_tile_begin;    // host mirrors of the three tables
  // A heavy tile may further be cut along the TARGET axis into consecutive "units" (same points, disjoint ascending
  // target ranges [jlo, jhi], full range = [0, 0xFFFFFFFF]); every unit is a tile of its own for the scheduler.  Unit 0
  // of a group keeps the tile's record slot (off + k0), the others get slots after the last point.
  std::vector<uint32_t> h_tile_jlo, h_tile_jhi, h_tile_slot;
  hitl::DevBuf<uint2> d_tile_j;                                  // (jlo, jhi)
  hitl::DevBuf<uint32_t> d_tile_slot;                            // record region of the unit = slot * cap
  hitl::DevBuf<uint2> d_groups;                                  // (first unit, units) of every tile that is split along the target axis
  uint32_t n_groups = 0;
  uint64_t n_slots = 0;                                          // n_points + extra slots of the split units
  uint32_t tiling_splits = 0;                                    // heavy tiles split so far (adaptive tiling)
  uint32_t split_rounds = 0, split_lo = 0, split_hi = 0;         // calls that were allowed to re-tile for the source range [split_lo, split_hi)
  int adaptive_tiling = 1;
  // split policy of the adaptive tiling (scheduling only; env HITL_SPLIT_LIMIT_DIV / HITL_MIN_TARGET_SPAN / HITL_SPLIT_ROUNDS override)
  uint32_t split_limit_div = 1;          // a tile heavier than 2 x (fair share / split_limit_div) is split (1: measured best on 1/8 shards of c2, profiles/diag_shard.py)
  uint32_t min_target_span = 64;         // target-axis ranges never get narrower than this many poses
  uint32_t max_split_rounds = 3;         // only the first max_split_rounds calls on a source range may re-tile (setup; W >= 3 warm-up steps cover it)
  int target_splitting = 1;              // heavy tiles may also be cut along the target axis (hitl_debug_set_tiling)
  int search_carveout = -1;              // preferred shared-memory carve-out in percent (-1: driver default)
  int search_variant = 0;                // 0: 16 CTAs/SM (32 regs), 1: 12 CTAs/SM (40 regs), 2: 10 CTAs/SM (hitl_debug_set_search_variant)
  // scheduling hint of the search: tiles sorted by the cycles the previous call spent on them
  hitl::DevBuf<uint32_t> d_tile_work, d_tile_order, d_tile_iota, d_tile_keys, d_tile_open, d_tile_end;
  hitl::DevBuf<uint8_t> d_sort_tmp;

  // ---- trees ----
  bool have_trees = false;
  int tree_builder = 0;                  // 0: device (kdtree_gpu.cu), 1: host threads (kdtree_build.cpp)   (hitl_debug_set_tree_builder)
  uint64_t tree_exact_segments = 0;      // segments of the last device build that needed the exact std::sort emulation (equal keys)
  hitl::DevBuf<float4> d_node_pm;        // px, py, bits(index | dim << 31), 0   (preorder, concatenated): one 16 B load per node visit
  hitl::DevBuf<float2> d_node_nn;        // nx, ny   (read only for nodes inside the query radius)
  hitl::DevBuf<hitl_kdnode> d_node_aos;  // staging for hitl_set_kdtrees / hitl_get_kdtrees (AoS <-> SoA on the device)
  hitl::DevBuf<uint32_t> d_node_compact; // staging for hitl_set_kdtrees_compact / hitl_get_kdtrees_compact (index | dim << 31 per node)
  hitl::DevBuf<uint32_t> d_pack_k, d_pack_idx;   // staging for hitl_get_stf16 (two 16-bit indices per word)

  // ---- per-call pose tables ----
  hitl::DevBuf<double> d_pose;           // x, y, theta
  hitl::DevBuf<hitl::PoseRec> d_rec;
  hitl::DevBuf<float4> d_wbox;           // world-frame AABB per scan, inflated (pair cull)
  hitl::DevBuf<float4> d_gbox;           // union of d_wbox over every 32 consecutive poses (first level of the target sweep)
  hitl::DevBuf<float4> d_src;            // source-side transform per pose: cos, sin, tx, ty
  // occupancy bitmaps (exact "no point of scan j within thr of q" test), rebuilt when thr or the scans change
  std::vector<float> h_aabb;             // 4 per scan
  hitl::DevBuf<hitl::GridRec> d_grid;
  hitl::DevBuf<uint32_t> d_occ;
  hitl::DevBuf<uint32_t> d_occ_fine;     // second level: cell = thr / 4, consulted only for points that pass the coarse level
  hitl::DevBuf<uint32_t> d_occ_mip;      // coarse bitmap reduced kMipCells x kMipCells (tile-box vs scan cull, one lane per target pose)
  hitl::DevBuf<uint32_t> d_moff;         // per scan: first word of its mip bitmap
  int mip_occupancy = 1;                 // (hitl_debug_set_fine_occupancy bit 2 set clears it)
  hitl::DevBuf<uint32_t> d_occ_dir;      // per coarse cell: 16 direction bins of the NODE normals near it (two cells per word); angle-gate prefilter
  hitl::DevBuf<float> d_nmax;            // per scan: largest node-normal length
  int dir_occupancy = 1;                 // direction prefilter on (hitl_debug_set_fine_occupancy bit 1 clears it)
  bool grid_valid = false;
  int fine_occupancy = 1;                // second-level bitmaps on (hitl_debug_set_fine_occupancy)
  float grid_thr = 0.f;

  // ---- search results ----
  hitl::DevBuf<uint32_t> d_raw_j, d_raw_k, d_raw_idx, d_tile_cnt;   // per-tile raw records
  hitl::DevBuf<uint32_t> d_srt_j, d_srt_k, d_srt_idx;               // per-pose sorted staging
  hitl::DevBuf<uint8_t> d_srt_flag;                                 // bit0 keep, bit1 first-of-pair
  hitl::DevBuf<uint64_t> d_order_state;  // look-back states of the single-pass ordering kernel (4 words per source pose)
  int order_two_pass = 0;                // HITL_ORDER_TWO_PASS=1: the older count / scan / place sequence (A/B check)
  hitl::DevBuf<uint64_t> d_pose_cnt;     // per pose: kept matches, kept pairs (2 per pose) then scanned
  hitl::DevBuf<uint64_t> d_pose_work;    // SM cycles per source pose of the last search (shard balancing)
  hitl::DevBuf<uint64_t> d_counters;     // [0] n_queries [1] n_traversals [2] raw matches [3] pairs [4] matches
  hitl::DevBuf<uint32_t> d_pair_i, d_pair_j, d_k, d_idx;
  hitl::DevBuf<uint64_t> d_pair_off;
  uint64_t n_pairs = 0, n_matches = 0;
  bool have_stf = false;
  // vo
  hitl::DevBuf<uint32_t> d_vo_sp, d_vo_sk, d_vo_tk;
  uint64_t n_vo = 0;

  // ---- world clouds / EM ----
  hitl::DevBuf<float2> d_world;
  bool have_world = false;
  hitl::DevBuf<float> d_poses_f;
  hitl::DevBuf<uint32_t> d_em_pose, d_em_idx, d_em_obs[2], d_em_cnt[2], d_em_setpose[2], d_em_slots[2], d_em_slotof;
  hitl::DevBuf<uint64_t> d_em_setoff[2];
  hitl::DevBuf<float2> d_em_xy;
  hitl::DevBuf<uint64_t> d_scan_state;   // decoupled look-back tile states
  hitl::DevBuf<uint32_t> d_ticket;
  hitl::DevBuf<float> d_bp_poses, d_bp_rw, d_bp_tw;   // scratch of hitl_backprop_poses
  hitl::DevBuf<float2> d_bp_cs;
  hitl::DevBuf<double> d_fit_partial;    // per-CTA partial sums of the device M-step (two buffers)
  hitl::DevBuf<uint8_t> d_fit_out;       // FitResults of the last hitl_em_refit / hitl_em_refit_chain (one per launch)
  hitl::DevBuf<float4> d_em_box;         // bounding box of every E-step chunk of the resident world clouds
  bool em_boxes_valid = false;           // recorded by the first E-step after the world clouds changed, used by the later ones
  bool em_cull = true;                   // hitl_debug_set_em_cull / HITL_EM_CULL=0: every E-step reads every chunk

  // ---- residual blocks ----
  uint64_t nb_odo = 0, nb_human = 0, nb_stf = 0, nb_p2lg = 0, nb_p2l = 0;
  bool stf_from_search = false;
  float stf_std = 0.05f, stf_corr = 0.025f;
  hitl::DevBuf<float> d_odo;             // 9 per block
  hitl::DevBuf<int32_t> d_hum_i;         // 2 per block
  hitl::DevBuf<double> d_hum_d;          // 4 per block
  hitl::DevBuf<uint32_t> d_blk_i, d_blk_j, d_blk_k, d_blk_idx;   // explicit stf blocks
  hitl::DevBuf<uint64_t> d_blk_off;
  hitl::DevBuf<uint32_t> d_p2lg_pose; hitl::DevBuf<uint64_t> d_p2lg_off;
  hitl::DevBuf<float2> d_p2lg_pts, d_p2lg_n; hitl::DevBuf<float> d_p2lg_o; hitl::DevBuf<uint8_t> d_p2lg_v;
  float p2lg_std = 1, p2lg_corr = 1;
  hitl::DevBuf<uint32_t> d_p2l_pose; hitl::DevBuf<float2> d_p2l_pts, d_p2l_n; hitl::DevBuf<float> d_p2l_o;
  hitl::DevBuf<uint8_t> d_p2l_v;
  float p2l_std = 1, p2l_corr = 1;
  hitl::DevBuf<double> d_r, d_J, d_neq, d_hoff;
  hitl::DevBuf<double2> d_trig;          // per pose (cos, sin), (x, y) of the current evaluation point

  int deterministic = 0;                 // hitl_set_deterministic: normal equations by per-pose gather instead of FP64 atomics
  bool inc_valid = false;                // per-pose incidence lists match the registered blocks
  uint64_t n_inc = 0;
  hitl::DevBuf<uint32_t> d_inc_key;      // [unsorted | sorted] pose of every (block, side)
  hitl::DevBuf<uint64_t> d_inc_ref;      // [unsorted | sorted] block reference: kind << 60 | side << 59 | block
  hitl::DevBuf<uint64_t> d_inc_off;      // per pose: first incidence
  hitl::DevBuf<double> d_cost_partial;
  bool neq_valid = false, eval_valid = false;   // d_neq / d_r + d_J hold the result of a completed evaluation of the current blocks
  // ---- multi-GPU exchange (comm.cu) ----
  bool upload_sharded = false;           // hitl_set_scans_sharded / hitl_set_kdtrees*_sharded: the next uploads move 1/world over PCIe each
  void* comm = nullptr;                  // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  hitl::DevBuf<uint64_t> d_comm_cnt;
  hitl::DevBuf<uint32_t> d_g_pi, d_g_pj; // gathered STF blocks (root only)
  hitl::DevBuf<double> d_g_r, d_g_J;

  // pinned staging for small read-backs
  uint64_t* h_pinned = nullptr;          // 64 x u64
};

namespace hitl {
int fail(hitl_ctx* c, int code, const char* what);
int cuda_fail(hitl_ctx* c, cudaError_t e, const char* where);
#define HITL_CUDA(call)                                              \
  do {                                                               \
    cudaError_t e__ = (call);                                        \
    if (e__ != cudaSuccess) return hitl::cuda_fail(ctx, e__, #call); \
  } while (0)
// Event pair around one named kernel launch (hitl_last_kernel_ms)
#define HITL_KERNEL_BEGIN(which) do { if (ctx->kev[which][0]) cudaEventRecord(ctx->kev[which][0], ctx->stream); } while (0)
#define HITL_KERNEL_END(which) do { if (ctx->kev[which][1]) { cudaEventRecord(ctx->kev[which][1], ctx->stream); ctx->kev_set[which] = true; } } while (0)
// Every ABI entry makes its context's device current first: several contexts (one per GPU) may live in one process, driven by
// different host threads, and a kernel launch goes to whatever device the calling thread last selected.
#define HITL_DEVICE(ctx)                                                       \
  do {                                                                         \
    cudaError_t e__ = cudaSetDevice((ctx)->device);                            \
    if (e__ != cudaSuccess) return hitl::cuda_fail(ctx, e__, "cudaSetDevice"); \
  } while (0)
#define HITL_LAUNCH_CHECK(name)                                      \
  do {                                                               \
    ctx->launches++;                                                 \
    cudaError_t e__ = cudaGetLastError();                            \
    if (e__ != cudaSuccess) return hitl::cuda_fail(ctx, e__, name);  \
  } while (0)

// Upload of a buffer that every rank of the job holds identically on its host (comm.cu): with a communicator and ctx->upload_sharded
// set, this rank copies only its 1/world slice across PCIe and the slices are exchanged in place with ncclAllGather over NVLink;
// otherwise a plain cudaMemcpyAsync.  `capacity_bytes` of the device buffer must be >= replicated_upload_capacity(ctx, bytes).
size_t replicated_upload_capacity(const hitl_ctx* ctx, size_t bytes);
int replicated_upload(hitl_ctx* ctx, void* dev, const void* host, size_t bytes);
int build_tiling(hitl_ctx* ctx, uint32_t max_len);
int upload_tiling(hitl_ctx* ctx);
uint32_t split_heavy_tiles(hitl_ctx* ctx, const std::vector<uint32_t>& h_work, const std::vector<uint32_t>& h_open, const std::vector<uint32_t>& h_end, uint32_t lo,
                           uint32_t hi, uint64_t limit, std::vector<uint32_t>* est);
constexpr uint32_t kFullRange = 0xFFFFFFFFu;
// device tree builder (kdtree_gpu.cu): all scans at once, reference-identical shape
int build_kdtrees_device(hitl_ctx* ctx, uint64_t* n_exact_out);
// host tree builder (kdtree_build.cpp)
void build_flat_kdtree(const float* pts_xy, const float* nrm_xy, uint32_t n, hitl_kdnode* out);
}  // namespace hitl
