// ctx.cu — context lifetime, scan / KD-tree residency (C ABI: include/hitl_gpu.h).
//
// Replaces JointOpt::BuildKDTrees (human_in_the_loop_slam/JointOptimization.cpp:514-537): scans
// and trees are uploaded once per session and stay in HBM; the reference keeps one heap node
// per point behind `kdtrees_`.
#include <string.h>
#include <algorithm>
#include <thread>
#include "hitl_internal.h"
#include "hitl_math.h"

namespace hitl {
int fail(hitl_ctx* c, int code, const char* what) {
  if (c) c->err = what;
  return code;
}
int cuda_fail(hitl_ctx* c, cudaError_t e, const char* where) {
  if (c) { c->err = std::string(where) + ": " + cudaGetErrorString(e); }
  cudaGetLastError();
  return HITL_ERR_CUDA;
}
int launch_scan_aabb(hitl_ctx* ctx);

// Uploads the host tile tables (h_tile_scan, h_tile_kl = k0 | len << 16, h_tile_begin).
int upload_tiling(hitl_ctx* ctx) {
  const uint32_t nt = (uint32_t)ctx->h_tile_scan.size();
  ctx->n_tiles = nt;
  // record slots and target-axis groups: unit 0 of a group (and every unsplit tile) writes at its points' own slots
  ctx->h_tile_slot.resize(nt);
  std::vector<uint2> groups;
  uint64_t extra = 0;
  for (uint32_t t = 0; t < nt;) {
    uint32_t e = t + 1;
    const bool split = ctx->h_tile_jlo[t] != 0 || ctx->h_tile_jhi[t] != kFullRange;
    if (split) while (e < nt && ctx->h_tile_scan[e] == ctx->h_tile_scan[t] && ctx->h_tile_kl[e] == ctx->h_tile_kl[t] && ctx->h_tile_jlo[e] == ctx->h_tile_jhi[e - 1] + 1) ++e;
    const uint32_t len = ctx->h_tile_kl[t] >> 16;
    ctx->h_tile_slot[t] = ctx->h_off[ctx->h_tile_scan[t]] + (ctx->h_tile_kl[t] & 0xFFFFu);
    for (uint32_t u = t + 1; u < e; ++u) { ctx->h_tile_slot[u] = (uint32_t)(ctx->n_points + extra); extra += len; }
    if (e - t > 1) groups.push_back(make_uint2(t, e - t));
    t = e;
  }
  ctx->n_slots = ctx->n_points + extra;
  ctx->n_groups = (uint32_t)groups.size();
  // per-tile cost estimates (0 = never searched) and the identity permutation the scheduler sorts by them
  HITL_CUDA(ctx->d_tile_work.ensure(nt)); HITL_CUDA(ctx->d_tile_order.ensure(nt)); HITL_CUDA(ctx->d_tile_iota.ensure(nt)); HITL_CUDA(ctx->d_tile_keys.ensure(nt));
  HITL_CUDA(ctx->d_tile_scan.ensure(nt)); HITL_CUDA(ctx->d_tile_k0.ensure(nt)); HITL_CUDA(ctx->d_tile_begin.ensure(ctx->n_poses + 1));
  HITL_CUDA(ctx->d_tile_j.ensure(nt)); HITL_CUDA(ctx->d_tile_slot.ensure(nt)); HITL_CUDA(ctx->d_groups.ensure(groups.size()));
  std::vector<uint32_t> iota(nt);
  std::vector<uint2> tj(nt);
  for (uint32_t t = 0; t < nt; ++t) { iota[t] = t; tj[t] = make_uint2(ctx->h_tile_jlo[t], ctx->h_tile_jhi[t]); }
  if (nt) {
    HITL_CUDA(cudaMemsetAsync(ctx->d_tile_work.p, 0, 4 * (size_t)nt, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_iota.p, iota.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_scan.p, ctx->h_tile_scan.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_k0.p, ctx->h_tile_kl.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_j.p, tj.data(), 8 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
    HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_slot.p, ctx->h_tile_slot.data(), 4 * (size_t)nt, cudaMemcpyHostToDevice, ctx->stream));
  }
  if (!groups.empty()) HITL_CUDA(cudaMemcpyAsync(ctx->d_groups.p, groups.data(), 8 * groups.size(), cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(ctx->d_tile_begin.p, ctx->h_tile_begin.data(), 4 * (size_t)(ctx->n_poses + 1), cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));   // the staging vectors are locals
  return HITL_OK;
}

// Uniform tiling: every scan cut into tiles of at most max_len points, each over the full target range.
int build_tiling(hitl_ctx* ctx, uint32_t max_len) {
  if (max_len < 1) max_len = 1;
  if (max_len > 32) max_len = 32;
  const uint32_t n_poses = ctx->n_poses;
  ctx->h_tile_scan.clear(); ctx->h_tile_kl.clear();
  ctx->h_tile_scan.reserve(ctx->n_points / max_len + n_poses); ctx->h_tile_kl.reserve(ctx->n_points / max_len + n_poses);
  ctx->h_tile_begin.assign(n_poses + 1, 0);
  for (uint32_t i = 0; i < n_poses; ++i) {
    ctx->h_tile_begin[i] = (uint32_t)ctx->h_tile_scan.size();
    const uint32_t n = ctx->h_off[i + 1] - ctx->h_off[i];
    for (uint32_t k0 = 0; k0 < n; k0 += max_len) { ctx->h_tile_scan.push_back(i); ctx->h_tile_kl.push_back(k0 | (std::min(max_len, n - k0) << 16)); }
  }
  ctx->h_tile_begin[n_poses] = (uint32_t)ctx->h_tile_scan.size();
  ctx->h_tile_jlo.assign(ctx->h_tile_scan.size(), 0u);
  ctx->h_tile_jhi.assign(ctx->h_tile_scan.size(), kFullRange);
  ctx->tiling_splits = 0;
  return upload_tiling(ctx);
}

// Splits every tile whose measured work (cycles / 64, h_work indexed by tile id, tiles in [lo, hi)) exceeds `limit`:
// first along the points into 2, 4 or 8 equal parts (never below 4 points), then — a tile is a sequential loop over
// the target poses, so a few points that never reach the cap keep it alive through all of them — along the TARGET
// axis into up to 16 consecutive ranges (never below 64 target poses) that different warps search concurrently
// (stf_split_merge_kernel re-applies the per-point cap across the ranges).  Returns the number of tiles that were split.
uint32_t split_heavy_tiles(hitl_ctx* ctx, const std::vector<uint32_t>& h_work, const std::vector<uint32_t>& h_open, const std::vector<uint32_t>& h_end, uint32_t lo,
                           uint32_t hi, uint64_t limit, std::vector<uint32_t>* est) {
  std::vector<uint32_t> scan, kl, jlo, jhi;
  const size_t cap0 = ctx->h_tile_scan.size() + 1024;
  est->clear(); est->reserve(cap0);
  scan.reserve(cap0); kl.reserve(cap0); jlo.reserve(cap0); jhi.reserve(cap0);
  std::vector<uint32_t> begin(ctx->n_poses + 1, 0);
  uint32_t n_split = 0, pose = 0;
  const uint32_t last_pose = ctx->n_poses ? ctx->n_poses - 1 : 0;
  for (uint32_t t = 0; t < (uint32_t)ctx->h_tile_scan.size(); ++t) {
    const uint32_t i = ctx->h_tile_scan[t], k0 = ctx->h_tile_kl[t] & 0xFFFFu, len = ctx->h_tile_kl[t] >> 16;
    const uint32_t a = ctx->h_tile_jlo[t], b = ctx->h_tile_jhi[t];
    while (pose <= i) begin[pose++] = (uint32_t)scan.size();
    const bool measured = t >= lo && t < hi;
    const uint64_t work = measured ? h_work[t - lo] : 0;
    uint32_t parts = 1, jparts = 1;
    uint32_t swept_end = std::min(b, last_pose);
    if (work > limit) {
      const bool is_unit = a != 0 || b != kFullRange;                       // units of a group keep their points (the group must stay aligned)
      if (!is_unit) while (parts < 8 && len / (parts * 2) >= 4 && work > limit * parts) parts *= 2;
      // The part of the target range the tile really swept: to its end when some point stayed below the cap, else up to the target
      // where its last point was capped (h_end).  Every range restarts the cap state, so a range redoes the (cheap) matching of the
      // points that cap early; the long tail that a few slow points cause is what gets divided.
      if (h_open[t - lo] == 0 && h_end[t - lo] > a) swept_end = std::min(swept_end, h_end[t - lo] - 1);
      if (ctx->target_splitting && work > limit * parts) {
        const uint32_t ja = a, jb = swept_end;
        const uint32_t span = jb >= ja ? jb - ja + 1 : 0;
        jparts = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(16, span / std::max(1u, ctx->min_target_span)), (work + limit * parts - 1) / (limit * parts));
        if (jparts < 1) jparts = 1;
      }
    }
    if (parts > 1 || jparts > 1) ++n_split;
    const uint32_t step = (len + parts - 1) / parts;
    const uint32_t w = (uint32_t)(work / (parts * jparts));   // children inherit an equal share of the measured work
    for (uint32_t p0 = 0; p0 < len; p0 += step) {
      if (jparts == 1) { scan.push_back(i); kl.push_back((k0 + p0) | (std::min(step, len - p0) << 16)); jlo.push_back(a); jhi.push_back(b); est->push_back(w); continue; }
      const uint32_t ja = a, jb = swept_end, span = jb - ja + 1;    // the ranges divide the swept part; the last one keeps the original end
      for (uint32_t q = 0; q < jparts; ++q) {
        const uint32_t qa = ja + (uint32_t)((uint64_t)span * q / jparts), qb = ja + (uint32_t)((uint64_t)span * (q + 1) / jparts) - 1;
        scan.push_back(i); kl.push_back((k0 + p0) | (std::min(step, len - p0) << 16));
        jlo.push_back(qa); jhi.push_back(q + 1 == jparts ? b : qb);           // the last range keeps the original (possibly open) end
        est->push_back(w);
      }
    }
  }
  while (pose <= ctx->n_poses) begin[pose++] = (uint32_t)scan.size();
  if (n_split) {
    ctx->h_tile_scan.swap(scan); ctx->h_tile_kl.swap(kl); ctx->h_tile_begin.swap(begin); ctx->h_tile_jlo.swap(jlo); ctx->h_tile_jhi.swap(jhi);
    ctx->tiling_splits += n_split;
  }
  return n_split;
}
}  // namespace hitl
using namespace hitl;

extern "C" int hitl_create(hitl_ctx** out, int device) {
  if (!out) return HITL_ERR_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0 || device < 0 || device >= count) { cudaGetLastError(); return HITL_ERR_CUDA; }
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return HITL_ERR_CUDA; }
  hitl_ctx* ctx = new hitl_ctx();
  ctx->device = device;
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  // scheduling knobs of the adaptive tiling (results never depend on them)
  if (const char* v = getenv("HITL_SPLIT_LIMIT_DIV")) ctx->split_limit_div = (uint32_t)std::max(1, atoi(v));
  if (const char* v = getenv("HITL_MIN_TARGET_SPAN")) ctx->min_target_span = (uint32_t)std::max(1, atoi(v));
  if (const char* v = getenv("HITL_ORDER_TWO_PASS")) ctx->order_two_pass = atoi(v) ? 1 : 0;
  if (const char* v = getenv("HITL_EM_CULL")) ctx->em_cull = atoi(v) != 0;
  if (const char* v = getenv("HITL_SEARCH_VARIANT")) ctx->search_variant = std::min(2, std::max(0, atoi(v)));   // occupancy / register trade-off (profiling)
  if (const char* v = getenv("HITL_SPLIT_ROUNDS")) ctx->max_split_rounds = (uint32_t)std::max(0, atoi(v));
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; cudaGetLastError(); return HITL_ERR_CUDA; }
  for (int i = 0; i < 4; ++i) { cudaEventCreate(&ctx->ev[i]); cudaEventCreate(&ctx->evx[i]); }
  for (int k = 0; k < HITL_K_COUNT; ++k) { cudaEventCreate(&ctx->kev[k][0]); cudaEventCreate(&ctx->kev[k][1]); }
  if (cudaMallocHost((void**)&ctx->h_pinned, 64 * sizeof(uint64_t)) != cudaSuccess) { hitl_destroy(ctx); cudaGetLastError(); return HITL_ERR_CUDA; }
  *out = ctx;
  return HITL_OK;
}

extern "C" void hitl_destroy(hitl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  hitl_comm_destroy(ctx);
  ctx->d_inc_key.release(); ctx->d_inc_ref.release(); ctx->d_inc_off.release(); ctx->d_cost_partial.release(); ctx->d_comm_cnt.release(); ctx->d_g_pi.release(); ctx->d_g_pj.release(); ctx->d_g_r.release(); ctx->d_g_J.release();
  ctx->d_off.release(); ctx->d_pts.release(); ctx->d_nrm.release(); ctx->d_aabb.release();
  ctx->d_tile_scan.release(); ctx->d_tile_k0.release(); ctx->d_tile_begin.release();
  ctx->d_tile_work.release(); ctx->d_tile_order.release(); ctx->d_tile_iota.release(); ctx->d_tile_keys.release(); ctx->d_sort_tmp.release();
  ctx->d_node_pm.release(); ctx->d_node_nn.release(); ctx->d_node_aos.release(); ctx->d_node_compact.release(); ctx->d_pack_k.release(); ctx->d_pack_idx.release();
  ctx->d_pose.release(); ctx->d_rec.release(); ctx->d_wbox.release(); ctx->d_gbox.release(); ctx->d_src.release(); ctx->d_grid.release(); ctx->d_occ.release(); ctx->d_occ_fine.release(); ctx->d_occ_dir.release(); ctx->d_nmax.release(); ctx->d_occ_mip.release(); ctx->d_moff.release(); ctx->d_tile_j.release(); ctx->d_tile_slot.release(); ctx->d_groups.release(); ctx->d_tile_open.release(); ctx->d_tile_end.release();
  ctx->d_raw_j.release(); ctx->d_raw_k.release(); ctx->d_raw_idx.release(); ctx->d_tile_cnt.release();
  ctx->d_srt_j.release(); ctx->d_srt_k.release(); ctx->d_srt_idx.release(); ctx->d_srt_flag.release();
  ctx->d_pose_cnt.release(); ctx->d_order_state.release(); ctx->d_counters.release(); ctx->d_pose_work.release();
  ctx->d_pair_i.release(); ctx->d_pair_j.release(); ctx->d_k.release(); ctx->d_idx.release(); ctx->d_pair_off.release();
  ctx->d_vo_sp.release(); ctx->d_vo_sk.release(); ctx->d_vo_tk.release();
  ctx->d_world.release(); ctx->d_poses_f.release(); ctx->d_em_pose.release(); ctx->d_em_idx.release(); ctx->d_em_xy.release();
  for (int f = 0; f < 2; ++f) { ctx->d_em_obs[f].release(); ctx->d_em_cnt[f].release(); ctx->d_em_setpose[f].release(); ctx->d_em_setoff[f].release(); ctx->d_em_slots[f].release(); }
  ctx->d_em_slotof.release();
  ctx->d_scan_state.release(); ctx->d_ticket.release(); ctx->d_fit_partial.release(); ctx->d_fit_out.release(); ctx->d_bp_poses.release(); ctx->d_bp_rw.release(); ctx->d_bp_tw.release(); ctx->d_bp_cs.release();
  ctx->d_odo.release(); ctx->d_hum_i.release(); ctx->d_hum_d.release();
  ctx->d_blk_i.release(); ctx->d_blk_j.release(); ctx->d_blk_k.release(); ctx->d_blk_idx.release(); ctx->d_blk_off.release();
  ctx->d_p2lg_pose.release(); ctx->d_p2lg_off.release(); ctx->d_p2lg_pts.release(); ctx->d_p2lg_n.release(); ctx->d_p2lg_o.release(); ctx->d_p2lg_v.release();
  ctx->d_p2l_pose.release(); ctx->d_p2l_pts.release(); ctx->d_p2l_n.release(); ctx->d_p2l_o.release(); ctx->d_p2l_v.release();
  ctx->d_r.release(); ctx->d_J.release(); ctx->d_neq.release(); ctx->d_hoff.release(); ctx->d_trig.release();
  if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
  for (int i = 0; i < 4; ++i) { if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]); if (ctx->evx[i]) cudaEventDestroy(ctx->evx[i]); }
  for (int k = 0; k < HITL_K_COUNT; ++k) for (int q = 0; q < 2; ++q) if (ctx->kev[k][q]) cudaEventDestroy(ctx->kev[k][q]);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* hitl_last_error(const hitl_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* hitl_stream(hitl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t hitl_launch_count(const hitl_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int hitl_sm_count(const hitl_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" int hitl_last_kernel_ms(hitl_ctx* ctx, int which, float* ms) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (which < 0 || which >= HITL_K_COUNT || !ms) return fail(ctx, HITL_ERR_ARG, "hitl_last_kernel_ms: bad argument");
  if (!ctx->kev_set[which]) return fail(ctx, HITL_ERR_STATE, "hitl_last_kernel_ms: that kernel has not been launched yet");
  HITL_CUDA(cudaEventSynchronize(ctx->kev[which][1]));
  HITL_CUDA(cudaEventElapsedTime(ms, ctx->kev[which][0], ctx->kev[which][1]));
  return HITL_OK;
}

// Page-locked host buffers for the caller's pose / scan / result arrays: DMA at PCIe rate and truly
// asynchronous copies.  Plain malloc'd buffers work everywhere too (staged by the driver).
extern "C" void* hitl_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void hitl_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int hitl_set_scans(hitl_ctx* ctx, uint32_t n_poses, const uint32_t* off, const float* pts_xy, const float* nrm_xy) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!off && n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_set_scans: null offsets");
  HITL_CUDA(cudaSetDevice(ctx->device));
  // Validate everything into locals first: a rejected call leaves the context exactly as it was.
  uint32_t max_scan = 0;
  for (uint32_t i = 1; i <= n_poses; ++i) {
    if (off[i] < off[i - 1]) return fail(ctx, HITL_ERR_ARG, "hitl_set_scans: offsets must be non-decreasing");
    max_scan = std::max(max_scan, off[i] - off[i - 1]);
  }
  if (n_poses && off[0] != 0) return fail(ctx, HITL_ERR_ARG, "hitl_set_scans: offsets must start at 0");
  if (max_scan > 65534) return fail(ctx, HITL_ERR_ARG, "hitl_set_scans: scans larger than 65534 points are not supported");
  const uint64_t n_points = n_poses ? off[n_poses] : 0;
  if (n_points && (!pts_xy || !nrm_xy)) return fail(ctx, HITL_ERR_ARG, "hitl_set_scans: null clouds");
  // Same scan partition as before (a re-upload of the same map): tiling and schedule hints stay valid.
  const bool same_partition = ctx->n_poses == n_poses && n_poses > 0 && ctx->h_off.size() == (size_t)n_poses + 1 &&
                              memcmp(ctx->h_off.data(), off, sizeof(uint32_t) * ((size_t)n_poses + 1)) == 0 && !ctx->h_tile_scan.empty();
  // ---- commit ----
  ctx->have_trees = false; ctx->have_stf = false; ctx->have_world = false; ctx->em_boxes_valid = false;
  // Residual blocks registered for the previous map carry pose / point indices that were range-checked against IT: they must be
  // registered again (hitl_eval / hitl_normal_eq would otherwise index the new, possibly smaller, map with stale indices).
  ctx->nb_odo = ctx->nb_human = ctx->nb_stf = ctx->nb_p2lg = ctx->nb_p2l = 0; ctx->stf_from_search = false;
  ctx->n_pairs = ctx->n_matches = 0; ctx->n_vo = 0; ctx->eval_valid = ctx->neq_valid = false; ctx->inc_valid = false;
  ctx->n_poses = n_poses;
  ctx->h_off.assign(n_poses + 1, 0);
  for (uint32_t i = 0; i <= n_poses && n_poses; ++i) ctx->h_off[i] = off[i];
  ctx->max_scan = max_scan;
  ctx->n_points = n_points;
  ctx->h_pts.clear(); ctx->h_nrm.clear();   // host copies are fetched lazily by hitl_build_kdtrees
  HITL_CUDA(ctx->d_off.ensure(n_poses + 1));
  const size_t cloud_elems = (replicated_upload_capacity(ctx, 8 * ctx->n_points) + 7) / 8;      // padded to world equal slices when the upload is sharded
  HITL_CUDA(ctx->d_pts.ensure(cloud_elems)); HITL_CUDA(ctx->d_nrm.ensure(cloud_elems));
  HITL_CUDA(cudaMemcpyAsync(ctx->d_off.p, ctx->h_off.data(), 4 * (size_t)(n_poses + 1), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->n_points) {
    int urc = replicated_upload(ctx, ctx->d_pts.p, pts_xy, 8 * ctx->n_points);
    if (urc) return urc;
    urc = replicated_upload(ctx, ctx->d_nrm.p, nrm_xy, 8 * ctx->n_points);
    if (urc) return urc;
  }
  // tiles: up to 32 consecutive points of one scan (the unit of work of the search; heavy tiles are split later)
  int rc = same_partition ? HITL_OK : hitl::build_tiling(ctx, 32);
  if (rc) return rc;
  rc = launch_scan_aabb(ctx);
  if (rc) return rc;
  ctx->h_aabb.assign(4 * (size_t)n_poses, 0.f);
  ctx->grid_valid = false;
  if (n_poses) HITL_CUDA(cudaMemcpyAsync(ctx->h_aabb.data(), ctx->d_aabb.p, 16 * (size_t)n_poses, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

// The same uploads for a multi-GPU job whose ranks all hold the map on their hosts (comm.cu: hitl_comm_init first): every rank passes the
// SAME arrays, each moves only its 1/world slice across PCIe, the slices travel over NVLink (ncclAllGather, in place).  Collective.
extern "C" int hitl_set_scans_sharded(hitl_ctx* ctx, uint32_t n_poses, const uint32_t* off, const float* pts_xy, const float* nrm_xy) {
  if (!ctx) return HITL_ERR_ARG;
  ctx->upload_sharded = true;
  const int rc = hitl_set_scans(ctx, n_poses, off, pts_xy, nrm_xy);
  ctx->upload_sharded = false;
  return rc;
}
extern "C" int hitl_set_kdtrees(hitl_ctx* ctx, const hitl_kdnode* nodes);
extern "C" int hitl_set_kdtrees_compact(hitl_ctx* ctx, const uint32_t* index_dim);
extern "C" int hitl_set_kdtrees_sharded(hitl_ctx* ctx, const hitl_kdnode* nodes) {
  if (!ctx) return HITL_ERR_ARG;
  ctx->upload_sharded = true;
  const int rc = hitl_set_kdtrees(ctx, nodes);
  ctx->upload_sharded = false;
  return rc;
}
extern "C" int hitl_set_kdtrees_compact_sharded(hitl_ctx* ctx, const uint32_t* index_dim) {
  if (!ctx) return HITL_ERR_ARG;
  ctx->upload_sharded = true;
  const int rc = hitl_set_kdtrees_compact(ctx, index_dim);
  ctx->upload_sharded = false;
  return rc;
}

namespace hitl {
// AoS hitl_kdnode (24 B) -> the resident SoA layout {float4 p|n, int32 index|dim<<31}; validates
// index / dim against the owning scan on the way (one thread per node, scan found by bisection).
__global__ void split_nodes_kernel(const hitl_kdnode* __restrict__ nodes, const uint32_t* __restrict__ off, uint32_t n_poses, uint64_t m,
                                   const float2* __restrict__ pts, const float2* __restrict__ nrm, float4* __restrict__ pm, float2* __restrict__ nn,
                                   uint32_t* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const hitl_kdnode nd = nodes[i];
  uint32_t lo = 0, hi = n_poses;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
  const uint32_t n = off[lo + 1] - off[lo];
  if (nd.index < 0 || (uint32_t)nd.index >= n || (nd.dim != 0 && nd.dim != 1)) { atomicOr(bad, 1u); return; }
  // a node IS a point of its scan (KDNodeValue = {point, normal, index}, kdtree.h:26-40): the occupancy levels of the search are built
  // from the resident scans, so a tree whose node differs from the point it names would make the culls inexact — refuse it
  const float2 p = pts[off[lo] + nd.index], v = nrm[off[lo] + nd.index];
  if (__float_as_uint(p.x) != __float_as_uint(nd.px) || __float_as_uint(p.y) != __float_as_uint(nd.py) || __float_as_uint(v.x) != __float_as_uint(nd.nx) ||
      __float_as_uint(v.y) != __float_as_uint(nd.ny))
    atomicOr(bad, 2u);
  pm[i] = make_float4(nd.px, nd.py, __uint_as_float(((uint32_t)nd.index & 0x7FFFFFFFu) | (nd.dim ? 0x80000000u : 0u)), 0.0f);
  nn[i] = make_float2(nd.nx, nd.ny);
}
__global__ void merge_nodes_kernel(const float4* __restrict__ pm, const float2* __restrict__ nn, uint64_t m, hitl_kdnode* __restrict__ nodes) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const float4 p = pm[i]; const float2 v = nn[i]; const uint32_t w = __float_as_uint(p.z);
  hitl_kdnode nd; nd.px = p.x; nd.py = p.y; nd.nx = v.x; nd.ny = v.y; nd.index = (int32_t)(w & 0x7FFFFFFFu); nd.dim = (int32_t)(w >> 31);
  nodes[i] = nd;
}
}  // namespace hitl

static int upload_trees(hitl_ctx* ctx, const hitl_kdnode* nodes) {
  const size_t m = ctx->n_points;
  HITL_CUDA(ctx->d_node_pm.ensure(m)); HITL_CUDA(ctx->d_node_nn.ensure(m));
  ctx->have_trees = false;
  if (m) {
    HITL_CUDA(ctx->d_node_aos.ensure((replicated_upload_capacity(ctx, sizeof(hitl_kdnode) * m) + sizeof(hitl_kdnode) - 1) / sizeof(hitl_kdnode))); HITL_CUDA(ctx->d_ticket.ensure(1));
    HITL_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, 4, ctx->stream));
    { const int urc = replicated_upload(ctx, ctx->d_node_aos.p, nodes, sizeof(hitl_kdnode) * m); if (urc) return urc; }
    split_nodes_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_node_aos.p, ctx->d_off.p, ctx->n_poses, m, ctx->d_pts.p, ctx->d_nrm.p,
                                                                            ctx->d_node_pm.p, ctx->d_node_nn.p, ctx->d_ticket.p);
    HITL_LAUNCH_CHECK("split_nodes_kernel");
    HITL_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_ticket.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
    const uint32_t badbits = *(const uint32_t*)ctx->h_pinned;
    if (badbits & 1u) return fail(ctx, HITL_ERR_ARG, "hitl_set_kdtrees: node index/dim out of range");
    if (badbits & 2u) return fail(ctx, HITL_ERR_ARG, "hitl_set_kdtrees: a node's point / normal differs from the scan point it names (trees must be trees of the uploaded scans)");
  }
  ctx->have_trees = true;
  ctx->grid_valid = false;      // the direction masks are built from the node normals
  return HITL_OK;
}

extern "C" int hitl_set_kdtrees(hitl_ctx* ctx, const hitl_kdnode* nodes) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_set_kdtrees: call hitl_set_scans first");
  if (!nodes && ctx->n_points) return fail(ctx, HITL_ERR_ARG, "hitl_set_kdtrees: null nodes");
  return upload_trees(ctx, nodes);
}

extern "C" int hitl_build_kdtrees(hitl_ctx* ctx) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_build_kdtrees: call hitl_set_scans first");
  if (ctx->tree_builder == 0) {
    ctx->have_trees = false;
    const int rc = build_kdtrees_device(ctx, &ctx->tree_exact_segments);
    if (rc == HITL_OK) { ctx->have_trees = true; ctx->grid_valid = false; }
    return rc;
  }
  std::vector<hitl_kdnode> nodes(ctx->n_points);
  const uint32_t n = ctx->n_poses;
  if (ctx->h_pts.size() != 2 * ctx->n_points) {   // the builder runs on the host: fetch the resident clouds
    ctx->h_pts.resize(2 * ctx->n_points); ctx->h_nrm.resize(2 * ctx->n_points);
    if (ctx->n_points) {
      HITL_CUDA(cudaMemcpyAsync(ctx->h_pts.data(), ctx->d_pts.p, 8 * ctx->n_points, cudaMemcpyDeviceToHost, ctx->stream));
      HITL_CUDA(cudaMemcpyAsync(ctx->h_nrm.data(), ctx->d_nrm.p, 8 * ctx->n_points, cudaMemcpyDeviceToHost, ctx->stream));
      HITL_CUDA(cudaStreamSynchronize(ctx->stream));
    }
  }
  unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
  if (n < 64) nt = 1;
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&, t]() {
      for (uint32_t i = t; i < n; i += nt) {
        const uint32_t o = ctx->h_off[i], c = ctx->h_off[i + 1] - o;
        build_flat_kdtree(ctx->h_pts.data() + 2 * (size_t)o, ctx->h_nrm.data() + 2 * (size_t)o, c, nodes.data() + o);
      }
    });
  for (auto& x : th) x.join();
  return upload_trees(ctx, nodes.data());
}

extern "C" int hitl_debug_set_tree_builder(hitl_ctx* ctx, int host) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  ctx->tree_builder = host ? 1 : 0;
  return HITL_OK;
}
extern "C" int hitl_debug_tree_stats(hitl_ctx* ctx, uint64_t* exact_segments) {
  if (!ctx || !exact_segments) return HITL_ERR_ARG;
  *exact_segments = ctx->tree_exact_segments;
  return HITL_OK;
}

extern "C" int hitl_kdtree_build_host(const float* pts_xy, const float* nrm_xy, uint32_t n, hitl_kdnode* out) {
  if (n && (!pts_xy || !nrm_xy || !out)) return HITL_ERR_ARG;
  if (n > 65534) return HITL_ERR_ARG;
  build_flat_kdtree(pts_xy, nrm_xy, n, out);
  return HITL_OK;
}

extern "C" int hitl_get_kdtrees(hitl_ctx* ctx, hitl_kdnode* out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_get_kdtrees: trees not built");
  const size_t m = ctx->n_points;
  if (!m) return HITL_OK;
  if (!out) return fail(ctx, HITL_ERR_ARG, "hitl_get_kdtrees: null output");
  HITL_CUDA(ctx->d_node_aos.ensure(m));
  merge_nodes_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_node_pm.p, ctx->d_node_nn.p, m, ctx->d_node_aos.p);
  HITL_LAUNCH_CHECK("merge_nodes_kernel");
  HITL_CUDA(cudaMemcpyAsync(out, ctx->d_node_aos.p, sizeof(hitl_kdnode) * m, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

// ---- compact tree format: 4 B per node instead of 24 ----------------------------------------------
// A flattened tree is a permutation of its scan's points plus one split bit per node; the points and normals are already resident
// (hitl_set_scans), so (index | dim << 31) per node in preorder is the whole tree: 14 MB instead of 86 MB across PCIe at c2.
namespace hitl {
__global__ void expand_nodes_kernel(const uint32_t* __restrict__ compact, const float2* __restrict__ pts, const float2* __restrict__ nrm,
                                    const uint32_t* __restrict__ off, uint32_t n_poses, uint64_t m, float4* __restrict__ pm, float2* __restrict__ nn,
                                    uint32_t* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t w = compact[i], index = w & 0x7FFFFFFFu;
  uint32_t lo = 0, hi = n_poses;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (off[mid] <= i) lo = mid; else hi = mid; }
  const uint32_t base = off[lo], n = off[lo + 1] - base;
  if (index >= n) { atomicOr(bad, 1u); return; }
  const float2 p = pts[base + index];
  pm[i] = make_float4(p.x, p.y, __uint_as_float(w), 0.0f);
  nn[i] = nrm[base + index];
}
__global__ void compact_nodes_kernel(const float4* __restrict__ pm, uint64_t m, uint32_t* __restrict__ compact) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) compact[i] = __float_as_uint(pm[i].z);
}
}  // namespace hitl

extern "C" int hitl_set_kdtrees_compact(hitl_ctx* ctx, const uint32_t* index_dim) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (ctx->h_off.empty()) return fail(ctx, HITL_ERR_STATE, "hitl_set_kdtrees_compact: call hitl_set_scans first");
  const size_t m = ctx->n_points;
  if (!index_dim && m) return fail(ctx, HITL_ERR_ARG, "hitl_set_kdtrees_compact: null nodes");
  HITL_CUDA(ctx->d_node_pm.ensure(m)); HITL_CUDA(ctx->d_node_nn.ensure(m));
  ctx->have_trees = false;
  if (m) {
    HITL_CUDA(ctx->d_node_compact.ensure((replicated_upload_capacity(ctx, 4 * m) + 3) / 4)); HITL_CUDA(ctx->d_ticket.ensure(4));
    HITL_CUDA(cudaMemsetAsync(ctx->d_ticket.p, 0, 4, ctx->stream));
    { const int urc = replicated_upload(ctx, ctx->d_node_compact.p, index_dim, 4 * m); if (urc) return urc; }
    expand_nodes_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_node_compact.p, ctx->d_pts.p, ctx->d_nrm.p, ctx->d_off.p, ctx->n_poses, m,
                                                                              ctx->d_node_pm.p, ctx->d_node_nn.p, ctx->d_ticket.p);
    HITL_LAUNCH_CHECK("expand_nodes_kernel");
    HITL_CUDA(cudaMemcpyAsync(ctx->h_pinned, ctx->d_ticket.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    HITL_CUDA(cudaStreamSynchronize(ctx->stream));
    if (*(const uint32_t*)ctx->h_pinned) return fail(ctx, HITL_ERR_ARG, "hitl_set_kdtrees_compact: node index out of range");
  }
  ctx->have_trees = true;
  ctx->grid_valid = false;
  return HITL_OK;
}

extern "C" int hitl_get_kdtrees_compact(hitl_ctx* ctx, uint32_t* index_dim_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (!ctx->have_trees) return fail(ctx, HITL_ERR_STATE, "hitl_get_kdtrees_compact: trees not built");
  const size_t m = ctx->n_points;
  if (!m) return HITL_OK;
  if (!index_dim_out) return fail(ctx, HITL_ERR_ARG, "hitl_get_kdtrees_compact: null output");
  HITL_CUDA(ctx->d_node_compact.ensure(m));
  compact_nodes_kernel<<<(uint32_t)((m + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_node_pm.p, m, ctx->d_node_compact.p);
  HITL_LAUNCH_CHECK("compact_nodes_kernel");
  HITL_CUDA(cudaMemcpyAsync(index_dim_out, ctx->d_node_compact.p, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

// ---- diagnostics ---------------------------------------------------------------------------------
namespace hitl {
__global__ void debug_sincos_kernel(const float* __restrict__ x, uint64_t n, float* __restrict__ s, float* __restrict__ c) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { s[i] = sinf_rn(x[i]); c[i] = cosf_rn(x[i]); }
}
__global__ void debug_relative_pose_kernel(const double* __restrict__ pose, uint32_t n, const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst,
                                           float* __restrict__ out6) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Aff2 a = pose_affine(pose[3 * src[i]], pose[3 * src[i] + 1], pose[3 * src[i] + 2]);
  const Aff2 b = pose_affine(pose[3 * dst[i]], pose[3 * dst[i] + 1], pose[3 * dst[i] + 2]);
  const Aff2 T = affine_mul(affine_inverse(b), a);
  float* o = out6 + 6 * (size_t)i;
  o[0] = T.m00; o[1] = T.m01; o[2] = T.m10; o[3] = T.m11; o[4] = T.tx; o[5] = T.ty;
}
}  // namespace hitl

extern "C" int hitl_debug_sincos(hitl_ctx* ctx, uint64_t n, const float* x, float* sin_out, float* cos_out) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n == 0) return HITL_OK;
  if (!x || !sin_out || !cos_out) return fail(ctx, HITL_ERR_ARG, "hitl_debug_sincos: null argument");
  TmpBuf<float> dx, ds, dc;
  HITL_CUDA(dx.ensure(n)); HITL_CUDA(ds.ensure(n)); HITL_CUDA(dc.ensure(n));
  HITL_CUDA(cudaMemcpyAsync(dx.p, x, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
  debug_sincos_kernel<<<(uint32_t)((n + 255) / 256), 256, 0, ctx->stream>>>(dx.p, n, ds.p, dc.p);
  HITL_LAUNCH_CHECK("debug_sincos_kernel");
  HITL_CUDA(cudaMemcpyAsync(sin_out, ds.p, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(cos_out, dc.p, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}

extern "C" int hitl_debug_relative_pose(hitl_ctx* ctx, const double* pose_array, uint32_t n_pairs, const uint32_t* src, const uint32_t* dst, float* out6) {
  if (!ctx) return HITL_ERR_ARG;
  HITL_DEVICE(ctx);
  if (n_pairs == 0) return HITL_OK;
  if (!pose_array || !src || !dst || !out6 || ctx->n_poses == 0) return fail(ctx, HITL_ERR_ARG, "hitl_debug_relative_pose: bad argument");
  for (uint32_t q = 0; q < n_pairs; ++q)
    if (src[q] >= ctx->n_poses || dst[q] >= ctx->n_poses) return fail(ctx, HITL_ERR_ARG, "hitl_debug_relative_pose: pose index out of range");
  TmpBuf<double> dp; TmpBuf<uint32_t> da, db; TmpBuf<float> dout;
  HITL_CUDA(dp.ensure(3 * (size_t)ctx->n_poses)); HITL_CUDA(da.ensure(n_pairs)); HITL_CUDA(db.ensure(n_pairs)); HITL_CUDA(dout.ensure(6 * (size_t)n_pairs));
  HITL_CUDA(cudaMemcpyAsync(dp.p, pose_array, 24 * (size_t)ctx->n_poses, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(da.p, src, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
  HITL_CUDA(cudaMemcpyAsync(db.p, dst, 4 * (size_t)n_pairs, cudaMemcpyHostToDevice, ctx->stream));
  debug_relative_pose_kernel<<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(dp.p, n_pairs, da.p, db.p, dout.p);
  HITL_LAUNCH_CHECK("debug_relative_pose_kernel");
  HITL_CUDA(cudaMemcpyAsync(out6, dout.p, 24 * (size_t)n_pairs, cudaMemcpyDeviceToHost, ctx->stream));
  HITL_CUDA(cudaStreamSynchronize(ctx->stream));
  return HITL_OK;
}
