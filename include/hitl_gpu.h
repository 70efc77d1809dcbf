/* hitl_gpu.h — C ABI of the B200 (sm_100a) back-end for HitL-SLAM's data-parallel hot path.
 *
 * Drop-in boundary: every entry point replaces one reference interface on the path
 * BASELINE.json:north_star names (paths relative to HitL-SLAM/src/ of ut-amrl/hitl-slam):
 *
 *   hitl_set_scans / hitl_build_kdtrees   JointOpt::BuildKDTrees            human_in_the_loop_slam/JointOptimization.cpp:514-537
 *                                         KDTree<float,2>::BuildKDTree      perception_tools/kdtree.cpp:37-139
 *   hitl_kd_query                         KDTree::FindNearestPointNormal    perception_tools/kdtree.cpp:141-197
 *                                         KDTree::FindNearestPoint          perception_tools/kdtree.cpp:220-273
 *   hitl_kd_neighbors                     KDTree::FindNeighborPoints        perception_tools/kdtree.cpp:199-218 (the node list, in push order)
 *   hitl_find_stf / hitl_get_stf          JointOpt::FindSTFCorrespondences  JointOptimization.cpp:561-642
 *   hitl_find_vo / hitl_get_vo            JointOpt::FindVisualOdometryCorrespondences  JointOptimization.cpp:432-468
 *   hitl_world_transform                  HitLSLAM::transformPointCloudsToWorldFrame   human_in_the_loop_slam/HitLSLAM.cpp:245-254
 *   hitl_verify_input                     HitLSLAM::verifyUserInput                    human_in_the_loop_slam/HitLSLAM.cpp:218-243
 *   hitl_em_inliers                       E-step of EMInput::AutomaticEndpointAdjustment  human_in_the_loop_slam/EMinput.cpp:207-218
 *   hitl_em_refit                         one E-step + M-step round: the above + EMInput::SegFitEM / segDistResidualEM  EMinput.cpp:107-191
 *   hitl_em_refit_chain                   the rounds of both strokes chained on the device  EMInput::AutomaticEndpointAdjustment EMinput.cpp:195-250
 *   hitl_em_assign                        EMInput::EstablishObservationSets EMinput.cpp:281-323
 *   hitl_set_*_blocks / hitl_eval         AutoDiffCostFunction<...>::Evaluate of the blocks added by
 *                                         AddSTFConstraints :539-559, AddOdometryConstraints :736-825,
 *                                         AddHumanConstraints :969-1054 (functors: residual_functors.h:768-848,
 *                                         1054-1133, 1299-1415; point-to-line :314-385, :557-622)
 *   hitl_normal_eq                        per-pose 3x3 J^T J / J^T r blocks of the same problem (input of the
 *                                         Gauss-Newton / LM step the host solver takes)
 *   hitl_comm_* / hitl_normal_eq_allreduce / hitl_gather_stf_blocks   the one exchange of a multi-GPU iteration (NCCL inside the
 *                                         boundary): OMP-over-source-poses of :575 becomes one rank per source range
 *
 * Conventions: plain pointers and sizes; the caller owns every host buffer; the context owns
 * all device memory; scans/trees are uploaded once per session, poses per call. Every call
 * returns a status (0 = ok) and never throws or aborts; hitl_last_error() describes the last
 * failure. There is no CPU fallback: without a CUDA device hitl_create fails.
 * One context per GPU per process; calls on one context must come from one thread at a time.
 */
#ifndef HITL_GPU_H_
#define HITL_GPU_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hitl_ctx hitl_ctx;

enum {
  HITL_OK = 0,
  HITL_ERR_ARG = 1,       /* bad argument */
  HITL_ERR_CUDA = 2,      /* CUDA runtime error (message in hitl_last_error) */
  HITL_ERR_STATE = 3,     /* call order: scans / trees / blocks not set */
  HITL_ERR_OVERFLOW = 4,  /* caller buffer too small */
  HITL_ERR_NCCL = 5       /* NCCL missing or a collective failed (hitl_comm_*, hitl_normal_eq_allreduce, hitl_gather_stf_blocks) */
};

/* ---- context ------------------------------------------------------------------------- */
int hitl_create(hitl_ctx** ctx, int device);
void hitl_destroy(hitl_ctx* ctx);
const char* hitl_last_error(const hitl_ctx* ctx);
/* cudaStream_t all kernels of this context are launched on (for external CUDA-event timing). */
void* hitl_stream(hitl_ctx* ctx);
/* Kernel launches issued by this context since creation (bench.py's gpu_launches). */
uint64_t hitl_launch_count(const hitl_ctx* ctx);
int hitl_sm_count(const hitl_ctx* ctx);
/* Duration (ms, CUDA events on the context's stream) of the LAST launch of one named kernel, for roofline accounting in bench.py:
 * the call that launched it has returned, i.e. the stream is idle.  HITL_ERR_STATE if that kernel has not run yet. */
enum { HITL_K_STF_SEARCH = 0, HITL_K_EVAL_STF = 1, HITL_K_EM_INLIERS = 2, HITL_K_EM_ASSIGN = 3, HITL_K_WORLD_TRANSFORM = 4, HITL_K_EM_FIT = 5, HITL_K_ALLREDUCE = 6, HITL_K_COUNT = 7 };
int hitl_last_kernel_ms(hitl_ctx* ctx, int which, float* ms);
/* Page-locked host memory for the caller's buffers (optional: any host pointer is accepted by every
 * call; pinned ones move at PCIe rate).  NULL on failure. */
void* hitl_host_alloc(size_t bytes);
void hitl_host_free(void* p);

/* ---- scans and KD-trees ----------------------------------------------------------------- */
/* Robot-frame point and normal clouds of all poses, concatenated; scan i owns
 * [scan_offsets[i], scan_offsets[i+1]). float2 AoS on the host, kept SoA-of-float2 in HBM. */
int hitl_set_scans(hitl_ctx* ctx, uint32_t n_poses, const uint32_t* scan_offsets, const float* pts_xy,
                   const float* nrm_xy);

/* One KD node, preorder-flattened: a subtree of n nodes rooted at position p has its left
 * child (n/2 nodes) at p+1 and its right child (n-1-n/2 nodes) at p+1+n/2. */
typedef struct {
  float px, py, nx, ny;
  int32_t index; /* index of the point inside its scan */
  int32_t dim;   /* splitting dimension, 0 or 1 */
} hitl_kdnode;

/* Build every scan's tree with the reference's algorithm (max-variance split, std::sort,
 * median n/2) and make it resident.  Built on the device, all scans at once, with the reference's exact
 * tree shape (equal keys included); hitl_debug_set_tree_builder(ctx, 1) selects the threaded host builder. */
int hitl_build_kdtrees(hitl_ctx* ctx);
/* Or adopt trees built elsewhere (same layout, concatenated by scan_offsets).  Every node must carry exactly the point and normal of
 * the scan entry its `index` names (bit for bit — a KD-tree of a scan holds that scan's points); HITL_ERR_ARG otherwise. */
int hitl_set_kdtrees(hitl_ctx* ctx, const hitl_kdnode* nodes);
int hitl_get_kdtrees(hitl_ctx* ctx, hitl_kdnode* nodes_out);
/* Compact form of the same trees: one 32-bit word per node in preorder, index | dim << 31.  The node's point and normal are the
 * scan's entries at `index`, which hitl_set_scans already made resident, so this is the whole tree at 4 B per node (24 B in hitl_kdnode). */
int hitl_set_kdtrees_compact(hitl_ctx* ctx, const uint32_t* index_dim);
int hitl_get_kdtrees_compact(hitl_ctx* ctx, uint32_t* index_dim_out);

/* The same builder for one scan without a context (pure host code; used by tests and by callers
 * that want to inspect a tree).  out must hold n nodes. */
int hitl_kdtree_build_host(const float* pts_xy, const float* nrm_xy, uint32_t n, hitl_kdnode* out);

/* Batched tree queries against scan `scan`.  mode 0 = FindNearestPointNormal, 1 = FindNearestPoint,
 * 2 = FindNeighborPoints (dist_out unused, index_out = number of neighbours).
 * index_out = -1 when the reference would leave neighbor_node untouched. */
int hitl_kd_query(hitl_ctx* ctx, uint32_t scan, uint32_t n_queries, const float* q_xy, float threshold, int mode,
                  float* dist_out, int32_t* index_out);

/* KDTree::FindNeighborPoints (kdtree.cpp:199-218) with its result list: for every query, count_out = number of nodes with
 * |node - q| < threshold and index_out[q * cap ...] = the scan point indices of the first `cap` of them in the reference's push order
 * (node, left subtree, right subtree); unused slots are -1.  cap = 0 counts only (index_out may be NULL). */
int hitl_kd_neighbors(hitl_ctx* ctx, uint32_t scan, uint32_t n_queries, const float* q_xy, float threshold, uint32_t cap,
                      int32_t* index_out, uint32_t* count_out);

/* ---- scan-to-scan correspondence search ------------------------------------------------ */
typedef struct {
  float point_match_threshold;          /* kPointMatchThreshold, config 0.15 */
  float min_cosine_angle;               /* cos(kMaxStfAngleError) computed once by the caller */
  int32_t max_correspondences_per_point; /* kMaxCorrespondencesPerPoint, config 6 */
  uint32_t num_skip_readings;           /* config 1 */
  uint32_t min_inter_pose_correspondence; /* kMinInterPoseCorrespondence = 10: keep pairs with MORE matches */
  uint32_t disable_culling;             /* debug: 1 = visit every pair exactly as the CPU loop does */
} hitl_stf_opts;

typedef struct {
  uint64_t n_pairs;       /* kept ordered pose pairs (= residual blocks) */
  uint64_t n_matches;     /* correspondences in kept pairs */
  uint64_t n_raw_matches; /* matches before the > min_inter_pose_correspondence filter */
  uint64_t n_queries;     /* KD queries the reference semantics execute (cap-skipped ones excluded) */
  uint64_t n_traversals;  /* queries that actually walked a tree on the GPU (rest proven empty by exact culling) */
  uint64_t n_tile_pairs;  /* (32-point source tile, target pose) pairs that survived the world-frame box test */
  float ms_search;        /* device time of the search kernel alone (CUDA events on the ctx stream) */
  float ms_total;         /* device time of the whole call: pose prep + search + ordering/compaction */
  uint64_t n_coarse_pass; /* (point, target) items that passed the coarse occupancy level (candidates of the fine level) */
  uint64_t n_in_radius;   /* tree walks that found a node inside the radius (the rest proved "no neighbour" the slow way) */
  uint64_t sum_tile_cycles; /* SM cycles the resident warps spent inside tiles, summed over all tiles of this call */
  uint64_t max_tile_cycles; /* ... and of the single most expensive tile: the kernel's critical path (a tile is a sequential loop) */
  uint32_t n_tiles;       /* work units (runs of <= 32 source points) the search kernel scheduled in this call */
  uint32_t n_tiles_next;  /* ... and after the adaptive split of heavy tiles that this call's measurements triggered */
  uint64_t n_gate_fail;   /* walks that found an in-radius node whose normal failed the angle gate (an executed query without a match) */
  uint64_t n_over_cap;    /* walks that matched but whose point had filled its cap earlier in the same batch (speculative, dropped) */
  uint64_t n_dir_culled;  /* queries whose walk was skipped because no node normal near them can pass the angle gate (exact prefilter) */
} hitl_stf_info;

/* Correspondences between source poses [src_lo, src_hi) ∩ [min_pose, max_pose] and all target
 * poses in [min_pose, max_pose].  src_lo = 0, src_hi = UINT32_MAX gives the reference call;
 * a narrower source range is one shard of it (results of consecutive shards concatenate to
 * the full result).  pose_array: x, y, theta doubles per pose (JointOpt::pose_array_).
 * Results stay resident for hitl_set_stf_blocks_from_search / hitl_get_stf. */
int hitl_find_stf(hitl_ctx* ctx, const double* pose_array, uint32_t min_pose, uint32_t max_pose, uint32_t src_lo,
                  uint32_t src_hi, const hitl_stf_opts* opts, hitl_stf_info* info);
/* CSR copy-out in the reference's order (pose_index0 asc, pose_index1 asc, points0 index asc):
 * pair_i/pair_j [n_pairs], pair_off [n_pairs+1], k/idx [n_matches]. */
int hitl_get_stf(hitl_ctx* ctx, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx);
/* The same with 16-bit point indices (a scan holds at most 65534 points): half the bytes of the two largest arrays across PCIe. */
int hitl_get_stf16(hitl_ctx* ctx, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint16_t* k, uint16_t* idx);

/* Load-balancing feedback: SM cycles the last hitl_find_stf spent on each source pose (0 outside the
 * searched source range).  work_per_pose holds n_poses entries.  Multi-GPU callers sum these over the
 * ranks and cut the next call's source ranges at equal work (hitl_slam_b200/sharding.py). */
int hitl_get_stf_work(hitl_ctx* ctx, uint64_t* work_per_pose);

/* Consecutive-pose matching with the Euclidean query. n_out = number of correspondences. */
int hitl_find_vo(hitl_ctx* ctx, const double* pose_array, int32_t min_pose, int32_t max_pose,
                 const hitl_stf_opts* opts, uint64_t* n_out);
int hitl_get_vo(hitl_ctx* ctx, uint32_t* source_pose, uint32_t* source_point, uint32_t* target_point);

/* ---- world-frame clouds and EM assignment ------------------------------------------------ */
/* world = Rotation2Df(theta) * p + t for every point; poses_xyt: 3 floats per pose (poses_).
 * Result stays resident for the EM calls; world_xy_out may be NULL. */
int hitl_world_transform(hitl_ctx* ctx, const float* poses_xyt, float* world_xy_out);
int hitl_set_world_clouds(hitl_ctx* ctx, const float* world_xy);

/* Input verification over the resident world clouds: points_verified = number of selected points (<= 8, sel_xy = x, y per point)
 * that have a world point with (w - s).norm() < threshold (the reference uses 0.05f), or 0 when either stroke of a 4-point input is
 * degenerate (sel[0] == sel[1] or sel[2] == sel[3]).  seen_mask (may be NULL): bit i = selected point i was seen. */
int hitl_verify_input(hitl_ctx* ctx, uint32_t n_selected, const float* sel_xy, float threshold, uint32_t* points_verified, uint32_t* seen_mask);

/* E-step: every world point with DistanceToLineSegment(seg) < threshold, in (pose, index) order.
 * seg = {p0x, p0y, p1x, p1y}. out_* may be NULL (count only); cap = capacity of the out arrays. */
int hitl_em_inliers(hitl_ctx* ctx, const float seg[4], double threshold, uint64_t cap, uint32_t* out_pose,
                    uint32_t* out_idx, float* out_xy, uint64_t* n_out);

/* One EM round of EMInput::AutomaticEndpointAdjustment entirely on the device (EMinput.cpp:195-250): the E-step above (inliers of
 * seg_in within inlier_threshold stay resident) and the M-step, SegFitEM's one-parameter Levenberg-Marquardt fit of the stroke's
 * direction to those inliers (EMinput.cpp:107-191: fixed midpoint and length, theta_0 = acos(|dx| / length), Ceres defaults,
 * <= max_iterations iterations; the reference passes 25), as ONE cooperative kernel whose every evaluation is a grid-wide
 * deterministic reduction.  seg_out = the refit endpoints as the reference rounds them to float.  Nothing but the 64-byte
 * result crosses PCIe. */
typedef struct {
  double theta, initial_cost, final_cost;
  uint64_t n_inliers;
  int32_t iterations, evaluations, termination; /* termination: 1 = a Ceres convergence test fired, 0 = iteration cap */
  float ms;                                     /* device time of E-step + M-step */
} hitl_em_fit_info;
int hitl_em_refit(hitl_ctx* ctx, const float seg_in[4], double inlier_threshold, int32_t max_iterations, float seg_out[4],
                  hitl_em_fit_info* info);
/* `rounds` (1..4) EM rounds of `n_strokes` (1..2) independent strokes with ONE host wait: round r of a stroke runs hitl_em_refit's
 * E-step + M-step on the stroke that round r-1 left in device memory (round 0: segs_in[4 * s ..]).  Results are indexed
 * slot = r * n_strokes + s: segs_out[4 * slot ..], info[slot] (info may be NULL; info[].ms = device time of the whole chain).
 * The loop of EMInput::AutomaticEndpointAdjustment (EMinput.cpp:195-250) stops a stroke once both endpoints moved <= 0.05 m;
 * a caller applies that rule to the returned sequence and ignores the rounds past convergence — every round it keeps is bit for
 * bit the round a one-call-per-round loop would have produced.  hitl_em_refit is the (1 stroke, 1 round) case.
 * E-steps after the first on the same world clouds skip, unread, the 2048-point chunks whose bounding box lies out of the
 * stroke's reach (exact: such a chunk has no inlier). */
int hitl_em_refit_chain(hitl_ctx* ctx, uint32_t n_strokes, const float* segs_in, double inlier_threshold, int32_t max_iterations,
                        uint32_t rounds, float* segs_out, hitl_em_fit_info* info);

/* Observation sets of both strokes: segs = {a0, a1, b0, b1} as 8 floats.  A pose is kept for a
 * stroke when MORE than min_obs of its points are within threshold (reference: 5).
 * For stroke f: set_pose[f][s], set_off[f][s..s+1] into obs[f][]. Capacities: n_poses, n_poses+1,
 * total points. */
int hitl_em_assign(hitl_ctx* ctx, const float segs[8], double threshold, uint32_t min_obs, uint32_t n_sets[2],
                   uint32_t* set_pose0, uint64_t* set_off0, uint32_t* obs0, uint32_t* set_pose1, uint64_t* set_off1,
                   uint32_t* obs1);

/* ---- residual blocks --------------------------------------------------------------------- */
/* Block order of hitl_eval = [odometry | human | stf | p2l_glob | p2l], each in the order given. */
/* STF blocks = the kept pairs of the last hitl_find_stf on this context (AddSTFConstraints). */
int hitl_set_stf_blocks_from_search(hitl_ctx* ctx, float laser_std_dev, float point_point_correlation_factor);
/* Or from an explicit CSR list. */
int hitl_set_stf_blocks(hitl_ctx* ctx, uint64_t n_pairs, const uint32_t* pair_i, const uint32_t* pair_j,
                        const uint64_t* pair_off, const uint32_t* k, const uint32_t* idx, float laser_std_dev,
                        float point_point_correlation_factor);
/* Odometry blocks between poses b and b+1: 9 floats each = axis_transform (row-major 2x2), radial,
 * tangential, angular std-dev, radial_translation, rotation (PoseConstraint's members). */
int hitl_set_odometry_blocks(hitl_ctx* ctx, uint32_t n_blocks, const float* consts9);
/* Human blocks: type_pose = {CorrectionType, constrained pose} per block, targets = {x, y, theta,
 * penalty_dir} doubles per block (what AddHumanConstraints passes to the functor constructors). */
int hitl_set_human_blocks(hitl_ctx* ctx, uint32_t n_blocks, const int32_t* type_pose, const double* targets4);
/* Point-to-line-glob blocks (one pose per block, CSR over points) and single point-to-line blocks. */
int hitl_set_p2l_glob_blocks(hitl_ctx* ctx, uint32_t n_blocks, const uint32_t* blk_pose, const uint64_t* blk_off,
                             const float* pts_xy, const float* line_normal_xy, const float* line_offset,
                             const uint8_t* valid, float std_dev, float correlation_factor);
int hitl_set_p2l_blocks(hitl_ctx* ctx, uint64_t n_blocks, const uint32_t* pose_idx, const float* pts_xy,
                        const float* line_normal_xy, const float* line_offset, const uint8_t* valid, float std_dev,
                        float correlation_factor);

typedef struct {
  uint64_t n_odometry, n_human, n_stf, n_p2l_glob, n_p2l;
  uint64_t n_residuals;  /* doubles in r_out */
  uint64_t n_jacobian;   /* doubles in J_out */
} hitl_eval_layout;
/* r_out / J_out layout per block kind (row-major [residual][param], pose blocks side by side):
 *   odometry : r 3, J 18 = [3x3 wrt pose b | 3x3 wrt pose b+1]
 *   human    : r 3 (unused rows 0), J 9 = [3x3 wrt constrained pose] (unused rows 0)
 *   stf      : r 2, J 12 = [2x3 wrt pose_index0 | 2x3 wrt pose_index1]
 *   p2l_glob : r 1, J 3 ;  p2l : r 1, J 3 */
int hitl_eval_layout_get(hitl_ctx* ctx, hitl_eval_layout* layout);
/* precision: 0 = FP64 (<=1e-9 rel. vs the CPU functors), 1 = FP32 mode (<=1e-5): projections and
 * derivative sums in FP32; pose trigonometry and the cancelling world-frame point difference of
 * the STF blocks stay FP64. J_out may be NULL. */
int hitl_eval(hitl_ctx* ctx, const double* pose_array, int precision, double* r_out, double* J_out, float* ms_out);
/* Normal-equation blocks of the whole problem at pose_array: H_diag [n_poses x 9] (J^T J, row-major
 * 3x3 per pose), g [n_poses x 3] (J^T r), H_off [n_binary_blocks x 9] (J_a^T J_b per odometry block,
 * then per stf block), cost = 1/2 sum r^2. Any output may be NULL. Results also stay resident
 * (hitl_normal_eq_device) for a device-side all-reduce. */
int hitl_normal_eq(hitl_ctx* ctx, const double* pose_array, double* H_diag, double* g, double* H_off, double* cost,
                   float* ms_out);
/* on = 1: hitl_normal_eq / hitl_normal_eq_allreduce become run-to-run reproducible to the last bit on one context: instead of FP64
 * atomics, every pose GATHERS the J^T J / J^T r of its blocks in a fixed order (per-pose incidence lists, rebuilt when blocks are
 * registered) and the cost is a fixed-shape sum.  Slightly slower (per-block r and J are written and read back once); r and J of every
 * block are resident afterwards as after hitl_eval.  Default 0. */
int hitl_set_deterministic(hitl_ctx* ctx, int on);
/* Device pointer + length (doubles) of the packed [H_diag | g | cost] buffer of the last hitl_normal_eq. */
int hitl_normal_eq_device(hitl_ctx* ctx, void** dev_ptr, uint64_t* n_doubles);

/* ---- multi-GPU exchange -------------------------------------------------------- */
/* One context per GPU (one process per GPU, or several contexts in one process driven by one host thread each).  The search
 * shards by source pose with no collective (hitl_find_stf's src_lo / src_hi); scans and trees are replicated.  The context owns
 * an NCCL communicator, bound at run time (libnccl.so.2); every failure of the layer returns HITL_ERR_NCCL.
 * hitl_comm_unique_id: on ONE rank, fills HITL_COMM_ID_BYTES (an ncclUniqueId) to be handed to the others out of band
 * (MPI / torch.distributed / a file).  hitl_comm_init is collective: every rank of the job calls it with the same id. */
#define HITL_COMM_ID_BYTES 128
int hitl_comm_unique_id(void* id_out);
int hitl_comm_init(hitl_ctx* ctx, const void* nccl_unique_id, int rank, int world);
int hitl_comm_destroy(hitl_ctx* ctx);
int hitl_comm_info(const hitl_ctx* ctx, int* rank, int* world, int* nccl_version);
/* The per-iteration collective (PostHumanOptimization's Gauss-Newton / LM loop, JointOptimization.cpp:1192-1208): when pose_array is
 * not NULL, evaluates the normal equations of THIS rank's blocks as hitl_normal_eq does, then — same stream, no host
 * synchronisation in between — one ncclAllReduce(sum, f64) over the packed resident buffer [H_diag n_poses x 9 | g n_poses x 3 |
 * cost], in place.  Afterwards every rank holds the whole problem's H_diag / g / cost (also resident: hitl_normal_eq_device);
 * H_off stays with the rank that owns each block.  Without a communicator (single GPU) it is hitl_normal_eq. */
/* Uploads for a job whose ranks all hold the same map on their hosts: identical arguments on every rank (collective); each rank moves
 * only its 1/world slice across PCIe and the slices are exchanged over NVLink (ncclAllGather, in place).  Without a communicator they
 * are the plain calls. */
int hitl_set_scans_sharded(hitl_ctx* ctx, uint32_t n_poses, const uint32_t* scan_offsets, const float* pts_xy, const float* nrm_xy);
int hitl_set_kdtrees_sharded(hitl_ctx* ctx, const hitl_kdnode* nodes);
int hitl_set_kdtrees_compact_sharded(hitl_ctx* ctx, const uint32_t* index_dim);
int hitl_normal_eq_allreduce(hitl_ctx* ctx, const double* pose_array, double* H_diag, double* g, double* cost, float* ms_out);
/* The Ceres-on-the-host feed: after hitl_eval (with Jacobians) on every rank, gathers the STF blocks of all ranks to `root` in rank
 * order (= the reference's block order when the ranks own ascending source ranges): pair_i / pair_j [total], r [total x 2],
 * J [total x 12] = 14 doubles per block, what SizedCostFunction<2,3,3>::Evaluate hands to Ceres.  n_blocks_per_rank [world] is
 * filled on every rank; the other outputs only on root (may be NULL elsewhere).  HITL_ERR_OVERFLOW when total > cap_blocks. */
int hitl_gather_stf_blocks(hitl_ctx* ctx, int root, uint64_t cap_blocks, uint64_t* n_blocks_per_rank, uint32_t* pair_i, uint32_t* pair_j,
                           double* r, double* J);

/* ---- COP-SLAM back-propagation (between EM and the joint optimisation) --------------------- */
/* The pose update of Backprop::BackPropagateError (Backprop.cpp:170-199), with the host loops' float operation sequence per
 * pose: for i in [lo, hi): rotate poses (i, hi] by rot_weights[i-lo] * theta about pose i (pose i's angle moves too), then
 * with trans = destination - pose[hi]: poses (i, hi] += trans_weights[i-lo] * trans.  poses_xyt: n_poses float triples on the
 * host, in/out; weights: hi - lo entries used.  ms_out (optional): device time of the kernel. */
int hitl_backprop_poses(hitl_ctx* ctx, uint32_t n_poses, float* poses_xyt, uint32_t lo, uint32_t hi, const float* rot_weights,
                        const float* trans_weights, float theta, const float* destination_xy, float* ms_out);

/* ---- diagnostics ------------------------------------------------------------------------- */
/* Device evaluation of the library's sinf/cosf (bit-identical to glibc 2.39's x86-64 FMA variant;
 * csrc/hitl_math.h) and of RelativePoseTransform (JointOptimization.cpp:296-305) for parity tests. */
/* SM cycles / 64 the last hitl_find_stf spent on each 32-point tile (tiles outside the searched range keep
 * stale values); profiling aid for the tile scheduler. */
int hitl_debug_tile_work(hitl_ctx* ctx, uint32_t cap, uint32_t* work_out, uint32_t* n_tiles_out);
/* Descriptors of the current tiles: source scan, first point | length << 16, target range [jlo, jhi] (0 .. 0xFFFFFFFF = all),
 * and the number of points the last search left below the cap at the end of the tile's range. Any pointer may be NULL. */
int hitl_debug_tile_desc(hitl_ctx* ctx, uint32_t cap, uint32_t* scan, uint32_t* k0_len, uint32_t* jlo, uint32_t* jhi, uint32_t* open_points);
/* Re-cuts every scan into tiles of at most max_len (1..32) points and sets the automatic splitting of heavy tiles:
 * adaptive 0 = off, 1 = along the points and along the target axis, 2 = along the points only.  target_parts > 1
 * additionally cuts EVERY tile into that many consecutive target ranges that are searched concurrently and merged
 * under the per-point cap.  The tiling is a scheduling choice; parity tests use this to prove it. */
int hitl_debug_set_tiling(hitl_ctx* ctx, uint32_t max_len, int adaptive, uint32_t target_parts);
/* Switches the second (fine, cell = threshold / 4) level of the occupancy cull (bit 0 of `on`) and the angle-gate direction
 * prefilter (bit 1 SET switches it OFF: on = 1 is the default, everything on; 0 = no fine level; 3 = no prefilter; 2 = neither;
 * bit 2 SET additionally switches the tile-box vs scan cull off, e.g. 5);
 * the bitmaps are rebuilt by the next search.  Culling is result-preserving; parity tests compare the settings and disable_culling = 1. */
int hitl_debug_set_fine_occupancy(hitl_ctx* ctx, int on);
/* E-step chunk cull (see hitl_em_refit_chain): on = 0 makes every E-step read every chunk.  Result-preserving; parity tests compare. */
int hitl_debug_set_em_cull(hitl_ctx* ctx, int on);
/* Occupancy / register trade-off of the search kernel: 0 = 16 CTAs per SM (32 registers), 1 = 12 (40), 2 = 10 (48);
 * smem_carveout_pct = preferred shared-memory carve-out of the unified L1 (percent, -1 = driver default). */
int hitl_debug_set_search_variant(hitl_ctx* ctx, int variant, int smem_carveout_pct);
/* hitl_build_kdtrees builds on the device by default (all scans at once, reference-identical shape); host = 1 selects the
 * threaded host builder instead (same trees; parity tests compare the two node for node).  hitl_debug_tree_stats: number of
 * tree segments of the last device build that contained equal keys and were re-sorted with the exact std::sort emulation. */
int hitl_debug_set_tree_builder(hitl_ctx* ctx, int host);
int hitl_debug_tree_stats(hitl_ctx* ctx, uint64_t* exact_segments);
int hitl_debug_sincos(hitl_ctx* ctx, uint64_t n, const float* x, float* sin_out, float* cos_out);
int hitl_debug_relative_pose(hitl_ctx* ctx, const double* pose_array, uint32_t n_pairs, const uint32_t* src,
                             const uint32_t* dst, float* out6);

#ifdef __cplusplus
}
#endif
#endif /* HITL_GPU_H_ */
