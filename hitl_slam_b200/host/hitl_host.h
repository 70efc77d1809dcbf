// hitl_host.h — C entry points of the host mirror library (libhitl_host.so): file formats and
// session-level helpers that sit above the C ABI of include/hitl_gpu.h.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
// .stfs.covars (loadPoseGraph, HitLSLAM_main.cpp:192-300 / SaveStfsandCovars, vector_mapping_main.cpp:1855-1928)
void* hitl_host_load_pose_graph(const char* path, uint64_t* n_poses, uint64_t* n_points);
// the same through a binary cache of the parsed graph (default cache_path: path + ".hitlcache"; stale caches are detected and rewritten)
void* hitl_host_load_pose_graph_cached(const char* path, const char* cache_path, uint64_t* n_poses, uint64_t* n_points, int* from_cache);
void hitl_host_pose_graph_get(void* h, float* poses, float* cov, uint32_t* off, float* pts, float* nrm);
void hitl_host_pose_graph_free(void* h);
int hitl_host_save_stfs_covars(const char* path, const char* map_name, double timestamp, uint32_t n_poses, const float* poses_xyt,
                               const float* cov9, const uint32_t* off, const float* obs_world_xy, const float* nrm_world_xy);
int hitl_host_save_poses(const char* path, uint32_t n_poses, const float* poses_xyt);
#ifdef __cplusplus
}
#endif
#ifdef __cplusplus
extern "C" {
#endif
float hitl_host_sinf(float x);
float hitl_host_cosf(float x);
uint64_t hitl_host_sincos_mismatches(uint64_t first, uint64_t count, uint64_t stride);
void hitl_host_relative_pose(const double* pose_array, uint32_t src, uint32_t dst, float* out6);
uint64_t hitl_host_stdsort_mismatches(const float* keys, uint32_t n, uint32_t* out_std, uint32_t* out_restated, uint32_t* heap_sorts);
#ifdef __cplusplus
}
#endif

// ---- session over the C++ mirror classes (host_capi.cpp; JointOpt / EMInput on one hitl_ctx) ----
#ifdef __cplusplus
extern "C" {
#endif
void* hitl_host_session_create(void* ctx);
void hitl_host_session_destroy(void* s);
const char* hitl_host_session_error(void* s);
int hitl_host_session_set_map(void* s, uint32_t n_poses, const float* poses_xyt, const uint32_t* off, const float* pts_xy, const float* nrm_xy);
int hitl_host_session_set_poses(void* s, const float* poses_xyt);
int hitl_host_session_get_poses(void* s, float* poses_xyt, double* pose_array);
int hitl_host_session_world_transform(void* s, int keep_host_copy);
int hitl_host_session_em_run(void* s, int correction_type, float sel_xy[8], int32_t info[6]);
int hitl_host_session_em_poses(void* s, int32_t* corrected, int32_t* anchor);
int hitl_host_session_add_constraints_from_em(void* s, uint32_t* n_out);
int hitl_host_session_add_constraints(void* s, uint32_t n, const int32_t* ids3, const float* deltas4);
int hitl_host_session_clear_constraints(void* s);
int hitl_host_session_verify_input(void* s, const float sel_xy[8], uint32_t* points_verified);
int hitl_host_app_exp_correct(int correction_type, const float sel_xy[8], uint32_t n_poses, float* poses_xyt, uint32_t n_corrected, const int32_t* corrected, float C3[3]);
int hitl_host_constraint_targets(int correction_type, const float sel_xy[8], uint32_t n_poses, const float* poses_xyt, uint32_t n_corrected, const int32_t* corrected,
                                 uint32_t n_anchor, const int32_t* anchor, int32_t* ids3, float* deltas4);
int hitl_host_backprop(void* ctx, uint32_t n_poses, float* poses_xyt, float* cov9, int32_t lo, int32_t hi, const float C3[3], float* device_ms);
int hitl_host_session_correct(void* s, int correction_type, float sel_xy[8], float* cov9, int solve, int32_t info[8], double ms[5], double summary[6]);
int hitl_host_session_solver_options(void* s, int which, int max_iterations, double function_tolerance, double gradient_tolerance, double parameter_tolerance,
                                     int precision, int verbose);
int hitl_host_session_joint_opt_run(void* s, int post, double summary[6]);
int hitl_host_session_solve(void* s, int mode, double summary[6]);
int hitl_host_session_copy_params(void* s);
int hitl_host_session_find_stf(void* s, uint64_t min_pose, uint64_t max_pose, uint64_t counts[3]);
int hitl_host_session_get_stf(void* s, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx);
int hitl_host_session_gradient(void* s, uint64_t cap, double* gradient, uint64_t* n_out, uint64_t jac_dims[3]);
int hitl_host_session_evaluate_block(void* s, int with_stf, uint64_t block, const double* pose_array_override, int32_t* nres, int32_t* nblocks,
                                     double* residuals, double* jac0, double* jac1, uint64_t* n_total_blocks);
int hitl_host_seg_fit_em(const double p1[2], const double p2[2], const double* data, int size, float out4[4]);
void hitl_host_odometry_consts(const float* poses_xyt, uint32_t n_poses, float* consts9);
void hitl_host_human_targets(const float* poses_xyt, uint32_t n_poses, uint32_t n, const int32_t* ids3, const float* deltas4, double* targets4);
int hitl_host_solver_selftest(double x[4], int max_iterations, int hold_x1, int force_cg, double out[4]);
int hitl_host_solver_chain_selftest(double* x, int n, int mode, double out[4]);
int hitl_host_load_log(const char* path, uint32_t cap_entries, uint32_t cap_points, int32_t* types, int32_t* undone, int32_t* npts, float* pts_xy,
                       uint32_t* n_entries, uint32_t* n_points);
int hitl_host_save_log(const char* path, uint32_t n_entries, const int32_t* types, const int32_t* undone, const int32_t* npts, const float* pts_xy);
#ifdef __cplusplus
}
#endif
