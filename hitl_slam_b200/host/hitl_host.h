// hitl_host.h — C entry points of the host mirror library (libhitl_host.so): file formats and
// session-level helpers that sit above the C ABI of include/hitl_gpu.h.
#pragma once
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
// .stfs.covars (loadPoseGraph, HitLSLAM_main.cpp:192-300 / SaveStfsandCovars, vector_mapping_main.cpp:1855-1928)
void* hitl_host_load_pose_graph(const char* path, uint64_t* n_poses, uint64_t* n_points);
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
#ifdef __cplusplus
}
#endif
