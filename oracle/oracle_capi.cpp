// oracle_capi.cpp — flat C entry points over hitl_oracle.hpp for ctypes (tests/, bench.py's
// cpu_baseline and --impl reference legs).  TEST INFRASTRUCTURE ONLY — never linked into or
// loaded by the product library.
#include "hitl_oracle.hpp"
#include <string.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

using namespace orc;

namespace {
struct Handle {
  ScanSet S;
  std::vector<GlobCorrespondence> stf;
  std::vector<PointCorrespondence> vo;
  uint64_t n_queries = 0;
};
void fill(std::vector<std::vector<V2> >* dst, uint32_t n, const uint32_t* off, const float* xy) {
  dst->resize(n);
  for (uint32_t i = 0; i < n; ++i) {
    (*dst)[i].resize(off[i + 1] - off[i]);
    for (uint32_t k = off[i]; k < off[i + 1]; ++k) (*dst)[i][k - off[i]] = V2(xy[2 * k], xy[2 * k + 1]);
  }
}
}  // namespace

extern "C" {

int orc_num_threads() {
#if defined(_OPENMP)
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void* orc_create(uint32_t n, const uint32_t* off, const float* pts, const float* nrm, int build_trees) {
  Handle* h = new Handle();
  fill(&h->S.points, n, off, pts);
  fill(&h->S.normals, n, off, nrm);
  h->S.trees.assign(n, NULL);
  if (build_trees) h->S.build_trees();
  return h;
}
void orc_destroy(void* p) { delete static_cast<Handle*>(p); }

// Preorder flattening: 6 words per node {px, py, nx, ny, index, dim}, scans concatenated.
void orc_flatten(void* p, float* pn /* 4 per node */, int32_t* idx, int32_t* dim) {
  Handle* h = static_cast<Handle*>(p);
  size_t o = 0;
  for (size_t i = 0; i < h->S.trees.size(); ++i) {
    if (!h->S.trees[i]) continue;
    std::vector<KDValue> v; std::vector<int> d;
    h->S.trees[i]->flatten(&v, &d);
    for (size_t k = 0; k < v.size(); ++k, ++o) {
      pn[4 * o] = v[k].point.x; pn[4 * o + 1] = v[k].point.y; pn[4 * o + 2] = v[k].normal.x; pn[4 * o + 3] = v[k].normal.y;
      idx[o] = v[k].index; dim[o] = d[k];
    }
  }
}

// mode 0: nearest_point_normal, 1: nearest_point. index = -1 when nothing was written.
void orc_query(void* p, uint32_t scan, uint32_t nq, const float* q, float thr, int mode, float* dist, int32_t* index) {
  Handle* h = static_cast<Handle*>(p);
  for (uint32_t i = 0; i < nq; ++i) {
    KDValue nb; nb.index = -1;
    float d = FLT_MAX;
    if (h->S.trees[scan]) {
      d = mode == 0 ? h->S.trees[scan]->nearest_point_normal(V2(q[2 * i], q[2 * i + 1]), thr, &nb)
                    : h->S.trees[scan]->nearest_point(V2(q[2 * i], q[2 * i + 1]), thr, &nb);
    }
    dist[i] = d; index[i] = nb.index;
  }
}
uint32_t orc_radius(void* p, uint32_t scan, float qx, float qy, float thr, int32_t* index, uint32_t cap) {
  Handle* h = static_cast<Handle*>(p);
  std::vector<KDValue> out;
  if (h->S.trees[scan]) h->S.trees[scan]->neighbor_points(V2(qx, qy), thr, &out);
  for (size_t i = 0; i < out.size() && i < cap; ++i) index[i] = out[i].index;
  return out.size();
}

void orc_relative_pose(const double* pose_array, uint32_t src, uint32_t dst, float* out6) {
  const Aff T = relative_pose_transform(pose_array, src, dst);
  out6[0] = T.L.m00; out6[1] = T.L.m01; out6[2] = T.L.m10; out6[3] = T.L.m11; out6[4] = T.t.x; out6[5] = T.t.y;
}

// counts[0] = kept pairs, counts[1] = kept matches, counts[2] = executed queries.
void orc_find_stf(void* p, const double* poses, uint64_t min_pose, uint64_t max_pose, float thr, float min_cos,
                  int cap, uint32_t skip, uint32_t min_corr, uint64_t src_lo, uint64_t src_hi, uint64_t* counts) {
  Handle* h = static_cast<Handle*>(p);
  StfOptions o; o.kPointMatchThreshold = thr; o.min_cosine_angle = min_cos; o.kMaxCorrespondencesPerPoint = cap;
  o.num_skip_readings = skip; o.kMinInterPoseCorrespondence = min_corr;
  find_stf(h->S, poses, min_pose, max_pose, o, &h->stf, &h->n_queries, src_lo, src_hi);
  uint64_t m = 0;
  for (size_t i = 0; i < h->stf.size(); ++i) m += h->stf[i].points0_indices.size();
  counts[0] = h->stf.size(); counts[1] = m; counts[2] = h->n_queries;
}
// The same with every src_stride-th source pose of [src_lo, src_hi) only (timing samples).
void orc_find_stf_strided(void* p, const double* poses, uint64_t min_pose, uint64_t max_pose, float thr, float min_cos,
                          int cap, uint32_t skip, uint32_t min_corr, uint64_t src_lo, uint64_t src_hi, uint64_t src_stride, uint64_t* counts) {
  Handle* h = static_cast<Handle*>(p);
  StfOptions o; o.kPointMatchThreshold = thr; o.min_cosine_angle = min_cos; o.kMaxCorrespondencesPerPoint = cap;
  o.num_skip_readings = skip; o.kMinInterPoseCorrespondence = min_corr;
  find_stf(h->S, poses, min_pose, max_pose, o, &h->stf, &h->n_queries, src_lo, src_hi, src_stride);
  uint64_t m = 0;
  for (size_t i = 0; i < h->stf.size(); ++i) m += h->stf[i].points0_indices.size();
  counts[0] = h->stf.size(); counts[1] = m; counts[2] = h->n_queries;
}
void orc_get_stf(void* p, uint32_t* pair_i, uint32_t* pair_j, uint64_t* pair_off, uint32_t* k, uint32_t* idx) {
  Handle* h = static_cast<Handle*>(p);
  uint64_t o = 0;
  for (size_t b = 0; b < h->stf.size(); ++b) {
    pair_i[b] = h->stf[b].pose_index0; pair_j[b] = h->stf[b].pose_index1; pair_off[b] = o;
    for (size_t m = 0; m < h->stf[b].points0_indices.size(); ++m, ++o) { k[o] = h->stf[b].points0_indices[m]; idx[o] = h->stf[b].points1_indices[m]; }
  }
  pair_off[h->stf.size()] = o;
}
uint64_t orc_find_vo(void* p, const double* poses, int min_pose, int max_pose, float thr, float min_cos) {
  Handle* h = static_cast<Handle*>(p);
  StfOptions o; o.kPointMatchThreshold = thr; o.min_cosine_angle = min_cos; o.kMaxCorrespondencesPerPoint = 0;
  o.num_skip_readings = 1; o.kMinInterPoseCorrespondence = 0;
  find_vo(h->S, poses, min_pose, max_pose, o, &h->vo);
  return h->vo.size();
}
void orc_get_vo(void* p, uint32_t* sp, uint32_t* sk, uint32_t* tk) {
  Handle* h = static_cast<Handle*>(p);
  for (size_t i = 0; i < h->vo.size(); ++i) { sp[i] = h->vo[i].source_pose; sk[i] = h->vo[i].source_point; tk[i] = h->vo[i].target_point; }
}

// ---- world transform + EM ---------------------------------------------------------------
void orc_world_transform(void* p, const float* poses_xyt, float* out_xy) {
  Handle* h = static_cast<Handle*>(p);
  std::vector<std::vector<V2> > w;
  world_transform(poses_xyt, h->S.points, &w);
  size_t o = 0;
  for (size_t i = 0; i < w.size(); ++i) for (size_t j = 0; j < w[i].size(); ++j, ++o) { out_xy[2 * o] = w[i][j].x; out_xy[2 * o + 1] = w[i][j].y; }
}
uint64_t orc_verify_input(uint32_t n, const uint32_t* off, const float* world_xy, uint32_t n_sel, const float* sel_xy, float thr, uint32_t* seen_mask) {
  std::vector<std::vector<V2> > w; fill(&w, n, off, world_xy);
  std::vector<V2> sel(n_sel);
  for (uint32_t i = 0; i < n_sel; ++i) sel[i] = V2(sel_xy[2 * i], sel_xy[2 * i + 1]);
  return verify_user_input(w, sel.data(), n_sel, seen_mask, thr);
}
uint64_t orc_em_inliers(uint32_t n, const uint32_t* off, const float* world_xy, const float seg[4], double thr,
                        uint32_t* out_pose, uint32_t* out_idx, uint64_t cap) {
  std::vector<std::vector<V2> > w; fill(&w, n, off, world_xy);
  std::vector<Inlier> in;
  em_inliers(w, V2(seg[0], seg[1]), V2(seg[2], seg[3]), thr, &in);
  for (size_t i = 0; i < in.size() && i < cap; ++i) { out_pose[i] = in[i].pose; out_idx[i] = in[i].index; }
  return in.size();
}
// Flattened observation sets: for each feature f in {0,1}: n_sets[f], then pose ids, CSR offsets and indices.
void orc_em_assign(uint32_t n, const uint32_t* off, const float* world_xy, const float segs[8], double thr, uint32_t min_obs,
                   uint32_t* n_sets, uint32_t* set_pose0, uint64_t* set_off0, uint32_t* obs0,
                   uint32_t* set_pose1, uint64_t* set_off1, uint32_t* obs1) {
  std::vector<std::vector<V2> > w; fill(&w, n, off, world_xy);
  const V2 sel[4] = {V2(segs[0], segs[1]), V2(segs[2], segs[3]), V2(segs[4], segs[5]), V2(segs[6], segs[7])};
  ObsSets a, b;
  establish_observation_sets(w, sel, thr, min_obs, &a, &b);
  auto dump = [](const ObsSets& s, uint32_t* pose, uint64_t* soff, uint32_t* obs) {
    uint64_t o = 0;
    for (size_t i = 0; i < s.size(); ++i) { pose[i] = s[i].first; soff[i] = o; for (size_t k = 0; k < s[i].second.size(); ++k) obs[o++] = s[i].second[k]; }
    soff[s.size()] = o;
  };
  n_sets[0] = a.size(); n_sets[1] = b.size();
  dump(a, set_pose0, set_off0, obs0); dump(b, set_pose1, set_off1, obs1);
}
// Full EMInput::Run: refit both strokes, assign, order. out_lists: corrected then anchor poses.
// ret[0..3] = n_corrected, n_anchor, backprop_start, backprop_end ; ret[4] = swapped ; ret[5] = EM rounds
void orc_em_run(uint32_t n, const uint32_t* off, const float* world_xy, float segs[8], int32_t* ret,
                int32_t* corrected, int32_t* anchor) {
  std::vector<std::vector<V2> > w; fill(&w, n, off, world_xy);
  V2 sel[4] = {V2(segs[0], segs[1]), V2(segs[2], segs[3]), V2(segs[4], segs[5]), V2(segs[6], segs[7])};
  const int rounds = automatic_endpoint_adjustment(w, sel);
  ObsSets a, b;
  establish_observation_sets(w, sel, 0.03, 5, &a, &b);
  OrderResult R = order_and_filter(a, b, sel);
  for (int i = 0; i < 4; ++i) { segs[2 * i] = sel[i].x; segs[2 * i + 1] = sel[i].y; }
  ret[0] = R.corrected_poses.size(); ret[1] = R.anchor_poses.size(); ret[2] = R.backprop_start; ret[3] = R.backprop_end;
  ret[4] = R.swapped; ret[5] = rounds;
  for (size_t i = 0; i < R.corrected_poses.size(); ++i) corrected[i] = R.corrected_poses[i];
  for (size_t i = 0; i < R.anchor_poses.size(); ++i) anchor[i] = R.anchor_poses[i];
}
void orc_seg_fit(const double p1[2], const double p2[2], const double* data, int size, float out[4]) {
  V2 a, b; seg_fit_em(p1, p2, data, size, &a, &b);
  out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}
float orc_distance_to_line_segment(const float s[4], float px, float py) { return distance_to_line_segment(V2(s[0], s[1]), V2(s[2], s[3]), V2(px, py)); }
double orc_dist_to_line_seg(const float s[4], float px, float py) { return dist_to_line_seg(V2(s[0], s[1]), V2(s[2], s[3]), V2(px, py)); }

// ---- residual blocks ----------------------------------------------------------------------
// STF blocks from the scan set + a CSR correspondence list (AddSTFConstraints, JointOptimization.cpp:539-559).
// r: 2 per block; J: 12 per block = [2x3 wrt pose i | 2x3 wrt pose j] row-major.
void orc_eval_stf(void* p, const double* poses, uint64_t n_pairs, const uint32_t* pair_i, const uint32_t* pair_j,
                  const uint64_t* pair_off, const uint32_t* k, const uint32_t* idx, float std_dev, float corr,
                  double* r, double* J, int parallel) {
  Handle* h = static_cast<Handle*>(p);
#if defined(_OPENMP)
#pragma omp parallel for schedule(dynamic, 16) if (parallel)
#endif
  for (uint64_t b = 0; b < n_pairs; ++b) {
    PointToPointGlob f; f.std_dev = std_dev; f.correlation_factor = corr;
    const uint32_t i = pair_i[b], j = pair_j[b];
    for (uint64_t m = pair_off[b]; m < pair_off[b + 1]; ++m) {
      f.points0.push_back(h->S.points[i][k[m]]); f.points1.push_back(h->S.points[j][idx[m]]);
      f.normals0.push_back(h->S.normals[i][k[m]]); f.normals1.push_back(h->S.normals[j][idx[m]]);
    }
    autodiff2<PointToPointGlob, 2>(f, poses + 3 * i, poses + 3 * j, r + 2 * b, J ? J + 12 * b : NULL, J ? J + 12 * b + 6 : NULL);
  }
}
// Odometry blocks (AddOdometryConstraints): constants from float poses, evaluated at double poses.
// consts: 9 floats per block; r: 3 per block; J: 18 per block = [3x3 wrt pose i-1 | 3x3 wrt pose i].
void orc_odometry_consts(const float* poses_xyt, uint32_t n, float* consts) {
  for (uint32_t i = 1; i < n; ++i) {
    const PoseConstraint pc = make_odometry_block(poses_xyt, i);
    float* c = consts + 9 * (i - 1);
    c[0] = pc.a00; c[1] = pc.a01; c[2] = pc.a10; c[3] = pc.a11; c[4] = pc.radial_std_dev; c[5] = pc.tangential_std_dev;
    c[6] = pc.angular_std_dev; c[7] = pc.radial_translation; c[8] = pc.rotation;
  }
}
void orc_eval_odometry(const float* consts, const double* poses, uint32_t n, double* r, double* J) {
  for (uint32_t i = 1; i < n; ++i) {
    const float* c = consts + 9 * (i - 1);
    PoseConstraint pc; pc.a00 = c[0]; pc.a01 = c[1]; pc.a10 = c[2]; pc.a11 = c[3]; pc.radial_std_dev = c[4];
    pc.tangential_std_dev = c[5]; pc.angular_std_dev = c[6]; pc.radial_translation = c[7]; pc.rotation = c[8];
    autodiff2<PoseConstraint, 3>(pc, poses + 3 * (i - 1), poses + 3 * i, r + 3 * (i - 1), J ? J + 18 * (i - 1) : NULL, J ? J + 18 * (i - 1) + 9 : NULL);
  }
}
// Human blocks. hc: per constraint {type, constrained, anchor} ints + {dpar, dperp, dang, pen} floats.
// blocks out: per block {type, pose} + 4 doubles. r: 3 slots per block (unused = 0); J: 9 per block.
void orc_human_blocks(const float* poses_xyt, uint32_t n, const int32_t* hc_i, const float* hc_f, int32_t* blk_i, double* blk_d) {
  for (uint32_t b = 0; b < n; ++b) {
    HumanConstraint c; c.constraint_type = hc_i[3 * b]; c.constrained_pose_id = hc_i[3 * b + 1]; c.anchor_pose_id = hc_i[3 * b + 2];
    c.delta_parallel = hc_f[4 * b]; c.delta_perpendicular = hc_f[4 * b + 1]; c.delta_angle = hc_f[4 * b + 2]; c.relative_penalty_dir = hc_f[4 * b + 3];
    const HumanBlock hb = make_human_block(poses_xyt, c);
    blk_i[2 * b] = hb.type; blk_i[2 * b + 1] = hb.pose;
    blk_d[4 * b] = hb.x_target; blk_d[4 * b + 1] = hb.y_target; blk_d[4 * b + 2] = hb.t_target; blk_d[4 * b + 3] = hb.penalty_dir;
  }
}
void orc_eval_human(uint32_t n, const int32_t* blk_i, const double* blk_d, const double* poses, double* r, double* J) {
  for (uint32_t b = 0; b < n; ++b) {
    HumanBlock hb; hb.type = blk_i[2 * b]; hb.pose = blk_i[2 * b + 1];
    hb.x_target = blk_d[4 * b]; hb.y_target = blk_d[4 * b + 1]; hb.t_target = blk_d[4 * b + 2]; hb.penalty_dir = blk_d[4 * b + 3];
    double rr[3] = {0, 0, 0}, JJ[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    autodiff1(hb, hb.num_residuals(), poses + 3 * hb.pose, rr, JJ);
    for (int i = 0; i < 3; ++i) r[3 * b + i] = rr[i];
    if (J) for (int i = 0; i < 9; ++i) J[9 * b + i] = JJ[i];
  }
}
// Point-to-line glob blocks: CSR over points; one pose per block. r: 1 per block, J: 3 per block.
void orc_eval_p2l_glob(uint32_t n_blocks, const uint32_t* blk_pose, const uint64_t* blk_off, const float* pts, const float* line_n,
                       const float* line_off, const uint8_t* valid, float std_dev, float corr, const double* poses, double* r, double* J) {
  for (uint32_t b = 0; b < n_blocks; ++b) {
    PointToLineGlob f; f.std_dev = std_dev; f.correlation_factor = corr;
    for (uint64_t m = blk_off[b]; m < blk_off[b + 1]; ++m) {
      f.points.push_back(V2(pts[2 * m], pts[2 * m + 1])); f.line_normals.push_back(V2(line_n[2 * m], line_n[2 * m + 1]));
      f.line_offsets.push_back(line_off[m]); f.valid.push_back(valid[m]);
    }
    autodiff1(f, 1, poses + 3 * blk_pose[b], r + b, J ? J + 3 * b : NULL);
  }
}
// Single point-to-line residuals: one per point.
void orc_eval_p2l(uint64_t n, const uint32_t* pose_idx, const float* pts, const float* line_n, const float* line_off,
                  const uint8_t* valid, float std_dev, float corr, const double* poses, double* r, double* J) {
  for (uint64_t m = 0; m < n; ++m) {
    PointToLine f; f.point = V2(pts[2 * m], pts[2 * m + 1]); f.line_normal = V2(line_n[2 * m], line_n[2 * m + 1]);
    f.line_offset = line_off[m]; f.valid = valid[m]; f.std_dev = std_dev; f.correlation_factor = corr;
    autodiff1(f, 1, poses + 3 * pose_idx[m], r + m, J ? J + 3 * m : NULL);
  }
}

// ---- file format ---------------------------------------------------------------------------
void* orc_load_pose_graph(const char* path, uint64_t* n_poses, uint64_t* n_points) {
  PoseGraph* g = new PoseGraph();
  if (!load_pose_graph(path, g)) { delete g; return NULL; }
  uint64_t np = 0; for (size_t i = 0; i < g->points.size(); ++i) np += g->points[i].size();
  *n_poses = g->points.size(); *n_points = np;
  return g;
}
void orc_pose_graph_get(void* p, float* poses, float* cov, uint32_t* off, float* pts, float* nrm) {
  PoseGraph* g = static_cast<PoseGraph*>(p);
  memcpy(poses, g->poses.data(), g->poses.size() * 4);
  memcpy(cov, g->covariances.data(), g->covariances.size() * 4);
  uint32_t o = 0;
  for (size_t i = 0; i < g->points.size(); ++i) {
    off[i] = o;
    for (size_t k = 0; k < g->points[i].size(); ++k, ++o) { pts[2 * o] = g->points[i][k].x; pts[2 * o + 1] = g->points[i][k].y; nrm[2 * o] = g->normals[i][k].x; nrm[2 * o + 1] = g->normals[i][k].y; }
  }
  off[g->points.size()] = o;
}
void orc_pose_graph_free(void* p) { delete static_cast<PoseGraph*>(p); }

// ---- explicit correction + back-propagation (f3) -------------------------------------------------
// poses: x, y, theta floats (in/out).  Returns 1 when a contiguous group was applied (C written), 0 otherwise.
int orc_app_exp_corrections(int type, const float* sel8, float* poses_xyt, uint32_t n_poses, const int32_t* corrected, uint32_t n_corrected, float* C3) {
  std::vector<Pose2Df> poses(n_poses);
  for (uint32_t i = 0; i < n_poses; ++i) { poses[i].translation = V2(poses_xyt[3 * i], poses_xyt[3 * i + 1]); poses[i].angle = poses_xyt[3 * i + 2]; }
  V2 sel[4];
  for (int k = 0; k < 4; ++k) sel[k] = V2(sel8[2 * k], sel8[2 * k + 1]);
  std::vector<int> corr(corrected, corrected + n_corrected);
  bool applied = false;
  app_exp_corrections(type, sel, &poses, corr, C3, &applied);
  for (uint32_t i = 0; i < n_poses; ++i) { poses_xyt[3 * i] = poses[i].translation.x; poses_xyt[3 * i + 1] = poses[i].translation.y; poses_xyt[3 * i + 2] = poses[i].angle; }
  return applied ? 1 : 0;
}
void orc_backprop(float* poses_xyt, float* cov9, uint32_t n_poses, int32_t lo, int32_t hi, const float* C3) {
  std::vector<Pose2Df> poses(n_poses);
  for (uint32_t i = 0; i < n_poses; ++i) { poses[i].translation = V2(poses_xyt[3 * i], poses_xyt[3 * i + 1]); poses[i].angle = poses_xyt[3 * i + 2]; }
  std::vector<float> cov(cov9, cov9 + 9 * (size_t)n_poses);
  backprop(&poses, &cov, lo, hi, C3);
  for (uint32_t i = 0; i < n_poses; ++i) { poses_xyt[3 * i] = poses[i].translation.x; poses_xyt[3 * i + 1] = poses[i].translation.y; poses_xyt[3 * i + 2] = poses[i].angle; }
  memcpy(cov9, cov.data(), 4 * cov.size());
}

float orc_sinf(float x) { return sinf(x); }
float orc_cosf(float x) { return cosf(x); }
double orc_angle_mod(double a) { return angle_mod_d(a); }

}  // extern "C"
