// pose_graph_io.cpp — `.stfs.covars` reader / writer of the host mirror.
//
// Reader mirrors loadPoseGraph (human_in_the_loop_slam/HitLSLAM_main.cpp:192-300): one text line
// per point, a new scan starts when (x, y, theta) differs from the previous line, observations
// are converted to the robot frame with R(-theta) * (p + (-t)) and — reference quirk kept —
// normals with the SAME expression (translated like points).  Values are parsed with strtof,
// i.e. the same correctly-rounded decimal->float conversion fscanf("%f") performs.
// Writer mirrors SaveStfsandCovars (episodic_non_markov_localization/vector_mapping_main.cpp:1855-1928):
//   "%.4f,%.4f,%.4f,%.4f,%.4f, %.4f,%.4f,%f, %f, %f, %f, %f, %f, %f, %f, %f\n" after two header lines.
// Built with -ffp-contract=off; sin/cos are hitl::sinf_rn / cosf_rn (bit-identical to glibc's).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "hitl_host.h"
#include "hitl_math.h"

namespace {
struct Graph {
  std::vector<float> poses, cov, pts, nrm;   // 3 / 9 per pose, 2 per point
  std::vector<uint32_t> off;
  std::string map_name;
  double timestamp = 0;
};

void flush_scan(Graph* g, std::vector<float>* pc, std::vector<float>* nc) {
  const size_t n = g->poses.size() / 3 - 1;
  const float th = -g->poses[3 * n + 2];
  const float s = hitl::sinf_rn(th), c = hitl::cosf_rn(th);
  const float lx = -g->poses[3 * n], ly = -g->poses[3 * n + 1];
  for (size_t i = 0; i < pc->size() / 2; ++i) {
    float ox, oy;
    hitl::rot_apply(c, s, (*pc)[2 * i] + lx, (*pc)[2 * i + 1] + ly, &ox, &oy);
    g->pts.push_back(ox); g->pts.push_back(oy);
    hitl::rot_apply(c, s, (*nc)[2 * i] + lx, (*nc)[2 * i + 1] + ly, &ox, &oy);
    g->nrm.push_back(ox); g->nrm.push_back(oy);
  }
  g->off.push_back((uint32_t)(g->pts.size() / 2));
  pc->clear(); nc->clear();
}

// Parses the 16 comma-separated floats of one line; returns the number parsed.
int parse_line(const char* s, float* v) {
  int n = 0;
  while (n < 16) {
    char* end;
    v[n] = strtof(s, &end);
    if (end == s) break;
    ++n;
    s = end;
    while (*s == ' ' || *s == '\t') ++s;
    if (*s == ',') ++s; else break;
  }
  return n;
}
}  // namespace

extern "C" {

void* hitl_host_load_pose_graph(const char* path, uint64_t* n_poses, uint64_t* n_points) {
  FILE* f = fopen(path, "r");
  if (!f) return NULL;
  Graph* g = new Graph();
  char* line = NULL; size_t cap = 0;
  bool ok = getline(&line, &cap, f) > 0;
  if (ok) { g->map_name = line; while (!g->map_name.empty() && (g->map_name.back() == '\n' || g->map_name.back() == '\r')) g->map_name.pop_back(); }
  ok = ok && getline(&line, &cap, f) > 0;
  if (ok) g->timestamp = strtod(line, NULL);
  if (!ok) { free(line); fclose(f); delete g; return NULL; }
  std::vector<float> pc, nc;
  float v[16];
  g->off.push_back(0);
  while (getline(&line, &cap, f) > 0) {
    if (parse_line(line, v) != 16) break;   // the reference's fscanf loop stops at the first malformed line
    const size_t n = g->poses.size() / 3;
    const bool first = n == 0;
    const bool changed = !first && (v[0] != g->poses[3 * n - 3] || v[1] != g->poses[3 * n - 2] || v[2] != g->poses[3 * n - 1]);
    if (changed) flush_scan(g, &pc, &nc);
    if (first || changed) {
      g->poses.insert(g->poses.end(), v, v + 3);
      g->cov.insert(g->cov.end(), v + 7, v + 16);
    }
    pc.push_back(v[3]); pc.push_back(v[4]); nc.push_back(v[5]); nc.push_back(v[6]);
  }
  if (!pc.empty()) flush_scan(g, &pc, &nc);
  free(line); fclose(f);
  *n_poses = g->poses.size() / 3; *n_points = g->pts.size() / 2;
  return g;
}

void hitl_host_pose_graph_get(void* h, float* poses, float* cov, uint32_t* off, float* pts, float* nrm) {
  Graph* g = static_cast<Graph*>(h);
  if (poses) memcpy(poses, g->poses.data(), 4 * g->poses.size());
  if (cov) memcpy(cov, g->cov.data(), 4 * g->cov.size());
  if (off) memcpy(off, g->off.data(), 4 * g->off.size());
  if (pts) memcpy(pts, g->pts.data(), 4 * g->pts.size());
  if (nrm) memcpy(nrm, g->nrm.data(), 4 * g->nrm.size());
}
void hitl_host_pose_graph_free(void* h) { delete static_cast<Graph*>(h); }

// One line per point; obs/normals are WORLD frame as the format requires (README.md:119-137).
int hitl_host_save_stfs_covars(const char* path, const char* map_name, double timestamp, uint32_t n_poses, const float* poses_xyt, const float* cov9,
                               const uint32_t* off, const float* obs_world_xy, const float* nrm_world_xy) {
  FILE* f = fopen(path, "w");
  if (!f) return 1;
  std::vector<char> buf(1 << 22);
  setvbuf(f, buf.data(), _IOFBF, buf.size());
  fprintf(f, "%s\n", map_name);
  fprintf(f, "%lf\n", timestamp);
  for (uint32_t i = 0; i < n_poses; ++i) {
    const float* p = poses_xyt + 3 * i; const float* c = cov9 + 9 * i;
    for (uint32_t k = off[i]; k < off[i + 1]; ++k)
      fprintf(f, "%.4f,%.4f,%.4f,%.4f,%.4f, %.4f,%.4f,%f, %f, %f, %f, %f, %f, %f, %f, %f\n", p[0], p[1], p[2], obs_world_xy[2 * k], obs_world_xy[2 * k + 1],
              nrm_world_xy[2 * k], nrm_world_xy[2 * k + 1], c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]);
  }
  const int rc = ferror(f) ? 2 : 0;
  fclose(f);
  return rc;
}

// Result writer (saveHitLResults, HitLSLAM_main.cpp:572-581): "%f %f %f\n" per pose.
int hitl_host_save_poses(const char* path, uint32_t n_poses, const float* poses_xyt) {
  FILE* f = fopen(path, "w");
  if (!f) return 1;
  for (uint32_t i = 0; i < n_poses; ++i) fprintf(f, "%f %f %f\n", poses_xyt[3 * i], poses_xyt[3 * i + 1], poses_xyt[3 * i + 2]);
  fclose(f);
  return 0;
}

}  // extern "C"
