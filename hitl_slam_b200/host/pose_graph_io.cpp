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
#include <thread>
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

// Decimal -> float, correctly rounded like strtof / fscanf("%f").  Fast path (Clinger): a plain decimal
// with at most 7 significant digits' worth of mantissa (< 2^24) and at most 10 fractional digits is the
// quotient of two exactly representable floats, and one IEEE division rounds it correctly.  Everything
// else (exponents, long mantissas, inf/nan, hex) goes to strtof.
inline float parse_float(const char* s, const char** end) {
  static const float kPow10[11] = {1e0f, 1e1f, 1e2f, 1e3f, 1e4f, 1e5f, 1e6f, 1e7f, 1e8f, 1e9f, 1e10f};
  const char* p = s;
  while (*p == ' ' || *p == '\t') ++p;
  // a number must start here: strtof would otherwise skip line breaks and read the NEXT line's first field
  if (!((*p >= '0' && *p <= '9') || *p == '-' || *p == '+' || *p == '.' || *p == 'i' || *p == 'I' || *p == 'n' || *p == 'N')) { *end = s; return 0.f; }
  const char* q = p;
  bool neg = false;
  if (*q == '-') { neg = true; ++q; } else if (*q == '+') ++q;
  uint32_t m = 0; int digits = 0, frac = 0; bool ok = true;
  while (*q >= '0' && *q <= '9') { m = m * 10 + (uint32_t)(*q - '0'); ++q; if (++digits > 8) ok = false; }
  if (*q == '.') {
    ++q;
    while (*q >= '0' && *q <= '9') { m = m * 10 + (uint32_t)(*q - '0'); ++q; ++frac; if (++digits > 8) ok = false; }
  }
  if (digits == 0 || !ok || frac > 10 || m >= (1u << 24) || *q == 'e' || *q == 'E' || *q == 'x' || *q == 'X' || *q == 'n' || *q == 'i') {
    char* e;
    const float v = strtof(s, &e);
    *end = e;
    return v;
  }
  *end = q;
  const float v = (float)m / kPow10[frac];
  return neg ? -v : v;
}

// Parses the 16 comma-separated floats of one line; returns the number parsed.
int parse_line(const char* s, float* v) {
  int n = 0;
  while (n < 16) {
    const char* end;
    v[n] = parse_float(s, &end);
    if (end == s) break;
    ++n;
    s = end;
    while (*s == ' ' || *s == '\t') ++s;
    if (*s == ',') ++s; else break;
  }
  return n;
}

// One worker's share of the file: the scans it saw (pose, covariance, first point), world-frame points and normals.
struct Piece {
  std::vector<float> poses, cov, pw, nw;   // 3 / 9 per scan, 2 per point (world frame, as stored in the file)
  std::vector<uint64_t> first;             // first point of each scan inside pw / nw
  bool stopped = false;                    // hit a malformed line: everything after it is ignored
};

void parse_piece(const char* b, const char* e, Piece* out) {
  float v[16];
  while (b < e) {
    const char* nl = (const char*)memchr(b, '\n', (size_t)(e - b));
    const char* le = nl ? nl : e;
    if (parse_line(b, v) != 16) { out->stopped = true; return; }   // the reference's fscanf loop stops at the first malformed line
    const size_t n = out->poses.size() / 3;
    if (n == 0 || v[0] != out->poses[3 * n - 3] || v[1] != out->poses[3 * n - 2] || v[2] != out->poses[3 * n - 1]) {
      out->poses.insert(out->poses.end(), v, v + 3);
      out->cov.insert(out->cov.end(), v + 7, v + 16);
      out->first.push_back(out->pw.size() / 2);
    }
    out->pw.push_back(v[3]); out->pw.push_back(v[4]); out->nw.push_back(v[5]); out->nw.push_back(v[6]);
    b = le + 1;
  }
}

unsigned io_threads() {
  const char* env = getenv("HITL_IO_THREADS");
  unsigned t = env ? (unsigned)atoi(env) : std::thread::hardware_concurrency();
  return t < 1 ? 1 : (t > 64 ? 64 : t);
}
}  // namespace

extern "C" {

// Whole file in memory, split at line boundaries into one piece per thread, parsed in parallel; pieces are stitched
// in order (a scan that straddles a boundary is merged: same pose on both sides), then every scan is converted to
// the robot frame in parallel.  Result is identical to the sequential getline loop.
void* hitl_host_load_pose_graph(const char* path, uint64_t* n_poses, uint64_t* n_points) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  std::string data;
  {
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); return NULL; }
    data.resize((size_t)sz);
    if (sz && fread(&data[0], 1, (size_t)sz, f) != (size_t)sz) { fclose(f); return NULL; }
    fclose(f);
  }
  const char* b = data.c_str();
  const char* e = b + data.size();
  const char* l1 = (const char*)memchr(b, '\n', (size_t)(e - b));
  if (!l1) return NULL;
  const char* l2 = (const char*)memchr(l1 + 1, '\n', (size_t)(e - l1 - 1));
  if (!l2) { if (l1 + 1 >= e) return NULL; l2 = e - 1; }
  Graph* g = new Graph();
  g->map_name.assign(b, l1);
  while (!g->map_name.empty() && g->map_name.back() == '\r') g->map_name.pop_back();
  g->timestamp = strtod(l1 + 1, NULL);
  const char* body = l2 + 1 <= e ? l2 + 1 : e;
  const unsigned nt = (size_t)(e - body) < (1u << 20) ? 1u : io_threads();
  std::vector<const char*> cut(nt + 1);
  cut[0] = body; cut[nt] = e;
  for (unsigned t = 1; t < nt; ++t) {
    const char* c = body + (size_t)(e - body) * t / nt;
    if (c < cut[t - 1]) c = cut[t - 1];
    const char* nl = c < e ? (const char*)memchr(c, '\n', (size_t)(e - c)) : NULL;
    cut[t] = nl ? nl + 1 : e;
  }
  std::vector<Piece> pieces(nt);
  {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(parse_piece, cut[t], cut[t + 1], &pieces[t]);
    parse_piece(cut[0], cut[1], &pieces[0]);
    for (auto& x : th) x.join();
  }
  // stitch: scan list (pose, cov, point range in the concatenated world arrays)
  std::vector<uint64_t> scan_begin;   // first point of each scan in the global numbering
  std::vector<uint64_t> base(nt + 1, 0);
  for (unsigned t = 0; t < nt; ++t) {
    const Piece& P = pieces[t];
    base[t + 1] = base[t] + P.pw.size() / 2;
    for (size_t k = 0; k < P.first.size(); ++k) {
      const size_t n = g->poses.size() / 3;
      const bool same = k == 0 && n > 0 && P.poses[0] == g->poses[3 * n - 3] && P.poses[1] == g->poses[3 * n - 2] && P.poses[2] == g->poses[3 * n - 1];
      if (same) continue;            // continuation of the scan the previous piece ended with
      g->poses.insert(g->poses.end(), P.poses.begin() + 3 * k, P.poses.begin() + 3 * k + 3);
      g->cov.insert(g->cov.end(), P.cov.begin() + 9 * k, P.cov.begin() + 9 * k + 9);
      scan_begin.push_back(base[t] + P.first[k]);
    }
    if (P.stopped) { for (unsigned u = t + 1; u <= nt; ++u) base[u] = base[t + 1]; pieces.resize(t + 1); break; }
  }
  const unsigned np = (unsigned)pieces.size();
  const uint64_t total = base[np];
  const size_t ns = scan_begin.size();
  g->pts.resize(2 * total); g->nrm.resize(2 * total);
  g->off.resize(ns + 1);
  for (size_t i = 0; i < ns; ++i) g->off[i] = (uint32_t)scan_begin[i];
  g->off[ns] = (uint32_t)total;
  // robot-frame conversion, one contiguous range of scans per thread
  auto convert = [&](size_t s_lo, size_t s_hi) {
    unsigned t = 0;
    for (size_t i = s_lo; i < s_hi; ++i) {
      const float th = -g->poses[3 * i + 2];
      const float sn = hitl::sinf_rn(th), cs = hitl::cosf_rn(th);
      const float lx = -g->poses[3 * i], ly = -g->poses[3 * i + 1];
      for (uint64_t k = g->off[i]; k < g->off[i + 1]; ++k) {
        while (k >= base[t + 1]) ++t;
        const Piece& P = pieces[t];
        const size_t q = (size_t)(k - base[t]);
        float ox, oy;
        hitl::rot_apply(cs, sn, P.pw[2 * q] + lx, P.pw[2 * q + 1] + ly, &ox, &oy);
        g->pts[2 * k] = ox; g->pts[2 * k + 1] = oy;
        hitl::rot_apply(cs, sn, P.nw[2 * q] + lx, P.nw[2 * q + 1] + ly, &ox, &oy);
        g->nrm[2 * k] = ox; g->nrm[2 * k + 1] = oy;
      }
    }
  };
  {
    const unsigned ct = ns < 64 ? 1u : nt;
    std::vector<std::thread> th;
    for (unsigned t = 1; t < ct; ++t) th.emplace_back(convert, ns * t / ct, ns * (t + 1) / ct);
    convert(0, ns / ct);
    for (auto& x : th) x.join();
  }
  *n_poses = ns; *n_points = total;
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

// ---- binary cache of a parsed pose graph ------------------------------------------------------------
// Parsing is the dominant wall-clock term outside the kernels at c3 / c5 sizes (one text line per point: 21.6 M - 216 M lines).
// The cache stores the PARSED graph (robot-frame clouds, poses, covariances, offsets: exactly what hitl_host_load_pose_graph
// returns) next to a fingerprint of the text file (size, mtime in ns, FNV-1a of the first and last 64 KB); a cache whose
// fingerprint does not match the text file is ignored and rewritten.  Layout: 64-byte header, then poses | cov | off | pts | nrm.
namespace {
struct CacheHeader {
  char magic[8];                  // "HITLPG01"
  uint64_t text_size, text_mtime_ns, text_hash, n_poses, n_points;
  double timestamp;
  uint64_t reserved;
};
static_assert(sizeof(CacheHeader) == 64, "cache header layout");
}  // namespace
}  // extern "C" (helpers below are C++)
#include <sys/stat.h>
namespace {
bool text_fingerprint(const char* path, uint64_t* size, uint64_t* mtime_ns, uint64_t* hash) {
  struct stat st;
  if (stat(path, &st) != 0) return false;
  *size = (uint64_t)st.st_size;
  *mtime_ns = (uint64_t)st.st_mtim.tv_sec * 1000000000ull + (uint64_t)st.st_mtim.tv_nsec;
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  uint64_t h = 1469598103934665603ull;
  std::vector<unsigned char> buf(65536);
  auto mix = [&](size_t n) { for (size_t i = 0; i < n; ++i) { h ^= buf[i]; h *= 1099511628211ull; } };
  size_t n = fread(buf.data(), 1, buf.size(), f);
  mix(n);
  if (*size > 2 * buf.size()) { fseek(f, -(long)buf.size(), SEEK_END); n = fread(buf.data(), 1, buf.size(), f); mix(n); }
  fclose(f);
  *hash = h;
  return true;
}
bool read_all(FILE* f, void* dst, size_t bytes) { return bytes == 0 || fread(dst, 1, bytes, f) == bytes; }
bool write_all(FILE* f, const void* src, size_t bytes) { return bytes == 0 || fwrite(src, 1, bytes, f) == bytes; }
}  // namespace
extern "C" {

// Loads `path` through its cache `cache_path` (NULL: path + ".hitlcache").  *from_cache (may be NULL) tells which way it went.
// A missing / stale / truncated cache falls back to the text parser and is rewritten (best effort: a read-only directory only
// costs the cache).  Returns the same handle type as hitl_host_load_pose_graph.
void* hitl_host_load_pose_graph_cached(const char* path, const char* cache_path, uint64_t* n_poses, uint64_t* n_points, int* from_cache) {
  if (from_cache) *from_cache = 0;
  if (!path || !n_poses || !n_points) return NULL;
  const std::string cp = cache_path ? std::string(cache_path) : std::string(path) + ".hitlcache";
  uint64_t size = 0, mtime = 0, hash = 0;
  if (!text_fingerprint(path, &size, &mtime, &hash)) return NULL;
  if (FILE* f = fopen(cp.c_str(), "rb")) {
    CacheHeader h;
    Graph* g = new Graph();
    bool ok = read_all(f, &h, sizeof(h)) && memcmp(h.magic, "HITLPG01", 8) == 0 && h.text_size == size && h.text_mtime_ns == mtime && h.text_hash == hash &&
              h.n_poses < (1ull << 32) && h.n_points < (1ull << 32);
    if (ok) {
      g->poses.resize(3 * h.n_poses); g->cov.resize(9 * h.n_poses); g->off.resize(h.n_poses + 1); g->pts.resize(2 * h.n_points); g->nrm.resize(2 * h.n_points);
      g->timestamp = h.timestamp;
      ok = read_all(f, g->poses.data(), 4 * g->poses.size()) && read_all(f, g->cov.data(), 4 * g->cov.size()) && read_all(f, g->off.data(), 4 * g->off.size()) &&
           read_all(f, g->pts.data(), 4 * g->pts.size()) && read_all(f, g->nrm.data(), 4 * g->nrm.size());
      // structural check: offsets start at 0, never decrease and end at n_points
      ok = ok && g->off[0] == 0 && g->off[h.n_poses] == h.n_points;
      for (uint64_t i = 0; ok && i < h.n_poses; ++i) ok = g->off[i] <= g->off[i + 1];
    }
    fclose(f);
    if (ok) { *n_poses = h.n_poses; *n_points = h.n_points; if (from_cache) *from_cache = 1; return g; }
    delete g;
  }
  Graph* g = static_cast<Graph*>(hitl_host_load_pose_graph(path, n_poses, n_points));
  if (!g) return NULL;
  const std::string tmp = cp + ".tmp";
  if (FILE* f = fopen(tmp.c_str(), "wb")) {
    CacheHeader h; memset(&h, 0, sizeof(h));
    memcpy(h.magic, "HITLPG01", 8);
    h.text_size = size; h.text_mtime_ns = mtime; h.text_hash = hash; h.n_poses = *n_poses; h.n_points = *n_points; h.timestamp = g->timestamp;
    const bool ok = write_all(f, &h, sizeof(h)) && write_all(f, g->poses.data(), 4 * g->poses.size()) && write_all(f, g->cov.data(), 4 * g->cov.size()) &&
                    write_all(f, g->off.data(), 4 * g->off.size()) && write_all(f, g->pts.data(), 4 * g->pts.size()) && write_all(f, g->nrm.data(), 4 * g->nrm.size());
    const bool closed = fclose(f) == 0;
    if (ok && closed) rename(tmp.c_str(), cp.c_str()); else remove(tmp.c_str());
  }
  return g;
}

// One line per point; obs/normals are WORLD frame as the format requires (README.md:119-137).
int hitl_host_save_stfs_covars(const char* path, const char* map_name, double timestamp, uint32_t n_poses, const float* poses_xyt, const float* cov9,
                               const uint32_t* off, const float* obs_world_xy, const float* nrm_world_xy) {
  FILE* f = fopen(path, "wb");
  if (!f) return 1;
  fprintf(f, "%s\n", map_name);
  fprintf(f, "%lf\n", timestamp);
  // Scans are formatted in parallel (contiguous pose ranges balanced by point count), then written in order.  The
  // pose prefix and covariance suffix of a line repeat for every point of a scan and are formatted once per scan.
  const unsigned nt = off[n_poses] < (1u << 16) ? 1u : io_threads();
  std::vector<uint32_t> cut(nt + 1, n_poses);
  cut[0] = 0;
  for (unsigned t = 1, i = 0; t < nt; ++t) {
    const uint64_t want = (uint64_t)off[n_poses] * t / nt;
    while (i < n_poses && off[i] < want) ++i;
    cut[t] = i;
  }
  std::vector<std::string> out(nt);
  auto format = [&](unsigned t) {
    std::string& o = out[t];
    o.reserve((size_t)(off[cut[t + 1]] - off[cut[t]]) * 150 + 64);
    char head[128], tail[256], mid[160];
    for (uint32_t i = cut[t]; i < cut[t + 1]; ++i) {
      const float* p = poses_xyt + 3 * i; const float* c = cov9 + 9 * i;
      const int nh = snprintf(head, sizeof(head), "%.4f,%.4f,%.4f,", p[0], p[1], p[2]);
      const int ntl = snprintf(tail, sizeof(tail), "%f, %f, %f, %f, %f, %f, %f, %f, %f\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8]);
      for (uint32_t k = off[i]; k < off[i + 1]; ++k) {
        const int nm = snprintf(mid, sizeof(mid), "%.4f,%.4f, %.4f,%.4f,", obs_world_xy[2 * k], obs_world_xy[2 * k + 1], nrm_world_xy[2 * k], nrm_world_xy[2 * k + 1]);
        o.append(head, (size_t)nh); o.append(mid, (size_t)nm); o.append(tail, (size_t)ntl);
      }
    }
  };
  {
    std::vector<std::thread> th;
    for (unsigned t = 1; t < nt; ++t) th.emplace_back(format, t);
    format(0);
    for (auto& x : th) x.join();
  }
  for (unsigned t = 0; t < nt; ++t)
    if (!out[t].empty() && fwrite(out[t].data(), 1, out[t].size(), f) != out[t].size()) { fclose(f); return 2; }
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
