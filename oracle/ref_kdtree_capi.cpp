// ref_kdtree_capi.cpp — C entry points over the REFERENCE's own KDTree<float,2>
// (perception_tools/kdtree.{h,cpp}, compiled from /root/reference where it lies, against
// oracle/shim/).  Output: oracle/_ref/libkdtree_ref.so.  TEST INFRASTRUCTURE ONLY.
#define private public   // read-only access to the node fields for flattening
#include "kdtree.h"
#undef private
#include <float.h>
#include <stdint.h>
#include <vector>

typedef KDTree<float, 2> Tree;
typedef KDNodeValue<float, 2> Val;

static void flatten(Tree* t, float* pn, int32_t* idx, int32_t* dim, size_t* o) {
  pn[4 * *o] = t->value_.point(0); pn[4 * *o + 1] = t->value_.point(1);
  pn[4 * *o + 2] = t->value_.normal(0); pn[4 * *o + 3] = t->value_.normal(1);
  idx[*o] = t->value_.index; dim[*o] = t->splitting_dimension_;
  ++*o;
  if (t->left_tree_) flatten(t->left_tree_, pn, idx, dim, o);
  if (t->right_tree_) flatten(t->right_tree_, pn, idx, dim, o);
}

extern "C" {
void* ref_kd_create(uint32_t n, const float* pts, const float* nrm) {
  std::vector<Val> v(n);
  for (uint32_t i = 0; i < n; ++i) {
    v[i].index = i;
    v[i].point = Eigen::Vector2f(pts[2 * i], pts[2 * i + 1]);
    v[i].normal = Eigen::Vector2f(nrm[2 * i], nrm[2 * i + 1]);
  }
  return new Tree(v);
}
void ref_kd_destroy(void* t) { delete static_cast<Tree*>(t); }
void ref_kd_flatten(void* t, float* pn, int32_t* idx, int32_t* dim) { size_t o = 0; flatten(static_cast<Tree*>(t), pn, idx, dim, &o); }
void ref_kd_query(void* t, uint32_t nq, const float* q, float thr, int mode, float* dist, int32_t* index) {
  Tree* T = static_cast<Tree*>(t);
  for (uint32_t i = 0; i < nq; ++i) {
    Val nb; nb.index = -1;
    const Eigen::Vector2f p(q[2 * i], q[2 * i + 1]);
    dist[i] = mode == 0 ? T->FindNearestPointNormal(p, thr, &nb) : T->FindNearestPoint(p, thr, &nb);
    index[i] = nb.index;
  }
}
uint32_t ref_kd_radius(void* t, float qx, float qy, float thr, int32_t* index, uint32_t cap) {
  std::vector<Val> out;
  static_cast<Tree*>(t)->FindNeighborPoints(Eigen::Vector2f(qx, qy), thr, &out);
  for (size_t i = 0; i < out.size() && i < cap; ++i) index[i] = out[i].index;
  return out.size();
}
}
