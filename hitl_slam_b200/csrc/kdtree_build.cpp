// kdtree_build.cpp — host-side KD-tree construction, emitted directly in the flat preorder
// layout the kernels traverse (include/hitl_gpu.h: hitl_kdnode).
//
// Shape contract (reference: perception_tools/kdtree.cpp:37-69 GetSplittingPlane, :106-139
// BuildKDTree): split dimension = larger sequential-float sum of squared deviations from the
// sequential-float mean (dimension 0 on ties), std::sort on that coordinate with a strict
// '<' comparator, node = element n/2, children built from the two sorted halves in their
// sorted order.  std::sort is not stable, so ties in a coordinate are resolved by
// libstdc++'s introsort exactly as in the reference build; sorting a sub-range in place
// performs the same comparison sequence as sorting a fresh copy of it.
//
// Built with -ffp-contract=off: the mean/deviation sums must not be fused.
#include <algorithm>
#include <vector>
#include "hitl_internal.h"

namespace hitl {
namespace {
struct Item { float px, py, nx, ny; int32_t index; };

inline int split_dim(const Item* v, uint32_t n) {
  float mx = 0.0f, my = 0.0f;
  for (uint32_t i = 0; i < n; ++i) { mx = mx + v[i].px; my = my + v[i].py; }
  mx = mx / static_cast<float>(n);
  my = my / static_cast<float>(n);
  float dx = 0.0f, dy = 0.0f;
  for (uint32_t i = 0; i < n; ++i) {
    dx = dx + (v[i].px - mx) * (v[i].px - mx);
    dy = dy + (v[i].py - my) * (v[i].py - my);
  }
  int dim = 0;
  float best = 0.0f;
  if (dx > best) { dim = 0; best = dx; }
  if (dy > best) { dim = 1; best = dy; }
  return dim;
}

void build_range(Item* v, uint32_t n, hitl_kdnode* out, uint32_t pos) {
  const int dim = split_dim(v, n);
  if (dim == 0) std::sort(v, v + n, [](const Item& a, const Item& b) { return a.px < b.px; });
  else std::sort(v, v + n, [](const Item& a, const Item& b) { return a.py < b.py; });
  const uint32_t mid = n / 2;
  hitl_kdnode& o = out[pos];
  o.px = v[mid].px; o.py = v[mid].py; o.nx = v[mid].nx; o.ny = v[mid].ny; o.index = v[mid].index; o.dim = dim;
  if (mid > 0) build_range(v, mid, out, pos + 1);
  if (mid + 1 < n) build_range(v + mid + 1, n - 1 - mid, out, pos + 1 + mid);
}
}  // namespace

void build_flat_kdtree(const float* pts_xy, const float* nrm_xy, uint32_t n, hitl_kdnode* out) {
  if (n == 0) return;
  std::vector<Item> v(n);
  for (uint32_t i = 0; i < n; ++i) {
    v[i].px = pts_xy[2 * i]; v[i].py = pts_xy[2 * i + 1];
    v[i].nx = nrm_xy[2 * i]; v[i].ny = nrm_xy[2 * i + 1];
    v[i].index = (int32_t)i;
  }
  build_range(v.data(), n, out, 0);
}
}  // namespace hitl
