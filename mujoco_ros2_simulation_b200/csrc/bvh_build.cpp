// bvh_build.cpp -- binned-SAH BVH2 builder (host, one-off).  See bvh_build.h.
#include "bvh_build.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <numeric>

namespace b2 {
namespace {

constexpr int kBins = 16;

struct Tmp {
  Aabb box;
  int left = -1, right = -1;  // indices into tmp (inner) or -1
  int first = 0, count = 0;   // range in `order`
};

inline void box_reset(Aabb& b) {
  for (int k = 0; k < 3; k++) {
    b.lo[k] = FLT_MAX;
    b.hi[k] = -FLT_MAX;
  }
}
inline void box_grow(Aabb& b, const Aabb& o) {
  for (int k = 0; k < 3; k++) {
    b.lo[k] = std::min(b.lo[k], o.lo[k]);
    b.hi[k] = std::max(b.hi[k], o.hi[k]);
  }
}
inline float box_area(const Aabb& b) {
  const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
  if (dx < 0 || dy < 0 || dz < 0) return 0.f;
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

struct Builder {
  const std::vector<Aabb>& prims;
  std::vector<float> cen;  // [n*3]
  std::vector<int32_t> order;
  std::vector<Tmp> tmp;
  int max_leaf;
  float cost_isect;
  int max_depth = 0;

  explicit Builder(const std::vector<Aabb>& p, int ml, float ci)
      : prims(p), max_leaf(ml), cost_isect(ci) {
    const int n = (int)p.size();
    cen.resize((size_t)n * 3);
    order.resize(n);
    std::iota(order.begin(), order.end(), 0);
    for (int i = 0; i < n; i++)
      for (int k = 0; k < 3; k++) cen[3 * i + k] = 0.5f * (p[i].lo[k] + p[i].hi[k]);
  }

  int build(int first, int count, int depth) {
    max_depth = std::max(max_depth, depth);
    const int id = (int)tmp.size();
    tmp.emplace_back();
    Aabb box, cbox;
    box_reset(box);
    box_reset(cbox);
    for (int i = first; i < first + count; i++) {
      box_grow(box, prims[order[i]]);
      for (int k = 0; k < 3; k++) {
        const float c = cen[3 * order[i] + k];
        cbox.lo[k] = std::min(cbox.lo[k], c);
        cbox.hi[k] = std::max(cbox.hi[k], c);
      }
    }
    tmp[id].box = box;
    tmp[id].first = first;
    tmp[id].count = count;
    if (count == 1) return id;

    // binned SAH over the three axes
    float best_cost = FLT_MAX;
    int best_axis = -1, best_split = -1;
    for (int axis = 0; axis < 3; axis++) {
      const float lo = cbox.lo[axis], ext = cbox.hi[axis] - cbox.lo[axis];
      if (!(ext > 0.f)) continue;
      Aabb bb[kBins];
      int bc[kBins] = {0};
      for (auto& b : bb) box_reset(b);
      const float scale = kBins / ext;
      for (int i = first; i < first + count; i++) {
        int b = (int)((cen[3 * order[i] + axis] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        bc[b]++;
        box_grow(bb[b], prims[order[i]]);
      }
      float right_area[kBins];
      int right_cnt[kBins];
      Aabb acc;
      box_reset(acc);
      int cnt = 0;
      for (int b = kBins - 1; b > 0; b--) {
        box_grow(acc, bb[b]);
        cnt += bc[b];
        right_area[b] = box_area(acc);
        right_cnt[b] = cnt;
      }
      box_reset(acc);
      cnt = 0;
      for (int b = 0; b < kBins - 1; b++) {
        box_grow(acc, bb[b]);
        cnt += bc[b];
        if (cnt == 0 || right_cnt[b + 1] == 0) continue;
        const float cost = box_area(acc) * cnt + right_area[b + 1] * right_cnt[b + 1];
        if (cost < best_cost) {
          best_cost = cost;
          best_axis = axis;
          best_split = b;
        }
      }
    }
    const float parent_area = std::max(box_area(box), 1e-30f);
    if (count <= max_leaf) {
      const float leaf_cost = cost_isect * count;
      const float split_cost = 1.0f + cost_isect * best_cost / parent_area;
      if (best_axis < 0 || leaf_cost <= split_cost) return id;  // stays a leaf
    }
    int mid;
    if (best_axis >= 0) {
      const float lo = cbox.lo[best_axis], ext = cbox.hi[best_axis] - cbox.lo[best_axis];
      const float scale = kBins / ext;
      auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](int p) {
        int b = (int)((cen[3 * p + best_axis] - lo) * scale);
        b = std::min(std::max(b, 0), kBins - 1);
        return b <= best_split;
      });
      mid = (int)(it - order.begin());
    } else {
      mid = first;  // all centroids coincide
    }
    if (mid == first || mid == first + count) {
      // degenerate: fall back to an index median (keeps coincident duplicates adjacent)
      mid = first + count / 2;
      std::sort(order.begin() + first, order.begin() + first + count);
    }
    const int l = build(first, mid - first, depth + 1);
    const int r = build(mid, first + count - mid, depth + 1);
    tmp[id].left = l;
    tmp[id].right = r;
    return id;
  }
};

inline void set_child(BvhNode& n, int which, const Aabb* box, int32_t link) {
  float* lo = which == 0 ? n.lo0 : n.lo1;
  float* hi = which == 0 ? n.hi0 : n.hi1;
  for (int k = 0; k < 3; k++) {
    lo[k] = box ? box->lo[k] : FLT_MAX;
    hi[k] = box ? box->hi[k] : -FLT_MAX;
  }
  (which == 0 ? n.c0 : n.c1) = link;
}

}  // namespace

BuiltBvh build_bvh(const std::vector<Aabb>& prims, int max_leaf, bool leaf_payload_geom,
                   float cost_isect) {
  BuiltBvh out;
  const int n = (int)prims.size();
  if (n == 0) return out;
  if (leaf_payload_geom) max_leaf = 1;
  max_leaf = std::min(std::max(max_leaf, 1), 4);
  Builder B(prims, max_leaf, cost_isect);
  B.tmp.reserve((size_t)2 * n);
  B.build(0, n, 1);
  out.order = B.order;
  out.max_depth = B.max_depth;

  auto leaf_link = [&](const Tmp& t) -> int32_t {
    if (leaf_payload_geom) return ~(int32_t)B.order[t.first];
    return ~(int32_t)((t.first << 2) | (t.count - 1));
  };

  // flatten inner nodes in pre-order so children follow their parent
  std::vector<int> inner_index(B.tmp.size(), -1);
  std::vector<int> stack;
  if (B.tmp[0].left < 0) {
    // single leaf: synthesise a root whose child 0 is the leaf
    BvhNode root{};
    set_child(root, 0, &B.tmp[0].box, leaf_link(B.tmp[0]));
    set_child(root, 1, nullptr, kEmptyChild);
    out.nodes.push_back(root);
  } else {
    stack.push_back(0);
    int next = 0;
    // first pass: assign indices in pre-order
    std::vector<int> pre;
    while (!stack.empty()) {
      const int t = stack.back();
      stack.pop_back();
      inner_index[t] = next++;
      pre.push_back(t);
      const Tmp& T = B.tmp[t];
      if (B.tmp[T.right].left >= 0) stack.push_back(T.right);
      if (B.tmp[T.left].left >= 0) stack.push_back(T.left);
    }
    out.nodes.resize(next);
    for (int t : pre) {
      const Tmp& T = B.tmp[t];
      BvhNode node{};
      const Tmp& L = B.tmp[T.left];
      const Tmp& R = B.tmp[T.right];
      set_child(node, 0, &L.box, L.left >= 0 ? inner_index[T.left] : leaf_link(L));
      set_child(node, 1, &R.box, R.left >= 0 ? inner_index[T.right] : leaf_link(R));
      out.nodes[inner_index[t]] = node;
    }
  }

  // refit schedule: height 0 = both children are leaves
  const int nn = (int)out.nodes.size();
  std::vector<int> height(nn, 0);
  for (int i = nn - 1; i >= 0; i--) {
    int h = 0;
    if (out.nodes[i].c0 >= 0) h = std::max(h, height[out.nodes[i].c0] + 1);
    if (out.nodes[i].c1 >= 0) h = std::max(h, height[out.nodes[i].c1] + 1);
    height[i] = h;
  }
  const int nlevel = nn ? *std::max_element(height.begin(), height.end()) + 1 : 0;
  out.refit_level_start.assign(nlevel + 1, 0);
  for (int i = 0; i < nn; i++) out.refit_level_start[height[i] + 1]++;
  for (int l = 0; l < nlevel; l++) out.refit_level_start[l + 1] += out.refit_level_start[l];
  out.refit_nodes.resize(nn);
  std::vector<int> cursor(out.refit_level_start.begin(), out.refit_level_start.end() - 1);
  for (int i = 0; i < nn; i++) out.refit_nodes[cursor[height[i]]++] = i;
  return out;
}

}  // namespace b2

// ---------------------------------------------------------------------------------------------
// BVH2 -> fanout-32 collapse (geom BVH).  A wide node starts as the two children of a BVH2 node and
// repeatedly replaces its largest-area inner child by that child's two children until it holds 32
// children or only geoms are left.
// ---------------------------------------------------------------------------------------------
namespace b2 {
namespace {

struct WItem {
  int32_t link;  // BVH2 link: >= 0 inner node, < 0 leaf payload (~geom)
  float area;
};

float child_area(const BvhNode& n, int which) {
  const float* lo = which == 0 ? n.lo0 : n.lo1;
  const float* hi = which == 0 ? n.hi0 : n.hi1;
  const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  if (dx < 0 || dy < 0 || dz < 0) return 0.f;
  return 2.f * (dx * dy + dy * dz + dz * dx);
}

}  // namespace

WideBvh collapse_wide(const BuiltBvh& bvh2) {
  WideBvh out;
  if (bvh2.nodes.empty()) return out;
  struct Pending {
    int32_t bvh2_node;
    int32_t wide_id;
    int depth;
  };
  // geoms under every BVH2 node (children always have a larger index than their parent)
  std::vector<int> leaves(bvh2.nodes.size(), 0);
  for (int i = (int)bvh2.nodes.size() - 1; i >= 0; i--) {
    const BvhNode& n = bvh2.nodes[(size_t)i];
    int c = 0;
    if (n.c0 != kEmptyChild) c += n.c0 >= 0 ? leaves[(size_t)n.c0] : 1;
    if (n.c1 != kEmptyChild) c += n.c1 >= 0 ? leaves[(size_t)n.c1] : 1;
    leaves[(size_t)i] = c;
  }
  std::vector<Pending> todo;
  out.nodes.push_back(WideNode{0, 0});
  todo.push_back({0, 0, 1});
  std::vector<int32_t> node_of_slot;
  while (!todo.empty()) {
    const Pending cur = todo.back();
    todo.pop_back();
    out.depth = std::max(out.depth, cur.depth);
    std::vector<WItem> items;
    auto add_children = [&](int32_t b2n) {
      const BvhNode& n = bvh2.nodes[(size_t)b2n];
      if (n.c0 != kEmptyChild) items.push_back({n.c0, child_area(n, 0)});
      if (n.c1 != kEmptyChild) items.push_back({n.c1, child_area(n, 1)});
    };
    add_children(cur.bvh2_node);
    // Default ("thin"): every node is opened greedily to 32 children, so the root of a ~500-geom
    // scene points at 32 clusters of ~16 geoms.  B2_FAT_CLUSTERS=1 instead keeps subtrees of <= 32
    // geoms as single flat nodes (fewer, fuller collect rounds) -- measured 2 % slower on B200
    // (profiles/r01_sweeps.txt: 5.49 vs 5.60 Grays/s): the finer cull of thin clusters wins.
    static const bool fat_clusters = [] {
      const char* v = std::getenv("B2_FAT_CLUSTERS");
      return v ? std::atoi(v) != 0 : false;
    }();
    const bool flat = !fat_clusters || leaves[(size_t)cur.bvh2_node] <= kWideFanout;
    while ((int)items.size() < kWideFanout) {
      int best = -1;
      for (int i = 0; i < (int)items.size(); i++) {
        if (items[i].link < 0) continue;
        if (!flat && leaves[(size_t)items[i].link] <= kWideFanout) continue;
        if (best < 0 || items[i].area > items[best].area) best = i;
      }
      if (best < 0) break;
      const int32_t b2n = items[best].link;
      items.erase(items.begin() + best);
      add_children(b2n);
    }
    // geoms first, in id order (keeps coincident duplicates adjacent), then inner children
    std::stable_sort(items.begin(), items.end(), [](const WItem& a, const WItem& b) {
      const bool la = a.link < 0, lb = b.link < 0;
      if (la != lb) return la;
      if (la) return (~a.link) < (~b.link);
      return false;
    });
    WideNode& wn = out.nodes[(size_t)cur.wide_id];
    wn.first = (int32_t)out.slot_link.size();
    wn.count = (int32_t)items.size();
    for (const WItem& it : items) {
      if (it.link < 0) {
        out.slot_link.push_back(it.link);
      } else {
        const int32_t wid = (int32_t)out.nodes.size();
        out.nodes.push_back(WideNode{0, 0});
        out.slot_link.push_back(wid);
        todo.push_back({it.link, wid, cur.depth + 1});
      }
      node_of_slot.push_back(cur.wide_id);
    }
  }
  // heights: children always have a larger id than their parent
  const int nn = (int)out.nodes.size();
  std::vector<int> height(nn, 0);
  for (int i = nn - 1; i >= 0; i--) {
    int h = 0;
    for (int s = out.nodes[i].first; s < out.nodes[i].first + out.nodes[i].count; s++)
      if (out.slot_link[s] >= 0) h = std::max(h, height[out.slot_link[s]] + 1);
    height[i] = h;
  }
  const int nlevel = *std::max_element(height.begin(), height.end()) + 1;
  const int nslot = (int)out.slot_link.size();
  out.refit_level_start.assign(nlevel + 1, 0);
  for (int s = 0; s < nslot; s++) out.refit_level_start[height[node_of_slot[s]] + 1]++;
  for (int l = 0; l < nlevel; l++) out.refit_level_start[l + 1] += out.refit_level_start[l];
  out.refit_slots.resize(nslot);
  std::vector<int> cursor(out.refit_level_start.begin(), out.refit_level_start.end() - 1);
  for (int s = 0; s < nslot; s++) out.refit_slots[cursor[height[node_of_slot[s]]]++] = s;
  return out;
}

}  // namespace b2
