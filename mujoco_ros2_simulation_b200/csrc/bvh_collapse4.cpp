// bvh_collapse4.cpp -- BVH2 -> fanout-4 collapse of the per-mesh triangle BVHs (host, one-off).
#include <algorithm>
#include <cfloat>
#include <vector>

#include "bvh_build.h"

namespace b2 {

namespace {

struct Item4 {
  int32_t link;  // BVH2 link: >= 0 inner node, < 0 leaf payload (~), never kEmptyChild
  float lo[3], hi[3];
  float area;
};

Item4 child_item(const BvhNode& n, int which) {
  Item4 it;
  it.link = which == 0 ? n.c0 : n.c1;
  const float* lo = which == 0 ? n.lo0 : n.lo1;
  const float* hi = which == 0 ? n.hi0 : n.hi1;
  for (int k = 0; k < 3; k++) {
    it.lo[k] = lo[k];
    it.hi[k] = hi[k];
  }
  const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
  it.area = 2.f * (dx * dy + dy * dz + dz * dx);
  return it;
}

}  // namespace

std::vector<MeshNode4> collapse4(const std::vector<BvhNode>& nodes2, int* depth) {
  std::vector<MeshNode4> out;
  int max_depth = 0;
  if (nodes2.empty()) {
    if (depth) *depth = 0;
    return out;
  }
  struct Pending {
    int32_t node2;
    int32_t id4;
    int depth;
  };
  std::vector<Pending> todo;
  out.push_back(MeshNode4{});
  todo.push_back({0, 0, 1});
  while (!todo.empty()) {
    const Pending cur = todo.back();
    todo.pop_back();
    max_depth = std::max(max_depth, cur.depth);
    std::vector<Item4> items;
    auto add_children = [&](int32_t n2) {
      const BvhNode& n = nodes2[(size_t)n2];
      if (n.c0 != kEmptyChild) items.push_back(child_item(n, 0));
      if (n.c1 != kEmptyChild) items.push_back(child_item(n, 1));
    };
    add_children(cur.node2);
    while ((int)items.size() < 4) {
      int best = -1;
      for (int i = 0; i < (int)items.size(); i++) {
        if (items[(size_t)i].link < 0) continue;
        // opening a node with a single child would not add a slot but is still a gain (one hop
        // less); opening one with two children needs a free slot
        if (best < 0 || items[(size_t)i].area > items[(size_t)best].area) best = i;
      }
      if (best < 0) break;
      const int32_t n2 = items[(size_t)best].link;
      items.erase(items.begin() + best);
      add_children(n2);
    }
    MeshNode4 node;
    for (int c = 0; c < 4; c++) {
      node.lox[c] = node.loy[c] = node.loz[c] = FLT_MAX;
      node.hix[c] = node.hiy[c] = node.hiz[c] = -FLT_MAX;
      node.child[c] = kEmptyChild;
      node.pad[c] = 0;
    }
    for (int c = 0; c < (int)items.size(); c++) {
      const Item4& it = items[(size_t)c];
      node.lox[c] = it.lo[0];
      node.loy[c] = it.lo[1];
      node.loz[c] = it.lo[2];
      node.hix[c] = it.hi[0];
      node.hiy[c] = it.hi[1];
      node.hiz[c] = it.hi[2];
      if (it.link < 0) {
        node.child[c] = it.link;
      } else {
        const int32_t id = (int32_t)out.size();
        out.push_back(MeshNode4{});
        node.child[c] = id;
        todo.push_back({it.link, id, cur.depth + 1});
      }
    }
    out[(size_t)cur.id4] = node;
  }
  if (depth) *depth = max_depth;
  return out;
}

}  // namespace b2
