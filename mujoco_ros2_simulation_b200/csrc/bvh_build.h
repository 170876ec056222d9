// bvh_build.h -- host-side binned-SAH BVH2 construction (K4, one-off per model / rebuild).
// Used for the per-mesh triangle BVHs (static, mesh-local frame) and for the topology of the
// geom BVH (whose boxes are then refit on the device every step).
#ifndef B2_BVH_BUILD_H
#define B2_BVH_BUILD_H

#include <vector>

#include "b2_types.h"

namespace b2 {

struct Aabb {
  float lo[3], hi[3];
};

struct BuiltBvh {
  std::vector<BvhNode> nodes;    // children always have a larger index than their parent
  std::vector<int32_t> order;    // primitive ids in leaf order
  // refit schedule for the geom BVH: nodes sorted by height (leaves-only nodes first)
  std::vector<int32_t> refit_nodes;
  std::vector<int32_t> refit_level_start;  // [nlevel+1] offsets into refit_nodes
  int max_depth = 0;
};

// Fanout-32 tree collapsed from a BVH2 whose leaves carry geom ids (the geom BVH).
struct WideBvh {
  std::vector<WideNode> nodes;       // node 0 = root
  std::vector<int32_t> slot_link;    // [nslot] >= 0: wide node id, < 0: ~geom id
  // refit schedule: slots sorted by the height of the node that owns them (height 0 first:
  // a slot of height h only depends on nodes of height < h)
  std::vector<int32_t> refit_slots;
  std::vector<int32_t> refit_level_start;  // [nlevel+1]
  int depth = 0;                           // nodes on the longest root-to-leaf path
};
WideBvh collapse_wide(const BuiltBvh& bvh2);

// Fanout-4 tree collapsed from a BVH2 whose leaves carry triangle ranges (the per-mesh BVHs):
// starting from a node's two children, the inner child with the largest surface area is replaced
// by its own children until there are four or none can be opened.  Node 0 = root, pre-order.
// depth (out): nodes on the longest root-to-leaf path.
std::vector<MeshNode4> collapse4(const std::vector<BvhNode>& nodes2, int* depth);

// prims: one box per primitive.  max_leaf: primitives per leaf (1 for geoms, <=4 for triangles).
// leaf_payload_geom: true  -> leaf payload = primitive id (max_leaf must be 1)
//                    false -> leaf payload = (first << 2) | (count - 1) into `order`
// cost_isect: SAH cost of one primitive test relative to one node visit.
BuiltBvh build_bvh(const std::vector<Aabb>& prims, int max_leaf, bool leaf_payload_geom,
                   float cost_isect);

}  // namespace b2
#endif
