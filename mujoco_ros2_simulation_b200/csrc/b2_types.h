// b2_types.h -- POD layouts shared by the host runtime and the sm_100a kernels.
//
// HBM layout of the scene mirror (DESIGN.md "Data layout"):
//   GeomStatic  [G]        static per model      32 B  size (f64x3), type, dataid
//   GeomPose    [W][G]     per world, per step   96 B  xpos (f64x3) + xmat (f64x9), 16-B aligned
//   WideChild   [W][S]     geom BVH child slots  32 B  fp32 box + link + geom type (fanout-32 tree)
//   MeshNode4   [sum Nm]   per-mesh triangle BVH 128 B four children's boxes, mesh-local, static
//   MeshTri     [sum T]    triangles in leaf order 48 B  three float4 (xyz + pad)
#ifndef B2_TYPES_H
#define B2_TYPES_H

#include <stdint.h>

#if defined(__CUDACC__)
#define B2_HD __host__ __device__ __forceinline__
#define B2_ALIGN(n) __align__(n)
#else
#define B2_HD inline
#define B2_ALIGN(n) alignas(n)
#endif

namespace b2 {

enum GeomType : int32_t {
  GEOM_PLANE = 0,
  GEOM_HFIELD = 1,
  GEOM_SPHERE = 2,
  GEOM_CAPSULE = 3,
  GEOM_ELLIPSOID = 4,
  GEOM_CYLINDER = 5,
  GEOM_BOX = 6,
  GEOM_MESH = 7
};

constexpr double kMinVal = 1e-15;  // mjMINVAL
constexpr int kEmptyChild = INT32_MIN;

// A BVH2 node holding the boxes of both children (so one 64-B fetch decides both descents).
// child >= 0: inner node index; child < 0: leaf, payload = ~child
//   geom BVH : payload = geom id
//   mesh BVH : payload = (first_triangle << 2) | (count - 1), triangles relative to the mesh base
// An absent child has an inverted box (lo = +FLT_MAX, hi = -FLT_MAX) and link kEmptyChild.
struct B2_ALIGN(16) BvhNode {
  float lo0[3], hi0[3];
  float lo1[3], hi1[3];
  int32_t c0, c1;
  int32_t pad0, pad1;
};
static_assert(sizeof(BvhNode) == 64, "BvhNode must be 64 B");

// The per-mesh triangle BVHs are collapsed to fanout 4 (bvh_build.cpp collapse4): half as many
// dependent node fetches per ray as the BVH2 -- the mesh kernels are bound by exactly that latency.
// Boxes are stored per axis for the four children (seven 16-B loads per node).
//   child >= 0: MeshNode4 index (mesh-relative); child < 0: leaf, payload = ~child =
//   (first_triangle << 2) | (count - 1); absent: kEmptyChild with an inverted box.
struct B2_ALIGN(16) MeshNode4 {
  float lox[4], loy[4], loz[4];
  float hix[4], hiy[4], hiz[4];
  int32_t child[4];
  int32_t pad[4];
};
static_assert(sizeof(MeshNode4) == 128, "MeshNode4 must be 128 B");

// The geom BVH is a wide tree (fanout <= 32) so that one warp tests all children of a node at
// once, one lane per child.  A child slot holds the child's box and its link
// (>= 0: wide node id, < 0: ~geom id).  Slots of one node are contiguous.
constexpr int kWideFanout = 32;
struct B2_ALIGN(16) WideChild {
  float lo[3];
  int32_t link;
  float hi[3];
  int32_t pad;
};
static_assert(sizeof(WideChild) == 32, "WideChild must be 32 B");

struct WideNode {
  int32_t first;  // first child slot
  int32_t count;  // number of child slots (1..32)
};

struct B2_ALIGN(16) GeomStatic {
  double size[3];
  int32_t type;
  int32_t dataid;
};
static_assert(sizeof(GeomStatic) == 32, "GeomStatic must be 32 B");

struct B2_ALIGN(16) GeomPose {
  double pos[3];
  double mat[9];
};
static_assert(sizeof(GeomPose) == 96, "GeomPose must be 96 B");

struct B2_ALIGN(16) MeshTri {
  float v0[4], v1[4], v2[4];
};
static_assert(sizeof(MeshTri) == 48, "MeshTri must be 48 B");

struct B2_ALIGN(16) MeshInfo {
  int32_t node_base;  // first MeshNode4 of this mesh in the mesh-node array
  int32_t tri_base;   // first MeshTri of this mesh
  int32_t nnode;
  int32_t ntri;
  // bounding sphere of the mesh's triangles in the mesh frame (centre, radius inflated by 1e-6): a ray
  // that misses it meets no triangle, so it needs no (ray, mesh) pair -- tighter than the geom_size box
  // for anything roundish, and decided by sign alone (sphere_miss)
  float bs_c[3];
  float bs_r;
};
static_assert(sizeof(MeshInfo) == 32, "MeshInfo must be 32 B");

// ray generation modes of the cast kernel
enum CastMode : int32_t {
  MODE_LIDAR = 0,     // dir = sensor_mat * local_dir[r], LaserScan epilogue
  MODE_MULTIRAY = 1,  // dir = vec[r] as given (mj_multiRay), (geomid, dist) epilogue
  MODE_CAMERA = 2     // dir from pinhole intrinsics, GL depth epilogue
};

// Scratch of the split mesh path (K2b-K2d).  K2 appends one PAIR per (ray, mesh geom whose box the
// ray enters) and one AFFECTED record per ray that has at least one pair; the mesh kernels trace the
// pairs with full warps, merge the distances with 64-bit atomicMin on the IEEE bit pattern (a
// non-negative double orders like its bits) and rewrite the outputs of the affected rays.
struct MeshSplit {
  int32_t* pair_slot;              // [pair_cap] affected-record slot of the pair's ray
  int32_t* pair_geom;              // [pair_cap] mesh geom id
  double* pair_x;                  // [pair_cap] mj_rayMesh result of the pair
  uint32_t* aff_ray;               // [aff_cap] ray index within the launch (scan * nray + r)
  unsigned long long* aff_bits;    // [aff_cap] current best distance bits (prim hit or +inf)
  unsigned long long* aff_prim;    // [aff_cap] distance bits of the primitive-only hit
  int32_t* aff_prim_id;            // [aff_cap]
  int32_t* aff_mesh_id;            // [aff_cap] lowest mesh geom id reaching aff_bits
  int32_t* counters;               // [4] npair, naff, overflow, unused
  int32_t pair_cap, aff_cap;
};
constexpr unsigned long long kInfBits = 0x7ff0000000000000ull;

// Scratch of the two-kernel cast (K2a collect -> K2b drain).  K2a writes, per group of tiles, the
// distance-sorted list of candidate geoms (as child-slot indices of the geom BVH); K2b walks it per
// tile.  A group K2a cannot finish in one pass (candidate overflow, rays outside the float range)
// gets count = -1 and is appended to ovf_list: the fused kernel traces those groups.
struct B2_ALIGN(8) CandRef {
  float key2;    // squared distance from the origin to the candidate's box, rounded down
  int32_t slot;  // child slot of the world's geom BVH (box, ~geom id, geom type)
};
struct CastLists {
  int32_t* count;     // [ngroup] candidates of the group, or -1
  CandRef* list;      // [ngroup][kCandCap]
  int32_t* ovf_list;  // [ngroup] groups left to the fused kernel
  int32_t* counters;  // [4] novf, tile cursor (K2b), group cursor (fused fallback), unused
};

// Everything one cast launch reads.  Passed by value as a __grid_constant__ kernel parameter.
struct CastParams {
  MeshSplit ms;  // used when the launch splits the mesh work off (ms.counters != nullptr)
  CastLists cl;  // used by the two-kernel cast (cl.count != nullptr)
  // scene mirror
  const WideChild* wslots;  // [nworld][nslot] geom BVH child slots (boxes refit per step)
  const WideNode* wnodes;   // [nwnode] geom BVH topology (node 0 = root)
  const GeomPose* pose;     // [nworld][ngeom]
  const GeomStatic* gstat;  // [ngeom]
  const int32_t* plane_ids; // [nplane] geoms kept outside the BVH
  const MeshInfo* mesh;     // [nmesh]
  const MeshNode4* mnodes;  // fanout-4 triangle BVHs of all meshes
  const MeshTri* mtris;
  int32_t nslot, nwnode, ngeom, nplane, nworld;
  // sensor
  const uint8_t* elim;      // [ngeom] ray_eliminate flags for this sensor
  const double* local_dir;  // [nray*3] (lidar table or multiray vec)
  int32_t nray, tile_w, mode;
  int32_t group_tiles;      // tiles whose rays share one collect: 4, or fewer for coarse sensors (0 = 4)
  double cutoff, range_min, range_max;
  // camera mode
  int32_t cam_w, cam_h;
  double cam_f, cam_cx, cam_cy;
  float cam_near, cam_far;
  // batch
  int32_t nscan;
  int32_t scan0;              // index of this launch's first scan in the caller's batch (world mapping)
  const int32_t* scan_world;  // nullable
  int32_t* work_counter;      // [1] next tile group of a persistent (dynamic) K2 launch; nullable
  const double* spos;         // [nscan*3]
  const double* smat;         // [nscan*9]
  // outputs (nullable)
  float* ranges;
  int32_t* geomid;
  uint16_t* geomid16;  // compact ids (0xFFFF = miss) for scenes with < 65535 geoms
  double* dist;
  float* xyz;
  float* gl_depth;
};

}  // namespace b2
#endif
