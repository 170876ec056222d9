// scene_build.h -- host-side, CUDA-free construction of the static scene data (K4 inputs):
// ray_eliminate flags, per-mesh triangle BVHs and the geom-BVH topology.  Shared by the runtime
// (runtime.cu) and by the CPU emulation harness under tests/emu (test infrastructure).
#ifndef B2_SCENE_BUILD_H
#define B2_SCENE_BUILD_H

#include <string>
#include <vector>

#include "b2_types.h"
#include "bvh_build.h"

namespace b2 {

struct HostModel {
  int ngeom = 0, nbody = 0, nmesh = 0, nmat = 0;
  std::vector<int32_t> geom_type, geom_bodyid, geom_group, geom_matid, geom_dataid, body_weldid;
  std::vector<double> geom_size;
  std::vector<float> geom_rgba, mat_rgba;
  std::vector<int32_t> plane_ids;  // geoms kept outside the BVH
  std::vector<int32_t> bvh_geoms;  // all other geoms
};

struct HostMeshes {
  std::vector<MeshInfo> infos;
  std::vector<MeshNode4> nodes;  // fanout-4 trees of all meshes, MeshInfo::node_base apart
  std::vector<MeshTri> tris;
};

// ray_eliminate for every geom, evaluated once per sensor
void eliminate_flags(const HostModel& m, const uint8_t* geomgroup, uint8_t flg_static,
                     int bodyexclude, std::vector<uint8_t>& out);

// returns false and sets err on bad input / unsupported depth
bool build_mesh_bvhs(int nmesh, const int32_t* vertadr, const int32_t* vertnum,
                     const int32_t* faceadr, const int32_t* facenum, const float* vert,
                     const int32_t* face, int max_leaf, int max_depth, HostMeshes& out,
                     std::string& err);

// geom-BVH topology from one world's poses: binned-SAH BVH2 collapsed to fanout 32; leaf links
// are ~geom id.  max_depth: deepest wide tree the traversal's frontier queue can hold.
bool build_geom_topology(const HostModel& m, const double* xpos, const double* xmat, int max_depth,
                         WideBvh& out, std::string& err);

}  // namespace b2
#endif
