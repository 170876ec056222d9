// kernels.cuh -- launch wrappers of the sm_100a kernels (defined in kernels.cu).
#ifndef B2_KERNELS_CUH
#define B2_KERNELS_CUH

#include <cuda_runtime.h>

#include "b2_types.h"

namespace b2 {

struct RefitParams {
  // inputs: either raw mjData arrays for worlds [0, nworld) of this launch ...
  const double* raw_xpos;  // [nworld][ngeom][3]   (nullable: refit from the packed poses)
  const double* raw_xmat;  // [nworld][ngeom][9]
  // ... written to / read from the packed scene mirror of worlds [world0, world0+nworld)
  GeomPose* pose;           // [nworld_total][ngeom]
  WideChild* wslots;        // [nworld_total][nslot] child slots of the wide geom BVH
  const GeomStatic* gstat;  // [ngeom]
  const WideNode* wnodes;   // [nwnode] shared topology
  const int32_t* slot_link; // [nslot] >= 0: wide node id, < 0: ~geom id
  const int32_t* refit_slots;  // [nslot] slots sorted by the height of their node
  const int32_t* level_start;  // [nlevel+1]
  int32_t ngeom, nslot, nlevel, world0, nworld;
};

struct DepthParams {
  const float* gl_depth;  // [ncam][H][W] bottom-left origin
  float* depth_out;       // [ncam][H][W] top-left origin (nullable)
  float* cloud;           // [ncam][H][W][point_floats] (nullable)
  int32_t ncam, width, height, point_floats;
  float znear, zfar, depth_scale;
  float cx, cy, inv_fx, inv_fy;
};

// K1: pack poses + refit the geom BVH of nworld worlds.  One CTA per world.
cudaError_t launch_refit(const RefitParams& p, cudaStream_t stream);

// K2: cast nscan*nray rays.  stage_tlas: copy the world's geom BVH into shared memory per CTA.
// tiles_per_block: 32-ray tiles one CTA walks through (multiple of the warps per CTA).
// nlaunch is incremented by the number of kernels enqueued (1, or 5 when the mesh work is split off)
cudaError_t launch_cast(const CastParams& p, bool stage_tlas, int tiles_per_block, int sm_count,
                        cudaStream_t stream, int* nlaunch);
// geometry of the tiling the cast kernel uses for a sensor (for host-side sizing)
int cast_tiles_per_scan(int nray);
int cast_threads_per_block();
size_t cast_stage_limit_bytes();
size_t cast_stage_bytes(int nslot, int nwnode, int ngeom);  // shared memory of the staged geom BVH

// K3: depth linearise + flip + PointCloud2 projection.
cudaError_t launch_depth_to_cloud(const DepthParams& p, cudaStream_t stream);

// N2: RGB8 row flip of ncam images (GL bottom-up -> Image top-down).
cudaError_t launch_rgb_flip(const uint8_t* in, uint8_t* out, int ncam, int width, int height,
                            cudaStream_t stream);

}  // namespace b2
#endif
