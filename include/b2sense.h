/*
 * b2sense.h -- C ABI of the B200 sensor-readout engine (libb2sense.so).
 *
 * This is the drop-in boundary under the two wrapper classes of the reference plugin
 * (ref = yassine-cherni/mujoco_ros2_simulation):
 *   MujocoLidar    ref:include/mujoco_ros2_control/mujoco_lidar.hpp:74-92,  ref:src/mujoco_lidar.cpp
 *   MujocoCameras  ref:include/mujoco_ros2_control/mujoco_cameras.hpp:82-101, ref:src/mujoco_cameras.cpp
 * and under the rangefinder work MuJoCo does inside mj_step
 * (ref:src/mujoco_system_interface.cpp:1691,1731 -> engine_sensor.c / engine_ray.c, MuJoCo 3.3.4).
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions cross the boundary;
 *   - every call returns B2_OK (0) or a negative b2_status; b2_last_error() gives the text
 *     (thread-local);
 *   - handles are opaque; every output buffer is caller-allocated;
 *   - "_device" entry points take device pointers, enqueue on the scene's stream and return
 *     without blocking; the plain entry points take HOST pointers and block until the result
 *     is in the caller's buffer: pinned (cudaHostAlloc / cudaHostRegister) buffers are written by
 *     the copy engine directly, pageable ones through the library's own pinned staging;
 *   - threads: one scene may be used from several threads at once (the plugin's physics, lidar
 *     and camera threads do): work on a scene is serialised internally.  A lidar or camera handle
 *     belongs to one thread at a time, like the wrapper object that owns it;
 *   - there is NO CPU fallback: every entry point fails with B2_ERR_CUDA when no sm_100 device
 *     is usable.
 *   - mjModel / mjData field names are kept so a binding is `desc.geom_type = m->geom_type;`.
 */
#ifndef B2SENSE_H
#define B2SENSE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_VERSION 100 /* 0.1.0 */
#define B2_NGROUP 6    /* mjNGROUP */

typedef enum b2_status {
  B2_OK = 0,
  B2_ERR_INVALID = -1,     /* bad argument (null handle, negative count, bad enum ...) */
  B2_ERR_CUDA = -2,        /* CUDA runtime error or no usable device */
  B2_ERR_UNSUPPORTED = -3, /* geom type the engine does not implement (hfield, sdf) */
  B2_ERR_NOMEM = -4
} b2_status;

/* mjtGeom values the engine accepts (mujoco/mjmodel.h) */
enum {
  B2_GEOM_PLANE = 0,
  B2_GEOM_SPHERE = 2,
  B2_GEOM_CAPSULE = 3,
  B2_GEOM_ELLIPSOID = 4,
  B2_GEOM_CYLINDER = 5,
  B2_GEOM_BOX = 6,
  B2_GEOM_MESH = 7
};

typedef struct b2_scene b2_scene;
typedef struct b2_lidar b2_lidar;
typedef struct b2_camera b2_camera;

/* Static model: the mjModel fields read by ray_eliminate / mju_rayGeom / mj_rayMesh
 * (SURVEY.md A.1).  Arrays are copied by b2_scene_create; the caller keeps ownership. */
typedef struct b2_model_desc {
  int32_t ngeom, nbody, nmesh, nmat;
  const int32_t* geom_type;    /* [ngeom] mjtGeom */
  const int32_t* geom_bodyid;  /* [ngeom] */
  const int32_t* geom_group;   /* [ngeom] */
  const int32_t* geom_matid;   /* [ngeom] -1: no material */
  const int32_t* geom_dataid;  /* [ngeom] mesh id for mesh geoms */
  const double* geom_size;     /* [ngeom*3] */
  const float* geom_rgba;      /* [ngeom*4] */
  const float* mat_rgba;       /* [nmat*4]  (may be NULL when nmat == 0) */
  const int32_t* body_weldid;  /* [nbody] */
  const int32_t* mesh_vertadr; /* [nmesh] */
  const int32_t* mesh_vertnum; /* [nmesh] */
  const int32_t* mesh_faceadr; /* [nmesh] */
  const int32_t* mesh_facenum; /* [nmesh] */
  const float* mesh_vert;      /* [sum(vertnum)*3] */
  const int32_t* mesh_face;    /* [sum(facenum)*3] vertex ids local to the mesh */
} b2_model_desc;

/* One lidar = a static table of ray directions in the sensor frame plus the mj_ray arguments the
 * rangefinder uses (engine_sensor.c) and the LaserScan limits (ref:src/mujoco_lidar.cpp:95-97). */
typedef struct b2_lidar_desc {
  int32_t nray;
  const double* local_dir;  /* [nray*3] ray i in the sensor frame (need not be unit length) */
  int32_t tile_w;           /* rays per scan line for 2-D tiling of 3-D lidars (0: 1-D) */
  const uint8_t* geomgroup; /* NULL or [B2_NGROUP] (mj_ray geomgroup) */
  uint8_t flg_static;       /* mj_ray flg_static (rangefinder: 1) */
  int32_t bodyexclude;      /* mj_ray bodyexclude (rangefinder: site_bodyid), -1: none */
  double cutoff;            /* sensor_cutoff; <= 0: none */
  double range_min;         /* LaserScan range_min (default 0.0) */
  double range_max;         /* LaserScan range_max (default 1000.0) */
} b2_lidar_desc;

/* One camera model (ref:src/mujoco_cameras.cpp:43-47,119-124,222-224). */
typedef struct b2_camera_desc {
  int32_t width, height; /* cam_resolution */
  double fovy;           /* cam_fovy, degrees */
  double znear, zfar;    /* vis.map.znear / zfar */
  double extent;         /* stat.extent */
  int32_t point_step;    /* PointCloud2 point_step: 12 (x,y,z) or 16 (x,y,z,pad) */
  /* geom groups the ray-cast renderer draws (mjvOption.geomgroup).  geomgroup_set == 0: the
   * reference's mjv_defaultOption (ref:src/mujoco_cameras.cpp:165) = groups 0-2 visible, 3-5 hidden */
  uint8_t geomgroup[B2_NGROUP];
  uint8_t geomgroup_set;
} b2_camera_desc;

/* ---- library ------------------------------------------------------------------------------ */
int b2_version(void);
int b2_device_count(void); /* number of CUDA devices, 0 if none / no driver */
const char* b2_last_error(void);

/* ---- scene mirror --------------------------------------------------------------------------
 * b2_scene_create: called from on_init after LoadModel (ref:src/mujoco_system_interface.cpp:677).
 * Copies the static model, builds the per-mesh triangle BVHs and allocates nworld pose slots.
 * The geom BVH topology is built from the first world of the first b2_scene_update_poses (worlds
 * that update does not cover start as copies of it) and refit on every later update.  stream: a cudaStream_t to enqueue on, or NULL for an internal one. */
int b2_scene_create(const b2_model_desc* model, int32_t nworld, int32_t device, void* stream,
                    b2_scene** out);
/* fails with B2_ERR_INVALID while lidars / cameras created on the scene are still alive */
int b2_scene_destroy(b2_scene* scene);

/* Per-step pose refresh: mjData geom_xpos [nworld*ngeom*3] / geom_xmat [nworld*ngeom*9] for
 * worlds [world0, world0+nworld).  Called by the physics thread right after mj_step
 * (ref:src/mujoco_system_interface.cpp:1691,1731) while it holds sim_mutex_.  After the first
 * update, a host update of ALL worlds only copies into a pinned "latest wins" slot -- no CUDA call,
 * no scene mutex, never waits for the GPU (microseconds); the next cast / render / sync uploads the
 * newest slot on a dedicated copy stream and refits before it runs.  Partial-range and device
 * updates are enqueued at once (upload on the copy stream, refit on the scene's stream). */
int b2_scene_update_poses(b2_scene* scene, int32_t world0, int32_t nworld,
                          const double* geom_xpos, const double* geom_xmat);
int b2_scene_update_poses_device(b2_scene* scene, int32_t world0, int32_t nworld,
                                 const double* d_geom_xpos, const double* d_geom_xmat);
/* Rebuild the shared geom-BVH topology from the current poses of `world` (optional; use when
 * the scene has drifted far from the configuration the topology was built on). */
int b2_scene_rebuild(b2_scene* scene, int32_t world);
int b2_scene_sync(b2_scene* scene); /* wait for everything enqueued on the scene's stream */

/* ---- lidar: replaces the rangefinder pass of mj_sensorPos and MujocoLidar::update steps 1-2
 * (ref:src/mujoco_lidar.cpp:247-271) ------------------------------------------------------- */
int b2_lidar_create(b2_scene* scene, const b2_lidar_desc* desc, b2_lidar** out);
int b2_lidar_destroy(b2_lidar* lidar);

/* nscan scans in one launch.  Scan s is cast in world scan_world[s] (NULL: world s % nworld)
 * from the sensor frame sensor_pos[s*3] / sensor_mat[s*9] (row-major, columns = frame axes).
 * Outputs, each [nscan*nray], any may be NULL:
 *   ranges : LaserScan.ranges -- (float)min(cutoff, dist), -1 outside [range_min, range_max]
 *   geomid : id of the geom hit, -1 for a miss (not affected by the range filter)
 *   dist   : unfiltered double distance (mj_ray return value after cutoff)
 *   xyz    : [nscan*nray*3] hit point in the sensor frame, NaN for a filtered / missed ray   */
int b2_lidar_cast(b2_lidar* lidar, int32_t nscan, const int32_t* scan_world,
                  const double* sensor_pos, const double* sensor_mat, float* ranges,
                  int32_t* geomid, double* dist, float* xyz);
int b2_lidar_cast_device(b2_lidar* lidar, int32_t nscan, const int32_t* d_scan_world,
                         const double* d_sensor_pos, const double* d_sensor_mat, float* d_ranges,
                         int32_t* d_geomid, double* d_dist, float* d_xyz);

/* The blocking call split in two, for callers that sample the sensor frame under their own
 * simulation mutex (ref:src/mujoco_lidar.cpp:247-250 holds sim_mutex_ while it copies sensordata):
 * _begin enqueues everything -- pending pose refresh, sensor poses, the cast, the read-back into
 * pinned staging -- under one hold of the scene mutex and returns without waiting for the GPU, so
 * calling it with the simulation mutex held pins sensor frame and geom poses to the same step; _end
 * waits and copies into the buffers given to _begin (which must stay valid).  One cast in flight per
 * lidar.  Single scans replay a captured CUDA graph from the second call of a shape on. */
int b2_lidar_cast_begin(b2_lidar* lidar, int32_t nscan, const int32_t* scan_world,
                        const double* sensor_pos, const double* sensor_mat, float* ranges,
                        int32_t* geomid, double* dist, float* xyz);
int b2_lidar_cast_end(b2_lidar* lidar);
/* ranges + geom ids as uint16 (0xFFFF = miss; needs ngeom < 65535): 6 B/ray over the host link
 * instead of 8 */
int b2_lidar_cast_u16(b2_lidar* lidar, int32_t nscan, const int32_t* scan_world,
                      const double* sensor_pos, const double* sensor_mat, float* ranges,
                      uint16_t* geomid16);
int b2_lidar_cast_u16_device(b2_lidar* lidar, int32_t nscan, const int32_t* d_scan_world,
                             const double* d_sensor_pos, const double* d_sensor_mat,
                             float* d_ranges, uint16_t* d_geomid16);
/* device time (us) of the last blocking / begin-end host cast: first enqueue to last kernel */
double b2_lidar_last_device_us(const b2_lidar* lidar);

/* mj_multiRay-shaped entry (MuJoCo 3.3.4 engine_ray.c) on `world`.
 * Host pointers, blocking.  Hits farther than cutoff are reported as misses. */
int b2_multiray(b2_scene* scene, int32_t world, const double pnt[3], const double* vec,
                const uint8_t* geomgroup, uint8_t flg_static, int32_t bodyexclude,
                int32_t* geomid, double* dist, int32_t nray, double cutoff);

/* ---- depth camera: replaces MujocoCameras::update step 3 (ref:src/mujoco_cameras.cpp:222-240)
 * and adds the PointCloud2 projection ------------------------------------------------------- */
int b2_camera_create(b2_scene* scene /* nullable */, int32_t device, const b2_camera_desc* desc,
                     b2_camera** out);
int b2_camera_destroy(b2_camera* cam);
/* CameraInfo K (3x3) and P (3x4) as register_cameras fills them (ref:src/mujoco_cameras.cpp:119-124) */
int b2_camera_info(const b2_camera* cam, double K[9], double P[12]);
/* ncam depth buffers of this camera model in one launch.
 *   gl_depth     [ncam*H*W] float, mjr_readPixels layout (bottom-left origin, 0 near .. 1 far)
 *   depth32fc1   [ncam*H*W] float, Image 32FC1 data, top-left origin, metres   (nullable)
 *   cloud        [ncam*H*W*point_step bytes] PointCloud2 data, x,y,z FLOAT32    (nullable)  */
int b2_depth_to_cloud(b2_camera* cam, int32_t ncam, const float* gl_depth, float* depth32fc1,
                      void* cloud);
int b2_depth_to_cloud_device(b2_camera* cam, int32_t ncam, const float* d_gl_depth,
                             float* d_depth32fc1, void* d_cloud);
/* N1: depth by ray casting -- fills a GL-layout depth buffer
 * [ncam*H*W] from the scene mirror for ncam camera poses (MuJoCo camera frame: looks along -Z,
 * +X right, +Y up), so the GL render + readback (ref:src/mujoco_cameras.cpp:202-219) is not needed. */
int b2_camera_render_depth_device(b2_camera* cam, int32_t ncam, const int32_t* d_cam_world,
                                  const double* d_cam_pos, const double* d_cam_mat,
                                  float* d_gl_depth);

/* Host-pointer convenience for the camera thread when no OpenGL readback exists: ray-casts the GL
 * depth of ncam camera poses (cam_world NULL: camera i sees world i % nworld), then runs the fused
 * linearise + flip + cloud step on it.  gl_depth / depth32fc1 / cloud: HOST buffers, each nullable.
 * Blocking. */
int b2_camera_render_to_cloud(b2_camera* cam, int32_t ncam, const int32_t* cam_world,
                              const double* cam_pos, const double* cam_mat, float* gl_depth,
                              float* depth32fc1, void* cloud);

/* begin / end form of the above (see b2_lidar_cast_begin): results are staged through pinned
 * memory, so the output buffers may be pageable (sensor_msgs data vectors) */
int b2_camera_render_to_cloud_begin(b2_camera* cam, int32_t ncam, const int32_t* cam_world,
                                    const double* cam_pos, const double* cam_mat, float* gl_depth,
                                    float* depth32fc1, void* cloud);
int b2_camera_render_to_cloud_end(b2_camera* cam);

/* N2: RGB8 row flip of ncam images of this camera model (ref:src/mujoco_cameras.cpp:243-249) for
 * colour buffers that already live on the device (e.g. CUDA-GL interop).  [ncam*H*W*3] bytes. */
int b2_rgb_flip_device(b2_camera* cam, int32_t ncam, const uint8_t* d_rgb_gl, uint8_t* d_rgb_out);

/* ---- multi-GPU: result gather over NVLink -----------------------
 * [NOT IN REFERENCE: the reference is a single process on one CPU.]  One process per GPU; worlds /
 * sensor-pose batches are sharded, nothing is exchanged while casting.  The gather is fused into
 * the cast: the owner of the gathered array creates a PEER BUFFER (device memory other processes
 * can map) and publishes its 64-byte handle; every other rank opens the handle and passes
 * `peer_ptr + its own offset` as the ranges / geomid / depth output of the `_device` entry points.
 * The kernels' own result stores (128-B float4 / int4 lines) then travel over NVLink / NVSwitch
 * while the traversal runs -- no staging copy and no separate collective.  The writer must
 * b2_scene_sync() (or order a stream event) before the owner reads; ordering between processes
 * is the caller's (one barrier per step). */
#define B2_PEER_HANDLE_BYTES 64
/* allocate `bytes` of device memory on `device` and export its handle */
int b2_peer_buffer_create(int32_t device, size_t bytes, void** d_ptr,
                          uint8_t handle[B2_PEER_HANDLE_BYTES]);
/* map another process's buffer into this process (current device = `device`); fails with
 * B2_ERR_CUDA when the two GPUs are not P2P-connected or the handle is stale */
int b2_peer_buffer_open(int32_t device, const uint8_t handle[B2_PEER_HANDLE_BYTES], void** d_ptr);
int b2_peer_buffer_close(int32_t device, void* d_ptr);   /* a pointer from _open  */
/* Bulk form of the gather: copy `bytes` of device memory (e.g. a finished block of results) to
 * another device pointer (e.g. a slice of a peer buffer) with the copy engine, on the scene's copy
 * stream, ordered after everything enqueued on the scene's stream so far -- the traversal of the
 * next block keeps all SMs while the block crosses NVLink.  b2_scene_join makes the scene's stream
 * wait for the copies enqueued so far (b2_scene_sync waits for both on the host). */
int b2_scene_copy_async(b2_scene* scene, void* d_dst, const void* d_src, size_t bytes);
int b2_scene_join(b2_scene* scene);
/* One-step-lagged form for a streaming loop with double-buffered sources: marks the copies
 * enqueued so far as "step s" and makes the scene's stream wait for the copies of step s-1 only, so
 * the tail of step s's transfers overlaps the first block of step s+1.  Finish with b2_scene_join. */
int b2_scene_copy_fence(b2_scene* scene);
int b2_peer_buffer_destroy(int32_t device, void* d_ptr); /* a pointer from _create */

/* ---- introspection used by bench.py / tests ------------------------------------------------ */
/* number of kernels this library has launched in this process (all scenes / cameras) */
long long b2_launch_count(void);
int b2_scene_stats(const b2_scene* scene, int32_t* n_bvh_nodes, int32_t* n_planes,
                   int64_t* n_mesh_nodes, int64_t* n_mesh_tris);

#ifdef __cplusplus
}
#endif
#endif
