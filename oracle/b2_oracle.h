/*
 * b2_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, double-precision restatement of the sensor-readout hot path of
 * yassine-cherni/mujoco_ros2_simulation:
 *   - the ray arithmetic MuJoCo 3.3.4 runs for every rangefinder inside mj_step
 *     (engine_ray.c / engine_sensor.c; pinned by ref:CMakeLists.txt:44, called at
 *     ref:src/mujoco_system_interface.cpp:1691,1731) -- restated from SURVEY.md Appendix A
 *     because the MuJoCo sources/binaries are absent from this environment;
 *   - MujocoLidar::update gather / float cast / range filter  (ref:src/mujoco_lidar.cpp:256-271);
 *   - MujocoCameras::update step 3, float32 linearise + flip  (ref:src/mujoco_cameras.cpp:222-240);
 *   - camera intrinsics                                       (ref:src/mujoco_cameras.cpp:119-124);
 *   - depth_image_proc::point_cloud_xyz semantics for the PointCloud2 projection.
 *
 * PARITY UNPINNED at the MuJoCo boundary: the reference holds no golden vector for this path
 *  and libmujoco cannot run here.  The oracle is pinned only by the
 * survey's hand-derived known answer for test_robot.xml and by analytic cases (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libb2sense.so) never links or calls it.
 */
#ifndef B2_ORACLE_H
#define B2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MINVAL 1e-15 /* mjMINVAL */
#define ORC_NGROUP 6     /* mjNGROUP */

/* mjtGeom */
enum {
  ORC_GEOM_PLANE = 0,
  ORC_GEOM_HFIELD = 1,
  ORC_GEOM_SPHERE = 2,
  ORC_GEOM_CAPSULE = 3,
  ORC_GEOM_ELLIPSOID = 4,
  ORC_GEOM_CYLINDER = 5,
  ORC_GEOM_BOX = 6,
  ORC_GEOM_MESH = 7
};

/* The mjModel fields the path reads. Arrays are borrowed. */
typedef struct orc_model {
  int32_t ngeom, nbody, nmesh, nmat;
  const int32_t* geom_type;   /* [ngeom] */
  const int32_t* geom_bodyid; /* [ngeom] */
  const int32_t* geom_group;  /* [ngeom] */
  const int32_t* geom_matid;  /* [ngeom] (-1: none) */
  const int32_t* geom_dataid; /* [ngeom] mesh id or -1 */
  const double* geom_size;    /* [ngeom*3] */
  const float* geom_rgba;     /* [ngeom*4] */
  const float* mat_rgba;      /* [nmat*4] */
  const int32_t* body_weldid; /* [nbody] */
  const int32_t* mesh_vertadr; /* [nmesh] */
  const int32_t* mesh_vertnum;
  const int32_t* mesh_faceadr;
  const int32_t* mesh_facenum;
  const float* mesh_vert;   /* [nvert*3] float, mesh-local */
  const int32_t* mesh_face; /* [nface*3] local vertex ids */
  /* per-mesh BVH, mjModel layout (filled by orc_build_mesh_bvh) */
  const int32_t* mesh_bvhadr; /* [nmesh] */
  const int32_t* mesh_bvhnum; /* [nmesh] */
  const double* bvh_aabb;     /* [nbvh*6] centre xyz, half xyz */
  const int32_t* bvh_child;   /* [nbvh*2] -1: none */
  const int32_t* bvh_nodeid;  /* [nbvh] leaf: local face id, else -1 */
} orc_model;

/* The mjData fields the path reads. */
typedef struct orc_data {
  const double* geom_xpos; /* [ngeom*3] */
  const double* geom_xmat; /* [ngeom*9] row-major */
} orc_data;

/* ---- A.3 primitive tests --------------------------------------------------------------- */
double orc_ray_geom(const double pos[3], const double mat[9], const double size[3],
                    const double pnt[3], const double vec[3], int geomtype);
/* n independent (geom, ray) cases: arrays [n*3] / [n*9] / [n]; out[n] (test harness convenience) */
void orc_ray_geom_batch(int n, const double* pos, const double* mat, const double* size,
                        const double* pnt, const double* vec, const int32_t* geomtype, double* out);

/* ---- A.4 mesh ---------------------------------------------------------------------------- */
/* number of BVH nodes a mesh with nface faces needs (2*nface-1) */
int orc_mesh_bvh_size(int nface);
/* builds the mjModel-layout BVH of one mesh; returns nodes written */
int orc_build_mesh_bvh(const float* vert, const int32_t* face, int nface, double* bvh_aabb,
                       int32_t* bvh_child, int32_t* bvh_nodeid);
double orc_ray_mesh(const orc_model* m, const orc_data* d, int geomid, const double pnt[3],
                    const double vec[3]);
/* same result without the BVH (every face tested): validates the BVH cull */
double orc_ray_mesh_brute(const orc_model* m, const orc_data* d, int geomid, const double pnt[3],
                          const double vec[3]);

/* ---- A.2 / A.5 / A.6 --------------------------------------------------------------------- */
int orc_ray_eliminate(const orc_model* m, int geomid, const uint8_t* geomgroup,
                      uint8_t flg_static, int bodyexclude);
double orc_ray(const orc_model* m, const orc_data* d, const double pnt[3], const double vec[3],
               const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, int32_t* geomid);
/* mj_multiRay shape; hits farther than cutoff are reported as misses (dist -1, geomid -1).
 * geomid may be NULL.  nthread <= 1: caller's thread. */
void orc_multi_ray(const orc_model* m, const orc_data* d, const double pnt[3], const double* vec,
                   const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, int32_t* geomid,
                   double* dist, int nray, double cutoff, int nthread);

/* the same with the cull chosen explicitly (orc_multi_ray = ORC_CULL_BODY_ANGULAR, A.6 as specified) */
#define ORC_CULL_NONE 0          /* per-ray brute force with precomputed eliminate flags */
#define ORC_CULL_BODY_ANGULAR 1  /* per-body azimuth / elevation rectangle (A.6) */
#define ORC_CULL_GEOM_SPHERE 2   /* per-geom bounding sphere vs the ray's line */
void orc_multi_ray_cull(const orc_model* m, const orc_data* d, const double pnt[3],
                        const double* vec, const uint8_t* geomgroup, uint8_t flg_static,
                        int bodyexclude, int32_t* geomid, double* dist, int nray, double cutoff,
                        int nthread, int cull);

/* ---- A.7 + lidar.cpp:256-271 --------------------------------------------------------------
 * One scan: world dir i = sensor_mat * local_dir[i], origin sensor_pos; mj_ray(geomgroup,
 * flg_static, bodyexclude); rangefinder cutoff (min(cutoff, v) when cutoff > 0); float cast;
 * (r < range_min || r > range_max) ? -1 : r.  dist_raw (nullable) receives the unfiltered double. */
void orc_lidar_scan(const orc_model* m, const orc_data* d, const double sensor_pos[3],
                    const double sensor_mat[9], const double* local_dir, int nray,
                    const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, double cutoff,
                    double range_min, double range_max, float* ranges, int32_t* geomid,
                    double* dist_raw, int nthread);
/* the same scan through orc_multi_ray_cull (one origin per scan): identical results */
void orc_lidar_scan_multiray(const orc_model* m, const orc_data* d, const double sensor_pos[3],
                             const double sensor_mat[9], const double* local_dir, int nray,
                             const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude,
                             double cutoff, double range_min, double range_max, float* ranges,
                             int32_t* geomid, double* dist_raw, int nthread, int cull);
/* lidar.cpp:92-93 */
int orc_lidar_num_rangefinders(double min_angle, double max_angle, double angle_increment);

/* ---- cameras.cpp:119-124, 222-240 and depth_image_proc semantics --------------------------- */
/* K[0]=K[4]=f, K[2]=W/2, K[5]=H/2, K[8]=1, rest 0 */
void orc_camera_intrinsics(double fovy_deg, int width, int height, double K[9]);
/* gl_depth: bottom-left origin, values in [0,1]; out: top-left origin, metres (float32 math) */
void orc_depth_linearise_flip(const float* gl_depth, int width, int height, double znear,
                              double zfar, double extent, float* depth_flipped);
/* xyz: [H*W*point_floats] (point_floats 3 or 4; 4th float = 0 pad). Non-finite depth -> NaN. */
void orc_depth_to_cloud(const float* depth_flipped, int width, int height, const double K[9],
                        int point_floats, float* cloud);

#ifdef __cplusplus
}
#endif
#endif
