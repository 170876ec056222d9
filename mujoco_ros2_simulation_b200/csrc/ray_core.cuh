// ray_core.cuh -- the ray arithmetic and BVH traversal of the cast kernel (K2) and the geom
// bounding boxes of the refit kernel (K1).
//
// Written as __host__ __device__ code: the sm_100a kernels in kernels.cu instantiate it with real
// warp votes; tests/emu compiles the very same functions for the CPU with a 1-lane "warp" so the
// traversal logic can be checked against the oracle without a GPU.  That emulation is test
// infrastructure only -- the product library contains no CPU path.
//
// Numerics: every accepted hit distance is computed in IEEE double with the operation order of
// MuJoCo 3.3.4 engine_ray.c, compiled with FMA contraction OFF
// (nvcc -fmad=false), so distances are bit-identical to the CPU oracle.  fp32 is used only for
// conservative culling (tile cone vs box, slab tests against outward-rounded boxes) and can never
// drop a hit the double test would accept.
//
// The warp-cooperative traversal itself lives in tile_trace.cuh (included at the end).
#ifndef B2_RAY_CORE_CUH
#define B2_RAY_CORE_CUH

#include <float.h>
#include <math.h>
#include <string.h>

#include "b2_types.h"

namespace b2 {

#if defined(__CUDA_ARCH__)
#define B2_ON_DEVICE 1
#else
#define B2_ON_DEVICE 0
#endif
// heavy leaf routines: one out-of-line copy on the device keeps the kernel's hot loops inside the
// instruction cache (the fully inlined first version was 140 KB of SASS and stalled on fetch)
#if defined(__CUDACC__) && defined(B2_INLINE_LEAF)
#define B2_LEAF static __host__ __device__ __forceinline__
#elif defined(__CUDACC__)
#define B2_LEAF static __host__ __device__ __noinline__
#else
#define B2_LEAF static inline
#endif

constexpr int kMeshStack = 64;    // mesh-BVH stack entries per lane
constexpr int kQueueCap = 160;    // collect frontier (wide node ids) per warp: 1 + 31 * depth
#ifndef B2_CAND_CAP
#define B2_CAND_CAP 96
#endif
constexpr int kCandCap = B2_CAND_CAP;  // candidate geoms per group between two drains

struct B2_ALIGN(16) F4 {
  float x, y, z, w;
};

B2_HD float int_as_f32(int32_t i) {
#if B2_ON_DEVICE
  return __int_as_float(i);
#else
  float f;
  memcpy(&f, &i, 4);
  return f;
#endif
}
B2_HD int32_t f32_as_int(float f) {
#if B2_ON_DEVICE
  return __float_as_int(f);
#else
  int32_t i;
  memcpy(&i, &f, 4);
  return i;
#endif
}

// ---------------------------------------------------------------------------------------------
// directed rounding
// ---------------------------------------------------------------------------------------------
B2_HD float f32_up(double x) {
#if B2_ON_DEVICE
  return __double2float_ru(x);
#else
  float f = (float)x;
  if ((double)f < x) f = nextafterf(f, INFINITY);
  return f;
#endif
}
B2_HD float f32_down(double x) {
#if B2_ON_DEVICE
  return __double2float_rd(x);
#else
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -INFINITY);
  return f;
#endif
}
B2_HD float fmul_up(float a, float b) {
#if B2_ON_DEVICE
  return __fmul_ru(a, b);
#else
  return f32_up((double)a * (double)b);  // the double product of two floats is exact
#endif
}
B2_HD float fmul_down(float a, float b) {
#if B2_ON_DEVICE
  return __fmul_rd(a, b);
#else
  return f32_down((double)a * (double)b);
#endif
}
B2_HD float fma_rn(float a, float b, float c) {
#if B2_ON_DEVICE
  return __fmaf_rn(a, b, c);
#else
  return fmaf(a, b, c);
#endif
}

// ---------------------------------------------------------------------------------------------
// read-only loads of scene records
// ---------------------------------------------------------------------------------------------
B2_HD BvhNode ld_node(const BvhNode* ptr) {
#if B2_ON_DEVICE
  BvhNode n;
  const float4* q = reinterpret_cast<const float4*>(ptr);
  float4* o = reinterpret_cast<float4*>(&n);
  o[0] = __ldg(q);
  o[1] = __ldg(q + 1);
  o[2] = __ldg(q + 2);
  o[3] = __ldg(q + 3);
  return n;
#else
  return *ptr;
#endif
}
// boxes and links of a fanout-4 mesh node (the pad quarter is not fetched)
B2_HD MeshNode4 ld_node4(const MeshNode4* ptr) {
#if B2_ON_DEVICE
  MeshNode4 n;
  const float4* q = reinterpret_cast<const float4*>(ptr);
  float4* o = reinterpret_cast<float4*>(&n);
#pragma unroll
  for (int k = 0; k < 7; k++) o[k] = __ldg(q + k);
  return n;
#else
  return *ptr;
#endif
}
// one child slot of the wide geom BVH: shared memory when staged, else the read-only path
template <bool STAGE>
B2_HD WideChild ld_wslot(const WideChild* ptr) {
#if B2_ON_DEVICE
  WideChild c;
  const float4* q = reinterpret_cast<const float4*>(ptr);
  float4* o = reinterpret_cast<float4*>(&c);
  if (STAGE) {
    o[0] = q[0];
    o[1] = q[1];
  } else {
    o[0] = __ldg(q);
    o[1] = __ldg(q + 1);
  }
  return c;
#else
  return *ptr;
#endif
}
B2_HD float rsqrt_f32(float x) {
#if B2_ON_DEVICE
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
B2_HD MeshTri ld_tri(const MeshTri* ptr) {
#if B2_ON_DEVICE
  MeshTri t;
  const float4* q = reinterpret_cast<const float4*>(ptr);
  float4* o = reinterpret_cast<float4*>(&t);
  o[0] = __ldg(q);
  o[1] = __ldg(q + 1);
  o[2] = __ldg(q + 2);
  return t;
#else
  return *ptr;
#endif
}
B2_HD GeomPose ld_pose(const GeomPose* ptr) {
#if B2_ON_DEVICE
  GeomPose P;
  const double2* q = reinterpret_cast<const double2*>(ptr);
  double2* o = reinterpret_cast<double2*>(&P);
#pragma unroll
  for (int i = 0; i < 6; i++) o[i] = __ldg(q + i);
  return P;
#else
  return *ptr;
#endif
}
B2_HD GeomStatic ld_gstat(const GeomStatic* ptr) {
#if B2_ON_DEVICE
  GeomStatic g;
  const double2* q = reinterpret_cast<const double2*>(ptr);
  double2* o = reinterpret_cast<double2*>(&g);
  o[0] = __ldg(q);
  o[1] = __ldg(q + 1);
  return g;
#else
  return *ptr;
#endif
}
B2_HD int32_t ld_geom_type(const GeomStatic* ptr) {
#if B2_ON_DEVICE
  return __ldg(&ptr->type);
#else
  return ptr->type;
#endif
}
B2_HD uint8_t ld_u8(const uint8_t* ptr) {
#if B2_ON_DEVICE
  return __ldg(ptr);
#else
  return *ptr;
#endif
}
// start pulling a record towards L1 (no-op on the host).  Opt-in (-DB2_PREFETCH): prefetching the
// candidates' poses during the collect measured 0.9 % SLOWER than not doing it
// (profiles/r01_sweeps.txt: 6.153 vs 6.209 Grays/s), so it is off.
B2_HD void prefetch_l1(const void* ptr) {
#if B2_ON_DEVICE && defined(B2_PREFETCH)
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
#else
  (void)ptr;
#endif
}

// ---------------------------------------------------------------------------------------------
// fp64 closed-form tests (A.3).  Return distance x >= 0 in units of |vec|, or -1.
// ---------------------------------------------------------------------------------------------
B2_HD double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

B2_HD void ray_map(const GeomPose& P, const double pnt[3], const double vec[3], double lpnt[3],
                   double lvec[3]) {
  const double d0 = pnt[0] - P.pos[0], d1 = pnt[1] - P.pos[1], d2 = pnt[2] - P.pos[2];
  lpnt[0] = P.mat[0] * d0 + P.mat[3] * d1 + P.mat[6] * d2;
  lpnt[1] = P.mat[1] * d0 + P.mat[4] * d1 + P.mat[7] * d2;
  lpnt[2] = P.mat[2] * d0 + P.mat[5] * d1 + P.mat[8] * d2;
  lvec[0] = P.mat[0] * vec[0] + P.mat[3] * vec[1] + P.mat[6] * vec[2];
  lvec[1] = P.mat[1] * vec[0] + P.mat[4] * vec[1] + P.mat[7] * vec[2];
  lvec[2] = P.mat[2] * vec[0] + P.mat[5] * vec[1] + P.mat[8] * vec[2];
}

// a*x^2 + 2*b*x + c = 0
// (kept inline: an out-of-line copy shrinks K2 by 10 % of its SASS but measured 0.7 % slower)
B2_HD double ray_quad(double a, double b, double c, double& x0, double& x1) {
  double det = b * b - a * c;
  if (det < kMinVal) {
    x0 = -1;
    x1 = -1;
    return -1;
  }
  det = sqrt(det);
  x0 = (-b - det) / a;
  x1 = (-b + det) / a;
  if (x0 >= 0) return x0;
  if (x1 >= 0) return x1;
  return -1;
}

// ray_quad's RETURN VALUE only (callers that never look at the second root): the same operations in
// the same order, but the second division is paid only when the first root is negative.
B2_HD double ray_quad_first(double a, double b, double c) {
  double det = b * b - a * c;
  if (det < kMinVal) return -1;
  det = sqrt(det);
  const double x0 = (-b - det) / a;
  if (x0 >= 0) return x0;
  const double x1 = (-b + det) / a;
  if (x1 >= 0) return x1;
  return -1;
}

B2_HD double ray_sphere(const double pos[3], double dist_sqr, const double pnt[3],
                        const double vec[3]) {
  const double d0 = pnt[0] - pos[0], d1 = pnt[1] - pos[1], d2 = pnt[2] - pos[2];
  const double a = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  const double b = vec[0] * d0 + vec[1] * d1 + vec[2] * d2;
  const double c = d0 * d0 + d1 * d1 + d2 * d2 - dist_sqr;
  return ray_quad_first(a, b, c);
}

// true iff ray_sphere(pos, dist_sqr, pnt, vec) < 0 -- the bounding-sphere pre-test of capsule /
// cylinder / box (and of mj_rayMesh's box) only looks at the sign.  a, b, c, det are computed as
// ray_sphere computes them; the square root and the two divisions are skipped when the sign is
// certain.  With a > 0 and s = sqrt(det) > 0 the roots x0 = (-b - s)/a <= x1 = (-b + s)/a are both
// negative iff s < b, and a quotient has the sign of its numerator (it cannot round to -0 while
// a < 1e290 and |numerator| >= 1e-20, which s >= sqrt(mjMINVAL) guarantees below):
//   b <= 0                   : x1 >= 0, a hit;
//   det < b*b * (1 - 1e-12)  : s < b even after the roundings of b*b, sqrt (each <= 2^-53), a miss;
//   det > b*b * (1 + 1e-12)  : s > b, a hit;
//   anything else (incl. non-finite or huge operands): the upstream arithmetic as written.
B2_HD bool sphere_miss(const double pos[3], double dist_sqr, const double pnt[3],
                       const double vec[3]) {
  const double d0 = pnt[0] - pos[0], d1 = pnt[1] - pos[1], d2 = pnt[2] - pos[2];
  const double a = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  const double b = vec[0] * d0 + vec[1] * d1 + vec[2] * d2;
  const double c = d0 * d0 + d1 * d1 + d2 * d2 - dist_sqr;
  const double bb = b * b;
  double det = bb - a * c;
  if (det < kMinVal) return true;
#ifndef B2_NO_SIGN_PRETEST
  if (a < 1e290 && det < 1e290 && fabs(b) < 1e140) {
    if (b <= 0) return false;
    if (det < bb * (1.0 - 1e-12)) return true;
    if (det > bb * (1.0 + 1e-12)) return false;
  }
#endif
  det = sqrt(det);
  const double x0 = (-b - det) / a;
  const double x1 = (-b + det) / a;
  return !(x0 >= 0) && !(x1 >= 0);
}

B2_HD void keep_min(double& x, double sol) {
  if (x < 0 || sol < x) x = sol;
}

// local-frame parts of the tests (after ray_map and, where upstream has one, the bounding-sphere
// pre-test)
B2_HD double local_plane(const double size[3], const double lpnt[3], const double lvec[3]) {
  if (lvec[2] > -kMinVal) return -1;
  const double x = -lpnt[2] / lvec[2];
  if (x < 0) return -1;
  const double p0 = lpnt[0] + x * lvec[0];
  const double p1 = lpnt[1] + x * lvec[1];
  if ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) return x;
  return -1;
}

B2_HD double local_capsule(const double size[3], const double lpnt[3], const double lvec[3]) {
  double x = -1, x0, x1;
  const double sq_size0 = size[0] * size[0];
  // round side
  double a = lvec[0] * lvec[0] + lvec[1] * lvec[1];
  double b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1];
  double c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - sq_size0;
  const double sol = ray_quad_first(a, b, c);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) keep_min(x, sol);
  a += lvec[2] * lvec[2];
  // caps: side = +1 top (accept z >= size1), side = -1 bottom (accept z <= -size1)
#pragma unroll 1
  for (int side = 1; side >= -1; side -= 2) {
    const double ldif[3] = {lpnt[0], lpnt[1], side > 0 ? lpnt[2] - size[1] : lpnt[2] + size[1]};
    b = dot3(lvec, ldif);
    c = dot3(ldif, ldif) - sq_size0;
    ray_quad(a, b, c, x0, x1);
    const double z0 = lpnt[2] + x0 * lvec[2], z1 = lpnt[2] + x1 * lvec[2];
    if (x0 >= 0 && (side > 0 ? z0 >= size[1] : z0 <= -size[1])) keep_min(x, x0);
    if (x1 >= 0 && (side > 0 ? z1 >= size[1] : z1 <= -size[1])) keep_min(x, x1);
  }
  return x;
}

B2_HD double local_ellipsoid(const double size[3], const double lpnt[3], const double lvec[3]) {
  const double s0 = 1 / (size[0] * size[0]), s1 = 1 / (size[1] * size[1]),
               s2 = 1 / (size[2] * size[2]);
  const double a = s0 * lvec[0] * lvec[0] + s1 * lvec[1] * lvec[1] + s2 * lvec[2] * lvec[2];
  const double b = s0 * lvec[0] * lpnt[0] + s1 * lvec[1] * lpnt[1] + s2 * lvec[2] * lpnt[2];
  const double c = s0 * lpnt[0] * lpnt[0] + s1 * lpnt[1] * lpnt[1] + s2 * lpnt[2] * lpnt[2] - 1;
  return ray_quad_first(a, b, c);
}

B2_HD double local_cylinder(const double size[3], const double lpnt[3], const double lvec[3]) {
  double x = -1;
  if (fabs(lvec[2]) > kMinVal) {
#pragma unroll 1
    for (int side = -1; side <= 1; side += 2) {
      const double sol = (side * size[1] - lpnt[2]) / lvec[2];
      if (sol >= 0) {
        const double p0 = lpnt[0] + sol * lvec[0];
        const double p1 = lpnt[1] + sol * lvec[1];
        if (p0 * p0 + p1 * p1 <= size[0] * size[0]) keep_min(x, sol);
      }
    }
  }
  const double a = lvec[0] * lvec[0] + lvec[1] * lvec[1];
  const double b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1];
  const double c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0];
  const double sol = ray_quad_first(a, b, c);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) keep_min(x, sol);
  return x;
}

// box faces: the axis loop is kept rolled by rotating (axis, other, other); the acceptance test is
// symmetric in the two other axes and keep_min is order independent, so every accepted `sol` is
// the same arithmetic as upstream's ifcs table.
B2_HD double local_box(const double size[3], const double lpnt[3], const double lvec[3]) {
  double x = -1;
  double a0 = lpnt[0], a1 = lpnt[1], a2 = lpnt[2];
  double b0 = lvec[0], b1 = lvec[1], b2 = lvec[2];
  double s0 = size[0], s1 = size[1], s2 = size[2];
  // The two face divisions of an axis are issued together (independent, so their fp64 latency
  // overlaps); the axis loop stays rolled unless B2_BOX_UNROLL is set -- the kernel is on the edge
  // of the instruction cache and the fully unrolled body cost more in fetch stalls than it gained
  // (profiles/r01l).  side * s0 for side = -1 / +1 is exactly -s0 / s0: arithmetic unchanged.
#ifdef B2_BOX_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
  for (int it = 0; it < 3; it++) {
    const bool ok = fabs(b0) > kMinVal;
    const double solm = (-s0 - a0) / b0;
    const double solp = (s0 - a0) / b0;
    if (ok && solm >= 0) {
      const double p0 = a1 + solm * b1;
      const double p1 = a2 + solm * b2;
      if (fabs(p0) <= s1 && fabs(p1) <= s2) keep_min(x, solm);
    }
    if (ok && solp >= 0) {
      const double p0 = a1 + solp * b1;
      const double p1 = a2 + solp * b2;
      if (fabs(p0) <= s1 && fabs(p1) <= s2) keep_min(x, solp);
    }
    double t = a0;
    a0 = a1;
    a1 = a2;
    a2 = t;
    t = b0;
    b0 = b1;
    b1 = b2;
    b2 = t;
    t = s0;
    s0 = s1;
    s1 = s2;
    s2 = t;
  }
  return x;
}

// a pose is fetched in two steps so that a test which ends at the bounding-sphere pre-test (or is a
// sphere) never touches the rotation: pos + mat[0] (two 16-B loads), then mat[1..8] (four)
B2_HD void ld_pose_head(const GeomPose* ptr, double pos[3], double& m0) {
#if B2_ON_DEVICE
  const double2* q = reinterpret_cast<const double2*>(ptr);
  const double2 a = __ldg(q), b = __ldg(q + 1);
  pos[0] = a.x;
  pos[1] = a.y;
  pos[2] = b.x;
  m0 = b.y;
#else
  pos[0] = ptr->pos[0];
  pos[1] = ptr->pos[1];
  pos[2] = ptr->pos[2];
  m0 = ptr->mat[0];
#endif
}
B2_HD void ld_pose_tail(const GeomPose* ptr, double m0, double mat[9]) {
  mat[0] = m0;
#if B2_ON_DEVICE
  const double2* q = reinterpret_cast<const double2*>(ptr);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double2 v = __ldg(q + 2 + i);
    mat[1 + 2 * i] = v.x;
    mat[2 + 2 * i] = v.y;
  }
#else
  for (int i = 1; i < 9; i++) mat[i] = ptr->mat[i];
#endif
}

// mju_rayGeom for one primitive geom (also the geom_size box of a mesh geom when type = GEOM_BOX
// is forced by the caller).  Out of line on the device: one copy serves every call site.
B2_LEAF double prim_test(int type, const GeomPose* gp, const GeomStatic* gsp, double px, double py,
                         double pz, double vx, double vy, double vz) {
  GeomPose P;
  ld_pose_head(gp, P.pos, P.mat[0]);
  const GeomStatic gs = ld_gstat(gsp);
  const double pnt[3] = {px, py, pz}, vec[3] = {vx, vy, vz};
  if (type == GEOM_SPHERE) return ray_sphere(P.pos, gs.size[0] * gs.size[0], pnt, vec);
  // bounding-sphere pre-tests of capsule / cylinder / box
  if (type == GEOM_CAPSULE || type == GEOM_CYLINDER || type == GEOM_BOX) {
    double ssz;
    if (type == GEOM_CAPSULE) {
      const double s = gs.size[0] + gs.size[1];
      ssz = s * s;
    } else if (type == GEOM_CYLINDER) {
      ssz = gs.size[0] * gs.size[0] + gs.size[1] * gs.size[1];
    } else {
      ssz = gs.size[0] * gs.size[0] + gs.size[1] * gs.size[1] + gs.size[2] * gs.size[2];
    }
    if (sphere_miss(P.pos, ssz, pnt, vec)) return -1;
  }
  ld_pose_tail(gp, P.mat[0], P.mat);
#ifndef B2_NO_PLANE_FAST
  if (type == GEOM_PLANE && gs.size[0] <= 0 && gs.size[1] <= 0) {
    // unbounded plane (the usual ground plane, tested for every ray): only the local z components
    // of ray_map decide; they are computed exactly as ray_map computes them
    const double d0 = px - P.pos[0], d1 = py - P.pos[1], d2 = pz - P.pos[2];
    const double lpz = P.mat[2] * d0 + P.mat[5] * d1 + P.mat[8] * d2;
    const double lvz = P.mat[2] * vx + P.mat[5] * vy + P.mat[8] * vz;
    if (lvz > -kMinVal) return -1;
    const double x = -lpz / lvz;
    return x < 0 ? -1 : x;
  }
#endif
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  switch (type) {
    case GEOM_PLANE:
      return local_plane(gs.size, lpnt, lvec);
    case GEOM_CAPSULE:
      return local_capsule(gs.size, lpnt, lvec);
    case GEOM_ELLIPSOID:
      return local_ellipsoid(gs.size, lpnt, lvec);
    case GEOM_CYLINDER:
      return local_cylinder(gs.size, lpnt, lvec);
    case GEOM_BOX:
      return local_box(gs.size, lpnt, lvec);
    default:
      return -1;
  }
}

// prim_test(GEOM_PLANE, ...) inline: the same arithmetic without the call, for kernels that test
// nothing but planes (K2a) or test them outside the candidate loop (K2b)
B2_HD double plane_test(const GeomPose* gp, const GeomStatic* gsp, const double pnt[3],
                        const double vec[3]) {
  const GeomPose P = ld_pose(gp);
  const GeomStatic gs = ld_gstat(gsp);
  if (gs.size[0] <= 0 && gs.size[1] <= 0) {
    const double d0 = pnt[0] - P.pos[0], d1 = pnt[1] - P.pos[1], d2 = pnt[2] - P.pos[2];
    const double lpz = P.mat[2] * d0 + P.mat[5] * d1 + P.mat[8] * d2;
    const double lvz = P.mat[2] * vec[0] + P.mat[5] * vec[1] + P.mat[8] * vec[2];
    if (lvz > -kMinVal) return -1;
    const double x = -lpz / lvz;
    return x < 0 ? -1 : x;
  }
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  return local_plane(gs.size, lpnt, lvec);
}

// ---------------------------------------------------------------------------------------------
// mesh triangle test (A.4)
// ---------------------------------------------------------------------------------------------
B2_HD void normalize3(double v[3]) {
  const double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n < kMinVal) {
    v[0] = 1;
    v[1] = 0;
    v[2] = 0;
  } else {
    const double inv = 1 / n;
    v[0] *= inv;
    v[1] *= inv;
    v[2] *= inv;
  }
}

B2_HD void mesh_ray_basis(const double lvec[3], double b0[3], double b1[3]) {
  b0[0] = b0[1] = b0[2] = 1;
  if (fabs(lvec[0]) >= fabs(lvec[1]) && fabs(lvec[0]) >= fabs(lvec[2])) {
    b0[0] = 0;
  } else if (fabs(lvec[1]) >= fabs(lvec[2])) {
    b0[1] = 0;
  } else {
    b0[2] = 0;
  }
  const double s = -dot3(lvec, b0) / dot3(lvec, lvec);
  b1[0] = b0[0] + lvec[0] * s;
  b1[1] = b0[1] + lvec[1] * s;
  b1[2] = b0[2] + lvec[2] * s;
  normalize3(b1);
  b0[0] = b1[1] * lvec[2] - b1[2] * lvec[1];
  b0[1] = b1[2] * lvec[0] - b1[0] * lvec[2];
  b0[2] = b1[0] * lvec[1] - b1[1] * lvec[0];
  normalize3(b0);
}

B2_HD double ray_triangle(const float* fv0, const float* fv1, const float* fv2,
                          const double lpnt[3], const double lvec[3], const double b0[3],
                          const double b1[3]) {
  const double v0[3] = {(double)fv0[0], (double)fv0[1], (double)fv0[2]};
  const double v1[3] = {(double)fv1[0], (double)fv1[1], (double)fv1[2]};
  const double v2[3] = {(double)fv2[0], (double)fv2[1], (double)fv2[2]};
  const double d0[3] = {v0[0] - lpnt[0], v0[1] - lpnt[1], v0[2] - lpnt[2]};
  const double d1[3] = {v1[0] - lpnt[0], v1[1] - lpnt[1], v1[2] - lpnt[2]};
  const double d2[3] = {v2[0] - lpnt[0], v2[1] - lpnt[1], v2[2] - lpnt[2]};
  const double p00 = dot3(b0, d0), p01 = dot3(b1, d0);
  const double p10 = dot3(b0, d1), p11 = dot3(b1, d1);
  const double p20 = dot3(b0, d2), p21 = dot3(b1, d2);
  if ((p00 > 0 && p10 > 0 && p20 > 0) || (p00 < 0 && p10 < 0 && p20 < 0) ||
      (p01 > 0 && p11 > 0 && p21 > 0) || (p01 < 0 && p11 < 0 && p21 < 0))
    return -1;
  const double A0 = p00 - p20, A1 = p10 - p20, A2 = p01 - p21, A3 = p11 - p21;
  const double bb0 = -p20, bb1 = -p21;
  const double det = A0 * A3 - A1 * A2;
  if (fabs(det) < kMinVal) return -1;
  const double t0 = (A3 * bb0 - A1 * bb1) / det;
  const double t1 = (-A2 * bb0 + A0 * bb1) / det;
  if (t0 < 0 || t1 < 0 || t0 + t1 > 1) return -1;
  const double e0[3] = {v0[0] - v2[0], v0[1] - v2[1], v0[2] - v2[2]};
  const double e1[3] = {v1[0] - v2[0], v1[1] - v2[1], v1[2] - v2[2]};
  const double e2[3] = {lpnt[0] - v2[0], lpnt[1] - v2[1], lpnt[2] - v2[2]};
  const double nrm[3] = {e0[1] * e1[2] - e0[2] * e1[1], e0[2] * e1[0] - e0[0] * e1[2],
                         e0[0] * e1[1] - e0[1] * e1[0]};
  const double denom = dot3(lvec, nrm);
  if (fabs(denom) < kMinVal) return -1;
  return -dot3(e2, nrm) / denom;
}

// ---------------------------------------------------------------------------------------------
// conservative fp32 image of an fp64 ray for slab culling, in FMA form:
//   t_near_k = pn_k * an_k - bn_k     (never above the true entry parameter)
//   t_far_k  = pf_k * af_k - bf_k     (never below the true exit parameter)
// an/af are 1/v shrunk/grown by 8 ulp (covers the float casts, the reciprocal and the final
// rounding of the fma), bn/bf are (origin +- 2 ulp) * an/af with directed rounding, so the only
// cancellation (pn*an against bn) happens inside one exactly-rounded fma.
// ---------------------------------------------------------------------------------------------
struct RayF {
  float an[3], bn[3], af[3], bf[3];
  bool neg[3];
};

B2_HD void rayf_setup(RayF& r, const double o[3], const double v[3]) {
  const float kEps = 4.76837158e-07f;  // 8 * 2^-24
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float of = (float)o[k];
    const float eo = fabsf(of) * 2.4e-7f + 1e-37f;  // >= 2 ulp, while |o - of| <= 0.5 ulp
    const float ohi = of + eo, olo = of - eo;
    const float vf = (float)v[k];
    const bool ng = v[k] < 0;
    const float av = fmaxf(fabsf(vf), 1e-36f);
    float inv = 1.0f / av;
    if (ng) inv = -inv;
    r.neg[k] = ng;
    r.an[k] = inv * (1.0f - kEps);
    r.af[k] = inv * (1.0f + kEps);
    // near plane uses the origin pushed along the ray, far plane the origin pulled back
    r.bn[k] = fmul_up(ng ? olo : ohi, r.an[k]);
    r.bf[k] = fmul_down(ng ? ohi : olo, r.af[k]);
  }
}

// true if the ray may touch the box for some t in [0, tmax]; tnear = conservative entry
B2_HD bool slab_test(const RayF& r, const float lo[3], const float hi[3], float tmax,
                     float& tnear) {
  float tn = 0.f, tf = tmax;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float pn = r.neg[k] ? hi[k] : lo[k];
    const float pf = r.neg[k] ? lo[k] : hi[k];
    tn = fmaxf(tn, fma_rn(pn, r.an[k], -r.bn[k]));
    tf = fminf(tf, fma_rn(pf, r.af[k], -r.bf[k]));
  }
  tnear = tn;
  return tn <= tf;
}

// ---------------------------------------------------------------------------------------------
// world-space bounding box of one geom (K1 leaf boxes), rounded outward to fp32
// ---------------------------------------------------------------------------------------------
B2_HD void geom_aabb(int type, const double size[3], const GeomPose& P, float lo[3], float hi[3]) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double m0 = P.mat[3 * k], m1 = P.mat[3 * k + 1], m2 = P.mat[3 * k + 2];
    double e;
    switch (type) {
      case GEOM_SPHERE:
        e = size[0];
        break;
      case GEOM_CAPSULE:
        e = fabs(m2) * size[1] + size[0];
        break;
      case GEOM_CYLINDER: {
        const double s = 1.0 - m2 * m2;
        e = fabs(m2) * size[1] + size[0] * sqrt(s > 0 ? s : 0.0);
        break;
      }
      case GEOM_ELLIPSOID: {
        const double a = m0 * size[0], b = m1 * size[1], c = m2 * size[2];
        e = sqrt(a * a + b * b + c * c);
        break;
      }
      default:  // box, mesh (geom_size box)
        e = fabs(m0) * size[0] + fabs(m1) * size[1] + fabs(m2) * size[2];
        break;
    }
    e = e * (1.0 + 1e-6) + 1e-9;
    lo[k] = f32_down(P.pos[k] - e);
    hi[k] = f32_up(P.pos[k] + e);
  }
}

// ---------------------------------------------------------------------------------------------
// closest-hit bookkeeping: A.5 update rule made order-independent (lowest geom id wins ties)
// ---------------------------------------------------------------------------------------------
struct Hit {
  double dist;  // < 0: none
  int32_t id;
};

B2_HD void hit_update(Hit& h, double nd, int32_t g) {
  if (nd >= 0 && (h.dist < 0 || nd < h.dist || (nd == h.dist && g < h.id))) {
    h.dist = nd;
    h.id = g;
  }
}

B2_HD float hit_tmax(const Hit& h, float tcap) {
  return h.dist < 0 ? tcap : fminf(tcap, f32_up(h.dist));
}

// ---------------------------------------------------------------------------------------------
// per-lane traversal of one mesh BVH (mj_rayMesh).  No warp-level operations: lanes call this
// for different meshes at the same time.  Out of line on the device.
// ---------------------------------------------------------------------------------------------
B2_LEAF double mesh_trace(const MeshInfo* mesh, const MeshNode4* mnodes, const MeshTri* mtris,
                          const GeomPose* gp, const GeomStatic* gsp, double px, double py,
                          double pz, double vx, double vy, double vz, float tcap) {
  const GeomPose P = ld_pose(gp);
  const double pnt[3] = {px, py, pz}, vec[3] = {vx, vy, vz};
  double lpnt[3], lvec[3], b0[3], b1[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  mesh_ray_basis(lvec, b0, b1);
  RayF rf;
  rayf_setup(rf, lpnt, lvec);
  const MeshInfo mi = mesh[gsp->dataid];
  const MeshNode4* nodes = mnodes + mi.node_base;
  const MeshTri* tris = mtris + mi.tri_base;

  double x = -1;
  float tmax = tcap;
  int32_t stack[kMeshStack];
  int sp = 0;
  int32_t cur = 0;
  while (true) {
    const MeshNode4 n = ld_node4(nodes + cur);
    // hit leaves are tested at once; of the hit inner children the nearest is entered next and
    // the others wait on the stack
    int32_t next = kEmptyChild;
    float tnext = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
      const int32_t link = n.child[c];
      if (link == kEmptyChild) continue;
      const float lo[3] = {n.lox[c], n.loy[c], n.loz[c]}, hi[3] = {n.hix[c], n.hiy[c], n.hiz[c]};
      float t;
      if (!slab_test(rf, lo, hi, tmax, t)) continue;
      if (link < 0) {
        const int32_t payload = ~link;
        const int first = payload >> 2, cnt = (payload & 3) + 1;
#pragma unroll 1
        for (int i = 0; i < cnt; i++) {
          const MeshTri tr = ld_tri(tris + first + i);
          const double sol = ray_triangle(tr.v0, tr.v1, tr.v2, lpnt, lvec, b0, b1);
          if (sol >= 0 && (x < 0 || sol < x)) {
            x = sol;
            tmax = fminf(tmax, f32_up(sol));
          }
        }
      } else if (next == kEmptyChild) {
        next = link;
        tnext = t;
      } else {
        const bool nearer = t < tnext;
        if (sp < kMeshStack) stack[sp++] = nearer ? next : link;
        if (nearer) {
          next = link;
          tnext = t;
        }
      }
    }
    if (next == kEmptyChild) {
      if (sp == 0) break;
      next = stack[--sp];
    }
    cur = next;
  }
  // mj_rayMesh's box pre-test on geom_size, applied after the fact (same result, paid on hits only)
  if (x >= 0 && prim_test(GEOM_BOX, gp, gsp, px, py, pz, vx, vy, vz) < 0) x = -1;
  return x;
}

// true if the ray certainly meets no triangle of mesh geom g: it misses the bounding sphere of the
// mesh (MeshInfo::bs_*, inflated by 1e-6 -- nine orders of magnitude above the rounding of the fp64
// triangle test, so even a hit that test would report through rounding lies inside it).  A cheaper
// and, for roundish meshes, tighter filter than mj_rayMesh's own geom_size box; mj_rayMesh's result
// for such a ray is -1 either way.
B2_HD bool mesh_bsphere_miss(const CastParams& p, const GeomPose* gp, const GeomStatic* gsp,
                             const double pnt[3], const double vec[3]) {
#ifdef B2_NO_MESH_SPHERE
  (void)p, (void)gp, (void)gsp, (void)pnt, (void)vec;
  return false;
#else
  const MeshInfo* mi = p.mesh + gsp->dataid;
#if B2_ON_DEVICE
  const float4 b = __ldg(reinterpret_cast<const float4*>(mi->bs_c));
  const double cx = b.x, cy = b.y, cz = b.z, r = b.w;
#else
  const double cx = mi->bs_c[0], cy = mi->bs_c[1], cz = mi->bs_c[2], r = mi->bs_r;
#endif
  const GeomPose P = ld_pose(gp);
  const double cw[3] = {P.pos[0] + (P.mat[0] * cx + P.mat[1] * cy + P.mat[2] * cz),
                        P.pos[1] + (P.mat[3] * cx + P.mat[4] * cy + P.mat[5] * cz),
                        P.pos[2] + (P.mat[6] * cx + P.mat[7] * cy + P.mat[8] * cz)};
  return sphere_miss(cw, r * r, pnt, vec);
#endif
}

// One geom leaf some lane's ray may touch (warp-uniform g); h = this lane's slab test passed.
// Primitives are tested here, all lanes on the same geom type.  Mesh geoms are handed to the warp
// policy W: the device kernel appends a (ray, mesh geom) PAIR to a compacted global list that the
// mesh kernels trace afterwards with full warps (mj_rayMesh per pair); the policy of the small /
// clean-up launches and of the CPU emulation traces the mesh in place.
template <class W>
B2_HD void leaf_visit(const CastParams& p, const GeomPose* pose, int g, int type, bool h, Hit& hit,
                      const double pnt[3], const double vec[3], float tcap, W& warp) {
  if (type == GEOM_MESH) {
    warp.mesh_pair(p, pose, g, h, hit, pnt, vec, tcap);
  } else if (h) {
    hit_update(hit,
               prim_test(type, pose + g, p.gstat + g, pnt[0], pnt[1], pnt[2], vec[0], vec[1],
                         vec[2]),
               g);
  }
}

// in-place mesh handling (no pair list): used by policies that do not split
B2_HD void mesh_inline(const CastParams& p, const GeomPose* pose, int g, bool h, Hit& hit,
                       const double pnt[3], const double vec[3], float tcap) {
  if (h && !mesh_bsphere_miss(p, pose + g, p.gstat + g, pnt, vec))
    hit_update(hit,
               mesh_trace(p.mesh, p.mnodes, p.mtris, pose + g, p.gstat + g, pnt[0], pnt[1], pnt[2],
                          vec[0], vec[1], vec[2], hit_tmax(hit, tcap)),
               g);
}

// ---------------------------------------------------------------------------------------------
// ray generation and the fused epilogues of the three cast modes (shared by the kernel and the
// CPU emulation harness)
// ---------------------------------------------------------------------------------------------
// R: the scan's sensor frame.  Returns false for a degenerate (zero-length) ray.
B2_HD bool cast_make_ray(const CastParams& p, int mode, const double R[9], int r, double l[3],
                         double vec[3]) {
  if (mode == MODE_CAMERA) {
    // MuJoCo camera frame: looks along -Z, +X right, +Y up; pixel centres, top-left origin
    const int v = r / p.cam_w, u = r - v * p.cam_w;
    l[0] = ((double)u + 0.5 - p.cam_cx) / p.cam_f;
    l[1] = -((double)v + 0.5 - p.cam_cy) / p.cam_f;
    l[2] = -1.0;
  } else {
    l[0] = p.local_dir[(size_t)r * 3 + 0];
    l[1] = p.local_dir[(size_t)r * 3 + 1];
    l[2] = p.local_dir[(size_t)r * 3 + 2];
  }
  if (mode == MODE_MULTIRAY) {
    vec[0] = l[0];
    vec[1] = l[1];
    vec[2] = l[2];
    return dot3(vec, vec) >= kMinVal;
  }
  // mju_mulMatVec3 order
  vec[0] = R[0] * l[0] + R[1] * l[1] + R[2] * l[2];
  vec[1] = R[3] * l[0] + R[4] * l[1] + R[5] * l[2];
  vec[2] = R[6] * l[0] + R[7] * l[1] + R[8] * l[2];
  return dot3(vec, vec) >= kMinVal * kMinVal;
}

B2_HD float quiet_nan() {
#if B2_ON_DEVICE
  return __int_as_float(0x7fc00000);
#else
  return NAN;
#endif
}

// engine_sensor.c apply_cutoff, then ref:src/mujoco_lidar.cpp:262 (double -> float) and :269-271
// (filter on the float against the double limits).  val = the mj_ray value after the cutoff.
B2_HD float lidar_range(const CastParams& p, double dist, double& val, bool& filtered) {
  val = dist;
  if (p.cutoff > 0 && val > p.cutoff) val = p.cutoff;
  float rng = (float)val;
  filtered = ((double)rng < p.range_min) || ((double)rng > p.range_max);
  if (filtered) rng = -1.0f;
  return rng;
}

B2_HD void cast_epilogue(const CastParams& p, int mode, int scan, int r, const Hit& hit,
                         const double l[3]) {
  const size_t idx = (size_t)scan * p.nray + r;
  if (mode == MODE_LIDAR) {
    double val;
    bool filtered;
    const float rng = lidar_range(p, hit.dist, val, filtered);
    if (p.ranges) p.ranges[idx] = rng;
    if (p.geomid) p.geomid[idx] = hit.id;
    if (p.geomid16) p.geomid16[idx] = (uint16_t)hit.id;  // -1 -> 0xFFFF
    if (p.dist) p.dist[idx] = val;
    if (p.xyz) {
      const bool ok = hit.id >= 0 && !filtered;
      const float qn = quiet_nan();
      p.xyz[idx * 3 + 0] = ok ? (float)(l[0] * val) : qn;
      p.xyz[idx * 3 + 1] = ok ? (float)(l[1] * val) : qn;
      p.xyz[idx * 3 + 2] = ok ? (float)(l[2] * val) : qn;
    }
  } else if (mode == MODE_MULTIRAY) {
    double d = hit.dist;
    int32_t id = hit.id;
    if (d > p.cutoff) {
      d = -1;
      id = -1;
    }
    if (p.dist) p.dist[idx] = d;
    if (p.geomid) p.geomid[idx] = id;
  } else {
    // GL window depth of the hit: d = (1 - near/z) / (1 - near/far), z = distance along -Z;
    // GL rows are bottom-up
    const int v = r / p.cam_w, u = r - v * p.cam_w;
    float d = 1.0f;
    if (hit.id >= 0) {
      const float z = (float)hit.dist;
      d = (1.0f - p.cam_near / z) / (1.0f - p.cam_near / p.cam_far);
      d = fminf(fmaxf(d, 0.0f), 1.0f);
    }
    p.gl_depth[(size_t)scan * p.nray + (size_t)(p.cam_h - 1 - v) * p.cam_w + u] = d;
  }
}

}  // namespace b2
#include "tile_trace.cuh"

#endif
