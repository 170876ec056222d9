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
// conservative culling (slab tests against outward-rounded boxes) and can never drop a hit the
// double test would accept.
#ifndef B2_RAY_CORE_CUH
#define B2_RAY_CORE_CUH

#include <float.h>
#include <math.h>

#include "b2_types.h"

namespace b2 {

#if defined(__CUDA_ARCH__)
#define B2_ON_DEVICE 1
#else
#define B2_ON_DEVICE 0
#endif

constexpr int kPacketStack = 64;  // geom-BVH stack entries per warp
constexpr int kMeshStack = 64;    // mesh-BVH stack entries per lane
constexpr int kMaxMeshCand = 12;  // deferred mesh candidates per lane

// ---------------------------------------------------------------------------------------------
// directed rounding double -> float
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

B2_HD double ray_sphere(const double pos[3], double dist_sqr, const double pnt[3],
                        const double vec[3]) {
  const double d0 = pnt[0] - pos[0], d1 = pnt[1] - pos[1], d2 = pnt[2] - pos[2];
  const double a = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  const double b = vec[0] * d0 + vec[1] * d1 + vec[2] * d2;
  const double c = d0 * d0 + d1 * d1 + d2 * d2 - dist_sqr;
  double x0, x1;
  return ray_quad(a, b, c, x0, x1);
}

B2_HD double ray_plane(const GeomPose& P, const double size[3], const double pnt[3],
                       const double vec[3]) {
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  if (lvec[2] > -kMinVal) return -1;
  const double x = -lpnt[2] / lvec[2];
  if (x < 0) return -1;
  const double p0 = lpnt[0] + x * lvec[0];
  const double p1 = lpnt[1] + x * lvec[1];
  if ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) return x;
  return -1;
}

B2_HD void keep_min(double& x, double sol) {
  if (x < 0 || sol < x) x = sol;
}

B2_HD double ray_capsule(const GeomPose& P, const double size[3], const double pnt[3],
                         const double vec[3]) {
  const double ssz = size[0] + size[1];
  if (ray_sphere(P.pos, ssz * ssz, pnt, vec) < 0) return -1;
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  double x = -1, x0, x1;
  const double sq_size0 = size[0] * size[0];
  // round side
  double a = lvec[0] * lvec[0] + lvec[1] * lvec[1];
  double b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1];
  double c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - sq_size0;
  const double sol = ray_quad(a, b, c, x0, x1);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) keep_min(x, sol);
  // top cap
  double ldif[3] = {lpnt[0], lpnt[1], lpnt[2] - size[1]};
  a += lvec[2] * lvec[2];
  b = dot3(lvec, ldif);
  c = dot3(ldif, ldif) - sq_size0;
  ray_quad(a, b, c, x0, x1);
  if (x0 >= 0 && lpnt[2] + x0 * lvec[2] >= size[1]) keep_min(x, x0);
  if (x1 >= 0 && lpnt[2] + x1 * lvec[2] >= size[1]) keep_min(x, x1);
  // bottom cap
  ldif[2] = lpnt[2] + size[1];
  b = dot3(lvec, ldif);
  c = dot3(ldif, ldif) - sq_size0;
  ray_quad(a, b, c, x0, x1);
  if (x0 >= 0 && lpnt[2] + x0 * lvec[2] <= -size[1]) keep_min(x, x0);
  if (x1 >= 0 && lpnt[2] + x1 * lvec[2] <= -size[1]) keep_min(x, x1);
  return x;
}

B2_HD double ray_ellipsoid(const GeomPose& P, const double size[3], const double pnt[3],
                           const double vec[3]) {
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  const double s0 = 1 / (size[0] * size[0]), s1 = 1 / (size[1] * size[1]),
               s2 = 1 / (size[2] * size[2]);
  const double a = s0 * lvec[0] * lvec[0] + s1 * lvec[1] * lvec[1] + s2 * lvec[2] * lvec[2];
  const double b = s0 * lvec[0] * lpnt[0] + s1 * lvec[1] * lpnt[1] + s2 * lvec[2] * lpnt[2];
  const double c = s0 * lpnt[0] * lpnt[0] + s1 * lpnt[1] * lpnt[1] + s2 * lpnt[2] * lpnt[2] - 1;
  double x0, x1;
  return ray_quad(a, b, c, x0, x1);
}

B2_HD double ray_cylinder(const GeomPose& P, const double size[3], const double pnt[3],
                          const double vec[3]) {
  const double ssz = size[0] * size[0] + size[1] * size[1];
  if (ray_sphere(P.pos, ssz, pnt, vec) < 0) return -1;
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  double x = -1;
  if (fabs(lvec[2]) > kMinVal) {
#pragma unroll
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
  double x0, x1;
  const double sol = ray_quad(a, b, c, x0, x1);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) keep_min(x, sol);
  return x;
}

// one face pair of the box test: axis i, the other two axes j,k
#define B2_BOX_AXIS(i, j, k)                                                        \
  if (fabs(lvec[i]) > kMinVal) {                                                    \
    _Pragma("unroll") for (int side = -1; side <= 1; side += 2) {                   \
      const double sol = (side * size[i] - lpnt[i]) / lvec[i];                      \
      if (sol >= 0) {                                                               \
        const double p0 = lpnt[j] + sol * lvec[j];                                  \
        const double p1 = lpnt[k] + sol * lvec[k];                                  \
        if (fabs(p0) <= size[j] && fabs(p1) <= size[k]) keep_min(x, sol);           \
      }                                                                             \
    }                                                                               \
  }

B2_HD double ray_box(const GeomPose& P, const double size[3], const double pnt[3],
                     const double vec[3]) {
  const double ssz = size[0] * size[0] + size[1] * size[1] + size[2] * size[2];
  if (ray_sphere(P.pos, ssz, pnt, vec) < 0) return -1;
  double lpnt[3], lvec[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  double x = -1;
  B2_BOX_AXIS(0, 1, 2)
  B2_BOX_AXIS(1, 0, 2)
  B2_BOX_AXIS(2, 0, 1)
  return x;
}

// mju_rayGeom for the primitive types (mesh handled by mesh_trace)
B2_HD double ray_geom(int type, const GeomPose& P, const double size[3], const double pnt[3],
                      const double vec[3]) {
  switch (type) {
    case GEOM_PLANE:
      return ray_plane(P, size, pnt, vec);
    case GEOM_SPHERE:
      return ray_sphere(P.pos, size[0] * size[0], pnt, vec);
    case GEOM_CAPSULE:
      return ray_capsule(P, size, pnt, vec);
    case GEOM_ELLIPSOID:
      return ray_ellipsoid(P, size, pnt, vec);
    case GEOM_CYLINDER:
      return ray_cylinder(P, size, pnt, vec);
    case GEOM_BOX:
      return ray_box(P, size, pnt, vec);
    default:
      return -1;
  }
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
// conservative fp32 image of an fp64 ray, for slab culling
// ---------------------------------------------------------------------------------------------
struct RayF {
  float on[3], of[3];  // origin to use with the near / far plane of each axis
  float in[3], ifr[3]; // 1/v shrunk (near) and grown (far) by the accumulated rounding bound
  bool neg[3];
};

B2_HD void rayf_setup(RayF& r, const double o[3], const double v[3]) {
  const float kEps = 4.76837158e-07f;  // 8 * 2^-24: covers sub, rcp, cast, mul and scale roundings
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float of = (float)o[k];
    const float eo = fabsf(of) * 2.4e-7f + 1e-37f;  // >= 2 ulp: |o - of| <= 0.5 ulp
    const float ohi = of + eo, olo = of - eo;
    const float vf = (float)v[k];
    const bool ng = v[k] < 0;
    const float av = fmaxf(fabsf(vf), 1e-36f);
    float inv = 1.0f / av;
    if (ng) inv = -inv;
    r.neg[k] = ng;
    r.in[k] = inv * (1.0f - kEps);
    r.ifr[k] = inv * (1.0f + kEps);
    r.on[k] = ng ? olo : ohi;
    r.of[k] = ng ? ohi : olo;
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
    tn = fmaxf(tn, (pn - r.on[k]) * r.in[k]);
    tf = fminf(tf, (pf - r.of[k]) * r.ifr[k]);
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
// for different meshes at the same time.
// ---------------------------------------------------------------------------------------------
template <class Ld>
B2_HD double mesh_trace(const CastParams& p, int g, const GeomStatic& gs, const GeomPose& P,
                        const double pnt[3], const double vec[3], float tcap, const Ld& ld) {
  double lpnt[3], lvec[3], b0[3], b1[3];
  ray_map(P, pnt, vec, lpnt, lvec);
  mesh_ray_basis(lvec, b0, b1);
  RayF rf;
  rayf_setup(rf, lpnt, lvec);
  const MeshInfo mi = p.mesh[gs.dataid];
  const BvhNode* nodes = p.mnodes + mi.node_base;
  const MeshTri* tris = p.mtris + mi.tri_base;

  double x = -1;
  float tmax = tcap;
  int32_t stack[kMeshStack];
  int sp = 0;
  int32_t cur = 0;
  while (true) {
    const BvhNode n = ld.node(nodes + cur);
    float t0, t1;
    bool h0 = slab_test(rf, n.lo0, n.hi0, tmax, t0);
    bool h1 = slab_test(rf, n.lo1, n.hi1, tmax, t1);
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const bool h = c ? h1 : h0;
      const int32_t link = c ? n.c1 : n.c0;
      if (h && link < 0) {
        const int32_t payload = ~link;
        const int first = payload >> 2, cnt = (payload & 3) + 1;
        for (int i = 0; i < cnt; i++) {
          const MeshTri tr = ld.tri(tris + first + i);
          const double sol = ray_triangle(tr.v0, tr.v1, tr.v2, lpnt, lvec, b0, b1);
          if (sol >= 0 && (x < 0 || sol < x)) {
            x = sol;
            tmax = fminf(tmax, f32_up(sol));
          }
        }
      }
    }
    h0 = h0 && n.c0 >= 0;
    h1 = h1 && n.c1 >= 0;
    if (h0 && h1) {
      const bool first0 = t0 <= t1;
      if (sp < kMeshStack) stack[sp++] = first0 ? n.c1 : n.c0;
      cur = first0 ? n.c0 : n.c1;
    } else if (h0) {
      cur = n.c0;
    } else if (h1) {
      cur = n.c1;
    } else {
      if (sp == 0) break;
      cur = stack[--sp];
    }
  }
  // mj_rayMesh's box pre-test on geom_size, applied after the fact (same result, paid on hits only)
  if (x >= 0 && ray_box(P, gs.size, pnt, vec) < 0) x = -1;
  return x;
}

// ---------------------------------------------------------------------------------------------
// one ray of a warp packet.  W supplies the warp votes and the per-warp stack; every lane of the
// warp must call trace_ray together (lanes without a ray pass active = false).
// ---------------------------------------------------------------------------------------------
template <class W, class Ld>
B2_HD Hit trace_ray(const CastParams& p, const BvhNode* tlas, const GeomPose* pose, bool active,
                    const double pnt[3], const double vec[3], float tcap, W& warp, const Ld& ld) {
  Hit hit;
  hit.dist = -1;
  hit.id = -1;

  // planes live outside the BVH (unbounded)
  for (int i = 0; i < p.nplane; i++) {
    const int g = p.plane_ids[i];
    if (p.elim[g]) continue;
    if (active) {
      const GeomPose P = ld.pose(pose + g);
      const GeomStatic gs = ld.gstat(p.gstat + g);
      hit_update(hit, ray_geom(GEOM_PLANE, P, gs.size, pnt, vec), g);
    }
  }

  RayF rf;
  rayf_setup(rf, pnt, vec);

  // deferred mesh candidates of this lane
  float cand_t[kMaxMeshCand];
  int32_t cand_g[kMaxMeshCand];
  int ncand = 0;

  if (p.nnode > 0) {
    int sp = 0;
    int32_t cur = 0;
    while (true) {
      const BvhNode n = ld.tlas_node(tlas + cur);
      const float tmax = hit_tmax(hit, tcap);
      float t0, t1;
      const bool h0 = active && slab_test(rf, n.lo0, n.hi0, tmax, t0);
      const bool h1 = active && slab_test(rf, n.lo1, n.hi1, tmax, t1);
      unsigned m0 = warp.ballot(h0), m1 = warp.ballot(h1);
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const int32_t link = c ? n.c1 : n.c0;
        const unsigned m = c ? m1 : m0;
        if (m != 0u && link < 0) {  // warp-uniform: a geom leaf some lane touches
          const int g = ~link;
          if (!p.elim[g]) {
            const GeomStatic gs = ld.gstat(p.gstat + g);
            const bool h = c ? h1 : h0;
            if (h) {
              if (gs.type == GEOM_MESH) {
                if (ncand < kMaxMeshCand) {
                  cand_t[ncand] = c ? t1 : t0;
                  cand_g[ncand] = g;
                  ncand++;
                } else {  // candidate list full: resolve this one now
                  const GeomPose P = ld.pose(pose + g);
                  hit_update(hit, mesh_trace(p, g, gs, P, pnt, vec, hit_tmax(hit, tcap), ld), g);
                }
              } else {
                const GeomPose P = ld.pose(pose + g);
                hit_update(hit, ray_geom(gs.type, P, gs.size, pnt, vec), g);
              }
            }
          }
        }
      }
      if (n.c0 < 0) m0 = 0u;
      if (n.c1 < 0) m1 = 0u;
      if (m0 != 0u && m1 != 0u) {
        // both inner children wanted: descend into the one most lanes enter first
        const unsigned near0 = warp.ballot(h0 && (!h1 || t0 <= t1));
        const unsigned near1 = warp.ballot(h1 && (!h0 || t1 < t0));
        const bool first0 = warp.popc(near0) >= warp.popc(near1);
        warp.push(sp, first0 ? n.c1 : n.c0);
        cur = first0 ? n.c0 : n.c1;
      } else if (m0 != 0u) {
        cur = n.c0;
      } else if (m1 != 0u) {
        cur = n.c1;
      } else {
        if (sp == 0) break;
        cur = warp.pop(sp);
      }
    }
  }

  // deferred meshes, nearest box first, each lane on its own
  while (ncand > 0) {
    int best = 0;
    for (int i = 1; i < ncand; i++)
      if (cand_t[i] < cand_t[best]) best = i;
    const float t = cand_t[best];
    const int g = cand_g[best];
    ncand--;
    cand_t[best] = cand_t[ncand];
    cand_g[best] = cand_g[ncand];
    const float tmax = hit_tmax(hit, tcap);
    if (t > tmax) continue;
    const GeomStatic gs = ld.gstat(p.gstat + g);
    const GeomPose P = ld.pose(pose + g);
    hit_update(hit, mesh_trace(p, g, gs, P, pnt, vec, tmax, ld), g);
  }
  return hit;
}

// ---------------------------------------------------------------------------------------------
// ray generation and the fused epilogues of the three cast modes (shared by the kernel and the
// CPU emulation harness)
// ---------------------------------------------------------------------------------------------
// o / R: the scan's sensor frame.  Returns false for a degenerate (zero-length) ray.
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

B2_HD void cast_epilogue(const CastParams& p, int mode, int scan, int r, const Hit& hit,
                         const double l[3]) {
  const size_t idx = (size_t)scan * p.nray + r;
  if (mode == MODE_LIDAR) {
    // engine_sensor.c apply_cutoff, then ref:src/mujoco_lidar.cpp:262 (double -> float) and
    // :269-271 (filter on the float against the double limits)
    double val = hit.dist;
    if (p.cutoff > 0 && val > p.cutoff) val = p.cutoff;
    float rng = (float)val;
    const bool filtered = ((double)rng < p.range_min) || ((double)rng > p.range_max);
    if (filtered) rng = -1.0f;
    if (p.ranges) p.ranges[idx] = rng;
    if (p.geomid) p.geomid[idx] = hit.id;
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
#endif
