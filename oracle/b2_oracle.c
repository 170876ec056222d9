/*
 * b2_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See b2_oracle.h.
 *
 * Every function cites the upstream routine it restates (MuJoCo 3.3.4 engine_ray.c /
 * engine_sensor.c as specified in SURVEY.md Appendix A) or the reference file:line it follows.
 * Compile with -ffp-contract=off: parity with the CUDA path is checked bit for bit, so no
 * operation here may be fused into an FMA.
 */
#include "b2_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * small vector helpers (mju_dot3 / mju_cross / mju_normalize3 operation order)
 * ---------------------------------------------------------------------------------------- */
static inline double dot3(const double a[3], const double b[3]) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

static inline void cross3(double r[3], const double a[3], const double b[3]) {
  r[0] = a[1] * b[2] - a[2] * b[1];
  r[1] = a[2] * b[0] - a[0] * b[2];
  r[2] = a[0] * b[1] - a[1] * b[0];
}

static inline void normalize3(double v[3]) {
  double n = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (n < ORC_MINVAL) {
    v[0] = 1;
    v[1] = 0;
    v[2] = 0;
  } else {
    double inv = 1 / n;
    v[0] *= inv;
    v[1] *= inv;
    v[2] *= inv;
  }
}

/* A.3 ray_map: lpnt = mat' * (pnt - pos), lvec = mat' * vec */
static void ray_map(const double pos[3], const double mat[9], const double pnt[3],
                    const double vec[3], double lpnt[3], double lvec[3]) {
  const double dif[3] = {pnt[0] - pos[0], pnt[1] - pos[1], pnt[2] - pos[2]};
  lpnt[0] = mat[0] * dif[0] + mat[3] * dif[1] + mat[6] * dif[2];
  lpnt[1] = mat[1] * dif[0] + mat[4] * dif[1] + mat[7] * dif[2];
  lpnt[2] = mat[2] * dif[0] + mat[5] * dif[1] + mat[8] * dif[2];
  lvec[0] = mat[0] * vec[0] + mat[3] * vec[1] + mat[6] * vec[2];
  lvec[1] = mat[1] * vec[0] + mat[4] * vec[1] + mat[7] * vec[2];
  lvec[2] = mat[2] * vec[0] + mat[5] * vec[1] + mat[8] * vec[2];
}

/* A.3 ray_quad: a*x^2 + 2*b*x + c = 0; det < mjMINVAL is a miss */
static double ray_quad(double a, double b, double c, double x[2]) {
  double det = b * b - a * c;
  if (det < ORC_MINVAL) {
    x[0] = -1;
    x[1] = -1;
    return -1;
  }
  det = sqrt(det);
  x[0] = (-b - det) / a;
  x[1] = (-b + det) / a;
  if (x[0] >= 0) return x[0];
  if (x[1] >= 0) return x[1];
  return -1;
}

/* A.3 plane: one-sided, local z = 0 */
static double ray_plane(const double pos[3], const double mat[9], const double size[3],
                        const double pnt[3], const double vec[3]) {
  double lpnt[3], lvec[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);
  if (lvec[2] > -ORC_MINVAL) return -1;
  const double x = -lpnt[2] / lvec[2];
  if (x < 0) return -1;
  const double p0 = lpnt[0] + x * lvec[0];
  const double p1 = lpnt[1] + x * lvec[1];
  if ((size[0] <= 0 || fabs(p0) <= size[0]) && (size[1] <= 0 || fabs(p1) <= size[1])) return x;
  return -1;
}

/* A.3 sphere: world frame, dist_sqr = r^2 */
static double ray_sphere(const double pos[3], double dist_sqr, const double pnt[3],
                         const double vec[3]) {
  const double dif[3] = {pnt[0] - pos[0], pnt[1] - pos[1], pnt[2] - pos[2]};
  const double a = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2];
  const double b = vec[0] * dif[0] + vec[1] * dif[1] + vec[2] * dif[2];
  const double c = dif[0] * dif[0] + dif[1] * dif[1] + dif[2] * dif[2] - dist_sqr;
  double xx[2];
  return ray_quad(a, b, c, xx);
}

/* A.3 capsule: radius size[0], half-length size[1] along local z */
static double ray_capsule(const double pos[3], const double mat[9], const double size[3],
                          const double pnt[3], const double vec[3]) {
  const double ssz = size[0] + size[1];
  if (ray_sphere(pos, ssz * ssz, pnt, vec) < 0) return -1;

  double lpnt[3], lvec[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);

  double x = -1, sol, xx[2];
  const double sq_size0 = size[0] * size[0];

  /* round side */
  double a = lvec[0] * lvec[0] + lvec[1] * lvec[1];
  double b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1];
  double c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - sq_size0;
  sol = ray_quad(a, b, c, xx);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) {
    if (x < 0 || sol < x) x = sol;
  }

  /* top cap */
  double ldif[3] = {lpnt[0], lpnt[1], lpnt[2] - size[1]};
  a += lvec[2] * lvec[2];
  b = dot3(lvec, ldif);
  c = dot3(ldif, ldif) - sq_size0;
  ray_quad(a, b, c, xx);
  for (int i = 0; i < 2; i++) {
    if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] >= size[1]) {
      if (x < 0 || xx[i] < x) x = xx[i];
    }
  }

  /* bottom cap */
  ldif[2] = lpnt[2] + size[1];
  b = dot3(lvec, ldif);
  c = dot3(ldif, ldif) - sq_size0;
  ray_quad(a, b, c, xx);
  for (int i = 0; i < 2; i++) {
    if (xx[i] >= 0 && lpnt[2] + xx[i] * lvec[2] <= -size[1]) {
      if (x < 0 || xx[i] < x) x = xx[i];
    }
  }
  return x;
}

/* A.3 ellipsoid */
static double ray_ellipsoid(const double pos[3], const double mat[9], const double size[3],
                            const double pnt[3], const double vec[3]) {
  double lpnt[3], lvec[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);
  const double s[3] = {1 / (size[0] * size[0]), 1 / (size[1] * size[1]), 1 / (size[2] * size[2])};
  const double a = s[0] * lvec[0] * lvec[0] + s[1] * lvec[1] * lvec[1] + s[2] * lvec[2] * lvec[2];
  const double b = s[0] * lvec[0] * lpnt[0] + s[1] * lvec[1] * lpnt[1] + s[2] * lvec[2] * lpnt[2];
  const double c =
      s[0] * lpnt[0] * lpnt[0] + s[1] * lpnt[1] * lpnt[1] + s[2] * lpnt[2] * lpnt[2] - 1;
  double xx[2];
  return ray_quad(a, b, c, xx);
}

/* A.3 cylinder: radius size[0], half-height size[1] */
static double ray_cylinder(const double pos[3], const double mat[9], const double size[3],
                           const double pnt[3], const double vec[3]) {
  const double ssz = size[0] * size[0] + size[1] * size[1];
  if (ray_sphere(pos, ssz, pnt, vec) < 0) return -1;

  double lpnt[3], lvec[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);

  double x = -1, sol;

  /* flat sides */
  if (fabs(lvec[2]) > ORC_MINVAL) {
    for (int side = -1; side <= 1; side += 2) {
      sol = (side * size[1] - lpnt[2]) / lvec[2];
      if (sol >= 0) {
        const double p0 = lpnt[0] + sol * lvec[0];
        const double p1 = lpnt[1] + sol * lvec[1];
        if (p0 * p0 + p1 * p1 <= size[0] * size[0]) {
          if (x < 0 || sol < x) x = sol;
        }
      }
    }
  }

  /* round side */
  const double a = lvec[0] * lvec[0] + lvec[1] * lvec[1];
  const double b = lvec[0] * lpnt[0] + lvec[1] * lpnt[1];
  const double c = lpnt[0] * lpnt[0] + lpnt[1] * lpnt[1] - size[0] * size[0];
  double xx[2];
  sol = ray_quad(a, b, c, xx);
  if (sol >= 0 && fabs(lpnt[2] + sol * lvec[2]) <= size[1]) {
    if (x < 0 || sol < x) x = sol;
  }
  return x;
}

/* A.3 box: half-sizes size */
static double ray_box(const double pos[3], const double mat[9], const double size[3],
                      const double pnt[3], const double vec[3]) {
  static const int ifcs[3][2] = {{1, 2}, {0, 2}, {0, 1}};
  const double ssz = size[0] * size[0] + size[1] * size[1] + size[2] * size[2];
  if (ray_sphere(pos, ssz, pnt, vec) < 0) return -1;

  double lpnt[3], lvec[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);

  double x = -1, sol;
  for (int i = 0; i < 3; i++) {
    if (fabs(lvec[i]) > ORC_MINVAL) {
      for (int side = -1; side <= 1; side += 2) {
        sol = (side * size[i] - lpnt[i]) / lvec[i];
        if (sol >= 0) {
          const double p0 = lpnt[ifcs[i][0]] + sol * lvec[ifcs[i][0]];
          const double p1 = lpnt[ifcs[i][1]] + sol * lvec[ifcs[i][1]];
          if (fabs(p0) <= size[ifcs[i][0]] && fabs(p1) <= size[ifcs[i][1]]) {
            if (x < 0 || sol < x) x = sol;
          }
        }
      }
    }
  }
  return x;
}

/* A.3 mju_rayGeom dispatcher.  HFIELD / SDF / MESH are not handled here (mesh: orc_ray_mesh). */
double orc_ray_geom(const double pos[3], const double mat[9], const double size[3],
                    const double pnt[3], const double vec[3], int geomtype) {
  switch (geomtype) {
    case ORC_GEOM_PLANE:
      return ray_plane(pos, mat, size, pnt, vec);
    case ORC_GEOM_SPHERE:
      return ray_sphere(pos, size[0] * size[0], pnt, vec);
    case ORC_GEOM_CAPSULE:
      return ray_capsule(pos, mat, size, pnt, vec);
    case ORC_GEOM_ELLIPSOID:
      return ray_ellipsoid(pos, mat, size, pnt, vec);
    case ORC_GEOM_CYLINDER:
      return ray_cylinder(pos, mat, size, pnt, vec);
    case ORC_GEOM_BOX:
      return ray_box(pos, mat, size, pnt, vec);
    default:
      return -1;
  }
}

void orc_ray_geom_batch(int n, const double* pos, const double* mat, const double* size,
                        const double* pnt, const double* vec, const int32_t* geomtype, double* out) {
  for (int i = 0; i < n; i++)
    out[i] = orc_ray_geom(pos + 3 * (size_t)i, mat + 9 * (size_t)i, size + 3 * (size_t)i,
                          pnt + 3 * (size_t)i, vec + 3 * (size_t)i, geomtype[i]);
}

/* ------------------------------------------------------------------------------------------
 * A.4 mesh
 * ---------------------------------------------------------------------------------------- */

/* ray_triangle: v widened float->double by the caller */
static double ray_triangle(double v[3][3], const double lpnt[3], const double lvec[3],
                           const double b0[3], const double b1[3]) {
  double dif[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) dif[i][j] = v[i][j] - lpnt[j];

  double planar[3][2];
  for (int i = 0; i < 3; i++) {
    planar[i][0] = dot3(b0, dif[i]);
    planar[i][1] = dot3(b1, dif[i]);
  }

  if ((planar[0][0] > 0 && planar[1][0] > 0 && planar[2][0] > 0) ||
      (planar[0][0] < 0 && planar[1][0] < 0 && planar[2][0] < 0) ||
      (planar[0][1] > 0 && planar[1][1] > 0 && planar[2][1] > 0) ||
      (planar[0][1] < 0 && planar[1][1] < 0 && planar[2][1] < 0)) {
    return -1;
  }

  const double A[4] = {planar[0][0] - planar[2][0], planar[1][0] - planar[2][0],
                       planar[0][1] - planar[2][1], planar[1][1] - planar[2][1]};
  const double b[2] = {-planar[2][0], -planar[2][1]};
  const double det = A[0] * A[3] - A[1] * A[2];
  if (fabs(det) < ORC_MINVAL) return -1;
  const double t0 = (A[3] * b[0] - A[1] * b[1]) / det;
  const double t1 = (-A[2] * b[0] + A[0] * b[1]) / det;
  if (t0 < 0 || t1 < 0 || t0 + t1 > 1) return -1;

  double e0[3], e1[3], e2[3], nrm[3];
  for (int j = 0; j < 3; j++) {
    e0[j] = v[0][j] - v[2][j];
    e1[j] = v[1][j] - v[2][j];
    e2[j] = lpnt[j] - v[2][j];
  }
  cross3(nrm, e0, e1);
  const double denom = dot3(lvec, nrm);
  if (fabs(denom) < ORC_MINVAL) return -1;
  return -dot3(e2, nrm) / denom;
}

/* mju_raySlab restated for an AABB already in the ray's (mesh-local) frame */
static int ray_slab(const double aabb[6], const double src[3], const double dir[3]) {
  double tmin = 0.0, tmax = INFINITY;
  const double invdir[3] = {1.0 / dir[0], 1.0 / dir[1], 1.0 / dir[2]};
  for (int k = 0; k < 3; k++) {
    const double lo = aabb[k] - aabb[k + 3], hi = aabb[k] + aabb[k + 3];
    const double t1 = (lo - src[k]) * invdir[k];
    const double t2 = (hi - src[k]) * invdir[k];
    const double mn = t1 < t2 ? t1 : t2;
    const double mx = t1 < t2 ? t2 : t1;
    tmin = tmin > mn ? tmin : mn;
    tmax = tmax < mx ? tmax : mx;
  }
  return tmin < tmax;
}

/* plane basis used by every triangle of one (ray, mesh geom) pair */
static void mesh_ray_basis(const double lvec[3], double b0[3], double b1[3]) {
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
  cross3(b0, b1, lvec);
  normalize3(b0);
}

static double mesh_face_test(const orc_model* m, int meshid, int face, const double lpnt[3],
                             const double lvec[3], const double b0[3], const double b1[3]) {
  const float* vert = m->mesh_vert + 3 * (size_t)m->mesh_vertadr[meshid];
  const int32_t* f = m->mesh_face + 3 * ((size_t)m->mesh_faceadr[meshid] + face);
  double v[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) v[i][j] = (double)vert[3 * f[i] + j];
  return ray_triangle(v, lpnt, lvec, b0, b1);
}

double orc_ray_mesh(const orc_model* m, const orc_data* d, int g, const double pnt[3],
                    const double vec[3]) {
  const double* pos = d->geom_xpos + 3 * g;
  const double* mat = d->geom_xmat + 9 * g;
  /* bounding-box pre-test on the geom's box (geom_size) */
  if (ray_box(pos, mat, m->geom_size + 3 * g, pnt, vec) < 0) return -1;

  const int meshid = m->geom_dataid[g];
  double lpnt[3], lvec[3], b0[3], b1[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);
  mesh_ray_basis(lvec, b0, b1);

  const int adr = m->mesh_bvhadr[meshid];
  double x = -1;
  int stack[128], nstack = 0;
  stack[nstack++] = 0;
  while (nstack) {
    const int node = stack[--nstack];
    if (!ray_slab(m->bvh_aabb + 6 * (size_t)(adr + node), lpnt, lvec)) continue;
    const int face = m->bvh_nodeid[adr + node];
    if (face != -1) {
      const double sol = mesh_face_test(m, meshid, face, lpnt, lvec, b0, b1);
      if (sol >= 0 && (x < 0 || sol < x)) x = sol;
      continue;
    }
    for (int i = 0; i < 2; i++) {
      const int c = m->bvh_child[2 * (adr + node) + i];
      if (c != -1 && nstack < 128) stack[nstack++] = c;
    }
  }
  return x;
}

double orc_ray_mesh_brute(const orc_model* m, const orc_data* d, int g, const double pnt[3],
                          const double vec[3]) {
  const double* pos = d->geom_xpos + 3 * g;
  const double* mat = d->geom_xmat + 9 * g;
  if (ray_box(pos, mat, m->geom_size + 3 * g, pnt, vec) < 0) return -1;
  const int meshid = m->geom_dataid[g];
  double lpnt[3], lvec[3], b0[3], b1[3];
  ray_map(pos, mat, pnt, vec, lpnt, lvec);
  mesh_ray_basis(lvec, b0, b1);
  double x = -1;
  for (int f = 0; f < m->mesh_facenum[meshid]; f++) {
    const double sol = mesh_face_test(m, meshid, f, lpnt, lvec, b0, b1);
    if (sol >= 0 && (x < 0 || sol < x)) x = sol;
  }
  return x;
}

/* ---- mesh BVH construction: top-down, longest-axis median split, one face per leaf -------- */
typedef struct {
  const float* vert;
  const int32_t* face;
  double* aabb;
  int32_t* child;
  int32_t* nodeid;
  int nnode;
  int axis; /* comparator state */
  const double* cen;
} bvh_builder;

static const double* g_sort_cen;
static int g_sort_axis;
static int cmp_face(const void* a, const void* b) {
  const double ca = g_sort_cen[3 * (*(const int*)a) + g_sort_axis];
  const double cb = g_sort_cen[3 * (*(const int*)b) + g_sort_axis];
  if (ca < cb) return -1;
  if (ca > cb) return 1;
  return (*(const int*)a) - (*(const int*)b);
}

static int bvh_build_rec(bvh_builder* B, int* ids, int n) {
  const int node = B->nnode++;
  double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  double clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = 0; i < n; i++) {
    const int32_t* f = B->face + 3 * ids[i];
    for (int k = 0; k < 3; k++) {
      for (int j = 0; j < 3; j++) {
        const double p = (double)B->vert[3 * f[k] + j];
        if (p < lo[j]) lo[j] = p;
        if (p > hi[j]) hi[j] = p;
      }
    }
    for (int j = 0; j < 3; j++) {
      const double c = B->cen[3 * ids[i] + j];
      if (c < clo[j]) clo[j] = c;
      if (c > chi[j]) chi[j] = c;
    }
  }
  for (int j = 0; j < 3; j++) {
    /* inflate flat boxes so the strict slab test can pass */
    if (hi[j] - lo[j] < 1e-14) {
      lo[j] -= 1e-14;
      hi[j] += 1e-14;
    }
    B->aabb[6 * node + j] = 0.5 * (lo[j] + hi[j]);
    B->aabb[6 * node + 3 + j] = 0.5 * (hi[j] - lo[j]);
    /* guard against centre/half rounding shaving the box */
    B->aabb[6 * node + 3 + j] *= 1.0 + 1e-12;
  }
  if (n == 1) {
    B->nodeid[node] = ids[0];
    B->child[2 * node] = -1;
    B->child[2 * node + 1] = -1;
    return node;
  }
  B->nodeid[node] = -1;
  int axis = 0;
  double ext = chi[0] - clo[0];
  for (int j = 1; j < 3; j++)
    if (chi[j] - clo[j] > ext) {
      ext = chi[j] - clo[j];
      axis = j;
    }
  g_sort_cen = B->cen;
  g_sort_axis = axis;
  qsort(ids, (size_t)n, sizeof(int), cmp_face);
  const int half = n / 2;
  const int l = bvh_build_rec(B, ids, half);
  const int r = bvh_build_rec(B, ids + half, n - half);
  B->child[2 * node] = l;
  B->child[2 * node + 1] = r;
  return node;
}

int orc_mesh_bvh_size(int nface) { return nface > 0 ? 2 * nface - 1 : 0; }

int orc_build_mesh_bvh(const float* vert, const int32_t* face, int nface, double* bvh_aabb,
                       int32_t* bvh_child, int32_t* bvh_nodeid) {
  if (nface <= 0) return 0;
  int* ids = (int*)malloc(sizeof(int) * (size_t)nface);
  double* cen = (double*)malloc(sizeof(double) * 3 * (size_t)nface);
  for (int i = 0; i < nface; i++) {
    ids[i] = i;
    for (int j = 0; j < 3; j++)
      cen[3 * i + j] = ((double)vert[3 * face[3 * i] + j] + (double)vert[3 * face[3 * i + 1] + j] +
                        (double)vert[3 * face[3 * i + 2] + j]) /
                       3.0;
  }
  bvh_builder B = {vert, face, bvh_aabb, bvh_child, bvh_nodeid, 0, 0, cen};
  bvh_build_rec(&B, ids, nface);
  free(ids);
  free(cen);
  return B.nnode;
}

/* ------------------------------------------------------------------------------------------
 * A.2 ray_eliminate
 * ---------------------------------------------------------------------------------------- */
int orc_ray_eliminate(const orc_model* m, int g, const uint8_t* geomgroup, uint8_t flg_static,
                      int bodyexclude) {
  if (m->geom_bodyid[g] == bodyexclude) return 1;
  if (m->geom_matid[g] < 0 && m->geom_rgba[4 * g + 3] == 0) return 1;
  if (m->geom_matid[g] >= 0 && m->mat_rgba[4 * m->geom_matid[g] + 3] == 0) return 1;
  if (!flg_static && m->body_weldid[m->geom_bodyid[g]] == 0) return 1;
  if (!geomgroup) return 0;
  int groupid = m->geom_group[g];
  if (groupid < 0) groupid = 0;
  if (groupid > ORC_NGROUP - 1) groupid = ORC_NGROUP - 1;
  return geomgroup[groupid] == 0;
}

/* ------------------------------------------------------------------------------------------
 * A.5 mj_ray: brute force over geoms, strict '<' so the lowest geom id wins exact ties
 * ---------------------------------------------------------------------------------------- */
static double ray_one_geom(const orc_model* m, const orc_data* d, int g, const double pnt[3],
                           const double vec[3]) {
  if (m->geom_type[g] == ORC_GEOM_MESH) return orc_ray_mesh(m, d, g, pnt, vec);
  return orc_ray_geom(d->geom_xpos + 3 * g, d->geom_xmat + 9 * g, m->geom_size + 3 * g, pnt, vec,
                      m->geom_type[g]);
}

double orc_ray(const orc_model* m, const orc_data* d, const double pnt[3], const double vec[3],
               const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, int32_t* geomid) {
  double dist = -1;
  int32_t id = -1;
  if (dot3(vec, vec) >= ORC_MINVAL * ORC_MINVAL) { /* |vec| < mjMINVAL is an error upstream */
    for (int g = 0; g < m->ngeom; g++) {
      if (orc_ray_eliminate(m, g, geomgroup, flg_static, bodyexclude)) continue;
      const double nd = ray_one_geom(m, d, g, pnt, vec);
      if (nd >= 0 && (nd < dist || dist < 0)) {
        dist = nd;
        id = g;
      }
    }
  }
  if (geomid) *geomid = id;
  return dist;
}

/* ------------------------------------------------------------------------------------------
 * A.6 mj_multiRay: one origin for all rays; per-geom eliminate flags computed once; a cull that
 * skips geoms a ray cannot reach; the surviving geoms are visited in index order with the A.5 update
 * rule, so the result equals orc_ray per ray (hits beyond `cutoff` are reported as misses).
 *   ORC_CULL_BODY_ANGULAR (A.6 as specified): per body the (azimuth, elevation) rectangle of its
 *     bounding box as seen from pnt; a ray whose own (azimuth, elevation) falls outside skips the
 *     whole body.  A body that contains the origin, carries an unbounded plane, or whose azimuth
 *     range exceeds pi (it wraps) keeps the full range; a body whose box is farther than cutoff is
 *     dropped.  The body box is the union of its geoms' world boxes (mjModel's body BVH root is
 *     not part of the scene mirror).  Elevation bounds are taken from the box's z range and its
 *     nearest / farthest horizontal distance, not from its 8 corners only: the corner extremes are
 *     not conservative for a box edge passing close to the vertical through pnt.
 *   ORC_CULL_GEOM_SPHERE: per geom, distance of its bounding sphere from the ray's line -- what the
 *     body cull degenerates to with one geom per body; the stronger accelerator on scenes whose
 *     bodies are spatially scattered.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_model* m;
  const orc_data* d;
  const double* pnt;
  const double* vec;
  const uint8_t* elim;
  int cull;
  /* ORC_CULL_GEOM_SPHERE */
  const double* bs_c;  /* per-geom bounding sphere centre - pnt */
  const double* bs_r2; /* radius^2, <0: no cull (plane) */
  /* ORC_CULL_BODY_ANGULAR */
  const double* body_ba;   /* [nbody*4] az_min, el_min, az_max, el_max */
  const uint8_t* body_on;  /* [nbody] body has a live geom within cutoff */
  const int32_t* body_adr; /* [nbody+1] into body_geom */
  const int32_t* body_geom; /* geoms of each body, ascending */
  int32_t* geomid;
  double* dist;
  int begin, end;
  double cutoff;
} multi_job;

static void multi_ray_range(const multi_job* J) {
  const orc_model* m = J->m;
  for (int i = J->begin; i < J->end; i++) {
    const double* v = J->vec + 3 * (size_t)i;
    double dist = -1;
    int32_t id = -1;
    const double vv = dot3(v, v);
    if (vv >= ORC_MINVAL) {
      if (J->cull == ORC_CULL_BODY_ANGULAR) {
        const double az = atan2(v[1], v[0]);
        const double el = atan2(v[2], sqrt(v[0] * v[0] + v[1] * v[1]));
        for (int b = 0; b < m->nbody; b++) {
          if (!J->body_on[b]) continue;
          const double* ba = J->body_ba + 4 * b;
          if (az < ba[0] || el < ba[1] || az > ba[2] || el > ba[3]) continue;
          for (int k = J->body_adr[b]; k < J->body_adr[b + 1]; k++) {
            const int g = J->body_geom[k];
            const double nd = ray_one_geom(m, J->d, g, J->pnt, v);
            /* bodies are visited in id order; geom ids need not be body-major, so the tie rule
             * is made explicit (identical to strict '<' for MuJoCo's body-major geom order) */
            if (nd >= 0 && (dist < 0 || nd < dist || (nd == dist && g < id))) {
              dist = nd;
              id = g;
            }
          }
        }
      } else {
        for (int g = 0; g < m->ngeom; g++) {
          if (J->elim[g]) continue;
          if (J->cull == ORC_CULL_GEOM_SPHERE && J->bs_r2[g] >= 0) {
            /* squared distance from sphere centre to the ray's line, and side test */
            const double* c = J->bs_c + 3 * g;
            const double cc = dot3(c, c), cv = dot3(c, v);
            if (cc > J->bs_r2[g]) {
              if (cv <= 0) continue;
              if (cc * vv - cv * cv > J->bs_r2[g] * vv * (1 + 1e-9)) continue;
            }
          }
          const double nd = ray_one_geom(m, J->d, g, J->pnt, v);
          if (nd >= 0 && (nd < dist || dist < 0)) {
            dist = nd;
            id = g;
          }
        }
      }
    }
    if (dist > J->cutoff) {
      dist = -1;
      id = -1;
    }
    J->dist[i] = dist;
    if (J->geomid) J->geomid[i] = id;
  }
}

static void* multi_ray_thread(void* arg) {
  multi_ray_range((const multi_job*)arg);
  return NULL;
}

/* radius of a sphere around geom_xpos that contains the geom; < 0: unbounded (plane) */
static double geom_bound_radius(const orc_model* m, int g) {
  const double* s = m->geom_size + 3 * g;
  switch (m->geom_type[g]) {
    case ORC_GEOM_SPHERE:
      return s[0];
    case ORC_GEOM_CAPSULE:
      return s[0] + s[1];
    case ORC_GEOM_CYLINDER:
      return sqrt(s[0] * s[0] + s[1] * s[1]);
    case ORC_GEOM_ELLIPSOID:
      return fmax(s[0], fmax(s[1], s[2]));
    case ORC_GEOM_BOX:
    case ORC_GEOM_MESH:
      return sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
    default:
      return -1;
  }
}

/* world box of geom g (lo[3], hi[3]); returns 0 for an unbounded geom */
static int geom_world_box(const orc_model* m, const orc_data* d, int g, double lo[3], double hi[3]) {
  const double* s = m->geom_size + 3 * g;
  const double* p = d->geom_xpos + 3 * g;
  const double* R = d->geom_xmat + 9 * g;
  const int t = m->geom_type[g];
  if (t == ORC_GEOM_PLANE) return 0;
  for (int k = 0; k < 3; k++) {
    const double r0 = R[3 * k], r1 = R[3 * k + 1], r2 = R[3 * k + 2];
    double e;
    if (t == ORC_GEOM_SPHERE) {
      e = s[0];
    } else if (t == ORC_GEOM_CAPSULE) {
      e = fabs(r2) * s[1] + s[0];
    } else if (t == ORC_GEOM_CYLINDER) {
      const double q = 1.0 - r2 * r2;
      e = fabs(r2) * s[1] + s[0] * sqrt(q > 0 ? q : 0.0);
    } else if (t == ORC_GEOM_ELLIPSOID) {
      e = sqrt(r0 * r0 * s[0] * s[0] + r1 * r1 * s[1] * s[1] + r2 * r2 * s[2] * s[2]);
    } else {
      e = fabs(r0) * s[0] + fabs(r1) * s[1] + fabs(r2) * s[2];
    }
    e = e * (1 + 1e-9) + 1e-12;
    lo[k] = p[k] - e;
    hi[k] = p[k] + e;
  }
  return 1;
}

void orc_multi_ray_cull(const orc_model* m, const orc_data* d, const double pnt[3],
                        const double* vec, const uint8_t* geomgroup, uint8_t flg_static,
                        int bodyexclude, int32_t* geomid, double* dist, int nray, double cutoff,
                        int nthread, int cull) {
  const size_t G = (size_t)m->ngeom, B = (size_t)m->nbody;
  uint8_t* elim = (uint8_t*)malloc(G + 1);
  double* bs_c = (double*)malloc(sizeof(double) * 3 * (G + 1));
  double* bs_r2 = (double*)malloc(sizeof(double) * (G + 1));
  double* body_ba = (double*)malloc(sizeof(double) * 4 * (B + 1));
  uint8_t* body_on = (uint8_t*)calloc(B + 1, 1);
  int32_t* body_adr = (int32_t*)calloc(B + 2, sizeof(int32_t));
  int32_t* body_geom = (int32_t*)malloc(sizeof(int32_t) * (G + 1));
  for (int g = 0; g < m->ngeom; g++) {
    elim[g] = (uint8_t)orc_ray_eliminate(m, g, geomgroup, flg_static, bodyexclude);
    double r = geom_bound_radius(m, g);
    if (r >= 0) {
      r *= 1 + 1e-9;
      bs_r2[g] = r * r;
    } else {
      bs_r2[g] = -1;
    }
    for (int j = 0; j < 3; j++) bs_c[3 * g + j] = d->geom_xpos[3 * g + j] - pnt[j];
  }
  if (cull == ORC_CULL_BODY_ANGULAR) {
    /* geoms of each body (not eliminated), ascending geom id */
    for (int g = 0; g < m->ngeom; g++)
      if (!elim[g]) body_adr[m->geom_bodyid[g] + 1]++;
    for (size_t b = 0; b < B; b++) body_adr[b + 1] += body_adr[b];
    int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (B + 1));
    for (size_t b = 0; b < B; b++) fill[b] = body_adr[b];
    for (int g = 0; g < m->ngeom; g++)
      if (!elim[g]) body_geom[fill[m->geom_bodyid[g]]++] = g;
    free(fill);
    for (int b = 0; b < m->nbody; b++) {
      double* ba = body_ba + 4 * b;
      ba[0] = -M_PI, ba[1] = -M_PI, ba[2] = M_PI, ba[3] = M_PI; /* full range */
      if (body_adr[b + 1] == body_adr[b]) continue;                /* no live geom */
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      int bounded = 1;
      for (int k = body_adr[b]; k < body_adr[b + 1]; k++) {
        double gl[3], gh[3];
        if (!geom_world_box(m, d, body_geom[k], gl, gh)) {
          bounded = 0;
          break;
        }
        for (int j = 0; j < 3; j++) {
          lo[j] = fmin(lo[j], gl[j]);
          hi[j] = fmax(hi[j], gh[j]);
        }
      }
      body_on[b] = 1;
      if (!bounded) continue;
      /* box relative to the origin of the rays */
      double near2 = 0;
      int inside = 1;
      for (int j = 0; j < 3; j++) {
        lo[j] -= pnt[j];
        hi[j] -= pnt[j];
        const double gap = fmax(fmax(lo[j], -hi[j]), 0.0);
        near2 += gap * gap;
        if (gap > 0) inside = 0;
      }
      if (near2 > cutoff * cutoff) { /* farther than cutoff: nothing of it can be reported */
        body_on[b] = 0;
        continue;
      }
      if (inside) continue;
      /* azimuth range from the 4 vertical edges; wrap-around or the vertical through pnt crossing
       * the box's footprint disable the cull */
      if (lo[0] <= 0 && hi[0] >= 0 && lo[1] <= 0 && hi[1] >= 0) continue;
      double az0 = 1e300, az1 = -1e300, rho_max = 0;
      for (int c = 0; c < 4; c++) {
        const double x = (c & 1) ? hi[0] : lo[0], y = (c & 2) ? hi[1] : lo[1];
        const double az = atan2(y, x);
        az0 = fmin(az0, az);
        az1 = fmax(az1, az);
        rho_max = fmax(rho_max, sqrt(x * x + y * y));
      }
      if (az1 - az0 > M_PI) continue; /* straddles the +-pi cut */
      const double gx = fmax(fmax(lo[0], -hi[0]), 0.0), gy = fmax(fmax(lo[1], -hi[1]), 0.0);
      const double rho_min = sqrt(gx * gx + gy * gy);
      /* elevation atan2(z, rho) grows with z; with rho it falls for z > 0 and grows for z < 0 */
      const double el1 = atan2(hi[2], hi[2] > 0 ? rho_min : rho_max);
      const double el0 = atan2(lo[2], lo[2] > 0 ? rho_max : rho_min);
      const double eps = 1e-9;
      ba[0] = az0 - eps, ba[1] = el0 - eps, ba[2] = az1 + eps, ba[3] = el1 + eps;
    }
  }
  if (nthread < 1) nthread = 1;
  if (nthread > nray) nthread = nray > 0 ? nray : 1;
  multi_job* jobs = (multi_job*)malloc(sizeof(multi_job) * (size_t)nthread);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthread);
  for (int t = 0; t < nthread; t++) {
    multi_job J = {m,       d,        pnt,      vec,       elim,   cull,  bs_c,
                   bs_r2,   body_ba,  body_on,  body_adr,  body_geom, geomid, dist,
                   (int)((long long)nray * t / nthread), (int)((long long)nray * (t + 1) / nthread),
                   cutoff};
    jobs[t] = J;
  }
  if (nthread == 1) {
    multi_ray_range(&jobs[0]);
  } else {
    for (int t = 0; t < nthread; t++) pthread_create(&th[t], NULL, multi_ray_thread, &jobs[t]);
    for (int t = 0; t < nthread; t++) pthread_join(th[t], NULL);
  }
  free(jobs);
  free(th);
  free(elim);
  free(bs_c);
  free(bs_r2);
  free(body_ba);
  free(body_on);
  free(body_adr);
  free(body_geom);
}

void orc_multi_ray(const orc_model* m, const orc_data* d, const double pnt[3], const double* vec,
                   const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, int32_t* geomid,
                   double* dist, int nray, double cutoff, int nthread) {
  orc_multi_ray_cull(m, d, pnt, vec, geomgroup, flg_static, bodyexclude, geomid, dist, nray, cutoff,
                     nthread, ORC_CULL_BODY_ANGULAR);
}

/* ------------------------------------------------------------------------------------------
 * A.7 rangefinder + MujocoLidar::update (ref:src/mujoco_lidar.cpp:256-271)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const orc_model* m;
  const orc_data* d;
  const double* pos;
  const double* mat;
  const double* local_dir;
  const uint8_t* geomgroup;
  uint8_t flg_static;
  int bodyexclude;
  double cutoff, range_min, range_max;
  float* ranges;
  int32_t* geomid;
  double* dist_raw;
  int begin, end;
} scan_job;

static void scan_range(const scan_job* J) {
  for (int i = J->begin; i < J->end; i++) {
    const double* l = J->local_dir + 3 * (size_t)i;
    const double* R = J->mat;
    /* world ray = sensor_mat * local_dir (mju_mulMatVec3 order) */
    const double v[3] = {R[0] * l[0] + R[1] * l[1] + R[2] * l[2],
                         R[3] * l[0] + R[4] * l[1] + R[5] * l[2],
                         R[6] * l[0] + R[7] * l[1] + R[8] * l[2]};
    int32_t id;
    /* engine_sensor.c rangefinder: mj_ray(site_xpos, site z axis, NULL, 1, site_bodyid, NULL) */
    double val = orc_ray(J->m, J->d, J->pos, v, J->geomgroup, J->flg_static, J->bodyexclude, &id);
    /* apply_cutoff for a positive-valued sensor */
    if (J->cutoff > 0 && val > J->cutoff) val = J->cutoff;
    if (J->dist_raw) J->dist_raw[i] = val;
    /* lidar.cpp:262 double -> float; :269-271 filter on the float against double limits */
    float r = (float)val;
    r = (float)((r < J->range_min || r > J->range_max) ? -1.0 : r);
    J->ranges[i] = r;
    if (J->geomid) J->geomid[i] = id;
  }
}

static void* scan_thread(void* arg) {
  scan_range((const scan_job*)arg);
  return NULL;
}

void orc_lidar_scan(const orc_model* m, const orc_data* d, const double sensor_pos[3],
                    const double sensor_mat[9], const double* local_dir, int nray,
                    const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude, double cutoff,
                    double range_min, double range_max, float* ranges, int32_t* geomid,
                    double* dist_raw, int nthread) {
  if (nthread < 1) nthread = 1;
  if (nthread > nray) nthread = nray > 0 ? nray : 1;
  scan_job* jobs = (scan_job*)malloc(sizeof(scan_job) * (size_t)nthread);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthread);
  for (int t = 0; t < nthread; t++) {
    scan_job J = {m,         d,         sensor_pos, sensor_mat, local_dir, geomgroup, flg_static,
                  bodyexclude, cutoff,  range_min,  range_max,  ranges,    geomid,    dist_raw,
                  (int)((long long)nray * t / nthread),
                  (int)((long long)nray * (t + 1) / nthread)};
    jobs[t] = J;
  }
  if (nthread == 1) {
    scan_range(&jobs[0]);
  } else {
    for (int t = 0; t < nthread; t++) pthread_create(&th[t], NULL, scan_thread, &jobs[t]);
    for (int t = 0; t < nthread; t++) pthread_join(th[t], NULL);
  }
  free(jobs);
  free(th);
}

/* The same scan computed the mj_multiRay way: all rays of the scan share the origin, so the cull
 * of orc_multi_ray_cull is prepared once per scan; then the rangefinder cutoff and the LaserScan
 * cast / filter as in orc_lidar_scan.  Results are identical to orc_lidar_scan. */
void orc_lidar_scan_multiray(const orc_model* m, const orc_data* d, const double sensor_pos[3],
                             const double sensor_mat[9], const double* local_dir, int nray,
                             const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude,
                             double cutoff, double range_min, double range_max, float* ranges,
                             int32_t* geomid, double* dist_raw, int nthread, int cull) {
  double* vec = (double*)calloc(3 * ((size_t)nray + 1), sizeof(double));
  double* dist = (double*)calloc((size_t)nray + 1, sizeof(double));
  const double* R = sensor_mat;
  for (int i = 0; i < nray; i++) {
    const double* l = local_dir + 3 * (size_t)i;
    vec[3 * (size_t)i + 0] = R[0] * l[0] + R[1] * l[1] + R[2] * l[2];
    vec[3 * (size_t)i + 1] = R[3] * l[0] + R[4] * l[1] + R[5] * l[2];
    vec[3 * (size_t)i + 2] = R[6] * l[0] + R[7] * l[1] + R[8] * l[2];
  }
  orc_multi_ray_cull(m, d, sensor_pos, vec, geomgroup, flg_static, bodyexclude, geomid, dist, nray,
                     INFINITY, nthread, cull);
  for (int i = 0; i < nray; i++) {
    double val = dist[i];
    if (cutoff > 0 && val > cutoff) val = cutoff;
    if (dist_raw) dist_raw[i] = val;
    float r = (float)val;
    r = (float)((r < range_min || r > range_max) ? -1.0 : r);
    ranges[i] = r;
  }
  free(vec);
  free(dist);
}

/* ref:src/mujoco_lidar.cpp:92-93 */
int orc_lidar_num_rangefinders(double min_angle, double max_angle, double angle_increment) {
  return (int)((max_angle - min_angle) / angle_increment) + 1;
}

/* ------------------------------------------------------------------------------------------
 * cameras
 * ---------------------------------------------------------------------------------------- */

/* ref:src/mujoco_cameras.cpp:119-124 */
void orc_camera_intrinsics(double fovy_deg, int width, int height, double K[9]) {
  for (int i = 0; i < 9; i++) K[i] = 0.0;
  const double focal_scaling = (1.0 / tan((fovy_deg * M_PI / 180.0) / 2.0)) * height / 2.0;
  K[0] = focal_scaling;
  K[2] = (double)width / 2.0;
  K[4] = focal_scaling;
  K[5] = (double)height / 2.0;
  K[8] = 1.0;
}

/* ref:src/mujoco_cameras.cpp:222-240 -- float32 throughout */
void orc_depth_linearise_flip(const float* gl_depth, int width, int height, double znear,
                              double zfar, double extent, float* depth_flipped) {
  const float near = (float)(znear * extent);
  const float far = (float)(zfar * extent);
  const float depth_scale = 1.0f - near / far;
  for (int h = 0; h < height; h++) {
    for (int w = 0; w < width; w++) {
      const size_t idx = (size_t)h * width + w;
      const size_t idx_flipped = (size_t)(height - 1 - h) * width + w;
      const float z = near / (1.0f - gl_depth[idx] * depth_scale);
      depth_flipped[idx_flipped] = z;
    }
  }
}

/* depth_image_proc::convertDepth<float> semantics: organised cloud, top-left
 * origin, optical frame; non-finite depth -> NaN point. */
void orc_depth_to_cloud(const float* depth_flipped, int width, int height, const double K[9],
                        int point_floats, float* cloud) {
  const float center_x = (float)K[2];
  const float center_y = (float)K[5];
  const float constant_x = 1.0f / (float)K[0];
  const float constant_y = 1.0f / (float)K[4];
  const float bad = NAN;
  for (int v = 0; v < height; v++) {
    for (int u = 0; u < width; u++) {
      const size_t idx = (size_t)v * width + u;
      const float depth = depth_flipped[idx];
      float* p = cloud + idx * (size_t)point_floats;
      if (!isfinite(depth)) {
        p[0] = p[1] = p[2] = bad;
      } else {
        p[0] = ((float)u - center_x) * depth * constant_x;
        p[1] = ((float)v - center_y) * depth * constant_y;
        p[2] = depth;
      }
      if (point_floats == 4) p[3] = 0.0f;
    }
  }
}
