// scene_build.cpp -- see scene_build.h.
#include "scene_build.h"

#include <cfloat>
#include <cmath>

#include "ray_core.cuh"

namespace b2 {

void eliminate_flags(const HostModel& m, const uint8_t* geomgroup, uint8_t flg_static,
                     int bodyexclude, std::vector<uint8_t>& out) {
  out.assign((size_t)m.ngeom, 0);
  for (int g = 0; g < m.ngeom; g++) {
    bool e = false;
    const int body = m.geom_bodyid[g];
    if (body == bodyexclude) e = true;
    const int matid = m.geom_matid[g];
    if (matid < 0 && m.geom_rgba[4 * (size_t)g + 3] == 0) e = true;
    if (matid >= 0 && matid < m.nmat && m.mat_rgba[4 * (size_t)matid + 3] == 0) e = true;
    if (!flg_static && m.body_weldid[body] == 0) e = true;
    if (geomgroup) {
      int grp = m.geom_group[g];
      grp = grp < 0 ? 0 : (grp > 5 ? 5 : grp);
      if (geomgroup[grp] == 0) e = true;
    }
    out[g] = e ? 1 : 0;
  }
}

bool build_mesh_bvhs(int nmesh, const int32_t* vertadr, const int32_t* vertnum,
                     const int32_t* faceadr, const int32_t* facenum, const float* vert_all,
                     const int32_t* face_all, int max_leaf, int max_depth, HostMeshes& out,
                     std::string& err) {
  out.infos.assign((size_t)nmesh, MeshInfo{});
  out.nodes.clear();
  out.tris.clear();
  for (int k = 0; k < nmesh; k++) {
    const int nf = facenum[k], nv = vertnum[k];
    const float* vert = vert_all + 3 * (size_t)vertadr[k];
    const int32_t* face = face_all + 3 * (size_t)faceadr[k];
    std::vector<Aabb> boxes((size_t)nf);
    for (int f = 0; f < nf; f++) {
      Aabb b;
      for (int j = 0; j < 3; j++) {
        b.lo[j] = INFINITY;
        b.hi[j] = -INFINITY;
      }
      for (int c = 0; c < 3; c++) {
        const int vi = face[3 * f + c];
        if (vi < 0 || vi >= nv) {
          err = "mesh_face index out of range";
          return false;
        }
        for (int j = 0; j < 3; j++) {
          b.lo[j] = std::fmin(b.lo[j], vert[3 * vi + j]);
          b.hi[j] = std::fmax(b.hi[j], vert[3 * vi + j]);
        }
      }
      // pad: the double triangle test accepts points within rounding of the triangle's plane
      for (int j = 0; j < 3; j++) {
        const float pad = 1e-6f * std::fmax(std::fabs(b.lo[j]), std::fabs(b.hi[j])) + 1e-9f;
        b.lo[j] = std::nextafter(b.lo[j] - pad, -INFINITY);
        b.hi[j] = std::nextafter(b.hi[j] + pad, INFINITY);
      }
      boxes[f] = b;
    }
    MeshInfo mi;
    mi.node_base = (int32_t)out.nodes.size();
    mi.tri_base = (int32_t)out.tris.size();
    mi.ntri = nf;
    {
      // bounding sphere of the vertices the faces use: centre = middle of their box, radius = the
      // farthest of them (double), inflated and rounded up
      double lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
      for (int f = 0; f < nf; f++)
        for (int c = 0; c < 3; c++)
          for (int j = 0; j < 3; j++) {
            const double v = vert[3 * face[3 * f + c] + j];
            lo[j] = std::fmin(lo[j], v);
            hi[j] = std::fmax(hi[j], v);
          }
      double r2 = 0.0;
      for (int j = 0; j < 3; j++) mi.bs_c[j] = nf > 0 ? (float)(0.5 * (lo[j] + hi[j])) : 0.f;
      for (int f = 0; f < nf; f++)
        for (int c = 0; c < 3; c++) {
          double d2 = 0.0;
          for (int j = 0; j < 3; j++) {
            const double d = (double)vert[3 * face[3 * f + c] + j] - (double)mi.bs_c[j];
            d2 += d * d;
          }
          r2 = std::fmax(r2, d2);
        }
      const double r = std::sqrt(r2) * (1.0 + 1e-6) + 1e-30;
      mi.bs_r = std::nextafter((float)r, INFINITY);
    }
    if (nf == 0) {  // empty mesh: a root with four absent children
      MeshNode4 root{};
      for (int c = 0; c < 4; c++) {
        root.lox[c] = root.loy[c] = root.loz[c] = FLT_MAX;
        root.hix[c] = root.hiy[c] = root.hiz[c] = -FLT_MAX;
        root.child[c] = kEmptyChild;
      }
      mi.nnode = 1;
      out.nodes.push_back(root);
      out.infos[k] = mi;
      continue;
    }
    BuiltBvh bvh = build_bvh(boxes, max_leaf, false, 3.0f);
    int depth4 = 0;
    const std::vector<MeshNode4> n4 = collapse4(bvh.nodes, &depth4);
    // every visited node pushes at most three postponed children
    if (3 * depth4 + 1 > max_depth) {
      err = "mesh BVH deeper than the traversal stack";
      return false;
    }
    mi.nnode = (int32_t)n4.size();
    out.infos[k] = mi;
    out.nodes.insert(out.nodes.end(), n4.begin(), n4.end());
    for (int i = 0; i < nf; i++) {
      const int f = bvh.order[i];
      MeshTri t;
      for (int j = 0; j < 3; j++) {
        t.v0[j] = vert[3 * face[3 * f + 0] + j];
        t.v1[j] = vert[3 * face[3 * f + 1] + j];
        t.v2[j] = vert[3 * face[3 * f + 2] + j];
      }
      t.v0[3] = t.v1[3] = t.v2[3] = 0.f;
      out.tris.push_back(t);
    }
  }
  return true;
}

bool build_geom_topology(const HostModel& m, const double* xpos, const double* xmat, int max_depth,
                         WideBvh& wide, std::string& err) {
  BuiltBvh out;
  std::vector<Aabb> boxes;
  boxes.reserve(m.bvh_geoms.size());
  for (int g : m.bvh_geoms) {
    GeomPose P;
    for (int k = 0; k < 3; k++) P.pos[k] = xpos[3 * (size_t)g + k];
    for (int k = 0; k < 9; k++) P.mat[k] = xmat[9 * (size_t)g + k];
    Aabb b;
    geom_aabb(m.geom_type[g], &m.geom_size[3 * (size_t)g], P, b.lo, b.hi);
    boxes.push_back(b);
  }
  out = build_bvh(boxes, 1, true, 4.0f);
  // leaf payloads index `boxes`; map them to geom ids
  for (auto& n : out.nodes) {
    if (n.c0 < 0 && n.c0 != kEmptyChild) n.c0 = ~m.bvh_geoms[(size_t)(~n.c0)];
    if (n.c1 < 0 && n.c1 != kEmptyChild) n.c1 = ~m.bvh_geoms[(size_t)(~n.c1)];
  }
  wide = collapse_wide(out);
  if (wide.depth > max_depth) {
    err = "geom BVH deeper than the traversal frontier allows";
    return false;
  }
  return true;
}

}  // namespace b2
