// emu.cpp -- CPU EMULATION HARNESS of the cast kernel's traversal logic.  TEST INFRASTRUCTURE ONLY.
//
// Compiles the product's __host__ __device__ traversal (csrc/ray_core.cuh + csrc/tile_trace.cuh:
// trace_tile, mesh_trace, the fp64 leaf tests, the conservative fp32 slabs and cone test,
// geom_aabb) and its host-side scene construction (csrc/scene_build.cpp, csrc/bvh_build.cpp) for
// the CPU, with a 1-lane "warp", so the logic that runs on the B200 can be compared with the
// oracle in the CPU-only test tier.  It is never linked into libb2sense.so and is not a fallback:
// the product has no CPU path.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b2sense.h"
#include "../../mujoco_ros2_simulation_b200/csrc/b2_types.h"
#include "../../mujoco_ros2_simulation_b200/csrc/bvh_build.h"
#include "../../mujoco_ros2_simulation_b200/csrc/ray_core.cuh"
#include "../../mujoco_ros2_simulation_b200/csrc/scene_build.h"

using namespace b2;

namespace {

long long g_emu_fallbacks = 0;  // groups the two-kernel emulation left to trace_group

struct HostWarp {
  F4 cand_a_[kCandCap], cand_b_[kCandCap];
  int32_t queue_[kQueueCap], order_[kCandCap];
  int size() const { return 1; }
  int lane() const { return 0; }
  unsigned lanemask_lt() const { return 0u; }
  unsigned ballot(bool p) const { return p ? 1u : 0u; }
  bool any(bool p) const { return p; }
  int popc(unsigned m) const { return __builtin_popcount(m); }
  void sync() const {}
  float sum(float v) const { return v; }
  float min(float v) const { return v; }
  float max(float v) const { return v; }
  int32_t* queue() { return queue_; }
  F4* cand_a() { return cand_a_; }
  F4* cand_b() { return cand_b_; }
  int32_t* order() { return order_; }
  RayState st_[kGroupTiles];
  int slot_ = -1;
  RayState* ray_state() { return st_; }
  CandRef cref_[kCandCap];
  CandRef* cref() { return cref_; }
  void set_slot(int v) { slot_ = v; }
  int get_slot() const { return slot_; }
  // the emulation traces meshes in place (the device's clean-up / small-launch policy)
  void mesh_pair(const CastParams& p, const GeomPose* pose, int g, bool h, Hit& hit,
                 const double pnt[3], const double vec[3], float tcap) {
    mesh_inline(p, pose, g, h, hit, pnt, vec, tcap);
  }
};

struct EmuScene {
  HostModel hm;
  HostMeshes meshes;
  std::vector<GeomStatic> gstat;
  std::vector<GeomPose> pose;
  WideBvh topo;
  std::vector<WideChild> slots;
};

// the refit kernel's per-world work, serially (kernels.cu refit_kernel)
void host_refit(EmuScene& S) {
  const int nslot = (int)S.topo.slot_link.size();
  S.slots.assign((size_t)nslot, WideChild{});
  const int nlevel = (int)S.topo.refit_level_start.size() - 1;
  for (int l = 0; l < nlevel; l++) {
    for (int i = S.topo.refit_level_start[l]; i < S.topo.refit_level_start[l + 1]; i++) {
      const int s = S.topo.refit_slots[i];
      const int32_t link = S.topo.slot_link[s];
      WideChild out{};
      out.link = link;
      if (link < 0) {
        const int g = ~link;
        geom_aabb(S.gstat[g].type, S.gstat[g].size, S.pose[g], out.lo, out.hi);
        out.pad = S.gstat[g].type;
      } else {
        const WideNode wn = S.topo.nodes[link];
        for (int k = 0; k < 3; k++) {
          out.lo[k] = FLT_MAX;
          out.hi[k] = -FLT_MAX;
        }
        for (int c = wn.first; c < wn.first + wn.count; c++)
          for (int k = 0; k < 3; k++) {
            out.lo[k] = fminf(out.lo[k], S.slots[c].lo[k]);
            out.hi[k] = fmaxf(out.hi[k], S.slots[c].hi[k]);
          }
      }
      S.slots[s] = out;
    }
  }
}

bool make_scene(const b2_model_desc* m, const double* xpos, const double* xmat,
                const double* topo_xpos, const double* topo_xmat, int mesh_leaf, EmuScene& S,
                std::string& err) {
  const size_t G = (size_t)m->ngeom;
  HostModel& hm = S.hm;
  hm.ngeom = m->ngeom;
  hm.nbody = m->nbody;
  hm.nmesh = m->nmesh;
  hm.nmat = m->nmat;
  hm.geom_type.assign(m->geom_type, m->geom_type + G);
  hm.geom_bodyid.assign(m->geom_bodyid, m->geom_bodyid + G);
  hm.geom_group.assign(m->geom_group, m->geom_group + G);
  hm.geom_matid.assign(m->geom_matid, m->geom_matid + G);
  hm.geom_dataid.assign(m->geom_dataid, m->geom_dataid + G);
  hm.geom_size.assign(m->geom_size, m->geom_size + 3 * G);
  hm.geom_rgba.assign(m->geom_rgba, m->geom_rgba + 4 * G);
  if (m->nmat > 0) hm.mat_rgba.assign(m->mat_rgba, m->mat_rgba + 4 * (size_t)m->nmat);
  hm.body_weldid.assign(m->body_weldid, m->body_weldid + m->nbody);
  S.gstat.resize(G);
  S.pose.resize(G);
  for (size_t g = 0; g < G; g++) {
    for (int k = 0; k < 3; k++) S.gstat[g].size[k] = m->geom_size[3 * g + k];
    S.gstat[g].type = m->geom_type[g];
    S.gstat[g].dataid = m->geom_dataid[g];
    if (m->geom_type[g] == GEOM_PLANE)
      hm.plane_ids.push_back((int32_t)g);
    else
      hm.bvh_geoms.push_back((int32_t)g);
    for (int k = 0; k < 3; k++) S.pose[g].pos[k] = xpos[3 * g + k];
    for (int k = 0; k < 9; k++) S.pose[g].mat[k] = xmat[9 * g + k];
  }
  if (!build_mesh_bvhs(m->nmesh, m->mesh_vertadr, m->mesh_vertnum, m->mesh_faceadr,
                       m->mesh_facenum, m->mesh_vert, m->mesh_face, mesh_leaf, kMeshStack - 2,
                       S.meshes, err))
    return false;
  // topology from (possibly different) poses, boxes refit to the actual poses: exactly what the
  // runtime does for worlds other than the one the topology was built on
  if (!build_geom_topology(hm, topo_xpos ? topo_xpos : xpos, topo_xmat ? topo_xmat : xmat,
                           kMaxWideDepth, S.topo, err))
    return false;
  host_refit(S);
  return true;
}

void fill_params(const EmuScene& S, CastParams& p) {
  p.wslots = S.slots.data();
  p.wnodes = S.topo.nodes.data();
  p.pose = S.pose.data();
  p.gstat = S.gstat.data();
  p.plane_ids = S.hm.plane_ids.data();
  p.mesh = S.meshes.infos.data();
  p.mnodes = S.meshes.nodes.data();
  p.mtris = S.meshes.tris.data();
  p.nslot = (int)S.slots.size();
  p.nwnode = (int)S.topo.nodes.size();
  p.ngeom = S.hm.ngeom;
  p.nplane = (int)S.hm.plane_ids.size();
  p.nworld = 1;
}

void run(const EmuScene& S, CastParams& p, const double o[3], const double* R) {
  (void)S;
  HostWarp warp;
  const float tcap = p.mode == MODE_MULTIRAY ? f32_up(p.cutoff) : FLT_MAX;
  const char* two_env = std::getenv("B2_EMU_TWO_KERNEL");
  const bool two = two_env && std::atoi(two_env) != 0;
  // groups of kGroupTiles one-ray "tiles": consecutive rays share one collect, as on the device
  for (int r0 = 0; r0 < p.nray; r0 += kGroupTiles) {
    const int nt = p.nray - r0 < kGroupTiles ? p.nray - r0 : kGroupTiles;
    auto gen = [&](int t, double vec[3], int& r) {
      double l[3];
      r = r0 + t;
      return cast_make_ray(p, p.mode, R, r, l, vec);
    };
    auto epi = [&](int r, const Hit& hit, int) {
      double l[3], vec[3];
      cast_make_ray(p, p.mode, R, r, l, vec);
      cast_epilogue(p, p.mode, 0, r, hit, l);
    };
    // B2_EMU_TWO_KERNEL=1: the two-kernel cast (split_trace.cuh) -- collect_group, then drain_tile
    // per tile, with the fused trace_group for the groups the single collect pass gives up on
    if (two) {
      CandRef list[kCandCap];
      const int n = collect_group<false>(p, p.wslots, p.wnodes, p.elim, p.pose, nt, o, tcap, warp,
                                         gen, list);
      if (n >= 0) {
        // staged (the device's default) and unstaged candidate access, alternating per group
        const bool staged = ((r0 / kGroupTiles) & 1) == 0;
        if (staged) stage_candidates(warp, p.wslots, n, list, warp.cand_a(), warp.cand_b());
        for (int t = 0; t < nt; t++) {
          if (staged)
            drain_tile<true>(p, p.wslots, p.elim, p.pose, t, o, tcap, warp, gen, epi, n, list,
                             warp.cand_a(), warp.cand_b());
          else
            drain_tile<false>(p, p.wslots, p.elim, p.pose, t, o, tcap, warp, gen, epi, n, list,
                              nullptr, nullptr);
        }
        continue;
      }
      g_emu_fallbacks++;
    }
    trace_group<false>(p, p.wslots, p.wnodes, p.elim, p.pose, nt, o, tcap, warp, gen, epi);
  }
}

}  // namespace

extern "C" {

// mirrors b2_multiray for one world; topo_xpos/topo_xmat (nullable) = poses the BVH topology is
// built on when different from the poses the rays are cast against
int emu_multiray(const b2_model_desc* m, const double* xpos, const double* xmat,
                 const double* topo_xpos, const double* topo_xmat, const double pnt[3],
                 const double* vec, const uint8_t* geomgroup, uint8_t flg_static, int bodyexclude,
                 int32_t* geomid, double* dist, int nray, double cutoff, int mesh_leaf) {
  EmuScene S;
  std::string err;
  if (!make_scene(m, xpos, xmat, topo_xpos, topo_xmat, mesh_leaf, S, err)) return -1;
  std::vector<uint8_t> elim;
  eliminate_flags(S.hm, geomgroup, flg_static, bodyexclude, elim);
  elim.push_back(0);
  CastParams p{};
  fill_params(S, p);
  p.elim = elim.data();
  p.local_dir = vec;
  p.nray = nray;
  p.mode = MODE_MULTIRAY;
  p.cutoff = cutoff;
  p.nscan = 1;
  p.geomid = geomid;
  p.dist = dist;
  run(S, p, pnt, nullptr);
  return 0;
}

// mirrors b2_lidar_cast for one scan in one world
int emu_lidar(const b2_model_desc* m, const double* xpos, const double* xmat,
              const b2_lidar_desc* d, const double pos[3], const double mat[9], float* ranges,
              int32_t* geomid, double* dist, float* xyz, int mesh_leaf) {
  EmuScene S;
  std::string err;
  if (!make_scene(m, xpos, xmat, nullptr, nullptr, mesh_leaf, S, err)) return -1;
  std::vector<uint8_t> elim;
  eliminate_flags(S.hm, d->geomgroup, d->flg_static, d->bodyexclude, elim);
  elim.push_back(0);
  CastParams p{};
  fill_params(S, p);
  p.elim = elim.data();
  p.local_dir = d->local_dir;
  p.nray = d->nray;
  p.mode = MODE_LIDAR;
  p.cutoff = d->cutoff;
  p.range_min = d->range_min;
  p.range_max = d->range_max;
  p.nscan = 1;
  p.ranges = ranges;
  p.geomid = geomid;
  p.dist = dist;
  p.xyz = xyz;
  run(S, p, pos, mat);
  return 0;
}

long long emu_fallbacks() { return g_emu_fallbacks; }

// the sign-only bounding-sphere pre-test: out[i] = sphere_miss(pos, r*r, pnt, vec)
void emu_sphere_miss_batch(int n, const double* pos, const double* radius, const double* pnt,
                           const double* vec, uint8_t* out) {
  for (int i = 0; i < n; i++)
    out[i] = sphere_miss(pos + 3 * (size_t)i, radius[i] * radius[i], pnt + 3 * (size_t)i,
                         vec + 3 * (size_t)i)
                 ? 1
                 : 0;
}

// n independent (geom, ray) cases through the product's prim_test (the kernels' leaf routine)
void emu_prim_batch(int n, const double* pos, const double* mat, const double* size,
                    const double* pnt, const double* vec, const int32_t* type, double* out) {
  for (int i = 0; i < n; i++) {
    GeomPose P;
    GeomStatic gs{};
    for (int k = 0; k < 3; k++) P.pos[k] = pos[3 * (size_t)i + k];
    for (int k = 0; k < 9; k++) P.mat[k] = mat[9 * (size_t)i + k];
    for (int k = 0; k < 3; k++) gs.size[k] = size[3 * (size_t)i + k];
    gs.type = type[i];
    const double* o = pnt + 3 * (size_t)i;
    const double* v = vec + 3 * (size_t)i;
    out[i] = prim_test(type[i], &P, &gs, o[0], o[1], o[2], v[0], v[1], v[2]);
  }
}

// BVH statistics for tests: child slots / depth of the wide geom BVH and total mesh nodes
int emu_bvh_stats(const b2_model_desc* m, const double* xpos, const double* xmat, int mesh_leaf,
                  int32_t* nslot, int32_t* depth, int64_t* mesh_nodes) {
  EmuScene S;
  std::string err;
  if (!make_scene(m, xpos, xmat, nullptr, nullptr, mesh_leaf, S, err)) return -1;
  *nslot = (int32_t)S.slots.size();
  *depth = S.topo.depth;
  *mesh_nodes = (int64_t)S.meshes.nodes.size();
  return 0;
}

// Property check of the tile cone (tests): builds the cone of `n` rays from one origin exactly as
// trace_tile does (here with a serial reduction) and counts boxes some ray's slab test touches but
// cone_box rejects.  Must be 0: the collect phase may never drop a box a lane needs.
// kept / touched (nullable): boxes the bound keeps / boxes some ray's slab test touches (tightness)
int emu_cone_check(const double pnt[3], const double* vec, int n, const float* boxes, int nbox,
                   int* kept, int* ntouched) {
  TileCone tc;
  float sx = 0, sy = 0, sz = 0;
  std::vector<float> df((size_t)n * 3);
  for (int i = 0; i < n; i++) {
    const float s = (float)dot3(vec + 3 * i, vec + 3 * i);
    const float inv = rsqrt_f32(s);
    for (int k = 0; k < 3; k++) df[3 * i + k] = (float)vec[3 * i + k] * inv;
    sx += df[3 * i];
    sy += df[3 * i + 1];
    sz += df[3 * i + 2];
  }
  const float an = sqrtf(sx * sx + sy * sy + sz * sz);
  tc.wide = !(an > 1e-3f);
  const float ian = tc.wide ? 0.f : 1.0f / an;
  tc.a[0] = sx * ian;
  tc.a[1] = sy * ian;
  tc.a[2] = sz * ian;
  float cmin = 1.0f;
  for (int i = 0; i < n; i++)
    cmin = fminf(cmin, df[3 * i] * tc.a[0] + df[3 * i + 1] * tc.a[1] + df[3 * i + 2] * tc.a[2]);
  tc.cos_t = fminf(fmaxf(cmin - 8e-6f, -1.0f), 1.0f);
  if (tc.cos_t < 0.2f) tc.wide = true;
  tc.sin_t = sqrtf(fmaxf(1.0f - tc.cos_t * tc.cos_t, 0.f)) + 8e-6f;
  tc.omag = 0.f;
  for (int k = 0; k < 3; k++) {
    tc.o[k] = (float)pnt[k];
    tc.oslack[k] = 2e-6f * fabsf(tc.o[k]) + 1e-30f;
    tc.omag += fabsf(tc.o[k]);
  }
  tc.far2 = INFINITY;
  int bad = 0;
  for (int b = 0; b < nbox; b++) {
    const float* lo = boxes + 6 * b;
    const float* hi = lo + 3;
    float key2;
    const bool keep = cone_box(tc, lo, hi, key2);
    bool touched = false;
    for (int i = 0; i < n && !touched; i++) {
      RayF rf;
      rayf_setup(rf, pnt, vec + 3 * i);
      float tn;
      if (slab_test(rf, lo, hi, FLT_MAX, tn)) {
        touched = true;
        // the ordering key must not exceed the true entry distance of any ray
        const float vlen = sqrtf((float)dot3(vec + 3 * i, vec + 3 * i));
        const double entry = (double)tn * vlen;
        if ((double)key2 > entry * entry * (1.0 + 1e-4) + 1e-12) bad++;
      }
    }
    if (touched && !keep) bad++;
    if (kept && keep) ++*kept;
    if (ntouched && touched) ++*ntouched;
  }
  return bad;
}

int emu_cone_violations(const double pnt[3], const double* vec, int n, const float* boxes,
                        int nbox) {
  return emu_cone_check(pnt, vec, n, boxes, nbox, nullptr, nullptr);
}

}  // extern "C"
