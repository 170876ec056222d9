// tile_trace.cuh -- warp-cooperative traversal of the wide geom BVH for a GROUP of ray tiles
// (included at the end of ray_core.cuh; same host/device rules).
//
// One warp owns a group of up to kGroupTiles tiles of 32 rays (one ray per lane per tile) that
// share an origin, e.g. 32 azimuths x 4 scan lines of a spinning lidar.
//   A. per tile: generate the ray, test the planes (kept outside the BVH: their hit bounds the
//      search distance), accumulate the group's bounding cone.  Per-ray state (best distance, geom
//      id, mesh slot) lives in shared memory between the phases, not in registers;
//   B. COLLECT once for the whole group: the warp walks the fanout-32 geom BVH; every round pops a
//      node and tests ALL its children at once, one lane per child slot, against the group cone and
//      against the farthest distance any ray of the group still needs; surviving inner children go
//      to a small frontier queue, surviving geoms to a shared-memory candidate list, which is then
//      rank-sorted by distance from the origin.  (One collect per tile cost half of the kernel's
//      instructions; neighbouring tiles see almost the same clusters -- profiles/r01.)
//   C. DRAIN per tile: the ray is regenerated and the candidates are visited nearest first by all
//      lanes together: per-lane conservative slab test, then the closed-form fp64 test of the
//      (warp-uniform) geom type; the loop ends as soon as no lane can improve.  If the candidate
//      list fills up before the frontier is empty it is drained early and the collect continues
//      with the tightened distance bound (incoherent groups stay correct, just slower);
//   D. mesh geoms a ray enters are not traced here: the warp policy appends (ray, mesh geom) pairs
//      to a compacted list and separate mesh kernels trace them with full warps;
//   E. epilogue per tile from the shared-memory state.
#ifndef B2_TILE_TRACE_CUH
#define B2_TILE_TRACE_CUH

namespace b2 {

constexpr int kMaxWideDepth = 5;  // queue: every round pops 1 node and pushes <= 32
#ifndef B2_GROUP_TILES
#define B2_GROUP_TILES 4
#endif
constexpr int kGroupTiles = B2_GROUP_TILES;  // tiles per collect
// cand_a doubles as the buffer of the group's unit directions (3 floats per ray) while the cone is built
constexpr int kCandA = kCandCap > kGroupTiles * 24 ? kCandCap : kGroupTiles * 24;

struct B2_ALIGN(16) RayState {
  double vec[3];  // the ray direction (generated once, in phase A)
  double dist;    // best distance so far (< 0: none)
  int32_t id;     // its geom
  int32_t slot;   // affected-record slot of the split mesh path (-1: none)
  int32_t r;      // ray index within the scan (-1: this lane has no ray in the tile)
  int32_t pad;
};

// ---------------------------------------------------------------------------------------------
// group cone vs box (conservative).  The cone (apex o, unit axis a, half-angle theta) contains
// every ray of the group; a box is kept if its bounding sphere may touch the cone and its nearest
// point is not farther than any lane still searches.  key2 = squared distance from the apex to
// the box, rounded down (ordering key of the candidate list).
// ---------------------------------------------------------------------------------------------
struct TileCone {
  float o[3], oslack[3];  // apex and 2e-6*|o_k| (rounding slack of lo - o)
  float a[3];
  float cos_t, sin_t;
  float omag;
  float far2;  // squared farthest distance (world units) any lane still needs
  bool wide;   // half-angle too large for the cone test: keep everything within far2
};

B2_HD bool cone_box(const TileCone& tc, const float lo[3], const float hi[3], float& key2) {
  float d2 = 0.f, r2 = 0.f, c[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float gap = fmaxf(fmaxf(lo[k] - tc.o[k], tc.o[k] - hi[k]), 0.f);
    const float g = fmaxf(gap * (1.0f - 2e-6f) - tc.oslack[k], 0.f);
    d2 += g * g;
    c[k] = 0.5f * (lo[k] + hi[k]) - tc.o[k];
    const float h = 0.5f * (hi[k] - lo[k]);
    r2 += h * h;
  }
  key2 = d2 * (1.0f - 1e-6f);
  if (key2 > tc.far2) return false;
  if (tc.wide) return true;
  const float along = c[0] * tc.a[0] + c[1] * tc.a[1] + c[2] * tc.a[2];
  // |c x a| : distance of the box centre from the cone axis, without cancellation
  const float px = c[1] * tc.a[2] - c[2] * tc.a[1];
  const float py = c[2] * tc.a[0] - c[0] * tc.a[2];
  const float pz = c[0] * tc.a[1] - c[1] * tc.a[0];
  const float perp = sqrtf(px * px + py * py + pz * pz);
  const float err = 4e-6f * (tc.omag + fabsf(along) + perp) + 1e-7f;
  const float rr = r2 * 1.0004f;
  // signed distance from the box centre to the cone's lateral surface
  const float ds = perp * tc.cos_t - along * tc.sin_t - err;
  if (ds > 0.f && ds * ds > rr) return false;
  // behind the apex the nearest feature of the cone is the apex itself
  const float e = along * tc.cos_t + perp * tc.sin_t + err;
  if (e < 0.f) {
    const float c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    const float cl = sqrtf(c2) - err;
    if (cl > 0.f && cl * cl > rr) return false;
  }
  return true;
}

// squared search limit of one lane in world units: (tmax * |vec|)^2, rounded up
B2_HD float lane_lim2(const Hit& hit, float tcap, float vlen_up) {
  const float l = hit_tmax(hit, tcap) * vlen_up;
  return l * l * (1.0f + 1e-6f);
}

// unit direction (float) and |vec| rounded up; false if |vec| is outside the float range
B2_HD bool unit_dir(const double vec[3], float df[3], float& vlen_up) {
  const float s = (float)dot3(vec, vec);
  if (!(s > 1e-30f && s < 1e30f)) return false;
  const float inv = rsqrt_f32(s);
  df[0] = (float)vec[0] * inv;
  df[1] = (float)vec[1] * inv;
  df[2] = (float)vec[2] * inv;
  vlen_up = s * inv * (1.0f + 4e-6f);
  return true;
}

// ---------------------------------------------------------------------------------------------
// One group of nt <= kGroupTiles tiles.  Every lane of the warp must call trace_group together.
//   gen(t, vec, r) -> bool  generates this lane's ray of tile t and its index r in the scan
//                           (false: degenerate ray, traced as a miss; r < 0: no ray at all)
//   epi(r, hit, slot)       consumes the result of ray r
// All rays of the group share the origin pnt.
//   wslots : this world's child slots (shared memory when STAGE), wnodes: topology (same)
// ---------------------------------------------------------------------------------------------
template <bool STAGE, class W, class Gen, class Epi>
B2_HD void trace_group(const CastParams& p, const WideChild* wslots, const WideNode* wnodes,
                       const uint8_t* elim, const GeomPose* pose, int nt, const double pnt[3],
                       float tcap, W& warp, const Gen& gen, const Epi& epi) {
  const int L = warp.size(), lane = warp.lane();
  RayState* st = warp.ray_state();  // [kGroupTiles * L], index t * L + lane
  int32_t* queue = warp.queue();
  F4* cand_a = warp.cand_a();  // lo.xyz, key2
  F4* cand_b = warp.cand_b();  // hi.xyz, geom id bits
  int32_t* order = warp.order();
  float* dfbuf = reinterpret_cast<float*>(cand_a);  // unit directions during phase A

  // ---- A. rays, planes, cone accumulation ---------------------------------------------------
  float sx = 0.f, sy = 0.f, sz = 0.f, lim2max = 0.f;
  bool odd = false, any_active = false;
  for (int t = 0; t < nt; t++) {
    double vec[3];
    int r = -1;
    const bool active = gen(t, vec, r);
    Hit hit;
    hit.dist = -1;
    hit.id = -1;
    for (int i = 0; i < p.nplane; i++) {
      const int g = p.plane_ids[i];
      if (elim[g]) continue;
      if (active)
        hit_update(hit,
                   prim_test(GEOM_PLANE, pose + g, p.gstat + g, pnt[0], pnt[1], pnt[2], vec[0],
                             vec[1], vec[2]),
                   g);
    }
    float df[3] = {0.f, 0.f, 0.f}, vlen_up = 0.f;
    if (active) {
      if (unit_dir(vec, df, vlen_up)) {
        lim2max = fmaxf(lim2max, lane_lim2(hit, tcap, vlen_up));
      } else {
        df[0] = df[1] = df[2] = 0.f;
        odd = true;
      }
    }
    sx += df[0];
    sy += df[1];
    sz += df[2];
    float* d = dfbuf + (size_t)(t * L + lane) * 3;
    d[0] = df[0];
    d[1] = df[1];
    d[2] = df[2];
    RayState rs;
    rs.vec[0] = active ? vec[0] : 0.0;  // an all-zero direction marks a ray that is not traced
    rs.vec[1] = active ? vec[1] : 0.0;
    rs.vec[2] = active ? vec[2] : 0.0;
    rs.dist = hit.dist;
    rs.id = hit.id;
    rs.slot = -1;
    rs.r = r;
    rs.pad = 0;
    st[t * L + lane] = rs;
    any_active = any_active || active;
  }

  if (p.nwnode != 0 && warp.any(any_active)) {
    // group cone: axis = mean unit direction, half-angle = largest deviation (+ slack)
    TileCone tc;
    odd = warp.any(odd);  // some |vec| outside the float range: no distance / cone culling
    const float ax = warp.sum(sx), ay = warp.sum(sy), az = warp.sum(sz);
    const float an = sqrtf(ax * ax + ay * ay + az * az);
    tc.wide = odd || !(an > 1e-3f);
    const float ian = tc.wide ? 0.f : 1.0f / an;
    tc.a[0] = ax * ian;
    tc.a[1] = ay * ian;
    tc.a[2] = az * ian;
    float cmin = 1.0f;
    for (int t = 0; t < nt; t++) {
      const float* d = dfbuf + (size_t)(t * L + lane) * 3;
      const bool has = d[0] != 0.f || d[1] != 0.f || d[2] != 0.f;
      if (has) cmin = fminf(cmin, d[0] * tc.a[0] + d[1] * tc.a[1] + d[2] * tc.a[2]);
    }
    cmin = warp.min(cmin);
    tc.cos_t = fminf(fmaxf(cmin - 8e-6f, -1.0f), 1.0f);
    if (tc.cos_t < 0.2f) tc.wide = true;
    tc.sin_t = sqrtf(fmaxf(1.0f - tc.cos_t * tc.cos_t, 0.f)) + 8e-6f;
    tc.omag = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      tc.o[k] = (float)pnt[k];
      tc.oslack[k] = 2e-6f * fabsf(tc.o[k]) + 1e-30f;
      tc.omag += fabsf(tc.o[k]);
    }
    tc.far2 = odd ? INFINITY : warp.max(lim2max);
    warp.sync();  // dfbuf (aliasing cand_a) has been consumed

    int nq = 1, nc = 0;
    if (lane == 0) queue[0] = 0;
    warp.sync();

    while (nq > 0 || nc > 0) {
      // ---- B. COLLECT while there is room for one more node's worth of geoms -------------------
      while (nq > 0 && nc <= kCandCap - kWideFanout) {
        const int32_t node = queue[nq - 1];
        warp.sync();  // everybody has read the top before it is overwritten
        nq--;
        const WideNode wn = wnodes[node];
        // one lane per child slot (a 1-lane emulated warp walks the slots in turn)
        for (int base = 0; base < wn.count; base += L) {
          const int k = base + lane;
          bool pass = false;
          float key2 = 0.f;
          WideChild ch;
          ch.link = 0;
          if (k < wn.count) {
            ch = ld_wslot<STAGE>(wslots + wn.first + k);
            pass = cone_box(tc, ch.lo, ch.hi, key2);
          }
          const bool inner = pass && ch.link >= 0;
          const bool leaf = pass && ch.link < 0 && !elim[~ch.link];
          const unsigned mi = warp.ballot(inner), ml = warp.ballot(leaf);
          if (inner) {
            const int pos = nq + warp.popc(mi & warp.lanemask_lt());
            if (pos < kQueueCap) queue[pos] = ch.link;
          }
          if (leaf) {
            const int pos = nc + warp.popc(ml & warp.lanemask_lt());
            F4 a, b;
            a.x = ch.lo[0];
            a.y = ch.lo[1];
            a.z = ch.lo[2];
            a.w = key2;
            b.x = ch.hi[0];
            b.y = ch.hi[1];
            b.z = ch.hi[2];
            // geom id (24 bits) + geom type (from the slot's pad word, written by the refit) so the
            // drain needs no global load to dispatch
            b.w = int_as_f32((~ch.link) | (ch.pad << 24));
            // the drain will want this geom's pose / size: start the L2 -> L1 fetch now, from all
            // appending lanes at once, instead of stalling on it one candidate at a time
            prefetch_l1(pose + (~ch.link));
            prefetch_l1(reinterpret_cast<const char*>(pose + (~ch.link)) + sizeof(GeomPose) - 1);
            prefetch_l1(p.gstat + (~ch.link));
            cand_a[pos] = a;
            cand_b[pos] = b;
          }
          nq += warp.popc(mi);
          nc += warp.popc(ml);
          if (nq > kQueueCap) nq = kQueueCap;  // unreachable: depth is bounded at build time
        }
        warp.sync();
      }

      // rank sort by key2 (ties by slot): order[rank] = slot
      for (int i = lane; i < nc; i += L) {
        const float ki = cand_a[i].w;
        int rank = 0;
        for (int j = 0; j < nc; j++) {
          const float kj = cand_a[j].w;
          rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
        }
        order[rank] = i;
      }
      warp.sync();

      // ---- C. DRAIN per tile: nearest first; all lanes test the same geom at the same time ------
      lim2max = 0.f;
      for (int t = 0; t < nt && nc > 0; t++) {
        RayState rs = st[t * L + lane];
        const double vec[3] = {rs.vec[0], rs.vec[1], rs.vec[2]};
        const bool active = vec[0] != 0.0 || vec[1] != 0.0 || vec[2] != 0.0;
        if (!warp.any(active)) continue;
        Hit hit;
        hit.dist = rs.dist;
        hit.id = rs.id;
        warp.set_slot(rs.slot);
        RayF rf;
        rayf_setup(rf, pnt, vec);
        float df[3], vlen_up = INFINITY;
        if (active && !odd) unit_dir(vec, df, vlen_up);
        float tmaxf = hit_tmax(hit, tcap);           // search limit in ray units, rounded up
        float lim2 = lane_lim2(hit, tcap, vlen_up);  // the same in squared world units
        for (int i = 0; i < nc; i++) {
          const int slot = order[i];
          const F4 a = cand_a[slot], b = cand_b[slot];
          const bool need = active && a.w <= lim2;
          if (!warp.any(need)) break;  // sorted: nobody can improve on a later candidate either
          const float lo[3] = {a.x, a.y, a.z}, hi[3] = {b.x, b.y, b.z};
          float tn = 0.f;
          const bool h = need && slab_test(rf, lo, hi, tmaxf, tn);
          if (!warp.any(h)) continue;
          const int32_t gt = f32_as_int(b.w);
          leaf_visit(p, pose, gt & 0xffffff, gt >> 24, h, hit, pnt, vec, tcap, warp);
          tmaxf = hit_tmax(hit, tcap);
          lim2 = lane_lim2(hit, tcap, vlen_up);
        }
        rs.dist = hit.dist;
        rs.id = hit.id;
        rs.slot = warp.get_slot();
        st[t * L + lane] = rs;
        if (active) lim2max = fmaxf(lim2max, lane_lim2(hit, tcap, vlen_up));
      }
      nc = 0;
      warp.sync();
      if (nq > 0 && !odd) tc.far2 = warp.max(lim2max);
    }
  }

  // ---- E. epilogue ------------------------------------------------------------------------------
  for (int t = 0; t < nt; t++) {
    const RayState rs = st[t * L + lane];
    Hit hit;
    hit.dist = rs.dist;
    hit.id = rs.id;
    if (rs.r >= 0) epi(rs.r, hit, rs.slot);
  }
}

}  // namespace b2

#include "split_trace.cuh"
#endif
