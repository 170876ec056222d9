// tile_trace.cuh -- warp-cooperative traversal of the wide geom BVH for one tile of rays
// (included at the end of ray_core.cuh; same host/device rules).
//
// One warp = one tile of up to 32 rays that share an origin.
//   1. planes (kept outside the BVH) are tested first: their hit bounds the search distance;
//   2. COLLECT: the warp walks the fanout-32 geom BVH once for the whole tile.  Every round pops one
//      node and tests ALL its children at once, one lane per child slot, against the tile's
//      bounding cone and against the farthest distance any lane still needs; surviving inner
//      children go to a small frontier queue, surviving geoms to a shared-memory candidate list;
//   3. DRAIN: candidates are rank-sorted by their distance from the origin and visited in that
//      order by all lanes together: per-lane conservative slab test, then the closed-form fp64
//      test of the (warp-uniform) geom type.  The loop ends as soon as no lane can improve.
//      If the candidate list fills up before the frontier is empty, it is drained early and the
//      collect continues with the tightened distance bound (incoherent tiles stay correct);
//   4. mesh geoms a lane's ray enters are not traced here: the warp policy appends (ray, mesh
//      geom) pairs to a compacted list and separate mesh kernels trace them with full warps.
#ifndef B2_TILE_TRACE_CUH
#define B2_TILE_TRACE_CUH

namespace b2 {

constexpr int kMaxWideDepth = 5;  // queue: every round pops 1 node and pushes <= 32

// ---------------------------------------------------------------------------------------------
// tile cone vs box (conservative).  The cone (apex o, unit axis a, half-angle theta) contains
// every ray of the tile; a box is kept if its bounding sphere may touch the cone and its nearest
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

// ---------------------------------------------------------------------------------------------
// one ray of a warp tile.  Every lane of the warp must call trace_tile together (lanes without a
// ray pass active = false); all rays of the tile share the origin pnt.
//   wslots : this world's child slots (shared memory when STAGE), wnodes: topology (same)
// ---------------------------------------------------------------------------------------------
template <bool STAGE, class W>
B2_HD Hit trace_tile(const CastParams& p, const WideChild* wslots, const WideNode* wnodes,
                     const GeomPose* pose, bool active, const double pnt[3], const double vec[3],
                     float tcap, W& warp) {
  Hit hit;
  hit.dist = -1;
  hit.id = -1;

  // 1. planes
  for (int i = 0; i < p.nplane; i++) {
    const int g = p.plane_ids[i];
    if (ld_u8(p.elim + g)) continue;
    if (active)
      hit_update(hit,
                 prim_test(GEOM_PLANE, pose + g, p.gstat + g, pnt[0], pnt[1], pnt[2], vec[0],
                           vec[1], vec[2]),
                 g);
  }
  if (p.nwnode == 0 || !warp.any(active)) return hit;

  RayF rf;
  rayf_setup(rf, pnt, vec);

  // 2. tile cone: axis = mean unit direction, half-angle = largest deviation (+ slack)
  TileCone tc;
  float df[3] = {0.f, 0.f, 0.f};
  float vlen_up = 0.f;
  bool odd = false;  // |vec| outside the float range: give up on distance culling for the tile
  if (active) {
    const float s = (float)dot3(vec, vec);
    if (s > 1e-30f && s < 1e30f) {
      const float inv = rsqrt_f32(s);
      df[0] = (float)vec[0] * inv;
      df[1] = (float)vec[1] * inv;
      df[2] = (float)vec[2] * inv;
      vlen_up = s * inv * (1.0f + 4e-6f);
    } else {
      odd = true;
    }
  }
  odd = warp.any(odd);
  const float ax = warp.sum(df[0]), ay = warp.sum(df[1]), az = warp.sum(df[2]);
  const float an = sqrtf(ax * ax + ay * ay + az * az);
  tc.wide = odd || !(an > 1e-3f);
  const float ian = tc.wide ? 0.f : 1.0f / an;
  tc.a[0] = ax * ian;
  tc.a[1] = ay * ian;
  tc.a[2] = az * ian;
  const float cmin = warp.min(active ? df[0] * tc.a[0] + df[1] * tc.a[1] + df[2] * tc.a[2] : 1.0f);
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
  if (odd) vlen_up = INFINITY;
  tc.far2 = warp.max(active ? lane_lim2(hit, tcap, vlen_up) : 0.f);

  int32_t* queue = warp.queue();
  F4* cand_a = warp.cand_a();  // lo.xyz, key2
  F4* cand_b = warp.cand_b();  // hi.xyz, geom id bits
  int32_t* order = warp.order();
  const int L = warp.size();
  int nq = 1, nc = 0;
  if (warp.lane() == 0) queue[0] = 0;
  warp.sync();

  while (nq > 0 || nc > 0) {
    // 3. COLLECT while there is room for one more node's worth of geoms
    while (nq > 0 && nc <= kCandCap - kWideFanout) {
      // one node per round (packing several half-full nodes into one round was measured slower:
      // the bookkeeping cost more than the idle lanes, profiles/r01h)
      const int32_t node = queue[nq - 1];
      warp.sync();  // everybody has read the top before it is overwritten
      nq--;
      const WideNode wn = wnodes[node];
      const int total = wn.count;
      // one lane per child slot (a 1-lane emulated warp walks the slots in turn)
      for (int base = 0; base < total; base += L) {
        const int k = base + warp.lane();
        bool pass = false;
        float key2 = 0.f;
        WideChild ch;
        ch.link = 0;
        if (k < total) {
          ch = ld_wslot<STAGE>(wslots + wn.first + k);
          pass = cone_box(tc, ch.lo, ch.hi, key2);
        }
        const bool inner = pass && ch.link >= 0;
        const bool leaf = pass && ch.link < 0 && !ld_u8(p.elim + (~ch.link));
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
          b.w = int_as_f32(~ch.link);
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
    for (int i = warp.lane(); i < nc; i += L) {
      const float ki = cand_a[i].w;
      int rank = 0;
      for (int j = 0; j < nc; j++) {
        const float kj = cand_a[j].w;
        rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
      }
      order[rank] = i;
    }
    warp.sync();

    // DRAIN: nearest first; all lanes test the same geom at the same time
    for (int i = 0; i < nc; i++) {
      const int slot = order[i];
      const F4 a = cand_a[slot], b = cand_b[slot];
      const bool need = active && a.w <= lane_lim2(hit, tcap, vlen_up);
      if (!warp.any(need)) break;  // sorted: nobody can improve on a later candidate either
      const float lo[3] = {a.x, a.y, a.z}, hi[3] = {b.x, b.y, b.z};
      float tn = 0.f;
      const bool h = need && slab_test(rf, lo, hi, hit_tmax(hit, tcap), tn);
      if (!warp.any(h)) continue;
      leaf_visit(p, pose, f32_as_int(b.w), h, hit, pnt, vec, tcap, warp);
    }
    nc = 0;
    warp.sync();
    if (nq > 0) tc.far2 = warp.max(active ? lane_lim2(hit, tcap, vlen_up) : 0.f);
  }

  return hit;
}

}  // namespace b2
#endif
