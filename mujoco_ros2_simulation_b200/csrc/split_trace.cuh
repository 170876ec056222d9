// split_trace.cuh -- the group traversal of tile_trace.cuh cut in two, so that each half can be
// its own kernel (included at the end of tile_trace.cuh; same host/device rules).
//
// Why: the fused kernel is ~64 KB of SASS and its warps, running de-synchronised in the persistent
// launch, stall on instruction fetch for a fifth to two fifths of their time (ncu r01q-r01s:
// no_instruction 2.0-4.9 per issue).  Cut in two, each kernel's hot code is about half of that:
//   collect_group  (K2a)  rays + planes only as far as the group's bounding cone and farthest search
//                         distance need them, ONE warp-cooperative pass over the fanout-32 geom BVH,
//                         rank sort, and the sorted candidate list goes to global memory -- fp32
//                         apart from the plane test, few registers;
//   drain_tile     (K2b)  per tile: ray, planes, then the candidates nearest first (conservative
//                         slab test, warp-uniform fp64 leaf test) and the fused epilogue -- all
//                         per-ray state stays in registers, no shared memory at all.
// A group the single pass cannot finish (candidate list full, rays outside the float range) is
// reported as -1 and traced by the fused trace_group instead.  Results do not depend on which path
// a group takes: the list K2b walks is exactly the candidate list of trace_group's first pass (same
// cone, same far bound, same order), and every accepted distance comes from the same double test.
#ifndef B2_SPLIT_TRACE_CUH
#define B2_SPLIT_TRACE_CUH

namespace b2 {

B2_HD CandRef ld_cand(const CandRef* ptr) {
#if B2_ON_DEVICE
  const int2 v = __ldg(reinterpret_cast<const int2*>(ptr));
  CandRef c;
  c.key2 = __int_as_float(v.x);
  c.slot = v.y;
  return c;
#else
  return *ptr;
#endif
}

// One group of nt <= kGroupTiles tiles: returns the number of candidates written to out[] (sorted by
// key2, ties by collect order), or -1 if the fused path must trace the group.
//   W supplies: size/lane/ballot/any/popc/lanemask_lt/sync/sum/min/max, queue(), order(),
//   cand_a() (used as the unit-direction buffer) and cref() (kCandCap staging entries).
template <bool STAGE, class W, class Gen>
B2_HD int collect_group(const CastParams& p, const WideChild* wslots, const WideNode* wnodes,
                        const uint8_t* elim, const GeomPose* pose, int nt, const double pnt[3],
                        float tcap, W& warp, const Gen& gen, CandRef* out) {
  const int L = warp.size(), lane = warp.lane();
  int32_t* queue = warp.queue();
  int32_t* order = warp.order();
  CandRef* cref = warp.cref();
  float* dfbuf = reinterpret_cast<float*>(warp.cand_a());

  // ---- A'. rays and planes, as far as the cone and the far bound need them ---------------------
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
        hit_update(hit, plane_test(pose + g, p.gstat + g, pnt, vec), g);
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
    any_active = any_active || active;
  }
  if (p.nwnode == 0 || !warp.any(any_active)) return 0;
  if (warp.any(odd)) return -1;

  // group cone (identical to trace_group's)
  TileCone tc;
  const float ax = warp.sum(sx), ay = warp.sum(sy), az = warp.sum(sz);
  const float an = sqrtf(ax * ax + ay * ay + az * az);
  tc.wide = !(an > 1e-3f);
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
  tc.far2 = warp.max(lim2max);
  warp.sync();

  // ---- B'. one collect pass ---------------------------------------------------------------------
  int nq = 1, nc = 0;
  if (lane == 0) queue[0] = 0;
  warp.sync();
  while (nq > 0) {
    if (nc > kCandCap - kWideFanout) return -1;  // warp-uniform: the list may not hold the next node
    const int32_t node = queue[nq - 1];
    warp.sync();
    nq--;
    const WideNode wn = wnodes[node];
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
        CandRef c;
        c.key2 = key2;
        c.slot = wn.first + k;
        cref[pos] = c;
      }
      nq += warp.popc(mi);
      nc += warp.popc(ml);
      if (nq > kQueueCap) nq = kQueueCap;  // unreachable: depth is bounded at build time
    }
    warp.sync();
  }

  // rank sort by key2 (ties by position), written out in order
  for (int i = lane; i < nc; i += L) {
    const float ki = cref[i].key2;
    int rank = 0;
    for (int j = 0; j < nc; j++) {
      const float kj = cref[j].key2;
      rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    order[rank] = i;
  }
  warp.sync();
  for (int i = lane; i < nc; i += L) out[i] = cref[order[i]];
  warp.sync();
  return nc;
}

// Copies a group's candidates (list entry + the slot's box, geom id and type) to the warp's shared
// arrays, one candidate per lane.  All lanes call it; the caller syncs the warp before and after.
template <class W>
B2_HD void stage_candidates(W& warp, const WideChild* wslots, int count, const CandRef* list, F4* ca,
                            F4* cb) {
  for (int i = warp.lane(); i < count; i += warp.size()) {
    const CandRef c = ld_cand(list + i);
    const WideChild ch = ld_wslot<false>(wslots + c.slot);
    F4 a, b;
    a.x = ch.lo[0];
    a.y = ch.lo[1];
    a.z = ch.lo[2];
    a.w = c.key2;
    b.x = ch.hi[0];
    b.y = ch.hi[1];
    b.z = ch.hi[2];
    b.w = int_as_f32((~ch.link) | (ch.pad << 24));
    ca[i] = a;
    cb[i] = b;
  }
}

// One tile of a group whose candidate list is list[0..count).  wslots: the world's child slots in
// GLOBAL memory.  Every lane of the warp must call drain_tile together.
//   STAGED: the group's candidates were copied to shared memory by stage_candidates (ca/cb, the
//   fused kernel's layout: lo.xyz + key2, hi.xyz + geom id | type << 24); else list + wslots are read.
template <bool STAGED, class W, class Gen, class Epi>
B2_HD void drain_tile(const CastParams& p, const WideChild* wslots, const uint8_t* elim,
                      const GeomPose* pose, int t, const double pnt[3], float tcap, W& warp,
                      const Gen& gen, const Epi& epi, int count, const CandRef* list,
                      const F4* ca, const F4* cb) {
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
      hit_update(hit, plane_test(pose + g, p.gstat + g, pnt, vec), g);
  }
  warp.set_slot(-1);
  if (count > 0 && warp.any(active)) {
    if (!active) vec[0] = vec[1] = vec[2] = 0.0;
    RayF rf;
    rayf_setup(rf, pnt, vec);
    float df[3], vlen_up = INFINITY;
    if (active && !unit_dir(vec, df, vlen_up)) vlen_up = INFINITY;
    float tmaxf = hit_tmax(hit, tcap);
    float lim2 = lane_lim2(hit, tcap, vlen_up);
    for (int i = 0; i < count; i++) {
      float key2, lo[3], hi[3];
      int32_t geom, type;
      if (STAGED) {
        const F4 a = ca[i], b = cb[i];
        key2 = a.w;
        lo[0] = a.x, lo[1] = a.y, lo[2] = a.z;
        hi[0] = b.x, hi[1] = b.y, hi[2] = b.z;
        const int32_t gt = f32_as_int(b.w);
        geom = gt & 0xffffff;
        type = gt >> 24;
      } else {
        const CandRef c = ld_cand(list + i);
        key2 = c.key2;
        if (!warp.any(active && key2 <= lim2)) break;
        const WideChild ch = ld_wslot<false>(wslots + c.slot);
        lo[0] = ch.lo[0], lo[1] = ch.lo[1], lo[2] = ch.lo[2];
        hi[0] = ch.hi[0], hi[1] = ch.hi[1], hi[2] = ch.hi[2];
        geom = ~ch.link;
        type = ch.pad;
      }
      const bool need = active && key2 <= lim2;
      if (!warp.any(need)) break;  // sorted: nobody can improve on a later candidate either
      float tn = 0.f;
      const bool h = need && slab_test(rf, lo, hi, tmaxf, tn);
      if (!warp.any(h)) continue;
      leaf_visit(p, pose, geom, type, h, hit, pnt, vec, tcap, warp);
      tmaxf = hit_tmax(hit, tcap);
      lim2 = lane_lim2(hit, tcap, vlen_up);
    }
  }
  if (r >= 0) epi(r, hit, warp.get_slot());
}

}  // namespace b2
#endif
