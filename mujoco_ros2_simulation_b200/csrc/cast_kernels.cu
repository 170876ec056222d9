// cast_kernels.cu -- K2: the ray-cast kernels (sm_100a).
//
//   K2a collect_kernel     one warp per group of 4 tiles x 32 rays: planes, bounding cone, ONE
//                          warp-cooperative pass over the fanout-32 geom BVH, sorted candidate list
//                          to global memory (split_trace.cuh collect_group).
//   K2b drain_kernel       persistent, one group per fetch: per tile the candidates nearest first --
//                          fp32 slab test, closed-form fp64 leaf test of the warp-uniform geom type --
//                          and the fused LaserScan / multiRay / GL-depth epilogue (drain_tile).
//                          Mesh geoms are not traced here: each (ray, mesh geom) whose box the ray
//                          enters becomes a PAIR in a compacted global list (SPLIT = true).
//       fallback_kernel    the groups K2a could not finish in one pass: fused trace_group.
//   K2  cast_kernel        the fused form (tile_trace.cuh trace_group: collect -> drain -> continue):
//                          launches below ~1 M rays, scenes without a geom BVH, and the clean-up
//                          pass that recomputes a chunk in place when the pair list overflowed.
//   K2c mesh_pairs_kernel  one thread per pair: mj_rayMesh through the mesh's fanout-4 tree,
//                          persistent while-while with refill; distances merge with a 64-bit atomicMin.
//   K2d mesh_ids_kernel    lowest mesh geom id among the pairs that reached the minimum.
//   K2e mesh_final_kernel  re-runs the epilogue for the rays a mesh changed.
//
// Compile with -fmad=false: distances must be bit-identical to the CPU oracle (ray_core.cuh).
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdlib>

#include "kernels.cuh"
#include "ray_core.cuh"

namespace b2 {

constexpr int kCastThreads = 256;
constexpr int kCastWarps = kCastThreads / 32;
constexpr size_t kStageLimit = 96 * 1024;  // largest geom BVH staged in shared memory per CTA
constexpr int kMeshThreads = 128;

int cast_tiles_per_scan(int nray) { return (nray + 31) / 32; }
int cast_threads_per_block() { return kCastThreads; }
size_t cast_stage_limit_bytes() { return kStageLimit; }
// child slots + topology + the sensor's ray_eliminate flags (one byte per geom)
size_t cast_stage_bytes(int nslot, int nwnode, int ngeom) {
  const size_t bvh = (size_t)nslot * sizeof(WideChild) + (size_t)nwnode * sizeof(WideNode);
  return ((bvh + 15) & ~(size_t)15) + (((size_t)ngeom + 15) & ~(size_t)15);
}

// per-warp shared-memory scratch of the cast kernel
struct WarpScratch {
  F4 cand_a[kCandA];         // candidate box lo.xyz + distance key
  F4 cand_b[kCandCap];       // candidate box hi.xyz + geom id bits
  RayState st[kGroupTiles * 32];  // per-ray state of the group between the phases
  int32_t queue[kQueueCap];  // collect frontier
  int32_t order[kCandCap];   // rank-sorted candidate slots
};

template <bool SPLIT>
struct DevWarp {
  WarpScratch* s;
  int lane_;
  int slot;  // affected-record slot of this lane's ray in the current tile (-1: none yet)
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ int lane() const { return lane_; }
  __device__ __forceinline__ unsigned lanemask_lt() const { return (1u << lane_) - 1u; }
  __device__ __forceinline__ unsigned ballot(bool pred) const {
    return __ballot_sync(0xffffffffu, pred);
  }
  __device__ __forceinline__ bool any(bool pred) const { return __any_sync(0xffffffffu, pred); }
  __device__ __forceinline__ int popc(unsigned m) const { return __popc(m); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ float sum(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ float min(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ float max(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ int32_t* queue() const { return s->queue; }
  __device__ __forceinline__ F4* cand_a() const { return s->cand_a; }
  __device__ __forceinline__ F4* cand_b() const { return s->cand_b; }
  __device__ __forceinline__ int32_t* order() const { return s->order; }
  __device__ __forceinline__ RayState* ray_state() const { return s->st; }
  __device__ __forceinline__ void set_slot(int v) { slot = v; }
  __device__ __forceinline__ int get_slot() const { return slot; }

  // a mesh geom some lane's ray enters (called by all lanes, g warp-uniform)
  __device__ __forceinline__ void mesh_pair(const CastParams& p, const GeomPose* pose, int g,
                                            bool h, Hit& hit, const double pnt[3],
                                            const double vec[3], float tcap) {
    if (!SPLIT) {
      mesh_inline(p, pose, g, h, hit, pnt, vec, tcap);
      return;
    }
    // mj_rayMesh's own pre-test (ray vs the geom_size box, fp64): a ray that misses it returns -1
    // upstream, so it needs no pair.  Roughly halves the pair list (the world AABB of a rotated
    // box is loose) and all lanes test the same geom here.
    // before that, the mesh's bounding sphere (sign only: no division, no root): for anything roundish
    // it rejects what the box lets through
    if (h) h = !mesh_bsphere_miss(p, pose + g, p.gstat + g, pnt, vec);
    if (h)
      h = prim_test(GEOM_BOX, pose + g, p.gstat + g, pnt[0], pnt[1], pnt[2], vec[0], vec[1],
                    vec[2]) >= 0;
    const unsigned m = ballot(h);
    if (m == 0u) return;
    const int leader = __ffs(m) - 1;
    // first pair of a ray: reserve its affected-record slot (warp-aggregated atomics)
    const unsigned mnew = ballot(h && slot < 0);
    int abase = 0, pbase = 0;
    if (lane_ == leader) {
      if (mnew) abase = atomicAdd(p.ms.counters + 1, __popc(mnew));
      pbase = atomicAdd(p.ms.counters + 0, __popc(m));
    }
    abase = __shfl_sync(0xffffffffu, abase, leader);
    pbase = __shfl_sync(0xffffffffu, pbase, leader);
    if (h) {
      if (slot < 0) slot = abase + __popc(mnew & lanemask_lt());
      const int idx = pbase + __popc(m & lanemask_lt());
      if (idx < p.ms.pair_cap && slot < p.ms.aff_cap) {
        p.ms.pair_slot[idx] = slot;
        p.ms.pair_geom[idx] = g;
      } else {
        p.ms.counters[2] = 1;  // overflow: the clean-up launch recomputes this chunk in place
      }
    }
  }
};

// ray index of (tile, lane): 8x4 tiles over (scan line, line index) for 3-D sensors, else 32x1
__device__ __forceinline__ int tile_ray(int tile, int lane, int tiles_x, int tile_w) {
  if (tiles_x > 0) {
    const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    return (ty * 4 + (lane >> 3)) * tile_w + tx * 8 + (lane & 7);
  }
  return tile * 32 + lane;
}

// the scan's sensor frame and ray r of it (same arithmetic in K2 and in the mesh kernels)
template <int MODE>
__device__ __forceinline__ bool scan_ray(const CastParams& p, int scan, int r, double o[3],
                                         double l[3], double vec[3]) {
  double R[9];
#pragma unroll
  for (int k = 0; k < 3; k++) o[k] = __ldg(p.spos + (size_t)scan * 3 + k);
  if (MODE != MODE_MULTIRAY) {
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = __ldg(p.smat + (size_t)scan * 9 + k);
  }
  return cast_make_ray(p, MODE, R, r, l, vec);
}

#ifndef B2_CAST_MINBLOCKS
#define B2_CAST_MINBLOCKS 2
#endif

struct TileGeom {
  int tiles_per_block, tiles_per_scan, chunks_per_scan, tiles_x;
  // persistent launch (single-world batches): warps fetch tile groups from p.work_counter
  int dynamic, groups_per_scan, ngroups;
  int group_tiles;  // tiles per group of this sensor (1, 2 or kGroupTiles = 4)
  // K2b writes a group's LaserScan results as float4 / int4 (one 128-B line per scan-line segment):
  // 3-D sensors whose groups are 4 tiles of one tile row, ranges / geomid the only outputs, 16-B
  // aligned (decided at launch)
  int vec_out;
};

// ray generator of one group: tile t of the group, this lane's ray.  The sensor frame is
// warp-uniform and L1-resident; it is reloaded on every call so that nothing but the ray itself
// is live across the traversal.
template <int MODE>
struct GroupGen {
  const CastParams& p;
  const TileGeom& tg;
  int scan, tile0, lane;
  __device__ __forceinline__ bool operator()(int t, double vec[3], int& r) const {
    const int rr = tile_ray(tile0 + t, lane, tg.tiles_x, p.tile_w);
    const bool in = rr < p.nray;
    r = in ? rr : -1;
    double o[3], l[3];
    return scan_ray<MODE>(p, scan, in ? rr : 0, o, l, vec) && in;
  }
};

// result sink of one group: the split path's affected record, then the fused epilogue
// K2b with vec_out: s_rng / s_id point at the warp's shared-memory patch and the tile's LaserScan
// results go there (tile-major, [tile][scan line][8 azimuths]) instead of to global memory; the warp
// writes the whole 32 x 4 patch after its fourth tile (store_patch).
template <int MODE, bool SPLIT>
struct GroupEpi {
  const CastParams& p;
  const TileGeom& tg;
  int scan, tile0, lane;
  float* s_rng = nullptr;    // [128], or nullptr: write to global memory
  int32_t* s_id = nullptr;   // [128]
  int t = 0;                 // tile of the group being drained (vec_out)
  __device__ __forceinline__ void operator()(int r, const Hit& hit, int slot) const {
    if (SPLIT && slot >= 0 && slot < p.ms.aff_cap) {
      const unsigned long long bits =
          hit.dist < 0 ? kInfBits : (unsigned long long)__double_as_longlong(hit.dist);
      p.ms.aff_ray[slot] = (uint32_t)((size_t)scan * p.nray + r);
      p.ms.aff_bits[slot] = bits;
      p.ms.aff_prim[slot] = bits;
      p.ms.aff_prim_id[slot] = hit.id;
      p.ms.aff_mesh_id[slot] = INT32_MAX;
    }
    if (MODE == MODE_LIDAR && s_rng != nullptr) {
      double val;
      bool filtered;
      s_rng[t * 32 + lane] = lidar_range(p, hit.dist, val, filtered);
      s_id[t * 32 + lane] = hit.id;
      return;
    }
    double le[3] = {0.0, 0.0, 0.0};
    if (MODE == MODE_LIDAR && p.xyz != nullptr) {
#pragma unroll
      for (int k = 0; k < 3; k++) le[k] = __ldg(p.local_dir + (size_t)r * 3 + k);
    }
    cast_epilogue(p, MODE, scan, r, hit, le);
  }
};

// the 32 x 4 patch of a group (4 tiles of 8 x 4 rays side by side): lane = (scan line, quad of 4
// azimuths) reads its quad from the tile that holds it and writes one float4 + one int4 -- each
// scan-line segment of the patch is one full 128-B line per output array
__device__ __forceinline__ void store_patch(const CastParams& p, const TileGeom& tg, int scan,
                                            int tile0, int lane, const float* s_rng,
                                            const int32_t* s_id) {
  const int line = lane >> 3, q = lane & 7;
  const int src = (q >> 1) * 32 + line * 8 + (q & 1) * 4;
  const int ty = tile0 / tg.tiles_x, tx0 = tile0 - ty * tg.tiles_x;
  const size_t idx = (size_t)scan * p.nray + (size_t)(ty * 4 + line) * p.tile_w + tx0 * 8 + q * 4;
  if (p.ranges)
    *reinterpret_cast<float4*>(p.ranges + idx) = *reinterpret_cast<const float4*>(s_rng + src);
  if (p.geomid)
    *reinterpret_cast<int4*>(p.geomid + idx) = *reinterpret_cast<const int4*>(s_id + src);
  if (p.geomid16) {
    const int4 v = *reinterpret_cast<const int4*>(s_id + src);
    uint2 w;
    w.x = ((uint32_t)v.x & 0xffffu) | ((uint32_t)v.y << 16);
    w.y = ((uint32_t)v.z & 0xffffu) | ((uint32_t)v.w << 16);
    *reinterpret_cast<uint2*>(p.geomid16 + idx) = w;
  }
}

#ifdef B2_MESH_STATS
__device__ unsigned long long g_tile_stats[32];
extern "C" int b2_debug_tile_stats(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_tile_stats, sizeof(g_tile_stats));
  if (reset) {
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(g_tile_stats, z, sizeof(z));
  }
  return 0;
}
#endif
// ONLY_IF_OVERFLOW: the clean-up launch -- exits at once unless the split pass overflowed
template <int MODE, bool STAGE, bool SPLIT, bool ONLY_IF_OVERFLOW>
__global__ void __launch_bounds__(kCastThreads, B2_CAST_MINBLOCKS)
    cast_kernel(const __grid_constant__ CastParams p, const TileGeom tg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  if (ONLY_IF_OVERFLOW) {
    if (p.ms.counters[2] == 0) return;
  }
  // dynamic == 2: the fall-back of the two-kernel cast -- only the groups K2a listed (often none)
  if (tg.dynamic == 2 && p.cl.counters[0] == 0) return;
  WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem_raw);
  unsigned char* s_stage = smem_raw + kCastWarps * sizeof(WarpScratch);

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a dynamic launch serves one world only (the launcher checks): stage the world of scan 0
  const int scan_b = tg.dynamic ? 0 : blockIdx.x / tg.chunks_per_scan;
  const int chunk = blockIdx.x - scan_b * tg.chunks_per_scan;
  const int world = p.scan_world ? p.scan_world[scan_b] : (p.scan0 + scan_b) % p.nworld;

  // geom BVH of this world: child slots (boxes) + topology, staged in shared memory per CTA
  const WideChild* wslots = p.wslots + (size_t)world * p.nslot;
  const WideNode* wnodes = p.wnodes;
  const uint8_t* elim = p.elim;
  if (STAGE) {
    const float4* src = reinterpret_cast<const float4*>(wslots);
    float4* dst = reinterpret_cast<float4*>(s_stage);
    for (int i = threadIdx.x; i < p.nslot * 2; i += kCastThreads) dst[i] = __ldg(src + i);
    int2* ndst = reinterpret_cast<int2*>(s_stage + (size_t)p.nslot * sizeof(WideChild));
    const int2* nsrc = reinterpret_cast<const int2*>(p.wnodes);
    for (int i = threadIdx.x; i < p.nwnode; i += kCastThreads) ndst[i] = __ldg(nsrc + i);
    // the sensor's ray_eliminate flags: read once per collected geom, keep them out of L2 latency
    const size_t bvh = (size_t)p.nslot * sizeof(WideChild) + (size_t)p.nwnode * sizeof(WideNode);
    uint8_t* edst = s_stage + ((bvh + 15) & ~(size_t)15);
    for (int i = threadIdx.x; i < p.ngeom; i += kCastThreads) edst[i] = __ldg(p.elim + i);
    __syncthreads();
    wslots = reinterpret_cast<const WideChild*>(s_stage);
    wnodes = reinterpret_cast<const WideNode*>(ndst);
    elim = edst;
  }
  const GeomPose* pose = p.pose + (size_t)world * p.ngeom;

  DevWarp<SPLIT> warp{scratch + warp_id, lane, -1};
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;

  // each warp walks groups of kGroupTiles consecutive tiles of this CTA's chunk
  // (static launch) or, in a persistent launch, whatever group the global work counter hands out
  // next -- no warp waits for the slowest group of its CTA
  for (int t = warp_id * tg.group_tiles;; t += kCastWarps * tg.group_tiles) {
    int scan, tile0, nt;
    if (tg.dynamic) {
      int g = 0;
      if (tg.dynamic == 2) {
        if (lane == 0) g = atomicAdd(p.cl.counters + 2, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= p.cl.counters[0]) break;
        g = p.cl.ovf_list[g];
      } else {
        if (lane == 0) g = atomicAdd(p.work_counter, 1);
        g = __shfl_sync(0xffffffffu, g, 0);
        if (g >= tg.ngroups) break;
      }
      scan = g / tg.groups_per_scan;
      tile0 = (g - scan * tg.groups_per_scan) * tg.group_tiles;
      nt = ::min(tg.group_tiles, tg.tiles_per_scan - tile0);
    } else {
      if (t >= tg.tiles_per_block) break;
      scan = scan_b;
      tile0 = chunk * tg.tiles_per_block + t;
      if (tile0 >= tg.tiles_per_scan) break;  // warp-uniform
      nt = ::min(tg.group_tiles, ::min(tg.tiles_per_block - t, tg.tiles_per_scan - tile0));
    }
    double o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = __ldg(p.spos + (size_t)scan * 3 + k);
    const GroupGen<MODE> gen{p, tg, scan, tile0, lane};
    const GroupEpi<MODE, SPLIT> epi{p, tg, scan, tile0, lane};
#ifdef B2_MESH_STATS
    const long long st_t0 = clock64();
#endif
    trace_group<STAGE>(p, wslots, wnodes, elim, pose, nt, o, tcap, warp, gen, epi);
#ifdef B2_MESH_STATS
    if (lane == 0 && !ONLY_IF_OVERFLOW) {  // per tile group of the fused kernel (tools/mesh_stats.py --tiles)
      const unsigned long long cyc = (unsigned long long)(clock64() - st_t0);
      atomicAdd(g_tile_stats + 0, 1ull);
      atomicAdd(g_tile_stats + 1, cyc);
      atomicMax(g_tile_stats + 2, cyc);
      atomicMax(g_tile_stats + 3, (unsigned long long)(clock64() & 0xffffffffffffll));
      atomicAdd(g_tile_stats + 16 + ::min(15, ::max(0, 64 - __clzll(cyc) - 8)), 1ull);  // log2 histogram of cycles / 256
    }
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// Two-kernel cast (split_trace.cuh): K2a collect_kernel -> K2b drain_kernel (+ fused fall-back)
// ---------------------------------------------------------------------------------------------
#ifndef B2_COLLECT_MINBLOCKS
#define B2_COLLECT_MINBLOCKS 3
#endif
// CTAs/SM of K2b (register cap): 2 / 3 / 4 / 5 = 126 / 80 / 64 / 48 registers measured 7.91 / 8.36 /
// 8.42 / 7.82 Grays/s -- without shared memory the L1 absorbs the spills and the extra warps hide
// the fp64 latency (the fused kernel could not make this trade: profiles/r01_sweeps.txt)
#ifndef B2_DRAIN_MINBLOCKS
#define B2_DRAIN_MINBLOCKS 4
#endif

struct CollectScratch {
  F4 cand_a[kCandA];  // unit directions of the group's rays while the cone is built
  CandRef cref[kCandCap];
  int32_t queue[kQueueCap];
  int32_t order[kCandCap];
};

struct CollectWarp {
  CollectScratch* s;
  int lane_;
  __device__ __forceinline__ int size() const { return 32; }
  __device__ __forceinline__ int lane() const { return lane_; }
  __device__ __forceinline__ unsigned lanemask_lt() const { return (1u << lane_) - 1u; }
  __device__ __forceinline__ unsigned ballot(bool pred) const {
    return __ballot_sync(0xffffffffu, pred);
  }
  __device__ __forceinline__ bool any(bool pred) const { return __any_sync(0xffffffffu, pred); }
  __device__ __forceinline__ int popc(unsigned m) const { return __popc(m); }
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ float sum(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  __device__ __forceinline__ float min(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ float max(float v) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
  }
  __device__ __forceinline__ int32_t* queue() const { return s->queue; }
  __device__ __forceinline__ F4* cand_a() const { return s->cand_a; }
  __device__ __forceinline__ CandRef* cref() const { return s->cref; }
  __device__ __forceinline__ int32_t* order() const { return s->order; }
};

// K2a: one warp per group of tiles; same tiling / persistent group fetch as cast_kernel
template <int MODE, bool STAGE>
__global__ void __launch_bounds__(kCastThreads, B2_COLLECT_MINBLOCKS)
    collect_kernel(const __grid_constant__ CastParams p, const TileGeom tg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CollectScratch* scratch = reinterpret_cast<CollectScratch*>(smem_raw);
  unsigned char* s_stage = smem_raw + kCastWarps * sizeof(CollectScratch);
  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int scan_b = tg.dynamic ? 0 : blockIdx.x / tg.chunks_per_scan;
  const int chunk = blockIdx.x - scan_b * tg.chunks_per_scan;
  const int world = p.scan_world ? p.scan_world[scan_b] : (p.scan0 + scan_b) % p.nworld;
  const WideChild* wslots = p.wslots + (size_t)world * p.nslot;
  const WideNode* wnodes = p.wnodes;
  const uint8_t* elim = p.elim;
  if (STAGE) {
    const float4* src = reinterpret_cast<const float4*>(wslots);
    float4* dst = reinterpret_cast<float4*>(s_stage);
    for (int i = threadIdx.x; i < p.nslot * 2; i += kCastThreads) dst[i] = __ldg(src + i);
    int2* ndst = reinterpret_cast<int2*>(s_stage + (size_t)p.nslot * sizeof(WideChild));
    const int2* nsrc = reinterpret_cast<const int2*>(p.wnodes);
    for (int i = threadIdx.x; i < p.nwnode; i += kCastThreads) ndst[i] = __ldg(nsrc + i);
    const size_t bvh = (size_t)p.nslot * sizeof(WideChild) + (size_t)p.nwnode * sizeof(WideNode);
    uint8_t* edst = s_stage + ((bvh + 15) & ~(size_t)15);
    for (int i = threadIdx.x; i < p.ngeom; i += kCastThreads) edst[i] = __ldg(p.elim + i);
    __syncthreads();
    wslots = reinterpret_cast<const WideChild*>(s_stage);
    wnodes = reinterpret_cast<const WideNode*>(ndst);
    elim = edst;
  }
  const GeomPose* pose = p.pose + (size_t)world * p.ngeom;
  CollectWarp warp{scratch + warp_id, lane};
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;

  for (int t = warp_id * tg.group_tiles;; t += kCastWarps * tg.group_tiles) {
    int scan, tile0, nt;
    if (tg.dynamic) {
      int g = 0;
      if (lane == 0) g = atomicAdd(p.work_counter, 1);
      g = __shfl_sync(0xffffffffu, g, 0);
      if (g >= tg.ngroups) break;
      scan = g / tg.groups_per_scan;
      tile0 = (g - scan * tg.groups_per_scan) * tg.group_tiles;
      nt = ::min(tg.group_tiles, tg.tiles_per_scan - tile0);
    } else {
      if (t >= tg.tiles_per_block) break;
      scan = scan_b;
      tile0 = chunk * tg.tiles_per_block + t;
      if (tile0 >= tg.tiles_per_scan) break;  // warp-uniform
      nt = ::min(tg.group_tiles, ::min(tg.tiles_per_block - t, tg.tiles_per_scan - tile0));
    }
    double o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = __ldg(p.spos + (size_t)scan * 3 + k);
    const GroupGen<MODE> gen{p, tg, scan, tile0, lane};
    const int gid = scan * tg.groups_per_scan + tile0 / tg.group_tiles;
    const int n = collect_group<STAGE>(p, wslots, wnodes, elim, pose, nt, o, tcap, warp, gen,
                                       p.cl.list + (size_t)gid * kCandCap);
    if (lane == 0) {
      p.cl.count[gid] = n;
      if (n < 0) p.cl.ovf_list[atomicAdd(p.cl.counters + 0, 1)] = gid;
    }
  }
}

// K2b: persistent; a warp takes one group of tiles per fetch from a global cursor (its tiles share
// the candidate list, which then stays in L1: per-tile fetch 8.41, per-group 9.62 Grays/s) and reads
// the list and the slots through L1 -- no shared memory.  B2_DRAIN_STAGED=1 copies the group's
// candidates to 3 KB of shared memory per warp first, as the fused kernel does: measured slower
// (8.69), the carve-out costs more L1 than the copies save.
#ifndef B2_DRAIN_STAGED
#define B2_DRAIN_STAGED 0
#endif
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(kCastThreads, B2_DRAIN_MINBLOCKS)
    drain_kernel(const __grid_constant__ CastParams p, const TileGeom tg) {
#if B2_DRAIN_STAGED
  __shared__ F4 s_ca[kCastWarps][kCandCap];
  __shared__ F4 s_cb[kCastWarps][kCandCap];
  F4* ca = s_ca[threadIdx.x >> 5];
  F4* cb = s_cb[threadIdx.x >> 5];
#else
  F4* ca = nullptr;
  F4* cb = nullptr;
#endif
  const int lane = threadIdx.x & 31;
  // result patch of the group a warp is draining (vec_out): 128 ranges + 128 geom ids
  __shared__ __align__(16) float s_rng[kCastWarps][kGroupTiles * 32];
  __shared__ __align__(16) int32_t s_id[kCastWarps][kGroupTiles * 32];
  float* my_rng = s_rng[threadIdx.x >> 5];
  int32_t* my_id = s_id[threadIdx.x >> 5];
  DevWarp<SPLIT> warp{nullptr, lane, -1};
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;
  for (;;) {
    int gid = 0;
    if (lane == 0) gid = atomicAdd(p.cl.counters + 1, 1);
    gid = __shfl_sync(0xffffffffu, gid, 0);
    if (gid >= tg.ngroups) break;
    const int count = __ldg(p.cl.count + gid);
    if (count < 0) continue;  // left to the fused kernel
    const int scan = gid / tg.groups_per_scan;
    const int tile0 = (gid - scan * tg.groups_per_scan) * tg.group_tiles;
    const int nt = ::min(tg.group_tiles, tg.tiles_per_scan - tile0);
    const int world = p.scan_world ? p.scan_world[scan] : (p.scan0 + scan) % p.nworld;
    const WideChild* wslots = p.wslots + (size_t)world * p.nslot;
    const GeomPose* pose = p.pose + (size_t)world * p.ngeom;
    const CandRef* list = p.cl.list + (size_t)gid * kCandCap;
    double o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = __ldg(p.spos + (size_t)scan * 3 + k);
    const GroupGen<MODE> gen{p, tg, scan, tile0, lane};
    GroupEpi<MODE, SPLIT> epi{p, tg, scan, tile0, lane};
#if B2_DRAIN_STAGED
    __syncwarp();  // the previous group's tiles are done with ca / cb
    stage_candidates(warp, wslots, count, list, ca, cb);
    __syncwarp();
#endif
    const bool vec = MODE == MODE_LIDAR && tg.vec_out && nt == kGroupTiles;
    if (vec) {
      epi.s_rng = my_rng;
      epi.s_id = my_id;
    }
#pragma unroll 1
    for (int t = 0; t < nt; t++) {
      epi.t = t;
      drain_tile<B2_DRAIN_STAGED != 0>(p, wslots, p.elim, pose, t, o, tcap, warp, gen, epi, count, list,
                                        ca, cb);
    }
    if (vec) {
      __syncwarp();
      store_patch(p, tg, scan, tile0, lane, my_rng, my_id);
      __syncwarp();  // the patch is reused by this warp's next group
    }
  }
}

// the groups K2a could not finish: fused trace_group, unstaged, one warp per listed group
template <int MODE, bool SPLIT>
__global__ void __launch_bounds__(kCastThreads, B2_CAST_MINBLOCKS)
    fallback_kernel(const __grid_constant__ CastParams p, const TileGeom tg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int novf = p.cl.counters[0];
  if (novf == 0) return;
  WarpScratch* scratch = reinterpret_cast<WarpScratch*>(smem_raw);
  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  DevWarp<SPLIT> warp{scratch + warp_id, lane, -1};
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;
  for (;;) {
    int idx = 0;
    if (lane == 0) idx = atomicAdd(p.cl.counters + 2, 1);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= novf) break;
    const int gid = p.cl.ovf_list[idx];
    const int scan = gid / tg.groups_per_scan;
    const int tile0 = (gid - scan * tg.groups_per_scan) * tg.group_tiles;
    const int nt = ::min(tg.group_tiles, tg.tiles_per_scan - tile0);
    const int world = p.scan_world ? p.scan_world[scan] : (p.scan0 + scan) % p.nworld;
    double o[3];
#pragma unroll
    for (int k = 0; k < 3; k++) o[k] = __ldg(p.spos + (size_t)scan * 3 + k);
    const GroupGen<MODE> gen{p, tg, scan, tile0, lane};
    const GroupEpi<MODE, SPLIT> epi{p, tg, scan, tile0, lane};
    trace_group<false>(p, p.wslots + (size_t)world * p.nslot, p.wnodes, p.elim,
                       p.pose + (size_t)world * p.ngeom, nt, o, tcap, warp, gen, epi);
  }
}

// ---------------------------------------------------------------------------------------------
// K2b: one thread per (ray, mesh geom) pair
// ---------------------------------------------------------------------------------------------
//
// Persistent "while-while" traversal (Aila & Laine style): every lane owns one pair at a time and
// refills itself from a global work counter when its pair is finished, so lanes with short
// traversals do not idle behind long ones.  Each outer round runs (1) refill, (2) up to six nodes
// of the mesh's fanout-4 tree per lane, stopping at a leaf, (3) the fp64 triangle tests of the lanes
// that stand on a leaf -- so the expensive triangle code runs with many lanes active instead of
// whenever a single lane happens to hit a leaf.
// The pair already passed mj_rayMesh's box pre-test in K2b, so no post-test is needed here.
#ifndef B2_MESH_MINBLOCKS
// CTAs/SM (register cap): 4 / 5 / 6 measured 5.44 / 5.50 / 5.47 Grays/s early on; later, with the
// kernel bound by node-fetch latency, 5 / 6 / 8 / 10 gave 6.57 / 6.52 / 6.64 / 6.57 (64 regs, 32 warps)
#define B2_MESH_MINBLOCKS 8
#endif
#ifndef B2_MESH_INNER_CAP
#define B2_MESH_INNER_CAP 6  // inner nodes a lane may descend before the warp serves the leaves
#endif
#ifndef B2_MESH_REFILL
#define B2_MESH_REFILL 16  // idle lanes per warp that trigger a refill (2..16 measured: 16 best)
#endif
#ifndef B2_MESH_SMEM_STACK
// stack entries per lane kept in shared memory (deeper ones in local memory).  0 = all in local memory:
// 12 measured 10.25 Grays/s against 10.34 for 0 -- the stack was never what the L1 misses were about
#define B2_MESH_SMEM_STACK 0
#endif
#ifndef B2_MESH_PREFETCH
// prefetch.global.L1 of (bit 0) a leaf's triangles when the descent reaches it, (bit 1) the runner-up
// child.  Off: 0 / 1 / 2 / 3 measured 10.25 / 10.01 / 10.08 / 9.89 Grays/s -- the kernel waits on the
// node it needs NOW; hints for later nodes only add instructions and evict lines
#define B2_MESH_PREFETCH 0
#endif
#ifdef B2_MESH_STATS
// debug build only: per-pair work of the mesh kernel (tools/mesh_stats.py)
__device__ unsigned long long g_mesh_stats[64];
extern "C" int b2_debug_mesh_stats(unsigned long long* out, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_mesh_stats, sizeof(g_mesh_stats));
  if (reset) {
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(g_mesh_stats, z, sizeof(z));
  }
  return 0;
}
#endif
__device__ __forceinline__ void mesh_prefetch(const void* ptr) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(ptr));
}
template <int MODE>
__global__ void __launch_bounds__(kMeshThreads, B2_MESH_MINBLOCKS)
    mesh_pairs_kernel(const __grid_constant__ CastParams p) {
  const int npair = ::min(p.ms.counters[0], p.ms.pair_cap);
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  constexpr int32_t kDone = kEmptyChild;
  // (Spreading a small launch's pairs over all resident warps -- 3 lanes per warp instead of 32 for
  // the ~11 k pairs of one scan -- did not shorten it: 94 vs 114 us under ncu, 160 us per scan either
  // way in graph replay.  A pair is a ~1400-instruction dependent chain; profiles/r02_sweeps.txt r02y.)
  bool have = false;
  int pair = 0, slot = 0;
  double lpnt[3], lvec[3], b0[3], b1[3], x = -1;
  RayF rf;
  float tmax = 0.f;
  const MeshNode4* nodes = nullptr;
  const MeshTri* tris = nullptr;
  int32_t cur = kDone;
  // traversal stack of (link, entry distance): the first entries of every lane in shared memory
  // ([depth][thread]: conflict-free), the rest -- deep trees only -- in local memory
  __shared__ uint2 s_stack[B2_MESH_SMEM_STACK > 0 ? B2_MESH_SMEM_STACK : 1][kMeshThreads];
  uint2 l_stack[kMeshStack - B2_MESH_SMEM_STACK];
  int sp = 0;
#ifdef B2_MESH_STATS
  unsigned st_nodes = 0, st_tris = 0;
  long long st_t0 = 0;
#endif
  auto push = [&](int32_t link, float t) {
    const uint2 e = make_uint2((uint32_t)link, __float_as_uint(t));
    if (sp < B2_MESH_SMEM_STACK) s_stack[sp][threadIdx.x] = e;
    else l_stack[sp - B2_MESH_SMEM_STACK] = e;
    sp++;
  };
  // next stacked node that can still beat tmax, or kDone
  auto pop = [&]() -> int32_t {
    while (sp > 0) {
      --sp;
      const uint2 e = sp < B2_MESH_SMEM_STACK ? s_stack[sp][threadIdx.x] : l_stack[sp - B2_MESH_SMEM_STACK];
      if (__uint_as_float(e.y) <= tmax) return (int32_t)e.x;
    }
    return kDone;
  };

  for (;;) {
    // (1) refill: idle lanes take the next pairs; wait until half of the warp is idle so the
    //     per-pair setup (ray_map, basis, reciprocals) does not run for one lane at a time
    const unsigned idle = __ballot_sync(0xffffffffu, !have);
    const int missing = __popc(idle);
    if (idle == 0xffffffffu || missing >= B2_MESH_REFILL) {
      const int leader = __ffs(idle) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(p.ms.counters + 3, missing);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (!have) {
        pair = base + __popc(idle & lt);
        if (pair < npair) {
          slot = p.ms.pair_slot[pair];
          const int g = p.ms.pair_geom[pair];
          const uint32_t ray = p.ms.aff_ray[slot];
          const int scan = (int)(ray / (uint32_t)p.nray);
          const int r = (int)(ray - (uint32_t)scan * (uint32_t)p.nray);
          const int world = p.scan_world ? p.scan_world[scan] : (p.scan0 + scan) % p.nworld;
          const GeomPose P = ld_pose(p.pose + (size_t)world * p.ngeom + g);
          double o[3], l[3], vec[3];
          scan_ray<MODE>(p, scan, r, o, l, vec);
          ray_map(P, o, vec, lpnt, lvec);
          mesh_ray_basis(lvec, b0, b1);
          rayf_setup(rf, lpnt, lvec);
          const MeshInfo mi = p.mesh[p.gstat[g].dataid];
          nodes = p.mnodes + mi.node_base;
          tris = p.mtris + mi.tri_base;
          // what the primitives and the other pairs of this ray found so far bounds the search
          const unsigned long long best =
              *reinterpret_cast<volatile unsigned long long*>(p.ms.aff_bits + slot);
          tmax = best == kInfBits ? tcap : fminf(tcap, f32_up(__longlong_as_double((long long)best)));
          x = -1;
          sp = 0;
          cur = 0;
          have = true;
#ifdef B2_MESH_STATS
          st_nodes = st_tris = 0;
          st_t0 = clock64();
#endif
        }
      }
    }
    if (!__any_sync(0xffffffffu, have)) break;

    // (2) inner nodes until this lane stands on a leaf (cur < 0) or has nothing left (kDone)
    // bounded descent: lanes that reached a leaf do not wait for the longest descent of the warp
    // (cap 2 / 3 / 4 / 6 / none measured 9.39 / 9.61 / 9.61 / 9.71 / 9.62 Grays/s)
    for (int it = 0; it < B2_MESH_INNER_CAP && have && cur >= 0; it++) {
      const MeshNode4 n = ld_node4(nodes + cur);
#ifdef B2_MESH_STATS
      st_nodes++;
#endif
      float t[4];
      int32_t lk[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const float lo[3] = {n.lox[c], n.loy[c], n.loz[c]}, hi[3] = {n.hix[c], n.hiy[c], n.hiz[c]};
        float tc;
        const bool h = n.child[c] != kEmptyChild && slab_test(rf, lo, hi, tmax, tc);
        t[c] = h ? tc : INFINITY;
        lk[c] = n.child[c];
      }
      // sort the four (entry distance, link) by distance; missed children (inf) end up last
#define B2_CAS(i, j)                      \
  {                                       \
    const bool sw = t[j] < t[i];          \
    const float ta = sw ? t[j] : t[i];    \
    const float tb = sw ? t[i] : t[j];    \
    const int32_t la = sw ? lk[j] : lk[i]; \
    const int32_t lb = sw ? lk[i] : lk[j]; \
    t[i] = ta;                            \
    t[j] = tb;                            \
    lk[i] = la;                           \
    lk[j] = lb;                           \
  }
      B2_CAS(0, 1) B2_CAS(2, 3) B2_CAS(0, 2) B2_CAS(1, 3) B2_CAS(1, 2)
#undef B2_CAS
      if (t[0] < INFINITY) {
        // enter the nearest; the others wait on the stack, farthest first
#pragma unroll
        for (int c = 3; c >= 1; c--) {
          if (t[c] < INFINITY && sp < kMeshStack) push(lk[c], t[c]);
        }
        cur = lk[0];
        if ((B2_MESH_PREFETCH & 2) && t[1] < INFINITY) {
          // the runner-up is the most likely next stop after the nearest subtree
          if (lk[1] >= 0) mesh_prefetch(nodes + lk[1]);
          else mesh_prefetch(tris + (~lk[1] >> 2));
        }
      } else {
        cur = pop();
      }
      if ((B2_MESH_PREFETCH & 1) && cur < 0 && cur != kDone) {
        // the leaf is served in phase (3), after the other lanes' descents: fetch its triangles now
        const MeshTri* t0 = tris + (~cur >> 2);
        mesh_prefetch(t0);
        if ((~cur & 3) >= 2) mesh_prefetch(t0 + 2);
      }
    }

    // (3) the leaf's triangles (fp64, ref: engine_ray.c ray_triangle)
    if (have && cur < 0 && cur != kDone) {
      const int32_t payload = ~cur;
      const int first = payload >> 2, cnt = (payload & 3) + 1;
#pragma unroll 1
      for (int i = 0; i < cnt; i++) {
        const MeshTri tr = ld_tri(tris + first + i);
        const double sol = ray_triangle(tr.v0, tr.v1, tr.v2, lpnt, lvec, b0, b1);
#ifdef B2_MESH_STATS
        st_tris++;
#endif
        if (sol >= 0 && (x < 0 || sol < x)) {
          x = sol;
          tmax = fminf(tmax, f32_up(sol));
        }
      }
      cur = pop();
    }

    // (4) pair finished: publish
    if (have && cur == kDone) {
#ifdef B2_MESH_STATS
      {
        const unsigned long long cyc = (unsigned long long)(clock64() - st_t0);
        atomicAdd(g_mesh_stats + 0, 1ull);
        atomicAdd(g_mesh_stats + 1, (unsigned long long)st_nodes);
        atomicAdd(g_mesh_stats + 2, (unsigned long long)st_tris);
        atomicAdd(g_mesh_stats + 3, cyc);
        atomicMax(g_mesh_stats + 4, (unsigned long long)st_nodes);
        atomicMax(g_mesh_stats + 5, (unsigned long long)st_tris);
        atomicMax(g_mesh_stats + 6, cyc);
        atomicAdd(g_mesh_stats + 7, x >= 0 ? 1ull : 0ull);
        atomicAdd(g_mesh_stats + 16 + ::min(15, 32 - __clz(st_nodes)), 1ull);          // log2 histogram of nodes
        atomicAdd(g_mesh_stats + 32 + ::min(15, 32 - __clz(st_tris)), 1ull);           // ... of triangles
        atomicAdd(g_mesh_stats + 48 + ::min(15, ::max(0, 64 - __clzll(cyc) - 8)), 1ull);  // ... of cycles / 256
      }
#endif
      p.ms.pair_x[pair] = x;
      if (x >= 0) atomicMin(p.ms.aff_bits + slot, (unsigned long long)__double_as_longlong(x));
      have = false;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K2b for SMALL launches (one scan: ~10 k pairs): G lanes per pair
// ---------------------------------------------------------------------------------------------
//
// One lane per pair makes a small launch as long as its heaviest pair: a ray that grazes a mesh walks
// ~60 nodes and ~30 triangles, an ~11 000-instruction dependent chain (~95 us), while the machine
// idles (tools/mesh_stats.py).  Here a GROUP of G lanes (8, 16 or 32: the fewer pairs, the wider)
// shares one pair and one traversal stack in shared memory.  Per round, quad q of the group pops
// entry q from the top of the stack; its four lanes test the four children of that node -- or up to
// four triangles of that leaf -- at once; hits are pushed back with ballot / popc addressing, the
// nearest child of the nearest node on top.  The walk is no longer strictly nearest-first, which
// costs a few extra nodes and is why the one-lane kernel stays the throughput path; the result
// cannot differ: it is the exact minimum over the fp64 triangle tests, and boxes only ever cull what
// lies beyond a hit already found.
constexpr int kCoopStackPerWarp = 512;  // stack entries per warp, split between its 32 / G groups
constexpr int kCoopReserve = 64;        // >= 3 * depth + 1 of any mesh tree (scene_build.cpp)
template <int MODE, int G>
__device__ __forceinline__ void mesh_pairs_coop(const CastParams& p, int npair, uint2* warp_stk) {
  const float tcap = (MODE == MODE_MULTIRAY) ? f32_up(p.cutoff) : FLT_MAX;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);   // lane within the group
  const int g0 = lane & ~(G - 1);  // first lane of the group
  const int q0 = lane & ~3;        // first lane of the quad
  const int quad = gl >> 2, child = lane & 3;
  constexpr int nq = G >> 2, ngroup = 32 / G;
  constexpr int cap = kCoopStackPerWarp / ngroup, cap_soft = cap - kCoopReserve;
  constexpr unsigned leaders = G == 32 ? 1u : (G == 16 ? 0x00010001u : 0x01010101u);
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << g0);
  uint2* stk = warp_stk + (g0 / G) * cap;

  bool have = false;
  int pair = 0, slot = 0, sp = 0;
  double lpnt[3], lvec[3], b0[3], b1[3], x = -1;
  RayF rf;
  float tmax = 0.f;
  const MeshNode4* nodes = nullptr;
  const MeshTri* tris = nullptr;

  for (;;) {
    // (1) refill: groups without a pair take the next ones (every lane of a group sets the pair up
    //     for itself: same loads, same arithmetic, same values)
    const unsigned idle = __ballot_sync(0xffffffffu, !have) & leaders;
    if (idle == leaders || __popc(idle) * 2 >= ngroup) {
      const int first = __ffs(idle) - 1;
      int base = 0;
      if (lane == first) base = atomicAdd(p.ms.counters + 3, __popc(idle));
      base = __shfl_sync(0xffffffffu, base, first);
      const bool fresh = !have;
      if (fresh) {
        pair = base + __popc(idle & ((1u << g0) - 1u));
        if (pair < npair) {
          slot = p.ms.pair_slot[pair];
          const int g = p.ms.pair_geom[pair];
          const uint32_t ray = p.ms.aff_ray[slot];
          const int scan = (int)(ray / (uint32_t)p.nray);
          const int r = (int)(ray - (uint32_t)scan * (uint32_t)p.nray);
          const int world = p.scan_world ? p.scan_world[scan] : (p.scan0 + scan) % p.nworld;
          const GeomPose P = ld_pose(p.pose + (size_t)world * p.ngeom + g);
          double o[3], l[3], vec[3];
          scan_ray<MODE>(p, scan, r, o, l, vec);
          ray_map(P, o, vec, lpnt, lvec);
          mesh_ray_basis(lvec, b0, b1);
          rayf_setup(rf, lpnt, lvec);
          const MeshInfo mi = p.mesh[p.gstat[g].dataid];
          nodes = p.mnodes + mi.node_base;
          tris = p.mtris + mi.tri_base;
          x = -1;
          have = true;
        }
      }
      // the bound other pairs of the ray have published so far: read by one lane, so that every lane
      // of the group starts from the same tmax
      unsigned long long best = kInfBits;
      if (fresh && have && gl == 0)
        best = *reinterpret_cast<volatile unsigned long long*>(p.ms.aff_bits + slot);
      best = __shfl_sync(0xffffffffu, best, g0);
      if (fresh && have) {
        tmax = best == kInfBits ? tcap : fminf(tcap, f32_up(__longlong_as_double((long long)best)));
        if (gl == 0) stk[0] = make_uint2(0u, 0u);  // the root, entry distance 0
        sp = 1;
      }
      __syncwarp();
    }
    if (!__any_sync(0xffffffffu, have)) break;

    // (2) pop: quad q takes entry q from the top; only one entry while the stack is near its cap
    int ntake = 0;
    if (have) {
      ntake = ::min(nq, sp);
      if (sp + 3 * nq > cap_soft) ntake = 1;
    }
    const bool mine = quad < ntake;
    uint2 e = make_uint2(0u, 0u);
    if (mine) e = stk[sp - 1 - quad];
    const int spn = sp - ntake;
    const bool live = mine && __uint_as_float(e.y) <= tmax;
    const int32_t link = (int32_t)e.x;
    __syncwarp();  // every pop has read its entry before the pushes below overwrite the top

    // (3) a node: this lane's child box; a leaf: this lane's triangle
    bool hitc = false;
    float tc = INFINITY;
    int32_t clink = kEmptyChild;
    if (live && link >= 0) {
      const MeshNode4* n = nodes + link;
      clink = __ldg(n->child + child);
      const float lo[3] = {__ldg(n->lox + child), __ldg(n->loy + child), __ldg(n->loz + child)};
      const float hi[3] = {__ldg(n->hix + child), __ldg(n->hiy + child), __ldg(n->hiz + child)};
      hitc = clink != kEmptyChild && slab_test(rf, lo, hi, tmax, tc);
      if (!hitc) tc = INFINITY;  // (slab_test leaves its entry distance behind on a miss too)
    }
    const bool at_leaf = live && link < 0;
    if (__any_sync(0xffffffffu, at_leaf)) {
      double sol = INFINITY;
      if (at_leaf) {
        const int32_t payload = ~link;
        const int first = payload >> 2, cnt = (payload & 3) + 1;
        if (child < cnt) {
          const MeshTri tr = ld_tri(tris + first + child);
          const double t = ray_triangle(tr.v0, tr.v1, tr.v2, lpnt, lvec, b0, b1);
          if (t >= 0) sol = t;
        }
      }
      // nearest hit of the round over the group (exact: min of doubles)
#pragma unroll
      for (int o = G >> 1; o > 0; o >>= 1) sol = fmin(sol, __shfl_xor_sync(0xffffffffu, sol, o));
      if (sol < INFINITY && (x < 0 || sol < x)) {
        x = sol;
        tmax = fminf(tmax, f32_up(sol));
      }
    }

    // (4) push the children that were hit: higher quads first, and within a quad the farthest first,
    //     so that the nearest child of the top entry ends on top
    const unsigned hm = __ballot_sync(0xffffffffu, hitc);
    if (hm) {        // (warp-uniform: the shuffles below name every lane)
      int nearer = 0;  // hit children of this quad that are nearer than this one
#pragma unroll
      for (int k = 1; k < 4; k++) {
        const float tj = __shfl_xor_sync(0xffffffffu, tc, k);  // a missed child carries +inf
        const int j = child ^ k;
        nearer += (tj < tc || (tj == tc && j < child && tj < INFINITY)) ? 1 : 0;
      }
      if (hitc) {
        const unsigned above = q0 + 4 >= 32 ? 0u : (hm & gmask & ~((1u << (q0 + 4)) - 1u));
        const int qh = __popc(hm & (0xfu << q0));
        const int idx = spn + __popc(above) + (qh - 1 - nearer);
        if (idx < cap) stk[idx] = make_uint2((uint32_t)clink, __float_as_uint(tc));
        else p.ms.counters[2] = 1;  // cannot happen (reserve >= 3 * depth + 1); the clean-up launch recomputes
      }
    }
    sp = ::min(spn + __popc(hm & gmask), cap);
    __syncwarp();

    // (5) pair finished: publish
    if (have && sp == 0) {
      if (gl == 0) {
        p.ms.pair_x[pair] = x;
        if (x >= 0) atomicMin(p.ms.aff_bits + slot, (unsigned long long)__double_as_longlong(x));
      }
      have = false;
    }
  }
}

#ifndef B2_COOP_MINBLOCKS
#define B2_COOP_MINBLOCKS 8
#endif
template <int MODE>
__global__ void __launch_bounds__(kMeshThreads, B2_COOP_MINBLOCKS)
    mesh_pairs_coop_kernel(const __grid_constant__ CastParams p, int g_req) {
  const int npair = ::min(p.ms.counters[0], p.ms.pair_cap);
  __shared__ uint2 s_stk[kMeshThreads / 32][kCoopStackPerWarp];
  uint2* warp_stk = s_stk[threadIdx.x >> 5];
  // group width: a whole warp per pair while that leaves every resident warp at most ~3 pairs
  // (one 64x2048 scan: 11 k pairs, 4736 warps; 133.7 us per scan against 134.5 at 16 lanes and 159.3
  // with one lane per pair), half a warp beyond (two scans: 200 / 207 / 219 us at 16 / 32 / 1 lanes)
  const int nwarp = (int)gridDim.x * (kMeshThreads / 32);
  const int G = (g_req == 8 || g_req == 16 || g_req == 32) ? g_req : (npair <= 3 * nwarp ? 32 : 16);
  if (G == 32) mesh_pairs_coop<MODE, 32>(p, npair, warp_stk);
  else if (G == 16) mesh_pairs_coop<MODE, 16>(p, npair, warp_stk);
  else mesh_pairs_coop<MODE, 8>(p, npair, warp_stk);
}

// K2c: lowest mesh geom id among the pairs that reached the ray's minimum (A.5 tie rule)
__global__ void __launch_bounds__(kMeshThreads) mesh_ids_kernel(const __grid_constant__ CastParams p) {
  const int npair = ::min(p.ms.counters[0], p.ms.pair_cap);
  for (int i = blockIdx.x * kMeshThreads + threadIdx.x; i < npair; i += gridDim.x * kMeshThreads) {
    const double x = p.ms.pair_x[i];
    if (x < 0) continue;
    const int slot = p.ms.pair_slot[i];
    if ((unsigned long long)__double_as_longlong(x) == p.ms.aff_bits[slot])
      atomicMin(p.ms.aff_mesh_id + slot, p.ms.pair_geom[i]);
  }
}

// K2d: rays whose closest hit a mesh changed (or tied): rewrite their outputs
template <int MODE>
__global__ void __launch_bounds__(kMeshThreads) mesh_final_kernel(const __grid_constant__ CastParams p) {
  const int naff = ::min(p.ms.counters[1], p.ms.aff_cap);
  for (int s = blockIdx.x * kMeshThreads + threadIdx.x; s < naff; s += gridDim.x * kMeshThreads) {
    const int mesh_id = p.ms.aff_mesh_id[s];
    if (mesh_id == INT32_MAX) continue;  // no mesh reached the minimum: K2's output stands
    const unsigned long long fin = p.ms.aff_bits[s], prim = p.ms.aff_prim[s];
    Hit hit;
    hit.dist = __longlong_as_double((long long)fin);
    hit.id = fin < prim ? mesh_id : ::min(p.ms.aff_prim_id[s], mesh_id);
    const uint32_t ray = p.ms.aff_ray[s];
    const int scan = (int)(ray / (uint32_t)p.nray), r = (int)(ray - (uint32_t)scan * (uint32_t)p.nray);
    double le[3] = {0.0, 0.0, 0.0};
    if (MODE == MODE_LIDAR && p.xyz != nullptr) {
#pragma unroll
      for (int k = 0; k < 3; k++) le[k] = __ldg(p.local_dir + (size_t)r * 3 + k);
    }
    cast_epilogue(p, MODE, scan, r, hit, le);
  }
}

// ---------------------------------------------------------------------------------------------
// launch
// ---------------------------------------------------------------------------------------------
template <class K>
static cudaError_t launch_tile_kernel(K k, const CastParams& p, const TileGeom& tg, unsigned nblocks,
                                      size_t smem, cudaStream_t stream) {
  cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  // ask for the largest shared-memory carve-out so several CTAs fit next to each other
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
  k<<<nblocks, kCastThreads, smem, stream>>>(p, tg);
  return cudaGetLastError();
}

// K2a -> K2b -> fused fall-back for the groups K2a gave up on
template <int MODE, bool SPLIT>
static cudaError_t launch_two_kernel(const CastParams& p, const TileGeom& tg, bool stage,
                                     long long nblocks_static, int sm_count, cudaStream_t stream) {
  cudaError_t err = cudaMemsetAsync(p.cl.counters, 0, 4 * sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  size_t smem = kCastWarps * sizeof(CollectScratch);
  if (stage) smem += cast_stage_bytes(p.nslot, p.nwnode, p.ngeom);
  const unsigned cgrid =
      tg.dynamic ? (unsigned)(sm_count * B2_COLLECT_MINBLOCKS) : (unsigned)nblocks_static;
  err = stage ? launch_tile_kernel(collect_kernel<MODE, true>, p, tg, cgrid, smem, stream)
              : launch_tile_kernel(collect_kernel<MODE, false>, p, tg, cgrid, smem, stream);
  if (err != cudaSuccess) return err;
  drain_kernel<MODE, SPLIT><<<(unsigned)(sm_count * B2_DRAIN_MINBLOCKS), kCastThreads, 0, stream>>>(
      p, tg);
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  // the groups K2a gave up on: the fused kernel.  One world: the persistent staged form of
  // cast_kernel reading its groups from the list; several worlds: unstaged, world per group.
  if (p.nworld == 1 && stage) {
    TileGeom tf = tg;
    tf.dynamic = 2;
    return launch_tile_kernel(cast_kernel<MODE, true, SPLIT, false>, p, tf,
                              (unsigned)(sm_count * B2_CAST_MINBLOCKS),
                              kCastWarps * sizeof(WarpScratch) +
                                  cast_stage_bytes(p.nslot, p.nwnode, p.ngeom),
                              stream);
  }
  return launch_tile_kernel(fallback_kernel<MODE, SPLIT>, p, tg,
                            (unsigned)(sm_count * B2_CAST_MINBLOCKS),
                            kCastWarps * sizeof(WarpScratch), stream);
}

// B2_VEC_OUT=0 turns the float4 / int4 result stores of K2b off (A/B runs)
static int env_int_local(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
static bool env_flag_vec_out() {
  const char* v = std::getenv("B2_VEC_OUT");
  return !(v && v[0] == '0');
}

template <int MODE>
static cudaError_t launch_cast_mode(const CastParams& p, bool stage, int tiles_per_block,
                                    int sm_count, cudaStream_t stream, int* nlaunch) {
  // 2-D tiling when the sensor is a grid of scan lines
  TileGeom tg;
  tg.tiles_x = 0;
  tg.tiles_per_scan = cast_tiles_per_scan(p.nray);
  if (p.tile_w > 0 && p.tile_w % 8 == 0 && p.nray % p.tile_w == 0) {
    const int rows = p.nray / p.tile_w;
    if (rows >= 4 && rows % 4 == 0) {
      tg.tiles_x = p.tile_w / 8;
      tg.tiles_per_scan = tg.tiles_x * (rows / 4);
    }
  }
  if (tiles_per_block < kCastWarps) tiles_per_block = kCastWarps;
  tg.tiles_per_block = tiles_per_block;
  tg.chunks_per_scan = (tg.tiles_per_scan + tiles_per_block - 1) / tiles_per_block;
  long long nblocks = (long long)p.nscan * tg.chunks_per_scan;
  if (nblocks <= 0) return cudaSuccess;
  if (nblocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
  // persistent launch: one resident wave of CTAs whose warps pull groups from the work counter
  tg.group_tiles = (p.group_tiles == 1 || p.group_tiles == 2) ? p.group_tiles : kGroupTiles;
  tg.groups_per_scan = (tg.tiles_per_scan + tg.group_tiles - 1) / tg.group_tiles;
  const long long ngroups = (long long)p.nscan * tg.groups_per_scan;
  const long long resident = (long long)sm_count * B2_CAST_MINBLOCKS;
  tg.dynamic = p.work_counter != nullptr && stage && p.nworld == 1 && ngroups <= 0x7fffffffLL &&
               ngroups >= 4 * resident * kCastWarps;
  tg.ngroups = (int)ngroups;
  tg.vec_out = env_flag_vec_out() && MODE == MODE_LIDAR && tg.tiles_x > 0 && tg.tiles_x % kGroupTiles == 0 &&
               tg.group_tiles == kGroupTiles && p.dist == nullptr && p.xyz == nullptr &&
               (reinterpret_cast<uintptr_t>(p.ranges) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(p.geomid) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(p.geomid16) & 7) == 0 && p.nray % 4 == 0;
  const long long nblocks_static = nblocks;
  if (tg.dynamic) nblocks = resident;
  size_t smem = kCastWarps * sizeof(WarpScratch);
  if (stage) smem += cast_stage_bytes(p.nslot, p.nwnode, p.ngeom);
  const bool split = p.ms.counters != nullptr;
  cudaError_t err;
  if (tg.dynamic) {
    err = cudaMemsetAsync(p.work_counter, 0, sizeof(int32_t), stream);
    if (err != cudaSuccess) return err;
  }
  const bool two = p.cl.count != nullptr;
  if (!split) {
    if (two) {
      *nlaunch += 3;
      return launch_two_kernel<MODE, false>(p, tg, stage, nblocks_static, sm_count, stream);
    }
    err = stage ? launch_tile_kernel(cast_kernel<MODE, true, false, false>, p, tg, (unsigned)nblocks,
                                     smem, stream)
                : launch_tile_kernel(cast_kernel<MODE, false, false, false>, p, tg,
                                     (unsigned)nblocks, smem, stream);
    *nlaunch += 1;
    return err;
  }
  err = cudaMemsetAsync(p.ms.counters, 0, 4 * sizeof(int32_t), stream);
  if (err != cudaSuccess) return err;
  if (two) {
    *nlaunch += 2;
    err = launch_two_kernel<MODE, true>(p, tg, stage, nblocks_static, sm_count, stream);
  } else {
    err = stage ? launch_tile_kernel(cast_kernel<MODE, true, true, false>, p, tg, (unsigned)nblocks,
                                     smem, stream)
                : launch_tile_kernel(cast_kernel<MODE, false, true, false>, p, tg,
                                     (unsigned)nblocks, smem, stream);
  }
  if (err != cudaSuccess) return err;
  const unsigned mgrid = (unsigned)(sm_count * 16);
  // persistent: exactly the CTAs that are resident at once
  // a launch of one or two scans (the plugin's duty cycle: one scan per publish) is as long as its
  // heaviest pair: G lanes per pair there.  From four 64x2048 scans on, the one-lane kernel wins
  // (tools/latency_probe.py --coop: 1 / 2 / 4 / 8 scans 134 / 200 / 305 / 441 us against 159 / 219 / 269 / 367)
  const int coop_rays = env_int_local("B2_MESH_COOP_RAYS", 1 << 18);
  const int coop_g = env_int_local("B2_MESH_COOP_G", 0);
  if ((long long)p.nscan * p.nray <= coop_rays)
    mesh_pairs_coop_kernel<MODE><<<(unsigned)(sm_count * B2_COOP_MINBLOCKS), kMeshThreads, 0, stream>>>(p, coop_g);
  else
    mesh_pairs_kernel<MODE><<<(unsigned)(sm_count * B2_MESH_MINBLOCKS), kMeshThreads, 0, stream>>>(p);
  mesh_ids_kernel<<<mgrid, kMeshThreads, 0, stream>>>(p);
  mesh_final_kernel<MODE><<<mgrid, kMeshThreads, 0, stream>>>(p);
  err = cudaGetLastError();
  if (err != cudaSuccess) return err;
  // clean-up: recompute the chunk in place if the pair list overflowed (exits at once otherwise);
  // always the static tiling -- it is (almost) never more than an empty launch
  tg.dynamic = 0;
  nblocks = nblocks_static;
  err = stage ? launch_tile_kernel(cast_kernel<MODE, true, false, true>, p, tg, (unsigned)nblocks,
                                   smem, stream)
              : launch_tile_kernel(cast_kernel<MODE, false, false, true>, p, tg, (unsigned)nblocks,
                                   smem, stream);
  *nlaunch += 5;
  return err;
}

cudaError_t launch_cast(const CastParams& p, bool stage_tlas, int tiles_per_block, int sm_count,
                        cudaStream_t stream, int* nlaunch) {
  if (p.nwnode == 0 || cast_stage_bytes(p.nslot, p.nwnode, p.ngeom) > kStageLimit) stage_tlas = false;
  switch (p.mode) {
    case MODE_LIDAR:
      return launch_cast_mode<MODE_LIDAR>(p, stage_tlas, tiles_per_block, sm_count, stream, nlaunch);
    case MODE_MULTIRAY:
      return launch_cast_mode<MODE_MULTIRAY>(p, stage_tlas, tiles_per_block, sm_count, stream,
                                             nlaunch);
    case MODE_CAMERA:
      return launch_cast_mode<MODE_CAMERA>(p, stage_tlas, tiles_per_block, sm_count, stream,
                                           nlaunch);
    default:
      return cudaErrorInvalidValue;
  }
}

}  // namespace b2
