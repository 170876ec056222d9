// runtime.cu -- host runtime and C ABI of libb2sense.so (include/b2sense.h).
//
// Scene mirror: static model copied once, per-world poses packed on the device by K1, geom BVH
// topology built on the host (binned SAH) from one world's poses and refit per step on the device.
// There is no CPU compute path: without a usable CUDA device every entry point returns B2_ERR_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b2sense.h"
#include "b2_types.h"
#include "bvh_build.h"
#include "kernels.cuh"
#include "ray_core.cuh"
#include "scene_build.h"

using namespace b2;

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define B2_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t e_ = (call);                                                               \
    if (e_ != cudaSuccess)                                                                 \
      return fail(B2_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));        \
  } while (0)

// bumped whenever a device / pinned buffer of the library is (re)allocated: captured CUDA graphs
// bake buffer addresses in, so a graph is only replayed while this is what it was at capture
std::atomic<unsigned long long> g_alloc_epoch{0};

template <class T>
struct DevBuf {
  T* ptr = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    g_alloc_epoch++;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&ptr, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

template <class T>
struct PinBuf {
  T* ptr = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    g_alloc_epoch++;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    cudaError_t e = cudaMallocHost(&ptr, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct b2_scene {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  int ngeom = 0, nbody = 0, nmesh = 0, nmat = 0, nworld = 0;
  HostModel hm;  // host copy of the static model
  // device: static
  DevBuf<GeomStatic> d_gstat;
  DevBuf<int32_t> d_plane_ids;
  DevBuf<MeshInfo> d_mesh;
  DevBuf<MeshNode4> d_mnodes;
  DevBuf<MeshTri> d_mtris;
  long long n_mesh_nodes = 0, n_mesh_tris = 0;
  // device: per world
  DevBuf<GeomPose> d_pose;
  DevBuf<WideChild> d_wslots;  // [nworld][nslot]
  // geom BVH topology (fanout-32 tree, shared by all worlds)
  bool topo_built = false;
  int nslot = 0, nwnode = 0, nlevel = 0, wide_depth = 0;
  DevBuf<WideNode> d_wnodes;
  DevBuf<int32_t> d_slot_link, d_refit_slots, d_level_start;
  // pose staging of the immediate (partial-range) path: double-buffered pinned host memory
  PinBuf<double> h_stage[2];
  cudaEvent_t stage_done[2] = {nullptr, nullptr};
  int stage_flip = 0;
  // device raw arrays (mjData layout), two halves used alternately; refit_done[f]: the refit that
  // read half f has finished (the next upload into it waits for that on the copy stream)
  DevBuf<double> d_raw;
  cudaEvent_t refit_done[2] = {nullptr, nullptr};
  int raw_flip = 0;
  // "latest wins" host mirror of mjData's geom poses (full-range updates after the first one): the
  // producer (physics thread) only copies into a pinned slot -- no CUDA call, no scene mutex; the
  // next consumer (cast / render / sync) uploads the newest slot on the copy stream and refits.
  struct PoseMirror {
    std::mutex idx;   // guards the indices below, never held across a memcpy or a CUDA call
    std::mutex prod;  // serialises producers
    PinBuf<double> slot[3];
    int latest = -1, uploading = -1;
    unsigned long long version = 0, uploaded = 0;
    cudaEvent_t upload_done = nullptr;
    bool upload_pending = false;
  } pm;
  int topo_version = 0;  // bumped by every topology (re)build: invalidates captured graphs
  std::atomic<int> dependents{0};  // live lidars / cameras bound to this scene
  bool capturing = false;  // the scene's stream is being captured into a graph (scene mutex held)
  cudaEvent_t sync_mark[2] = {nullptr, nullptr};  // b2_scene_sync
  cudaEvent_t copy_fork = nullptr, copy_join = nullptr;  // b2_scene_copy_async / b2_scene_join
  cudaEvent_t copy_fence[2] = {nullptr, nullptr};        // b2_scene_copy_fence (one-step lag)
  bool copy_fence_used[2] = {false, false};
  int copy_fence_flip = 0;
  // scratch of the split mesh path (sized for one chunk of rays)
  DevBuf<int32_t> d_pair_slot, d_pair_geom, d_aff_prim_id, d_aff_mesh_id, d_counters;
  DevBuf<double> d_pair_x;
  DevBuf<uint32_t> d_aff_ray;
  DevBuf<unsigned long long> d_aff_bits, d_aff_prim;
  // scratch of the two-kernel cast: per-group candidate lists
  DevBuf<int32_t> d_cl_count, d_cl_ovf;
  DevBuf<CandRef> d_cl_list;
  bool has_mesh_geoms = false;
  // The plugin's physics, lidar and camera threads share one scene: every enqueue on the scene's
  // stream (and every use of the scene-owned scratch) is serialised by this mutex.
  std::recursive_mutex mtx;
  // scratch for b2_multiray
  DevBuf<double> d_mr_vec, d_mr_dist, d_mr_pnt;
  DevBuf<int32_t> d_mr_id;
  DevBuf<uint8_t> d_mr_elim;
};

// One host cast between b2_lidar_cast_begin and b2_lidar_cast_end
struct HostCast {
  bool active = false;
  bool small = false;   // one piece, linear on the scene's stream (CUDA-graph cached)
  bool staged = false;  // results land in the lidar's pinned staging first
  int nscan = 0;
  size_t n = 0;  // rays
  float* ranges = nullptr;
  int32_t* geomid = nullptr;
  uint16_t* geomid16 = nullptr;
  double* dist = nullptr;
  float* xyz = nullptr;
  std::vector<int> bounds;  // scan bounds of the pieces (bulk path)
};

struct LidarGraph {
  int nscan = 0, mask = 0, topo_version = -1;
  bool has_world = false;
  bool seen = false;  // the same shape ran once un-captured (all allocations exist)
  int kernels = 0;    // kernels inside the captured graph (b2_launch_count accounting)
  unsigned long long alloc_epoch = 0;  // g_alloc_epoch when `seen` / captured
  cudaGraphExec_t exec = nullptr;
};

struct b2_lidar {
  b2_scene* scene = nullptr;
  int nray = 0, tile_w = 0;
  int group_tiles = 4;  // tiles per collect group, chosen from the angular width of the ray table
  double cutoff = 0, range_min = 0, range_max = 1000;
  DevBuf<double> d_local_dir;
  DevBuf<uint8_t> d_elim;
  // scratch of the host-pointer entry point
  DevBuf<double> d_spos, d_smat, d_dist;
  DevBuf<double> d_in;  // small path: pos | mat | world of the call in one block (one H2D copy)
  DevBuf<int32_t> d_world, d_geomid;
  DevBuf<uint16_t> d_geomid16;
  DevBuf<float> d_ranges, d_xyz;
  // pinned staging: sensor poses in, results out (when the caller's buffers are pageable, and always
  // on the small path so that the whole call is one replayable graph over fixed addresses)
  PinBuf<double> h_in;
  PinBuf<unsigned char> h_out;
  size_t off_ranges = 0, off_geomid = 0, off_geomid16 = 0, off_dist = 0, off_xyz = 0;
  std::vector<cudaEvent_t> piece_done;  // D2H of piece c has landed
  cudaEvent_t all_done = nullptr;       // everything of the current host cast (small path)
  // cast time of the previous blocking call (events around its kernels): decides how finely the
  // next one is pipelined against the read-back
  cudaEvent_t cast_t0 = nullptr, cast_t1 = nullptr;
  double cast_ns_per_ray = 0.0;
  double last_device_us = 0.0;  // device time of the last host cast (first enqueue .. last kernel)
  HostCast pend;
  LidarGraph graph;
  bool counted = false;  // registered in scene->dependents
};

struct b2_camera {
  b2_scene* scene = nullptr;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  b2_camera_desc desc{};
  double K[9];
  float near_f = 0, far_f = 0, depth_scale = 0;
  DevBuf<float> d_in, d_depth, d_cloud;
  DevBuf<uint8_t> d_elim;  // geoms the renderer does not draw (invisible, hidden groups)
  DevBuf<double> d_pos, d_mat;
  DevBuf<int32_t> d_world;
  // host render (begin .. end): pinned staging in / out, completion event
  PinBuf<double> h_in;
  PinBuf<unsigned char> h_out;
  cudaEvent_t done = nullptr;
  struct {
    bool active = false;
    int ncam = 0;
    float* gl_depth = nullptr;
    float* depth32fc1 = nullptr;
    void* cloud = nullptr;
    size_t off_gl = 0, off_depth = 0, off_cloud = 0;
  } pend;
};

static int flush_pending(b2_scene* s);  // newest poses of the host mirror -> device (+ refit)

namespace {

int build_meshes(b2_scene* s, const b2_model_desc* m) {
  HostMeshes hmesh;
  std::string err;
  if (!build_mesh_bvhs(s->nmesh, m->mesh_vertadr, m->mesh_vertnum, m->mesh_faceadr, m->mesh_facenum,
                       m->mesh_vert, m->mesh_face, env_int("B2_MESH_LEAF", 2), kMeshStack - 2, hmesh,
                       err))
    return fail(B2_ERR_INVALID, err);
  s->n_mesh_nodes = (long long)hmesh.nodes.size();
  s->n_mesh_tris = (long long)hmesh.tris.size();
  if (s->nmesh > 0) {
    B2_CUDA(s->d_mesh.reserve(hmesh.infos.size()));
    B2_CUDA(cudaMemcpy(s->d_mesh.ptr, hmesh.infos.data(), hmesh.infos.size() * sizeof(MeshInfo),
                       cudaMemcpyHostToDevice));
    B2_CUDA(s->d_mnodes.reserve(hmesh.nodes.size() + 1));
    B2_CUDA(cudaMemcpy(s->d_mnodes.ptr, hmesh.nodes.data(), hmesh.nodes.size() * sizeof(MeshNode4),
                       cudaMemcpyHostToDevice));
    B2_CUDA(s->d_mtris.reserve(hmesh.tris.size() + 1));
    if (!hmesh.tris.empty())
      B2_CUDA(cudaMemcpy(s->d_mtris.ptr, hmesh.tris.data(), hmesh.tris.size() * sizeof(MeshTri),
                         cudaMemcpyHostToDevice));
  }
  return B2_OK;
}

// topology of the geom BVH from one world's poses (host arrays, [ngeom*3] / [ngeom*9])
int build_topology(b2_scene* s, const double* xpos, const double* xmat) {
  WideBvh w;
  std::string err;
  if (!build_geom_topology(s->hm, xpos, xmat, kMaxWideDepth, w, err))
    return fail(B2_ERR_UNSUPPORTED, err);
  const int nslot = (int)w.slot_link.size(), nwnode = (int)w.nodes.size();
  s->nslot = nslot;
  s->nwnode = nwnode;
  s->wide_depth = w.depth;
  s->nlevel = (int)w.refit_level_start.size() - 1;
  if (s->nlevel < 0) s->nlevel = 0;
  if (nslot > 0) {
    B2_CUDA(s->d_wnodes.reserve((size_t)nwnode));
    B2_CUDA(cudaMemcpyAsync(s->d_wnodes.ptr, w.nodes.data(), (size_t)nwnode * sizeof(WideNode),
                            cudaMemcpyHostToDevice, s->stream));
    B2_CUDA(s->d_slot_link.reserve((size_t)nslot));
    B2_CUDA(cudaMemcpyAsync(s->d_slot_link.ptr, w.slot_link.data(), (size_t)nslot * sizeof(int32_t),
                            cudaMemcpyHostToDevice, s->stream));
    B2_CUDA(s->d_refit_slots.reserve((size_t)nslot));
    B2_CUDA(cudaMemcpyAsync(s->d_refit_slots.ptr, w.refit_slots.data(),
                            (size_t)nslot * sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
    B2_CUDA(s->d_level_start.reserve((size_t)s->nlevel + 1));
    B2_CUDA(cudaMemcpyAsync(s->d_level_start.ptr, w.refit_level_start.data(),
                            ((size_t)s->nlevel + 1) * sizeof(int32_t), cudaMemcpyHostToDevice,
                            s->stream));
    B2_CUDA(s->d_wslots.reserve((size_t)nslot * s->nworld));
    // the host vectors die at return: the copies above must have consumed them
    B2_CUDA(cudaStreamSynchronize(s->stream));
  }
  s->topo_built = true;
  s->topo_version++;
  return B2_OK;
}

int run_refit(b2_scene* s, int world0, int nworld, const double* d_xpos, const double* d_xmat) {
  RefitParams rp{};
  rp.raw_xpos = d_xpos;
  rp.raw_xmat = d_xmat;
  rp.pose = s->d_pose.ptr;
  rp.wslots = s->d_wslots.ptr;
  rp.gstat = s->d_gstat.ptr;
  rp.wnodes = s->d_wnodes.ptr;
  rp.slot_link = s->d_slot_link.ptr;
  rp.refit_slots = s->d_refit_slots.ptr;
  rp.level_start = s->d_level_start.ptr;
  rp.ngeom = s->ngeom;
  rp.nslot = s->nslot;
  rp.nlevel = s->nlevel;
  rp.world0 = world0;
  rp.nworld = nworld;
  B2_CUDA(launch_refit(rp, s->stream));
  g_launches++;
  return B2_OK;
}

void fill_cast_scene(const b2_scene* s, CastParams& p) {
  p.wslots = s->d_wslots.ptr;
  p.wnodes = s->d_wnodes.ptr;
  p.pose = s->d_pose.ptr;
  p.gstat = s->d_gstat.ptr;
  p.plane_ids = s->d_plane_ids.ptr;
  p.mesh = s->d_mesh.ptr;
  p.mnodes = s->d_mnodes.ptr;
  p.mtris = s->d_mtris.ptr;
  p.nslot = s->nslot;
  p.nwnode = s->nwnode;
  p.ngeom = s->ngeom;
  p.nplane = (int)s->hm.plane_ids.size();
  p.nworld = s->nworld;
}

// CTA sizing: enough CTAs to fill the machine several times over, but as many tiles per CTA as
// that allows so the shared-memory staging of the geom BVH is amortised.
void cast_shape(const b2_scene* s, const CastParams& p, bool& stage, int& tiles_per_block) {
  const int warps = cast_threads_per_block() / 32;
  const long long tiles = (long long)cast_tiles_per_scan(p.nray) * p.nscan;
  const long long want_blocks = (long long)s->sm_count * 16;
  // a warp walks groups of g tiles (4, or fewer for coarse sensors and small launches): the unit of
  // work of a CTA is warps * g tiles.  (A single 64x2048 scan with g = 1 used to get 128 CTAs of 32
  // tiles -- less than one CTA per SM; with the unit following g it gets 512 CTAs of 8.)
  const int g = (p.group_tiles == 1 || p.group_tiles == 2) ? p.group_tiles : 4;
  const int unit = warps * g;
  long long tpb = tiles / want_blocks;
  tpb = (tpb / unit) * unit;
  if (tpb < unit) tpb = unit;
  if (tpb > 2 * unit) tpb = 2 * unit;
  tiles_per_block = env_int("B2_TILES_PER_BLOCK", (int)tpb);
  const size_t bytes = p.nwnode > 0 ? cast_stage_bytes(p.nslot, p.nwnode, p.ngeom) : 0;
  stage = bytes > 0 && bytes <= cast_stage_limit_bytes() &&
          (tiles_per_block >= 2 * unit || bytes <= 4096);
  const int force = env_int("B2_STAGE", -1);
  if (force == 0) stage = false;
  if (force == 1 && bytes > 0 && bytes <= cast_stage_limit_bytes()) stage = true;
}

// Enqueue one cast (any mode) on `stream`.  Large launches on scenes with mesh geoms are cut into
// chunks of whole scans so that the scratch of the split mesh path (K2b-K2d) stays bounded; each
// chunk runs K2 -> K2b -> K2c -> K2d (+ the overflow clean-up guard).
int run_cast(b2_scene* s, CastParams p, cudaStream_t stream) {
  if (p.nscan <= 0 || p.nray <= 0) return B2_OK;
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  {
    int rc = flush_pending(s);  // newest poses of the host mirror, if the producer left any
    if (rc != B2_OK) return rc;
  }
  const long long total = (long long)p.nscan * p.nray;
  bool split = s->has_mesh_geoms && total >= 65536;
  const int force = env_int("B2_SPLIT", -1);
  if (force == 0) split = false;
  if (force == 1 && s->has_mesh_geoms) split = true;
  int scans_per_chunk = p.nscan;
  // counters[0..3]: split mesh path; counters[4]: group cursor of a persistent K2 launch
  // counters[8..11]: the two-kernel cast (CastLists)
  B2_CUDA(s->d_counters.reserve(16));
  p.work_counter = env_int("B2_DYNAMIC", 1) != 0 ? s->d_counters.ptr + 4 : nullptr;
  // two-kernel cast (K2a collect -> K2b drain): large launches on scenes with a geom BVH.  Below
  // ~1 M rays a cast is a chain of latencies, not a throughput problem, and the fused kernel's single
  // launch wins (tools/crossover.py on B200: 64x2048 lidar 0.5 M rays 269 vs 246 us, 1 M 310 vs 329,
  // 8 M 1057 vs 1415; a coarse 64x256 table breaks even at ~4 M)
  const bool two = env_int("B2_TWO_KERNEL", 1) != 0 && s->nwnode > 0 &&
                   total >= env_int("B2_TWO_MIN_RAYS", 1 << 20);
  // measured on B200 (profiles/r01_sweeps.txt): 8M rays/chunk 5.0, 32M 5.4, whole launch 5.6
  // Grays/s -- fewer, longer kernels lose less in launch gaps and tails; scratch is 68 B/ray
  const long long chunk_rays = env_int("B2_CHUNK_MRAYS", 144) * 1000000LL;
  if (split || two)
    scans_per_chunk = (int)std::max(1LL, std::min((long long)p.nscan, chunk_rays / p.nray));
  // small launches: fewer tiles per group until every resident warp has a group -- a cast of a few
  // thousand tiles is a chain of latencies, and 4 tiles per warp in sequence make it 4x longer
  {
    int g = (p.group_tiles == 1 || p.group_tiles == 2) ? p.group_tiles : kGroupTiles;
    const long long tiles = cast_tiles_per_scan(p.nray);
    const long long resident_warps = (long long)s->sm_count * 2 * (cast_threads_per_block() / 32);
    if (env_int("B2_SMALL_SPLIT", 1) != 0)
      while (g > 1 && (long long)p.nscan * ((tiles + g - 1) / g) < resident_warps) g >>= 1;
    p.group_tiles = g;
  }
  p.cl = CastLists{};
  if (two) {
    const size_t gt = (size_t)p.group_tiles;
    const size_t groups_per_scan = ((size_t)cast_tiles_per_scan(p.nray) + gt - 1) / gt;
    const size_t ngroup = groups_per_scan * (size_t)scans_per_chunk;
    B2_CUDA(s->d_cl_count.reserve(ngroup));
    B2_CUDA(s->d_cl_ovf.reserve(ngroup));
    B2_CUDA(s->d_cl_list.reserve(ngroup * kCandCap));
    p.cl.count = s->d_cl_count.ptr;
    p.cl.ovf_list = s->d_cl_ovf.ptr;
    p.cl.list = s->d_cl_list.ptr;
    p.cl.counters = s->d_counters.ptr + 8;
  }
  if (split) {
    const size_t aff_cap = (size_t)scans_per_chunk * p.nray;
    const size_t pair_cap = std::min<size_t>(2 * aff_cap + 1024, 0x7fffffff);
    if (aff_cap > 0x7fffffff) return fail(B2_ERR_INVALID, "scan too large for the split mesh path");
    B2_CUDA(s->d_pair_slot.reserve(pair_cap));
    B2_CUDA(s->d_pair_geom.reserve(pair_cap));
    B2_CUDA(s->d_pair_x.reserve(pair_cap));
    B2_CUDA(s->d_aff_ray.reserve(aff_cap));
    B2_CUDA(s->d_aff_bits.reserve(aff_cap));
    B2_CUDA(s->d_aff_prim.reserve(aff_cap));
    B2_CUDA(s->d_aff_prim_id.reserve(aff_cap));
    B2_CUDA(s->d_aff_mesh_id.reserve(aff_cap));
    p.ms.pair_slot = s->d_pair_slot.ptr;
    p.ms.pair_geom = s->d_pair_geom.ptr;
    p.ms.pair_x = s->d_pair_x.ptr;
    p.ms.aff_ray = s->d_aff_ray.ptr;
    p.ms.aff_bits = s->d_aff_bits.ptr;
    p.ms.aff_prim = s->d_aff_prim.ptr;
    p.ms.aff_prim_id = s->d_aff_prim_id.ptr;
    p.ms.aff_mesh_id = s->d_aff_mesh_id.ptr;
    p.ms.counters = s->d_counters.ptr;
    p.ms.pair_cap = (int32_t)pair_cap;
    p.ms.aff_cap = (int32_t)aff_cap;
  } else {
    p.ms = MeshSplit{};
  }
  const CastParams base = p;
  for (int s0 = 0; s0 < base.nscan; s0 += scans_per_chunk) {
    const int ns = std::min(scans_per_chunk, base.nscan - s0);
    const size_t roff = (size_t)s0 * base.nray;
    p.nscan = ns;
    p.scan_world = base.scan_world ? base.scan_world + s0 : nullptr;
    p.spos = base.spos + 3 * (size_t)s0;
    p.smat = base.smat ? base.smat + 9 * (size_t)s0 : nullptr;
    p.ranges = base.ranges ? base.ranges + roff : nullptr;
    p.geomid = base.geomid ? base.geomid + roff : nullptr;
    p.geomid16 = base.geomid16 ? base.geomid16 + roff : nullptr;
    p.dist = base.dist ? base.dist + roff : nullptr;
    p.xyz = base.xyz ? base.xyz + 3 * roff : nullptr;
    p.gl_depth = base.gl_depth ? base.gl_depth + roff : nullptr;
    p.scan0 = base.scan0 + s0;  // scans without a world table map to world (scan0 + s) % nworld
    bool stage;
    int tpb;
    cast_shape(s, p, stage, tpb);
    int nl = 0;
    B2_CUDA(launch_cast(p, stage, tpb, s->sm_count, stream, &nl));
    g_launches += nl;
    if (two && env_int("B2_DEBUG_LISTS", 0) != 0) {
      // diagnostics only (blocks): how many groups the collect pass handed to the fused fall-back
      int32_t c[4] = {0, 0, 0, 0};
      cudaStreamSynchronize(stream);
      cudaMemcpy(c, p.cl.counters, sizeof(c), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[b2] two-kernel cast: %d scans x %d rays, %d groups fell back to the fused kernel\n",
              p.nscan, p.nray, c[0]);
      if (p.ms.counters) {
        cudaMemcpy(c, p.ms.counters, sizeof(c), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[b2] split mesh path: %d (ray, mesh) pairs on %d rays, overflow %d\n", c[0], c[1], c[2]);
      }
    }
  }
  return B2_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------
extern "C" {

int b2_version(void) { return B2_VERSION; }

int b2_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char* b2_last_error(void) { return g_err.c_str(); }

long long b2_launch_count(void) { return g_launches.load(); }

// ---------------------------------------------------------------------------------------------
// scene
// ---------------------------------------------------------------------------------------------
int b2_scene_create(const b2_model_desc* m, int32_t nworld, int32_t device, void* stream,
                    b2_scene** out) {
  if (!m || !out) return fail(B2_ERR_INVALID, "b2_scene_create: null argument");
  *out = nullptr;
  if (m->ngeom < 0 || m->nbody < 1 || m->nmesh < 0 || m->nmat < 0 || nworld < 1)
    return fail(B2_ERR_INVALID, "b2_scene_create: bad counts");
  if (m->ngeom >= (1 << 24))  // candidate records pack the geom id into 24 bits
    return fail(B2_ERR_UNSUPPORTED, "b2_scene_create: more than 16M geoms");
  if (m->ngeom > 0 && (!m->geom_type || !m->geom_bodyid || !m->geom_group || !m->geom_matid ||
                       !m->geom_dataid || !m->geom_size || !m->geom_rgba))
    return fail(B2_ERR_INVALID, "b2_scene_create: null geom array");
  if (!m->body_weldid) return fail(B2_ERR_INVALID, "b2_scene_create: null body_weldid");
  if (m->nmat > 0 && !m->mat_rgba) return fail(B2_ERR_INVALID, "b2_scene_create: null mat_rgba");
  if (m->nmesh > 0 && (!m->mesh_vertadr || !m->mesh_vertnum || !m->mesh_faceadr ||
                       !m->mesh_facenum || !m->mesh_vert || !m->mesh_face))
    return fail(B2_ERR_INVALID, "b2_scene_create: null mesh array");
  for (int g = 0; g < m->ngeom; g++) {
    const int t = m->geom_type[g];
    if (t == GEOM_HFIELD || t < 0 || t > GEOM_MESH)
      return fail(B2_ERR_UNSUPPORTED, "geom type not supported by the ray engine (hfield/sdf)");
    if (m->geom_bodyid[g] < 0 || m->geom_bodyid[g] >= m->nbody)
      return fail(B2_ERR_INVALID, "geom_bodyid out of range");
    if (t == GEOM_MESH && (m->geom_dataid[g] < 0 || m->geom_dataid[g] >= m->nmesh))
      return fail(B2_ERR_INVALID, "mesh geom with bad geom_dataid");
    if (m->geom_matid[g] >= m->nmat) return fail(B2_ERR_INVALID, "geom_matid out of range");
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B2_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device index out of range");
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(B2_ERR_CUDA, "device is not sm_100 class: libb2sense is built for sm_100a only");

  b2_scene* s = new b2_scene();
  s->device = device;
  s->sm_count = prop.multiProcessorCount;
  s->ngeom = m->ngeom;
  s->nbody = m->nbody;
  s->nmesh = m->nmesh;
  s->nmat = m->nmat;
  s->nworld = nworld;
  auto cleanup_fail = [&](int code) {
    b2_scene_destroy(s);
    return code;
  };
  if (stream) {
    s->stream = (cudaStream_t)stream;
  } else {
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess)
      return cleanup_fail(fail(B2_ERR_CUDA, "cudaStreamCreate failed"));
    s->own_stream = true;
  }
  if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
    return cleanup_fail(fail(B2_ERR_CUDA, "cudaStreamCreate failed"));
  for (int i = 0; i < 2; i++)
    if (cudaEventCreateWithFlags(&s->stage_done[i], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&s->refit_done[i], cudaEventDisableTiming) != cudaSuccess)
      return cleanup_fail(fail(B2_ERR_CUDA, "cudaEventCreate failed"));
  if (cudaEventCreateWithFlags(&s->pm.upload_done, cudaEventDisableTiming) != cudaSuccess)
    return cleanup_fail(fail(B2_ERR_CUDA, "cudaEventCreate failed"));

  const size_t G = (size_t)m->ngeom;
  HostModel& hm = s->hm;
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

  std::vector<GeomStatic> gs(G);
  for (size_t g = 0; g < G; g++) {
    for (int k = 0; k < 3; k++) gs[g].size[k] = m->geom_size[3 * g + k];
    gs[g].type = m->geom_type[g];
    gs[g].dataid = m->geom_dataid[g];
    if (m->geom_type[g] == GEOM_MESH) s->has_mesh_geoms = true;
    if (m->geom_type[g] == GEOM_PLANE)
      hm.plane_ids.push_back((int32_t)g);
    else
      hm.bvh_geoms.push_back((int32_t)g);
  }
  auto up = [&](cudaError_t e) { return e == cudaSuccess; };
  bool ok = up(s->d_gstat.reserve(G + 1)) && up(s->d_plane_ids.reserve(hm.plane_ids.size() + 1)) &&
            up(s->d_pose.reserve(G * nworld + 1));
  if (ok && G > 0)
    ok = up(cudaMemcpy(s->d_gstat.ptr, gs.data(), G * sizeof(GeomStatic), cudaMemcpyHostToDevice));
  if (ok && !hm.plane_ids.empty())
    ok = up(cudaMemcpy(s->d_plane_ids.ptr, hm.plane_ids.data(),
                       hm.plane_ids.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (!ok) {
    cudaError_t e = cudaGetLastError();
    return cleanup_fail(fail(B2_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e)));
  }
  int rc = build_meshes(s, m);
  if (rc != B2_OK) return cleanup_fail(rc);
  *out = s;
  return B2_OK;
}

int b2_scene_destroy(b2_scene* s) {
  if (!s) return B2_OK;
  if (s->dependents.load() > 0)
    return fail(B2_ERR_INVALID, "b2_scene_destroy: lidars / cameras of this scene are still alive");
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->copy_stream) {
    cudaStreamSynchronize(s->copy_stream);
    cudaStreamDestroy(s->copy_stream);
  }
  s->d_gstat.release();
  s->d_plane_ids.release();
  s->d_mesh.release();
  s->d_mnodes.release();
  s->d_mtris.release();
  s->d_pose.release();
  s->d_wslots.release();
  s->d_wnodes.release();
  s->d_slot_link.release();
  s->d_refit_slots.release();
  s->d_level_start.release();
  s->d_raw.release();
  s->d_pair_slot.release();
  s->d_pair_geom.release();
  s->d_pair_x.release();
  s->d_aff_ray.release();
  s->d_aff_bits.release();
  s->d_aff_prim.release();
  s->d_aff_prim_id.release();
  s->d_aff_mesh_id.release();
  s->d_counters.release();
  s->d_cl_count.release();
  s->d_cl_ovf.release();
  s->d_cl_list.release();
  s->d_mr_vec.release();
  s->d_mr_dist.release();
  s->d_mr_pnt.release();
  s->d_mr_id.release();
  s->d_mr_elim.release();
  for (int i = 0; i < 2; i++) {
    s->h_stage[i].release();
    if (s->stage_done[i]) cudaEventDestroy(s->stage_done[i]);
    if (s->refit_done[i]) cudaEventDestroy(s->refit_done[i]);
  }
  for (int i = 0; i < 2; i++)
    if (s->sync_mark[i]) cudaEventDestroy(s->sync_mark[i]);
  if (s->copy_fork) cudaEventDestroy(s->copy_fork);
  if (s->copy_join) cudaEventDestroy(s->copy_join);
  for (int i = 0; i < 2; i++)
    if (s->copy_fence[i]) cudaEventDestroy(s->copy_fence[i]);
  for (int i = 0; i < 3; i++) s->pm.slot[i].release();
  if (s->pm.upload_done) cudaEventDestroy(s->pm.upload_done);
  if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return B2_OK;
}

}  // extern "C"

// Upload the newest slot of the "latest wins" host mirror (if any) and refit.  Called by every
// consumer with the scene mutex held.  H2D rides the copy stream; the refit waits for it on the
// scene's stream; the upload into a d_raw half waits for the refit that last read that half.
static int flush_pending(b2_scene* s) {
  // never inside a stream capture: an upload + refit baked into a replayed graph would re-install
  // stale poses on every replay.  The caller flushed before it began capturing; whatever the
  // producer left since then is picked up by the next consumer.
  if (s->capturing) return B2_OK;
  b2_scene::PoseMirror& pm = s->pm;
  {
    std::lock_guard<std::mutex> g(pm.idx);
    if (pm.version == pm.uploaded) return B2_OK;
  }
  if (pm.upload_pending) {
    // the previous upload must have consumed its slot before the slot indices move on
    B2_CUDA(cudaEventSynchronize(pm.upload_done));
    pm.upload_pending = false;
  }
  int slot;
  unsigned long long v;
  {
    std::lock_guard<std::mutex> g(pm.idx);
    slot = pm.latest;
    pm.uploading = slot;
    v = pm.version;
  }
  const size_t G = (size_t)s->ngeom, W = (size_t)s->nworld;
  const size_t n3 = G * 3 * W, n9 = G * 9 * W, half = G * 12 * W;
  B2_CUDA(s->d_raw.reserve(2 * half));
  const int f = s->raw_flip;
  s->raw_flip ^= 1;
  double* d_dst = s->d_raw.ptr + (size_t)f * half;
  B2_CUDA(cudaStreamWaitEvent(s->copy_stream, s->refit_done[f], 0));
  B2_CUDA(cudaMemcpyAsync(d_dst, pm.slot[slot].ptr, (n3 + n9) * sizeof(double),
                          cudaMemcpyHostToDevice, s->copy_stream));
  B2_CUDA(cudaEventRecord(pm.upload_done, s->copy_stream));
  pm.upload_pending = true;
  B2_CUDA(cudaStreamWaitEvent(s->stream, pm.upload_done, 0));
  int rc = run_refit(s, 0, s->nworld, d_dst, d_dst + n3);
  if (rc != B2_OK) return rc;
  B2_CUDA(cudaEventRecord(s->refit_done[f], s->stream));
  pm.uploaded = v;
  return B2_OK;
}

// Worlds that have not been updated yet mirror the poses of the world the topology was built on, so
// that a cast against them walks valid boxes (their results are that world's, not garbage).
static int seed_all_worlds(b2_scene* s, int src_world) {
  if (s->nworld <= 1 || s->ngeom == 0) return B2_OK;
  const size_t G = (size_t)s->ngeom;
  // world `src_world` first into world 0, then doubling copies
  if (src_world != 0)
    B2_CUDA(cudaMemcpyAsync(s->d_pose.ptr, s->d_pose.ptr + G * src_world, G * sizeof(GeomPose),
                            cudaMemcpyDeviceToDevice, s->stream));
  for (size_t have = 1; have < (size_t)s->nworld; have *= 2) {
    const size_t n = std::min(have, (size_t)s->nworld - have);
    B2_CUDA(cudaMemcpyAsync(s->d_pose.ptr + G * have, s->d_pose.ptr, n * G * sizeof(GeomPose),
                            cudaMemcpyDeviceToDevice, s->stream));
  }
  return run_refit(s, 0, s->nworld, nullptr, nullptr);
}

static int update_poses_common(b2_scene* s, int32_t world0, int32_t nworld, const double* xpos,
                               const double* xmat, bool on_device) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  if (!xpos || !xmat) return fail(B2_ERR_INVALID, "null pose array");
  if (world0 < 0 || nworld < 1 || world0 + nworld > s->nworld)
    return fail(B2_ERR_INVALID, "world range out of bounds");
  if (s->ngeom == 0) return B2_OK;
  const size_t G = (size_t)s->ngeom;
  const size_t n3 = G * 3 * nworld, n9 = G * 9 * nworld;

  // ---- the per-step refresh of the plugin (host arrays, all worlds, topology already built):
  //      copy into a pinned slot and return -- no CUDA call, no scene mutex, never waits for the GPU
  if (!on_device && world0 == 0 && nworld == s->nworld && s->topo_built && s->pm.slot[2].ptr) {
    b2_scene::PoseMirror& pm = s->pm;
    std::lock_guard<std::mutex> prod(pm.prod);
    int w = 0;
    {
      std::lock_guard<std::mutex> g(pm.idx);
      while (w == pm.latest || w == pm.uploading) w++;
    }
    std::memcpy(pm.slot[w].ptr, xpos, n3 * sizeof(double));
    std::memcpy(pm.slot[w].ptr + n3, xmat, n9 * sizeof(double));
    {
      std::lock_guard<std::mutex> g(pm.idx);
      pm.latest = w;
      pm.version++;
    }
    return B2_OK;
  }

  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(cudaSetDevice(s->device));
  {
    int rc = flush_pending(s);  // keep the order of updates
    if (rc != B2_OK) return rc;
  }
  const double* d_xpos = xpos;
  const double* d_xmat = xmat;
  const double* h_first_pos = xpos;
  const double* h_first_mat = xmat;
  std::vector<double> first;  // first world's poses on the host for the topology build
  int f = -1;
  if (!on_device) {
    // pinned double buffer: the caller may reuse its arrays as soon as we return.  The upload
    // rides the copy stream; the refit waits for it on the scene's stream.
    const int hf = s->stage_flip;
    s->stage_flip ^= 1;
    B2_CUDA(s->h_stage[hf].reserve(n3 + n9));
    B2_CUDA(cudaEventSynchronize(s->stage_done[hf]));
    std::memcpy(s->h_stage[hf].ptr, xpos, n3 * sizeof(double));
    std::memcpy(s->h_stage[hf].ptr + n3, xmat, n9 * sizeof(double));
    const size_t half = G * 12 * (size_t)s->nworld;
    B2_CUDA(s->d_raw.reserve(2 * half));
    f = s->raw_flip;
    s->raw_flip ^= 1;
    double* d_dst = s->d_raw.ptr + (size_t)f * half;
    B2_CUDA(cudaStreamWaitEvent(s->copy_stream, s->refit_done[f], 0));
    B2_CUDA(cudaMemcpyAsync(d_dst, s->h_stage[hf].ptr, (n3 + n9) * sizeof(double),
                            cudaMemcpyHostToDevice, s->copy_stream));
    B2_CUDA(cudaEventRecord(s->stage_done[hf], s->copy_stream));
    B2_CUDA(cudaStreamWaitEvent(s->stream, s->stage_done[hf], 0));
    d_xpos = d_dst;
    d_xmat = d_dst + n3;
  } else if (!s->topo_built) {
    first.resize(G * 12);
    B2_CUDA(cudaMemcpyAsync(first.data(), xpos, G * 3 * sizeof(double), cudaMemcpyDeviceToHost,
                            s->stream));
    B2_CUDA(cudaMemcpyAsync(first.data() + G * 3, xmat, G * 9 * sizeof(double),
                            cudaMemcpyDeviceToHost, s->stream));
    B2_CUDA(cudaStreamSynchronize(s->stream));
    h_first_pos = first.data();
    h_first_mat = first.data() + G * 3;
  }
  const bool first_update = !s->topo_built;
  if (first_update) {
    int rc = build_topology(s, h_first_pos, h_first_mat);
    if (rc != B2_OK) return rc;
  }
  int rc = run_refit(s, world0, nworld, d_xpos, d_xmat);
  if (rc != B2_OK) return rc;
  if (f >= 0) B2_CUDA(cudaEventRecord(s->refit_done[f], s->stream));
  if (first_update) {
    if (nworld < s->nworld) {
      rc = seed_all_worlds(s, world0);
      if (rc != B2_OK) return rc;
      // the caller's worlds again, on top of the seed
      rc = run_refit(s, world0, nworld, d_xpos, d_xmat);
      if (rc != B2_OK) return rc;
      if (f >= 0) B2_CUDA(cudaEventRecord(s->refit_done[f], s->stream));
    }
    // the pinned slots of the "latest wins" mirror, allocated once, outside the per-step path
    // (bounded: bulk multi-world simulations feed device arrays and never need them)
    const size_t slot_elems = G * 12 * (size_t)s->nworld;
    if (!on_device && slot_elems * sizeof(double) <= ((size_t)64 << 20))
      for (int i = 0; i < 3; i++) B2_CUDA(s->pm.slot[i].reserve(slot_elems));
  }
  return B2_OK;
}

extern "C" {

int b2_scene_update_poses(b2_scene* s, int32_t world0, int32_t nworld, const double* geom_xpos,
                          const double* geom_xmat) {
  return update_poses_common(s, world0, nworld, geom_xpos, geom_xmat, false);
}

int b2_scene_update_poses_device(b2_scene* s, int32_t world0, int32_t nworld,
                                 const double* d_geom_xpos, const double* d_geom_xmat) {
  return update_poses_common(s, world0, nworld, d_geom_xpos, d_geom_xmat, true);
}

int b2_scene_rebuild(b2_scene* s, int32_t world) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  if (world < 0 || world >= s->nworld) return fail(B2_ERR_INVALID, "world out of range");
  if (!s->topo_built) return fail(B2_ERR_INVALID, "b2_scene_rebuild before any pose update");
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(cudaSetDevice(s->device));
  if (s->ngeom == 0) return B2_OK;
  {
    int rc = flush_pending(s);
    if (rc != B2_OK) return rc;
  }
  const size_t G = (size_t)s->ngeom;
  std::vector<GeomPose> hp(G);
  B2_CUDA(cudaMemcpyAsync(hp.data(), s->d_pose.ptr + G * world, G * sizeof(GeomPose),
                          cudaMemcpyDeviceToHost, s->stream));
  B2_CUDA(cudaStreamSynchronize(s->stream));
  std::vector<double> xpos(G * 3), xmat(G * 9);
  for (size_t g = 0; g < G; g++) {
    for (int k = 0; k < 3; k++) xpos[3 * g + k] = hp[g].pos[k];
    for (int k = 0; k < 9; k++) xmat[9 * g + k] = hp[g].mat[k];
  }
  int rc = build_topology(s, xpos.data(), xmat.data());
  if (rc != B2_OK) return rc;
  return run_refit(s, 0, s->nworld, nullptr, nullptr);
}

int b2_scene_sync(b2_scene* s) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  B2_CUDA(cudaSetDevice(s->device));
  {
    // markers are recorded under the scene mutex and waited for outside it: another thread may be
    // capturing the scene's stream meanwhile, and a stream-wide synchronize would break its capture
    std::lock_guard<std::recursive_mutex> lock(s->mtx);
    int rc = flush_pending(s);
    if (rc != B2_OK) return rc;
    for (int i = 0; i < 2; i++)
      if (!s->sync_mark[i]) B2_CUDA(cudaEventCreateWithFlags(&s->sync_mark[i], cudaEventDisableTiming));
    B2_CUDA(cudaEventRecord(s->sync_mark[0], s->stream));
    B2_CUDA(cudaEventRecord(s->sync_mark[1], s->copy_stream));
  }
  B2_CUDA(cudaEventSynchronize(s->sync_mark[0]));
  B2_CUDA(cudaEventSynchronize(s->sync_mark[1]));
  return B2_OK;
}

int b2_scene_stats(const b2_scene* s, int32_t* n_bvh_nodes, int32_t* n_planes,
                   int64_t* n_mesh_nodes, int64_t* n_mesh_tris) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  if (n_bvh_nodes) *n_bvh_nodes = s->nslot;
  if (n_planes) *n_planes = (int32_t)s->hm.plane_ids.size();
  if (n_mesh_nodes) *n_mesh_nodes = s->n_mesh_nodes;
  if (n_mesh_tris) *n_mesh_tris = s->n_mesh_tris;
  return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// lidar
// ---------------------------------------------------------------------------------------------
namespace {

// Tiles per collect group for a ray table: the largest of 4 / 2 / 1 whose groups all fit a cone of
// half-angle 10 deg.  The 64x2048 table gives 3 deg at 4 tiles per group; a coarse table (64x256:
// 23 deg at 4) would hand a third of its groups' over-long candidate lists to the fall-back.
// Measured (tools/coarse_probe.py, 8.4 M rays, Grays/s at 4 / 2 / 1 tiles per group):
//   64x128: 2.19 / 2.82 / 3.90   64x256: 3.48 / 5.07 / 4.86   64x512: 5.98 / 6.19 / 5.54
//   64x2048: 7.91 / 7.43 / 6.29.
// Sampled on up to 64 groups spread over the table, same tile -> ray mapping as the kernels.
int choose_group_tiles(const double* dir, int nray, int tile_w, size_t bvh_geoms) {
  const int forced = env_int("B2_GROUP_TILES_RT", 0);
  if (forced == 1 || forced == 2 || forced == 4) return forced;
  if (nray <= 0 || !dir) return kGroupTiles;
  // a scene of a few geoms has nothing to cull: narrower groups would only repeat the collect
  // (4096 arm worlds x 1024-ray planar lidar: 24.5 Grays/s at 4 tiles per group, 17.7 at 1)
  if (bvh_geoms <= 64) return kGroupTiles;
  int tiles_x = 0, tiles = (nray + 31) / 32;
  if (tile_w > 0 && tile_w % 8 == 0 && nray % tile_w == 0) {
    const int rows = nray / tile_w;
    if (rows >= 4 && rows % 4 == 0) {
      tiles_x = tile_w / 8;
      tiles = tiles_x * (rows / 4);
    }
  }
  auto ray_of = [&](int tile, int lane) {
    if (tiles_x > 0) {
      const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
      return (ty * 4 + (lane >> 3)) * tile_w + tx * 8 + (lane & 7);
    }
    return tile * 32 + lane;
  };
  const double cos_limit = std::cos(10.0 * 3.14159265358979323846 / 180.0);
  for (int gt = kGroupTiles; gt >= 2; gt >>= 1) {
    const int groups = (tiles + gt - 1) / gt;
    const int step = std::max(1, groups / 64);
    bool ok = true;
    for (int g = 0; g < groups && ok; g += step) {
      double ax = 0, ay = 0, az = 0;
      std::vector<double> u;
      for (int t = g * gt; t < std::min(tiles, (g + 1) * gt); t++)
        for (int lane = 0; lane < 32; lane++) {
          const int r = ray_of(t, lane);
          if (r >= nray) continue;
          const double x = dir[3 * (size_t)r], y = dir[3 * (size_t)r + 1], z = dir[3 * (size_t)r + 2];
          const double n = std::sqrt(x * x + y * y + z * z);
          if (!(n > 0)) continue;
          u.push_back(x / n), u.push_back(y / n), u.push_back(z / n);
          ax += x / n, ay += y / n, az += z / n;
        }
      const double an = std::sqrt(ax * ax + ay * ay + az * az);
      if (u.empty()) continue;
      if (!(an > 1e-9)) {
        ok = false;
        break;
      }
      for (size_t i = 0; i < u.size() && ok; i += 3)
        if ((u[i] * ax + u[i + 1] * ay + u[i + 2] * az) / an < cos_limit) ok = false;
    }
    if (ok) return gt;
  }
  return 1;
}

}  // namespace

int b2_lidar_create(b2_scene* s, const b2_lidar_desc* d, b2_lidar** out) {
  if (!s || !d || !out) return fail(B2_ERR_INVALID, "b2_lidar_create: null argument");
  *out = nullptr;
  if (d->nray < 0 || (d->nray > 0 && !d->local_dir))
    return fail(B2_ERR_INVALID, "b2_lidar_create: bad ray table");
  if (d->bodyexclude >= s->nbody) return fail(B2_ERR_INVALID, "bodyexclude out of range");
  B2_CUDA(cudaSetDevice(s->device));
  b2_lidar* l = new b2_lidar();
  l->scene = s;
  l->nray = d->nray;
  l->tile_w = d->tile_w;
  l->group_tiles = choose_group_tiles(d->local_dir, d->nray, d->tile_w, s->hm.bvh_geoms.size());
  l->cutoff = d->cutoff;
  l->range_min = d->range_min;
  l->range_max = d->range_max;
  std::vector<uint8_t> elim;
  eliminate_flags(s->hm, d->geomgroup, d->flg_static, d->bodyexclude, elim);
  cudaError_t e = l->d_local_dir.reserve((size_t)d->nray * 3 + 1);
  if (e == cudaSuccess) e = l->d_elim.reserve(elim.size() + 1);
  if (e == cudaSuccess && d->nray > 0)
    e = cudaMemcpy(l->d_local_dir.ptr, d->local_dir, (size_t)d->nray * 3 * sizeof(double),
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess && !elim.empty())
    e = cudaMemcpy(l->d_elim.ptr, elim.data(), elim.size(), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l->all_done, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreate(&l->cast_t0);
  if (e == cudaSuccess) e = cudaEventCreate(&l->cast_t1);
  if (e != cudaSuccess) {
    b2_lidar_destroy(l);
    return fail(B2_ERR_CUDA, std::string("b2_lidar_create: ") + cudaGetErrorString(e));
  }
  s->dependents++;
  l->counted = true;
  *out = l;
  return B2_OK;
}

int b2_lidar_destroy(b2_lidar* l) {
  if (!l) return B2_OK;
  cudaSetDevice(l->scene->device);
  {
    std::lock_guard<std::recursive_mutex> lock(l->scene->mtx);  // nobody is capturing the stream
    cudaStreamSynchronize(l->scene->stream);
    cudaStreamSynchronize(l->scene->copy_stream);
  }
  l->d_local_dir.release();
  l->d_elim.release();
  l->d_spos.release();
  l->d_in.release();
  l->d_smat.release();
  l->d_dist.release();
  l->d_world.release();
  l->d_geomid.release();
  l->d_ranges.release();
  l->d_xyz.release();
  l->d_geomid16.release();
  l->h_in.release();
  l->h_out.release();
  for (cudaEvent_t ev : l->piece_done)
    if (ev) cudaEventDestroy(ev);
  if (l->all_done) cudaEventDestroy(l->all_done);
  if (l->graph.exec) cudaGraphExecDestroy(l->graph.exec);
  if (l->cast_t0) cudaEventDestroy(l->cast_t0);
  if (l->cast_t1) cudaEventDestroy(l->cast_t1);
  if (l->counted) l->scene->dependents--;
  delete l;
  return B2_OK;
}

// scan0: index of the first of these scans in the caller's batch (scans without a world table map
// to world (scan0 + s) % nworld)
static int lidar_launch(b2_lidar* l, int32_t nscan, int32_t scan0, const int32_t* d_world,
                        const double* d_pos, const double* d_mat, float* d_ranges,
                        int32_t* d_geomid, uint16_t* d_geomid16, double* d_dist, float* d_xyz,
                        cudaStream_t stream) {
  b2_scene* s = l->scene;
  if (nscan == 0 || l->nray == 0) return B2_OK;
  if (s->ngeom > 0 && !s->topo_built)
    return fail(B2_ERR_INVALID, "cast before the first b2_scene_update_poses");
  if (d_geomid16 && s->ngeom >= 65535)
    return fail(B2_ERR_INVALID, "uint16 geom ids need a model with fewer than 65535 geoms");
  CastParams p{};
  fill_cast_scene(s, p);
  p.elim = l->d_elim.ptr;
  p.local_dir = l->d_local_dir.ptr;
  p.nray = l->nray;
  p.tile_w = l->tile_w;
  p.group_tiles = l->group_tiles;
  p.mode = MODE_LIDAR;
  p.cutoff = l->cutoff;
  p.range_min = l->range_min;
  p.range_max = l->range_max;
  p.nscan = nscan;
  p.scan0 = scan0;
  p.scan_world = d_world;
  p.spos = d_pos;
  p.smat = d_mat;
  p.ranges = d_ranges;
  p.geomid = d_geomid;
  p.geomid16 = d_geomid16;
  p.dist = d_dist;
  p.xyz = d_xyz;
  return run_cast(s, p, stream);
}

int b2_lidar_cast_device(b2_lidar* l, int32_t nscan, const int32_t* d_scan_world,
                         const double* d_sensor_pos, const double* d_sensor_mat, float* d_ranges,
                         int32_t* d_geomid, double* d_dist, float* d_xyz) {
  if (!l) return fail(B2_ERR_INVALID, "null lidar");
  if (nscan < 0) return fail(B2_ERR_INVALID, "negative nscan");
  if (nscan > 0 && (!d_sensor_pos || !d_sensor_mat)) return fail(B2_ERR_INVALID, "null sensor pose");
  B2_CUDA(cudaSetDevice(l->scene->device));
  return lidar_launch(l, nscan, 0, d_scan_world, d_sensor_pos, d_sensor_mat, d_ranges, d_geomid,
                      nullptr, d_dist, d_xyz, l->scene->stream);
}

int b2_lidar_cast_u16_device(b2_lidar* l, int32_t nscan, const int32_t* d_scan_world,
                             const double* d_sensor_pos, const double* d_sensor_mat,
                             float* d_ranges, uint16_t* d_geomid16) {
  if (!l) return fail(B2_ERR_INVALID, "null lidar");
  if (nscan < 0) return fail(B2_ERR_INVALID, "negative nscan");
  if (nscan > 0 && (!d_sensor_pos || !d_sensor_mat)) return fail(B2_ERR_INVALID, "null sensor pose");
  B2_CUDA(cudaSetDevice(l->scene->device));
  return lidar_launch(l, nscan, 0, d_scan_world, d_sensor_pos, d_sensor_mat, d_ranges, nullptr,
                      d_geomid16, nullptr, nullptr, l->scene->stream);
}

namespace {

bool host_ptr_is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a{};
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

constexpr long long kSmallCastRays = 1 << 20;  // below this a host cast is one piece (and a graph)

// the enqueue of one host cast, small path: H2D of the poses from pinned staging, the cast, D2H of
// the results into pinned staging -- all on the scene's stream, all addresses fixed per lidar
int enqueue_small(b2_lidar* l, const HostCast& hc, bool has_world, cudaStream_t stream) {
  const size_t ns = (size_t)hc.nscan, n = hc.n;
  // sensor frames (and the world table): one copy of the staging block, laid out pos | mat | world
  double* d_in = l->d_in.ptr;
  B2_CUDA(cudaMemcpyAsync(d_in, l->h_in.ptr, ns * 12 * sizeof(double) + (has_world ? ns * sizeof(int32_t) : 0),
                          cudaMemcpyHostToDevice, stream));
  int rc = lidar_launch(l, hc.nscan, 0, has_world ? reinterpret_cast<const int32_t*>(d_in + ns * 12) : nullptr,
                        d_in, d_in + ns * 3, hc.ranges ? l->d_ranges.ptr : nullptr,
                        hc.geomid ? l->d_geomid.ptr : nullptr,
                        hc.geomid16 ? l->d_geomid16.ptr : nullptr, hc.dist ? l->d_dist.ptr : nullptr,
                        hc.xyz ? l->d_xyz.ptr : nullptr, stream);
  if (rc != B2_OK) return rc;
  unsigned char* o = l->h_out.ptr;
  if (hc.ranges)
    B2_CUDA(cudaMemcpyAsync(o + l->off_ranges, l->d_ranges.ptr, n * sizeof(float),
                            cudaMemcpyDeviceToHost, stream));
  if (hc.geomid)
    B2_CUDA(cudaMemcpyAsync(o + l->off_geomid, l->d_geomid.ptr, n * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, stream));
  if (hc.geomid16)
    B2_CUDA(cudaMemcpyAsync(o + l->off_geomid16, l->d_geomid16.ptr, n * sizeof(uint16_t),
                            cudaMemcpyDeviceToHost, stream));
  if (hc.dist)
    B2_CUDA(cudaMemcpyAsync(o + l->off_dist, l->d_dist.ptr, n * sizeof(double),
                            cudaMemcpyDeviceToHost, stream));
  if (hc.xyz)
    B2_CUDA(cudaMemcpyAsync(o + l->off_xyz, l->d_xyz.ptr, n * 3 * sizeof(float),
                            cudaMemcpyDeviceToHost, stream));
  return B2_OK;
}

int lidar_cast_begin(b2_lidar* l, int32_t nscan, const int32_t* scan_world, const double* sensor_pos,
                     const double* sensor_mat, float* ranges, int32_t* geomid, uint16_t* geomid16,
                     double* dist, float* xyz) {
  if (!l) return fail(B2_ERR_INVALID, "null lidar");
  if (l->pend.active) return fail(B2_ERR_INVALID, "b2_lidar_cast_begin: the previous cast was not ended");
  if (nscan < 0) return fail(B2_ERR_INVALID, "negative nscan");
  HostCast hc;
  hc.nscan = nscan;
  hc.n = (size_t)nscan * l->nray;
  hc.ranges = ranges;
  hc.geomid = geomid;
  hc.geomid16 = geomid16;
  hc.dist = dist;
  hc.xyz = xyz;
  if (nscan == 0 || l->nray == 0) {
    hc.active = true;
    l->pend = hc;
    return B2_OK;
  }
  if (!sensor_pos || !sensor_mat) return fail(B2_ERR_INVALID, "null sensor pose");
  b2_scene* s = l->scene;
  if (geomid16 && s->ngeom >= 65535)
    return fail(B2_ERR_INVALID, "uint16 geom ids need a model with fewer than 65535 geoms");
  if (s->ngeom > 0 && !s->topo_built)
    return fail(B2_ERR_INVALID, "cast before the first b2_scene_update_poses");
  B2_CUDA(cudaSetDevice(s->device));
  // One scene-mutex hold for the whole enqueue: every piece of this cast sees the same pose
  // snapshot (a concurrent refresh lands before or after all of them in stream order), and a
  // caller that holds its own simulation mutex across this call gets sensor frames and geom poses
  // of the same step.  Nothing below waits for the GPU.
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  {
    int rc = flush_pending(s);
    if (rc != B2_OK) return rc;
  }
  const size_t n = hc.n, ns = (size_t)nscan;
  B2_CUDA(l->d_spos.reserve(ns * 3));
  B2_CUDA(l->d_smat.reserve(ns * 9));
  if (scan_world) B2_CUDA(l->d_world.reserve(ns));
  if (ranges) B2_CUDA(l->d_ranges.reserve(n));
  if (geomid) B2_CUDA(l->d_geomid.reserve(n));
  if (geomid16) B2_CUDA(l->d_geomid16.reserve(n));
  if (dist) B2_CUDA(l->d_dist.reserve(n));
  if (xyz) B2_CUDA(l->d_xyz.reserve(n * 3));
  hc.small = (long long)n <= env_int("B2_SMALL_CAST_RAYS", (int)kSmallCastRays);
  hc.staged = hc.small || !(host_ptr_is_pinned(ranges) && host_ptr_is_pinned(geomid) &&
                            host_ptr_is_pinned(geomid16) && host_ptr_is_pinned(dist) &&
                            host_ptr_is_pinned(xyz));
  if (hc.staged) {
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    l->off_ranges = off, off += ranges ? up(n * sizeof(float)) : 0;
    l->off_geomid = off, off += geomid ? up(n * sizeof(int32_t)) : 0;
    l->off_geomid16 = off, off += geomid16 ? up(n * sizeof(uint16_t)) : 0;
    l->off_dist = off, off += dist ? up(n * sizeof(double)) : 0;
    l->off_xyz = off, off += xyz ? up(n * 3 * sizeof(float)) : 0;
    B2_CUDA(l->h_out.reserve(off + 256));
  }
  B2_CUDA(cudaEventRecord(l->cast_t0, s->stream));
  if (hc.small) {
    // ---- small path: one piece; the whole sequence is captured into a CUDA graph the second time
    //      the same shape is asked for (a single scan is ~6 launches + copies: launch-gap bound)
    B2_CUDA(l->h_in.reserve(ns * 13 + 16));
    B2_CUDA(l->d_in.reserve(ns * 13 + 16));
    std::memcpy(l->h_in.ptr, sensor_pos, ns * 3 * sizeof(double));
    std::memcpy(l->h_in.ptr + ns * 3, sensor_mat, ns * 9 * sizeof(double));
    if (scan_world) std::memcpy(l->h_in.ptr + ns * 12, scan_world, ns * sizeof(int32_t));
    const int mask = (ranges ? 1 : 0) | (geomid ? 2 : 0) | (geomid16 ? 4 : 0) | (dist ? 8 : 0) |
                     (xyz ? 16 : 0);
    LidarGraph& g = l->graph;
    const bool same = g.nscan == nscan && g.mask == mask && g.topo_version == s->topo_version &&
                      g.has_world == (scan_world != nullptr) && g.alloc_epoch == g_alloc_epoch.load();
    const bool use_graph = env_int("B2_GRAPH", 1) != 0;
    if (use_graph && same && g.exec) {
      B2_CUDA(cudaGraphLaunch(g.exec, s->stream));
      g_launches += g.kernels;
    } else if (use_graph && same && g.seen) {
      cudaGraph_t graph = nullptr;
      const long long l0 = g_launches.load();
      B2_CUDA(cudaStreamBeginCapture(s->stream, cudaStreamCaptureModeThreadLocal));
      s->capturing = true;
      int rc = enqueue_small(l, hc, scan_world != nullptr, s->stream);
      s->capturing = false;
      cudaError_t e = cudaStreamEndCapture(s->stream, &graph);
      g.kernels = (int)(g_launches.load() - l0);  // counted once here = the first replay
      if (rc != B2_OK || e != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        g.seen = false;  // do not try again for this shape
        g.nscan = -1;
        rc = enqueue_small(l, hc, scan_world != nullptr, s->stream);
        if (rc != B2_OK) return rc;
      } else {
        e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
          g.exec = nullptr;
          cudaGetLastError();
          rc = enqueue_small(l, hc, scan_world != nullptr, s->stream);
          if (rc != B2_OK) return rc;
        } else {
          B2_CUDA(cudaGraphLaunch(g.exec, s->stream));
        }
      }
    } else {
      if (g.exec) {
        cudaGraphExecDestroy(g.exec);
        g.exec = nullptr;
      }
      g.nscan = nscan;
      g.mask = mask;
      g.topo_version = s->topo_version;
      g.has_world = scan_world != nullptr;
      g.seen = true;
      int rc = enqueue_small(l, hc, scan_world != nullptr, s->stream);
      if (rc != B2_OK) return rc;
      g.alloc_epoch = g_alloc_epoch.load();  // after the plain run: every buffer it needs exists now
    }
    B2_CUDA(cudaEventRecord(l->cast_t1, s->stream));
    B2_CUDA(cudaEventRecord(l->all_done, s->stream));
    hc.active = true;
    l->pend = hc;
    return B2_OK;
  }

  // ---- bulk path: cast piece k on the scene's stream while piece k-1 drains over PCIe on the copy
  //      stream (the result read-back is up to 8 B/ray and would otherwise serialise behind the kernel)
  B2_CUDA(cudaMemcpyAsync(l->d_spos.ptr, sensor_pos, ns * 3 * sizeof(double), cudaMemcpyHostToDevice,
                          s->stream));
  B2_CUDA(cudaMemcpyAsync(l->d_smat.ptr, sensor_mat, ns * 9 * sizeof(double), cudaMemcpyHostToDevice,
                          s->stream));
  if (scan_world)
    B2_CUDA(cudaMemcpyAsync(l->d_world.ptr, scan_world, ns * sizeof(int32_t), cudaMemcpyHostToDevice,
                            s->stream));
  // How many pieces: when the read-back is the longer leg it should start early and never wait
  // (16 pieces: 6.3 of the 6.85 Grays/s the 54.8 GB/s link allows at 8 B/ray; 4 / 8 / 32 gave
  // 5.8 / 6.1 / 5.8); when the cast is the longer leg every extra piece only adds launch tails
  // (ranges only, 4 B/ray: 4 / 8 / 16 gave 8.0 / 7.8 / 7.2).  The cast time per ray is known from
  // the previous call on this lidar; the link is taken as 50 GB/s.
  const double out_bytes = (ranges ? 4.0 : 0.0) + (geomid ? 4.0 : 0.0) + (geomid16 ? 2.0 : 0.0) +
                           (dist ? 8.0 : 0.0) + (xyz ? 12.0 : 0.0);
  const double copy_ns_per_ray = out_bytes / 50.0;
  const double cast_ns_per_ray = l->cast_ns_per_ray > 0 ? l->cast_ns_per_ray : 0.12;
  const bool copy_bound = copy_ns_per_ray > cast_ns_per_ray;
  int nchunk = env_int("B2_E2E_CHUNKS", copy_bound ? 16 : 0);
  if (nchunk > nscan) nchunk = nscan;
  std::vector<int>& bounds = hc.bounds;
  if (nchunk >= 1) {
    // read-back bound: equal pieces, except that the first and the last one are cut into quarters:
    // nothing overlaps the cast of the first piece nor the copy of the last one
    for (int c = 0; c < nchunk; c++) {
      const int b0 = (int)((long long)nscan * c / nchunk), b1 = (int)((long long)nscan * (c + 1) / nchunk);
      if (bounds.empty() || b0 > bounds.back()) bounds.push_back(b0);
      if (nchunk > 1 && (c == 0 || c == nchunk - 1)) {
        for (int q = 1; q < 4; q++) {
          const int b = b0 + (int)((long long)(b1 - b0) * q / 4);
          if (b > bounds.back() && b < nscan) bounds.push_back(b);
        }
      }
    }
  } else {
    // cast bound: every extra piece costs the launch tails of its kernels (~0.2 ms), and only the
    // LAST piece's read-back is exposed: three quarters, then halving pieces -- the first copy starts
    // after a quarter of the cast, the exposed one is a sixteenth of the data
    static const int kNum[] = {0, 4, 8, 12, 14, 15};
    for (int k = 0; k < 6; k++) {
      const int b = (int)((long long)nscan * kNum[k] / 16);
      if (bounds.empty() || b > bounds.back()) bounds.push_back(b);
    }
  }
  if (bounds.empty()) bounds.push_back(0);
  bounds.push_back(nscan);
  while (l->piece_done.size() + 1 < bounds.size()) {
    cudaEvent_t ev = nullptr;
    B2_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    l->piece_done.push_back(ev);
  }
  unsigned char* o = l->h_out.ptr;
  for (size_t c = 0; c + 1 < bounds.size(); c++) {
    const int s0 = bounds[c], s1 = bounds[c + 1];
    const int npc = s1 - s0;
    const size_t off = (size_t)s0 * l->nray, cnt = (size_t)npc * l->nray;
    int rc = lidar_launch(l, npc, s0, scan_world ? l->d_world.ptr + s0 : nullptr,
                          l->d_spos.ptr + 3 * (size_t)s0, l->d_smat.ptr + 9 * (size_t)s0,
                          ranges ? l->d_ranges.ptr + off : nullptr,
                          geomid ? l->d_geomid.ptr + off : nullptr,
                          geomid16 ? l->d_geomid16.ptr + off : nullptr,
                          dist ? l->d_dist.ptr + off : nullptr, xyz ? l->d_xyz.ptr + 3 * off : nullptr,
                          s->stream);
    if (rc != B2_OK) return rc;
    cudaEvent_t ev = l->piece_done[c];
    B2_CUDA(cudaEventRecord(ev, s->stream));
    B2_CUDA(cudaStreamWaitEvent(s->copy_stream, ev, 0));
    const bool st = hc.staged;
    if (ranges)
      B2_CUDA(cudaMemcpyAsync(st ? (void*)(o + l->off_ranges + off * 4) : (void*)(ranges + off),
                              l->d_ranges.ptr + off, cnt * sizeof(float), cudaMemcpyDeviceToHost,
                              s->copy_stream));
    if (geomid)
      B2_CUDA(cudaMemcpyAsync(st ? (void*)(o + l->off_geomid + off * 4) : (void*)(geomid + off),
                              l->d_geomid.ptr + off, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost,
                              s->copy_stream));
    if (geomid16)
      B2_CUDA(cudaMemcpyAsync(st ? (void*)(o + l->off_geomid16 + off * 2) : (void*)(geomid16 + off),
                              l->d_geomid16.ptr + off, cnt * sizeof(uint16_t), cudaMemcpyDeviceToHost,
                              s->copy_stream));
    if (dist)
      B2_CUDA(cudaMemcpyAsync(st ? (void*)(o + l->off_dist + off * 8) : (void*)(dist + off),
                              l->d_dist.ptr + off, cnt * sizeof(double), cudaMemcpyDeviceToHost,
                              s->copy_stream));
    if (xyz)
      B2_CUDA(cudaMemcpyAsync(st ? (void*)(o + l->off_xyz + off * 12) : (void*)(xyz + 3 * off),
                              l->d_xyz.ptr + 3 * off, cnt * 3 * sizeof(float), cudaMemcpyDeviceToHost,
                              s->copy_stream));
    // the same event, re-recorded on the copy stream: piece c has landed on the host
    B2_CUDA(cudaEventRecord(ev, s->copy_stream));
  }
  B2_CUDA(cudaEventRecord(l->cast_t1, s->stream));
  hc.active = true;
  l->pend = hc;
  return B2_OK;
}

int lidar_cast_end(b2_lidar* l) {
  if (!l) return fail(B2_ERR_INVALID, "null lidar");
  if (!l->pend.active) return fail(B2_ERR_INVALID, "b2_lidar_cast_end without b2_lidar_cast_begin");
  HostCast hc = l->pend;
  l->pend = HostCast{};
  if (hc.nscan == 0 || l->nray == 0) return B2_OK;
  b2_scene* s = l->scene;
  B2_CUDA(cudaSetDevice(s->device));
  const unsigned char* o = l->h_out.ptr;
  const size_t n = hc.n;
  if (hc.small) {
    B2_CUDA(cudaEventSynchronize(l->all_done));
    if (hc.ranges) std::memcpy(hc.ranges, o + l->off_ranges, n * sizeof(float));
    if (hc.geomid) std::memcpy(hc.geomid, o + l->off_geomid, n * sizeof(int32_t));
    if (hc.geomid16) std::memcpy(hc.geomid16, o + l->off_geomid16, n * sizeof(uint16_t));
    if (hc.dist) std::memcpy(hc.dist, o + l->off_dist, n * sizeof(double));
    if (hc.xyz) std::memcpy(hc.xyz, o + l->off_xyz, n * 3 * sizeof(float));
  } else {
    for (size_t c = 0; c + 1 < hc.bounds.size(); c++) {
      B2_CUDA(cudaEventSynchronize(l->piece_done[c]));
      if (!hc.staged) continue;
      // pageable destination: drain the landed piece while the later ones are still in flight
      const size_t off = (size_t)hc.bounds[c] * l->nray;
      const size_t cnt = (size_t)(hc.bounds[c + 1] - hc.bounds[c]) * l->nray;
      if (hc.ranges) std::memcpy(hc.ranges + off, o + l->off_ranges + off * 4, cnt * 4);
      if (hc.geomid) std::memcpy(hc.geomid + off, o + l->off_geomid + off * 4, cnt * 4);
      if (hc.geomid16) std::memcpy(hc.geomid16 + off, o + l->off_geomid16 + off * 2, cnt * 2);
      if (hc.dist) std::memcpy(hc.dist + off, o + l->off_dist + off * 8, cnt * 8);
      if (hc.xyz) std::memcpy(hc.xyz + 3 * off, o + l->off_xyz + off * 12, cnt * 12);
    }
  }
  float cast_ms = 0.f;
  if (cudaEventElapsedTime(&cast_ms, l->cast_t0, l->cast_t1) == cudaSuccess && cast_ms > 0.f) {
    l->cast_ns_per_ray = (double)cast_ms * 1e6 / (double)n;
    l->last_device_us = (double)cast_ms * 1e3;
  } else {
    cudaGetLastError();
  }
  return B2_OK;
}

}  // namespace

int b2_lidar_cast_begin(b2_lidar* l, int32_t nscan, const int32_t* scan_world,
                        const double* sensor_pos, const double* sensor_mat, float* ranges,
                        int32_t* geomid, double* dist, float* xyz) {
  return lidar_cast_begin(l, nscan, scan_world, sensor_pos, sensor_mat, ranges, geomid, nullptr, dist,
                          xyz);
}

int b2_lidar_cast_end(b2_lidar* l) { return lidar_cast_end(l); }

int b2_lidar_cast(b2_lidar* l, int32_t nscan, const int32_t* scan_world, const double* sensor_pos,
                  const double* sensor_mat, float* ranges, int32_t* geomid, double* dist,
                  float* xyz) {
  int rc = lidar_cast_begin(l, nscan, scan_world, sensor_pos, sensor_mat, ranges, geomid, nullptr,
                            dist, xyz);
  if (rc != B2_OK) return rc;
  return lidar_cast_end(l);
}

int b2_lidar_cast_u16(b2_lidar* l, int32_t nscan, const int32_t* scan_world,
                      const double* sensor_pos, const double* sensor_mat, float* ranges,
                      uint16_t* geomid16) {
  int rc = lidar_cast_begin(l, nscan, scan_world, sensor_pos, sensor_mat, ranges, nullptr, geomid16,
                            nullptr, nullptr);
  if (rc != B2_OK) return rc;
  return lidar_cast_end(l);
}

double b2_lidar_last_device_us(const b2_lidar* l) { return l ? l->last_device_us : 0.0; }

// ---------------------------------------------------------------------------------------------
// mj_multiRay-shaped entry
// ---------------------------------------------------------------------------------------------
int b2_multiray(b2_scene* s, int32_t world, const double pnt[3], const double* vec,
                const uint8_t* geomgroup, uint8_t flg_static, int32_t bodyexclude, int32_t* geomid,
                double* dist, int32_t nray, double cutoff) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  if (nray < 0) return fail(B2_ERR_INVALID, "negative nray");
  if (nray == 0) return B2_OK;
  if (!pnt || !vec || !dist) return fail(B2_ERR_INVALID, "b2_multiray: null argument");
  if (world < 0 || world >= s->nworld) return fail(B2_ERR_INVALID, "world out of range");
  if (s->ngeom > 0 && !s->topo_built)
    return fail(B2_ERR_INVALID, "b2_multiray before the first b2_scene_update_poses");
  std::lock_guard<std::recursive_mutex> lock(s->mtx);  // scene-owned scratch until the sync below
  B2_CUDA(cudaSetDevice(s->device));
  std::vector<uint8_t> elim;
  eliminate_flags(s->hm, geomgroup, flg_static, bodyexclude, elim);
  B2_CUDA(s->d_mr_elim.reserve(elim.size() + 1));
  B2_CUDA(s->d_mr_vec.reserve((size_t)nray * 3));
  B2_CUDA(s->d_mr_dist.reserve((size_t)nray));
  B2_CUDA(s->d_mr_id.reserve((size_t)nray + 1));
  B2_CUDA(s->d_mr_pnt.reserve(3));
  if (!elim.empty())
    B2_CUDA(cudaMemcpyAsync(s->d_mr_elim.ptr, elim.data(), elim.size(), cudaMemcpyHostToDevice,
                            s->stream));
  B2_CUDA(cudaMemcpyAsync(s->d_mr_vec.ptr, vec, (size_t)nray * 3 * sizeof(double),
                          cudaMemcpyHostToDevice, s->stream));
  B2_CUDA(cudaMemcpyAsync(s->d_mr_pnt.ptr, pnt, 3 * sizeof(double), cudaMemcpyHostToDevice,
                          s->stream));
  int32_t* d_world = s->d_mr_id.ptr + nray;  // one spare slot holds the world index
  B2_CUDA(cudaMemcpyAsync(d_world, &world, sizeof(int32_t), cudaMemcpyHostToDevice, s->stream));
  CastParams p{};
  fill_cast_scene(s, p);
  p.elim = s->d_mr_elim.ptr;
  p.local_dir = s->d_mr_vec.ptr;
  p.nray = nray;
  p.tile_w = 0;
  p.mode = MODE_MULTIRAY;
  p.cutoff = cutoff;
  p.nscan = 1;
  p.scan_world = d_world;
  p.spos = s->d_mr_pnt.ptr;
  p.smat = nullptr;
  p.geomid = s->d_mr_id.ptr;
  p.dist = s->d_mr_dist.ptr;
  {
    int rc = run_cast(s, p, s->stream);
    if (rc != B2_OK) return rc;
  }
  B2_CUDA(cudaMemcpyAsync(dist, s->d_mr_dist.ptr, (size_t)nray * sizeof(double),
                          cudaMemcpyDeviceToHost, s->stream));
  if (geomid)
    B2_CUDA(cudaMemcpyAsync(geomid, s->d_mr_id.ptr, (size_t)nray * sizeof(int32_t),
                            cudaMemcpyDeviceToHost, s->stream));
  B2_CUDA(cudaStreamSynchronize(s->stream));
  return B2_OK;
}

// ---------------------------------------------------------------------------------------------
// camera
// ---------------------------------------------------------------------------------------------
int b2_camera_create(b2_scene* scene, int32_t device, const b2_camera_desc* d, b2_camera** out) {
  if (!d || !out) return fail(B2_ERR_INVALID, "b2_camera_create: null argument");
  *out = nullptr;
  if (d->width < 1 || d->height < 1) return fail(B2_ERR_INVALID, "bad camera resolution");
  if (d->point_step != 12 && d->point_step != 16)
    return fail(B2_ERR_INVALID, "point_step must be 12 or 16");
  if (!(d->fovy > 0 && d->fovy < 180)) return fail(B2_ERR_INVALID, "bad fovy");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B2_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  }
  if (scene) device = scene->device;
  if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device index out of range");
  B2_CUDA(cudaSetDevice(device));
  b2_camera* c = new b2_camera();
  c->scene = scene;
  c->device = device;
  c->desc = *d;
  if (scene) {
    c->stream = scene->stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
      delete c;
      return fail(B2_ERR_CUDA, "cudaStreamCreate failed");
    }
    c->own_stream = true;
  }
  // ref:src/mujoco_cameras.cpp:119-124
  for (int i = 0; i < 9; i++) c->K[i] = 0.0;
  const double focal_scaling = (1.0 / std::tan((d->fovy * M_PI / 180.0) / 2.0)) * d->height / 2.0;
  c->K[0] = focal_scaling;
  c->K[2] = (double)d->width / 2.0;
  c->K[4] = focal_scaling;
  c->K[5] = (double)d->height / 2.0;
  c->K[8] = 1.0;
  // ref:src/mujoco_cameras.cpp:222-224
  c->near_f = (float)(d->znear * d->extent);
  c->far_f = (float)(d->zfar * d->extent);
  c->depth_scale = 1.0f - c->near_f / c->far_f;
  if (scene) scene->dependents++;
  *out = c;
  return B2_OK;
}

int b2_camera_destroy(b2_camera* c) {
  if (!c) return B2_OK;
  cudaSetDevice(c->device);
  {
    std::unique_lock<std::recursive_mutex> lock;
    if (c->scene) lock = std::unique_lock<std::recursive_mutex>(c->scene->mtx);
    if (c->stream) cudaStreamSynchronize(c->stream);
  }
  c->d_in.release();
  c->d_depth.release();
  c->d_cloud.release();
  c->d_elim.release();
  c->d_pos.release();
  c->d_mat.release();
  c->d_world.release();
  c->h_in.release();
  c->h_out.release();
  if (c->done) cudaEventDestroy(c->done);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->scene) c->scene->dependents--;
  delete c;
  return B2_OK;
}

int b2_camera_info(const b2_camera* c, double K[9], double P[12]) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (K)
    for (int i = 0; i < 9; i++) K[i] = c->K[i];
  if (P) {
    for (int i = 0; i < 12; i++) P[i] = 0.0;
    P[0] = c->K[0];
    P[2] = c->K[2];
    P[5] = c->K[4];
    P[6] = c->K[5];
    P[10] = 1.0;
  }
  return B2_OK;
}

int b2_depth_to_cloud_device(b2_camera* c, int32_t ncam, const float* d_gl_depth,
                             float* d_depth32fc1, void* d_cloud) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (ncam < 0) return fail(B2_ERR_INVALID, "negative ncam");
  if (ncam == 0) return B2_OK;
  if (!d_gl_depth) return fail(B2_ERR_INVALID, "null depth buffer");
  B2_CUDA(cudaSetDevice(c->device));
  // a camera bound to a scene enqueues on the scene's stream: serialise with the scene's other users
  std::unique_lock<std::recursive_mutex> lock;
  if (c->scene) lock = std::unique_lock<std::recursive_mutex>(c->scene->mtx);
  DepthParams p{};
  p.gl_depth = d_gl_depth;
  p.depth_out = d_depth32fc1;
  p.cloud = (float*)d_cloud;
  p.ncam = ncam;
  p.width = c->desc.width;
  p.height = c->desc.height;
  p.point_floats = c->desc.point_step / 4;
  p.znear = c->near_f;
  p.zfar = c->far_f;
  p.depth_scale = c->depth_scale;
  p.cx = (float)c->K[2];
  p.cy = (float)c->K[5];
  p.inv_fx = 1.0f / (float)c->K[0];
  p.inv_fy = 1.0f / (float)c->K[4];
  B2_CUDA(launch_depth_to_cloud(p, c->stream));
  g_launches++;
  return B2_OK;
}

int b2_depth_to_cloud(b2_camera* c, int32_t ncam, const float* gl_depth, float* depth32fc1,
                      void* cloud) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (ncam < 0) return fail(B2_ERR_INVALID, "negative ncam");
  if (ncam == 0) return B2_OK;
  if (!gl_depth) return fail(B2_ERR_INVALID, "null depth buffer");
  B2_CUDA(cudaSetDevice(c->device));
  const size_t npix = (size_t)ncam * c->desc.width * c->desc.height;
  const size_t pf = (size_t)c->desc.point_step / 4;
  if (!c->done) B2_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  {
    // a camera bound to a scene shares the scene's stream: enqueue everything under the scene mutex,
    // wait on an event outside it (no stream-wide synchronize: see b2_scene_sync)
    std::unique_lock<std::recursive_mutex> lock;
    if (c->scene) lock = std::unique_lock<std::recursive_mutex>(c->scene->mtx);
    B2_CUDA(c->d_in.reserve(npix));
    if (depth32fc1) B2_CUDA(c->d_depth.reserve(npix));
    if (cloud) B2_CUDA(c->d_cloud.reserve(npix * pf));
    B2_CUDA(cudaMemcpyAsync(c->d_in.ptr, gl_depth, npix * sizeof(float), cudaMemcpyHostToDevice,
                            c->stream));
    int rc = b2_depth_to_cloud_device(c, ncam, c->d_in.ptr, depth32fc1 ? c->d_depth.ptr : nullptr,
                                      cloud ? c->d_cloud.ptr : nullptr);
    if (rc != B2_OK) return rc;
    if (depth32fc1)
      B2_CUDA(cudaMemcpyAsync(depth32fc1, c->d_depth.ptr, npix * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
    if (cloud)
      B2_CUDA(cudaMemcpyAsync(cloud, c->d_cloud.ptr, npix * pf * sizeof(float),
                              cudaMemcpyDeviceToHost, c->stream));
    B2_CUDA(cudaEventRecord(c->done, c->stream));
  }
  B2_CUDA(cudaEventSynchronize(c->done));
  return B2_OK;
}

int b2_camera_render_depth_device(b2_camera* c, int32_t ncam, const int32_t* d_cam_world,
                                  const double* d_cam_pos, const double* d_cam_mat,
                                  float* d_gl_depth) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (!c->scene) return fail(B2_ERR_INVALID, "camera was created without a scene");
  if (ncam < 0) return fail(B2_ERR_INVALID, "negative ncam");
  if (ncam == 0) return B2_OK;
  if (!d_cam_pos || !d_cam_mat || !d_gl_depth) return fail(B2_ERR_INVALID, "null argument");
  b2_scene* s = c->scene;
  if (s->ngeom > 0 && !s->topo_built)
    return fail(B2_ERR_INVALID, "render before the first b2_scene_update_poses");
  B2_CUDA(cudaSetDevice(s->device));
  if (c->d_elim.cap == 0) {
    // what mjv_updateScene draws with the reference's mjv_defaultOption
    // (ref:src/mujoco_cameras.cpp:165): geom groups 0-2 on, 3-5 off (collision geoms live in group 3
    // in the converter's output), alpha-0 geoms invisible; no body / static exclusion
    static const uint8_t kDefaultGroups[B2_NGROUP] = {1, 1, 1, 0, 0, 0};
    std::vector<uint8_t> elim;
    eliminate_flags(s->hm, c->desc.geomgroup_set ? c->desc.geomgroup : kDefaultGroups, 1, -1, elim);
    B2_CUDA(c->d_elim.reserve(elim.size() + 1));
    if (!elim.empty())
      B2_CUDA(cudaMemcpy(c->d_elim.ptr, elim.data(), elim.size(), cudaMemcpyHostToDevice));
  }
  CastParams p{};
  fill_cast_scene(s, p);
  p.elim = c->d_elim.ptr;
  p.local_dir = nullptr;
  p.nray = c->desc.width * c->desc.height;
  p.tile_w = c->desc.width;
  p.mode = MODE_CAMERA;
  p.cam_w = c->desc.width;
  p.cam_h = c->desc.height;
  p.cam_f = c->K[0];
  p.cam_cx = c->K[2];
  p.cam_cy = c->K[5];
  p.cam_near = c->near_f;
  p.cam_far = c->far_f;
  p.nscan = ncam;
  p.scan_world = d_cam_world;
  p.spos = d_cam_pos;
  p.smat = d_cam_mat;
  p.gl_depth = d_gl_depth;
  return run_cast(s, p, c->stream);
}

int b2_rgb_flip_device(b2_camera* c, int32_t ncam, const uint8_t* d_rgb_gl, uint8_t* d_rgb_out) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (ncam < 0) return fail(B2_ERR_INVALID, "negative ncam");
  if (ncam == 0) return B2_OK;
  if (!d_rgb_gl || !d_rgb_out) return fail(B2_ERR_INVALID, "null image buffer");
  B2_CUDA(cudaSetDevice(c->device));
  // shares the scene's stream: never enqueue while another thread captures it into a graph
  std::unique_lock<std::recursive_mutex> lock;
  if (c->scene) lock = std::unique_lock<std::recursive_mutex>(c->scene->mtx);
  B2_CUDA(launch_rgb_flip(d_rgb_gl, d_rgb_out, ncam, c->desc.width, c->desc.height, c->stream));
  g_launches++;
  return B2_OK;
}

int b2_camera_render_to_cloud_begin(b2_camera* c, int32_t ncam, const int32_t* cam_world,
                                    const double* cam_pos, const double* cam_mat, float* gl_depth,
                                    float* depth32fc1, void* cloud) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (!c->scene) return fail(B2_ERR_INVALID, "camera was created without a scene");
  if (c->pend.active) return fail(B2_ERR_INVALID, "render_to_cloud_begin: the previous render was not ended");
  if (ncam < 0) return fail(B2_ERR_INVALID, "negative ncam");
  c->pend.ncam = ncam;
  c->pend.gl_depth = gl_depth;
  c->pend.depth32fc1 = depth32fc1;
  c->pend.cloud = cloud;
  if (ncam == 0) {
    c->pend.active = true;
    return B2_OK;
  }
  if (!cam_pos || !cam_mat) return fail(B2_ERR_INVALID, "null camera pose");
  b2_scene* s = c->scene;
  B2_CUDA(cudaSetDevice(c->device));
  if (!c->done) B2_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
  const size_t npix = (size_t)ncam * c->desc.width * c->desc.height;
  const size_t pf = (size_t)c->desc.point_step / 4;
  const size_t nc = (size_t)ncam;
  // one scene-mutex hold for the whole enqueue (poses, render, linearise + cloud, read-back): see
  // lidar_cast_begin.  Results are staged through pinned memory; _end copies them out.
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(c->d_pos.reserve(nc * 3));
  B2_CUDA(c->d_mat.reserve(nc * 9));
  B2_CUDA(c->d_in.reserve(npix));
  if (depth32fc1) B2_CUDA(c->d_depth.reserve(npix));
  if (cloud) B2_CUDA(c->d_cloud.reserve(npix * pf));
  B2_CUDA(c->h_in.reserve(nc * 13 + 16));
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t off = 0;
  c->pend.off_gl = off, off += gl_depth ? up(npix * 4) : 0;
  c->pend.off_depth = off, off += depth32fc1 ? up(npix * 4) : 0;
  c->pend.off_cloud = off, off += cloud ? up(npix * pf * 4) : 0;
  B2_CUDA(c->h_out.reserve(off + 256));
  std::memcpy(c->h_in.ptr, cam_pos, nc * 3 * sizeof(double));
  std::memcpy(c->h_in.ptr + nc * 3, cam_mat, nc * 9 * sizeof(double));
  int32_t* d_world = nullptr;
  if (cam_world) {
    B2_CUDA(c->d_world.reserve(nc));
    std::memcpy(c->h_in.ptr + nc * 12, cam_world, nc * sizeof(int32_t));
    B2_CUDA(cudaMemcpyAsync(c->d_world.ptr, c->h_in.ptr + nc * 12, nc * sizeof(int32_t),
                            cudaMemcpyHostToDevice, c->stream));
    d_world = c->d_world.ptr;
  }
  B2_CUDA(cudaMemcpyAsync(c->d_pos.ptr, c->h_in.ptr, nc * 3 * sizeof(double), cudaMemcpyHostToDevice,
                          c->stream));
  B2_CUDA(cudaMemcpyAsync(c->d_mat.ptr, c->h_in.ptr + nc * 3, nc * 9 * sizeof(double),
                          cudaMemcpyHostToDevice, c->stream));
  int rc = b2_camera_render_depth_device(c, ncam, d_world, c->d_pos.ptr, c->d_mat.ptr, c->d_in.ptr);
  if (rc != B2_OK) return rc;
  if (depth32fc1 || cloud) {
    rc = b2_depth_to_cloud_device(c, ncam, c->d_in.ptr, depth32fc1 ? c->d_depth.ptr : nullptr,
                                  cloud ? c->d_cloud.ptr : nullptr);
    if (rc != B2_OK) return rc;
  }
  unsigned char* o = c->h_out.ptr;
  if (gl_depth)
    B2_CUDA(cudaMemcpyAsync(o + c->pend.off_gl, c->d_in.ptr, npix * sizeof(float),
                            cudaMemcpyDeviceToHost, c->stream));
  if (depth32fc1)
    B2_CUDA(cudaMemcpyAsync(o + c->pend.off_depth, c->d_depth.ptr, npix * sizeof(float),
                            cudaMemcpyDeviceToHost, c->stream));
  if (cloud)
    B2_CUDA(cudaMemcpyAsync(o + c->pend.off_cloud, c->d_cloud.ptr, npix * pf * sizeof(float),
                            cudaMemcpyDeviceToHost, c->stream));
  B2_CUDA(cudaEventRecord(c->done, c->stream));
  c->pend.active = true;
  return B2_OK;
}

int b2_camera_render_to_cloud_end(b2_camera* c) {
  if (!c) return fail(B2_ERR_INVALID, "null camera");
  if (!c->pend.active) return fail(B2_ERR_INVALID, "render_to_cloud_end without render_to_cloud_begin");
  c->pend.active = false;
  if (c->pend.ncam == 0) return B2_OK;
  B2_CUDA(cudaSetDevice(c->device));
  B2_CUDA(cudaEventSynchronize(c->done));
  const size_t npix = (size_t)c->pend.ncam * c->desc.width * c->desc.height;
  const size_t pf = (size_t)c->desc.point_step / 4;
  const unsigned char* o = c->h_out.ptr;
  if (c->pend.gl_depth) std::memcpy(c->pend.gl_depth, o + c->pend.off_gl, npix * 4);
  if (c->pend.depth32fc1) std::memcpy(c->pend.depth32fc1, o + c->pend.off_depth, npix * 4);
  if (c->pend.cloud) std::memcpy(c->pend.cloud, o + c->pend.off_cloud, npix * pf * 4);
  return B2_OK;
}

int b2_camera_render_to_cloud(b2_camera* c, int32_t ncam, const int32_t* cam_world,
                              const double* cam_pos, const double* cam_mat, float* gl_depth,
                              float* depth32fc1, void* cloud) {
  int rc = b2_camera_render_to_cloud_begin(c, ncam, cam_world, cam_pos, cam_mat, gl_depth, depth32fc1,
                                           cloud);
  if (rc != B2_OK) return rc;
  return b2_camera_render_to_cloud_end(c);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU: peer buffers (CUDA IPC).  The gather itself is the cast kernels' result stores.
// ---------------------------------------------------------------------------------------------
static_assert(sizeof(cudaIpcMemHandle_t) == B2_PEER_HANDLE_BYTES, "handle size");

int b2_peer_buffer_create(int32_t device, size_t bytes, void** d_ptr, uint8_t* handle) {
  if (!d_ptr || !handle || bytes == 0) return fail(B2_ERR_INVALID, "b2_peer_buffer_create: bad argument");
  *d_ptr = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B2_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device index out of range");
  B2_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  B2_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(B2_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  std::memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return B2_OK;
}

int b2_peer_buffer_open(int32_t device, const uint8_t* handle, void** d_ptr) {
  if (!d_ptr || !handle) return fail(B2_ERR_INVALID, "b2_peer_buffer_open: null argument");
  *d_ptr = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(B2_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "device index out of range");
  B2_CUDA(cudaSetDevice(device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  B2_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return B2_OK;
}

int b2_scene_copy_async(b2_scene* s, void* d_dst, const void* d_src, size_t bytes) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  if (bytes == 0) return B2_OK;
  if (!d_dst || !d_src) return fail(B2_ERR_INVALID, "b2_scene_copy_async: null pointer");
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(cudaSetDevice(s->device));
  if (!s->copy_fork) B2_CUDA(cudaEventCreateWithFlags(&s->copy_fork, cudaEventDisableTiming));
  B2_CUDA(cudaEventRecord(s->copy_fork, s->stream));
  B2_CUDA(cudaStreamWaitEvent(s->copy_stream, s->copy_fork, 0));
  B2_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, s->copy_stream));
  return B2_OK;
}

int b2_scene_join(b2_scene* s) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(cudaSetDevice(s->device));
  if (!s->copy_join) B2_CUDA(cudaEventCreateWithFlags(&s->copy_join, cudaEventDisableTiming));
  B2_CUDA(cudaEventRecord(s->copy_join, s->copy_stream));
  B2_CUDA(cudaStreamWaitEvent(s->stream, s->copy_join, 0));
  return B2_OK;
}

int b2_scene_copy_fence(b2_scene* s) {
  if (!s) return fail(B2_ERR_INVALID, "null scene");
  std::lock_guard<std::recursive_mutex> lock(s->mtx);
  B2_CUDA(cudaSetDevice(s->device));
  for (int i = 0; i < 2; i++)
    if (!s->copy_fence[i]) B2_CUDA(cudaEventCreateWithFlags(&s->copy_fence[i], cudaEventDisableTiming));
  const int i = s->copy_fence_flip;
  s->copy_fence_flip ^= 1;
  B2_CUDA(cudaEventRecord(s->copy_fence[i], s->copy_stream));
  if (s->copy_fence_used[i ^ 1]) B2_CUDA(cudaStreamWaitEvent(s->stream, s->copy_fence[i ^ 1], 0));
  s->copy_fence_used[i] = true;
  return B2_OK;
}

int b2_peer_buffer_close(int32_t device, void* d_ptr) {
  if (!d_ptr) return B2_OK;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return B2_OK;
}

int b2_peer_buffer_destroy(int32_t device, void* d_ptr) {
  if (!d_ptr) return B2_OK;
  B2_CUDA(cudaSetDevice(device));
  B2_CUDA(cudaFree(d_ptr));
  return B2_OK;
}

}  // extern "C"
