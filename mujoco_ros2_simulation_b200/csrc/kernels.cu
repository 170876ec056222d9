// kernels.cu -- hand-written sm_100a kernels of the sensor-readout engine.
//
//   K1 refit_kernel           pose packing + bottom-up refit of the per-world geom BVH
//   (K2, the ray-cast kernels, live in cast_kernels.cu)
//   K3 depth_to_cloud_kernel  depth linearise + vertical flip + PointCloud2 projection (streaming)
//
// Compile with -fmad=false: distances must be bit-identical to the CPU oracle (see ray_core.cuh).
// No tensor cores: nothing here is a dense contraction.
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "ray_core.cuh"

namespace b2 {

// =============================================================================================
// K1: refit
// =============================================================================================
__global__ void __launch_bounds__(128) refit_kernel(const RefitParams p) {
  const int w_local = blockIdx.x;
  const int world = p.world0 + w_local;
  GeomPose* pose = p.pose + (size_t)world * p.ngeom;
  WideChild* slots = p.wslots + (size_t)world * p.nslot;

  // 1. pack mjData geom_xpos / geom_xmat into 16-B aligned GeomPose records
  if (p.raw_xpos != nullptr) {
    const double* xp = p.raw_xpos + (size_t)w_local * p.ngeom * 3;
    const double* xm = p.raw_xmat + (size_t)w_local * p.ngeom * 9;
    // element-wise copy keeps the global reads and writes coalesced: GeomPose is 12 doubles
    for (int i = threadIdx.x; i < p.ngeom * 12; i += blockDim.x) {
      const int g = i / 12, k = i - g * 12;
      const double v = k < 3 ? xp[g * 3 + k] : xm[g * 9 + (k - 3)];
      reinterpret_cast<double*>(pose + g)[k] = v;
    }
    __syncthreads();
  }

  // 2. bottom-up over child slots: a slot of level l is a geom (box from its pose) or an inner
  //    node whose own slots all belong to levels < l (written by this CTA in an earlier pass)
  for (int l = 0; l < p.nlevel; l++) {
    const int beg = p.level_start[l], end = p.level_start[l + 1];
    for (int i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const int s = p.refit_slots[i];
      const int32_t link = p.slot_link[s];
      WideChild out;
      out.link = link;
      out.pad = 0;
      if (link < 0) {
        const int g = ~link;
        const GeomStatic gs = p.gstat[g];
        geom_aabb(gs.type, gs.size, pose[g], out.lo, out.hi);
        out.pad = gs.type;  // travels with the slot so the cast needs no lookup to dispatch
      } else {
        const WideNode wn = p.wnodes[link];
        out.lo[0] = out.lo[1] = out.lo[2] = FLT_MAX;
        out.hi[0] = out.hi[1] = out.hi[2] = -FLT_MAX;
        for (int c = wn.first; c < wn.first + wn.count; c++) {
          const WideChild ch = slots[c];
#pragma unroll
          for (int k = 0; k < 3; k++) {
            out.lo[k] = fminf(out.lo[k], ch.lo[k]);
            out.hi[k] = fmaxf(out.hi[k], ch.hi[k]);
          }
        }
      }
      slots[s] = out;
    }
    __syncthreads();
  }
}

cudaError_t launch_refit(const RefitParams& p, cudaStream_t stream) {
  if (p.nworld <= 0) return cudaSuccess;
  int threads = 32;
  while (threads < 128 && threads < p.ngeom * 2) threads *= 2;
  refit_kernel<<<p.nworld, threads, 0, stream>>>(p);
  return cudaGetLastError();
}

// =============================================================================================
// K3: depth linearise + flip + PointCloud2 projection
//   per pixel: 4 B read, 4 B (32FC1) + point_step B (cloud) written; nothing is re-read.
//   Each thread owns 4 consecutive pixels of one output row; the xyz triples are transposed
//   through shared memory so every global store instruction writes 512 contiguous bytes.
// =============================================================================================
constexpr int kDepthThreads = 256;

template <int PF>  // floats per point: 3 or 4
__global__ void __launch_bounds__(kDepthThreads) depth_to_cloud_kernel(const DepthParams p) {
  __shared__ __align__(16) float s_pts[kDepthThreads / 32][32 * 4 * PF];
  const int quads_per_row = p.width >> 2;
  const long long quads_per_cam = (long long)quads_per_row * p.height;
  const long long total = quads_per_cam * p.ncam;
  const int lane = threadIdx.x & 31, warp_id = threadIdx.x >> 5;
  const float qnan = __int_as_float(0x7fc00000);
  float* sw = s_pts[warp_id];

  // warp-contiguous blocks of 32 quads (128 pixels)
  const long long warp_base0 = ((long long)blockIdx.x * (kDepthThreads / 32) + warp_id) * 32;
  const long long warp_stride = (long long)gridDim.x * kDepthThreads;
  for (long long wb = warp_base0; wb < total; wb += warp_stride) {
    const long long q = wb + lane;
    const bool ok = q < total;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    int v = 0, u0 = 0;
    long long opix = 0;
    if (ok) {
      const int cam = (int)(q / quads_per_cam);
      const long long rem = q - (long long)cam * quads_per_cam;
      v = (int)(rem / quads_per_row);
      u0 = (int)(rem - (long long)v * quads_per_row) << 2;
      const long long cam_pix = (long long)cam * p.width * p.height;
      // GL rows are bottom-up: output row v reads GL row H-1-v (ref:src/mujoco_cameras.cpp:234)
      const float4 d = __ldcs(reinterpret_cast<const float4*>(
          p.gl_depth + cam_pix + (long long)(p.height - 1 - v) * p.width + u0));
      // ref:src/mujoco_cameras.cpp:235  near / (1.0f - d * depth_scale), float32, unfused
      z[0] = p.znear / (1.0f - d.x * p.depth_scale);
      z[1] = p.znear / (1.0f - d.y * p.depth_scale);
      z[2] = p.znear / (1.0f - d.z * p.depth_scale);
      z[3] = p.znear / (1.0f - d.w * p.depth_scale);
      opix = cam_pix + (long long)v * p.width + u0;
      if (p.depth_out)
        __stcs(reinterpret_cast<float4*>(p.depth_out + opix), make_float4(z[0], z[1], z[2], z[3]));
    }
    if (p.cloud) {
      // depth_image_proc::convertDepth<float>: x = (u - cx) * z * (1/fx), y likewise, NaN if !finite
      float pt[4 * PF];
      const float yv = (float)v - p.cy;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const bool fin = isfinite(z[i]);
        const float xu = (float)(u0 + i) - p.cx;
        pt[i * PF + 0] = fin ? xu * z[i] * p.inv_fx : qnan;
        pt[i * PF + 1] = fin ? yv * z[i] * p.inv_fy : qnan;
        pt[i * PF + 2] = fin ? z[i] : qnan;
        if (PF == 4) pt[i * PF + 3] = 0.0f;
      }
      float4* s4 = reinterpret_cast<float4*>(sw);
#pragma unroll
      for (int j = 0; j < PF; j++)
        s4[lane * PF + j] = make_float4(pt[4 * j], pt[4 * j + 1], pt[4 * j + 2], pt[4 * j + 3]);
      __syncwarp();
      // the warp's 32 quads are contiguous in the output: 32*4*PF floats from pixel wb*4
      const long long nvalid = (total - wb) < 32 ? (total - wb) : 32;  // quads
      float4* g4 = reinterpret_cast<float4*>(p.cloud + (wb * 4) * PF);
#pragma unroll
      for (int j = 0; j < PF; j++) {
        const int e = j * 32 + lane;  // float4 index within the warp's block
        if (e < nvalid * PF) __stcs(g4 + e, s4[e]);
      }
      __syncwarp();
    }
  }
}

// generic fallback for widths that are not a multiple of 4: one pixel per thread
template <int PF>
__global__ void __launch_bounds__(kDepthThreads) depth_to_cloud_scalar_kernel(const DepthParams p) {
  const long long per_cam = (long long)p.width * p.height;
  const long long total = per_cam * p.ncam;
  const float qnan = __int_as_float(0x7fc00000);
  for (long long i = (long long)blockIdx.x * kDepthThreads + threadIdx.x; i < total;
       i += (long long)gridDim.x * kDepthThreads) {
    const int cam = (int)(i / per_cam);
    const long long rem = i - (long long)cam * per_cam;
    const int v = (int)(rem / p.width), u = (int)(rem - (long long)v * p.width);
    const float d = p.gl_depth[(long long)cam * per_cam + (long long)(p.height - 1 - v) * p.width + u];
    const float z = p.znear / (1.0f - d * p.depth_scale);
    if (p.depth_out) p.depth_out[i] = z;
    if (p.cloud) {
      const bool fin = isfinite(z);
      float* o = p.cloud + i * PF;
      o[0] = fin ? ((float)u - p.cx) * z * p.inv_fx : qnan;
      o[1] = fin ? ((float)v - p.cy) * z * p.inv_fy : qnan;
      o[2] = fin ? z : qnan;
      if (PF == 4) o[3] = 0.0f;
    }
  }
}

cudaError_t launch_depth_to_cloud(const DepthParams& p, cudaStream_t stream) {
  const long long npix = (long long)p.ncam * p.width * p.height;
  if (npix <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool vec4 = (p.width % 4 == 0) && (((uintptr_t)p.gl_depth & 15) == 0) &&
                    (((uintptr_t)p.depth_out & 15) == 0) && (((uintptr_t)p.cloud & 15) == 0);
  const long long work = vec4 ? npix / 4 : npix;
  long long blocks = (work + kDepthThreads - 1) / kDepthThreads;
  const long long cap = (long long)sms * 16;  // grid-stride beyond 16 CTAs per SM
  if (blocks > cap) blocks = cap;
  if (vec4) {
    if (p.point_floats == 4)
      depth_to_cloud_kernel<4><<<(unsigned)blocks, kDepthThreads, 0, stream>>>(p);
    else
      depth_to_cloud_kernel<3><<<(unsigned)blocks, kDepthThreads, 0, stream>>>(p);
  } else {
    if (p.point_floats == 4)
      depth_to_cloud_scalar_kernel<4><<<(unsigned)blocks, kDepthThreads, 0, stream>>>(p);
    else
      depth_to_cloud_scalar_kernel<3><<<(unsigned)blocks, kDepthThreads, 0, stream>>>(p);
  }
  return cudaGetLastError();
}

// =============================================================================================
// N2: RGB8 row flip (ref:src/mujoco_cameras.cpp:243-249): GL rows are bottom-up, Image rows are
// top-down.  Pure copy, 6 B/pixel; 16-B vectors when a row is a multiple of 16 B.
// =============================================================================================
__global__ void __launch_bounds__(256) rgb_flip_kernel(const uint8_t* __restrict__ in,
                                                       uint8_t* __restrict__ out, int ncam,
                                                       int row_bytes, int height, int vec16) {
  const long long per_cam = (long long)row_bytes * height;
  if (vec16) {
    const int row_v = row_bytes >> 4;
    const long long total = (long long)ncam * height * row_v;
    const uint4* src = reinterpret_cast<const uint4*>(in);
    uint4* dst = reinterpret_cast<uint4*>(out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const long long cam = i / ((long long)height * row_v);
      const long long rem = i - cam * (long long)height * row_v;
      const int h = (int)(rem / row_v), x = (int)(rem - (long long)h * row_v);
      dst[cam * (per_cam >> 4) + (long long)(height - 1 - h) * row_v + x] = __ldcs(src + i);
    }
  } else {
    const long long total = (long long)ncam * per_cam;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const long long cam = i / per_cam;
      const long long rem = i - cam * per_cam;
      const int h = (int)(rem / row_bytes), x = (int)(rem - (long long)h * row_bytes);
      out[cam * per_cam + (long long)(height - 1 - h) * row_bytes + x] = in[i];
    }
  }
}

cudaError_t launch_rgb_flip(const uint8_t* in, uint8_t* out, int ncam, int width, int height,
                            cudaStream_t stream) {
  const int row_bytes = width * 3;
  const long long bytes = (long long)ncam * row_bytes * height;
  if (bytes <= 0) return cudaSuccess;
  const int vec16 = (row_bytes % 16 == 0) && (((uintptr_t)in & 15) == 0) && (((uintptr_t)out & 15) == 0);
  const long long work = vec16 ? bytes / 16 : bytes;
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  rgb_flip_kernel<<<(unsigned)blocks, 256, 0, stream>>>(in, out, ncam, row_bytes, height, vec16);
  return cudaGetLastError();
}

}  // namespace b2
