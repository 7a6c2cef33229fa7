// Stage kernels of the render path that are integer / indexing / HBM-bound work:
// ray generation, SMPL-box slab intersection, stratified sampling, world->pose, blend-weight volume
// sampling, forward / inverse LBS, alpha compositing, and the stable compaction they feed.
//
// Bit-exact targets (rays, near/far/mask, z_vals, sample points, pose points, trilinear pnorm) use
// _rn intrinsics in the reference's op order; see common.cuh.
#include "common.cuh"

namespace aninerf {

thread_local char g_err[512] = {0};
std::atomic<long long> g_launches{0};

// =============================================================================================
// 1. rays: get_rays (if_nerf_data_utils.py:64-89) + float32 cast (:328-329)
// =============================================================================================
struct CamDev {
  double Kinv[9], R[9], T[3], o[3];
  int H, W;
};

__global__ void __launch_bounds__(256) gen_rays_kernel(CamDev c, float *__restrict__ ray_o, float *__restrict__ ray_d) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t n = (int64_t)c.H * c.W;
  if (idx >= n) return;
  // pixel coordinates are float32 aranges in the reference (exact integers)
  double u = (double)(float)(idx % c.W), v = (double)(float)(idx / c.W);
  double cam[3];
#pragma unroll
  for (int k = 0; k < 3; ++k)   // np.dot(xy1, inv(K).T)
    cam[k] = __dadd_rn(__dadd_rn(__dmul_rn(u, c.Kinv[3 * k]), __dmul_rn(v, c.Kinv[3 * k + 1])), c.Kinv[3 * k + 2]);
  double q[3] = {__dsub_rn(cam[0], c.T[0]), __dsub_rn(cam[1], c.T[1]), __dsub_rn(cam[2], c.T[2])};
  double d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {  // np.dot(pixel_camera - T, R) - rays_o
    double w = __dadd_rn(__dadd_rn(__dmul_rn(q[0], c.R[k]), __dmul_rn(q[1], c.R[3 + k])), __dmul_rn(q[2], c.R[6 + k]));
    d[k] = __dsub_rn(w, c.o[k]);
  }
  double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[0], d[0]), __dmul_rn(d[1], d[1])), __dmul_rn(d[2], d[2])));
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ray_d[idx * 3 + k] = (float)__ddiv_rn(d[k], nrm);
    ray_o[idx * 3 + k] = (float)c.o[k];
  }
}

// =============================================================================================
// 2. near / far / mask_at_box: get_near_far (if_nerf_data_utils.py:156-196)
// =============================================================================================
struct BoxDev {
  double b[6];   // min_x,min_y,min_z,max_x,max_y,max_z after the float64 +-0.01 (:168)
};

__device__ __forceinline__ bool near_far_one(const BoxDev &B, const float of[3], const float df[3], float &near, float &far) {
  const double eps = 1e-6;
  double o[3] = {(double)of[0], (double)of[1], (double)of[2]};
  double d[3] = {(double)df[0], (double)df[1], (double)df[2]};
  double lo[3] = {__dsub_rn(B.b[0], eps), __dsub_rn(B.b[1], eps), __dsub_rn(B.b[2], eps)};
  double hi[3] = {__dadd_rn(B.b[3], eps), __dadd_rn(B.b[4], eps), __dadd_rn(B.b[5], eps)};
  int hits = 0;
  double dist[2] = {0.0, 0.0};
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    int a = p % 3;
    double t = __ddiv_rn(__dsub_rn(B.b[p], o[a]), d[a]);
    double P[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) P[k] = __dadd_rn(__dmul_rn(t, d[k]), o[k]);
    bool in = P[0] >= lo[0] && P[0] <= hi[0] && P[1] >= lo[1] && P[1] <= hi[1] && P[2] >= lo[2] && P[2] <= hi[2];
    if (in) {
      if (hits < 2) {
        double q0 = __dsub_rn(P[0], o[0]), q1 = __dsub_rn(P[1], o[1]), q2 = __dsub_rn(P[2], o[2]);
        dist[hits] = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(q0, q0), __dmul_rn(q1, q1)), __dmul_rn(q2, q2)));
      }
      ++hits;
    }
  }
  if (hits != 2) {
    near = 0.f;
    far = 0.f;
    return false;
  }
  // np.linalg.norm of the float32 direction is evaluated in float32 (:190)
  float nd = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(df[0], df[0]), __fmul_rn(df[1], df[1])), __fmul_rn(df[2], df[2])));
  double d0 = __ddiv_rn(dist[0], (double)nd), d1 = __ddiv_rn(dist[1], (double)nd);
  near = (float)fmin(d0, d1);
  far = (float)fmax(d0, d1);
  return true;
}

__global__ void __launch_bounds__(256) near_far_kernel(BoxDev B, const float *__restrict__ ray_o, const float *__restrict__ ray_d,
                                                       int64_t n, float *__restrict__ near, float *__restrict__ far,
                                                       uint8_t *__restrict__ mask) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float o[3] = {ray_o[3 * i], ray_o[3 * i + 1], ray_o[3 * i + 2]};
  float d[3] = {ray_d[3 * i], ray_d[3 * i + 1], ray_d[3 * i + 2]};
  float nr, fr;
  bool m = near_far_one(B, o, d, nr, fr);
  near[i] = nr;
  far[i] = fr;
  mask[i] = m ? 1 : 0;
}

// =============================================================================================
// 3. stable compaction by a byte mask (block counts -> single-block scan -> scatter)
// =============================================================================================
constexpr int CB = 1024;   // items per compaction block

__global__ void __launch_bounds__(256) count_mask_kernel(const uint8_t *__restrict__ mask, int64_t n, int32_t *__restrict__ block_counts) {
  __shared__ int s[8];
  int64_t base = (int64_t)blockIdx.x * CB;
  int c = 0;
  for (int k = threadIdx.x; k < CB; k += 256) {
    int64_t i = base + k;
    c += (i < n && mask[i]) ? 1 : 0;
  }
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s[w];
    block_counts[blockIdx.x] = t;
  }
}

// exclusive scan of `m` int32 counts by ONE block of 1024 threads; total -> *total
__global__ void __launch_bounds__(1024) scan_counts_kernel(const int32_t *__restrict__ counts, int64_t m, int32_t *__restrict__ offsets,
                                                           int32_t *__restrict__ total) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < m; base += 1024) {
    int64_t i = base + threadIdx.x;
    int v = i < m ? counts[i] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if ((threadIdx.x & 31) >= o) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      int ws = warp_sums[threadIdx.x];
      int winc = ws;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (threadIdx.x >= o) winc += t;
      }
      warp_sums[threadIdx.x] = winc - ws;   // exclusive warp prefix
    }
    __syncthreads();
    int excl = carry + warp_sums[threadIdx.x >> 5] + inc - v;
    if (i < m) offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(256) scatter_rays_kernel(const float *__restrict__ ray_o, const float *__restrict__ ray_d,
                                                           const float *__restrict__ near, const float *__restrict__ far,
                                                           const uint8_t *__restrict__ mask, int64_t n,
                                                           const int32_t *__restrict__ block_offsets, float *__restrict__ o_out,
                                                           float *__restrict__ d_out, float *__restrict__ near_out,
                                                           float *__restrict__ far_out, int32_t *__restrict__ index_out) {
  __shared__ int warp_base[8];
  int64_t base = (int64_t)blockIdx.x * CB;
  int running = block_offsets[blockIdx.x];
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = 0; k0 < CB; k0 += 256) {
    int64_t i = base + k0 + threadIdx.x;
    bool f = i < n && mask[i];
    unsigned bal = __ballot_sync(0xffffffffu, f);
    if (lane == 0) warp_base[warp] = __popc(bal);
    __syncthreads();
    int before = 0, tot = 0;
    for (int w = 0; w < 8; ++w) {
      int c = warp_base[w];
      if (w < warp) before += c;
      tot += c;
    }
    if (f) {
      int dst = running + before + __popc(bal & ((1u << lane) - 1u));
      if (o_out) { o_out[3 * dst] = ray_o[3 * i]; o_out[3 * dst + 1] = ray_o[3 * i + 1]; o_out[3 * dst + 2] = ray_o[3 * i + 2]; }
      if (d_out) { d_out[3 * dst] = ray_d[3 * i]; d_out[3 * dst + 1] = ray_d[3 * i + 1]; d_out[3 * dst + 2] = ray_d[3 * i + 2]; }
      if (near_out) near_out[dst] = near[i];
      if (far_out) far_out[dst] = far[i];
      if (index_out) index_out[dst] = (int32_t)i;
    }
    running += tot;
    __syncthreads();
  }
}

// =============================================================================================
// 4. sample points / z / dists (tpose_renderer.py:14-69), materialising variant
// =============================================================================================
__global__ void __launch_bounds__(256) sample_points_kernel(const float *__restrict__ ray_o, const float *__restrict__ ray_d,
                                                            const float *__restrict__ near, const float *__restrict__ far,
                                                            const float *__restrict__ t_vals, const float *__restrict__ t_rand,
                                                            int64_t n_rays, int S, float *__restrict__ pts, float *__restrict__ z_vals,
                                                            float *__restrict__ dists) {
  __shared__ float st[ANINERF_MAX_SAMPLES], s1mt[ANINERF_MAX_SAMPLES];
  if (threadIdx.x < S) {
    float t = t_vals[threadIdx.x];
    st[threadIdx.x] = t;
    s1mt[threadIdx.x] = __fsub_rn(1.0f, t);
  }
  __syncthreads();
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_rays * S) return;
  int64_t r = idx / S;
  int s = (int)(idx - r * S);
  float nr = near[r], fr = far[r];
  auto zat = [&](int k) -> float {
    float z = z_lerp(nr, fr, st[k], s1mt[k]);
    if (t_rand) {
      float zp = k > 0 ? z_lerp(nr, fr, st[k - 1], s1mt[k - 1]) : z;
      float zn = k < S - 1 ? z_lerp(nr, fr, st[k + 1], s1mt[k + 1]) : z;
      z = z_jitter(zp, z, zn, k == 0, k == S - 1, t_rand[r * S + k]);
    }
    return z;
  };
  float z = zat(s);
  if (z_vals) z_vals[idx] = z;
  if (pts) {
#pragma unroll
    for (int k = 0; k < 3; ++k) pts[idx * 3 + k] = __fadd_rn(ray_o[r * 3 + k], __fmul_rn(ray_d[r * 3 + k], z));
  }
  if (dists) dists[idx] = s < S - 1 ? __fsub_rn(zat(s + 1), z) : __fsub_rn(z, zat(s - 1));
}

// =============================================================================================
// 5. world -> pose (blend_utils.py:6-16)
// =============================================================================================
__global__ void __launch_bounds__(256) world_to_pose_kernel(const float *__restrict__ wpts, int64_t n, const float *__restrict__ R,
                                                            const float *__restrict__ Th, float *__restrict__ ppts) {
  __shared__ RigidFrame f;
  if (threadIdx.x < 9) f.R[threadIdx.x] = R[threadIdx.x];
  if (threadIdx.x < 3) f.Th[threadIdx.x] = Th[threadIdx.x];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x, y, z;
  world_to_pose(f, wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2], x, y, z);
  ppts[3 * i] = x;
  ppts[3 * i + 1] = y;
  ppts[3 * i + 2] = z;
}

// =============================================================================================
// 6. blend-weight volume sampling (blend_utils.py:119-149), all 25 channels, reference layout
// =============================================================================================
// One warp per group of 32 points.  Phase 1 (lane = point): the bit-exact trilinear corner set (weights and voxel offsets in ATen's
// order) of each point goes to shared memory.  Phase 2 (lane = channel, 25 of 32 lanes): per point the 16 words are read back with four
// broadcast LDS.128, then the 8 corner rows (100 B each) are read with coalesced 100-byte requests and accumulated in ATen's order;
// out (n,25) rows likewise.  Two points are in flight per iteration (16 independent row loads).
// Round 1 broadcast the corner set with 16 shuffles per point: ncu (profiles/r02_stage_kernels_summary.md) showed the LSU data pipe --
// which serves SHFL, LDS and the L1 wavefronts of LDG alike -- 87 % busy, two thirds of it shuffles.
constexpr int SBW_STRIDE = 20;     // words per point in shared memory (16 + 4 pad: conflict-free 128-bit stores at an 80-byte stride)
__global__ void __launch_bounds__(256) sample_bw_kernel(const float *__restrict__ pts, int64_t n, const float *__restrict__ vol,
                                                        const float *__restrict__ bounds, int X, int Y, int Z,
                                                        float *__restrict__ out) {
  __shared__ VolumeGrid g;
  __shared__ __align__(16) float s_corner[8][32 * SBW_STRIDE];
  if (threadIdx.x < 3) {
    g.lo[threadIdx.x] = bounds[threadIdx.x];
    g.ext[threadIdx.x] = __fsub_rn(bounds[3 + threadIdx.x], bounds[threadIdx.x]);
    g.dim[threadIdx.x] = threadIdx.x == 0 ? X : (threadIdx.x == 1 ? Y : Z);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  float *sc = s_corner[wib];
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
    int64_t mine = base + lane;
    float w[8];
    int off[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      w[k] = 0.f;
      off[k] = -1;
    }
    if (mine < n) trilinear_corners(g, pts[3 * mine], pts[3 * mine + 1], pts[3 * mine + 2], w, off);
    __syncwarp();                                   // the previous iteration's readers are done
    float4 *dst = reinterpret_cast<float4 *>(sc + lane * SBW_STRIDE);
    dst[0] = make_float4(w[0], w[1], w[2], w[3]);
    dst[1] = make_float4(w[4], w[5], w[6], w[7]);
    dst[2] = make_float4(__int_as_float(off[0]), __int_as_float(off[1]), __int_as_float(off[2]), __int_as_float(off[3]));
    dst[3] = make_float4(__int_as_float(off[4]), __int_as_float(off[5]), __int_as_float(off[6]), __int_as_float(off[7]));
    __syncwarp();
    const int cnt = (int)min((int64_t)32, n - base);
    const bool ch = lane < ANINERF_BW_CH;
    for (int p = 0; p < cnt; p += 2) {
      // two points per iteration: their 16 corner rows are independent loads
      float wv[2][8];
      int ov[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4 *src = reinterpret_cast<const float4 *>(sc + min(p + u, cnt - 1) * SBW_STRIDE);
        const float4 a = src[0], b = src[1], c = src[2], d = src[3];
        wv[u][0] = a.x; wv[u][1] = a.y; wv[u][2] = a.z; wv[u][3] = a.w;
        wv[u][4] = b.x; wv[u][5] = b.y; wv[u][6] = b.z; wv[u][7] = b.w;
        ov[u][0] = __float_as_int(c.x); ov[u][1] = __float_as_int(c.y); ov[u][2] = __float_as_int(c.z); ov[u][3] = __float_as_int(c.w);
        ov[u][4] = __float_as_int(d.x); ov[u][5] = __float_as_int(d.y); ov[u][6] = __float_as_int(d.z); ov[u][7] = __float_as_int(d.w);
      }
      float v[2][8];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int k = 0; k < 8; ++k) v[u][k] = (ch && ov[u][k] >= 0) ? __ldg(vol + (int64_t)ov[u][k] * ANINERF_BW_CH + lane) : 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (p + u < cnt) {
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (ov[u][k] >= 0) acc = __fadd_rn(acc, __fmul_rn(v[u][k], wv[u][k]));      // (an outside corner is skipped, as ATen does)
          if (ch) out[(base + p + u) * ANINERF_BW_CH + lane] = acc;
        }
      }
    }
  }
}

// =============================================================================================
// 7. forward / inverse LBS (blend_utils.py:41-59, 77-90): bone matrices staged in shared memory
// =============================================================================================
template <bool INVERSE>
__global__ void __launch_bounds__(256) lbs_kernel(const float *__restrict__ pts, const float *__restrict__ bw, int64_t n,
                                                  const float *__restrict__ A, float *__restrict__ out) {
  __shared__ float sA[ANINERF_N_BONES][12];
  for (int k = threadIdx.x; k < ANINERF_N_BONES * 12; k += blockDim.x) sA[k / 12][k % 12] = A[(k / 12) * 16 + (k % 12)];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) M[j] = 0.f;
  const float4 *row = reinterpret_cast<const float4 *>(bw + i * ANINERF_N_BONES);   // 96-byte rows: 16 B aligned
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    float4 w4 = __ldg(row + q);
    float ww[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int j = 0; j < 12; ++j) M[j] = fmaf(ww[e], sA[q * 4 + e][j], M[j]);
  }
  float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  if (INVERSE) {
    float qx = x - M[3], qy = y - M[7], qz = z - M[11];
    float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], k = M[10];
    float c00 = e * k - f * h, c01 = c * h - b * k, c02 = b * f - c * e;
    float c10 = f * g - d * k, c11 = a * k - c * g, c12 = c * d - a * f;
    float c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
    float det = a * c00 + b * c10 + c * c20;
    float inv = 1.0f / det;
    out[3 * i] = (c00 * qx + c01 * qy + c02 * qz) * inv;
    out[3 * i + 1] = (c10 * qx + c11 * qy + c12 * qz) * inv;
    out[3 * i + 2] = (c20 * qx + c21 * qy + c22 * qz) * inv;
  } else {
    out[3 * i] = M[0] * x + M[1] * y + M[2] * z + M[3];
    out[3 * i + 1] = M[4] * x + M[5] * y + M[6] * z + M[7];
    out[3 * i + 2] = M[8] * x + M[9] * y + M[10] * z + M[11];
  }
}

// =============================================================================================
// 8. alpha compositing (nerf_net_utils.py:6-36): one warp per ray, S/32 samples per lane,
//    exclusive product scan of (1 - alpha + 1e-10) with shuffles.
// =============================================================================================
// Fused compositing + image gather over NVLink peer memory: when peers.world > 0 the ray's (rgb, acc, depth) row is stored
// straight into EVERY rank's frame-ordered image buffer (peer-mapped symmetric memory), 20 B per ray and peer, instead of
// an all_gather + reorder afterwards.  Local ray r belongs to local chunk r / chunk_rays = global chunk i*world + rank.
struct PeerScatter {
  float *maps[ANINERF_MAX_PEERS];
  int world, rank, chunk_rays;
};

template <int SPL>   // samples per lane (S = 32*SPL), lane owns samples [lane*SPL, lane*SPL+SPL)
__global__ void __launch_bounds__(256) composite_kernel(const float4 *__restrict__ raw, const float *__restrict__ z_vals,
                                                        const float *__restrict__ near, const float *__restrict__ far,
                                                        const float *__restrict__ t_vals, int64_t n_rays, int white_bkgd,
                                                        float *__restrict__ rgb_map, float *__restrict__ acc_map,
                                                        float *__restrict__ depth_map, float *__restrict__ disp_map,
                                                        float *__restrict__ weights, PeerScatter peers,
                                                        const uint32_t *__restrict__ mask_words, const int32_t *__restrict__ block_offsets) {
  // mask_words != nullptr: `raw` holds only the ACTIVE samples, compacted in ascending sample order (the render-only path never
  // materialises the dense (n,4) buffer); a ray's rows start at block_offsets[its 2048-sample block] + the active samples of
  // the block before it, and an inactive sample contributes exactly what a zero row of the dense buffer would.
  constexpr int S = 32 * SPL;
  int lane = threadIdx.x & 31;
  int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const bool valid = ray < n_rays;
  if (!valid) {
    if (peers.world == 0) return;
    ray = n_rays - 1;            // the peer scatter below needs the whole block at its __syncthreads: compute a dummy, write nothing
  }
  float r = 0.f, g = 0.f, b = 0.f, acc = 0.f, dep = 0.f;
  bool empty = false;            // compact mode: a ray without a single active sample composites to exact zeros (w = 0 * T)
  if (mask_words) {
    uint32_t any = 0u;
#pragma unroll
    for (int j = 0; j < SPL; ++j) any |= __ldg(mask_words + ray * SPL + j);
    empty = any == 0u;
  }
  if (!empty) {
  float4 c[SPL];
  float z[SPL];
  if (mask_words) {
    constexpr int WPB = 2048 / 32;                          // mask words per 2048-sample block (MB below)
    const int64_t w0 = ray * SPL;                           // first mask word of the ray (S/32 = SPL words per ray)
    const int64_t blk = (ray * S) / 2048;
    const int64_t bw0 = blk * WPB;
    int before = 0;
    for (int64_t i = bw0 + lane; i < w0; i += 32) before += __popc(__ldg(mask_words + i));
#pragma unroll
    for (int o = 16; o; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    int64_t pos = (int64_t)__ldg(block_offsets + blk) + before;
    uint32_t mw[SPL];
#pragma unroll
    for (int j = 0; j < SPL; ++j) mw[j] = __ldg(mask_words + w0 + j);
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      const int sidx = lane * SPL + j;                      // sample index within the ray
      const int word = sidx >> 5, bit = sidx & 31;
      int64_t p = pos;
      uint32_t m = 0u;
#pragma unroll
      for (int q = 0; q < SPL; ++q) {      // constant indices only (no local-memory copy of mw)
        if (q < word) p += __popc(mw[q]);
        if (q == word) m = mw[q];
      }
      p += __popc(m & ((1u << bit) - 1u));
      c[j] = ((m >> bit) & 1u) ? __ldg(raw + p) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
#pragma unroll
    for (int j = 0; j < SPL; ++j) c[j] = __ldg(raw + ray * S + lane * SPL + j);
  }
  if (z_vals) {
#pragma unroll
    for (int j = 0; j < SPL; ++j) z[j] = __ldg(z_vals + ray * S + lane * SPL + j);
  } else {
    float nr = near[ray], fr = far[ray];
#pragma unroll
    for (int j = 0; j < SPL; ++j) {
      float t = t_vals[lane * SPL + j];
      z[j] = z_lerp(nr, fr, t, __fsub_rn(1.0f, t));
    }
  }
  // local transmittance products
  float f[SPL];
  float prod = 1.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    f[j] = __fadd_rn(__fsub_rn(1.0f, c[j].w), 1e-10f);
    prod *= f[j];
  }
  float inc = prod;   // inclusive scan over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc *= t;
  }
  float T = __shfl_up_sync(0xffffffffu, inc, 1);
  if (lane == 0) T = 1.f;
#pragma unroll
  for (int j = 0; j < SPL; ++j) {
    float w = c[j].w * T;
    if (weights && valid) weights[ray * S + lane * SPL + j] = w;
    r = fmaf(w, c[j].x, r);
    g = fmaf(w, c[j].y, g);
    b = fmaf(w, c[j].z, b);
    acc += w;
    dep = fmaf(w, z[j], dep);
    T *= f[j];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    r += __shfl_xor_sync(0xffffffffu, r, o);
    g += __shfl_xor_sync(0xffffffffu, g, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    dep += __shfl_xor_sync(0xffffffffu, dep, o);
  }
  }   // !empty
  __shared__ __align__(16) float s_rows[8 * 5];
  if (lane == 0) {
    if (white_bkgd) {
      r += 1.f - acc;
      g += 1.f - acc;
      b += 1.f - acc;
    }
    if (valid) {
      if (rgb_map) {
        rgb_map[3 * ray] = r;
        rgb_map[3 * ray + 1] = g;
        rgb_map[3 * ray + 2] = b;
      }
      if (acc_map) acc_map[ray] = acc;
      if (depth_map) depth_map[ray] = dep;
      if (disp_map) disp_map[ray] = 1.0f / fmaxf(1e-10f, dep / acc);
    }
    if (peers.world > 0) {
      float *row = s_rows + (threadIdx.x >> 5) * 5;
      row[0] = r;
      row[1] = g;
      row[2] = b;
      row[3] = acc;
      row[4] = dep;
    }
  }
  if (peers.world > 0) {
    // the block's 8 rays are 8 consecutive rows of one chunk: 160 contiguous, 16-byte aligned bytes per peer -> ten float4
    // stores per peer (NVLink moves them as full packets; per-float remote stores cost ~5x the time)
    __syncthreads();
    const int64_t ray0 = (int64_t)blockIdx.x * 8;
    const int nvalid = (int)min((int64_t)8, n_rays - ray0);
    const int64_t lc = ray0 / peers.chunk_rays;
    const int64_t row0 = (lc * peers.world + peers.rank) * peers.chunk_rays + (ray0 - lc * peers.chunk_rays);
    // (constant indices only: a dynamically indexed kernel-parameter array is copied to local memory by every thread)
    if (nvalid == 8) {
      if (threadIdx.x < 10 * peers.world) {
        const int kk = threadIdx.x / 10, q = threadIdx.x % 10;
        float *base = nullptr;
#pragma unroll
        for (int k = 0; k < ANINERF_MAX_PEERS; ++k)
          if (k == kk) base = peers.maps[k];
        reinterpret_cast<float4 *>(base + row0 * 5)[q] = reinterpret_cast<const float4 *>(s_rows)[q];
      }
    } else {
      for (int i = threadIdx.x; i < peers.world * nvalid * 5; i += blockDim.x) {
        const int kk = i / (nvalid * 5), j = i % (nvalid * 5);
        float *base = nullptr;
#pragma unroll
        for (int k = 0; k < ANINERF_MAX_PEERS; ++k)
          if (k == kk) base = peers.maps[k];
        base[row0 * 5 + j] = s_rows[j];
      }
    }
  }
}

// =============================================================================================
// 9. fused render front end: mask pass over all samples (sample -> pose -> trilinear pnorm ->
//    threshold), per-2048-ray-chunk argmin forcing, stable compaction of the active samples
//    (tpose_nerf_network.py:143-157 over tpose_renderer.py:14-69)
// =============================================================================================
constexpr int MB = 2048;   // samples per mask block (8 per thread, 256 threads)

// ---- multi-view silhouette culling (prepare_inside_pts, tpose_renderer_mmsk.py:14-57) ----------
// A world point survives when it projects into the (dilated) body mask of EVERY training view:
//   cam = pts @ R^T + T ; img = cam @ K^T ; uv = round(img.xy / img.z) clamped to the image ; msk[v,u] != 0
// torch.matmul((1,m,3),(1,3,3)^T) on the CPU is bit-equal to the FMA chain below (checked in
// tests/test_oracle_golden.py); round() is round-half-to-even (rintf); the rest is integer work.
constexpr int SIL_MAX_STAGED = 32;   // views whose K / RT are staged in shared memory
struct SilDev {
  const uint8_t *msks;   // (V,H,W)
  const float *Ks;       // (V,3,3)
  const float *RT;       // (V,4,4)
  int V, H, W;
};
struct SilShared {
  float K[SIL_MAX_STAGED][9];
  float RT[SIL_MAX_STAGED][12];
};

__device__ __forceinline__ void stage_silhouettes(const SilDev &s, SilShared *sh) {
  if (!s.msks) return;
  const int V = min(s.V, SIL_MAX_STAGED);
  for (int i = threadIdx.x; i < V * 9; i += blockDim.x) sh->K[i / 9][i % 9] = s.Ks[i];
  for (int i = threadIdx.x; i < V * 12; i += blockDim.x) sh->RT[i / 12][i % 12] = s.RT[(i / 12) * 16 + i % 12];
}

__device__ __forceinline__ bool inside_view(const float *K, const float *RT, const uint8_t *msk, int H, int W, float x, float y, float z) {
  float c[3], q[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
    c[i] = __fadd_rn(__fmaf_rn(z, RT[4 * i + 2], __fmaf_rn(y, RT[4 * i + 1], __fmul_rn(x, RT[4 * i]))), RT[4 * i + 3]);
#pragma unroll
  for (int i = 0; i < 3; ++i) q[i] = __fmaf_rn(c[2], K[3 * i + 2], __fmaf_rn(c[1], K[3 * i + 1], __fmul_rn(c[0], K[3 * i])));
  long long u = __float2ll_rn(__fdiv_rn(q[0], q[2]));   // .round().long(): half-to-even, then the integer cast
  long long v = __float2ll_rn(__fdiv_rn(q[1], q[2]));
  u = u < 0 ? 0 : (u > W - 1 ? W - 1 : u);
  v = v < 0 ? 0 : (v > H - 1 ? H - 1 : v);
  return __ldg(msk + v * W + u) != 0;
}

__device__ __forceinline__ bool inside_all_views(const SilDev &s, const SilShared *sh, float x, float y, float z) {
  const int64_t plane = (int64_t)s.H * s.W;
  for (int v = 0; v < s.V; ++v) {
    bool in = v < SIL_MAX_STAGED ? inside_view(sh->K[v], sh->RT[v], s.msks + v * plane, s.H, s.W, x, y, z)
                                 : false;
    if (v >= SIL_MAX_STAGED) {   // beyond the staged set: read the matrices from global memory
      float K[9], RT[12];
      for (int i = 0; i < 9; ++i) K[i] = __ldg(s.Ks + v * 9 + i);
      for (int i = 0; i < 12; ++i) RT[i] = __ldg(s.RT + v * 16 + i);
      in = inside_view(K, RT, s.msks + v * plane, s.H, s.W, x, y, z);
    }
    if (!in) return false;
  }
  return true;
}

__global__ void __launch_bounds__(256) inside_kernel(const float *__restrict__ wpts, int64_t n, SilDev s, uint8_t *__restrict__ inside) {
  __shared__ SilShared sh;
  stage_silhouettes(s, &sh);
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  inside[i] = inside_all_views(s, &sh, wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2]) ? 1 : 0;
}

struct SampleSetup {
  const float *ray_o, *ray_d, *near, *far, *t_vals, *t_rand;
  int64_t n_rays;
  int S;
  const float *R, *Th, *bounds;   // device: (3,3), (3,), (2,3) -- no host round trip per frame
  int dim[3];
  float norm_th;
  SilDev sil;                     // sil.msks == nullptr: no silhouette culling
};

// stage the per-frame rigid transform and volume grid in shared memory (call before __syncthreads)
__device__ __forceinline__ void stage_frame(const float *R, const float *Th, const float *bounds, const int dim[3], RigidFrame *f,
                                            VolumeGrid *g) {
  if (threadIdx.x < 9) f->R[threadIdx.x] = R[threadIdx.x];
  if (threadIdx.x < 3) {
    f->Th[threadIdx.x] = Th[threadIdx.x];
    g->lo[threadIdx.x] = bounds[threadIdx.x];
    g->ext[threadIdx.x] = __fsub_rn(bounds[3 + threadIdx.x], bounds[threadIdx.x]);   // blend_utils.py:133
    g->dim[threadIdx.x] = dim[threadIdx.x];
  }
}

__device__ __forceinline__ void sample_at(const SampleSetup &p, const float *st, const float *s1mt, int64_t r, int s, float &z,
                                          float &wx, float &wy, float &wz) {
  float nr = __ldg(p.near + r), fr = __ldg(p.far + r);
  z = z_lerp(nr, fr, st[s], s1mt[s]);
  if (p.t_rand) {
    float zp = s > 0 ? z_lerp(nr, fr, st[s - 1], s1mt[s - 1]) : z;
    float zn = s < p.S - 1 ? z_lerp(nr, fr, st[s + 1], s1mt[s + 1]) : z;
    z = z_jitter(zp, z, zn, s == 0, s == p.S - 1, __ldg(p.t_rand + r * p.S + s));
  }
  wx = __fadd_rn(__ldg(p.ray_o + 3 * r), __fmul_rn(__ldg(p.ray_d + 3 * r), z));
  wy = __fadd_rn(__ldg(p.ray_o + 3 * r + 1), __fmul_rn(__ldg(p.ray_d + 3 * r + 1), z));
  wz = __fadd_rn(__ldg(p.ray_o + 3 * r + 2), __fmul_rn(__ldg(p.ray_d + 3 * r + 2), z));
}

// order-preserving float -> uint32 key
__device__ __forceinline__ uint32_t float_key(float v) {
  uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// dist: the (X,Y,Z) plane of channel 24 (distance to the SMPL surface)
__global__ void __launch_bounds__(256) mask_kernel(SampleSetup p, const float *__restrict__ dist, int64_t chunk_samples,
                                                   uint32_t *__restrict__ mask_words, int32_t *__restrict__ block_counts,
                                                   unsigned long long *__restrict__ chunk_argmin) {
  __shared__ float st[ANINERF_MAX_SAMPLES], s1mt[ANINERF_MAX_SAMPLES];
  __shared__ int s_cnt[8];
  __shared__ unsigned long long s_min[8];
  __shared__ RigidFrame s_frame;
  __shared__ VolumeGrid s_grid;
  __shared__ SilShared s_sil;
  stage_frame(p.R, p.Th, p.bounds, p.dim, &s_frame, &s_grid);
  stage_silhouettes(p.sil, &s_sil);
  if (threadIdx.x < p.S) {
    float t = p.t_vals[threadIdx.x];
    st[threadIdx.x] = t;
    s1mt[threadIdx.x] = __fsub_rn(1.0f, t);
  }
  __syncthreads();
  const uint32_t n = (uint32_t)(p.n_rays * p.S);            // < 2^31 (checked by the host)
  const uint32_t base = blockIdx.x * (uint32_t)MB;
  const int log2S = p.S == 64 ? 6 : 5;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = 0;
  unsigned long long best = ~0ull;
  const uint32_t chunk_base = base % (uint32_t)chunk_samples;   // blocks never straddle chunks
#pragma unroll 2
  for (int it = 0; it < MB / 256; ++it) {
    const uint32_t i = base + it * 256 + threadIdx.x;
    bool act = false;
    if (i < n) {
      const int64_t r = i >> log2S;
      const int s = (int)(i & (uint32_t)(p.S - 1));
      float z, wx, wy, wz, px, py, pz;
      sample_at(p, st, s1mt, r, s, z, wx, wy, wz);
      // tpose_renderer_mmsk.py:71-90: only samples inside every training-view silhouette reach the network
      // (and are candidates of its per-chunk argmin forcing)
      if (!p.sil.msks || inside_all_views(p.sil, &s_sil, wx, wy, wz)) {
        world_to_pose(s_frame, wx, wy, wz, px, py, pz);
        float w[8];
        int off[8];
        trilinear_corners_clamped(s_grid, px, py, pz, w, off);
        float pn = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) pn = __fadd_rn(pn, __fmul_rn(__ldg(dist + off[k]), w[k]));
        act = pn < p.norm_th;
        unsigned long long key = ((unsigned long long)float_key(pn) << 32) | (unsigned long long)(chunk_base + it * 256 + threadIdx.x);
        best = key < best ? key : best;
      }
    }
    unsigned bal = __ballot_sync(0xffffffffu, act);
    if (lane == 0) {
      mask_words[(base + it * 256) / 32 + warp] = bal;
      cnt += __popc(bal);
    }
  }
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if (lane == 0) {
    s_cnt[warp] = cnt;
    s_min[warp] = best;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    unsigned long long b = ~0ull;
    for (int w = 0; w < 8; ++w) {
      t += s_cnt[w];
      b = s_min[w] < b ? s_min[w] : b;
    }
    block_counts[blockIdx.x] = t;
    atomicMin(chunk_argmin + base / chunk_samples, b);
  }
}

// ONE block after the mask pass: (1) a chunk without active sample gets its argmin(pnorm) sample forced on
// (tpose_nerf_network.py:154), (2) exclusive scan of the per-block counts -> compacted offsets and n_active, (3) per-chunk start
// rows.  (Three single-block launches before; their fixed ~20 us is a visible share of a ray-tiled frame at 8 GPUs.)
__global__ void __launch_bounds__(1024) finalize_mask_kernel(int64_t n_chunks, int64_t blocks_per_chunk, int64_t n_blocks, int64_t chunk_samples,
                                                             const unsigned long long *__restrict__ chunk_argmin, uint32_t *__restrict__ mask_words,
                                                             int32_t *__restrict__ block_counts, int32_t *__restrict__ block_offsets,
                                                             int32_t *__restrict__ total, int32_t *__restrict__ chunk_offsets) {
  __shared__ int warp_sums[32];
  __shared__ int carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // (1) one warp per chunk
  for (int64_t c = warp; c < n_chunks; c += 32) {
    const int64_t b0 = c * blocks_per_chunk, b1 = min(n_blocks, b0 + blocks_per_chunk);
    int t = 0;
    for (int64_t b = b0 + lane; b < b1; b += 32) t += block_counts[b];
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0 && t == 0 && chunk_argmin[c] != ~0ull) {   // ~0: silhouette culling left the chunk empty, the network is not called
      const int64_t i = c * chunk_samples + (int64_t)(chunk_argmin[c] & 0xffffffffull);
      mask_words[i / 32] |= 1u << (i % 32);
      block_counts[i / MB] += 1;
    }
  }
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  // (2) exclusive scan of the block counts
  for (int64_t base = 0; base < n_blocks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < n_blocks ? block_counts[i] : 0;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      const int ws = warp_sums[threadIdx.x];
      int winc = ws;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (threadIdx.x >= o) winc += t;
      }
      warp_sums[threadIdx.x] = winc - ws;
    }
    __syncthreads();
    const int excl = carry + warp_sums[warp] + inc - v;
    if (i < n_blocks) block_offsets[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
  // (3) chunk c starts at the offset of its first block
  if (chunk_offsets)
    for (int64_t c = threadIdx.x; c <= n_chunks; c += 1024) {
      const int64_t b = c * blocks_per_chunk;
      chunk_offsets[c] = b < n_blocks ? block_offsets[b] : carry;
    }
}

// scatter the active samples in ascending index order: index, pose point, view direction, dist
__global__ void __launch_bounds__(256) compact_samples_kernel(SampleSetup p, const uint32_t *__restrict__ mask_words,
                                                              const int32_t *__restrict__ block_offsets, int32_t *__restrict__ index,
                                                              float *__restrict__ ppts, float *__restrict__ viewdir,
                                                              float *__restrict__ dists) {
  __shared__ float st[ANINERF_MAX_SAMPLES], s1mt[ANINERF_MAX_SAMPLES];
  __shared__ int word_base[MB / 32];
  __shared__ RigidFrame s_frame;
  __shared__ VolumeGrid s_grid;
  stage_frame(p.R, p.Th, p.bounds, p.dim, &s_frame, &s_grid);
  if (threadIdx.x < p.S) {
    float t = p.t_vals[threadIdx.x];
    st[threadIdx.x] = t;
    s1mt[threadIdx.x] = __fsub_rn(1.0f, t);
  }
  int64_t n = p.n_rays * p.S;
  int64_t base = (int64_t)blockIdx.x * MB;
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ int first_half_total;
  if (threadIdx.x < MB / 32) {
    // exclusive prefix over the block's 64 mask words: two warps, 32 words each
    int64_t wi = base / 32 + threadIdx.x;
    int c = wi * 32 < n ? __popc(mask_words[wi]) : 0;
    int inc = c;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    word_base[threadIdx.x] = inc - c;
    if (threadIdx.x == 31) first_half_total = inc;
  }
  __syncthreads();
  int blk = block_offsets[blockIdx.x];
  for (int it = 0; it < MB / 256; ++it) {
    int64_t i = base + it * 256 + threadIdx.x;
    if (i >= n) continue;
    int word = it * 8 + warp;
    uint32_t bits = mask_words[base / 32 + word];
    if (!((bits >> lane) & 1u)) continue;
    int dst = blk + word_base[word] + (word >= 32 ? first_half_total : 0) + __popc(bits & ((1u << lane) - 1u));
    int64_t r = i / p.S;
    int s = (int)(i - r * p.S);
    float z, wx, wy, wz, px, py, pz;
    sample_at(p, st, s1mt, r, s, z, wx, wy, wz);
    world_to_pose(s_frame, wx, wy, wz, px, py, pz);
    index[dst] = (int32_t)i;
    ppts[3 * dst] = px;
    ppts[3 * dst + 1] = py;
    ppts[3 * dst + 2] = pz;
    viewdir[3 * dst] = __ldg(p.ray_d + 3 * r);
    viewdir[3 * dst + 1] = __ldg(p.ray_d + 3 * r + 1);
    viewdir[3 * dst + 2] = __ldg(p.ray_d + 3 * r + 2);
    // dists = diff(z), last interval duplicated (tpose_renderer.py:63-65)
    float zo, ax, ay, az;
    float d;
    if (s < p.S - 1) {
      sample_at(p, st, s1mt, r, s + 1, zo, ax, ay, az);
      d = __fsub_rn(zo, z);
    } else {
      sample_at(p, st, s1mt, r, s - 1, zo, ax, ay, az);
      d = __fsub_rn(z, zo);
    }
    dists[dst] = d;
  }
}

// split the reference (X,Y,Z,25) volume into a 24-channel weight plane (96-byte rows, float4
// aligned) and the distance plane the mask pass gathers from
__global__ void __launch_bounds__(256) split_volume_kernel(const float *__restrict__ vol, int64_t voxels, float *__restrict__ w24,
                                                           float *__restrict__ dist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= voxels * ANINERF_BW_CH) return;
  int64_t v = i / ANINERF_BW_CH;
  int c = (int)(i - v * ANINERF_BW_CH);
  float x = vol[i];
  if (c < ANINERF_N_BONES) w24[v * ANINERF_N_BONES + c] = x;
  else dist[v] = x;
}

// gather world points -> pose points + pnorm mask for the density query (calculate_alpha)
struct PointSetup {
  const float *R, *Th, *bounds;
  int dim[3];
};

__global__ void __launch_bounds__(256) mask_points_kernel(const float *__restrict__ wpts, int64_t n, PointSetup ps,
                                                          float norm_th, const float *__restrict__ dist, int64_t chunk_pts,
                                                          uint8_t *__restrict__ mask, unsigned long long *__restrict__ chunk_argmin,
                                                          float *__restrict__ ppts) {
  __shared__ RigidFrame frame;
  __shared__ VolumeGrid grid;
  stage_frame(ps.R, ps.Th, ps.bounds, ps.dim, &frame, &grid);
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long best = ~0ull;
  if (i < n) {
    float px, py, pz;
    world_to_pose(frame, wpts[3 * i], wpts[3 * i + 1], wpts[3 * i + 2], px, py, pz);
    ppts[3 * i] = px;
    ppts[3 * i + 1] = py;
    ppts[3 * i + 2] = pz;
    float w[8];
    int off[8];
    trilinear_corners_clamped(grid, px, py, pz, w, off);
    float pn = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) pn = __fadd_rn(pn, __fmul_rn(__ldg(dist + off[k]), w[k]));
    mask[i] = pn < norm_th ? 1 : 0;
    best = ((unsigned long long)float_key(pn) << 32) | (unsigned long long)(uint32_t)(i % chunk_pts);
  }
  // blocks never straddle chunks when chunk_pts % 256 == 0 (checked by the host)
  for (int o = 16; o; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
    best = t < best ? t : best;
  }
  if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(chunk_argmin + ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / chunk_pts, best);
}

__global__ void force_argmin_points_kernel(int64_t n_chunks, int64_t chunk_pts, int64_t n, const unsigned long long *__restrict__ chunk_argmin,
                                           uint8_t *__restrict__ mask) {
  // one warp per chunk: OR-reduce the chunk's mask, force the argmin on if empty
  int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (c >= n_chunks) return;
  int64_t b0 = c * chunk_pts, b1 = min(n, b0 + chunk_pts);
  int any = 0;
  for (int64_t i = b0 + lane; i < b1 && !any; i += 32) any |= mask[i];
  any = __any_sync(0xffffffffu, any);
  if (!any && lane == 0) mask[b0 + (int64_t)(chunk_argmin[c] & 0xffffffffull)] = 1;
}

__global__ void __launch_bounds__(256) gather_points_kernel(const float *__restrict__ src, const int32_t *__restrict__ index,
                                                            const int32_t *__restrict__ count, float *__restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *count) return;
  int64_t j = index[i];
  dst[3 * i] = src[3 * j];
  dst[3 * i + 1] = src[3 * j + 1];
  dst[3 * i + 2] = src[3 * j + 2];
}

__global__ void __launch_bounds__(256) scatter_scalar_kernel(const float *__restrict__ src, const int32_t *__restrict__ index,
                                                             const int32_t *__restrict__ count, float *__restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *count) return;
  dst[index[i]] = src[i];
}

}  // namespace aninerf

using namespace aninerf;

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int aninerf_version(void) { return ANINERF_ABI_VERSION; }
const char *aninerf_last_error(void) { return g_err; }
int64_t aninerf_launch_count(void) { return (int64_t)g_launches.load(); }

int aninerf_gen_rays(const aninerf_camera *cam, float *ray_o, float *ray_d, void *stream) {
  ANI_CHECK_ARG(cam && ray_o && ray_d && cam->H > 0 && cam->W > 0);
  CamDev c;
  memcpy(c.Kinv, cam->Kinv, sizeof(c.Kinv));
  memcpy(c.R, cam->R, sizeof(c.R));
  memcpy(c.T, cam->T, sizeof(c.T));
  c.H = cam->H;
  c.W = cam->W;
  for (int k = 0; k < 3; ++k)   // rays_o = -(R^T T)  (:78)
    c.o[k] = -((cam->R[k] * cam->T[0] + cam->R[3 + k] * cam->T[1]) + cam->R[6 + k] * cam->T[2]);
  int64_t n = (int64_t)c.H * c.W;
  gen_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c, ray_o, ray_d);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_near_far(const float *bounds_host, const float *ray_o, const float *ray_d, int64_t n, float *near, float *far,
                     uint8_t *mask, void *stream) {
  ANI_CHECK_ARG(bounds_host && ray_o && ray_d && near && far && mask && n >= 0);
  if (n == 0) return ANINERF_OK;
  BoxDev B;
  for (int k = 0; k < 3; ++k) {
    B.b[k] = (double)bounds_host[k] + (-0.01);
    B.b[3 + k] = (double)bounds_host[3 + k] + 0.01;
  }
  near_far_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, ray_o, ray_d, n, near, far, mask);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int64_t aninerf_compact_workspace_bytes(int64_t n) {
  int64_t blocks = (n + CB - 1) / CB + 1;
  return align_up(blocks * 4, 256) * 2 + 256;
}

int aninerf_compact_rays(const float *ray_o, const float *ray_d, const float *near, const float *far, const uint8_t *mask, int64_t n,
                         float *ray_o_out, float *ray_d_out, float *near_out, float *far_out, int32_t *index_out, int32_t *count,
                         void *workspace, int64_t workspace_bytes, void *stream) {
  ANI_CHECK_ARG(mask && count && workspace && n >= 0);
  if (workspace_bytes < aninerf_compact_workspace_bytes(n)) return fail(ANINERF_ENOMEM, "%s: workspace too small%s", __func__);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ANI_CUDA(cudaMemsetAsync(count, 0, 4, st));
    return ANINERF_OK;
  }
  int64_t blocks = (n + CB - 1) / CB;
  int32_t *counts = (int32_t *)workspace;
  int32_t *offsets = (int32_t *)((char *)workspace + align_up((blocks + 1) * 4, 256));
  count_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(mask, n, counts);
  ANI_LAUNCHED();
  scan_counts_kernel<<<1, 1024, 0, st>>>(counts, blocks, offsets, count);
  ANI_LAUNCHED();
  scatter_rays_kernel<<<(unsigned)blocks, 256, 0, st>>>(ray_o, ray_d, near, far, mask, n, offsets, ray_o_out, ray_d_out, near_out,
                                                         far_out, index_out);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_sample_points(const float *ray_o, const float *ray_d, const float *near, const float *far, const float *t_vals,
                          const float *t_rand, int64_t n_rays, int32_t S, float *pts, float *z_vals, float *dists, void *stream) {
  ANI_CHECK_ARG(ray_o && ray_d && near && far && t_vals && n_rays >= 0 && S >= 2 && S <= ANINERF_MAX_SAMPLES);
  if (n_rays == 0) return ANINERF_OK;
  int64_t n = n_rays * S;
  sample_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ray_o, ray_d, near, far, t_vals, t_rand, n_rays, S,
                                                                                      pts, z_vals, dists);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_world_to_pose(const float *wpts, int64_t n, const float *R, const float *Th, float *ppts, void *stream) {
  ANI_CHECK_ARG(wpts && R && Th && ppts && n >= 0);
  if (n == 0) return ANINERF_OK;
  world_to_pose_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wpts, n, R, Th, ppts);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_sample_blend_weights(const float *pts, int64_t n, const float *vol, const int32_t dims[3], const float *bounds, float *out,
                                 void *stream) {
  ANI_CHECK_ARG(pts && vol && dims && bounds && out && n >= 0 && dims[0] > 0 && dims[1] > 0 && dims[2] > 0);
  if (n == 0) return ANINERF_OK;
  int64_t warps = (n + 31) / 32;
  int64_t blocks = (warps + 7) / 8;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  sample_bw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pts, n, vol, bounds, dims[0], dims[1], dims[2], out);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_inside_all_views(const float *wpts, int64_t n, const aninerf_silhouettes *sil, uint8_t *inside, void *stream) {
  ANI_CHECK_ARG(wpts && sil && inside && n >= 0);
  ANI_CHECK_ARG(sil->msks && sil->Ks && sil->RT && sil->n_views > 0 && sil->H > 0 && sil->W > 0);
  if (n == 0) return ANINERF_OK;
  SilDev s{sil->msks, sil->Ks, sil->RT, sil->n_views, sil->H, sil->W};
  inside_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(wpts, n, s, inside);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_inverse_lbs(const float *ppts, const float *bw, int64_t n, const float *A, float *tpts, void *stream) {
  ANI_CHECK_ARG(ppts && bw && A && tpts && n >= 0);
  if (n == 0) return ANINERF_OK;
  lbs_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ppts, bw, n, A, tpts);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_forward_lbs(const float *tpts, const float *bw, int64_t n, const float *A, float *ppts, void *stream) {
  ANI_CHECK_ARG(tpts && bw && A && ppts && n >= 0);
  if (n == 0) return ANINERF_OK;
  lbs_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tpts, bw, n, A, ppts);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_composite(const float *raw, const float *z_vals, int64_t n_rays, int32_t S, int32_t white_bkgd, float *rgb_map,
                      float *acc_map, float *depth_map, float *disp_map, float *weights, void *stream) {
  ANI_CHECK_ARG(raw && z_vals && n_rays >= 0 && (S == 32 || S == 64));
  if (n_rays == 0) return ANINERF_OK;
  unsigned blocks = (unsigned)((n_rays + 7) / 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 64)
    composite_kernel<2><<<blocks, 256, 0, st>>>((const float4 *)raw, z_vals, nullptr, nullptr, nullptr, n_rays, white_bkgd, rgb_map,
                                                acc_map, depth_map, disp_map, weights, PeerScatter{}, nullptr, nullptr);
  else
    composite_kernel<1><<<blocks, 256, 0, st>>>((const float4 *)raw, z_vals, nullptr, nullptr, nullptr, n_rays, white_bkgd, rgb_map,
                                                acc_map, depth_map, disp_map, weights, PeerScatter{}, nullptr, nullptr);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// internal entry points used by render.cu (same translation-unit family, C++ linkage)
// ---------------------------------------------------------------------------------------------
namespace aninerf {

int launch_split_volume(const float *vol, const int32_t dims[3], float *w24, float *dist, cudaStream_t st) {
  int64_t voxels = (int64_t)dims[0] * dims[1] * dims[2];
  int64_t n = voxels * ANINERF_BW_CH;
  split_volume_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(vol, voxels, w24, dist);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

struct FrontEndBuffers {
  uint32_t *mask_words;
  int32_t *block_counts, *block_offsets;
  unsigned long long *chunk_argmin;
};

// mask + force + scan + compact.  Returns the compacted arrays and *n_active on the device.
int launch_front_end(const float *ray_o, const float *ray_d, const float *near, const float *far, const float *t_vals,
                     const float *t_rand, int64_t n_rays, int S, int chunk_rays, const float *R, const float *Th,
                     const float *bounds, const int32_t dims[3], const float *dist_plane, float norm_th, FrontEndBuffers fb,
                     int32_t *index, float *ppts, float *viewdir, float *dists, int32_t *n_active, int32_t *chunk_offsets,
                     const aninerf_silhouettes *sil, cudaStream_t st) {
  SampleSetup p;
  p.sil = SilDev{nullptr, nullptr, nullptr, 0, 0, 0};
  if (sil) p.sil = SilDev{sil->msks, sil->Ks, sil->RT, sil->n_views, sil->H, sil->W};
  p.ray_o = ray_o; p.ray_d = ray_d; p.near = near; p.far = far; p.t_vals = t_vals; p.t_rand = t_rand;
  p.n_rays = n_rays; p.S = S; p.norm_th = norm_th;
  p.R = R; p.Th = Th; p.bounds = bounds;
  for (int a = 0; a < 3; ++a) p.dim[a] = dims[a];
  int64_t n = n_rays * S;
  int64_t n_blocks = (n + MB - 1) / MB;
  int64_t chunk_samples = (int64_t)chunk_rays * S;
  int64_t n_chunks = (n + chunk_samples - 1) / chunk_samples;
  int64_t blocks_per_chunk = chunk_samples / MB;
  ANI_CUDA(cudaMemsetAsync(fb.chunk_argmin, 0xff, n_chunks * 8, st));
  mask_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(p, dist_plane, chunk_samples, fb.mask_words, fb.block_counts, fb.chunk_argmin);
  ANI_LAUNCHED();
  finalize_mask_kernel<<<1, 1024, 0, st>>>(n_chunks, blocks_per_chunk, n_blocks, chunk_samples, fb.chunk_argmin, fb.mask_words, fb.block_counts,
                                           fb.block_offsets, n_active, chunk_offsets);
  ANI_LAUNCHED();
  compact_samples_kernel<<<(unsigned)n_blocks, 256, 0, st>>>(p, fb.mask_words, fb.block_offsets, index, ppts, viewdir, dists);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int launch_composite_fused(const float *raw, const float *near, const float *far, const float *t_vals, const float *z_vals, int64_t n_rays,
                           int S, int white_bkgd, float *rgb_map, float *acc_map, float *depth_map, const aninerf_peer_gather *peers,
                           int chunk_rays, const uint32_t *mask_words, const int32_t *block_offsets, cudaStream_t st) {
  unsigned blocks = (unsigned)((n_rays + 7) / 8);
  PeerScatter ps{};
  if (peers) {
    ps.world = peers->world;
    ps.rank = peers->rank;
    ps.chunk_rays = chunk_rays;
    for (int k = 0; k < peers->world; ++k) ps.maps[k] = (float *)peers->maps[k];
  }
  if (S == 64)
    composite_kernel<2><<<blocks, 256, 0, st>>>((const float4 *)raw, z_vals, near, far, t_vals, n_rays, white_bkgd, rgb_map, acc_map,
                                                depth_map, nullptr, nullptr, ps, mask_words, block_offsets);
  else
    composite_kernel<1><<<blocks, 256, 0, st>>>((const float4 *)raw, z_vals, near, far, t_vals, n_rays, white_bkgd, rgb_map, acc_map,
                                                depth_map, nullptr, nullptr, ps, mask_words, block_offsets);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int launch_mask_points(const float *wpts, int64_t n, const float *R, const float *Th, const float *bounds,
                       const int32_t dims[3], const float *dist_plane, float norm_th, int64_t chunk_pts, uint8_t *mask,
                       unsigned long long *chunk_argmin, float *ppts, cudaStream_t st) {
  PointSetup ps;
  ps.R = R; ps.Th = Th; ps.bounds = bounds;
  for (int a = 0; a < 3; ++a) ps.dim[a] = dims[a];
  int64_t n_chunks = (n + chunk_pts - 1) / chunk_pts;
  ANI_CUDA(cudaMemsetAsync(chunk_argmin, 0xff, n_chunks * 8, st));
  mask_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(wpts, n, ps, norm_th, dist_plane, chunk_pts, mask, chunk_argmin, ppts);
  ANI_LAUNCHED();
  force_argmin_points_kernel<<<(unsigned)((n_chunks * 32 + 127) / 128), 128, 0, st>>>(n_chunks, chunk_pts, n, chunk_argmin, mask);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int launch_gather_points(const float *src, const int32_t *index, const int32_t *count, int64_t cap, float *dst, cudaStream_t st) {
  if (cap == 0) return ANINERF_OK;
  gather_points_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(src, index, count, dst);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int launch_scatter_scalar(const float *src, const int32_t *index, const int32_t *count, int64_t cap, float *dst, cudaStream_t st) {
  if (cap == 0) return ANINERF_OK;
  scatter_scalar_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, st>>>(src, index, count, dst);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // namespace aninerf
