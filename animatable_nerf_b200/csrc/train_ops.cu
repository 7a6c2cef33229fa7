// Element-wise / per-point kernels of the training step (tpose_trainer.py:21-73 over Network.forward,
// tpose_nerf_network.py:139-215): the pieces between the dense layers, forward where the fused render path
// has no stand-alone form, and the backward of every differentiable stage -- positional encoding,
// softmax(log(smpl_bw) + delta), inverse LBS, trilinear volume sampling w.r.t. the coordinates,
// the tail of Network.forward, raw2outputs, and the two losses.  All fp32, point-major rows.
#include "common.cuh"

namespace aninerf {

// ---------------------------------------------------------------------------------------------
// positional encoding (embedder.py:11-36): [x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pe_forward_kernel(const float *__restrict__ x, int64_t n, int L, float *__restrict__ out, int64_t ld) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int64_t p = i / 3;
  const int c = (int)(i - p * 3);
  const float v = x[i];
  float *o = out + p * ld;
  o[c] = v;
  for (int f = 0; f < L; ++f) {
    float s, co;
    sincosf(v * (float)(1 << f), &s, &co);   // the product is exact (power of two)
    o[3 + 6 * f + c] = s;
    o[3 + 6 * f + 3 + c] = co;
  }
}

// dx_c = dPE[c] + sum_f 2^f (cos(2^f x_c) dPE[3+6f+c] - sin(2^f x_c) dPE[6+6f+c])
__global__ void __launch_bounds__(256) pe_backward_kernel(const float *__restrict__ x, const float *__restrict__ dpe, int64_t ld, int64_t n, int L,
                                                          float *__restrict__ dx, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int64_t p = i / 3;
  const int c = (int)(i - p * 3);
  const float v = x[i];
  const float *g = dpe + p * ld;
  float acc = g[c];
  for (int f = 0; f < L; ++f) {
    const float sc = (float)(1 << f);
    float s, co;
    sincosf(v * sc, &s, &co);
    acc += sc * (co * g[3 + 6 * f + c] - s * g[6 + 6 * f + c]);
  }
  dx[i] = accumulate ? dx[i] + acc : acc;
}

// ---------------------------------------------------------------------------------------------
// bw = softmax(log(smpl_bw + 1e-9) + delta) over 24 bones (tpose_nerf_network.py:74-76)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bw_softmax_forward_kernel(const float *__restrict__ init, int64_t ld_init, const float *__restrict__ delta,
                                                                 int64_t n, float *__restrict__ bw) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v[ANINERF_N_BONES], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < ANINERF_N_BONES; ++k) {
    v[k] = logf(init[i * ld_init + k] + 1e-9f) + delta[i * ANINERF_N_BONES + k];
    mx = fmaxf(mx, v[k]);
  }
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < ANINERF_N_BONES; ++k) {
    v[k] = expf(v[k] - mx);
    sum += v[k];
  }
  const float inv = 1.0f / sum;
#pragma unroll
  for (int k = 0; k < ANINERF_N_BONES; ++k) bw[i * ANINERF_N_BONES + k] = v[k] * inv;
}

// d logits_k = bw_k (dbw_k - sum_j dbw_j bw_j);  d init_k = d logits_k / (init_k + 1e-9)
__global__ void __launch_bounds__(256) bw_softmax_backward_kernel(const float *__restrict__ init, int64_t ld_init, const float *__restrict__ bw,
                                                                  const float *__restrict__ dbw, int64_t n, float *__restrict__ d_delta,
                                                                  float *__restrict__ d_init) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < ANINERF_N_BONES; ++k) dot += dbw[i * ANINERF_N_BONES + k] * bw[i * ANINERF_N_BONES + k];
#pragma unroll
  for (int k = 0; k < ANINERF_N_BONES; ++k) {
    const float dl = bw[i * ANINERF_N_BONES + k] * (dbw[i * ANINERF_N_BONES + k] - dot);
    d_delta[i * ANINERF_N_BONES + k] = dl;
    if (d_init) d_init[i * ANINERF_N_BONES + k] = dl / (init[i * ld_init + k] + 1e-9f);
  }
}

// ---------------------------------------------------------------------------------------------
// inverse LBS backward (blend_utils.py:41-59): x_c = M^-1 (x - t), M = sum w_k R_k, t = sum w_k t_k
//   d w_k = -(M^-T g) . (R_k x_c + t_k)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) inverse_lbs_backward_kernel(const float *__restrict__ bw, const float *__restrict__ A,
                                                                   const float *__restrict__ tpts, const float *__restrict__ d_tpts, int64_t n,
                                                                   float *__restrict__ d_bw, int accumulate) {
  __shared__ float sA[ANINERF_N_BONES][12];
  for (int k = threadIdx.x; k < ANINERF_N_BONES * 12; k += blockDim.x) sA[k / 12][k % 12] = A[(k / 12) * 16 + (k % 12)];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M[9];
#pragma unroll
  for (int j = 0; j < 9; ++j) M[j] = 0.f;
  for (int k = 0; k < ANINERF_N_BONES; ++k) {
    const float w = bw[i * ANINERF_N_BONES + k];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) M[3 * r + c] = fmaf(w, sA[k][4 * r + c], M[3 * r + c]);
  }
  const float a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], kk = M[8];
  const float c00 = e * kk - f * h, c01 = c * h - b * kk, c02 = b * f - c * e;
  const float c10 = f * g - d * kk, c11 = a * kk - c * g, c12 = c * d - a * f;
  const float c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
  const float inv = 1.0f / (a * c00 + b * c10 + c * c20);
  const float gx = d_tpts[3 * i], gy = d_tpts[3 * i + 1], gz = d_tpts[3 * i + 2];
  // u = M^-T g : (M^-1)_{rc} = C_rc * inv  ->  u_c = sum_r C_rc g_r * inv
  const float ux = (c00 * gx + c10 * gy + c20 * gz) * inv;
  const float uy = (c01 * gx + c11 * gy + c21 * gz) * inv;
  const float uz = (c02 * gx + c12 * gy + c22 * gz) * inv;
  const float x = tpts[3 * i], y = tpts[3 * i + 1], z = tpts[3 * i + 2];
  for (int k = 0; k < ANINERF_N_BONES; ++k) {
    const float vx = sA[k][0] * x + sA[k][1] * y + sA[k][2] * z + sA[k][3];
    const float vy = sA[k][4] * x + sA[k][5] * y + sA[k][6] * z + sA[k][7];
    const float vz = sA[k][8] * x + sA[k][9] * y + sA[k][10] * z + sA[k][11];
    const float r = -(ux * vx + uy * vy + uz * vz);
    d_bw[i * ANINERF_N_BONES + k] = accumulate ? d_bw[i * ANINERF_N_BONES + k] + r : r;
  }
}

// ---------------------------------------------------------------------------------------------
// trilinear volume sampling, gradient w.r.t. the query point (ATen grid_sampler_3d_backward with
// align_corners=True, padding border: the coordinate gradient is zero where the coordinate is clipped),
// chained through the normalisation of blend_utils.py:131-139:  u_a = ((p_a - lo_a)/ext_a) * (dim_a - 1)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sample_bw_backward_kernel(const float *__restrict__ pts, int64_t n, const float *__restrict__ vol,
                                                                 const float *__restrict__ bounds, int X, int Y, int Z,
                                                                 const float *__restrict__ d_out, float *__restrict__ d_pts, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int dim[3] = {X, Y, Z};
  float fl[3], fr[3], mult[3];
  int i0[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = bounds[a], ext = bounds[3 + a] - bounds[a];
    float nrm = (pts[3 * i + a] - lo) / ext;
    nrm = nrm * 2.0f - 1.0f;
    const float lim = (float)(dim[a] - 1);
    float u = ((nrm + 1.0f) * 0.5f) * lim;
    float gm = lim * 0.5f;                       // d u / d nrm
    if (u <= 0.f) { u = 0.f; gm = 0.f; }         // clip_coordinates_set_grad
    else if (u >= lim) { u = lim; gm = 0.f; }
    const float f = floorf(u);
    i0[a] = (int)f;
    fl[a] = (f + 1.0f) - u;
    fr[a] = u - f;
    mult[a] = gm * 2.0f / ext;                   // d u / d p
  }
  float g[3] = {0.f, 0.f, 0.f};
  for (int kx = 0; kx < 2; ++kx)
    for (int ky = 0; ky < 2; ++ky)
      for (int kz = 0; kz < 2; ++kz) {
        const int xi = i0[0] + kx, yi = i0[1] + ky, zi = i0[2] + kz;
        if (xi >= X || yi >= Y || zi >= Z) continue;
        const float *row = vol + ((int64_t)(xi * Y + yi) * Z + zi) * ANINERF_BW_CH;
        float dot = 0.f;
#pragma unroll
        for (int ch = 0; ch < ANINERF_N_BONES; ++ch) dot = fmaf(__ldg(row + ch), d_out[i * ANINERF_N_BONES + ch], dot);
        const float wx = kx ? fr[0] : fl[0], wy = ky ? fr[1] : fl[1], wz = kz ? fr[2] : fl[2];
        g[0] += (kx ? 1.f : -1.f) * wy * wz * dot;
        g[1] += (ky ? 1.f : -1.f) * wx * wz * dot;
        g[2] += (kz ? 1.f : -1.f) * wx * wy * dot;
      }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float r = g[a] * mult[a];
    d_pts[3 * i + a] = accumulate ? d_pts[3 * i + a] + r : r;
  }
}

// ---------------------------------------------------------------------------------------------
// tail of Network.forward (tpose_nerf_network.py:186-212): tbounds masking, activations, scatter
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nerf_tail_forward_kernel(const float *__restrict__ sigma, const float *__restrict__ rgb,
                                                                const float *__restrict__ tpts, const float *__restrict__ tbounds,
                                                                const float *__restrict__ dists, const int32_t *__restrict__ index, int64_t n,
                                                                float4 *__restrict__ raw_full, float *__restrict__ sigma_masked) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = tpts[3 * i], y = tpts[3 * i + 1], z = tpts[3 * i + 2];
  const bool inside = x > tbounds[0] && x < tbounds[3] && y > tbounds[1] && y < tbounds[4] && z > tbounds[2] && z < tbounds[5];
  const float sg = inside ? sigma[i] : 0.f;
  sigma_masked[i] = sg;
  const float al = 1.0f - expf(-fmaxf(sg, 0.f) * dists[i]);
  raw_full[index[i]] = make_float4(1.0f / (1.0f + expf(-rgb[3 * i])), 1.0f / (1.0f + expf(-rgb[3 * i + 1])), 1.0f / (1.0f + expf(-rgb[3 * i + 2])), al);
}

__global__ void __launch_bounds__(256) nerf_tail_backward_kernel(const float4 *__restrict__ d_raw_full, const float4 *__restrict__ raw_full,
                                                                 const int32_t *__restrict__ index, const float *__restrict__ sigma_masked,
                                                                 const float *__restrict__ tpts, const float *__restrict__ tbounds,
                                                                 const float *__restrict__ dists, int64_t n, float *__restrict__ d_sigma,
                                                                 float *__restrict__ d_rgb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 g = d_raw_full[index[i]], r = raw_full[index[i]];
  d_rgb[3 * i] = g.x * r.x * (1.0f - r.x);
  d_rgb[3 * i + 1] = g.y * r.y * (1.0f - r.y);
  d_rgb[3 * i + 2] = g.z * r.z * (1.0f - r.z);
  const float x = tpts[3 * i], y = tpts[3 * i + 1], z = tpts[3 * i + 2];
  const bool inside = x > tbounds[0] && x < tbounds[3] && y > tbounds[1] && y < tbounds[4] && z > tbounds[2] && z < tbounds[5];
  const float sg = sigma_masked[i];
  // alpha[outside] = 0 cuts the gradient; relu'(s) = 0 for s <= 0;  d/ds (1 - exp(-s d)) = d exp(-s d)
  d_sigma[i] = (inside && sg > 0.f) ? g.w * dists[i] * expf(-sg * dists[i]) : 0.f;
}

// stage-2 trainer (aninerf_animation_trainer.py:73-82): alpha[outside] = 0 with outside = !(tpose strictly inside tbounds
// && pnorm < norm_th); pnorm = channel 24 of the sampled posed-volume rows (leading dimension ld), or null
__global__ void __launch_bounds__(256) mask_sigma_kernel(const float *__restrict__ sigma, const float *__restrict__ tpts, const float *__restrict__ tbounds,
                                                         const float *__restrict__ pnorm, int64_t ld, float norm_th, int64_t n,
                                                         float *__restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = tpts[3 * i], y = tpts[3 * i + 1], z = tpts[3 * i + 2];
  bool inside = x > tbounds[0] && x < tbounds[3] && y > tbounds[1] && y < tbounds[4] && z > tbounds[2] && z < tbounds[5];
  if (pnorm) inside = inside && pnorm[i * ld] < norm_th;
  out[i] = inside ? sigma[i] : 0.f;
}

// ---------------------------------------------------------------------------------------------
// raw2outputs backward (nerf_net_utils.py:6-36), gradient of rgb_map only (the loss uses nothing else):
//   w_i = a_i T_i, T_i = prod_{j<i} (1 - a_j + 1e-10);  e_i = g.c_i (- sum g with a white background)
//   d c_i = w_i g ;  d a_i = T_i e_i - (sum_{j>i} w_j e_j) / (1 - a_i + 1e-10) = T_i (e_i - B_i)
// one warp per ray, lane owns SPL consecutive samples
// ---------------------------------------------------------------------------------------------
template <int SPL>
__global__ void __launch_bounds__(256) composite_backward_kernel(const float4 *__restrict__ raw, const float *__restrict__ d_rgb_map, int64_t n_rays,
                                                                 int white_bkgd, float4 *__restrict__ d_raw) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n_rays) return;
  const float g0 = d_rgb_map[3 * ray], g1 = d_rgb_map[3 * ray + 1], g2 = d_rgb_map[3 * ray + 2];
  const float gsum = white_bkgd ? g0 + g1 + g2 : 0.f;
  float4 v[SPL];
  float om[SPL];
  float prod = 1.f;
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    v[s] = raw[ray * (32 * SPL) + lane * SPL + s];
    om[s] = 1.0f - v[s].w + 1e-10f;
    prod *= om[s];
  }
  // exclusive product scan over lanes
  float incl = prod;
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl *= t;
  }
  float T = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) T = 1.f;
  // d a_i = T_i (e_i - B_i),  B_i = sum_{j>i} a_j e_j prod_{i<k<j} om_k  (the colour the ray would still collect after sample i
  // with unit transmittance): a backward linear recurrence B_{i-1} = a_i e_i + om_i B_i -- no division by (1 - a_i), which is
  // ill-conditioned for nearly opaque samples.  Across lanes: suffix scan of the affine maps B_before = p + q * B_after.
  float Ts[SPL], e[SPL];
#pragma unroll
  for (int s = 0; s < SPL; ++s) {
    Ts[s] = T;
    e[s] = g0 * v[s].x + g1 * v[s].y + g2 * v[s].z - gsum;
    T *= om[s];
  }
  float p = 0.f, q = 1.f;
#pragma unroll
  for (int s = SPL - 1; s >= 0; --s) {
    p = v[s].w * e[s] + om[s] * p;
    q = om[s] * q;
  }
  float P = p, Q = q;
  for (int o = 1; o < 32; o <<= 1) {
    const float P2 = __shfl_down_sync(0xffffffffu, P, o), Q2 = __shfl_down_sync(0xffffffffu, Q, o);
    if (lane + o < 32) {
      P = P + Q * P2;
      Q = Q * Q2;
    }
  }
  float B = __shfl_down_sync(0xffffffffu, P, 1);   // B after this lane's last sample
  if (lane == 31) B = 0.f;
#pragma unroll
  for (int s = SPL - 1; s >= 0; --s) {
    const float w = v[s].w * Ts[s];
    float4 d;
    d.x = w * g0;
    d.y = w * g1;
    d.z = w * g2;
    d.w = Ts[s] * (e[s] - B);
    d_raw[ray * (32 * SPL) + lane * SPL + s] = d;
    B = v[s].w * e[s] + om[s] * B;
  }
}

// ---------------------------------------------------------------------------------------------
// losses (tpose_trainer.py:48-63).  Single-block kernels: fixed-order reductions, training sizes are small.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float *red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

// img_loss = mean((rgb_map[mask] - rgb[mask])^2);  d rgb_map = 2 (rgb_map - rgb) / (3 n_mask) on masked rays
__global__ void __launch_bounds__(1024) img_loss_kernel(const float *__restrict__ rgb_map, const float *__restrict__ rgb_gt, const uint8_t *__restrict__ mask,
                                                        int64_t n_rays, float *__restrict__ loss, float *__restrict__ d_rgb_map) {
  __shared__ float red[32];
  float cnt = 0.f, sq = 0.f;
  for (int64_t r = threadIdx.x; r < n_rays; r += blockDim.x)
    if (mask[r]) {
      cnt += 1.f;
      for (int c = 0; c < 3; ++c) {
        const float d = rgb_map[3 * r + c] - rgb_gt[3 * r + c];
        sq += d * d;
      }
    }
  const float n_mask = block_sum(cnt, red);
  const float total = block_sum(sq, red);
  const float denom = 3.f * n_mask;
  if (threadIdx.x == 0) *loss = total / denom;
  for (int64_t r = threadIdx.x; r < n_rays; r += blockDim.x)
    for (int c = 0; c < 3; ++c) d_rgb_map[3 * r + c] = mask[r] ? 2.f * (rgb_map[3 * r + c] - rgb_gt[3 * r + c]) / denom : 0.f;
}

// alpha_ind of tpose_nerf_network.py:192-194: sigma_masked > train_th, plus the first arg-max row of every chunk.
// One block per chunk (rows [chunk_offsets[c], chunk_offsets[c+1])).  n_sel accumulates the selected rows.
__global__ void __launch_bounds__(256) select_rows_kernel(const float *__restrict__ sigma_masked, const int32_t *__restrict__ chunk_offsets,
                                                          float train_th, uint8_t *__restrict__ sel, int32_t *__restrict__ n_sel) {
  __shared__ unsigned long long best[8];
  __shared__ int cnts[8];
  const int b = chunk_offsets[blockIdx.x], e = chunk_offsets[blockIdx.x + 1];
  unsigned long long mine = 0ull;
  int cnt = 0;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    const float s = sigma_masked[i];
    const bool on = s > train_th;
    sel[i] = on ? 1 : 0;
    cnt += on ? 1 : 0;
    uint32_t u = __float_as_uint(s);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);                        // order-preserving key
    const unsigned long long key = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(i - b));   // max key = max value, then FIRST index
    mine = key > mine ? key : mine;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, mine, o);
    mine = t > mine ? t : mine;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    best[threadIdx.x >> 5] = mine;
    cnts[threadIdx.x >> 5] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0 && e > b) {
    unsigned long long m = 0ull;
    int c = 0;
    for (int w = 0; w < 8; ++w) {
      m = best[w] > m ? best[w] : m;
      c += cnts[w];
    }
    const int arg = b + (int)(0xffffffffu - (uint32_t)(m & 0xffffffffull));
    if (!sel[arg]) {
      sel[arg] = 1;
      c += 1;
    }
    atomicAdd(n_sel, c);
  }
}

// bw_loss = smooth_l1(pbw[sel], tbw[sel]) (mean, beta = 1);  d pbw = g, d tbw = -g on the selected rows, 0 elsewhere
__global__ void __launch_bounds__(1024) bw_loss_kernel(const float *__restrict__ pbw, const float *__restrict__ tbw, const uint8_t *__restrict__ sel,
                                                       const int32_t *__restrict__ n_sel, int64_t n, float *__restrict__ loss,
                                                       float *__restrict__ d_pbw, float *__restrict__ d_tbw) {
  __shared__ float red[32];
  const float denom = (float)(*n_sel) * (float)ANINERF_N_BONES;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < n * ANINERF_N_BONES; i += blockDim.x) {
    float g = 0.f;
    if (sel[i / ANINERF_N_BONES]) {
      const float d = pbw[i] - tbw[i];
      const float ad = fabsf(d);
      acc += ad < 1.f ? 0.5f * d * d : ad - 0.5f;
      g = (ad < 1.f ? d : (d > 0.f ? 1.f : -1.f)) / denom;
    }
    d_pbw[i] = g;
    d_tbw[i] = -g;
  }
  const float total = block_sum(acc, red);
  if (threadIdx.x == 0) *loss = total / denom;
}

// ---- the selected rows (ascending order) -> the (n_sel, 24) `pbw` / `tbw` outputs of the contract (tpose_nerf_network.py:195-196) ----
// per chunk: number of selected rows
__global__ void __launch_bounds__(256) count_sel_kernel(const uint8_t *__restrict__ sel, const int32_t *__restrict__ chunk_offsets,
                                                        int32_t *__restrict__ counts) {
  __shared__ int cnts[8];
  const int b = chunk_offsets[blockIdx.x], e = chunk_offsets[blockIdx.x + 1];
  int cnt = 0;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) cnt += sel[i] ? 1 : 0;
  for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) cnts[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int w = 0; w < 8; ++w) c += cnts[w];
    counts[blockIdx.x] = c;
  }
}

// exclusive scan of the per-chunk counts in place (+ the total at [n]); one block, chunks are few (a 1024x1024 frame has ~120)
__global__ void __launch_bounds__(1024) scan_sel_kernel(int32_t *__restrict__ counts, int n) {
  __shared__ int part[1024];
  const int per = (n + 1023) / 1024;
  const int b = threadIdx.x * per, e = min(n, b + per);
  int s = 0;
  for (int i = b; i < e; ++i) s += counts[i];
  part[threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int t = 0; t < 1024; ++t) {
      const int v = part[t];
      part[t] = run;
      run += v;
    }
    counts[n] = run;
  }
  __syncthreads();
  int run = part[threadIdx.x];
  for (int i = b; i < e; ++i) {
    const int v = counts[i];
    counts[i] = run;
    run += v;
  }
}

// stable gather of the selected 24-float rows of a chunk: positions by ballot + prefix inside 256-row tiles, rows copied as
// six float4 by all threads (coalesced on both sides)
__global__ void __launch_bounds__(256) gather_sel_rows_kernel(const uint8_t *__restrict__ sel, const int32_t *__restrict__ chunk_offsets,
                                                              const int32_t *__restrict__ sel_offsets, const float4 *__restrict__ src_a,
                                                              const float4 *__restrict__ src_b, float4 *__restrict__ dst_a,
                                                              float4 *__restrict__ dst_b) {
  constexpr int Q = ANINERF_N_BONES / 4;             // float4 per row
  __shared__ int pos[256];
  __shared__ int warp_cnt[8];
  const int b = chunk_offsets[blockIdx.x], e = chunk_offsets[blockIdx.x + 1];
  int base = sel_offsets[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t0 = b; t0 < e; t0 += 256) {
    const int i = t0 + (int)threadIdx.x;
    const bool on = i < e && sel[i] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < 8; ++w) {
      before += w < warp ? warp_cnt[w] : 0;
      total += warp_cnt[w];
    }
    pos[threadIdx.x] = on ? before + __popc(bal & ((1u << lane) - 1u)) : -1;
    __syncthreads();
    const int rows = min(256, e - t0);
    for (int j = threadIdx.x; j < rows * Q; j += 256) {
      const int r = j / Q, q = j - r * Q;
      const int p = pos[r];
      if (p >= 0) {
        const int64_t s = (int64_t)(t0 + r) * Q + q, d = (int64_t)(base + p) * Q + q;
        dst_a[d] = src_a[s];
        if (src_b) dst_b[d] = src_b[s];
      }
    }
    base += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) fill_kernel(float *__restrict__ p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace aninerf

using namespace aninerf;

static inline unsigned blocks_for(int64_t n, int per = 256) { return (unsigned)((n + per - 1) / per); }

extern "C" {

int aninerf_pe_forward(const float *x, int64_t n, int32_t n_freq, float *out, int64_t ld, void *stream) {
  ANI_CHECK_ARG(x && out && n >= 0 && n_freq >= 0 && n_freq <= 16 && ld >= 3 + 6 * n_freq);
  if (n == 0) return ANINERF_OK;
  pe_forward_kernel<<<blocks_for(n * 3), 256, 0, (cudaStream_t)stream>>>(x, n, n_freq, out, ld);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_pe_backward(const float *x, const float *d_pe, int64_t ld, int64_t n, int32_t n_freq, float *d_x, int32_t accumulate, void *stream) {
  ANI_CHECK_ARG(x && d_pe && d_x && n >= 0 && n_freq >= 0 && n_freq <= 16 && ld >= 3 + 6 * n_freq);
  if (n == 0) return ANINERF_OK;
  pe_backward_kernel<<<blocks_for(n * 3), 256, 0, (cudaStream_t)stream>>>(x, d_pe, ld, n, n_freq, d_x, accumulate);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_bw_softmax_forward(const float *init, int64_t ld_init, const float *delta, int64_t n, float *bw, void *stream) {
  ANI_CHECK_ARG(init && delta && bw && n >= 0 && ld_init >= ANINERF_N_BONES);
  if (n == 0) return ANINERF_OK;
  bw_softmax_forward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(init, ld_init, delta, n, bw);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_bw_softmax_backward(const float *init, int64_t ld_init, const float *bw, const float *d_bw, int64_t n, float *d_delta, float *d_init,
                                void *stream) {
  ANI_CHECK_ARG(init && bw && d_bw && d_delta && n >= 0 && ld_init >= ANINERF_N_BONES);
  if (n == 0) return ANINERF_OK;
  bw_softmax_backward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(init, ld_init, bw, d_bw, n, d_delta, d_init);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_inverse_lbs_backward(const float *bw, const float *A, const float *tpts, const float *d_tpts, int64_t n, float *d_bw, int32_t accumulate,
                                 void *stream) {
  ANI_CHECK_ARG(bw && A && tpts && d_tpts && d_bw && n >= 0);
  if (n == 0) return ANINERF_OK;
  inverse_lbs_backward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(bw, A, tpts, d_tpts, n, d_bw, accumulate);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_sample_blend_weights_backward(const float *pts, int64_t n, const float *vol, const int32_t dims[3], const float *bounds,
                                          const float *d_out, float *d_pts, int32_t accumulate, void *stream) {
  ANI_CHECK_ARG(pts && vol && dims && bounds && d_out && d_pts && n >= 0);
  if (n == 0) return ANINERF_OK;
  sample_bw_backward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(pts, n, vol, bounds, dims[0], dims[1], dims[2], d_out, d_pts, accumulate);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_nerf_tail_forward(const float *sigma, const float *rgb, const float *tpts, const float *tbounds, const float *dists, const int32_t *index,
                              int64_t n, float *raw_full, float *sigma_masked, void *stream) {
  ANI_CHECK_ARG(sigma && rgb && tpts && tbounds && dists && index && raw_full && sigma_masked && n >= 0);
  if (n == 0) return ANINERF_OK;
  nerf_tail_forward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(sigma, rgb, tpts, tbounds, dists, index, n, (float4 *)raw_full,
                                                                           sigma_masked);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_nerf_tail_backward(const float *d_raw_full, const float *raw_full, const int32_t *index, const float *sigma_masked, const float *tpts,
                               const float *tbounds, const float *dists, int64_t n, float *d_sigma, float *d_rgb, void *stream) {
  ANI_CHECK_ARG(d_raw_full && raw_full && index && sigma_masked && tpts && tbounds && dists && d_sigma && d_rgb && n >= 0);
  if (n == 0) return ANINERF_OK;
  nerf_tail_backward_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>((const float4 *)d_raw_full, (const float4 *)raw_full, index, sigma_masked,
                                                                            tpts, tbounds, dists, n, d_sigma, d_rgb);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_mask_sigma(const float *sigma, const float *tpts, const float *tbounds, const float *pnorm, int64_t ld, float norm_th, int64_t n,
                       float *out, void *stream) {
  ANI_CHECK_ARG(sigma && tpts && tbounds && out && n >= 0);
  if (n == 0) return ANINERF_OK;
  mask_sigma_kernel<<<blocks_for(n), 256, 0, (cudaStream_t)stream>>>(sigma, tpts, tbounds, pnorm, ld, norm_th, n, out);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_composite_backward(const float *raw, const float *d_rgb_map, int64_t n_rays, int32_t n_samples, int32_t white_bkgd, float *d_raw,
                               void *stream) {
  ANI_CHECK_ARG(raw && d_rgb_map && d_raw && n_rays >= 0 && (n_samples == 32 || n_samples == 64));
  if (n_rays == 0) return ANINERF_OK;
  const unsigned blocks = (unsigned)((n_rays + 7) / 8);
  if (n_samples == 64)
    composite_backward_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)raw, d_rgb_map, n_rays, white_bkgd, (float4 *)d_raw);
  else
    composite_backward_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float4 *)raw, d_rgb_map, n_rays, white_bkgd, (float4 *)d_raw);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_img_loss(const float *rgb_map, const float *rgb_gt, const uint8_t *mask, int64_t n_rays, float *loss, float *d_rgb_map, void *stream) {
  ANI_CHECK_ARG(rgb_map && rgb_gt && mask && loss && d_rgb_map && n_rays > 0);
  img_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(rgb_map, rgb_gt, mask, n_rays, loss, d_rgb_map);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_select_rows(const float *sigma_masked, const int32_t *chunk_offsets, int32_t n_chunks, float train_th, uint8_t *sel, int32_t *n_sel,
                        void *stream) {
  ANI_CHECK_ARG(sigma_masked && chunk_offsets && sel && n_sel && n_chunks >= 0);
  ANI_CUDA(cudaMemsetAsync(n_sel, 0, 4, (cudaStream_t)stream));
  if (n_chunks == 0) return ANINERF_OK;
  select_rows_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)stream>>>(sigma_masked, chunk_offsets, train_th, sel, n_sel);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_gather_selected_rows(const uint8_t *sel, const int32_t *chunk_offsets, int32_t n_chunks, const float *src_a, const float *src_b,
                                 float *dst_a, float *dst_b, int32_t *sel_offsets, void *stream) {
  ANI_CHECK_ARG(sel && chunk_offsets && src_a && dst_a && sel_offsets && n_chunks >= 0 && (!src_b || dst_b));
  ANI_CHECK_ARG((((uintptr_t)src_a | (uintptr_t)src_b | (uintptr_t)dst_a | (uintptr_t)dst_b) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_chunks == 0) {
    ANI_CUDA(cudaMemsetAsync(sel_offsets, 0, 4, st));
    return ANINERF_OK;
  }
  count_sel_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(sel, chunk_offsets, sel_offsets);
  ANI_LAUNCHED();
  scan_sel_kernel<<<1, 1024, 0, st>>>(sel_offsets, n_chunks);
  ANI_LAUNCHED();
  gather_sel_rows_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(sel, chunk_offsets, sel_offsets, (const float4 *)src_a, (const float4 *)src_b,
                                                             (float4 *)dst_a, (float4 *)dst_b);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

int aninerf_bw_loss(const float *pbw, const float *tbw, const uint8_t *sel, const int32_t *n_sel, int64_t n, float *loss, float *d_pbw, float *d_tbw,
                    void *stream) {
  ANI_CHECK_ARG(pbw && tbw && sel && n_sel && loss && d_pbw && d_tbw && n > 0);
  bw_loss_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pbw, tbw, sel, n_sel, n, loss, d_pbw, d_tbw);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // extern "C"

// =============================================================================================
// K-nearest-vertex blend weights of the extended networks:
// sample_blend_closest_points, lib/utils/sample_utils.py:323-349 over pytorch3d.ops.knn_points (K = 5): the K nearest SMPL
// vertices of every sample, inverse-distance weights w_k = (1/(d_k + eps)) / sum, the weighted blend weights and distance.
// All vertices (6890 x 16 B = 110 KB) are staged in shared memory; a thread scans them for one point (warp-uniform broadcast
// reads) keeping the K best in registers.  Squared distances are (dx*dx + dy*dy) + dz*dz, separately rounded, ties to the lower
// vertex index -- the order of the brute-force oracle.
// =============================================================================================
namespace aninerf {

constexpr int KNN_MAX_K = 8;
constexpr int KNN_TILE = 7168;          // vertices per shared-memory tile (112 KB of float4)

template <int K>
__global__ void __launch_bounds__(256) knn_blend_kernel(const float *__restrict__ pts, int64_t n, const float *__restrict__ verts, int n_verts,
                                                        const float *__restrict__ values, float eps, float *__restrict__ bw_out,
                                                        float *__restrict__ dist_out) {
  extern __shared__ float4 s_v[];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < n;
  float px = 0.f, py = 0.f, pz = 0.f;
  if (live) {
    px = pts[3 * i];
    py = pts[3 * i + 1];
    pz = pts[3 * i + 2];
  }
  float bd[K];
  int bi[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    bd[k] = INFINITY;
    bi[k] = 0;
  }
  for (int t0 = 0; t0 < n_verts; t0 += KNN_TILE) {
    const int cnt = min(KNN_TILE, n_verts - t0);
    __syncthreads();
    for (int v = threadIdx.x; v < cnt; v += blockDim.x)
      s_v[v] = make_float4(verts[3 * (int64_t)(t0 + v)], verts[3 * (int64_t)(t0 + v) + 1], verts[3 * (int64_t)(t0 + v) + 2], 0.f);
    __syncthreads();
    for (int v = 0; v < cnt; ++v) {
      const float4 q = s_v[v];
      const float dx = __fsub_rn(px, q.x), dy = __fsub_rn(py, q.y), dz = __fsub_rn(pz, q.z);
      const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d < bd[K - 1]) {
        // insertion into the sorted K best (strict <: an equal distance keeps the earlier vertex in front)
        bd[K - 1] = d;
        bi[K - 1] = t0 + v;
#pragma unroll
        for (int k = K - 1; k > 0; --k) {
          if (bd[k] < bd[k - 1]) {
            const float td = bd[k];
            bd[k] = bd[k - 1];
            bd[k - 1] = td;
            const int ti = bi[k];
            bi[k] = bi[k - 1];
            bi[k - 1] = ti;
          }
        }
      }
    }
  }
  if (!live) return;
  // guard_knn_points: dists = sqrt(d2); disp = 1 / (dists + eps); weights = disp / sum(disp)   (sample_utils.py:311-313, 340-342)
  float dist[K], w[K], sum = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    dist[k] = sqrtf(bd[k]);
    w[k] = 1.0f / (dist[k] + eps);
    sum += w[k];
  }
  float dw = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    w[k] = w[k] / sum;
    dw = fmaf(dist[k], w[k], dw);
  }
  if (dist_out) dist_out[i] = dw;
  float acc[ANINERF_N_BONES];
#pragma unroll
  for (int c = 0; c < ANINERF_N_BONES; ++c) acc[c] = 0.f;
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float4 *row = reinterpret_cast<const float4 *>(values + (int64_t)bi[k] * ANINERF_N_BONES);
#pragma unroll
    for (int q = 0; q < ANINERF_N_BONES / 4; ++q) {
      const float4 r = __ldg(row + q);
      acc[4 * q] = fmaf(r.x, w[k], acc[4 * q]);
      acc[4 * q + 1] = fmaf(r.y, w[k], acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(r.z, w[k], acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(r.w, w[k], acc[4 * q + 3]);
    }
  }
  float4 *o = reinterpret_cast<float4 *>(bw_out + i * ANINERF_N_BONES);
#pragma unroll
  for (int q = 0; q < ANINERF_N_BONES / 4; ++q) o[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
}

template <int K>
static int launch_knn(const float *pts, int64_t n, const float *verts, int n_verts, const float *values, float eps, float *bw_out, float *dist_out,
                      cudaStream_t st) {
  static bool configured = false;
  const int smem = KNN_TILE * (int)sizeof(float4);
  if (!configured) {
    ANI_CUDA(cudaFuncSetAttribute(knn_blend_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  knn_blend_kernel<K><<<(unsigned)((n + 255) / 256), 256, smem, st>>>(pts, n, verts, n_verts, values, eps, bw_out, dist_out);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // namespace aninerf

extern "C" int aninerf_knn_blend_weights(const float *pts, int64_t n, const float *verts, int32_t n_verts, const float *values, int32_t K, float eps,
                                         float *bw_out, float *dist_out, void *stream) {
  ANI_CHECK_ARG(pts && verts && values && bw_out && n >= 0 && n_verts >= K && K >= 1 && K <= aninerf::KNN_MAX_K);
  ANI_CHECK_ARG((((uintptr_t)values | (uintptr_t)bw_out) & 15) == 0);
  if (n == 0) return ANINERF_OK;
  cudaStream_t st = (cudaStream_t)stream;
  switch (K) {
    case 1: return aninerf::launch_knn<1>(pts, n, verts, n_verts, values, eps, bw_out, dist_out, st);
    case 5: return aninerf::launch_knn<5>(pts, n, verts, n_verts, values, eps, bw_out, dist_out, st);
    case 8: return aninerf::launch_knn<8>(pts, n, verts, n_verts, values, eps, bw_out, dist_out, st);
    default: return aninerf::fail(ANINERF_EINVAL, "%s: K must be 1, 5 or 8%s", __func__);
  }
}
