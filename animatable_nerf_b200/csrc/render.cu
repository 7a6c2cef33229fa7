// The fused path behind Renderer.render(batch) (tpose_renderer.py:159-186 over Network.forward,
// tpose_nerf_network.py:139-215) and the density query (Network.calculate_alpha, :105-137):
// a stream-ordered chain of this library's kernels with every count kept on the device.
#include <utility>
#include <vector>

#include "common.cuh"

namespace aninerf {

struct FrontEndBuffers {
  uint32_t *mask_words;
  int32_t *block_counts, *block_offsets;
  unsigned long long *chunk_argmin;
};

int launch_split_volume(const float *vol, const int32_t dims[3], float *w24, float *dist, cudaStream_t st);
int launch_front_end(const float *ray_o, const float *ray_d, const float *near, const float *far, const float *t_vals,
                     const float *t_rand, int64_t n_rays, int S, int chunk_rays, const float *R, const float *Th,
                     const float *bounds, const int32_t dims[3], const float *dist_plane, float norm_th, FrontEndBuffers fb,
                     int32_t *index, float *ppts, float *viewdir, float *dists, int32_t *n_active, int32_t *chunk_offsets,
                     const aninerf_silhouettes *sil, cudaStream_t st);
int launch_composite_fused(const float *raw, const float *near, const float *far, const float *t_vals, const float *z_vals, int64_t n_rays,
                           int S, int white_bkgd, float *rgb_map, float *acc_map, float *depth_map, const aninerf_peer_gather *peers,
                           int chunk_rays, const uint32_t *mask_words, const int32_t *block_offsets, cudaStream_t st);
int launch_mask_points(const float *wpts, int64_t n, const float *R, const float *Th, const float *bounds, const int32_t dims[3],
                       const float *dist_plane, float norm_th, int64_t chunk_pts, uint8_t *mask, unsigned long long *chunk_argmin,
                       float *ppts, cudaStream_t st);
int launch_gather_points(const float *src, const int32_t *index, const int32_t *count, int64_t cap, float *dst, cudaStream_t st);
int launch_scatter_scalar(const float *src, const int32_t *index, const int32_t *count, int64_t cap, float *dst, cudaStream_t st);
int bw_forward_impl(aninerf_net *net, int field, int latent_index, const int64_t *latent_dev, const float *pts, const float *smpl_bw, const float *vol_w24,
                    const int32_t dims[3], const float *bounds, int64_t n, const int32_t *n_dev, const float *A, float *bw_out,
                    float *tpts_out, int precision, cudaStream_t st);
int nerf_forward_impl(aninerf_net *net, int latent_index, const int64_t *latent_dev, const float *pts, const float *viewdir, int64_t n, const int32_t *n_dev,
                      float *sigma_out, float *rgb_out, const float *dists, const float *tbounds, const int32_t *index, float *raw_out,
                      float *sigma_masked_out, int precision, cudaStream_t st, int density_only);

// ---- optional per-stage timing (CUDA events on the launching stream) ---------------------------
enum { ST_SPLIT = 0, ST_CLEAR, ST_MASK, ST_SCAN, ST_COMPACT, ST_BW_POSE, ST_BW_CANON, ST_NERF, ST_COMPOSITE, ST_COUNT };
static bool g_profile = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_pending[ST_COUNT];
static std::vector<cudaEvent_t> g_free_events;
static double g_ms[ST_COUNT];
static long long g_calls[ST_COUNT];

static cudaEvent_t take_event() {
  cudaEvent_t e;
  if (!g_free_events.empty()) {
    e = g_free_events.back();
    g_free_events.pop_back();
  } else {
    cudaEventCreate(&e);
  }
  return e;
}

struct StageTimer {   // records start on construction, stop on destruction
  int stage;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  StageTimer(int stage_, cudaStream_t st_) : stage(stage_), st(st_) {
    if (g_profile) {
      a = take_event();
      b = take_event();
      cudaEventRecord(a, st);
    }
  }
  ~StageTimer() {
    if (a) {
      cudaEventRecord(b, st);
      g_pending[stage].push_back({a, b});
    }
  }
};

struct Carver {   // bump allocator over the caller's workspace (256-byte aligned pieces)
  char *base;
  int64_t off = 0, cap;
  Carver(void *p, int64_t c) : base((char *)p), cap(c) {}
  template <class T>
  T *take(int64_t count) {
    int64_t bytes = align_up(count * (int64_t)sizeof(T), 256);
    T *r = base ? (T *)(base + off) : nullptr;
    off += bytes;
    return r;
  }
};

struct RenderScratch {
  float *w24_p, *dist_p, *w24_t, *dist_t;
  FrontEndBuffers fb;
  int32_t *index;
  float *ppts, *viewdir, *dists, *tpts, *z_vals;
  float *raw_c;      // render-only: compact (n',4) rows of the active samples instead of the caller's dense raw
};

static int64_t carve_render(Carver &c, int64_t n_rays, int S, int want_bw, int64_t pv, int64_t tv, RenderScratch &s) {
  int64_t n = n_rays * S;
  int64_t n_blocks = (n + 2047) / 2048;
  int64_t n_chunks = n_blocks + 1;   // upper bound for any chunk size >= 2048 samples
  s.w24_p = c.take<float>(pv * ANINERF_N_BONES);
  s.dist_p = c.take<float>(pv);
  s.w24_t = c.take<float>(want_bw ? tv * ANINERF_N_BONES : 0);
  s.dist_t = c.take<float>(want_bw ? tv : 0);
  s.fb.mask_words = c.take<uint32_t>(n_blocks * 64);
  s.fb.block_counts = c.take<int32_t>(n_blocks + 1);
  s.fb.block_offsets = c.take<int32_t>(n_blocks + 1);
  s.fb.chunk_argmin = c.take<unsigned long long>(n_chunks);
  s.index = c.take<int32_t>(n);
  s.ppts = c.take<float>(n * 3);
  s.viewdir = c.take<float>(n * 3);
  s.dists = c.take<float>(n);
  s.tpts = c.take<float>(n * 3);
  s.z_vals = c.take<float>(n);
  s.raw_c = c.take<float>(want_bw ? 0 : n * 4);
  return c.off;
}

}  // namespace aninerf

using namespace aninerf;

static int render_rays_impl(aninerf_net *net, const aninerf_frame *fr, const aninerf_render_params *pr, const aninerf_silhouettes *sil,
                            const aninerf_peer_gather *peers, const float *ray_o, const float *ray_d, const float *near, const float *far,
                            const float *t_vals, const float *t_rand, int64_t n_rays, const aninerf_render_outputs *out, void *workspace,
                            int64_t workspace_bytes, void *stream);

extern "C" {

int64_t aninerf_render_workspace_bytes(int64_t n_rays, int32_t n_samples, int32_t want_bw, int64_t pbw_voxels, int64_t tbw_voxels) {
  Carver c(nullptr, 0);
  RenderScratch s;
  return carve_render(c, n_rays, n_samples, want_bw, pbw_voxels, tbw_voxels, s);
}

int aninerf_render_rays(aninerf_net *net, const aninerf_frame *fr, const aninerf_render_params *pr, const float *ray_o, const float *ray_d,
                        const float *near, const float *far, const float *t_vals, const float *t_rand, int64_t n_rays,
                        const aninerf_render_outputs *out, void *workspace, int64_t workspace_bytes, void *stream) {
  return render_rays_impl(net, fr, pr, nullptr, nullptr, ray_o, ray_d, near, far, t_vals, t_rand, n_rays, out, workspace, workspace_bytes, stream);
}

int aninerf_render_rays_culled(aninerf_net *net, const aninerf_frame *fr, const aninerf_render_params *pr, const aninerf_silhouettes *sil,
                               const float *ray_o, const float *ray_d, const float *near, const float *far, const float *t_vals,
                               const float *t_rand, int64_t n_rays, const aninerf_render_outputs *out, void *workspace,
                               int64_t workspace_bytes, void *stream) {
  ANI_CHECK_ARG(sil && sil->msks && sil->Ks && sil->RT && sil->n_views > 0 && sil->H > 0 && sil->W > 0);
  ANI_CHECK_ARG(pr && !pr->want_bw);
  return render_rays_impl(net, fr, pr, sil, nullptr, ray_o, ray_d, near, far, t_vals, t_rand, n_rays, out, workspace, workspace_bytes, stream);
}

int aninerf_render_rays_tiled(aninerf_net *net, const aninerf_frame *fr, const aninerf_render_params *pr, const aninerf_silhouettes *sil,
                              const aninerf_peer_gather *peers, const float *ray_o, const float *ray_d, const float *near, const float *far,
                              const float *t_vals, const float *t_rand, int64_t n_rays, const aninerf_render_outputs *out, void *workspace,
                              int64_t workspace_bytes, void *stream) {
  if (sil) {
    ANI_CHECK_ARG(sil->msks && sil->Ks && sil->RT && sil->n_views > 0 && sil->H > 0 && sil->W > 0);
    ANI_CHECK_ARG(pr && !pr->want_bw);
  }
  if (peers) {
    ANI_CHECK_ARG(peers->world >= 1 && peers->world <= ANINERF_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world);
    for (int k = 0; k < peers->world; ++k) ANI_CHECK_ARG(peers->maps[k] != nullptr);
  }
  return render_rays_impl(net, fr, pr, sil, peers, ray_o, ray_d, near, far, t_vals, t_rand, n_rays, out, workspace, workspace_bytes, stream);
}

}  // extern "C"

static int render_rays_impl(aninerf_net *net, const aninerf_frame *fr, const aninerf_render_params *pr, const aninerf_silhouettes *sil,
                            const aninerf_peer_gather *peers, const float *ray_o, const float *ray_d, const float *near, const float *far,
                            const float *t_vals, const float *t_rand, int64_t n_rays, const aninerf_render_outputs *out, void *workspace,
                            int64_t workspace_bytes, void *stream) {
  ANI_CHECK_ARG(net && fr && pr && out && ray_o && ray_d && near && far && t_vals && workspace && n_rays >= 0);
  ANI_CHECK_ARG(fr->A && fr->R && fr->Th && fr->pbw && fr->pbounds && fr->tbounds);
  ANI_CHECK_ARG(out->rgb_map && out->acc_map && out->depth_map && out->n_active);
  ANI_CHECK_ARG(out->raw || !pr->want_bw);     // the dense raw buffer is an output of the training contract only
  const int S = pr->n_samples;
  ANI_CHECK_ARG(S == 32 || S == 64);
  ANI_CHECK_ARG(pr->chunk_rays > 0 && ((int64_t)pr->chunk_rays * S) % 2048 == 0);
  ANI_CHECK_ARG(n_rays * S < (int64_t)2147483647);
  if (pr->want_bw) ANI_CHECK_ARG(fr->tbw && out->pbw_all && out->tbw_all && out->sigma_masked);
  cudaStream_t st = (cudaStream_t)stream;
  if (n_rays == 0) {
    ANI_CUDA(cudaMemsetAsync(out->n_active, 0, 4, st));
    return ANINERF_OK;
  }
  const int64_t pv = (int64_t)fr->pbw_dims[0] * fr->pbw_dims[1] * fr->pbw_dims[2];
  const int64_t tv = pr->want_bw ? (int64_t)fr->tbw_dims[0] * fr->tbw_dims[1] * fr->tbw_dims[2] : 0;
  Carver c(workspace, workspace_bytes);
  RenderScratch s;
  if (carve_render(c, n_rays, S, pr->want_bw, pv, tv, s) > workspace_bytes)
    return fail(ANINERF_ENOMEM, "%s: workspace too small%s", __func__);
  const int64_t n = n_rays * S;
  int rc;
  // per-frame volume split (weights plane for the MLP head, distance plane for the mask pass)
  {
    StageTimer t(ST_SPLIT, st);
    if ((rc = launch_split_volume(fr->pbw, fr->pbw_dims, s.w24_p, s.dist_p, st))) return rc;
    if (pr->want_bw && (rc = launch_split_volume(fr->tbw, fr->tbw_dims, s.w24_t, s.dist_t, st))) return rc;
  }
  const bool dense = out->raw != nullptr;
  if (dense) {
    StageTimer t(ST_CLEAR, st);
    ANI_CUDA(cudaMemsetAsync(out->raw, 0, n * 16, st));
  }
  int32_t *index = out->active_index ? out->active_index : s.index;
  // 1. sample -> pose -> pnorm mask -> per-chunk argmin forcing -> stable compaction
  {
    StageTimer t(ST_MASK, st);
    if ((rc = launch_front_end(ray_o, ray_d, near, far, t_vals, t_rand, n_rays, S, pr->chunk_rays, fr->R, fr->Th, fr->pbounds, fr->pbw_dims,
                               s.dist_p, pr->norm_th, s.fb, index, s.ppts, s.viewdir, s.dists, out->n_active, out->chunk_offsets, sil, st)))
      return rc;
  }
  // 2. neural blend weights at the posed points + inverse LBS -> canonical points
  const int bw_field = pr->novel_pose ? ANINERF_FIELD_NOVEL_BW : ANINERF_FIELD_BW;
  // latent indices: host values, or (no host round trip) the batch's device int64 tensors + an offset
  const int64_t *bw_lat_dev = pr->novel_pose ? fr->bw_latent_index_dev : fr->latent_index_dev;
  const int bw_latent = pr->novel_pose ? (bw_lat_dev ? 0 : fr->bw_latent_index) : (bw_lat_dev ? 1 : fr->latent_index + 1);
  const int nerf_latent = fr->latent_index_dev ? 0 : fr->latent_index;
  const int bw_prec = pr->bw_precision == 1 ? 1 : 3;
  {
    StageTimer t(ST_BW_POSE, st);
    if ((rc = bw_forward_impl(net, bw_field, bw_latent, bw_lat_dev, s.ppts, nullptr, s.w24_p, fr->pbw_dims, fr->pbounds, n, out->n_active, fr->A,
                              pr->want_bw ? out->pbw_all : nullptr, s.tpts, bw_prec, st)))
      return rc;
  }
  // 3. (training contract only) blend weights of the canonical points, latent index 0
  if (pr->want_bw) {
    StageTimer t(ST_BW_CANON, st);
    if ((rc = bw_forward_impl(net, ANINERF_FIELD_BW, 0, nullptr, s.tpts, nullptr, s.w24_t, fr->tbw_dims, fr->tbounds, n, out->n_active, nullptr,
                              out->tbw_all, nullptr, bw_prec, st)))
      return rc;
  }
  // 4. canonical NeRF field + tail of Network.forward, scattered into the dense raw buffer
  {
    StageTimer t(ST_NERF, st);
    if ((rc = nerf_forward_impl(net, nerf_latent, fr->latent_index_dev, s.tpts, s.viewdir, n, out->n_active, nullptr, nullptr, s.dists, fr->tbounds,
                                dense ? index : nullptr, dense ? out->raw : s.raw_c, pr->want_bw ? out->sigma_masked : nullptr, pr->nerf_precision == 3 ? 3 : 1, st, 0)))
      return rc;
  }
  // 5. compositing
  StageTimer t(ST_COMPOSITE, st);
  const float *z = nullptr;
  if (t_rand) {
    if ((rc = aninerf_sample_points(ray_o, ray_d, near, far, t_vals, t_rand, n_rays, S, nullptr, s.z_vals, nullptr, stream))) return rc;
    z = s.z_vals;
  }
  return launch_composite_fused(dense ? out->raw : s.raw_c, near, far, t_vals, z, n_rays, S, pr->white_bkgd, out->rgb_map, out->acc_map,
                                out->depth_map, peers, pr->chunk_rays, dense ? nullptr : s.fb.mask_words, dense ? nullptr : s.fb.block_offsets, st);
}

extern "C" {

int64_t aninerf_front_end_workspace_bytes(int64_t n_rays, int32_t n_samples, int64_t pbw_voxels) {
  Carver c(nullptr, 0);
  RenderScratch s;
  return carve_render(c, n_rays, n_samples, 0, pbw_voxels, 0, s);
}

// The front end of the fused path on its own (training step): sample -> pose -> pnorm mask -> per-chunk argmin forcing ->
// stable compaction.  Outputs have room for n_rays*S rows; *n_active / chunk_offsets stay on the device.
int aninerf_front_end(const aninerf_frame *fr, const aninerf_render_params *pr, const float *ray_o, const float *ray_d, const float *near,
                      const float *far, const float *t_vals, const float *t_rand, int64_t n_rays, int32_t *index, float *ppts, float *viewdir,
                      float *dists, float *z_vals, int32_t *n_active, int32_t *chunk_offsets, void *workspace, int64_t workspace_bytes,
                      void *stream) {
  ANI_CHECK_ARG(fr && pr && ray_o && ray_d && near && far && t_vals && index && ppts && viewdir && dists && n_active && workspace && n_rays > 0);
  ANI_CHECK_ARG(fr->R && fr->Th && fr->pbw && fr->pbounds);
  const int S = pr->n_samples;
  ANI_CHECK_ARG(S == 32 || S == 64);
  ANI_CHECK_ARG(pr->chunk_rays > 0 && ((int64_t)pr->chunk_rays * S) % 2048 == 0 && n_rays * S < (int64_t)2147483647);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t pv = (int64_t)fr->pbw_dims[0] * fr->pbw_dims[1] * fr->pbw_dims[2];
  Carver c(workspace, workspace_bytes);
  RenderScratch s;
  if (carve_render(c, n_rays, S, 0, pv, 0, s) > workspace_bytes) return fail(ANINERF_ENOMEM, "%s: workspace too small%s", __func__);
  int rc;
  if ((rc = launch_split_volume(fr->pbw, fr->pbw_dims, s.w24_p, s.dist_p, st))) return rc;
  if ((rc = launch_front_end(ray_o, ray_d, near, far, t_vals, t_rand, n_rays, S, pr->chunk_rays, fr->R, fr->Th, fr->pbounds, fr->pbw_dims, s.dist_p,
                             pr->norm_th, s.fb, index, ppts, viewdir, dists, n_active, chunk_offsets, nullptr, st)))
    return rc;
  if (z_vals) return aninerf_sample_points(ray_o, ray_d, near, far, t_vals, t_rand, n_rays, S, nullptr, z_vals, nullptr, stream);
  return ANINERF_OK;
}

int aninerf_profile_enable(int32_t on) {
  g_profile = on != 0;
  return ANINERF_OK;
}

int aninerf_profile_read(double *ms_out, int64_t *calls_out, int32_t reset) {
  ANI_CHECK_ARG(ms_out && calls_out);
  ANI_CUDA(cudaDeviceSynchronize());
  for (int s = 0; s < ST_COUNT; ++s) {
    for (auto &pr : g_pending[s]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) {
        g_ms[s] += ms;
        g_calls[s] += 1;
      }
      g_free_events.push_back(pr.first);
      g_free_events.push_back(pr.second);
    }
    g_pending[s].clear();
    ms_out[s] = g_ms[s];
    calls_out[s] = g_calls[s];
    if (reset) {
      g_ms[s] = 0.0;
      g_calls[s] = 0;
    }
  }
  return ANINERF_OK;
}

int64_t aninerf_query_workspace_bytes(int64_t n, int64_t pbw_voxels) {
  Carver c(nullptr, 0);
  c.take<float>(pbw_voxels * ANINERF_N_BONES);
  c.take<float>(pbw_voxels);
  c.take<uint8_t>(n);
  c.take<unsigned long long>(n / 32 + 2);
  c.take<float>(n * 3);
  c.take<float>(n * 3);
  c.take<float>(n * 3);
  c.take<float>(n);
  c.take<int32_t>(n);
  c.take<char>(aninerf_compact_workspace_bytes(n));
  return c.off;
}

int aninerf_query_alpha(aninerf_net *net, const aninerf_frame *fr, const float *wpts, int64_t n, int64_t chunk_pts, float norm_th,
                        int32_t novel_pose, int32_t bw_precision, float *sigma_out, int32_t *n_active, void *workspace,
                        int64_t workspace_bytes, void *stream) {
  ANI_CHECK_ARG(net && fr && wpts && sigma_out && n_active && workspace && n >= 0 && n < (int64_t)2147483647);
  ANI_CHECK_ARG(chunk_pts > 0 && chunk_pts % 256 == 0);
  ANI_CHECK_ARG(fr->A && fr->R && fr->Th && fr->pbw && fr->pbounds);
  if (workspace_bytes < aninerf_query_workspace_bytes(n, (int64_t)fr->pbw_dims[0] * fr->pbw_dims[1] * fr->pbw_dims[2]))
    return fail(ANINERF_ENOMEM, "%s: workspace too small%s", __func__);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    ANI_CUDA(cudaMemsetAsync(n_active, 0, 4, st));
    return ANINERF_OK;
  }
  const int64_t pv = (int64_t)fr->pbw_dims[0] * fr->pbw_dims[1] * fr->pbw_dims[2];
  Carver c(workspace, workspace_bytes);
  float *w24 = c.take<float>(pv * ANINERF_N_BONES);
  float *dist = c.take<float>(pv);
  uint8_t *mask = c.take<uint8_t>(n);
  unsigned long long *argmin = c.take<unsigned long long>(n / 32 + 2);
  float *ppts_all = c.take<float>(n * 3);
  float *ppts = c.take<float>(n * 3);
  float *tpts = c.take<float>(n * 3);
  float *sigma = c.take<float>(n);
  int32_t *index = c.take<int32_t>(n);
  void *cws = c.take<char>(aninerf_compact_workspace_bytes(n));
  int rc;
  if ((rc = launch_split_volume(fr->pbw, fr->pbw_dims, w24, dist, st))) return rc;
  if ((rc = launch_mask_points(wpts, n, fr->R, fr->Th, fr->pbounds, fr->pbw_dims, dist, norm_th, chunk_pts, mask, argmin, ppts_all, st)))
    return rc;
  if ((rc = aninerf_compact_rays(nullptr, nullptr, nullptr, nullptr, mask, n, nullptr, nullptr, nullptr, nullptr, index, n_active, cws,
                                 aninerf_compact_workspace_bytes(n), stream)))
    return rc;
  if ((rc = launch_gather_points(ppts_all, index, n_active, n, ppts, st))) return rc;
  const int bw_field = novel_pose ? ANINERF_FIELD_NOVEL_BW : ANINERF_FIELD_BW;
  const int64_t *bw_lat_dev = novel_pose ? fr->bw_latent_index_dev : fr->latent_index_dev;
  const int bw_latent = novel_pose ? (bw_lat_dev ? 0 : fr->bw_latent_index) : (bw_lat_dev ? 1 : fr->latent_index + 1);
  if ((rc = bw_forward_impl(net, bw_field, bw_latent, bw_lat_dev, ppts, nullptr, w24, fr->pbw_dims, fr->pbounds, n, n_active, fr->A, nullptr, tpts,
                            bw_precision == 1 ? 1 : 3, st)))
    return rc;
  // density only (TPoseHuman.calculate_alpha, tpose_nerf_network.py:241-250): the kernel stops after the trunk + alpha_fc
  if ((rc = nerf_forward_impl(net, fr->latent_index_dev ? 0 : fr->latent_index, fr->latent_index_dev, tpts, tpts, n, n_active, sigma, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 1,
                              st, 1)))
    return rc;
  ANI_CUDA(cudaMemsetAsync(sigma_out, 0, n * 4, st));
  return launch_scatter_scalar(sigma, index, n_active, n, sigma_out, st);
}

}  // extern "C"
