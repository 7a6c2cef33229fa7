/*
 * aninerf_b200.h -- C ABI of libaninerf_b200.so: the B200 (sm_100a) implementation of
 * Animatable NeRF's per-ray render hot path.
 *
 * The reference (xx-peach/animatable_nerf) has no FFI layer: the seam is duck-typed Python
 * (`Renderer.render(batch)`, `Network.*`).  Each entry point below names the reference
 * function (file:line under /root/reference) whose arithmetic it replaces; the Python mirror of
 * the reference interface (animatable_nerf_b200/tpose_renderer.py, tpose_nerf_network.py) binds
 * these with ctypes.  See INTEGRATION.md for the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - the caller owns all memory (inputs, outputs, workspace); the library allocates only the
 *    packed network weights held by an `aninerf_net` object;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises
 *    the device and never throws; it returns 0 on success or a negative ANINERF_E* code, with a
 *    message available from aninerf_last_error() (thread local);
 *  - point arrays are point-major: xyz (n,3), blend weights (n,24|25), raw (n,4);
 *  - float means IEEE fp32; "bit-exact" stages use round-to-nearest intrinsics in a fixed order.
 */
#ifndef ANINERF_B200_H
#define ANINERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANINERF_ABI_VERSION 1
#define ANINERF_N_BONES 24
#define ANINERF_BW_CH 25          /* 24 skinning weights + distance-to-surface channel */
#define ANINERF_MAX_SAMPLES 64    /* cfg.N_samples of every aninerf config (configs/aninerf_s9p.yaml:58) */
#define ANINERF_MAX_PEERS 8       /* GPUs of one NVSwitch box */
#define ANINERF_CHUNK_RAYS 2048   /* tpose_renderer.py:170 -- a semantic unit (per-chunk argmin forcing) */

enum {
  ANINERF_OK = 0,
  ANINERF_EINVAL = -1,    /* bad argument */
  ANINERF_ECUDA = -2,     /* a CUDA runtime call failed */
  ANINERF_ENOMEM = -3,    /* workspace too small / allocation failed */
  ANINERF_ESTATE = -4     /* weights not loaded etc. */
};

int aninerf_version(void);
const char *aninerf_last_error(void);

/* ------------------------------------------------------------------------------------------
 * Stage 1: rays and SMPL-box intersection (fp64 arithmetic, bit-exact targets)
 * ---------------------------------------------------------------------------------------- */

/* Camera: world->camera rotation R (row major), translation T, and inv(K) (row major), fp64. */
typedef struct {
  double Kinv[9];
  double R[9];
  double T[3];
  int32_t H, W;
} aninerf_camera;

/* get_rays, lib/utils/if_nerf/if_nerf_data_utils.py:64-89, followed by the float32 cast of its
 * callers (:328-329).  ray_o, ray_d: (H*W,3) fp32. */
int aninerf_gen_rays(const aninerf_camera *cam_host, float *ray_o, float *ray_d, void *stream);

/* get_near_far, if_nerf_data_utils.py:156-196.  bounds_host: (2,3) fp32 on the HOST.
 * near/far: (n,) fp32 written for EVERY ray (0 where mask==0); mask: (n,) uint8. */
int aninerf_near_far(const float *bounds_host, const float *ray_o, const float *ray_d, int64_t n,
                     float *near, float *far, uint8_t *mask, void *stream);

/* Stable compaction of rays by mask (the boolean indexing of get_rays_within_bounds,
 * if_nerf_data_utils.py:334-336).  Outputs hold up to n rows; *count (device int32) receives the
 * number kept.  workspace: aninerf_compact_workspace_bytes(n) bytes. */
int64_t aninerf_compact_workspace_bytes(int64_t n);
int aninerf_compact_rays(const float *ray_o, const float *ray_d, const float *near, const float *far,
                         const uint8_t *mask, int64_t n, float *ray_o_out, float *ray_d_out,
                         float *near_out, float *far_out, int32_t *index_out, int32_t *count,
                         void *workspace, int64_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage 1b: stratified sample points, Renderer.get_wsampling_points + get_density_color,
 * lib/networks/renderer/tpose_renderer.py:14-69
 * ---------------------------------------------------------------------------------------- */
/* t_vals: (S,) fp32 = torch.linspace(0,1,S) (passed in: it is not i/(S-1) rounded).
 * t_rand: (n_rays,S) fp32 or NULL (cfg.perturb jitter, drawn by the caller on the host RNG as
 * the reference does, tpose_renderer.py:35).  Any output pointer may be NULL.
 * pts (n_rays*S,3), z_vals (n_rays,S), dists (n_rays*S,) [last interval duplicated, :64-65]. */
int aninerf_sample_points(const float *ray_o, const float *ray_d, const float *near, const float *far,
                          const float *t_vals, const float *t_rand, int64_t n_rays, int32_t n_samples,
                          float *pts, float *z_vals, float *dists, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage 2: inverse linear-blend skinning pieces, lib/utils/blend_utils.py
 * ---------------------------------------------------------------------------------------- */
/* world_points_to_pose_points, blend_utils.py:6-16: (p - Th) @ R.  R (3,3), Th (3,). */
int aninerf_world_to_pose(const float *wpts, int64_t n, const float *R, const float *Th, float *ppts,
                          void *stream);

/* pts_sample_blend_weights, blend_utils.py:119-149 (F.grid_sample trilinear / border /
 * align_corners=True over a channels-last volume).  vol: (X,Y,Z,25) fp32 exactly as the reference
 * batch holds it; bounds: (2,3).  out: (n,25) point-major (the reference returns (1,25,n)). */
int aninerf_sample_blend_weights(const float *pts, int64_t n, const float *vol, const int32_t dims_host[3],
                                 const float *bounds, float *out, void *stream);

/* pose_points_to_tpose_points, blend_utils.py:41-59 (inverse LBS) and tpose_points_to_pose_points,
 * blend_utils.py:77-90 (forward LBS).  bw: (n,24) point-major; A: (24,4,4). */
int aninerf_inverse_lbs(const float *ppts, const float *bw, int64_t n, const float *A, float *tpts, void *stream);
int aninerf_forward_lbs(const float *tpts, const float *bw, int64_t n, const float *A, float *ppts, void *stream);

/* ------------------------------------------------------------------------------------------
 * Multi-view silhouette culling, Renderer.prepare_inside_pts,
 * lib/networks/renderer/tpose_renderer_mmsk.py:14-57 (batch keys `msks`, `Ks`, `RT`, `H`, `W` of
 * lib/datasets/tpose_novel_view_dataset.py:191)
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *msks;  /* (V,H,W) uint8, dilated training-view body masks */
  const float *Ks;      /* (V,3,3) intrinsics */
  const float *RT;      /* (V,4,4) world->camera */
  int32_t n_views, H, W;
} aninerf_silhouettes;

/* inside[i] = 1 iff world point i projects (round-half-even, clamped to the image) onto a non-zero
 * mask pixel in EVERY view.  Bit-exact against the reference's CPU path. */
int aninerf_inside_all_views(const float *wpts, int64_t n, const aninerf_silhouettes *sil_host, uint8_t *inside,
                             void *stream);

/* ------------------------------------------------------------------------------------------
 * Stage 5: alpha compositing, raw2outputs, lib/networks/renderer/nerf_net_utils.py:6-36
 * ---------------------------------------------------------------------------------------- */
/* raw: (n_rays,S,4) = (r,g,b,alpha) already activated; z_vals: (n_rays,S).
 * rgb_map (n_rays,3), acc_map, depth_map, disp_map (n_rays,), weights (n_rays,S): any may be NULL. */
int aninerf_composite(const float *raw, const float *z_vals, int64_t n_rays, int32_t n_samples,
                      int32_t white_bkgd, float *rgb_map, float *acc_map, float *depth_map,
                      float *disp_map, float *weights, void *stream);

/* ------------------------------------------------------------------------------------------
 * Stages 3+4: the two MLPs on tcgen05 tensor cores
 * ---------------------------------------------------------------------------------------- */
typedef struct aninerf_net aninerf_net;   /* opaque: packed weights of one Network on one device */

enum { ANINERF_FIELD_BW = 0, ANINERF_FIELD_NOVEL_BW = 1, ANINERF_FIELD_NERF = 2, ANINERF_N_FIELDS = 3 };

int aninerf_net_create(aninerf_net **out);
int aninerf_net_destroy(aninerf_net *net);

/* One dense layer as the kernel consumes it (after the host-side folding described in DESIGN.md):
 * W (n_out, k_in) row-major fp32, bias_table (n_tables, n_out) fp32 -- one bias row per latent
 * index (the per-frame latent code is constant over a frame, so its contribution
 * W[:, latent cols] @ latent is folded into the bias).  k_in is given WITHOUT padding.
 * These are HOST pointers: loading is a once-per-checkpoint operation, the library packs the
 * matrices into its tensor-core operand images on the host and uploads them. */
typedef struct {
  const float *W;
  const float *bias_table;
  int32_t n_out, k_in, n_tables, relu;
} aninerf_layer;

/* Upload one field.  Layer lists (k_in x n_out):
 *  BW / NOVEL_BW (calculate_neural_blend_weights, tpose_nerf_network.py:55-77 / :304-315):
 *     63x256, 4 x 256x256, (63+256)x256 [skip: PE first, then hidden], 2 x 256x256, 256x24
 *  NERF (TPoseHuman.calculate_alpha_rgb, tpose_nerf_network.py:252-275):
 *     63x256, 4 x 256x256, (63+256)x256, 2 x 256x256, (256+27)x128 [feature_fc o latent_fc o view_fc
 *     folded], then heads passed separately: alpha_fc (1,256)+(1,), rgb_fc (3,128)+(3,).
 * All pointers here are HOST fp32.  The call synchronises `stream` (it is not on the hot path). */
int aninerf_net_load_field(aninerf_net *net, int32_t field, const aninerf_layer *layers_host, int32_t n_layers,
                           const float *alpha_w_host, const float *alpha_b_host, const float *rgb_w_host,
                           const float *rgb_b_host, void *stream);

/* Blend-weight field forward (+ optional fused inverse LBS epilogue).
 *  pts (n,3): query points (pose space for pbw / canonical space for tbw)
 *  smpl_bw (n,24): initial SMPL weights (first 24 channels of aninerf_sample_blend_weights' rows,
 *     repacked to 24-float rows); the fused render path samples them in-kernel instead
 *  n_dev: optional device int32 holding the live row count (<= n) so the host never syncs
 *  bw_out (n,24) or NULL; A (24,4,4) + tpts_out (n,3) or NULL: inverse-LBS epilogue
 *     (Network.pose_points_to_tpose_points, tpose_nerf_network.py:79-100).
 * precision: 3 = bf16x3 split products (fp32-equivalent, the 1e-5 gate), 1 = single bf16 pass. */
int aninerf_bw_forward(aninerf_net *net, int32_t field, int32_t latent_index, const float *pts,
                       const float *smpl_bw, int64_t n, const int32_t *n_dev, const float *A, float *bw_out,
                       float *tpts_out, int32_t precision, void *stream);

/* Canonical NeRF field forward: raw density + raw colour (pre-activation), optional fused tail of
 * Network.forward (tpose_nerf_network.py:186-212): tbounds test, sigmoid, 1-exp(-relu(sigma)*dist),
 * scatter to the dense raw buffer through `index`.
 *  pts (n,3) canonical points; viewdir (n,3) world-space unit directions
 *  sigma_out (n,), rgb_out (n,3): pre-activation outputs, may be NULL
 *  tail: dists (n,), tbounds (2,3), index (n,) int32, raw_out (n_total,4) pre-zeroed; sigma_masked_out
 *  (n,) optional (sigma after the tbounds masking, the `alpha` the reference thresholds at :192). */
int aninerf_nerf_forward(aninerf_net *net, int32_t latent_index, const float *pts, const float *viewdir,
                         int64_t n, const int32_t *n_dev, float *sigma_out, float *rgb_out,
                         const float *dists, const float *tbounds, const int32_t *index, float *raw_out,
                         float *sigma_masked_out, int32_t precision, void *stream);

/* ------------------------------------------------------------------------------------------
 * The fused path: Renderer.render(batch), tpose_renderer.py:159-186 over
 * Network.forward, tpose_nerf_network.py:139-215
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  const float *A;        /* (24,4,4) */
  const float *R;        /* (3,3) */
  const float *Th;       /* (3,) */
  const float *pbw;      /* (X,Y,Z,25) posed-space blend-weight volume */
  const float *tbw;      /* (X',Y',Z',25) canonical volume */
  const float *pbounds;  /* (2,3) */
  const float *tbounds;  /* (2,3) */
  int32_t pbw_dims[3];
  int32_t tbw_dims[3];
  int32_t latent_index;     /* batch['latent_index'] */
  int32_t bw_latent_index;  /* batch['bw_latent_index'] (novel pose) */
  /* optional: the same two indices as DEVICE int64 scalars (how the reference batch holds them after
   * .cuda()); when non-NULL they are read by the kernels and the host values above are ignored, so the
   * host never waits on a device->host copy per frame */
  const int64_t *latent_index_dev;
  const int64_t *bw_latent_index_dev;
} aninerf_frame;

typedef struct {
  int32_t n_samples;      /* cfg.N_samples (<= 64, multiple of 32) */
  int32_t chunk_rays;     /* 2048 */
  float norm_th;          /* cfg.norm_th */
  int32_t white_bkgd;     /* cfg.white_bkgd */
  int32_t novel_pose;     /* cfg.test_novel_pose: use the NOVEL_BW field with bw_latent_index */
  int32_t want_bw;        /* also evaluate the canonical tbw pass and emit pbw/tbw/sigma rows */
  int32_t bw_precision;   /* 3 (default) or 1 */
  int32_t nerf_precision; /* 1 (default) or 3 */
} aninerf_render_params;

typedef struct {
  float *rgb_map;    /* (n_rays,3) */
  float *acc_map;    /* (n_rays,) */
  float *depth_map;  /* (n_rays,) */
  float *raw;        /* (n_rays*S,4) dense, zero off the active set: required with want_bw (a training-contract output);
                        NULL in render-only mode: the active rows stay compact in the workspace and are composited
                        through the mask words -- same maps bit for bit, no 16 B/sample buffer, no memset */
  /* want_bw outputs, each with room for n_rays*S rows (only the first *n_active are written) */
  float *pbw_all;      /* (n_active,24) */
  float *tbw_all;      /* (n_active,24) */
  float *sigma_masked; /* (n_active,) */
  int32_t *active_index; /* (n_active,) flat sample index of each active row; may be NULL */
  int32_t *n_active;   /* device int32: number of active samples (n'), REQUIRED */
  int32_t *chunk_offsets; /* (n_chunks+1,) device int32: start row of every 2048-ray chunk in the
                             compacted order; may be NULL */
} aninerf_render_outputs;

int64_t aninerf_render_workspace_bytes(int64_t n_rays, int32_t n_samples, int32_t want_bw, int64_t pbw_voxels,
                                       int64_t tbw_voxels);

/* ray_o, ray_d (n_rays,3); near, far (n_rays,); t_vals (S,); t_rand (n_rays,S) or NULL. */
int aninerf_render_rays(aninerf_net *net, const aninerf_frame *frame_host, const aninerf_render_params *params_host,
                        const float *ray_o, const float *ray_d, const float *near, const float *far,
                        const float *t_vals, const float *t_rand, int64_t n_rays,
                        const aninerf_render_outputs *out_host, void *workspace, int64_t workspace_bytes,
                        void *stream);

/* tpose_renderer_mmsk.Renderer.render (tpose_renderer_mmsk.py:99-166): the same frame render with the
 * samples culled by the training-view silhouettes BEFORE the network (so the per-chunk argmin forcing
 * of Network.forward runs over the survivors only, and a chunk without survivors evaluates nothing).
 * want_bw must be 0 (the reference returns maps only). */
int aninerf_render_rays_culled(aninerf_net *net, const aninerf_frame *frame_host, const aninerf_render_params *params_host,
                               const aninerf_silhouettes *sil_host, const float *ray_o, const float *ray_d, const float *near,
                               const float *far, const float *t_vals, const float *t_rand, int64_t n_rays,
                               const aninerf_render_outputs *out_host, void *workspace, int64_t workspace_bytes,
                               void *stream);

/* Ray-tiled multi-GPU render with the image gather FUSED into the compositing kernel: rank `rank` of `world` renders the
 * 2048-ray chunks  rank, rank+world, ...  of the frame (its rays arrive concatenated in that order) and the compositing
 * kernel stores every ray's (rgb, acc, depth) row -- 20 B -- straight into the frame-ordered (n_rays_frame, 5) image buffer
 * of EVERY rank: maps[k] is rank k's buffer, peer-mapped over NVLink/NVSwitch (e.g. torch symmetric memory).  The caller
 * follows the call with one cross-rank barrier; there is no all_gather and no reorder kernel.  sil_host / peers_host may be
 * NULL (then this is aninerf_render_rays[_culled]). */
typedef struct {
  void *maps[ANINERF_MAX_PEERS];
  int32_t world, rank;
} aninerf_peer_gather;
int aninerf_render_rays_tiled(aninerf_net *net, const aninerf_frame *frame_host, const aninerf_render_params *params_host,
                              const aninerf_silhouettes *sil_host, const aninerf_peer_gather *peers_host, const float *ray_o,
                              const float *ray_d, const float *near, const float *far, const float *t_vals, const float *t_rand,
                              int64_t n_rays, const aninerf_render_outputs *out_host, void *workspace, int64_t workspace_bytes,
                              void *stream);

/* Network.calculate_alpha (= get_alpha), tpose_nerf_network.py:105-137: density-only query of
 * world points (norm_th hard-coded 0.1 by the reference -- passed in), processed in chunks of
 * chunk_pts points (131072 in aninerf_mesh_renderer.py:35).  sigma_out (n,), zero where masked. */
int aninerf_query_alpha(aninerf_net *net, const aninerf_frame *frame_host, const float *wpts, int64_t n,
                        int64_t chunk_pts, float norm_th, int32_t novel_pose, int32_t bw_precision,
                        float *sigma_out, int32_t *n_active, void *workspace, int64_t workspace_bytes,
                        void *stream);
int64_t aninerf_query_workspace_bytes(int64_t n, int64_t pbw_voxels);

/* ------------------------------------------------------------------------------------------
 * Training step: tpose_trainer.NetworkWrapper.forward (lib/train/trainers/tpose_trainer.py:21-73) and the
 * backward pass torch.autograd derives for it in the reference (Trainer.train, trainer.py:62-66).
 * The Python mirror (animatable_nerf_b200/tpose_trainer.py) chains these entries layer by layer and
 * keeps every activation in fp32.
 * ---------------------------------------------------------------------------------------- */

/* The front end of the fused path alone: stratified samples (+jitter) -> pose space -> pnorm < norm_th mask ->
 * per-2048-ray-chunk argmin forcing -> stable compaction (tpose_renderer.py:14-69, tpose_nerf_network.py:143-157).
 * index (n',) flat sample index; ppts, viewdir (n',3); dists (n',); z_vals (n_rays,S) or NULL; buffers sized n_rays*S. */
int64_t aninerf_front_end_workspace_bytes(int64_t n_rays, int32_t n_samples, int64_t pbw_voxels);
int aninerf_front_end(const aninerf_frame *frame_host, const aninerf_render_params *params_host, const float *ray_o, const float *ray_d,
                      const float *near, const float *far, const float *t_vals, const float *t_rand, int64_t n_rays, int32_t *index,
                      float *ppts, float *viewdir, float *dists, float *z_vals, int32_t *n_active, int32_t *chunk_offsets,
                      void *workspace, int64_t workspace_bytes, void *stream);

/* One dense product on the tcgen05 tensor cores, fp32 in / fp32 out, every product as bf16x3 (fp32-equivalent):
 *     C[M,N] = epilogue( sum_s A_s[M,K_s] * B_s[N,K_s]^T )
 * It replaces F.conv1d (kernel 1) forward, its data gradient and its weight gradient (tpose_nerf_network.py:68-72,
 * 256-274 and their autograd).  Operand element (row r, k) of segment s is ptr[r*row_stride + k*k_stride], so X, X^T, W, W^T
 * and column ranges of W are read in place; two segments give the skip / view layers' concatenated inputs.
 * epilogue: + bias[n]; + C (accumulate); ReLU; * (relu_mask[m,n] > 0).  split_k > 1 (one segment, plain epilogue):
 * the K range is split over CTAs into `workspace` and reduced in fixed order (deterministic weight gradients). */
typedef struct {
  const float *A;
  int64_t a_row_stride, a_k_stride;
  const float *B;
  int64_t b_row_stride, b_k_stride;
  int32_t K;
} aninerf_gemm_seg;
typedef struct {
  aninerf_gemm_seg seg[2];
  int32_t n_seg, M, N;
  float *C;
  int64_t ldc;
  const float *bias;
  const float *relu_mask;
  int64_t ld_mask;
  int32_t relu, accumulate, split_k;
} aninerf_gemm;
int64_t aninerf_gemm_workspace_bytes(const aninerf_gemm *g_host);
int aninerf_gemm_x3(const aninerf_gemm *g_host, void *workspace, int64_t workspace_bytes, void *stream);
/* out[n] (+)= sum_m X[m*ld + n] (bias gradients), fixed summation order; workspace >= ceil(M/256)*N floats. */
int aninerf_colsum(const float *X, int64_t ld, int64_t M, int32_t N, float *out, int32_t accumulate, void *workspace,
                   int64_t workspace_bytes, void *stream);

/* Positional encoding (lib/networks/embedder.py:11-36) as a stand-alone op and its gradient w.r.t. x. out/d_pe rows have
 * leading dimension ld >= 3 + 6*n_freq. */
int aninerf_pe_forward(const float *x, int64_t n, int32_t n_freq, float *out, int64_t ld, void *stream);
int aninerf_pe_backward(const float *x, const float *d_pe, int64_t ld, int64_t n, int32_t n_freq, float *d_x, int32_t accumulate,
                        void *stream);
/* bw = softmax(log(init + 1e-9) + delta) over the 24 bones (tpose_nerf_network.py:74-76); init rows have leading dimension
 * ld_init (25 for the sampled volume rows).  Backward: d_delta (n,24) and, when non-NULL, d_init (n,24). */
int aninerf_bw_softmax_forward(const float *init, int64_t ld_init, const float *delta, int64_t n, float *bw, void *stream);
int aninerf_bw_softmax_backward(const float *init, int64_t ld_init, const float *bw, const float *d_bw, int64_t n, float *d_delta,
                                float *d_init, void *stream);
/* Gradient of pose_points_to_tpose_points (blend_utils.py:41-59) w.r.t. the blend weights. */
int aninerf_inverse_lbs_backward(const float *bw, const float *A, const float *tpts, const float *d_tpts, int64_t n, float *d_bw,
                                 int32_t accumulate, void *stream);
/* Gradient of pts_sample_blend_weights (blend_utils.py:119-149) w.r.t. the query points, first 24 channels
 * (grid_sampler_3d_backward, align_corners, border padding).  d_out (n,24). */
int aninerf_sample_blend_weights_backward(const float *pts, int64_t n, const float *vol, const int32_t dims_host[3], const float *bounds,
                                          const float *d_out, float *d_pts, int32_t accumulate, void *stream);
/* Tail of Network.forward (tpose_nerf_network.py:186-212) and its backward.  raw_full (n_total,4) pre-zeroed. */
int aninerf_nerf_tail_forward(const float *sigma, const float *rgb, const float *tpts, const float *tbounds, const float *dists,
                              const int32_t *index, int64_t n, float *raw_full, float *sigma_masked, void *stream);
int aninerf_nerf_tail_backward(const float *d_raw_full, const float *raw_full, const int32_t *index, const float *sigma_masked,
                               const float *tpts, const float *tbounds, const float *dists, int64_t n, float *d_sigma, float *d_rgb,
                               void *stream);
/* Stage-2 trainer (lib/train/trainers/aninerf_animation_trainer.py:73-82): out = sigma where the canonical point lies strictly
 * inside tbounds and (pnorm given) pnorm[i*ld] < norm_th, else 0. */
int aninerf_mask_sigma(const float *sigma, const float *tpts, const float *tbounds, const float *pnorm, int64_t ld, float norm_th,
                       int64_t n, float *out, void *stream);
/* Backward of raw2outputs (nerf_net_utils.py:6-36) for a gradient arriving on rgb_map. d_raw (n_rays,S,4). */
int aninerf_composite_backward(const float *raw, const float *d_rgb_map, int64_t n_rays, int32_t n_samples, int32_t white_bkgd,
                               float *d_raw, void *stream);
/* img_loss = mean((rgb_map[mask] - rgb[mask])^2) and its gradient (tpose_trainer.py:60-63). loss: device float. */
int aninerf_img_loss(const float *rgb_map, const float *rgb_gt, const uint8_t *mask, int64_t n_rays, float *loss, float *d_rgb_map,
                     void *stream);
/* alpha_ind of tpose_nerf_network.py:192-194 for every chunk: sigma_masked > train_th plus the chunk's first arg-max row. */
int aninerf_select_rows(const float *sigma_masked, const int32_t *chunk_offsets, int32_t n_chunks, float train_th, uint8_t *sel,
                        int32_t *n_sel, void *stream);
/* The boolean-mask gathers `pbw[alpha_ind]`, `tbw[alpha_ind]` of tpose_nerf_network.py:195-196 for every chunk at once: the rows
 * of src_a / src_b (24 floats each, 16-byte aligned) whose `sel` byte is set, in ascending row order, packed into dst_a / dst_b
 * (src_b / dst_b may be NULL).  sel_offsets: (n_chunks + 1) int32 scratch; on return [c] = first output row of chunk c and
 * [n_chunks] = total selected rows (stays on the device).  dst must have room for chunk_offsets[n_chunks] rows. */
int aninerf_gather_selected_rows(const uint8_t *sel, const int32_t *chunk_offsets, int32_t n_chunks, const float *src_a,
                                 const float *src_b, float *dst_a, float *dst_b, int32_t *sel_offsets, void *stream);
/* bw_loss = smooth_l1(pbw[sel], tbw[sel]) (tpose_trainer.py:48-51) and its gradients (n,24), zero on unselected rows. */
int aninerf_bw_loss(const float *pbw, const float *tbw, const uint8_t *sel, const int32_t *n_sel, int64_t n, float *loss, float *d_pbw,
                    float *d_tbw, void *stream);

/* K-nearest-vertex blend weights of the extended (`aligned_*`, `anisdf_*`) networks: `sample_blend_closest_points`,
 * lib/utils/sample_utils.py:323-349 over pytorch3d.ops.knn_points.  pts (n,3), verts (n_verts,3), values (n_verts,24) 16-byte
 * aligned; K in {1, 5, 8} (the reference uses 5), eps = 1e-8.  bw_out (n,24) = sum_k w_k values[idx_k] with
 * w_k = (1/(d_k + eps)) / sum_j 1/(d_j + eps), d = Euclidean distance to the K nearest vertices; dist_out (n) = sum_k d_k w_k
 * (may be NULL). */
int aninerf_knn_blend_weights(const float *pts, int64_t n, const float *verts, int32_t n_verts, const float *values, int32_t K, float eps,
                              float *bw_out, float *dist_out, void *stream);

/* Mesh extraction from the density cube: `mcubes.marching_cubes(cube, cfg.mesh_th)` of
 * lib/networks/renderer/aninerf_mesh_renderer.py:40 (PyMCubes 0.1.0: Lorensen-Cline marching cubes, corner bit set when
 * value <= iso, vertices linearly interpolated in float64 index coordinates).  cube: device (X,Y,Z) float32, x-major.
 * verts: device (cap_verts,3) float64, ordered by owning grid point then axis; tris: device (cap_tris,3) int32 indices into verts,
 * ordered by cell then case-table order, wound towards the <= iso side.  counts: device int32[2] = {vertices, triangles} the
 * cube yields -- when a count exceeds its capacity the surplus was dropped: read the counts, enlarge, call again. */
int64_t aninerf_marching_cubes_workspace_bytes(int32_t X, int32_t Y, int32_t Z);
int aninerf_marching_cubes(const float *cube, int32_t X, int32_t Y, int32_t Z, double iso, double *verts, int64_t cap_verts,
                           int32_t *tris, int64_t cap_tris, int32_t *counts, void *workspace, int64_t workspace_bytes, void *stream);

/* Optional per-stage device timing of aninerf_render_rays (CUDA events on the launching stream).
 * Stage order: split volumes, clear raw, mask+scan+compact front end, (unused), (unused), blend-weight
 * field at posed points (+LBS), blend-weight field at canonical points, NeRF field (+tail), compositing.
 * aninerf_profile_read synchronises the device, returns accumulated milliseconds and call counts. */
#define ANINERF_N_STAGES 9
int aninerf_profile_enable(int32_t on);
int aninerf_profile_read(double *ms_out_host, int64_t *calls_out_host, int32_t reset);

/* Bring-up aid: when set to a device buffer of >= 128 uint64, the MLP kernels write a clock64
 * timeline of block 0's first tile into it (slot map in csrc/mlp_tcgen05.cu).  NULL disables. */
int aninerf_debug_set_trace(unsigned long long *device_buf);

/* Counts launches of this library's kernels since process start (bench.py's gpu_launches). */
int64_t aninerf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* ANINERF_B200_H */
