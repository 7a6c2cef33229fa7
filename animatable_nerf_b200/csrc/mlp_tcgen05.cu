// The two dense fields of the path -- the neural blend-weight MLP (tpose_nerf_network.py:55-77,
// :304-315) and the canonical NeRF MLP (tpose_nerf_network.py:252-275) -- as ONE persistent,
// warp-specialised tcgen05 kernel family for sm_100a.
//
// Work unit: a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2).  Each CTA owns 128 samples per tile
// slot: their bf16 A operand lives in its shared memory (K-major, no-swizzle core-matrix layout
// [K/8][128 rows][8]) and their fp32 accumulator in its TMEM (128 lanes x 256 columns per slot).
// Weights stream from L2 through a ring of 16 KB stages filled by bulk TMA copies (cp.async.bulk)
// of pre-packed operand images; each CTA of the pair loads HALF of every chunk (the M=256 MMA reads
// B from both CTAs), which halves the L2 -> SM weight traffic per sample.  One elected thread of the
// leader CTA issues tcgen05.mma (M=256, N<=256, K=16); eight epilogue warps per CTA (two threads per
// row, each owning half of the columns) read the accumulator back with tcgen05.ld, apply bias+ReLU,
// re-quantise to bf16 (hi, and lo for the split-precision mode) and write the next layer's A operand
// in place.  Positional encoding is generated in-kernel straight into the A operand; the last
// epilogue is the field's head (softmax + inverse LBS, or alpha/rgb activation + tbounds masking +
// scatter).
//
// NT = 2 (single-pass precision): each CTA holds TWO tile slots, each with its own eight epilogue warps, and
// ping-pongs them -- while the tensor core runs layer l of slot 1 the warps of slot 0 drain its layer l -- so
// MMA and epilogue overlap.  NT = 1 for the split-precision mode (its A operand, hi+lo, fills shared memory):
// there the accumulator is double-buffered, the epilogue publishes the next A operand in quarters and every
// 256-wide layer runs as two N = 128 halves, so the tensor pipe never waits for an epilogue (see QP below).
//
// Precision modes: NPASS=1 single bf16 product; NPASS=3 "bf16x3": x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// with fp32 accumulation (fp32-equivalent; the blend-weight field needs it for the 1e-5 gate).
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tcgen05.cuh"

namespace aninerf {

// ------------------------------------------------------------------------------------------------
// layout constants
// ------------------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int CHUNK_BYTES = TILE_M * 16;     // one 8-wide K chunk of the A tile: 128 rows x 16 B
constexpr int PE_CHUNK0 = 0;                 // A chunks 0..7  : PE(xyz) 63 + pad; reused for PE(viewdir) once layer 5 has run
constexpr int HID_CHUNK0 = 8;                // A chunks 8..39 : hidden 256
constexpr int A_CHUNKS = 40;
constexpr int A_BYTES = A_CHUNKS * CHUNK_BYTES;   // 80 KB per tile slot (per hi / lo plane)
constexpr int STAGE_BYTES = 16384;
constexpr int MAX_LAYERS = 9;
constexpr int MAX_STEPS = 128;
constexpr int BIAS_FLOATS = MAX_LAYERS * 256;
constexpr int SMEM_LIMIT = 232448;           // 227 KB
constexpr int VIEW_LAYER_WRITE = 6;          // PE(viewdir) is written during this layer's epilogue (layer 5 was the last PE(xyz) reader)

struct Step {          // one weight-ring stage worth of MMAs
  uint32_t w_off;      // byte offset of this step's operand image in the packed buffer: [CTA0: hi, lo][CTA1: hi, lo]
  uint32_t bytes;      // bytes of the whole step (all CTAs of the pair)
  uint16_t a_chunk;    // first A chunk consumed
  uint16_t n_k16;      // K=16 MMAs (per pass) in this step
  uint16_t layer;
  uint16_t flags;      // 1: first step of its (sub-)layer: fresh accumulator; 2: last step of its (sub-)layer: commit its acc barrier;
                       // 4: (N-split, half B) the hidden quarters 0,1 of the A operand have been read for the last time: commit `bread`
};

struct LayerDev {
  int32_t n_pad;       // MMA N (multiple of 32)
  int32_t n_out;       // real outputs
  int32_t relu;
  int32_t bias_off;    // float offset of this layer's bias table inside `bias`
  int32_t n_tables;
  int32_t step0;       // first step of the layer
  int32_t n_steps;
};

struct FieldDev {
  const uint8_t *image;    // packed weights for this precision
  const Step *steps;
  int32_t n_steps;
  int32_t n_layers;
  LayerDev layers[MAX_LAYERS];
  const float *bias;       // all bias tables
  const float *head;       // NeRF: alpha_w[256], alpha_b, rgb_w[3][128], rgb_b[3]
};

struct MlpArgs {
  FieldDev f;
  int32_t latent_index;
  const int64_t *latent_dev;   // optional device int64 (batch['latent_index'] as the reference holds it): index = *latent_dev + latent_index
  const float *pts;        // (n,3)
  const float *viewdir;    // (n,3) NeRF
  int64_t n;
  const int32_t *n_dev;
  // BW head
  const float *smpl_bw;    // (n,24) or null
  const float *vol_w24;    // (X,Y,Z,24) or null
  const float *grid_bounds;   // device (2,3)
  int32_t grid_dim[3];
  const float *A;          // (24,4,4) or null
  float *bw_out;           // (n,24) or null
  float *tpts_out;         // (n,3) or null
  // NeRF head
  float *sigma_out, *rgb_out;
  const float *dists;
  const float *tbounds;
  const int32_t *index;
  float *raw_out;
  float *sigma_masked_out;
  unsigned long long *trace;   // bring-up: clock64 timeline of one unit tile of block 0 (null = off)
  int32_t trace_iter;          // which of block 0's tiles is traced (0 = first); slots 160+i: start of its i-th tile
};

// trace slots: 0 tile start, 1 PE done; per (layer l, slot t) base 8 + 16*l + 8*t: +0 rows wait begin,
// +1 rows woke, +2 rows epilogue done (arrived); +4 MMA waits a_ready, +5 MMA woke, +6 MMA issued the layer
#define ANI_TRACE(slot)                                                              \
  do {                                                                               \
    if (tracing) args.trace[(slot)] = (unsigned long long)clock64();                 \
  } while (0)

// write 8 consecutive K elements of `row` (one 16-byte core-matrix row) into A chunk `chunk`;
// RELU (single-pass mode only): x holds pre-activation values, the ReLU is applied by the conversion
template <int NPASS, bool RELU = false>
__device__ __forceinline__ void store_chunk(uint8_t *a_hi, uint8_t *a_lo, int chunk, int row, const float (&x)[8]) {
  static_assert(!(RELU && NPASS == 3), "the split-precision residual needs the activated fp32 value");
  uint4 h;
  if (RELU) {
    h.x = pack_bf16_relu(x[0], x[1]);
    h.y = pack_bf16_relu(x[2], x[3]);
    h.z = pack_bf16_relu(x[4], x[5]);
    h.w = pack_bf16_relu(x[6], x[7]);
  } else {
    h.x = pack_bf16(x[0], x[1]);
    h.y = pack_bf16(x[2], x[3]);
    h.z = pack_bf16(x[4], x[5]);
    h.w = pack_bf16(x[6], x[7]);
  }
  *reinterpret_cast<uint4 *>(a_hi + chunk * CHUNK_BYTES + row * 16) = h;
  if (NPASS == 3) {
    uint4 l;
    l.x = pack_bf16_residual(x[0], x[1], h.x);
    l.y = pack_bf16_residual(x[2], x[3], h.y);
    l.z = pack_bf16_residual(x[4], x[5], h.z);
    l.w = pack_bf16_residual(x[6], x[7], h.w);
    *reinterpret_cast<uint4 *>(a_lo + chunk * CHUNK_BYTES + row * 16) = l;
  }
}

// NeRF positional encoding of a 3-vector (embedder.py:11-36): [x, sin(2^0 x), cos(2^0 x), ...] in 3-wide
// blocks; writes the 8-wide K chunks [CB, CE) of the encoding into A chunks chunk0+CB ... (each of the
// two threads that share a row takes half of the chunks)
template <int NPASS, int L, int CB, int CE>
__device__ __forceinline__ void write_pe(uint8_t *a_hi, uint8_t *a_lo, int chunk0, int row, float px, float py, float pz) {
  constexpr int NV = 3 + 6 * L;                                   // 63 or 27
  constexpr int J0 = CB * 8, J1 = CE * 8;                         // value range [J0, J1)
  constexpr int F0 = J0 <= 3 ? 0 : (J0 - 3) / 6;
  constexpr int F1 = ((J1 < NV ? J1 : NV) - 1 - 3) / 6;           // last frequency touched
  const float p[3] = {px, py, pz};
  float sn[F1 - F0 + 1][3], cs[F1 - F0 + 1][3];
  // sin / cos of 2^f * x for power-of-two frequencies: x / 2pi once, in double-float (Cody-Waite split of 1/2pi);
  // scaling by 2^f and removing whole turns are then EXACT, and the remaining angle in [-pi, pi] goes to the
  // SFU (abs error ~6e-7, independent of the frequency; libm sincosf costs ~8x the instructions)
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float inv_hi = 0.15915494f, inv_lo = 6.4206382e-09f;
    const float t_hi = __fmul_rn(p[c], inv_hi);
    const float t_lo = __fmaf_rn(p[c], inv_lo, __fmaf_rn(p[c], inv_hi, -t_hi));
#pragma unroll
    for (int f = F0; f <= F1; ++f) {
      const float sc = (float)(1 << f);
      const float a = t_hi * sc;                       // exact
      const float turns = (a - rintf(a)) + t_lo * sc;  // fractional turns, |.| <= 0.5 (+ tiny)
      const float ang = turns * 6.2831855f;
      sn[f - F0][c] = __sinf(ang);
      cs[f - F0][c] = __cosf(ang);
    }
  }
#pragma unroll
  for (int ch = CB; ch < CE; ++ch) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int v = ch * 8 + j;
      if (v < 3) x[j] = p[v];
      else if (v >= NV) x[j] = 0.f;
      else {
        const int f = (v - 3) / 6, r = (v - 3) % 6;
        x[j] = r < 3 ? sn[f - F0][r] : cs[f - F0][r - 3];
      }
    }
    store_chunk<NPASS>(a_hi, a_lo, chunk0 + ch, row, x);
  }
}

// One trilinear corner for the SMPL-weight gather of the blend-weight head: same geometry as trilinear_corners() (align_corners,
// border clamp) with the normalisation folded into one multiply by (dim-1)/ext -- the weights feed a 1e-5-gated quantity, not a
// bit-exact mask, and the exact form's three IEEE divisions and rounding-order chain cost ~3k cycles of dependent latency per
// row on the two-warps-per-scheduler epilogue threads.  gs: lo[3], scale[3] in shared memory.  The voxel index is always valid
// (an out-of-range corner has weight exactly 0).  Corner k in ATen order: k&1 east, k>>1&1 south, k>>2 bottom.
__device__ __forceinline__ void fast_corner(const float *gs, const int32_t dim[3], float px, float py, float pz, int k, float &w, int &off) {
  const float p[3] = {px, py, pz};
  const int up[3] = {(k >> 2) & 1, (k >> 1) & 1, k & 1};   // axis 0 (X) <-> bottom, 1 (Y) <-> south, 2 (Z) <-> east
  float wa[3];
  int idx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lim = (float)(dim[a] - 1);
    const float u = fminf(lim, fmaxf((p[a] - gs[a]) * gs[3 + a], 0.0f));
    const float f = floorf(u);
    const float fr = u - f;
    idx[a] = up[a] ? min((int)f + 1, dim[a] - 1) : (int)f;
    wa[a] = up[a] ? fr : 1.0f - fr;
  }
  w = (wa[2] * wa[1]) * wa[0];
  off = (idx[0] * dim[1] + idx[1]) * dim[2] + idx[2];
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int ROW_WARPS = 8;                       // two threads per row: each takes half of the columns
constexpr int ROW_THREADS = ROW_WARPS * 32;        // 256
// Threads of a CTA: ROW_THREADS per tile slot (NT = 2: each slot has its OWN eight epilogue warps, so the two slots' epilogues run
// concurrently instead of queueing behind each other on the same threads), then one weight-producer warp and two warps for MMA
// issue (one thread per tile slot) / TMEM alloc / relays.
constexpr int n_threads(int nt) { return ROW_THREADS * nt + 96; }
constexpr int XCHG_BYTES = TILE_M * 4 * 4;         // per-row exchange between the two column halves
constexpr int PROD_LANES = 8;                      // producer lanes take turns issuing the bulk copies: one thread keeps only
                                                   // ~one copy in flight (20-28 B/cycle, tools/bench_stream.cu); several
                                                   // threads overlap theirs (4 threads: 85-110 B/cycle)

template <int NPASS, bool NERF, int NT>
struct Cfg {
  static constexpr int A_TOTAL = A_BYTES * NT * (NPASS == 3 ? 2 : 1);
  static constexpr int HEAD_BYTES = NERF ? 2576 : 1152;
  static constexpr int STEP_BYTES = MAX_STEPS * (int)sizeof(Step);
  static constexpr int FIXED = A_TOTAL + BIAS_FLOATS * 4 + HEAD_BYTES + XCHG_BYTES * NT + STEP_BYTES + 256;
  static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM = FIXED + STAGES * STAGE_BYTES;
  static_assert(STAGES >= 2, "weight ring needs at least two stages");
  static constexpr int TMEM_COLS = 512;                         // NT = 2: one accumulator per slot; NT = 1: two, alternating by layer
  static constexpr int N_AREADY = NT == 1 ? 4 : NT;             // NT = 1: one per 64-column quarter of the hidden layer (quarter pipelining)
  static constexpr int N_ACC = NT == 1 ? 3 : NT;                // NT = 1: acc of half A, acc of half B, `bread` (N-split)
  // offsets
  static constexpr int OFF_A_HI = 0;                            // slot t at t * A_BYTES
  static constexpr int OFF_A_LO = A_BYTES * NT;                 // only when NPASS == 3
  static constexpr int OFF_RING = A_TOTAL;
  static constexpr int OFF_BIAS = OFF_RING + STAGES * STAGE_BYTES;
  static constexpr int OFF_HEAD = OFF_BIAS + BIAS_FLOATS * 4;
  static constexpr int OFF_XCHG = OFF_HEAD + HEAD_BYTES;
  static constexpr int OFF_STEPS = OFF_XCHG + XCHG_BYTES * NT;
  static constexpr int OFF_BAR = OFF_STEPS + STEP_BYTES;
};

template <int NPASS, bool NERF, int PAIR, int NT>
__global__ void __launch_bounds__(n_threads(NT), 1) mlp_kernel(const __grid_constant__ MlpArgs args) {
  constexpr int N_THREADS = n_threads(NT);
  constexpr int RW = ROW_WARPS * NT;                 // row (epilogue) warps; warp RW: producer; RW+1, RW+2: MMA issue / relays
  using C = Cfg<NPASS, NERF, NT>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *ring = smem + C::OFF_RING;
  float *s_bias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
  float *s_head = reinterpret_cast<float *>(smem + C::OFF_HEAD);
  float *s_xchg = reinterpret_cast<float *>(smem + C::OFF_XCHG);
  Step *s_steps = reinterpret_cast<Step *>(smem + C::OFF_STEPS);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  // barrier slots: [0,S) full, [S,2S) empty, [2S,2S+NT) a_ready (leader), [2S+NT,2S+2NT) acc
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + C::STAGES);
  const uint32_t bar_a_ready = smem_u32(bars + 2 * C::STAGES);
  const uint32_t bar_acc = smem_u32(bars + 2 * C::STAGES + C::N_AREADY);
  // peer CTA only: its rows arrive here (cheap CTA-local arrives); one relay thread forwards each completed phase
  // to the leader's a_ready with a single cluster-scope arrive
  const uint32_t bar_a_local = smem_u32(bars + 2 * C::STAGES + C::N_AREADY + C::N_ACC);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * C::STAGES + 2 * C::N_AREADY + C::N_ACC);
  static_assert((2 * C::STAGES + 2 * C::N_AREADY + C::N_ACC + 1) * 8 <= 216, "barrier block overflow");
  // QP (NT == 1): the epilogue of layer l publishes the next layer's A operand quarter by quarter (a_ready[q]) and the
  // accumulator alternates between two TMEM buffers, so the MMAs of layer l+1 start after the first quarter is written
  // N-split (QP only): a 256-wide layer runs as two N=128 halves (A: output columns 0-127, then B), each with its own acc
  // barrier.  The epilogue of half A (= quarters 0,1 of the next layer's K) overlaps the MMAs of half B, and the next layer's
  // half A starts on quarters 0,1 while the epilogue of half B still produces quarters 2,3: the tensor pipe never waits for an
  // epilogue.  The A operand is updated IN PLACE, so the epilogue of half A may only overwrite quarters 0,1 once half B's MMAs
  // have read them: half B commits `bread` (bar_acc + 16) after its K steps over those quarters.
  constexpr bool QP = NT == 1;
  // XT (blend-weight field, QP): cross-tile prefetch.  The next tile's input encoding is written into the PE chunks during the
  // idle window of the LAST layer (the PE chunks are dead since layer 5), so the MMA issuer rolls from layer 8 straight into the
  // next tile's layer 0 while the epilogue warps are still busy with this tile's head; the accumulator buffer alternates with
  // (layer + tile) so that layer 0 never lands on the buffer the head is reading.
  constexpr bool XT = QP && !NERF;
  // last weight-stream position consumed from each ring stage: with one MMA thread per slot a thread only waits
  // on the `full` phases of its OWN steps, and a parity wait is only sound once the previous phase is known complete
  volatile uint32_t *s_last = reinterpret_cast<volatile uint32_t *>(smem + C::OFF_BAR + 216);
  float *s_grid = reinterpret_cast<float *>(smem + C::OFF_BAR + 232);   // lo[3], (dim-1)/ext [3] of the SMPL-weight volume
  static_assert(C::STAGES * 4 <= 16, "s_last overlaps s_grid");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_units = (int)gridDim.x / PAIR;            // clusters walking the unit-tile list
  const int unit = (int)blockIdx.x / PAIR;
  const FieldDev &F = args.f;
  const int64_t n_valid = args.n_dev ? (int64_t)min((int64_t)*args.n_dev, args.n) : args.n;
  const int64_t n_tiles = (n_valid + TILE_M - 1) / TILE_M;
  constexpr int UT = PAIR * NT;                         // 128-row tiles per unit tile
  const int64_t n_utiles = (n_tiles + UT - 1) / UT;

  // ---- one-time setup --------------------------------------------------------------------
  const int latent = args.latent_index + (args.latent_dev ? (int)__ldg(args.latent_dev) : 0);
  for (int i = threadIdx.x; i < BIAS_FLOATS; i += N_THREADS) {
    int l = i >> 8, j = i & 255;
    float b = 0.f;
    if (l < F.n_layers && j < F.layers[l].n_out) {
      int t = min(max(latent, 0), F.layers[l].n_tables - 1);
      b = F.bias[F.layers[l].bias_off + t * F.layers[l].n_out + j];
    }
    s_bias[i] = b;
  }
  for (int i = threadIdx.x; i < F.n_steps; i += N_THREADS) s_steps[i] = F.steps[i];
  if (NERF) {
    for (int i = threadIdx.x; i < 644; i += N_THREADS) s_head[i] = F.head[i];
  } else if (args.A) {
    for (int i = threadIdx.x; i < 288; i += N_THREADS) s_head[i] = args.A[(i / 12) * 16 + (i % 12)];
  }
  if (!NERF && threadIdx.x < 3 && args.grid_bounds) {
    const float lo = args.grid_bounds[threadIdx.x];
    s_grid[threadIdx.x] = lo;
    s_grid[3 + threadIdx.x] = (float)(args.grid_dim[threadIdx.x] - 1) / (args.grid_bounds[3 + threadIdx.x] - lo);
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(bar_full + 8 * s, (leader && PAIR == 2) ? 2 : 1);   // own bytes landed (+ the peer's relay on the leader)
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int t = 0; t < C::N_AREADY; ++t) {
      mbar_init(bar_a_ready + 8 * t, ROW_THREADS + (PAIR == 2 ? 1 : 0));   // the leader's rows + the peer's relay
      mbar_init(bar_a_local + 8 * t, ROW_THREADS);
    }
    for (int t = 0; t < C::N_ACC; ++t) mbar_init(bar_acc + 8 * t, 1);
    for (int s = 0; s < C::STAGES; ++s) s_last[s] = 0xffffffffu;
    fence_barrier_init();
  }
  if (warp == RW + 1) tmem_alloc<PAIR>(smem_u32(tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();        // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == RW) {
    // ===== weight producer: this CTA's half of every operand image chunk ======================
    if (lane < PROD_LANES) {
      uint32_t stage = 0, phase = 0, turn = 0;
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        for (int l = 0; l < F.n_layers; ++l) {
          for (int t = 0; t < NT; ++t) {
            for (int s = F.layers[l].step0; s < F.layers[l].step0 + F.layers[l].n_steps; ++s) {
              mbar_wait(bar_empty + 8 * stage, phase ^ 1, 1);               // all producer lanes stay in step (no divergent spin)
              if (turn == (uint32_t)lane) {
                const uint32_t bytes = s_steps[s].bytes / PAIR;             // this CTA's half of the chunk
                mbar_expect_tx(bar_full + 8 * stage, bytes);
                bulk_g2s(smem_u32(ring + stage * STAGE_BYTES), F.image + s_steps[s].w_off + cta_rank * bytes, bytes, bar_full + 8 * stage);
              }
              __syncwarp((1u << PROD_LANES) - 1u);
              turn = (turn + 1) % PROD_LANES;
              if (++stage == C::STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp > RW) {
    const int my_slot = warp - (RW + 1);
    if (leader && my_slot < NT) {
      // ===== MMA issuers (leader CTA): one thread per tile slot, so that neither slot's issue stream
      // waits behind the other's epilogue and the per-step barrier/commit overhead is split in two ====
      uint32_t g = 0, a_phase = 0;                                  // g: position in the shared weight stream
      const uint32_t a_lbo = (uint32_t)CHUNK_BYTES, a_sbo = 128u;   // K-direction / 8-row-group strides
      const int t = my_slot;
      const uint32_t a_hi = smem_u32(smem + C::OFF_A_HI + t * A_BYTES), a_lo = smem_u32(smem + C::OFF_A_LO + t * A_BYTES);
      uint32_t tile_it = 0;
      for (int64_t ut = unit; ut < n_utiles; ut += n_units, ++tile_it) {
        const bool tracing = args.trace && blockIdx.x == 0 && ut == (int64_t)args.trace_iter * n_units && lane == 0;
        for (int l = 0; l < F.n_layers; ++l) {
          const int n_pad = F.layers[l].n_pad;
          const int n_sub = (QP && n_pad == 256) ? 128 : n_pad;         // N of one MMA (N-split halves of a 256-wide layer)
          const int s0 = F.layers[l].step0, n_layer_steps = F.layers[l].n_steps;
          const uint32_t idesc = instr_desc(n_sub, TILE_M * PAIR);
          const uint32_t b_k_stride = (uint32_t)(n_sub / PAIR) * 16u;   // bytes between K core matrices (this CTA's rows)
          const uint32_t b_lbo = b_k_stride, b_sbo = 128u;
          g += (uint32_t)(t * n_layer_steps);                           // the earlier slots' copies of this layer
          uint32_t acc = tmem_base + (uint32_t)(QP ? ((l + (XT ? tile_it : 0u)) & 1u) * 256 : t * 256);
          int sub = 0;
          ANI_TRACE(8 + 16 * l + 8 * t + 4);
          // the A operand is ready (QP: its first quarter) and the accumulator has been drained
          int next_q = 1;
          mbar_wait<PAIR == 2>(bar_a_ready + 8 * (QP ? 0 : t), a_phase, 2 + 10 * t);
          ANI_TRACE(8 + 16 * l + 8 * t + 5);
          for (int s = s0; s < s0 + n_layer_steps; ++s, ++g) {
            const Step st = s_steps[s];
            if (QP && st.a_chunk >= HID_CHUNK0) {
              const int q_last = (st.a_chunk - HID_CHUNK0 + 2 * st.n_k16 - 1) >> 3;   // hidden chunk c belongs to quarter c / 8
              for (; next_q <= q_last; ++next_q) mbar_wait<PAIR == 2>(bar_a_ready + 8 * next_q, a_phase, 6);
            }
            const uint32_t stage = g % C::STAGES, phase = (g / C::STAGES) & 1u;
            if (NT > 1 && g >= (uint32_t)C::STAGES) {
              // the previous use of this stage (possibly the other slot's) must have been consumed first
              long long t0 = clock64();
              while (s_last[stage] != g - C::STAGES) {
                if (clock64() - t0 > 2000000000ll) {
                  printf("aninerf mlp: ring order timeout (block %d slot %d g %u)\n", blockIdx.x, t, g);
                  __trap();
                }
              }
            }
            mbar_wait(bar_full + 8 * stage, phase, 3 + 10 * t);   // TMA-written operands: CTA-scope acquire is enough for the async proxy
            tc_fence_after();
            const uint32_t b_hi = smem_u32(ring + stage * STAGE_BYTES);
            const uint32_t b_lo = b_hi + (uint32_t)st.n_k16 * 2u * b_k_stride;
            if (elect_one()) {
              for (int k = 0; k < st.n_k16; ++k) {
                const uint32_t a_off = (uint32_t)(st.a_chunk + 2 * k) * CHUNK_BYTES;
                const uint64_t adh = smem_desc(a_hi + a_off, a_lbo, a_sbo);
                const uint64_t bdh = smem_desc(b_hi + (uint32_t)k * 2u * b_k_stride, b_lbo, b_sbo);
                const uint32_t fresh = ((st.flags & 1) && k == 0) ? 0u : 1u;
                umma_bf16<PAIR>(acc, adh, bdh, idesc, fresh);
                if (NPASS == 3) {
                  const uint64_t adl = smem_desc(a_lo + a_off, a_lbo, a_sbo);
                  const uint64_t bdl = smem_desc(b_lo + (uint32_t)k * 2u * b_k_stride, b_lbo, b_sbo);
                  umma_bf16<PAIR>(acc, adl, bdh, idesc, 1u);
                  umma_bf16<PAIR>(acc, adh, bdl, idesc, 1u);
                }
              }
              if (NT > 1) s_last[stage] = g;
              umma_commit<PAIR>(bar_empty + 8 * stage);        // frees the ring stage (both CTAs) once these MMAs retire
              if (st.flags & 4) umma_commit<PAIR>(bar_acc + 16);                  // half B is done with the A quarters 0,1
              if (st.flags & 2) umma_commit<PAIR>(bar_acc + 8 * (QP ? sub : t));   // (half-)layer done: its accumulator columns are ready
            }
            if (st.flags & 2) {
              ++sub;
              acc += (uint32_t)n_sub;       // the next half's accumulator columns
            }
            __syncwarp();
          }
          ANI_TRACE(8 + 16 * l + 8 * t + 6);
          if (QP)
            for (; next_q < 4; ++next_q) mbar_wait<PAIR == 2>(bar_a_ready + 8 * next_q, a_phase, 7);   // keep every quarter's phase in step
          a_phase ^= 1;
          g += (uint32_t)((NT - 1 - t) * n_layer_steps);      // the later slots' copies of this layer
        }
      }
    } else if (lane == 0 && PAIR == 2 && !leader && my_slot == 1) {
      // ===== relay (peer CTA): forward "this CTA's rows have written their A operand" to the leader =====
      const uint32_t a_ready_leader = mapa_u32(bar_a_ready, 0);
      uint32_t ph = 0;
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        for (int l = 0; l < F.n_layers; ++l) {
          for (int b = 0; b < C::N_AREADY; ++b) {
            mbar_wait(bar_a_local + 8 * b, ph, 8);
            mbar_arrive_cluster(a_ready_leader + 8 * b);
          }
          ph ^= 1;
        }
      }
    } else if (lane == 0 && PAIR == 2 && !leader && my_slot == 0) {
      // ===== relay (peer CTA): tell the leader when this CTA's half of a stage has landed ========
      uint32_t stage = 0, phase = 0;
      const uint32_t full_leader = mapa_u32(bar_full, 0);
      const int64_t total = (int64_t)F.n_steps * NT;
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        for (int64_t s = 0; s < total; ++s) {
          mbar_wait(bar_full + 8 * stage, phase, 4);
          mbar_arrive_cluster(full_leader + 8 * stage);
          if (++stage == C::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ===== row threads: input encoding, per-layer epilogues, heads =============================
    const int my_t = warp / ROW_WARPS;                 // the tile slot this warp serves (always 0 when NT == 1)
    const int rwarp = warp % ROW_WARPS;
    const int row = (rwarp & 3) * 32 + lane;           // == TMEM lane (warp % 4 selects the lane quadrant); warps w and w+4 share a row
    const int half = rwarp >> 2;                       // which half of the columns this thread owns
    const int bar_id = 1 + my_t;                       // named barrier of this slot's 256 row threads
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);   // (ROW_WARPS is a multiple of 4: warp & 3 == rwarp & 3)
    // rows always arrive CTA-locally (per-warp aggregated arrives were measured slower: the __syncwarp lengthens every
    // quarter of the epilogue by more than the serialised arrives cost)
    const uint32_t a_arrive = leader ? bar_a_ready : bar_a_local;
    uint32_t acc_phase = 0, acc_b_phase = 0;
    uint32_t tile_it = 0;
    bool pe_ready = false;                       // XT: this tile's encoding was written during the previous tile's last layer
    int64_t ngi = 0;
    bool nvalid = false;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (int64_t ut = unit; ut < n_utiles; ut += n_units, ++tile_it) {
      const bool tracing = args.trace && blockIdx.x == 0 && ut == (int64_t)args.trace_iter * n_units && threadIdx.x == 0;
      if (args.trace && blockIdx.x == 0 && threadIdx.x == 0 && ut / n_units < 90) args.trace[160 + ut / n_units] = (unsigned long long)clock64();
      int64_t gi[NT];
      bool valid[NT];
      float px[NT], py[NT], pz[NT], sigma[NT];
      float smpl[NT][ANINERF_N_BONES / 2];     // blend-weight head: the row's initial SMPL weights, 12 of the 24 bones per thread
      ANI_TRACE(0);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        if (NT > 1 && t != my_t) continue;
        uint8_t *a_hi = smem + C::OFF_A_HI + t * A_BYTES, *a_lo = smem + C::OFF_A_LO + t * A_BYTES;
        sigma[t] = 0.f;
        if (XT && pe_ready) {                          // encoded (and published) during the previous tile's last layer
          gi[t] = ngi;
          valid[t] = nvalid;
          px[t] = nx;
          py[t] = ny;
          pz[t] = nz;
          continue;
        }
        gi[t] = ((ut * NT + t) * PAIR + cta_rank) * TILE_M + row;
        valid[t] = gi[t] < n_valid;
        px[t] = py[t] = pz[t] = 0.f;
        if (valid[t]) {
          px[t] = __ldg(args.pts + 3 * gi[t]);
          py[t] = __ldg(args.pts + 3 * gi[t] + 1);
          pz[t] = __ldg(args.pts + 3 * gi[t] + 2);
        }
        if (half == 0) write_pe<NPASS, 10, 0, 4>(a_hi, a_lo, PE_CHUNK0, row, px[t], py[t], pz[t]);
        else write_pe<NPASS, 10, 4, 8>(a_hi, a_lo, PE_CHUNK0, row, px[t], py[t], pz[t]);
        fence_proxy_async();
#pragma unroll
        for (int q = 0; q < (QP ? 4 : 1); ++q) {       // QP: the input encoding readies every quarter's barrier for layer 0
          const int bi = QP ? q : t;
          mbar_arrive(a_arrive + 8 * bi);
        }
      }
      ANI_TRACE(1);

      for (int l = 0; l < F.n_layers; ++l) {
        const bool last = l == F.n_layers - 1;
        const float *bias = s_bias + l * 256;
        const int n_pad = F.layers[l].n_pad;
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          if (NT > 1 && t != my_t) continue;
          uint8_t *a_hi = smem + C::OFF_A_HI + t * A_BYTES, *a_lo = smem + C::OFF_A_LO + t * A_BYTES;
          float *xchg = s_xchg + t * (TILE_M * 4);           // this slot's exchange between the row's two threads
          const uint32_t t_acc = t_lane + (uint32_t)(QP ? ((l + (XT ? tile_it : 0u)) & 1u) * 256 : t * 256);
          if (!NERF && l < 8) {
            // Initial SMPL weights of this row (this thread: bones 12*half .. +11): ONE trilinear corner per layer, fetched in
            // the idle window before the layer's accumulator is ready and accumulated in registers (ATen's corner order).
            // The whole gather is 98 KB per tile from L2, which also feeds the 1 MB weight stream: done in one burst it ran at
            // the L2 bandwidth roof for 7-8k cycles and delayed the layer it shared the window with; 12 KB per window hides.
            if (l == 0) {
              ANI_TRACE(4);
#pragma unroll
              for (int k = 0; k < ANINERF_N_BONES / 2; ++k) smpl[t][k] = 0.f;
            }
            if (valid[t]) {
              if (args.smpl_bw) {
                if (l == 7) {
                  const float4 *r4 = reinterpret_cast<const float4 *>(args.smpl_bw + gi[t] * ANINERF_N_BONES) + half * 3;
#pragma unroll
                  for (int q = 0; q < 3; ++q) {
                    float4 w4 = __ldg(r4 + q);
                    smpl[t][4 * q] = w4.x;
                    smpl[t][4 * q + 1] = w4.y;
                    smpl[t][4 * q + 2] = w4.z;
                    smpl[t][4 * q + 3] = w4.w;
                  }
                }
              } else {
                float wc;
                int oc;
                fast_corner(s_grid, args.grid_dim, px[t], py[t], pz[t], l, wc, oc);
                const float4 *r4 = reinterpret_cast<const float4 *>(args.vol_w24 + (int64_t)oc * ANINERF_N_BONES) + half * 3;
                float4 c4[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) c4[q] = __ldg(r4 + q);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                  smpl[t][4 * q] = fmaf(c4[q].x, wc, smpl[t][4 * q]);
                  smpl[t][4 * q + 1] = fmaf(c4[q].y, wc, smpl[t][4 * q + 1]);
                  smpl[t][4 * q + 2] = fmaf(c4[q].z, wc, smpl[t][4 * q + 2]);
                  smpl[t][4 * q + 3] = fmaf(c4[q].w, wc, smpl[t][4 * q + 3]);
                }
              }
            }
            if (l == 7) ANI_TRACE(5);
          }
          ANI_TRACE(8 + 16 * l + 8 * t);
          mbar_wait(bar_acc + 8 * (QP ? 0 : t), acc_phase, 5 + 10 * t + 100 * l);
          tc_fence_after();
          if (XT && last) {
            // The next tile's input encoding.  It must come AFTER this layer's accumulator barrier: an mbarrier arrival is not tagged
            // with a phase, and only the completed MMAs of layer 8 prove that every thread's layer-8 arrivals on a_ready[] are in.
            pe_ready = ut + n_units < n_utiles;
            if (pe_ready) {
              ngi = (((ut + n_units) * NT + t) * PAIR + cta_rank) * TILE_M + row;
              nvalid = ngi < n_valid;
              nx = ny = nz = 0.f;
              if (nvalid) {
                nx = __ldg(args.pts + 3 * ngi);
                ny = __ldg(args.pts + 3 * ngi + 1);
                nz = __ldg(args.pts + 3 * ngi + 2);
              }
              if (half == 0) write_pe<NPASS, 10, 0, 4>(a_hi, a_lo, PE_CHUNK0, row, nx, ny, nz);
              else write_pe<NPASS, 10, 4, 8>(a_hi, a_lo, PE_CHUNK0, row, nx, ny, nz);
              fence_proxy_async();
#pragma unroll
              for (int q = 0; q < 4; ++q) mbar_arrive(a_arrive + 8 * q);
            }
          }
          ANI_TRACE(8 + 16 * l + 8 * t + 1);
          if (!last) {
            if (NERF && l == VIEW_LAYER_WRITE) {
              // layer 5 was the last reader of PE(xyz): the PE chunks now take PE(viewdir) for the view layer
              float vx = 0.f, vy = 0.f, vz = 0.f;
              if (valid[t]) {
                vx = __ldg(args.viewdir + 3 * gi[t]);
                vy = __ldg(args.viewdir + 3 * gi[t] + 1);
                vz = __ldg(args.viewdir + 3 * gi[t] + 2);
              }
              if (half == 0) write_pe<NPASS, 4, 0, 2>(a_hi, a_lo, PE_CHUNK0, row, vx, vy, vz);
              else write_pe<NPASS, 4, 2, 4>(a_hi, a_lo, PE_CHUNK0, row, vx, vy, vz);
            }
            // hidden layer: bias + ReLU -> bf16 (hi/lo) -> A chunks 8..39 in place; this thread: 128 of the 256 columns
            // (hidden layers are 256 wide: 4 groups of 32 columns per thread), software-pipelined: the TMEM load of
            // group g+1 is in flight while group g is converted and stored
            const bool alpha_layer = NERF && (l == F.n_layers - 2);
            uint32_t va[32], vb[32];
            auto process = [&](const uint32_t (&v)[32], int g) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                float x[8];
                const float4 b0 = *reinterpret_cast<const float4 *>(bias + g * 32 + q * 8);
                const float4 b1 = *reinterpret_cast<const float4 *>(bias + g * 32 + q * 8 + 4);
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                if (NPASS == 1 && !alpha_layer) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[q * 8 + j]) + bb[j];
                  store_chunk<NPASS, NPASS == 1>(a_hi, a_lo, HID_CHUNK0 + g * 4 + q, row, x);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) x[j] = fmaxf(__uint_as_float(v[q * 8 + j]) + bb[j], 0.f);
                  if (alpha_layer) {
                    const float4 w0 = *reinterpret_cast<const float4 *>(s_head + g * 32 + q * 8);
                    const float4 w1 = *reinterpret_cast<const float4 *>(s_head + g * 32 + q * 8 + 4);
                    float sg = sigma[t];
                    sg = fmaf(x[0], w0.x, sg);
                    sg = fmaf(x[1], w0.y, sg);
                    sg = fmaf(x[2], w0.z, sg);
                    sg = fmaf(x[3], w0.w, sg);
                    sg = fmaf(x[4], w1.x, sg);
                    sg = fmaf(x[5], w1.y, sg);
                    sg = fmaf(x[6], w1.z, sg);
                    sg = fmaf(x[7], w1.w, sg);
                    sigma[t] = sg;
                  }
                  store_chunk<NPASS>(a_hi, a_lo, HID_CHUNK0 + g * 4 + q, row, x);
                }
              }
            };
            // publish the columns written so far (QP: after every group = one quarter of the next layer's K)
            auto publish = [&](int q) {
              tc_fence_before();
              fence_proxy_async();
              mbar_arrive(a_arrive + 8 * q);
            };
            // this thread's i-th group of 32 columns: QP interleaves the row's two threads inside every quarter
            const int ga = QP ? half : half * 4, gs = QP ? 2 : 1;
            if (QP) {
              // half A of the layer (output columns 0-127 = quarters 0,1 of the next K); half B's MMAs run meanwhile and still
              // read the old quarters 0,1 until `bread`
              tmem_ld32(t_acc + ga * 32, va);
              tmem_ld_wait();
              tmem_ld32(t_acc + (ga + gs) * 32, vb);
              mbar_wait(bar_acc + 16, acc_b_phase, 9 + 100 * l);
              process(va, ga);
              publish(0);
              tmem_ld_wait();
              process(vb, ga + gs);
              publish(1);
              // half B (output columns 128-255 = quarters 2,3)
              mbar_wait(bar_acc + 8, acc_b_phase, 10 + 100 * l);
              tc_fence_after();
              acc_b_phase ^= 1;
              tmem_ld32(t_acc + (ga + 2 * gs) * 32, va);
              tmem_ld_wait();
              tmem_ld32(t_acc + (ga + 3 * gs) * 32, vb);
              process(va, ga + 2 * gs);
              publish(2);
              tmem_ld_wait();
              process(vb, ga + 3 * gs);
            } else {
              tmem_ld32(t_acc + ga * 32, va);
              tmem_ld_wait();
              tmem_ld32(t_acc + (ga + gs) * 32, vb);
              process(va, ga);
              tmem_ld_wait();
              tmem_ld32(t_acc + (ga + 2 * gs) * 32, va);
              process(vb, ga + gs);
              tmem_ld_wait();
              tmem_ld32(t_acc + (ga + 3 * gs) * 32, vb);
              process(va, ga + 2 * gs);
              tmem_ld_wait();
              process(vb, ga + 3 * gs);
            }
            if (alpha_layer && half == 1) xchg[row * 4] = sigma[t];   // read by the row's other thread after the next acc barrier
            publish(QP ? 3 : t);
            ANI_TRACE(8 + 16 * l + 8 * t + 2);
          } else if (!NERF) {
            // ---- blend-weight head: softmax(log(smpl_bw + 1e-9) + delta), fused inverse LBS --------
            // The row's two threads take 12 bones each and meet three times (max, sum, skinning matrix) through a scratch
            // area in the A operand, which is dead until the next tile's input encoding.
            uint32_t v[32];
            tmem_ld32(t_acc, v);
            tmem_ld_wait();
            tc_fence_before();
            float *scr = reinterpret_cast<float *>(a_hi + 8 * CHUNK_BYTES);   // exchange scratch: the hidden chunks (all MMAs have retired)
            constexpr int HB = ANINERF_N_BONES / 2;
            const int k0 = half * HB;
            float bw[HB];
            float mx = -INFINITY;
#pragma unroll
            for (int k = 0; k < HB; ++k) {
              const float d = __uint_as_float(half ? v[HB + k] : v[k]);
              bw[k] = logf(smpl[t][k] + 1e-9f) + (d + bias[k0 + k]);
              mx = fmaxf(mx, bw[k]);
            }
            scr[half * TILE_M + row] = mx;
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");
            mx = fmaxf(scr[row], scr[TILE_M + row]);
            float part = 0.f;
#pragma unroll
            for (int k = 0; k < HB; ++k) {
              bw[k] = expf(bw[k] - mx);
              part += bw[k];
            }
            scr[(2 + half) * TILE_M + row] = part;
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");
            const float inv_sum = 1.0f / (scr[2 * TILE_M + row] + scr[3 * TILE_M + row]);
#pragma unroll
            for (int k = 0; k < HB; ++k) bw[k] *= inv_sum;
            if (valid[t] && args.bw_out) {
              float4 *o4 = reinterpret_cast<float4 *>(args.bw_out + gi[t] * ANINERF_N_BONES) + half * 3;
#pragma unroll
              for (int q = 0; q < 3; ++q) o4[q] = make_float4(bw[4 * q], bw[4 * q + 1], bw[4 * q + 2], bw[4 * q + 3]);
            }
            if (args.tpts_out) {
              float M[12];
#pragma unroll
              for (int j = 0; j < 12; ++j) M[j] = 0.f;
#pragma unroll
              for (int k = 0; k < HB; ++k)
#pragma unroll
                for (int j = 0; j < 12; ++j) M[j] = fmaf(bw[k], s_head[(k0 + k) * 12 + j], M[j]);
              if (half == 1) {
#pragma unroll
                for (int j = 0; j < 12; ++j) scr[(4 + j) * TILE_M + row] = M[j];
              }
              asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");
              if (half == 0 && valid[t]) {
#pragma unroll
                for (int j = 0; j < 12; ++j) M[j] += scr[(4 + j) * TILE_M + row];
                float qx = px[t] - M[3], qy = py[t] - M[7], qz = pz[t] - M[11];
                float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], kk = M[10];
                float c00 = e * kk - f * h, c01 = c * h - b * kk, c02 = b * f - c * e;
                float c10 = f * g - d * kk, c11 = a * kk - c * g, c12 = c * d - a * f;
                float c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
                float inv = 1.0f / (a * c00 + b * c10 + c * c20);
                args.tpts_out[3 * gi[t]] = (c00 * qx + c01 * qy + c02 * qz) * inv;
                args.tpts_out[3 * gi[t] + 1] = (c10 * qx + c11 * qy + c12 * qz) * inv;
                args.tpts_out[3 * gi[t] + 2] = (c20 * qx + c21 * qy + c22 * qz) * inv;
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");   // the scratch is the next tile's PE operand
          } else {
            // ---- NeRF head: view layer (ReLU) -> rgb_fc in fp32; alpha from the layer-7 epilogue ---
            float rgb[3] = {0.f, 0.f, 0.f};
            const int g0 = half * (n_pad / 64);
            for (int g = g0; g < g0 + n_pad / 64; ++g) {
              uint32_t v[32];
              tmem_ld32(t_acc + g * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float x = fmaxf(__uint_as_float(v[j]) + bias[g * 32 + j], 0.f);
                rgb[0] = fmaf(x, s_head[257 + g * 32 + j], rgb[0]);
                rgb[1] = fmaf(x, s_head[257 + 128 + g * 32 + j], rgb[1]);
                rgb[2] = fmaf(x, s_head[257 + 256 + g * 32 + j], rgb[2]);
              }
            }
            tc_fence_before();
            float my_sigma = sigma[t];
            if (half == 1) {
              xchg[row * 4 + 1] = rgb[0];
              xchg[row * 4 + 2] = rgb[1];
              xchg[row * 4 + 3] = rgb[2];
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");   // the row's two threads meet
            if (half == 0) {
              my_sigma += xchg[row * 4] + s_head[256];
              rgb[0] += xchg[row * 4 + 1] + s_head[257 + 384];
              rgb[1] += xchg[row * 4 + 2] + s_head[257 + 385];
              rgb[2] += xchg[row * 4 + 3] + s_head[257 + 386];
              if (valid[t]) {
                const int64_t o = gi[t];
                if (args.sigma_out) args.sigma_out[o] = my_sigma;
                if (args.rgb_out) {
                  args.rgb_out[3 * o] = rgb[0];
                  args.rgb_out[3 * o + 1] = rgb[1];
                  args.rgb_out[3 * o + 2] = rgb[2];
                }
                if (args.raw_out) {
                  // tail of Network.forward (tpose_nerf_network.py:186-212)
                  bool inside = px[t] > args.tbounds[0] && px[t] < args.tbounds[3] && py[t] > args.tbounds[1] && py[t] < args.tbounds[4] &&
                                pz[t] > args.tbounds[2] && pz[t] < args.tbounds[5];
                  float sg = inside ? my_sigma : 0.f;
                  if (args.sigma_masked_out) args.sigma_masked_out[o] = sg;
                  float al = 1.0f - expf(-fmaxf(sg, 0.f) * __ldg(args.dists + o));
                  float4 rv = make_float4(1.0f / (1.0f + expf(-rgb[0])), 1.0f / (1.0f + expf(-rgb[1])), 1.0f / (1.0f + expf(-rgb[2])), al);
                  reinterpret_cast<float4 *>(args.raw_out)[args.index ? (int64_t)__ldg(args.index + o) : o] = rv;   // dense scatter, or compact rows
                }
              }
            }
            asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(ROW_THREADS) : "memory");   // s_xchg is rewritten by the next slot / tile
          }
        }
        acc_phase ^= 1;
      }
      ANI_TRACE(3);
    }
  }

  // ---- teardown --------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();        // the leader's MMAs read the peer's shared memory: leave together
  if (warp == RW + 1) {
    tc_fence_after();
    tmem_dealloc<PAIR>(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: packing and the net object
// ------------------------------------------------------------------------------------------------
static inline uint16_t f2bf(float f) {   // round to nearest even
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static inline float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// source columns [src0, src0+len) of a layer's weight matrix, zero-padded to `pad` K elements,
// multiplied against the A chunks starting at `a_chunk0`
struct Segment { int src0, len, pad, a_chunk0; };

struct HostLayer {
  int n_out, n_pad, k_in, relu, n_tables;
  std::vector<Segment> segs;
};

struct FieldImage {          // per (field, precision)
  uint8_t *image = nullptr;
  Step *steps = nullptr;
  int n_steps = 0;
  int step0[MAX_LAYERS] = {0}, n_layer_steps[MAX_LAYERS] = {0};
};

struct FieldHost {
  bool loaded = false;
  int n_layers = 0;
  LayerDev layers[MAX_LAYERS];
  float *bias = nullptr;
  float *head = nullptr;
  FieldImage img[2];         // [0]: NPASS=1, [1]: NPASS=3
};

}  // namespace aninerf

struct aninerf_net {
  aninerf::FieldHost fields[ANINERF_N_FIELDS];
};

namespace aninerf {

// CTAs per tcgen05.mma (cta_group): CTA pairs, each CTA holds half of every weight chunk
constexpr int kPair = 2;

static void free_field(FieldHost &f) {
  cudaFree(f.bias);
  cudaFree(f.head);
  for (auto &im : f.img) {
    cudaFree(im.image);
    cudaFree(im.steps);
    im = FieldImage();
  }
  f = FieldHost();
}

// architecture tables (tpose_nerf_network.py:12-38, 219-239, 279-294 after folding)
static int describe_layers(int field, const aninerf_layer *L, int n_layers, std::vector<HostLayer> &out) {
  const bool nerf = field == ANINERF_FIELD_NERF;
  if (n_layers != 9) return fail(ANINERF_EINVAL, "%s: a field has 9 dense layers after folding%s", "aninerf_net_load_field");
  for (int l = 0; l < 9; ++l) {
    HostLayer h;
    h.n_out = L[l].n_out;
    h.k_in = L[l].k_in;
    h.relu = L[l].relu;
    h.n_tables = L[l].n_tables;
    int want_k, want_n;
    if (l == 0) {
      want_k = 63; want_n = 256; h.segs = {{0, 63, 64, PE_CHUNK0}};
    } else if (l == 5) {     // skip layer: [PE(xyz), hidden]
      want_k = 63 + 256; want_n = 256; h.segs = {{0, 63, 64, PE_CHUNK0}, {63, 256, 256, HID_CHUNK0}};
    } else if (l < 8) {
      want_k = 256; want_n = 256; h.segs = {{0, 256, 256, HID_CHUNK0}};
    } else if (nerf) {       // folded view layer: [hidden, PE(viewdir)]; PE(viewdir) sits in the old PE(xyz) chunks
      want_k = 256 + 27; want_n = 128; h.segs = {{256, 27, 32, PE_CHUNK0}, {0, 256, 256, HID_CHUNK0}};
    } else {
      want_k = 256; want_n = ANINERF_N_BONES; h.segs = {{0, 256, 256, HID_CHUNK0}};
    }
    if (h.k_in != want_k || h.n_out != want_n || !L[l].W || !L[l].bias_table || h.n_tables < 1)
      return fail(ANINERF_EINVAL, "%s: layer shape does not match the aninerf architecture%s", "aninerf_net_load_field");
    h.n_pad = (h.n_out + 31) / 32 * 32;
    out.push_back(h);
  }
  return ANINERF_OK;
}

static int build_image(const aninerf_layer *L, const std::vector<HostLayer> &H, int npass, FieldImage &im, cudaStream_t st) {
  const int hi_max = npass == 3 ? STAGE_BYTES / 2 : STAGE_BYTES;   // per CTA
  std::vector<Step> steps;
  std::vector<uint8_t> image;
  for (size_t l = 0; l < H.size(); ++l) {
    const HostLayer &h = H[l];
    // the split-precision kernel (NT = 1) runs a 256-wide layer as two N=128 halves (see the kernel's N-split note)
    const int n_subs = (npass == 3 && h.n_pad == 256) ? 2 : 1;
    const int n_sub = h.n_pad / n_subs;
    const int n_half = n_sub / kPair;                       // weight rows held by each CTA of the pair, per MMA
    const int per_step = std::max(1, hi_max / (n_half * 32));
    im.step0[l] = (int)steps.size();
    for (int sub = 0; sub < n_subs; ++sub) {
    bool bread_set = false;
    for (size_t si = 0; si < h.segs.size(); ++si) {
      const Segment &sg = h.segs[si];
      const int k16_total = sg.pad / 16;
      for (int k0 = 0; k0 < k16_total; k0 += per_step) {
        const int nk = std::min(per_step, k16_total - k0);
        Step s;
        s.w_off = (uint32_t)image.size();
        const size_t hi_bytes = (size_t)nk * 2 * n_half * 16;            // per CTA
        const size_t cta_bytes = hi_bytes * (npass == 3 ? 2 : 1);
        s.bytes = (uint32_t)(cta_bytes * kPair);
        s.a_chunk = (uint16_t)(sg.a_chunk0 + 2 * k0);
        s.n_k16 = (uint16_t)nk;
        s.layer = (uint16_t)l;
        s.flags = (uint16_t)(((si == 0 && k0 == 0) ? 1 : 0) | ((si + 1 == h.segs.size() && k0 + nk >= k16_total) ? 2 : 0));
        if (n_subs == 2 && sub == 1 && !bread_set) {
          // half B: after this step the hidden quarters 0,1 (A chunks HID_CHUNK0 .. HID_CHUNK0+15) are not read again
          const bool reads_hidden = sg.a_chunk0 >= HID_CHUNK0;
          const bool later_hidden = !reads_hidden && si + 1 < h.segs.size();      // a PE segment followed by the hidden segment
          const bool covers = reads_hidden && (sg.a_chunk0 + 2 * (k0 + nk)) >= HID_CHUNK0 + 16;
          const bool last_of_all = si + 1 == h.segs.size() && k0 + nk >= k16_total;
          if ((covers || last_of_all) && !later_hidden) {
            s.flags |= 4;
            bread_set = true;
          }
        }
        image.resize(image.size() + s.bytes, 0);
        for (int r = 0; r < kPair; ++r) {           // image = [CTA0: hi, lo][CTA1: hi, lo]
          uint16_t *hi = reinterpret_cast<uint16_t *>(image.data() + s.w_off + r * cta_bytes);
          uint16_t *lo = reinterpret_cast<uint16_t *>(image.data() + s.w_off + r * cta_bytes + hi_bytes);
          for (int c = 0; c < nk * 2; ++c)          // 8-wide K chunk
            for (int nn = 0; nn < n_half; ++nn)
              for (int j = 0; j < 8; ++j) {
                const int n = sub * n_sub + r * n_half + nn;
                const int k = (k0 * 2 + c) * 8 + j;         // K index inside the segment
                float w = 0.f;
                if (n < h.n_out && k < sg.len) w = L[l].W[(size_t)n * h.k_in + sg.src0 + k];
                const uint16_t wh = f2bf(w);
                const size_t o = ((size_t)c * n_half + nn) * 8 + j;
                hi[o] = wh;
                if (npass == 3) lo[o] = f2bf(w - bf2f(wh));
              }
        }
        steps.push_back(s);
      }
    }
    }
    im.n_layer_steps[l] = (int)steps.size() - im.step0[l];
  }
  if (steps.size() > (size_t)MAX_STEPS) return fail(ANINERF_EINVAL, "%s: internal: too many steps%s", __func__);
  for (auto &s : steps)
    if (s.bytes / kPair > (uint32_t)STAGE_BYTES || ((s.bytes / kPair) & 15u)) return fail(ANINERF_EINVAL, "%s: internal: bad step size%s", __func__);
  ANI_CUDA(cudaMalloc(&im.image, image.size()));
  ANI_CUDA(cudaMalloc(&im.steps, steps.size() * sizeof(Step)));
  ANI_CUDA(cudaMemcpyAsync(im.image, image.data(), image.size(), cudaMemcpyHostToDevice, st));
  ANI_CUDA(cudaMemcpyAsync(im.steps, steps.data(), steps.size() * sizeof(Step), cudaMemcpyHostToDevice, st));
  ANI_CUDA(cudaStreamSynchronize(st));   // the std::vectors go out of scope
  im.n_steps = (int)steps.size();
  return ANINERF_OK;
}

template <int NPASS, bool NERF, int NT>
static int launch_mlp(const MlpArgs &a, cudaStream_t st) {
  using C = Cfg<NPASS, NERF, NT>;
  static bool configured = false;
  if (!configured) {
    ANI_CUDA(cudaFuncSetAttribute(mlp_kernel<NPASS, NERF, kPair, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  const int64_t rows_per_unit = (int64_t)TILE_M * kPair * NT;
  const int64_t utiles = (a.n + rows_per_unit - 1) / rows_per_unit;
  const int units = (int)std::min<int64_t>(utiles, sm_count() / kPair);
  if (units <= 0) return ANINERF_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(units * kPair));
  cfg.blockDim = dim3(n_threads(NT));
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ANI_CUDA(cudaLaunchKernelEx(&cfg, mlp_kernel<NPASS, NERF, kPair, NT>, a));
  ANI_LAUNCHED();
  return ANINERF_OK;
}

static unsigned long long *g_trace = nullptr;
static int g_trace_iter = 0;
static int g_trace_field = -1;   // -1: every field's kernel writes the trace; else only this field

static int fill_field(const aninerf_net *net, int field, int precision, MlpArgs &a) {
  if (!net) return fail(ANINERF_EINVAL, "%s: null net%s", __func__);
  if (field < 0 || field >= ANINERF_N_FIELDS) return fail(ANINERF_EINVAL, "%s: bad field%s", __func__);
  if (precision != 1 && precision != 3) return fail(ANINERF_EINVAL, "%s: precision must be 1 or 3%s", __func__);
  const FieldHost &f = net->fields[field];
  if (!f.loaded) return fail(ANINERF_ESTATE, "%s: field weights not loaded%s", __func__);
  const FieldImage &im = f.img[precision == 3 ? 1 : 0];
  a.f.image = im.image;
  a.f.steps = im.steps;
  a.f.n_steps = im.n_steps;
  a.f.n_layers = f.n_layers;
  memcpy(a.f.layers, f.layers, sizeof(f.layers));
  for (int l = 0; l < f.n_layers; ++l) {
    a.f.layers[l].step0 = im.step0[l];
    a.f.layers[l].n_steps = im.n_layer_steps[l];
  }
  a.f.bias = f.bias;
  a.f.head = f.head;
  a.trace = (g_trace_field < 0 || g_trace_field == field) ? g_trace : nullptr;
  a.trace_iter = g_trace_iter;
  return ANINERF_OK;
}

// internal C++ entry points shared with render.cu
int bw_forward_impl(aninerf_net *net, int field, int latent_index, const int64_t *latent_dev, const float *pts, const float *smpl_bw, const float *vol_w24,
                    const int32_t dims[3], const float *bounds, int64_t n, const int32_t *n_dev, const float *A, float *bw_out,
                    float *tpts_out, int precision, cudaStream_t st) {
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_field(net, field, precision, a);
  if (rc) return rc;
  a.latent_index = latent_index;
  a.latent_dev = latent_dev;
  a.pts = pts;
  a.n = n;
  a.n_dev = n_dev;
  a.smpl_bw = smpl_bw;
  a.vol_w24 = vol_w24;
  if (!smpl_bw) {
    a.grid_bounds = bounds;
    for (int k = 0; k < 3; ++k) a.grid_dim[k] = dims[k];
  }
  a.A = A;
  a.bw_out = bw_out;
  a.tpts_out = tpts_out;
  return precision == 3 ? launch_mlp<3, false, 1>(a, st) : launch_mlp<1, false, 2>(a, st);
}

int nerf_forward_impl(aninerf_net *net, int latent_index, const int64_t *latent_dev, const float *pts, const float *viewdir, int64_t n, const int32_t *n_dev,
                      float *sigma_out, float *rgb_out, const float *dists, const float *tbounds, const int32_t *index, float *raw_out,
                      float *sigma_masked_out, int precision, cudaStream_t st) {
  MlpArgs a;
  memset(&a, 0, sizeof(a));
  int rc = fill_field(net, ANINERF_FIELD_NERF, precision, a);
  if (rc) return rc;
  a.latent_index = latent_index;
  a.latent_dev = latent_dev;
  a.pts = pts;
  a.viewdir = viewdir;
  a.n = n;
  a.n_dev = n_dev;
  a.sigma_out = sigma_out;
  a.rgb_out = rgb_out;
  a.dists = dists;
  a.tbounds = tbounds;
  a.index = index;
  a.raw_out = raw_out;
  a.sigma_masked_out = sigma_masked_out;
  return precision == 3 ? launch_mlp<3, true, 1>(a, st) : launch_mlp<1, true, 2>(a, st);
}

}  // namespace aninerf

using namespace aninerf;

extern "C" {

int aninerf_debug_set_trace(unsigned long long *device_buf) {
  g_trace = device_buf;
  const char *e = getenv("ANINERF_TRACE_ITER");   // which of block 0's tiles to trace (default: the first)
  g_trace_iter = e ? atoi(e) : 0;
  e = getenv("ANINERF_TRACE_FIELD");
  g_trace_field = e ? atoi(e) : -1;
  return ANINERF_OK;
}

int aninerf_net_create(aninerf_net **out) {
  ANI_CHECK_ARG(out);
  *out = new aninerf_net();
  return ANINERF_OK;
}

int aninerf_net_destroy(aninerf_net *net) {
  if (!net) return ANINERF_OK;
  for (auto &f : net->fields) free_field(f);
  delete net;
  return ANINERF_OK;
}

int aninerf_net_load_field(aninerf_net *net, int32_t field, const aninerf_layer *layers, int32_t n_layers, const float *alpha_w,
                           const float *alpha_b, const float *rgb_w, const float *rgb_b, void *stream) {
  ANI_CHECK_ARG(net && layers && field >= 0 && field < ANINERF_N_FIELDS);
  cudaStream_t st = (cudaStream_t)stream;
  const bool nerf = field == ANINERF_FIELD_NERF;
  if (nerf) ANI_CHECK_ARG(alpha_w && alpha_b && rgb_w && rgb_b);
  std::vector<HostLayer> H;
  int rc = describe_layers(field, layers, n_layers, H);
  if (rc) return rc;
  FieldHost &f = net->fields[field];
  ANI_CUDA(cudaStreamSynchronize(st));   // nothing may still be reading the old images
  free_field(f);
  f.n_layers = n_layers;
  // bias tables
  std::vector<float> bias;
  for (int l = 0; l < n_layers; ++l) {
    f.layers[l].n_pad = H[l].n_pad;
    f.layers[l].n_out = H[l].n_out;
    f.layers[l].relu = H[l].relu;
    f.layers[l].bias_off = (int)bias.size();
    f.layers[l].n_tables = H[l].n_tables;
    bias.insert(bias.end(), layers[l].bias_table, layers[l].bias_table + (size_t)H[l].n_tables * H[l].n_out);
  }
  ANI_CUDA(cudaMalloc(&f.bias, bias.size() * 4));
  ANI_CUDA(cudaMemcpyAsync(f.bias, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, st));
  if (nerf) {
    std::vector<float> head(644, 0.f);
    memcpy(head.data(), alpha_w, 256 * 4);
    head[256] = alpha_b[0];
    memcpy(head.data() + 257, rgb_w, 384 * 4);
    memcpy(head.data() + 257 + 384, rgb_b, 3 * 4);
    ANI_CUDA(cudaMalloc(&f.head, head.size() * 4));
    ANI_CUDA(cudaMemcpyAsync(f.head, head.data(), head.size() * 4, cudaMemcpyHostToDevice, st));
    ANI_CUDA(cudaStreamSynchronize(st));
  }
  ANI_CUDA(cudaStreamSynchronize(st));
  rc = build_image(layers, H, 1, f.img[0], st);
  if (rc) return rc;
  rc = build_image(layers, H, 3, f.img[1], st);
  if (rc) return rc;
  f.loaded = true;
  return ANINERF_OK;
}

int aninerf_bw_forward(aninerf_net *net, int32_t field, int32_t latent_index, const float *pts, const float *smpl_bw, int64_t n,
                       const int32_t *n_dev, const float *A, float *bw_out, float *tpts_out, int32_t precision, void *stream) {
  ANI_CHECK_ARG(net && pts && smpl_bw && n >= 0 && (field == ANINERF_FIELD_BW || field == ANINERF_FIELD_NOVEL_BW));
  ANI_CHECK_ARG(!tpts_out || A);
  if (n == 0) return ANINERF_OK;
  return bw_forward_impl(net, field, latent_index, nullptr, pts, smpl_bw, nullptr, nullptr, nullptr, n, n_dev, A, bw_out, tpts_out, precision,
                         (cudaStream_t)stream);
}

int aninerf_nerf_forward(aninerf_net *net, int32_t latent_index, const float *pts, const float *viewdir, int64_t n, const int32_t *n_dev,
                         float *sigma_out, float *rgb_out, const float *dists, const float *tbounds, const int32_t *index, float *raw_out,
                         float *sigma_masked_out, int32_t precision, void *stream) {
  ANI_CHECK_ARG(net && pts && viewdir && n >= 0);
  ANI_CHECK_ARG(!raw_out || (dists && tbounds && index));
  if (n == 0) return ANINERF_OK;
  return nerf_forward_impl(net, latent_index, nullptr, pts, viewdir, n, n_dev, sigma_out, rgb_out, dists, tbounds, index, raw_out, sigma_masked_out,
                           precision, (cudaStream_t)stream);
}

}  // extern "C"
