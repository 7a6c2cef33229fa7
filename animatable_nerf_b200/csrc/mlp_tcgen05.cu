// The two dense fields of the path -- the neural blend-weight MLP (tpose_nerf_network.py:55-77,
// :304-315) and the canonical NeRF MLP (tpose_nerf_network.py:252-275) -- as ONE persistent,
// warp-specialised tcgen05 kernel family for sm_100a.
//
// Work unit: a CTA PAIR (2-CTA cluster, tcgen05 cta_group::2, M = 256).  Each CTA owns 128 samples; their fp32 accumulator lives
// in its TMEM (128 lanes x 256 columns, two buffers alternating by layer).  THE HIDDEN ACTIVATIONS NEVER LEAVE TENSOR MEMORY: the
// epilogue of layer l reads its accumulator buffer with tcgen05.ld, applies bias + ReLU, re-quantises to bf16 (hi, and lo for the
// split-precision mode) and writes the packed pairs with tcgen05.st IN PLACE over the columns it has just read -- a retired
// accumulator buffer is exactly big enough for the next layer's A operand (x1: 128 of its 256 columns, x3: all of them) -- and
// layer l+1 is issued in the .ts form (A from tensor memory, `tcgen05.mma [d], [a], b_desc`), accumulating into the OTHER
// buffer.  Only the 64-wide input encodings (PE(xyz), PE(viewdir)) are shared-memory operands (K-major SWIZZLE_128B blocks).
// Weights stream from L2 through a ring of stages (split precision: 4 x 32 KB, single pass: 2 x 64 KB + a 16 KB auxiliary slot)
// filled by bulk TMA copies (cp.async.bulk) of pre-packed SWIZZLE_128B blocks; each CTA of the pair loads HALF of every block (the
// M=256 MMA reads B from both CTAs), which halves the L2 -> SM weight traffic per sample.  One ELECTED thread of the leader CTA
// issues tcgen05.mma (M=256, N<=256, K=16); eight epilogue warps per CTA (two threads per row, each owning half of every
// 64-column quarter).  Positional encoding is generated in-kernel straight into the A operand; the last epilogue is the field's
// head (softmax + inverse LBS, or alpha/rgb activation + tbounds masking + scatter).
//
// Why: tools/bench_mma.cu (profiles/r02_mma_microbench_*.log), DESIGN.md section 5.  (1) Issued from an elected lane with
// loop-invariant descriptors an M=256 K=16 MMA retires in 129.4 (N=256) / 68.5 (N=128) / 28 (N=32, A in tensor memory) cycles in
// every operand layout; issued from `lane == 0` of a divergent warp the compiler's ELECT / BRA.U.ANY waterfall loop makes the
// ISSUE cost ~105 cycles per MMA (what looked like a hardware floor in the first version of the benchmark and of this kernel).
// (2) The ONE thread that issues the MMAs is the scarcest resource of the kernel: every dependent scalar instruction, mbarrier
// try_wait (~100 cycles even when complete) and commit between two MMAs is tensor-pipe idle time once the ~7-deep MMA queue
// has drained.  (3) With A in shared memory the operand reads, the epilogue's writes and the weight ring share one 128 B/cycle
// port; with A in tensor memory the port carries the weights only.
//
// Pipelining (QP, "quarter pipelining"): the epilogue publishes the next layer's operand one 64-column K-block at a time
// (a_ready[q]), so the MMAs of layer l+1 start as soon as the first quarter is written and run while the epilogue produces the
// other three.
//
// Precision modes: NPASS=1 single bf16 product; NPASS=3 "bf16x3": x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// with fp32 accumulation (fp32-equivalent; the blend-weight field needs it for the 1e-5 gate).
#include <cuda_bf16.h>
#include <stdlib.h>

#include <algorithm>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "tcgen05.cuh"

namespace aninerf {

// ------------------------------------------------------------------------------------------------
// layout constants
// ------------------------------------------------------------------------------------------------
constexpr int TILE_M = 128;
constexpr int KB_BYTES = TILE_M * 128;       // one K-block of the A tile: 128 rows x 64 bf16 (SWIZZLE_128B), 16 KB
constexpr int PE_CHUNK0 = 0;                 // shared-memory A block 0 (chunks of 8 K elements 0..7): PE(xyz) 63 + pad
constexpr int VIEW_CHUNK0 = 8;               // shared-memory A block 1 (chunks 8..15): PE(viewdir) 27 + pad (NeRF field)
constexpr int KB_PE = 0, KB_HID0 = 1, KB_VIEW = 5;   // Step::a_kb: 0 PE(xyz) block, 1..4 hidden quarters (tensor memory), 5 PE(viewdir) block
// Ring stage.  Every stage boundary costs the MMA-issuing thread a commit, a decode and a `full` wait, during which the tensor pipe
// drains unless enough MMAs are queued (profiles/r02_mlp_trace.log: ~350 cycles of pipe idle time per boundary with 8 MMAs per stage).
// Single pass: a stage is a whole 256-wide layer (four K-blocks, 16 MMAs, 2.07 k cycles); only two such stages fit, so the small
// weight blocks of the input encodings (PE(xyz) in layers 0 and 5, PE(viewdir) in the view layer: 8-16 KB) travel through their
// own AUXILIARY slot instead of taking a ring stage each (they did at first: the 64 KB stage behind them then arrived ~1 k cycles
// late, twice per tile).  Split precision: a stage is the hi + lo block of one K-block (12 MMAs, 1.55 k cycles), four stages.
constexpr int stage_bytes(int npass) { return npass == 1 ? 65536 : 32768; }
constexpr int aux_bytes(int npass) { return npass == 1 ? 16384 : 0; }
constexpr int MAX_LAYERS = 9;
constexpr int MAX_STEPS = 68;
constexpr int SMEM_LIMIT = 232448;           // 227 KB
constexpr int XT_LAYER = 6;                  // the next tile's PE(xyz) is written in the idle window before this layer's accumulator is ready
constexpr int VIEW_LAYER_WRITE = 1;          // PE(viewdir) (its own operand block) is written in the idle window before this layer's accumulator is ready

struct Step {          // one weight-ring stage worth of MMAs: n_kb consecutive 64-wide K-blocks of the weight planes it holds
  uint32_t w_off;      // byte offset of this step's operand blocks in the packed buffer: [CTA0 half][CTA1 half]
  uint8_t a_kb;        // first A K-block consumed (0: PE, 1..4: hidden, 5: PE(viewdir))
  uint8_t n_kb;        // K-blocks in this stage; per block the stage holds [hi plane][lo plane] (those present)
  uint8_t k16_last;    // K=16 MMAs per pass for the last block (4, or 2 for the 27-wide view encoding); the others take 4
  uint8_t flags;       // 1: first step of its layer: fresh accumulator; 2: last step of its layer: commit the acc barrier;
                       // 4: the stage holds the HIGH weight plane: passes x_hi * w_hi (+ x_lo * w_hi in split precision);
                       // 8: the stage holds the LOW weight plane (split precision): pass x_hi * w_lo
                       // 16: the step's (small) operand block travels through the AUXILIARY slot, not through the ring
};
static_assert(sizeof(Step) == 8, "Step is packed into 8 bytes: the table lives in shared memory");
enum { STEP_FIRST = 1, STEP_LAST = 2, STEP_HI = 4, STEP_LO = 8, STEP_AUX = 16 };

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
  int32_t debug;               // bring-up (ANINERF_DEBUG_MLP, timing experiments only -- results are garbage): 1 = the producer signals `full`
                               // without copying (no weight traffic); 2 = the epilogue skips its tensor-memory stores
  int32_t density_only;        // NeRF field: stop after the trunk (TPoseHuman.calculate_alpha, tpose_nerf_network.py:241-250): sigma_out only
  unsigned long long *trace;   // bring-up: clock64 timeline of one unit tile of block 0 (null = off)
  int32_t trace_iter;          // which of block 0's tiles is traced (0 = first); slots 160+i: start of its i-th tile
};

// trace slots: 0 tile start, 1 PE done; per layer l base 8 + 16*l: +0 rows wait begin, +1 rows woke, +2 rows epilogue done
// (arrived), +7 rows published the first half of quarter 0, +3 MMA obtained it; +4 MMA starts the layer, +5 MMA has the layer's first weight stage, +6 MMA issued the layer; +8+q rows published
// quarter q of the NEXT layer's operand; +12+q MMA obtained quarter q of THIS layer's operand
#define ANI_TRACE(slot)                                                              \
  do {                                                                               \
    if (TRACE && tracing) trace_base[(slot)] = (unsigned long long)clock64();        \
  } while (0)

// write 8 consecutive K elements of `row` (one 16-byte chunk) into A chunk `chunk` (K-block chunk / 8, SWIZZLE_128B);
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
  const uint32_t off = (uint32_t)(chunk >> 3) * KB_BYTES + sw128_chunk_off(row, chunk & 7);
  *reinterpret_cast<uint4 *>(a_hi + off) = h;
  if (NPASS == 3) {
    uint4 l;
    l.x = pack_bf16_residual(x[0], x[1], h.x);
    l.y = pack_bf16_residual(x[2], x[3], h.y);
    l.z = pack_bf16_residual(x[4], x[5], h.z);
    l.w = pack_bf16_residual(x[6], x[7], h.w);
    *reinterpret_cast<uint4 *>(a_lo + off) = l;
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
constexpr int ROW_WARPS = 8;                       // two threads per row: each takes half of every 64-column quarter
constexpr int ROW_THREADS = ROW_WARPS * 32;        // 256
// Threads of a CTA: the row (epilogue) threads, then one weight-producer warp and two warps for MMA issue / TMEM alloc / relays.
constexpr int N_THREADS = ROW_THREADS + 96;
constexpr int XCHG_BYTES = TILE_M * 4 * 4;         // per-row exchange between the two column halves
constexpr int PROD_LANES = 8;                      // producer lanes share every stage's bulk copy: one thread keeps only
                                                   // ~one copy in flight (20-28 B/cycle, tools/bench_stream.cu); several
                                                   // threads overlap theirs (4 threads: 85-110 B/cycle)

template <int NPASS, bool NERF>
struct Cfg {
  static constexpr int A_KB = NERF ? 2 : 1;                       // shared-memory operand blocks: PE(xyz) [, PE(viewdir)]
  static constexpr int A_PLANE = A_KB * KB_BYTES;                 // per hi / lo plane
  static constexpr int A_TOTAL = A_PLANE * (NPASS == 3 ? 2 : 1);
  static constexpr int SCR_BYTES = NERF ? 0 : TILE_M * 16 * 4;    // blend-weight head: exchange between the row's two threads
  static constexpr int HEAD_BYTES = NERF ? 2576 : 1152;
  static constexpr int XCHG = NERF ? XCHG_BYTES : 0;             // the blend-weight head exchanges through the dead A operand
  static constexpr int STEP_BYTES = MAX_STEPS * (int)sizeof(Step);
  // (one layer's bias at a time, staged by the row threads in the idle window before the layer's accumulator is ready: with the
  // whole 9 KB table resident the split-precision blend-weight field would have three ring stages instead of four)
  static constexpr int BIAS_BYTES = 256 * 4;
  static constexpr int AUX_BYTES = aux_bytes(NPASS);
  static constexpr int FIXED = A_TOTAL + AUX_BYTES + SCR_BYTES + HEAD_BYTES + XCHG + STEP_BYTES + BIAS_BYTES + 256;
  static constexpr int STAGE_BYTES = stage_bytes(NPASS);
  static constexpr int STAGES_RAW = (SMEM_LIMIT - FIXED) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 4 ? 4 : STAGES_RAW;
  static_assert(STAGES >= 2, "weight ring needs at least two stages");
  static constexpr int TMEM_COLS = 512;                         // two accumulators, alternating by layer
  // offsets (A planes and the ring first: SWIZZLE_128B blocks need 1024-byte alignment)
  static constexpr int OFF_A_HI = 0;
  static constexpr int OFF_A_LO = A_PLANE;                      // only when NPASS == 3
  static constexpr int OFF_RING = A_TOTAL;
  static constexpr int OFF_AUX = OFF_RING + STAGES * STAGE_BYTES;
  static constexpr int OFF_SCR = OFF_AUX + AUX_BYTES;
  static constexpr int OFF_HEAD = OFF_SCR + SCR_BYTES;
  static constexpr int OFF_XCHG = OFF_HEAD + HEAD_BYTES;
  static constexpr int OFF_STEPS = OFF_XCHG + XCHG;
  static constexpr int OFF_BIAS = OFF_STEPS + STEP_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
  static constexpr int SMEM = OFF_BAR + 256;
  static_assert(SMEM <= SMEM_LIMIT, "shared memory budget");
  static_assert(OFF_RING % 1024 == 0, "ring stages must be 1024-byte aligned");
};

// TRACE: the clock64 timeline of tools/gpu_diag.py.  A separate instantiation: the predicated stamps alone cost 1.5 % (split precision)
// to 2.2 % (single pass) of the tile time when compiled in (A/B on one GPU: 61.0 k -> 60.1 k, 29.6 k -> 29.0 k cycles per tile).
template <int NPASS, bool NERF, int PAIR, bool TRACE>
__global__ void __launch_bounds__(N_THREADS, 1) mlp_kernel(const __grid_constant__ MlpArgs args) {
  constexpr int RW = ROW_WARPS;                      // row (epilogue) warps; warp RW: producer; RW+1, RW+2: MMA issue / relays
  using C = Cfg<NPASS, NERF>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *ring = smem + C::OFF_RING;
  float *s_head = reinterpret_cast<float *>(smem + C::OFF_HEAD);
  float *s_bias = reinterpret_cast<float *>(smem + C::OFF_BIAS);
  float *s_xchg = reinterpret_cast<float *>(smem + C::OFF_XCHG);
  Step *s_steps = reinterpret_cast<Step *>(smem + C::OFF_STEPS);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + C::OFF_BAR);
  // barrier slots: [0,S) full, [S,2S) empty, [2S,2S+5) a_ready (leader CTA: [q] = K-block quarter q of the hidden operand is
  // written, [4] = the first half of quarter 0 is; every epilogue warp of BOTH CTAs arrives once per phase -- the peer's warps
  // with a remote arrive, no relay), [2S+5] acc
  const uint32_t bar_full = smem_u32(bars);
  const uint32_t bar_empty = smem_u32(bars + C::STAGES);
  const uint32_t bar_a_ready = smem_u32(bars + 2 * C::STAGES);
  const uint32_t bar_acc = smem_u32(bars + 2 * C::STAGES + 5);
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * C::STAGES + 6);
  const uint32_t bar_aux_full = smem_u32(bars + 2 * C::STAGES + 7), bar_aux_empty = smem_u32(bars + 2 * C::STAGES + 8);
  static_assert((2 * C::STAGES + 9) * 8 <= 216, "barrier block overflow");
  // XT: cross-tile prefetch.  The next tile's input encoding is written into the PE(xyz) block (dead since layer 5) in layer
  // XT_LAYER's idle window and published right after the LAST layer's accumulator barrier, so the MMA issuer rolls from the last
  // layer straight into the next tile's layer 0 while the epilogue warps are still busy with this tile's head.
  constexpr bool XT = true;
  // Quarter 0 of the hidden operand is published in two halves (split precision only: there the MMAs of half a quarter, 0.8 k cycles,
  // cover the second half's conversion; in single-pass mode the epilogue is the critical path and the extra publish costs more than
  // the earlier start gains: 32.7 k -> 35.0 k cycles per tile measured).
  constexpr bool SPLIT_Q0 = NPASS == 3;
  constexpr int VCHUNK0 = VIEW_CHUNK0;
  float *s_grid = reinterpret_cast<float *>(smem + C::OFF_BAR + 224);   // lo[3], (dim-1)/ext [3] of the SMPL-weight volume

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR == 2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int n_units = (int)gridDim.x / PAIR;            // clusters walking the unit-tile list
  const int unit = (int)blockIdx.x / PAIR;
  const FieldDev &F = args.f;
  const int n_layers = (NERF && args.density_only) ? F.n_layers - 1 : F.n_layers;   // density query: the trunk + alpha only
  const int64_t n_valid = args.n_dev ? (int64_t)min((int64_t)*args.n_dev, args.n) : args.n;
  const int64_t n_tiles = (n_valid + TILE_M - 1) / TILE_M;
  const int64_t n_utiles = (n_tiles + PAIR - 1) / PAIR;

  // ---- one-time setup --------------------------------------------------------------------
  const int latent = args.latent_index + (args.latent_dev ? (int)__ldg(args.latent_dev) : 0);
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
    mbar_init(bar_aux_full, (leader && PAIR == 2) ? 2 : 1);
    mbar_init(bar_aux_empty, 1);
    for (int q = 0; q < 5; ++q) mbar_init(bar_a_ready + 8 * q, ROW_WARPS * PAIR);   // one arrival per epilogue warp of the pair
    mbar_init(bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == RW + 1) tmem_alloc<PAIR>(smem_u32(tmem_slot), C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();        // the peer's barriers exist before anything arrives on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == RW) {
    // ===== weight producer: this CTA's half of every operand block ==============================
    if (lane < PROD_LANES) {
      uint32_t stage = 0, phase = 0, aux_phase = 0;
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        for (int l = 0; l < n_layers; ++l) {
          for (int s = F.layers[l].step0; s < F.layers[l].step0 + F.layers[l].n_steps; ++s) {
            const Step ps = s_steps[s];
            const bool aux = C::AUX_BYTES > 0 && (ps.flags & STEP_AUX) != 0;
            const uint32_t fb = aux ? bar_aux_full : bar_full + 8 * stage, eb = aux ? bar_aux_empty : bar_empty + 8 * stage;
            const uint32_t dst = aux ? smem_u32(smem + C::OFF_AUX) : smem_u32(ring + stage * C::STAGE_BYTES);
            mbar_wait(eb, (aux ? aux_phase : phase) ^ 1, 1);               // all producer lanes stay in step (no divergent spin)
            const uint32_t bytes = (uint32_t)(F.layers[l].n_pad / PAIR) * 128u * ps.n_kb *
                                   (((ps.flags & STEP_HI) ? 1u : 0u) + ((ps.flags & STEP_LO) ? 1u : 0u));   // this CTA's half of the stage
            // lane 0 announces the stage's bytes, then every lane copies one eighth of it: one thread keeps only about one bulk copy
            // in flight, eight overlap theirs (every stage size is a multiple of 8 x 1 KB)
            if (lane == 0) {
              if (args.debug & 1) mbar_arrive(fb);
              else mbar_expect_tx(fb, bytes);
            }
            __syncwarp((1u << PROD_LANES) - 1u);
            if (!(args.debug & 1)) {
              const uint32_t part = bytes / PROD_LANES;
              bulk_g2s(dst + lane * part, F.image + ps.w_off + cta_rank * bytes + lane * part, part, fb);
            }
            if (aux) {
              aux_phase ^= 1;
            } else if (++stage == C::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp > RW) {
    const int role = warp - (RW + 1);
    if (leader && role == 0 && elect_one()) {
      // ===== MMA issuer (leader CTA): ONE thread.  Everything it does between two MMAs is on the critical path of the tensor pipe
      // (tools/bench_mma.cu: a few dozen dependent scalar instructions per MMA already halve the rate), so the loop is kept to:
      // per ring stage one decode + one `full` wait + one commit; per K-block one descriptor per operand and (first touch of a
      // quarter) one a_ready wait; per K=16 slice a constant +32 bytes / +8 columns. ====================================
      uint32_t stage = 0, full_phase = 0, aux_phase = 0, a_phase = 0, lc = 0;   // ring position; lc: running layer count (accumulator buffer)
      const uint32_t a_hi = smem_u32(smem + C::OFF_A_HI), a_lo = smem_u32(smem + C::OFF_A_LO);
      const uint32_t ring_u32 = smem_u32(ring);
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        const bool tracing = TRACE && args.trace && blockIdx.x == 0 && ut == (int64_t)args.trace_iter * n_units;
        unsigned long long *const trace_base = args.trace;
        for (int l = 0; l < n_layers; ++l, ++lc) {
          const int n_pad = F.layers[l].n_pad;
          const int s0 = F.layers[l].step0, n_layer_steps = F.layers[l].n_steps;
          const uint32_t idesc = instr_desc(n_pad, TILE_M * PAIR);
          const uint32_t acc = tmem_base + (lc & 1u) * 256u;
          const uint32_t a_prev = tmem_base + ((lc & 1u) ^ 1u) * 256u;      // the previous layer's buffer: this layer's hidden operand
          const uint32_t blk = (uint32_t)(n_pad / PAIR) * 128u;             // one K-block of this CTA's weight rows, one plane
          ANI_TRACE(8 + 16 * l + 4);
          // K-block q of the hidden operand is ready once a_ready[q] completes.  Layer 0 reads the input encoding only: its four
          // barriers complete together and are all consumed up front (never after the commit: the epilogue's arrivals for layer 1
          // must not overtake).  The accumulator buffer was drained two layers ago.
          int next_q = 0;
          bool got_first = false;                                 // a_ready[4]: the first half of quarter 0
          if (l == 0) {
            for (; next_q < 4; ++next_q) mbar_wait(bar_a_ready + 8 * next_q, a_phase, 2 + next_q);
            if (SPLIT_Q0) mbar_wait(bar_a_ready + 8 * 4, a_phase, 6);
            got_first = true;
          }
          uint32_t fresh = 0u;
          for (int s = s0; s < s0 + n_layer_steps; ++s) {
            const Step st = s_steps[s];
            const bool hi_stage = (st.flags & STEP_HI) != 0, lo_stage = NPASS == 3 && (st.flags & STEP_LO) != 0;
            const uint32_t per_kb = blk * ((hi_stage ? 1u : 0u) + (lo_stage ? 1u : 0u));
            const bool aux = C::AUX_BYTES > 0 && (st.flags & STEP_AUX) != 0;          // a small encoding block in the auxiliary slot
            // (TMA-written operands: CTA-scope acquire is enough for the async proxy)
            mbar_wait(aux ? bar_aux_full : bar_full + 8 * stage, aux ? aux_phase : full_phase, 3);
            tc_fence_after();
            if (s == s0) ANI_TRACE(8 + 16 * l + 5);
            uint32_t b_blk = aux ? smem_u32(smem + C::OFF_AUX) : ring_u32 + stage * (uint32_t)C::STAGE_BYTES;
            for (int kb = 0; kb < (int)st.n_kb; ++kb, b_blk += per_kb) {
              const int akb = (int)st.a_kb + kb;
              const bool hidden = akb >= KB_HID0 && akb < KB_HID0 + 4;
              const uint64_t bdh = smem_desc_sw128(b_blk);
              const uint64_t bdl = smem_desc_sw128(b_blk + (hi_stage ? blk : 0u));
              if (hidden) {
                // A from tensor memory: quarter q of the PREVIOUS layer's accumulator buffer, re-quantised in place by its epilogue.
                // Slice k (16 K elements) comes from 16 fp32 columns [64q + 16k, +16): their owner wrote the hi pairs over the first 8
                // of them and (split precision) the lo pairs over the other 8.  Quarter 0 is published in two halves -- slices
                // {0, 2} (the first 16 columns of each of the row's two threads), then {1, 3} -- so that the layer starts as early
                // as possible: the window between the previous layer's last MMA and this layer's first one is tensor-pipe idle time.
                const int q = akb - KB_HID0;
                const uint32_t ta = a_prev + (uint32_t)q * 64u;
                auto issue_slices = [&](int kfirst, int kstep) {
                  if (hi_stage) {
#pragma unroll
                    for (int k = kfirst; k < 4; k += kstep) {
                      const uint32_t tk = ta + (uint32_t)(k * 16);
                      umma_bf16_ts<PAIR>(acc, tk, bdh + (uint64_t)(2 * k), idesc, fresh);
                      fresh = 1u;
                      if (NPASS == 3) umma_bf16_ts<PAIR>(acc, tk + 8u, bdh + (uint64_t)(2 * k), idesc, 1u);
                    }
                  }
                  if (lo_stage) {
#pragma unroll
                    for (int k = kfirst; k < 4; k += kstep) {
                      umma_bf16_ts<PAIR>(acc, ta + (uint32_t)(k * 16), bdl + (uint64_t)(2 * k), idesc, fresh);
                      fresh = 1u;
                    }
                  }
                };
                if (SPLIT_Q0 && q == 0) {
                  if (!got_first) {
                    mbar_wait(bar_a_ready + 8 * 4, a_phase, 6);        // (CTA-scope acquire: a cluster-scope one invalidates the L1)
                    tc_fence_after();
                    got_first = true;
                    ANI_TRACE(8 + 16 * l + 3);
                  }
                  issue_slices(0, 2);
                  if (next_q == 0) {
                    mbar_wait(bar_a_ready, a_phase, 2);
                    tc_fence_after();
                    ANI_TRACE(8 + 16 * l + 12);
                    next_q = 1;
                  }
                  issue_slices(1, 2);
                } else {
                  for (; next_q <= q; ++next_q) {
                    mbar_wait(bar_a_ready + 8 * next_q, a_phase, 2 + next_q);
                    tc_fence_after();
                    ANI_TRACE(8 + 16 * l + 12 + next_q);
                  }
                  issue_slices(0, 1);
                }

              } else {
                // the input encodings: shared-memory operand blocks (ready since the tile's start / the previous layers)
                const bool full4 = !(kb + 1 == (int)st.n_kb && st.k16_last == 2);
                const uint32_t a_off = akb == KB_PE ? 0u : (uint32_t)KB_BYTES;
                const uint64_t adh = smem_desc_sw128(a_hi + a_off);
                const uint64_t adl = smem_desc_sw128(a_lo + a_off);
                if (hi_stage) {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    if (k < 2 || full4) {
                      umma_bf16<PAIR>(acc, adh + (uint64_t)(2 * k), bdh + (uint64_t)(2 * k), idesc, k == 0 ? fresh : 1u);
                      if (NPASS == 3) umma_bf16<PAIR>(acc, adl + (uint64_t)(2 * k), bdh + (uint64_t)(2 * k), idesc, 1u);
                    }
                  }
                  fresh = 1u;
                }
                if (lo_stage) {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < 2 || full4) umma_bf16<PAIR>(acc, adh + (uint64_t)(2 * k), bdl + (uint64_t)(2 * k), idesc, k == 0 ? fresh : 1u);
                  fresh = 1u;
                }
              }
            }
            umma_commit<PAIR>(aux ? bar_aux_empty : bar_empty + 8 * stage);   // frees the slot (both CTAs) once these MMAs retire
            if (st.flags & STEP_LAST) umma_commit<PAIR>(bar_acc);   // layer done: its accumulator is ready
            if (aux) {
              aux_phase ^= 1u;
            } else if (++stage == C::STAGES) {
              stage = 0;
              full_phase ^= 1u;
            }
          }
          ANI_TRACE(8 + 16 * l + 6);
          for (; next_q < 4; ++next_q) mbar_wait(bar_a_ready + 8 * next_q, a_phase, 7);   // keep every quarter's phase in step
          if (SPLIT_Q0 && !got_first) mbar_wait(bar_a_ready + 8 * 4, a_phase, 7);
          a_phase ^= 1;
        }
      }
    } else if (lane == 0 && PAIR == 2 && !leader && role == 0) {
      // ===== relay (peer CTA): tell the leader when this CTA's half of a stage has landed ========
      uint32_t stage = 0, phase = 0, aux_phase = 0;
      const uint32_t full_leader = mapa_u32(bar_full, 0), aux_full_leader = mapa_u32(bar_aux_full, 0);
      for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
        for (int l = 0; l < n_layers; ++l) {
          for (int s = F.layers[l].step0; s < F.layers[l].step0 + F.layers[l].n_steps; ++s) {
            if (C::AUX_BYTES > 0 && (s_steps[s].flags & STEP_AUX)) {
              mbar_wait(bar_aux_full, aux_phase, 4);
              mbar_arrive_remote(aux_full_leader);
              aux_phase ^= 1;
            } else {
              mbar_wait(bar_full + 8 * stage, phase, 4);
              mbar_arrive_remote(full_leader + 8 * stage);
              if (++stage == C::STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else {
    // ===== row threads: input encoding, per-layer epilogues, heads =============================
    const int row = (warp & 3) * 32 + lane;            // == TMEM lane (warp % 4 selects the lane quadrant); warps w and w+4 share a row
    const int half = warp >> 2;                        // which half of every 64-column quarter this thread owns
    const uint32_t t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    // One arrival per warp and quarter, straight on the leader's barrier (the peer's warps through the cluster's shared-memory
    // window): 16 arrivals per phase instead of 2 x 256 + a relay hop.
    const uint32_t a_arrive = leader ? bar_a_ready : mapa_u32(bar_a_ready, 0);
    auto arrive_warp = [&](int q) {          // every lane has fenced its own writes; lane 0 signals for the warp
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(a_arrive + 8 * q);
        else mbar_arrive_remote(a_arrive + 8 * q);
      }
    };
    uint8_t *a_hi = smem + C::OFF_A_HI, *a_lo = smem + C::OFF_A_LO;
    uint32_t acc_phase = 0, lc = 0;
    const bool debug_no_st = (args.debug & 2) != 0;
    auto load_bias = [&](int l) {                        // this thread's entry of layer l's bias row (latent-code tables: row `latent`)
      const int rt = (int)threadIdx.x;
      const float *gb = F.bias + F.layers[l].bias_off + min(max(latent, 0), F.layers[l].n_tables - 1) * F.layers[l].n_out;
      return rt < F.layers[l].n_out ? __ldg(gb + rt) : 0.f;
    };
    float bias_next = load_bias(0);
    bool pe_ready = false;                       // XT: this tile's encoding was written during the previous tile's last layer
    int64_t ngi = 0;
    bool nvalid = false;
    float nx = 0.f, ny = 0.f, nz = 0.f;
    for (int64_t ut = unit; ut < n_utiles; ut += n_units) {
      const bool tracing = TRACE && args.trace && blockIdx.x < 2 && ut == (int64_t)args.trace_iter * n_units && threadIdx.x == 0;
      unsigned long long *const trace_base = args.trace + (blockIdx.x == 1 ? 256 : 0);     // the peer CTA of cluster 0: slots 256.. (its own clock)
      if (TRACE && args.trace && blockIdx.x == 0 && threadIdx.x == 0 && ut / n_units < 90) args.trace[160 + ut / n_units] = (unsigned long long)clock64();
      int64_t gi;
      bool valid;
      float px, py, pz, sigma = 0.f;
      float smpl[ANINERF_N_BONES / 2];     // blend-weight head: the row's initial SMPL weights, 12 of the 24 bones per thread
      ANI_TRACE(0);
      if (XT && pe_ready) {                          // encoded (and published) during the previous tile's last layer
        gi = ngi;
        valid = nvalid;
        px = nx;
        py = ny;
        pz = nz;
      } else {
        gi = (ut * PAIR + cta_rank) * TILE_M + row;
        valid = gi < n_valid;
        px = py = pz = 0.f;
        if (valid) {
          px = __ldg(args.pts + 3 * gi);
          py = __ldg(args.pts + 3 * gi + 1);
          pz = __ldg(args.pts + 3 * gi + 2);
        }
        if (half == 0) write_pe<NPASS, 10, 0, 4>(a_hi, a_lo, PE_CHUNK0, row, px, py, pz);
        else write_pe<NPASS, 10, 4, 8>(a_hi, a_lo, PE_CHUNK0, row, px, py, pz);
        fence_proxy_async();
#pragma unroll
        for (int q = 0; q < (SPLIT_Q0 ? 5 : 4); ++q) arrive_warp(q);   // the input encoding readies every quarter's barrier for layer 0
      }
      ANI_TRACE(1);

      for (int l = 0; l < n_layers; ++l, ++lc) {
        const bool last = l == n_layers - 1;
        const int n_pad = F.layers[l].n_pad;
        {
          // this layer's bias row (fetched after the previous layer's last publish) -> shared memory (the previous layer's readers
          // are done: first barrier)
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
          s_bias[threadIdx.x] = bias_next;                     // row threads are threads 0..255
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
        }
        const float *bias = s_bias;
        float *xchg = s_xchg;                                // exchange between the row's two threads (NeRF field)
        const uint32_t t_acc = t_lane + (lc & 1u) * 256u;
        if (!NERF && l >= 1) {
          // Initial SMPL weights of this row (this thread: bones 12*half .. +11): ONE trilinear corner per layer (two in layer 1's
          // window), fetched in the idle window before the layer's accumulator is ready and accumulated in registers (ATen's
          // corner order).  The whole gather is 98 KB per tile from L2, which also feeds the 1 MB weight stream:
          // done in one burst it ran at the L2 bandwidth roof for 7-8k cycles and delayed the layer it shared the window with;
          // 12 KB per window hides.  Not in layer 0's window: with the cross-tile prefetch that layer's MMAs are done before the
          // tile starts, and the gather's L2 latency (~2 k cycles) sat on the critical path.
          if (l == 1) {
            ANI_TRACE(4);
#pragma unroll
            for (int k = 0; k < ANINERF_N_BONES / 2; ++k) smpl[k] = 0.f;
          }
          if (valid) {
            if (args.smpl_bw) {
              if (l == 7) {
                const float4 *r4 = reinterpret_cast<const float4 *>(args.smpl_bw + gi * ANINERF_N_BONES) + half * 3;
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                  float4 w4 = __ldg(r4 + q);
                  smpl[4 * q] = w4.x;
                  smpl[4 * q + 1] = w4.y;
                  smpl[4 * q + 2] = w4.z;
                  smpl[4 * q + 3] = w4.w;
                }
              }
            } else if (l < 8) {
              // window of layer 1: corners 0 and 1; layers 2..7: corner l (the last layer's window is short and holds the logs)
              for (int corner = l == 1 ? 0 : l; corner <= l; ++corner) {
                float wc;
                int oc;
                fast_corner(s_grid, args.grid_dim, px, py, pz, corner, wc, oc);
                const float4 *r4 = reinterpret_cast<const float4 *>(args.vol_w24 + (int64_t)oc * ANINERF_N_BONES) + half * 3;
                float4 c4[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) c4[q] = __ldg(r4 + q);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                  smpl[4 * q] = fmaf(c4[q].x, wc, smpl[4 * q]);
                  smpl[4 * q + 1] = fmaf(c4[q].y, wc, smpl[4 * q + 1]);
                  smpl[4 * q + 2] = fmaf(c4[q].z, wc, smpl[4 * q + 2]);
                  smpl[4 * q + 3] = fmaf(c4[q].w, wc, smpl[4 * q + 3]);
                }
              }
            }
          }
          if (l == 8) {
            // the head's log(smpl_bw + 1e-9), still inside the last layer's window (12 logf per thread: ~1 k cycles off the tile's tail)
#pragma unroll
            for (int k = 0; k < ANINERF_N_BONES / 2; ++k) smpl[k] = logf(smpl[k] + 1e-9f);
            ANI_TRACE(5);
          }
        }
        if (NERF && l == VIEW_LAYER_WRITE && !args.density_only) {
          // PE(viewdir) -> its own operand block, while this layer's MMAs run.  Its last reader was the PREVIOUS tile's view layer,
          // complete since that tile's last accumulator barrier; its next reader is issued after this tile's later publishes.
          float vx = 0.f, vy = 0.f, vz = 0.f;
          if (valid) {
            vx = __ldg(args.viewdir + 3 * gi);
            vy = __ldg(args.viewdir + 3 * gi + 1);
            vz = __ldg(args.viewdir + 3 * gi + 2);
          }
          if (half == 0) write_pe<NPASS, 4, 0, 2>(a_hi, a_lo, VCHUNK0, row, vx, vy, vz);
          else write_pe<NPASS, 4, 2, 4>(a_hi, a_lo, VCHUNK0, row, vx, vy, vz);
          fence_proxy_async();           // (shared-memory operand: visible to the tensor core before the next publishes)
        }
        if (XT && l == XT_LAYER) {
          // The NEXT tile's input encoding -> the PE(xyz) block, in the idle window of this layer: the block's last reader (the skip
          // layer 5) completed before this thread passed that layer's accumulator barrier.  (Written after the LAST layer's barrier,
          // as at first, its ~1.8 k cycles sat on the tile's critical path: barrier -> encoding -> head -> next tile's epilogue.)
          pe_ready = ut + n_units < n_utiles;
          if (pe_ready) {
            ngi = ((ut + n_units) * PAIR + cta_rank) * TILE_M + row;
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
          }
        }
        ANI_TRACE(8 + 16 * l);
        mbar_wait(bar_acc, acc_phase, 5 + 100 * l);
        acc_phase ^= 1;
        tc_fence_after();
        if (XT && last && pe_ready) {
          // The next tile's input encoding (written in layer XT_LAYER's window) is PUBLISHED here.  That must come AFTER this layer's
          // accumulator barrier: an mbarrier arrival is not tagged with a phase, and only the completed MMAs of the last layer prove
          // that every thread's arrivals on a_ready[] for it are in.
#pragma unroll
          for (int q = 0; q < (SPLIT_Q0 ? 5 : 4); ++q) arrive_warp(q);
        }
        ANI_TRACE(8 + 16 * l + 1);
        if (last) bias_next = load_bias(0);                    // the next tile's first layer (in flight during the head)
        const bool alpha_layer = NERF && l == 7;             // the trunk's last layer: alpha_fc is an fp32 dot in its epilogue
        // (the alpha variant is a separate instantiation of the epilogue: as a run-time branch inside every 8-column group it cost
        // a BSSY / BSYNC pair, a not-taken jump over ~35 instructions and instruction-fetch stalls per group -- ncu: `no_inst`,
        // `branch_resolving` on the publish path of the single-pass NeRF field)
        auto hidden_epilogue = [&](auto alpha_tag) {
          constexpr bool ALPHA = decltype(alpha_tag)::value;
          // hidden layer: bias + ReLU -> bf16 pairs (hi / lo) -> written with tcgen05.st over the accumulator columns just read.
          // This thread owns 32 of every quarter's 64 columns; each run of 16 fp32 columns [c, c+16) (= one K=16 slice of the next
          // layer's operand) becomes hi pairs in [c, c+8) and (split precision) lo pairs in [c+8, c+16).  Quarters are published
          // one at a time (= one K-block of the next layer), quarter 0 in two halves; software-pipelined: the TMEM load of the
          // next columns is in flight while the current ones are converted, stored and published.
          uint32_t va[32], vb[32];
          // signal barrier b (0..3: quarter b complete; 4: first half of quarter 0) once every store issued so far has landed
          auto publish = [&](int b) {
            tmem_st_wait();
            tc_fence_before();
            arrive_warp(b);
            ANI_TRACE(8 + 16 * l + (b < 4 ? 8 + b : 7));
          };
          // (measured, kept out: publishing a run only after the NEXT run's conversion, to cover the store latency, made the first
          // publish of a layer 200 cycles later and the whole tile 1-2 % slower)
          auto process16 = [&](const uint32_t (&v)[32], int off, int q, int publish_after) {
            const int col0 = q * 64 + half * 32 + off;         // first fp32 column of the run
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int j4 = 0; j4 < 2; ++j4) {
              const float4 b0 = *reinterpret_cast<const float4 *>(bias + col0 + j4 * 8);
              const float4 b1 = *reinterpret_cast<const float4 *>(bias + col0 + j4 * 8 + 4);
              const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
              float x[8];
              if (NPASS == 1 && !ALPHA) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[off + j4 * 8 + j]) + bb[j];
#pragma unroll
                for (int j = 0; j < 4; ++j) hi[j4 * 4 + j] = pack_bf16_relu(x[2 * j], x[2 * j + 1]);      // the ReLU rides on the conversion
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = fmaxf(__uint_as_float(v[off + j4 * 8 + j]) + bb[j], 0.f);
                if (ALPHA) {
                  const float4 w0 = *reinterpret_cast<const float4 *>(s_head + col0 + j4 * 8);
                  const float4 w1 = *reinterpret_cast<const float4 *>(s_head + col0 + j4 * 8 + 4);
                  // (four independent chains: one dependent chain of 128 FMAs per layer is 0.5 k cycles of latency on the publish path)
                  const float p0 = fmaf(x[2], w0.z, x[0] * w0.x), p1 = fmaf(x[3], w0.w, x[1] * w0.y);
                  const float p2 = fmaf(x[6], w1.z, x[4] * w1.x), p3 = fmaf(x[7], w1.w, x[5] * w1.y);
                  sigma += (p0 + p1) + (p2 + p3);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  hi[j4 * 4 + j] = pack_bf16(x[2 * j], x[2 * j + 1]);
                  if (NPASS == 3) lo[j4 * 4 + j] = pack_bf16_residual(x[2 * j], x[2 * j + 1], hi[j4 * 4 + j]);
                }
              }
            }
            if (!debug_no_st) {
              tmem_st8(t_acc + col0, hi);
              if (NPASS == 3) tmem_st8(t_acc + col0 + 8, lo);
            }
            if (publish_after >= 0) publish(publish_after);
          };
          // (publish_after: 0..3: quarter complete with this run; 4: first half of quarter 0; -1: nothing)
          if (SPLIT_Q0) {
            tmem_ld16<0>(t_acc + half * 32, va);
            tmem_ld_wait();
            tmem_ld16<16>(t_acc + half * 32 + 16, va);
            process16(va, 0, 0, 4);
          } else {
            tmem_ld32(t_acc + half * 32, va);
            tmem_ld_wait();
            process16(va, 0, 0, -1);
          }
          tmem_ld_wait();
          tmem_ld32(t_acc + 64 + half * 32, vb);
          process16(va, 16, 0, 0);
          tmem_ld_wait();
          tmem_ld32(t_acc + 128 + half * 32, va);
          process16(vb, 0, 1, -1);
          process16(vb, 16, 1, 1);
          tmem_ld_wait();
          tmem_ld32(t_acc + 192 + half * 32, vb);
          process16(va, 0, 2, -1);
          process16(va, 16, 2, 2);
          tmem_ld_wait();
          process16(vb, 0, 3, -1);
          process16(vb, 16, 3, -1);
          if (ALPHA && half == 1) xchg[row * 4] = sigma;   // read by the row's other thread after the next acc barrier (ordered by the publish)
          publish(3);
          ANI_TRACE(8 + 16 * l + 2);
          bias_next = load_bias(l + 1);                        // (off the publish path; consumed at the next layer's start)
        };
        if (!last) {
          if constexpr (NERF) {
            if (alpha_layer) hidden_epilogue(std::true_type{});
            else hidden_epilogue(std::false_type{});
          } else {
            hidden_epilogue(std::false_type{});
          }
        } else if (!NERF) {
          // ---- blend-weight head: softmax(log(smpl_bw + 1e-9) + delta), fused inverse LBS --------
          // The row's two threads take 12 bones each and meet three times (max, sum, skinning matrix) through a scratch
          // area in the A operand, which is dead until the next tile's layer-0 epilogue.
          uint32_t v[32];
          tmem_ld32(t_acc, v);
          tmem_ld_wait();
          tc_fence_before();
          float *scr = reinterpret_cast<float *>(smem + C::OFF_SCR);   // exchange scratch of the row's two threads
          constexpr int HB = ANINERF_N_BONES / 2;
          const int k0 = half * HB;
          float bw[HB];
          float mx = -INFINITY;
#pragma unroll
          for (int k = 0; k < HB; ++k) {
            const float d = __uint_as_float(half ? v[HB + k] : v[k]);
            bw[k] = smpl[k] + (d + bias[k0 + k]);             // smpl[]: already log(smpl_bw + 1e-9)
            mx = fmaxf(mx, bw[k]);
          }
          scr[half * TILE_M + row] = mx;
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
          mx = fmaxf(scr[row], scr[TILE_M + row]);
          float part = 0.f;
#pragma unroll
          for (int k = 0; k < HB; ++k) {
            bw[k] = expf(bw[k] - mx);
            part += bw[k];
          }
          scr[(2 + half) * TILE_M + row] = part;
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
          const float inv_sum = 1.0f / (scr[2 * TILE_M + row] + scr[3 * TILE_M + row]);
#pragma unroll
          for (int k = 0; k < HB; ++k) bw[k] *= inv_sum;
          if (valid && args.bw_out) {
            float4 *o4 = reinterpret_cast<float4 *>(args.bw_out + gi * ANINERF_N_BONES) + half * 3;
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
            asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
            if (half == 0 && valid) {
#pragma unroll
              for (int j = 0; j < 12; ++j) M[j] += scr[(4 + j) * TILE_M + row];
              float qx = px - M[3], qy = py - M[7], qz = pz - M[11];
              float a = M[0], b = M[1], c = M[2], d = M[4], e = M[5], f = M[6], g = M[8], h = M[9], kk = M[10];
              float c00 = e * kk - f * h, c01 = c * h - b * kk, c02 = b * f - c * e;
              float c10 = f * g - d * kk, c11 = a * kk - c * g, c12 = c * d - a * f;
              float c20 = d * h - e * g, c21 = b * g - a * h, c22 = a * e - b * d;
              float inv = 1.0f / (a * c00 + b * c10 + c * c20);
              args.tpts_out[3 * gi] = (c00 * qx + c01 * qy + c02 * qz) * inv;
              args.tpts_out[3 * gi + 1] = (c10 * qx + c11 * qy + c12 * qz) * inv;
              args.tpts_out[3 * gi + 2] = (c20 * qx + c21 * qy + c22 * qz) * inv;
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");   // the scratch is rewritten by the next tile's head
        } else if (args.density_only) {
          // ---- density query (TPoseHuman.calculate_alpha): the trunk's last layer, alpha_fc as an fp32 dot; no colour branch ----
          float sg = 0.f;
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            uint32_t v[32];
            tmem_ld32(t_acc + q * 64 + half * 32, v);
            tmem_ld_wait();
            const int col0 = q * 64 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {            // same association as the full path's layer-7 epilogue
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = fmaxf(__uint_as_float(v[j + i]) + bias[col0 + j + i], 0.f);
              const float *w = s_head + col0 + j;
              const float p0 = fmaf(x[2], w[2], x[0] * w[0]), p1 = fmaf(x[3], w[3], x[1] * w[1]);
              const float p2 = fmaf(x[6], w[6], x[4] * w[4]), p3 = fmaf(x[7], w[7], x[5] * w[5]);
              sg += (p0 + p1) + (p2 + p3);
            }
          }
          tc_fence_before();
          if (half == 1) xchg[row * 4] = sg;
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
          if (half == 0 && valid && args.sigma_out) args.sigma_out[gi] = sg + xchg[row * 4] + s_head[256];
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");
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
          float my_sigma = sigma;
          if (half == 1) {
            xchg[row * 4 + 1] = rgb[0];
            xchg[row * 4 + 2] = rgb[1];
            xchg[row * 4 + 3] = rgb[2];
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");   // the row's two threads meet
          if (half == 0) {
            my_sigma += xchg[row * 4] + s_head[256];
            rgb[0] += xchg[row * 4 + 1] + s_head[257 + 384];
            rgb[1] += xchg[row * 4 + 2] + s_head[257 + 385];
            rgb[2] += xchg[row * 4 + 3] + s_head[257 + 386];
            if (valid) {
              const int64_t o = gi;
              if (args.sigma_out) args.sigma_out[o] = my_sigma;
              if (args.rgb_out) {
                args.rgb_out[3 * o] = rgb[0];
                args.rgb_out[3 * o + 1] = rgb[1];
                args.rgb_out[3 * o + 2] = rgb[2];
              }
              if (args.raw_out) {
                // tail of Network.forward (tpose_nerf_network.py:186-212)
                bool inside = px > args.tbounds[0] && px < args.tbounds[3] && py > args.tbounds[1] && py < args.tbounds[4] &&
                              pz > args.tbounds[2] && pz < args.tbounds[5];
                float sg = inside ? my_sigma : 0.f;
                if (args.sigma_masked_out) args.sigma_masked_out[o] = sg;
                float al = 1.0f - expf(-fmaxf(sg, 0.f) * __ldg(args.dists + o));
                float4 rv = make_float4(1.0f / (1.0f + expf(-rgb[0])), 1.0f / (1.0f + expf(-rgb[1])), 1.0f / (1.0f + expf(-rgb[2])), al);
                reinterpret_cast<float4 *>(args.raw_out)[args.index ? (int64_t)__ldg(args.index + o) : o] = rv;   // dense scatter, or compact rows
              }
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory");   // s_xchg is rewritten by the next tile
        }
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

// source columns [src0, src0+len) of a layer's weight matrix, zero-padded to whole 64-wide K-blocks, multiplied against the
// A K-blocks starting at `a_kb0`; `k16_last`: K=16 MMAs issued for the last block (4, or 2 for the 27-wide view encoding)
struct Segment { int src0, len, a_kb0, k16_last; };

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
    // (the ENCODING's K-block first wherever a layer also reads one: those MMAs need nothing from the previous layer's epilogue, so
    // they run in the window between its accumulator barrier and its first published quarter, when the tensor pipe is otherwise idle)
    if (l == 0) {
      want_k = 63; want_n = 256; h.segs = {{0, 63, 0, 4}};
    } else if (l == 5) {     // skip layer: [PE(xyz), hidden]
      want_k = 63 + 256; want_n = 256; h.segs = {{0, 63, 0, 4}, {63, 256, 1, 4}};
    } else if (l < 8) {
      want_k = 256; want_n = 256; h.segs = {{0, 256, 1, 4}};
    } else if (nerf) {       // folded view layer: [hidden, PE(viewdir)]; PE(viewdir) K-block: see VIEW_CHUNK0
      want_k = 256 + 27; want_n = 128; h.segs = {{256, 27, -1, 2}, {0, 256, 1, 4}};
    } else {
      want_k = 256; want_n = ANINERF_N_BONES; h.segs = {{0, 256, 1, 4}};
    }
    if (h.k_in != want_k || h.n_out != want_n || !L[l].W || !L[l].bias_table || h.n_tables < 1)
      return fail(ANINERF_EINVAL, "%s: layer shape does not match the aninerf architecture%s", "aninerf_net_load_field");
    h.n_pad = (h.n_out + 31) / 32 * 32;
    out.push_back(h);
  }
  return ANINERF_OK;
}

static int build_image(const aninerf_layer *L, const std::vector<HostLayer> &H, int npass, bool nerf, FieldImage &im, cudaStream_t st) {
  std::vector<Step> steps;
  std::vector<uint8_t> image;
  const int view_kb = KB_VIEW;
  (void)nerf;
  const int planes = npass == 3 ? 2 : 1;
  for (size_t l = 0; l < H.size(); ++l) {
    const HostLayer &h = H[l];
    const int n_half = h.n_pad / kPair;                     // weight rows held by each CTA of the pair
    const size_t blk = (size_t)n_half * 128;                // one SWIZZLE_128B K-block (64 K elements) of one CTA's rows, one plane
    const int cap = (int)(stage_bytes(npass) / blk);               // plane-blocks per ring stage
    // A stage holds whole K-blocks with all their planes when they fit (narrow layers: the 24-wide head of the blend-weight field is
    // ONE stage), else one plane of one K-block (256-wide layers in split precision: the hi and the lo block travel separately)
    const bool split_planes = cap < planes;
    const int kb_per_step = split_planes ? 1 : cap / planes;
    im.step0[l] = (int)steps.size();
    for (size_t si = 0; si < h.segs.size(); ++si) {
      const Segment &sg = h.segs[si];
      const int n_kb = (sg.len + 63) / 64;
      const int a_kb0 = sg.a_kb0 < 0 ? view_kb : sg.a_kb0;
      for (int kb0 = 0; kb0 < n_kb; kb0 += kb_per_step) {
        const int nkb = std::min(kb_per_step, n_kb - kb0);
        for (int part = 0; part < (split_planes ? planes : 1); ++part) {
          Step s;
          memset(&s, 0, sizeof(s));
          const int p0 = split_planes ? part : 0, p1 = split_planes ? part + 1 : planes;       // planes in this stage
          const size_t cta_bytes = blk * (size_t)(p1 - p0) * nkb;
          s.w_off = (uint32_t)image.size();
          const uint32_t step_bytes = (uint32_t)(cta_bytes * kPair);
          if (cta_bytes > (size_t)stage_bytes(npass) || (cta_bytes & 1023u)) return fail(ANINERF_EINVAL, "%s: internal: bad step size%s", __func__);
          s.a_kb = (uint8_t)(a_kb0 + kb0);
          s.n_kb = (uint8_t)nkb;
          s.k16_last = (uint8_t)(kb0 + nkb == n_kb ? sg.k16_last : 4);
          const bool first = si == 0 && kb0 == 0 && part == 0;
          const bool last = si + 1 == h.segs.size() && kb0 + nkb == n_kb && part + 1 == (split_planes ? planes : 1);
          s.flags = (uint8_t)((first ? STEP_FIRST : 0) | (last ? STEP_LAST : 0) | (p0 == 0 ? STEP_HI : 0) | (p1 == 2 ? STEP_LO : 0));
          // single pass: the small blocks of the input encodings go through the auxiliary slot (see stage_bytes)
          if (npass == 1 && (a_kb0 == KB_PE || a_kb0 == KB_VIEW) && cta_bytes <= (size_t)aux_bytes(npass)) s.flags |= STEP_AUX;
          image.resize(image.size() + step_bytes, 0);
          for (int r = 0; r < kPair; ++r)             // image = [CTA0 rows][CTA1 rows]; per K-block: [hi plane][lo plane]
            for (int kb = 0; kb < nkb; ++kb)
              for (int pl = p0; pl < p1; ++pl) {
                uint8_t *dst = image.data() + s.w_off + r * cta_bytes + ((size_t)kb * (p1 - p0) + (pl - p0)) * blk;
                for (int nn = 0; nn < n_half; ++nn)
                  for (int k = 0; k < 64; ++k) {
                    const int n = r * n_half + nn;
                    const int ks = (kb0 + kb) * 64 + k;               // K index inside the segment
                    float w = 0.f;
                    if (n < h.n_out && ks < sg.len) w = L[l].W[(size_t)n * h.k_in + sg.src0 + ks];
                    const uint16_t wh = f2bf(w);
                    const uint16_t v = pl == 0 ? wh : f2bf(w - bf2f(wh));
                    *reinterpret_cast<uint16_t *>(dst + sw128_chunk_off(nn, k >> 3) + (k & 7) * 2) = v;
                  }
              }
          steps.push_back(s);
        }
      }
    }
    im.n_layer_steps[l] = (int)steps.size() - im.step0[l];
  }
  if (steps.size() > (size_t)MAX_STEPS) return fail(ANINERF_EINVAL, "%s: internal: too many steps%s", __func__);
  ANI_CUDA(cudaMalloc(&im.image, image.size()));
  ANI_CUDA(cudaMalloc(&im.steps, steps.size() * sizeof(Step)));
  ANI_CUDA(cudaMemcpyAsync(im.image, image.data(), image.size(), cudaMemcpyHostToDevice, st));
  ANI_CUDA(cudaMemcpyAsync(im.steps, steps.data(), steps.size() * sizeof(Step), cudaMemcpyHostToDevice, st));
  ANI_CUDA(cudaStreamSynchronize(st));   // the std::vectors go out of scope
  im.n_steps = (int)steps.size();
  return ANINERF_OK;
}

template <int NPASS, bool NERF, bool TRACE = false>
static int launch_mlp(const MlpArgs &a, cudaStream_t st) {
  using C = Cfg<NPASS, NERF>;
  if (!TRACE && a.trace) return launch_mlp<NPASS, NERF, true>(a, st);      // diagnostics build of the same kernel
  static bool configured = false;
  if (!configured) {
    ANI_CUDA(cudaFuncSetAttribute(mlp_kernel<NPASS, NERF, kPair, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    configured = true;
  }
  const int64_t rows_per_unit = (int64_t)TILE_M * kPair;
  const int64_t utiles = (a.n + rows_per_unit - 1) / rows_per_unit;
  const int units = (int)std::min<int64_t>(utiles, sm_count() / kPair);
  if (units <= 0) return ANINERF_OK;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(units * kPair));
  cfg.blockDim = dim3(N_THREADS);
  cfg.dynamicSmemBytes = C::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kPair;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ANI_CUDA(cudaLaunchKernelEx(&cfg, mlp_kernel<NPASS, NERF, kPair, TRACE>, a));
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
  {
    static const char *dbg = getenv("ANINERF_DEBUG_MLP");
    a.debug = dbg ? atoi(dbg) : 0;
  }
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
  return precision == 3 ? launch_mlp<3, false>(a, st) : launch_mlp<1, false>(a, st);
}

int nerf_forward_impl(aninerf_net *net, int latent_index, const int64_t *latent_dev, const float *pts, const float *viewdir, int64_t n, const int32_t *n_dev,
                      float *sigma_out, float *rgb_out, const float *dists, const float *tbounds, const int32_t *index, float *raw_out,
                      float *sigma_masked_out, int precision, cudaStream_t st, int density_only) {
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
  a.density_only = density_only;
  return precision == 3 ? launch_mlp<3, true>(a, st) : launch_mlp<1, true>(a, st);
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
  rc = build_image(layers, H, 1, nerf, f.img[0], st);
  if (rc) return rc;
  rc = build_image(layers, H, 3, nerf, f.img[1], st);
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
                           precision, (cudaStream_t)stream, 0);
}

}  // extern "C"
