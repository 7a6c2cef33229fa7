// Split-precision tcgen05 GEMM of the training step: the forward, data-gradient and weight-gradient
// products of every 1x1 Conv1d of the two fields (tpose_nerf_network.py:12-38, 219-239; autograd of
// F.conv1d in the reference) as ONE kernel,
//
//     C[M,N] = epilogue( sum_s  A_s[M,K_s] * B_s[N,K_s]^T ),      fp32 in, fp32 out,
//
// with every product evaluated as bf16x3 (x_hi*w_hi + x_lo*w_hi + x_hi*w_lo, fp32 accumulation in
// TMEM): fp32-equivalent results from the bf16 tensor pipe, so the training step matches the fp32
// reference to ~1e-6 relative instead of the ~1e-2 of plain bf16 mixed precision.
//
// Operands are addressed through (row stride, k stride) pairs, so X, X^T, W and W^T are all read in
// place; up to two K segments implement the skip layer's concat ([PE, hidden] @ [W[:, :63], W[:, 63:]])
// and the view layer without materialising the concatenation.  A CTA owns a 128 x BN output tile:
// all 256 threads stage 32-wide K chunks (global fp32 -> bf16 hi/lo -> shared memory in the UMMA
// no-swizzle K-major core-matrix layout [K/8][rows][8]) into a 3-stage ring, one thread issues the
// tcgen05.mma's of a staged chunk and commits them to the stage's mbarrier, and the staging of the
// next chunks overlaps those MMAs.  Epilogue: tcgen05.ld -> (+bias) -> (+C) -> ReLU / ReLU-mask ->
// store.  Weight gradients reduce over the samples: split-K over blockIdx.z into a partial buffer
// and a fixed-order reduction kernel (deterministic, no atomics).
#include "common.cuh"
#include "tcgen05.cuh"

namespace aninerf {

constexpr int G_BM = 128;          // rows of the output tile = TMEM lanes
constexpr int G_KC = 64;           // K elements per stage (12 MMAs per __syncthreads: the small products of a training step are latency-bound)
constexpr int G_STAGES = 3;        // (2 for the 256-wide tile: 96 KB per stage)
constexpr int G_THREADS = 256;

struct GemmDev {
  const float *A[2], *B[2];
  long long a_rs[2], a_ks[2], b_rs[2], b_ks[2];
  int K[2];
  int n_seg;
  int M, N;
  float *C;
  long long ldc;
  const float *bias;        // (N,) added per output column, or null
  const float *mask;        // (M,N) with leading dimension ldm: result *= (mask > 0), or null
  long long ldm;
  int relu;
  int accumulate;           // result += C (before the mask)
  int k_per_split;          // split-K (gridDim.z > 1, one segment): this CTA reduces [z*k_per_split, (z+1)*k_per_split)
  float *partial;           // (gridDim.z, M, N) when split
};

template <int BN>
struct GemmSmem {
  static constexpr int A_STAGE = G_BM * G_KC * 2;       // bytes of one hi (or lo) plane
  static constexpr int B_STAGE = BN * G_KC * 2;
  static constexpr int STAGE = 2 * (A_STAGE + B_STAGE);
  static constexpr int STAGES = BN > 128 ? 2 : G_STAGES;
  static constexpr int BYTES = STAGES * STAGE + 64;
  static_assert(BYTES <= 232448, "shared memory budget");
};

// One K chunk of an operand -- rows [r0, r0+ROWS) x k [k0, k0+32) of X (zero outside [0,R) x [0,K)) -- in two steps, so that the
// global loads of chunk c+1 are in flight while chunk c is converted, published and multiplied:
//   load():  global fp32 -> registers;   store(): registers -> bf16 hi/lo -> shared memory (UMMA no-swizzle K-major core matrices)
template <int ROWS>
struct OperandChunk {
  static constexpr int TASKS = ROWS * (G_KC / 8);
  static constexpr int PER_THREAD = (TASKS + G_THREADS - 1) / G_THREADS;
  float x[PER_THREAD][8];

  __device__ __forceinline__ void load(const float *X, long long rs, long long ks, int r0, int R, int k0, int K) {
    const bool row_contig = rs == 1 && ks != 1;   // consecutive threads walk the contiguous direction
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int t = threadIdx.x + i * G_THREADS;
      if (TASKS % G_THREADS != 0 && t >= TASKS) break;
      const int row = row_contig ? t % ROWS : t / (G_KC / 8);
      const int kg = row_contig ? t / ROWS : t % (G_KC / 8);
      const int r = r0 + row, k = k0 + kg * 8;
      const float *src = X + (long long)r * rs + (long long)k * ks;
      if (ks == 1 && r < R && k + 8 <= K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        // K-contiguous, 16-byte aligned, fully inside: two float4 loads
        const float4 v0 = __ldg(reinterpret_cast<const float4 *>(src)), v1 = __ldg(reinterpret_cast<const float4 *>(src) + 1);
        x[i][0] = v0.x; x[i][1] = v0.y; x[i][2] = v0.z; x[i][3] = v0.w;
        x[i][4] = v1.x; x[i][5] = v1.y; x[i][6] = v1.z; x[i][7] = v1.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[i][j] = (r < R && k + j < K) ? __ldg(src + (long long)j * ks) : 0.f;
      }
    }
  }

  __device__ __forceinline__ void store(uint8_t *hi, uint8_t *lo, long long rs, long long ks) const {
    const bool row_contig = rs == 1 && ks != 1;
#pragma unroll
    for (int i = 0; i < PER_THREAD; ++i) {
      const int t = threadIdx.x + i * G_THREADS;
      if (TASKS % G_THREADS != 0 && t >= TASKS) break;
      const int row = row_contig ? t % ROWS : t / (G_KC / 8);
      const int kg = row_contig ? t / ROWS : t % (G_KC / 8);
      uint4 h, l;
      h.x = pack_bf16(x[i][0], x[i][1]);
      h.y = pack_bf16(x[i][2], x[i][3]);
      h.z = pack_bf16(x[i][4], x[i][5]);
      h.w = pack_bf16(x[i][6], x[i][7]);
      l.x = pack_bf16_residual(x[i][0], x[i][1], h.x);
      l.y = pack_bf16_residual(x[i][2], x[i][3], h.y);
      l.z = pack_bf16_residual(x[i][4], x[i][5], h.z);
      l.w = pack_bf16_residual(x[i][6], x[i][7], h.w);
      const int off = kg * (ROWS * 16) + row * 16;
      *reinterpret_cast<uint4 *>(hi + off) = h;
      *reinterpret_cast<uint4 *>(lo + off) = l;
    }
  }
};

template <int BN>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_x3_kernel(const __grid_constant__ GemmDev g) {
  using S = GemmSmem<BN>;
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int NS = S::STAGES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + NS * S::STAGE);   // [0,STAGES): stage consumed; [STAGES]: accumulator complete
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + NS + 1);
  const uint32_t bar0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * G_BM, n0 = blockIdx.y * BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s <= NS; ++s) mbar_init(bar0 + 8 * s, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<1>(smem_u32(tmem_slot), BN < 32 ? 32 : BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = instr_desc(BN, G_BM);

  // K range of this CTA (split-K uses segment 0 only)
  int k_lo = 0, k_hi = g.K[0];
  if (gridDim.z > 1) {
    k_lo = blockIdx.z * g.k_per_split;
    k_hi = min(g.K[0], k_lo + g.k_per_split);
  }
  // the chunk list: segment 0 over [k_lo, k_hi), then segment 1 over [0, K1)
  const int n0_chunks = (max(k_hi - k_lo, 0) + G_KC - 1) / G_KC;
  const int n1_chunks = g.n_seg > 1 ? (g.K[1] + G_KC - 1) / G_KC : 0;
  const int n_chunks = n0_chunks + n1_chunks;
  OperandChunk<G_BM> ra;
  OperandChunk<BN> rb;
  auto load_chunk = [&](int c) {
    const int seg = c < n0_chunks ? 0 : 1;
    const int k0 = seg == 0 ? k_lo + c * G_KC : (c - n0_chunks) * G_KC;
    const int ke = seg == 0 ? k_hi : g.K[1];
    ra.load(g.A[seg], g.a_rs[seg], g.a_ks[seg], m0, g.M, k0, ke);
    rb.load(g.B[seg], g.b_rs[seg], g.b_ks[seg], n0, g.N, k0, ke);
  };
  if (n_chunks > 0) load_chunk(0);
  int chunk = 0;
  for (; chunk < n_chunks; ++chunk) {
    const int s = chunk % NS;
    const int seg = chunk < n0_chunks ? 0 : 1;
    if (chunk >= NS) mbar_wait(bar0 + 8 * s, (uint32_t)((chunk / NS - 1) & 1), 20);   // MMAs of chunk - STAGES have read the stage
    uint8_t *base = smem + s * S::STAGE;
    uint8_t *a_hi = base, *a_lo = base + S::A_STAGE, *b_hi = base + 2 * S::A_STAGE, *b_lo = b_hi + S::B_STAGE;
    ra.store(a_hi, a_lo, g.a_rs[seg], g.a_ks[seg]);
    rb.store(b_hi, b_lo, g.b_rs[seg], g.b_ks[seg]);
    if (chunk + 1 < n_chunks) load_chunk(chunk + 1);          // in flight while this chunk is published and multiplied
    fence_proxy_async();
    __syncthreads();
    // (an ELECTED lane: inside `if (threadIdx.x == 0)` ptxas wraps every MMA in a waterfall loop -- ~105 cycles of issue per MMA,
    // on the critical path of the chunk loop)
    if (threadIdx.x < 32 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int k = 0; k < G_KC / 16; ++k) {
        const uint64_t ah = smem_desc(smem_u32(a_hi) + k * 2 * (G_BM * 16), G_BM * 16, 128);
        const uint64_t al = smem_desc(smem_u32(a_lo) + k * 2 * (G_BM * 16), G_BM * 16, 128);
        const uint64_t bh = smem_desc(smem_u32(b_hi) + k * 2 * (BN * 16), BN * 16, 128);
        const uint64_t bl = smem_desc(smem_u32(b_lo) + k * 2 * (BN * 16), BN * 16, 128);
        umma_bf16<1>(tmem, ah, bh, idesc, (chunk == 0 && k == 0) ? 0u : 1u);
        umma_bf16<1>(tmem, al, bh, idesc, 1u);
        umma_bf16<1>(tmem, ah, bl, idesc, 1u);
      }
      umma_commit<1>(bar0 + 8 * s);
    }
  }
  if (threadIdx.x < 32 && elect_one()) umma_commit<1>(bar0 + 8 * NS);   // arrives once every MMA above has retired (same lane as the issuer)
  const bool any = chunk > 0;
  if (any) mbar_wait(bar0 + 8 * NS, 0, 21);
  tc_fence_after();

  // ---- epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (= rows); warps w and w+4 split the columns --------------
  constexpr int COLS_PER_HALF = BN >= 64 ? BN / 2 : BN;
  const int half = warp >> 2;
  if (BN >= 64 || half == 0) {
    const int m = m0 + (warp & 3) * 32 + lane;
    const bool split = gridDim.z > 1;
    float *out = split ? g.partial + (long long)blockIdx.z * g.M * g.N : g.C;
    const long long ldo = split ? g.N : g.ldc;
    for (int c = half * COLS_PER_HALF; c < (half + 1) * COLS_PER_HALF; c += 32) {
      uint32_t v[32];
      if (any) {
        tmem_ld32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
      }
      if (m < g.M) {
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        float *dst = out + (long long)m * ldo + n0 + c;
        const bool full = n0 + c + 32 <= g.N;
        const bool vec = full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0);
        if (!split) {
          if (g.bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (full || n0 + c + j < g.N) x[j] += __ldg(g.bias + n0 + c + j);
          }
          if (g.accumulate) {
            if (vec) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 o = reinterpret_cast<const float4 *>(dst)[q];
                x[4 * q] += o.x; x[4 * q + 1] += o.y; x[4 * q + 2] += o.z; x[4 * q + 3] += o.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || n0 + c + j < g.N) x[j] += dst[j];
            }
          }
          if (g.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          if (g.mask) {
            const float *mk = g.mask + (long long)m * g.ldm + n0 + c;
            if (full && ((reinterpret_cast<uintptr_t>(mk) & 15) == 0)) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 o = __ldg(reinterpret_cast<const float4 *>(mk) + q);
                x[4 * q] = o.x > 0.f ? x[4 * q] : 0.f;
                x[4 * q + 1] = o.y > 0.f ? x[4 * q + 1] : 0.f;
                x[4 * q + 2] = o.z > 0.f ? x[4 * q + 2] : 0.f;
                x[4 * q + 3] = o.w > 0.f ? x[4 * q + 3] : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (full || n0 + c + j < g.N) x[j] = __ldg(mk + j) > 0.f ? x[j] : 0.f;
            }
          }
        }
        if (vec) {
#pragma unroll
          for (int q = 0; q < 8; ++q) reinterpret_cast<float4 *>(dst)[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (full || n0 + c + j < g.N) dst[j] = x[j];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, BN < 32 ? 32 : BN);
  }
}

// C[m,n] (+)= sum_z partial[z][m][n] in fixed order (deterministic weight gradients)
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float *__restrict__ partial, int splits, long long mn, int N, float *__restrict__ C,
                                                            long long ldc, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += partial[(long long)z * mn + i];
  const long long m = i / N, n = i - m * N;
  float *dst = C + m * ldc + n;
  *dst = accumulate ? *dst + s : s;
}

// out[n] (+)= sum_m X[m*ld + n]: bias gradients.  Two fixed-order phases: per-block partial sums over a row slab, then a
// serial sum over the slabs.
constexpr int CS_ROWS = 256;
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float *__restrict__ X, long long ld, int M, int N, float *__restrict__ part) {
  __shared__ float red[8][33];
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r = threadIdx.x >> 5;
  const int mb = blockIdx.y * CS_ROWS;
  float s = 0.f;
  if (n < N)
    for (int m = mb + r; m < min(M, mb + CS_ROWS); m += 8) s += X[(long long)m * ld + n];
  red[r][threadIdx.x & 31] = s;
  __syncthreads();
  if (r == 0 && n < N) {
    float t = 0.f;
    for (int q = 0; q < 8; ++q) t += red[q][threadIdx.x];
    part[(long long)blockIdx.y * N + n] = t;
  }
}
__global__ void colsum_final_kernel(const float *__restrict__ part, int slabs, int N, float *__restrict__ out, int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int z = 0; z < slabs; ++z) s += part[(long long)z * N + n];
  out[n] = accumulate ? out[n] + s : s;
}

template <int BN>
static int launch_gemm(const GemmDev &g, int splits, cudaStream_t st) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  if (!configured) {
    ANI_CUDA(cudaFuncSetAttribute(gemm_x3_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES));
    configured = true;
  }
  dim3 grid((unsigned)((g.M + G_BM - 1) / G_BM), (unsigned)((g.N + BN - 1) / BN), (unsigned)splits);
  gemm_x3_kernel<BN><<<grid, G_THREADS, S::BYTES, st>>>(g);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // namespace aninerf

using namespace aninerf;

extern "C" {

int64_t aninerf_gemm_workspace_bytes(const aninerf_gemm *p) {
  if (!p || p->split_k <= 1) return 0;
  return (int64_t)p->split_k * p->M * p->N * 4;
}

int aninerf_gemm_x3(const aninerf_gemm *p, void *workspace, int64_t workspace_bytes, void *stream) {
  ANI_CHECK_ARG(p && p->C && p->M >= 0 && p->N > 0 && p->n_seg >= 1 && p->n_seg <= 2);
  if (p->M == 0) return ANINERF_OK;
  GemmDev g;
  memset(&g, 0, sizeof(g));
  for (int s = 0; s < p->n_seg; ++s) {
    ANI_CHECK_ARG(p->seg[s].A && p->seg[s].B && p->seg[s].K >= 0);
    g.A[s] = p->seg[s].A;
    g.B[s] = p->seg[s].B;
    g.a_rs[s] = p->seg[s].a_row_stride;
    g.a_ks[s] = p->seg[s].a_k_stride;
    g.b_rs[s] = p->seg[s].b_row_stride;
    g.b_ks[s] = p->seg[s].b_k_stride;
    g.K[s] = p->seg[s].K;
  }
  g.n_seg = p->n_seg;
  g.M = p->M;
  g.N = p->N;
  g.C = p->C;
  g.ldc = p->ldc;
  g.bias = p->bias;
  g.mask = p->relu_mask;
  g.ldm = p->ld_mask;
  g.relu = p->relu;
  g.accumulate = p->accumulate;
  int splits = p->split_k > 1 ? p->split_k : 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (splits > 1) {
    ANI_CHECK_ARG(p->n_seg == 1 && !p->bias && !p->relu && !p->relu_mask);
    if (!workspace || workspace_bytes < aninerf_gemm_workspace_bytes(p)) return fail(ANINERF_ENOMEM, "%s: split-K workspace too small%s", __func__);
    int per = (p->seg[0].K + splits - 1) / splits;
    per = (per + G_KC - 1) / G_KC * G_KC;
    splits = (p->seg[0].K + per - 1) / per;
    if (splits < 1) splits = 1;
    g.k_per_split = per;
    g.partial = (float *)workspace;
  }
  int rc;
  // output tile width: the widest tile that still yields ~one CTA per SM pair (small training batches have few 128-row tiles)
  int bn = p->N <= 32 ? 32 : p->N <= 64 ? 64 : p->N <= 128 ? 128 : 256;
  {
    const long long m_tiles = (p->M + G_BM - 1) / G_BM;
    while (bn > 64 && m_tiles * ((p->N + bn - 1) / bn) * splits < 96) bn /= 2;
  }
  if (splits > 1) {
    // grid.z > 1 selects the partial-buffer path inside the kernel
    rc = bn == 32 ? launch_gemm<32>(g, splits, st) : bn == 64 ? launch_gemm<64>(g, splits, st) : bn == 128 ? launch_gemm<128>(g, splits, st)
                                                                                                              : launch_gemm<256>(g, splits, st);
    if (rc) return rc;
    const long long mn = (long long)p->M * p->N;
    splitk_reduce_kernel<<<(unsigned)((mn + 255) / 256), 256, 0, st>>>(g.partial, splits, mn, p->N, p->C, p->ldc, p->accumulate);
    ANI_LAUNCHED();
    return ANINERF_OK;
  }
  g.k_per_split = 0;
  return bn == 32 ? launch_gemm<32>(g, 1, st) : bn == 64 ? launch_gemm<64>(g, 1, st) : bn == 128 ? launch_gemm<128>(g, 1, st)
                                                                                                  : launch_gemm<256>(g, 1, st);
}

int aninerf_colsum(const float *X, int64_t ld, int64_t M, int32_t N, float *out, int32_t accumulate, void *workspace, int64_t workspace_bytes,
                   void *stream) {
  ANI_CHECK_ARG(X && out && M >= 0 && N > 0 && M < (int64_t)2147483647);
  cudaStream_t st = (cudaStream_t)stream;
  const int slabs = (int)((M + CS_ROWS - 1) / CS_ROWS);
  if (slabs == 0) {
    if (!accumulate) ANI_CUDA(cudaMemsetAsync(out, 0, (size_t)N * 4, st));
    return ANINERF_OK;
  }
  if (!workspace || workspace_bytes < (int64_t)slabs * N * 4) return fail(ANINERF_ENOMEM, "%s: workspace too small%s", __func__);
  colsum_partial_kernel<<<dim3((unsigned)((N + 31) / 32), (unsigned)slabs), 256, 0, st>>>(X, ld, (int)M, N, (float *)workspace);
  ANI_LAUNCHED();
  colsum_final_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>((const float *)workspace, slabs, N, out, accumulate);
  ANI_LAUNCHED();
  return ANINERF_OK;
}

}  // extern "C"
