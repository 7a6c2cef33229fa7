// Micro-benchmarks behind the design choices of csrc/mlp_tcgen05.cu (sm_100a):
//   1. tcgen05.mma issue / retire rate (cycles per K=16 MMA) for cta_group::2 M=256 (and cta_group::1 M=128), N = 64 / 128 / 256,
//      with the A operand in shared memory in the SWIZZLE_NONE core-matrix layout, in the SWIZZLE_128B layout, or in TENSOR MEMORY
//      (the .ts form: `tcgen05.mma [d], [a_tmem], b_desc, ...`);
//   2. tcgen05.ld / tcgen05.st throughput (32x32b.x32, 4 KB per warp instruction) with 4 / 8 / 16 warps;
//   3. the same MMA streams with epilogue-like traffic (tcgen05.ld + st.shared, tcgen05.ld + tcgen05.st) running next to them;
//   4. the cost of `fence.proxy.async` after 16 st.shared.v4 per thread, and of 256 threads arriving on one mbarrier.
// One thread-block cluster of 2 CTAs by default; `bench_mma <n_clusters>` runs the same on many SM pairs at once (clock / power
// effects).  Prints one line per measurement.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/bench_mma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#define CK(x)                                                                       \
  do {                                                                              \
    cudaError_t e_ = (x);                                                           \
    if (e_ != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                      \
    }                                                                               \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded: returns false on timeout (a protocol bug must not hang the box)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity))
    if (clock64() - t0 > 400000000ll) return false;
  return true;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int PAIR>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  if (PAIR == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
}
template <int PAIR>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if (PAIR == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_ss(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (PAIR == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t bd, uint32_t idesc, uint32_t acc) {
  if (PAIR == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (PAIR == 2)
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr), "r"(v[0]),
      "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]),
      "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_NONE: core matrix = 8 rows x 16 B; lbo = stride between the two K core matrices of a K=16 slice, sbo = 8-row group stride
__device__ __forceinline__ uint64_t desc_none(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// K-major SWIZZLE_128B: rows of 128 B (64 bf16), 8-row groups of 1024 B (sbo), layout type 2 in bits [61,64)
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int n, int m) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

enum { SRC_SS_NONE = 0, SRC_SS_SW128 = 1, SRC_TS = 2 };
__device__ int g_commit_every = 0;      // > 0: tcgen05.commit to a scratch mbarrier after every n MMAs (as a weight ring does per stage)
__device__ int g_wait_every = 0;        // > 0: the issuing thread also try_waits an already-completed mbarrier every n MMAs
enum { SIDE_NONE = 0, SIDE_LD_STS = 1, SIDE_LD_STTM = 2, SIDE_LD_ONLY = 3 };

struct Result {
  unsigned long long cycles;   // of the measured region (thread 0 of CTA 0 of the cluster)
  unsigned long long aux;      // side-traffic iterations completed / per-op cycles
  int ok;
};

constexpr int A_BYTES = 128 * 256 * 2;      // 128 rows x K=256 bf16 = 64 KB
constexpr int B_BYTES = 128 * 256 * 2;      // up to N/2 = 128 rows per CTA x K=256

// mma_test: thread (warp 8, lane 0) of the leader issues `n_mma` MMAs of the given shape back to back (K walks over 16 slices of a
// K=256 operand, repeatedly), commits, waits.  Warps 0..side_warps-1 run the side traffic until the MMAs are done.
// ELECT: the issuing thread is chosen with elect.sync (ptxas then knows the region is executed by ONE thread and issues UTCHMMA
// straight from uniform registers); without it (`lane == 0`) every MMA sits in a compiler-generated ELECT / BRA.U.ANY waterfall
// loop behind R2UR moves, and the ISSUE costs ~105 cycles per MMA -- which is what limited every N < 256 shape in the first
// version of this benchmark, and the single-pass field of the MLP kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}
template <int PAIR, bool ELECT>
__global__ void __launch_bounds__(320, 1) mma_kernel(int src, int n, int n_mma, int side, int side_warps, Result *res) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                          // 64 KB
  uint8_t *sB = smem + A_BYTES;                // 64 KB
  uint8_t *sS = smem + A_BYTES + B_BYTES;      // 64 KB scratch for the side traffic's st.shared
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  __shared__ volatile int done_flag;
  __shared__ unsigned long long side_iters;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR == 2 ? cluster_ctarank() : 0u;
  // operands: a non-trivial bit pattern (bf16 values around 1)
  for (int i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3f803f80u + (i * 2654435761u & 0x007f007fu);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_init(smem_u32(&bars[2]), 1);
    done_flag = 0;
    side_iters = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc<PAIR>(smem_u32(&tmem_slot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TS: fill the A columns (256..383) with something
  if (warp < 4) {
    uint32_t v[32];
    for (int j = 0; j < 32; ++j) v[j] = 0x3f803f80u + j;
    for (int c = 0; c < 4; ++c) tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c * 32, v);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();
  tc_fence_after();
  long long t0 = 0, t1 = 0;
  bool ok = true;
  if (warp == 8) {
    if (rank == 0 && (ELECT ? elect_one() : lane == 0)) {
      const uint32_t idesc = instr_desc(n, 128 * PAIR);
      const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
      const int nb = n / PAIR;                                  // B rows held by each CTA
      const int commit_every = g_commit_every, wait_every = g_wait_every;      // (in MMAs; multiples of 16)
      const uint64_t ad_sw = desc_sw128(a0), bd_sw = desc_sw128(b0);
      const uint64_t ad_n = desc_none(a0, 2048, 128), bd_n = desc_none(b0, nb * 16, 128);
      const uint32_t b_kb = (uint32_t)(nb * 128) >> 4, b_k16 = (uint32_t)(2 * nb * 16) >> 4;      // descriptor address units (16 B)
      t0 = clock64();
      // sixteen K=16 slices per pass, every descriptor a constant offset from a loop-invariant base: the issue loop of the MLP kernel
      for (int i = 0; i < n_mma; i += 16) {
        if (wait_every > 0 && (i & (wait_every - 1)) == 0) (void)mbar_try_wait(smem_u32(&bars[2]), 1);
        if (src == SRC_TS) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_ts<PAIR>(tmem, tmem + 256 + k * 8, bd_sw + (uint64_t)((k >> 2) * b_kb + (k & 3) * 2), idesc, (i | k) ? 1u : 0u);
        } else if (src == SRC_SS_SW128) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_ss<PAIR>(tmem, ad_sw + (uint64_t)((k >> 2) * 1024 + (k & 3) * 2), bd_sw + (uint64_t)((k >> 2) * b_kb + (k & 3) * 2), idesc,
                          (i | k) ? 1u : 0u);
        } else {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_ss<PAIR>(tmem, ad_n + (uint64_t)(k * 256), bd_n + (uint64_t)(k * b_k16), idesc, (i | k) ? 1u : 0u);
        }
        if (commit_every > 0 && ((i + 16) & (commit_every - 1)) == 0) umma_commit<PAIR>(smem_u32(&bars[1]));
      }
      umma_commit<PAIR>(smem_u32(&bars[0]));
      const long long t_issued = clock64();
      ok = mbar_wait(smem_u32(&bars[0]), 0);
      t1 = clock64();
      done_flag = 1;
      res[blockIdx.x / PAIR].cycles = (unsigned long long)(t1 - t0);
      res[blockIdx.x / PAIR].aux = (unsigned long long)(t_issued - t0);
      res[blockIdx.x / PAIR].ok = ok ? 1 : 0;
    } else if (rank != 0 && lane == 0 && PAIR == 2) {
      ok = mbar_wait(smem_u32(&bars[0]), 0);                    // the multicast commit arrives here too
      done_flag = 1;
    }
  } else if (warp < side_warps && side != SIDE_NONE) {
    // epilogue-like traffic on the accumulator columns 0..255 of this CTA (reads race with the MMAs' writes: timing only)
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const int row = (warp & 3) * 32 + lane;
    unsigned long long it = 0;
    uint32_t v[32];
    uint32_t sink = 0;
    while (!done_flag) {
      const int c = (int)(it & 7) * 32;
      tmem_ld32(t_lane + c, v);
      tmem_ld_wait();
      if (side == SIDE_LD_STS) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 h = make_uint4(v[q * 8] + v[q * 8 + 1], v[q * 8 + 2] + v[q * 8 + 3], v[q * 8 + 4] + v[q * 8 + 5], v[q * 8 + 6] + v[q * 8 + 7]);
          *reinterpret_cast<uint4 *>(sS + (((it & 7) * 4 + q) * 2048 + row * 16)) = h;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      } else if (side == SIDE_LD_STTM) {
        uint32_t h[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) h[q] = v[2 * q] + v[2 * q + 1];
        tmem_st16(t_lane + 384 + (int)(it & 7) * 16, h);          // columns 384..511: not read by the TS MMAs
        tmem_st_wait();
      } else {
        sink += v[0] + v[31];
      }
      ++it;
    }
    if (sink == 0x12345678u) sS[0] = 1;
    if (lane == 0) atomicAdd(&side_iters, it);
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR == 2) cluster_sync_all();
  if (threadIdx.x == 0 && rank == 0 && side != SIDE_NONE) res[blockIdx.x / PAIR].aux = side_iters;
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<PAIR>(tmem, 512);
  }
}

// tcgen05.ld / tcgen05.st throughput: `warps` warps each move `iters` x 4 KB
__global__ void __launch_bounds__(512, 1) tmem_rw_kernel(int warps, int iters, int store, Result *res) {
  __shared__ uint32_t tmem_slot;
  __shared__ long long t_begin[16], t_end[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<1>(smem_u32(&tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t v[32];
  for (int j = 0; j < 32; ++j) v[j] = j + lane;
  uint32_t sink = 0;
  __syncthreads();
  if (warp < warps) {
    const uint32_t t_lane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const long long a = clock64();
    for (int i = 0; i < iters; ++i) {
      const int c = ((i + (warp >> 2)) & 15) * 32;
      if (store) {
        tmem_st32(t_lane + c, v);
      } else {
        tmem_ld32(t_lane + c, v);
      }
      if ((i & 3) == 3) {                         // a few instructions in flight, as the pipelined epilogue does
        if (store) tmem_st_wait();
        else tmem_ld_wait();
        sink += v[5];
      }
    }
    if (store) tmem_st_wait();
    else tmem_ld_wait();
    const long long b = clock64();
    if (lane == 0) {
      t_begin[warp] = a;
      t_end[warp] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long a = t_begin[0], b = t_end[0];
    for (int w = 1; w < warps; ++w) {
      a = t_begin[w] < a ? t_begin[w] : a;
      b = t_end[w] > b ? t_end[w] : b;
    }
    res[blockIdx.x].cycles = (unsigned long long)(b - a);
    res[blockIdx.x].aux = sink;
    res[blockIdx.x].ok = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<1>(tmem, 512);
  }
}

// hand-off costs: (a) fence.proxy.async after `n_sts` st.shared.v4 per thread, 256 threads; (b) 256 threads arriving on one mbarrier
__global__ void __launch_bounds__(320, 1) handoff_kernel(int n_sts, Result *res) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ long long t_first, t_last_fence, t_seen;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    t_first = 0x7fffffffffffffffll;
    t_last_fence = 0;
  }
  __syncthreads();
  if (warp < 8) {
    const int row = threadIdx.x & 127, half = threadIdx.x >> 7;
    const long long a = clock64();
    for (int q = 0; q < n_sts; ++q)
      *reinterpret_cast<uint4 *>(smem + ((half * n_sts + q) * 2048 + row * 16)) = make_uint4(q, row, half, 7);
    const long long b = clock64();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const long long c = clock64();
    mbar_arrive(smem_u32(&bar));
    atomicMin((unsigned long long *)&t_first, (unsigned long long)a);
    atomicMax((unsigned long long *)&t_last_fence, (unsigned long long)c);
    if (threadIdx.x == 0) {
      res[0].cycles = (unsigned long long)(b - a);      // issue time of the stores (one thread)
      res[0].aux = (unsigned long long)(c - b);         // the fence (one thread)
    }
  } else if (threadIdx.x == 256) {
    const bool ok = mbar_wait(smem_u32(&bar), 0);
    t_seen = clock64();
    res[1].ok = ok;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    res[0].ok = 1;
    res[1].cycles = (unsigned long long)(t_seen - t_first);        // first store issued -> waiter sees the phase complete
    res[1].aux = (unsigned long long)(t_seen - t_last_fence);      // last fence retired -> waiter released (arrive + wake-up)
  }
}

// ---- correctness of the SWIZZLE_128B K-major layout used by csrc/mlp_tcgen05.cu: D[256 x N] = A[256 x K] * B[N x K]^T with
// cta_group::2 (CTA r holds A rows [128r, 128r+128) and B rows [N/2 r, N/2 (r+1))), K = 64 * kblocks, operands written in the
// layout the kernel's epilogue / weight packer use:
//   element (row, k) of a K-block (64 wide) at  (row/8)*1024 + (row%8)*128 + (((k/8) ^ (row%8)) * 16) + (k%8)*2
__device__ __forceinline__ uint32_t sw128_off(int row, int k) {   // byte offset inside one K-block of `rows` rows
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 3) ^ (row & 7)) & 7) << 4) + (k & 7) * 2);
}
__global__ void __launch_bounds__(320, 1) verify_kernel(const uint16_t *A, const uint16_t *B, int n, int kblocks, float *D, int *status, int ts) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sA = smem;                          // kblocks x 16 KB
  uint8_t *sB = smem + 4 * 16384;              // kblocks x (n/2 rows x 128 B)
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int K = 64 * kblocks, nb = n / 2;
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int row = i / K, k = i % K;
    *reinterpret_cast<uint16_t *>(sA + (k >> 6) * 16384 + sw128_off(row, k & 63)) = A[(size_t)(rank * 128 + row) * K + k];
  }
  for (int i = threadIdx.x; i < nb * K; i += blockDim.x) {
    const int row = i / K, k = i % K;
    *reinterpret_cast<uint16_t *>(sB + (k >> 6) * (nb * 128) + sw128_off(row, k & 63)) = B[(size_t)(rank * nb + row) * K + k];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9) tmem_alloc<2>(smem_u32(&tmem_slot), 512);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (ts && warp < 4) {
    // A operand in tensor memory: lane = row, 32-bit column c of the operand = K elements (2c, 2c+1), element 2c in the low half
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < K / 2; c0 += 16) {
      uint32_t v[16];
      for (int j = 0; j < 16; ++j) {
        const int k = 2 * (c0 + j);
        v[j] = (uint32_t)A[(size_t)(rank * 128 + row) * K + k] | ((uint32_t)A[(size_t)(rank * 128 + row) * K + k + 1] << 16);
      }
      tmem_st16(tmem + ((uint32_t)(warp * 32) << 16) + 256 + c0, v);
    }
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 8 && lane == 0) {
    bool ok = true;
    if (rank == 0) {
      const uint32_t idesc = instr_desc(n, 256);
      for (int i = 0; i < 4 * kblocks; ++i) {
        const uint64_t ad = desc_sw128(smem_u32(sA) + (i >> 2) * 16384 + (i & 3) * 32);
        const uint64_t bd = desc_sw128(smem_u32(sB) + (i >> 2) * (nb * 128) + (i & 3) * 32);
        if (ts) umma_ts<2>(tmem, tmem + 256 + i * 8, bd, idesc, i ? 1u : 0u);
        else umma_ss<2>(tmem, ad, bd, idesc, i ? 1u : 0u);
      }
      umma_commit<2>(smem_u32(&bars[0]));
    }
    ok = mbar_wait(smem_u32(&bars[0]), 0);
    if (!ok) atomicExch(status, 1);
  }
  __syncthreads();
  tc_fence_after();
  if (warp < 4) {
    for (int c = 0; c < n; c += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) D[(size_t)(rank * 128 + warp * 32 + lane) * n + c + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<2>(tmem, 512);
  }
}

static uint16_t f2bf_host(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static float bf2f_host(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static void run_verify() {
  printf("## 0. SWIZZLE_128B K-major operand layout: numerical check of D = A B^T (cta_group::2, M = 256)\n");
  for (int ts = 0; ts < 2; ++ts)
  for (int n : {256, 128, 32})
    for (int kblocks : {1, 4}) {
      const int K = 64 * kblocks;
      uint16_t *hA = (uint16_t *)malloc(256 * K * 2), *hB = (uint16_t *)malloc(n * K * 2);
      uint32_t seed = 12345u + n + kblocks;
      auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 65536.0f - 0.5f; };
      for (int i = 0; i < 256 * K; ++i) hA[i] = f2bf_host(rnd());
      for (int i = 0; i < n * K; ++i) hB[i] = f2bf_host(rnd());
      uint16_t *dA, *dB;
      float *dD;
      int *dS;
      CK(cudaMalloc(&dA, 256 * K * 2));
      CK(cudaMalloc(&dB, n * K * 2));
      CK(cudaMalloc(&dD, 256 * n * 4));
      CK(cudaMalloc(&dS, 4));
      CK(cudaMemcpy(dA, hA, 256 * K * 2, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(dB, hB, n * K * 2, cudaMemcpyHostToDevice));
      CK(cudaMemset(dD, 0xff, 256 * n * 4));
      CK(cudaMemset(dS, 0, 4));
      const int smem = 8 * 16384;
      CK(cudaFuncSetAttribute(verify_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2);
      cfg.blockDim = dim3(320);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, verify_kernel, (const uint16_t *)dA, (const uint16_t *)dB, n, kblocks, dD, dS, ts));
      CK(cudaDeviceSynchronize());
      float *hD = (float *)malloc(256 * n * 4);
      int st = 0;
      CK(cudaMemcpy(hD, dD, 256 * n * 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
      double worst = 0;
      for (int m = 0; m < 256; ++m)
        for (int j = 0; j < n; ++j) {
          double acc = 0;
          for (int k = 0; k < K; ++k) acc += (double)bf2f_host(hA[m * K + k]) * (double)bf2f_host(hB[j * K + k]);
          const double e = fabs(acc - (double)hD[m * n + j]);
          worst = e > worst ? e : worst;
        }
      printf("%s N=%3d K=%3d: max |D - A B^T| = %.3e %s%s\n", ts ? "A in TMEM (.ts, packed pairs, K even in the low half)" : "A in smem  (SWIZZLE_128B)", n, K, worst, worst < 1e-4 ? "OK" : "MISMATCH", st ? " (TIMEOUT)" : "");
      free(hA); free(hB); free(hD);
      cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
    }
}

template <int PAIR, bool ELECT = true>
static void run_mma(const char *name, int src, int n, int n_mma, int side, int side_warps, int clusters, Result *d_res) {
  const int smem = A_BYTES + B_BYTES + 65536;
  CK(cudaFuncSetAttribute(mma_kernel<PAIR, ELECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * PAIR);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CK(cudaMemset(d_res, 0, sizeof(Result) * 256));
  for (int rep = 0; rep < 2; ++rep) {          // second run = warm
    CK(cudaLaunchKernelEx(&cfg, mma_kernel<PAIR, ELECT>, src, n, n_mma, side, side_warps, d_res));
    CK(cudaDeviceSynchronize());
  }
  Result h[256];
  CK(cudaMemcpy(h, d_res, sizeof(Result) * 256, cudaMemcpyDeviceToHost));
  double mean = 0, mx = 0;
  int ok = 1;
  for (int c = 0; c < clusters; ++c) {
    mean += (double)h[c].cycles / clusters;
    mx = h[c].cycles > mx ? (double)h[c].cycles : mx;
    ok &= h[c].ok;
  }
  const double floor_cyc = 128.0 * n / 256.0;    // max(M,128) * N / (256 * cta_group) per CTA pair member
  printf("%-44s N=%3d  %7.1f cycles/MMA (max %7.1f; floor %5.1f -> %5.1f%% of peak)  aux %llu  %s\n", name, n, mean / n_mma, mx / n_mma, floor_cyc,
         100.0 * floor_cyc / (mean / n_mma), (unsigned long long)h[0].aux, ok ? "" : "TIMEOUT");
}

int main(int argc, char **argv) {
  const int clusters = argc > 1 ? atoi(argv[1]) : 1;
  Result *d_res;
  CK(cudaMalloc(&d_res, sizeof(Result) * 256));
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  printf("# %s, %d SMs, %d cluster(s) of 2 CTAs\n", p.name, p.multiProcessorCount, clusters);
  if (argc > 2 && !strcmp(argv[2], "verify")) {
    run_verify();
    return 0;
  }
  run_verify();
  const int NM = 256;
  const char *srcs[3] = {"A smem SWIZZLE_NONE", "A smem SWIZZLE_128B", "A in TMEM (.ts)"};
  printf("## 1. MMA rate, nothing else running, issued by an ELECTED lane (aux = cycles to ISSUE the %d MMAs)\n", NM);
  for (int src = 0; src < 3; ++src)
    for (int n : {256, 128, 64, 32}) {
      char name[96];
      snprintf(name, sizeof(name), "cta_group::2 M=256, %s", srcs[src]);
      run_mma<2>(name, src, n, NM, SIDE_NONE, 0, clusters, d_res);
    }
  for (int src = 0; src < 3; ++src)
    for (int n : {256, 128}) {
      char name[96];
      snprintf(name, sizeof(name), "cta_group::1 M=128, %s", srcs[src]);
      run_mma<1>(name, src, n, NM, SIDE_NONE, 0, clusters, d_res);
    }
  printf("## 1a. the same issued by `lane == 0` of a divergent warp (compiler-generated waterfall loop around every MMA)\n");
  for (int src : {1, 2})
    for (int n : {256, 128, 32}) {
      char name[96];
      snprintf(name, sizeof(name), "lane==0: cta_group::2 M=256, %s", srcs[src]);
      run_mma<2, false>(name, src, n, NM, SIDE_NONE, 0, clusters, d_res);
    }
  printf("## 1b. the same with a tcgen05.commit (and an mbarrier try_wait by the issuing thread) every n MMAs, as a weight ring does per stage\n");
  for (int every : {0, 16, 64}) {
    CK(cudaMemcpyToSymbol(g_commit_every, &every, sizeof(int)));
    for (int w : {0, 1}) {
      const int we = w ? every : 0;
      CK(cudaMemcpyToSymbol(g_wait_every, &we, sizeof(int)));
      for (int src : {1, 2}) {
        char name[96];
        snprintf(name, sizeof(name), "commit every %d%s, %s", every, w ? " + try_wait" : "", srcs[src]);
        run_mma<2>(name, src, 256, NM, SIDE_NONE, 0, clusters, d_res);
      }
    }
  }
  {
    const int zero = 0;
    CK(cudaMemcpyToSymbol(g_commit_every, &zero, sizeof(int)));
    CK(cudaMemcpyToSymbol(g_wait_every, &zero, sizeof(int)));
  }
  printf("## 2. MMA rate with epilogue-like traffic next to it (aux = side iterations of 4 KB per warp while %d MMAs ran)\n", 1024);
  const char *sides[4] = {"", "8 warps tcgen05.ld + st.shared + fence", "8 warps tcgen05.ld + tcgen05.st", "8 warps tcgen05.ld only"};
  for (int src : {0, 2})
    for (int n : {256, 128})
      for (int side : {1, 2, 3}) {
        char name[96];
        snprintf(name, sizeof(name), "%s | %s", src == 0 ? "SS none" : "TS", sides[side]);
        run_mma<2>(name, src, n, 1024, side, 8, clusters, d_res);
      }
  printf("## 3. tcgen05.ld / tcgen05.st throughput (32x32b.x32 = 4 KB per warp instruction), one CTA\n");
  for (int store = 0; store < 2; ++store)
    for (int warps : {1, 4, 8, 16}) {
      const int iters = 512;
      CK(cudaMemset(d_res, 0, sizeof(Result) * 256));
      for (int rep = 0; rep < 2; ++rep) {
        tmem_rw_kernel<<<1, 512>>>(warps, iters, store, d_res);
        CK(cudaDeviceSynchronize());
      }
      Result h;
      CK(cudaMemcpy(&h, d_res, sizeof(Result), cudaMemcpyDeviceToHost));
      printf("%s %2d warps: %8llu cycles for %d x 4 KB per warp -> %6.1f B/cycle/SM, %6.1f cycles per instruction per warp\n", store ? "tcgen05.st" : "tcgen05.ld",
             warps, h.cycles, iters, (double)warps * iters * 4096 / h.cycles, (double)h.cycles / iters);
    }
  printf("## 4. hand-off costs (256 threads)\n");
  CK(cudaFuncSetAttribute(handoff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 16 * 2048));
  for (int n_sts : {0, 4, 16}) {
    CK(cudaMemset(d_res, 0, sizeof(Result) * 256));
    for (int rep = 0; rep < 2; ++rep) {
      handoff_kernel<<<1, 320, 2 * 16 * 2048>>>(n_sts, d_res);
      CK(cudaDeviceSynchronize());
    }
    Result h[2];
    CK(cudaMemcpy(h, d_res, sizeof(Result) * 2, cudaMemcpyDeviceToHost));
    printf("%2d st.shared.v4 per thread: stores %llu cycles, fence.proxy.async %llu cycles (thread 0); first store -> waiter released %llu cycles; last fence -> released %llu cycles %s\n",
           n_sts, h[0].cycles, h[0].aux, h[1].cycles, h[1].aux, h[1].ok ? "" : "TIMEOUT");
  }
  CK(cudaFree(d_res));
  return 0;
}
