// Micro-benchmark: how fast can one SM pull an L2-resident operand image into shared memory?
// (bulk TMA copies of various sizes / ring depths; all SMs reading the same image vs distinct ones)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// one thread per CTA drives a ring of `stages` x `chunk` bytes, `split` bulk copies per stage
__global__ void stream_kernel(const uint8_t *img, size_t img_bytes, size_t per_cta_stride, int stages, int chunk, int split, int iters,
                              unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[16];
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint8_t *base = img + (size_t)blockIdx.x * per_cta_stride;
    size_t off = 0;
    long long t0 = clock64();
    int issued = 0, done = 0;
    uint32_t phase_bits = 0;
    while (done < iters) {
      while (issued < iters && issued - done < stages) {
        int s = issued % stages;
        uint32_t bar = smem_u32(&bars[s]);
        mbar_expect_tx(bar, chunk);
        int piece = chunk / split;
        for (int q = 0; q < split; ++q) bulk_g2s(smem_u32(smem + (size_t)s * chunk + q * piece), base + off + q * piece, piece, bar);
        off += chunk;
        if (off + chunk > img_bytes) off = 0;
        ++issued;
      }
      int s = done % stages;
      while (!mbar_try_wait(smem_u32(&bars[s]), (phase_bits >> s) & 1)) {}
      phase_bits ^= 1u << s;
      ++done;
    }
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

// P producer warps per CTA, each with its own ring of `stages` x `chunk`
__global__ void stream_multi_kernel(const uint8_t *img, size_t img_bytes, int stages, int chunk, int iters, int lanes,
                                    unsigned long long *cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bars[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nprod = (blockDim.x >> 5) * lanes;
  const int pid = warp * lanes + lane;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages * nprod; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (lane < lanes) {
    uint8_t *ring = smem + (size_t)pid * stages * chunk;
    uint64_t *mybars = bars + pid * stages;
    size_t off = (size_t)pid * 65536;
    int issued = 0, done = 0;
    uint32_t phase_bits = 0;
    while (done < iters) {
      while (issued < iters && issued - done < stages) {
        int s = issued % stages;
        uint32_t bar = smem_u32(&mybars[s]);
        mbar_expect_tx(bar, chunk);
        bulk_g2s(smem_u32(ring + (size_t)s * chunk), img + off, chunk, bar);
        off += chunk;
        if (off + chunk > img_bytes) off = 0;
        ++issued;
      }
      int s = done % stages;
      while (!mbar_try_wait(smem_u32(&mybars[s]), (phase_bits >> s) & 1)) {}
      phase_bits ^= 1u << s;
      ++done;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
}

int main() {
  const size_t img_bytes = 2 << 20;
  uint8_t *img;
  cudaMalloc(&img, img_bytes * 40);
  cudaMemset(img, 1, img_bytes * 40);
  unsigned long long *cyc;
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct Cfg { int stages, chunk, split; } cfgs[] = {{3, 16384, 1}, {7, 16384, 1}, {12, 16384, 1}, {3, 16384, 4}, {3, 16384, 16}, {6, 32768, 1},
                                                      {6, 32768, 8}, {12, 8192, 1}, {24, 4096, 1}, {8, 4096, 1}, {3, 65536, 1}, {3, 65536, 16}};
  for (int distinct = 0; distinct < 1; ++distinct)
    for (int grid : {148})
      for (auto c : cfgs) {
        int iters = (int)((64u << 20) / c.chunk / (grid == 1 ? 1 : 4));
        size_t stride = distinct ? (img_bytes / 8) : 0;     // distinct: 256 KB apart (still L2 resident: 37 MB total window)
        size_t window = distinct ? img_bytes / 8 : img_bytes;
        stream_kernel<<<grid, 32, c.stages * c.chunk>>>(img, window, stride, c.stages, c.chunk, c.split, iters, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        unsigned long long h[148];
        cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double mx = 0;
        for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("distinct=%d grid=%3d stages=%2d chunk=%5d split=%2d : %.1f B/cycle/SM\n", distinct, grid, c.stages, c.chunk, c.split,
               (double)iters * c.chunk / mx);
      }
  cudaFuncSetAttribute(stream_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  struct M { int warps, lanes, stages, chunk; } ms[] = {{1, 1, 3, 16384}, {2, 1, 3, 16384}, {4, 1, 3, 16384}, {4, 1, 2, 16384}, {8, 1, 2, 8192},
                                                        {1, 4, 3, 16384}, {1, 8, 3, 8192}, {1, 32, 2, 2048}, {4, 4, 2, 4096}, {4, 1, 3, 8192}};
  for (int grid : {1, 148})
    for (auto m : ms) {
      int iters = 2000;
      stream_multi_kernel<<<grid, m.warps * 32, m.warps * m.lanes * m.stages * m.chunk>>>(img, img_bytes, m.stages, m.chunk, iters, m.lanes, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      unsigned long long h[148];
      cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
      double mx = 0;
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("multi grid=%3d warps=%d lanes=%2d stages=%d chunk=%5d : %.1f B/cycle/SM\n", grid, m.warps, m.lanes, m.stages, m.chunk,
             (double)iters * m.chunk * m.warps * m.lanes / mx);
    }
  return 0;
}
