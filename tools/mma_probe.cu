// Micro-probe (bring-up tool, not product): cycles per tcgen05.mma (M=128, K=16, fp16, SS mode) as a function of
// N, of the A-descriptor start alignment (halo mode uses starts that are not 1024-byte aligned) and of the
// 8-row-group stride (SBO).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../digipathai_b200/csrc/ptx.cuh"
using namespace dp;

// Pattern probe: like the real issue loop, every k4 block may use a different weight tile (b_cycle distinct tiles of
// n rows), a different A start (a_cycle distinct row offsets) and a different accumulator (d_cycle column groups).
__global__ void probe_pattern(int n, int sbo, int iters, int a_cycle, int b_cycle, int d_cycle, long long* out,
                              int commit_every = 0) {
  extern __shared__ uint8_t raw[];
  const uint32_t ra = smem_u32(raw);
  uint8_t* smem = raw + (((ra + 1023u) & ~1023u) - ra);
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy_bar;   // target of the in-loop commits: nobody waits on it
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); mbar_init(&dummy_bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = make_idesc_f16(n);
    const uint64_t a0 = (static_cast<uint64_t>(sw128_desc_hi(sbo)) << 32) | sw128_desc_lo(smem_u32(smem));
    const uint64_t b0 = (static_cast<uint64_t>(sw128_desc_hi(1024)) << 32) | sw128_desc_lo(smem_u32(smem) + 48 * 1024);
    const uint32_t b_tile_u = (n * 128) >> 4;
    umma_f16_ss_k4(tm, a0, b0, idesc, 0);
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    int ia = 0, ib = 0, id = 0;
    for (int i = 0; i < iters; ++i) {
      umma_f16_ss_k4(tm + id * n, a0 + ia * 8 * 11, b0 + ib * b_tile_u, idesc, 1);
      if (++ia == a_cycle) ia = 0;
      if (++ib == b_cycle) ib = 0;
      if (++id == d_cycle) id = 0;
      if (commit_every && (i + 1) % commit_every == 0) umma_commit(&dummy_bar);
    }
    long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

__global__ void probe(int n, int a_row_off, int sbo, int iters, int unroll_k, int spin_mode, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t ra = smem_u32(raw);
  uint8_t* smem = raw + (((ra + 1023u) & ~1023u) - ra);
  __shared__ uint64_t bar;
  __shared__ uint64_t done_bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 180 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); mbar_init(&done_bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_f16(n);
    const uint32_t a_hi = sw128_desc_hi(sbo), b_hi = sw128_desc_hi(1024);
    const uint32_t a_lo = sw128_desc_lo(smem_u32(smem) + a_row_off * 128);
    const uint32_t b_lo = sw128_desc_lo(smem_u32(smem) + 96 * 1024);
    // warm
    umma_f16_ss_parts(tm, a_lo, a_hi, b_lo, b_hi, idesc, 0);
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      umma_f16_ss_parts(tm, a_lo, a_hi, b_lo, b_hi, idesc, 1);
      if (unroll_k) {
        umma_f16_ss_parts(tm, a_lo + 2, a_hi, b_lo + 2, b_hi, idesc, 1);
        umma_f16_ss_parts(tm, a_lo + 4, a_hi, b_lo + 4, b_hi, idesc, 1);
        umma_f16_ss_parts(tm, a_lo + 6, a_hi, b_lo + 6, b_hi, idesc, 1);
      }
    }
    long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    mbar_arrive(&done_bar);
  } else if (warp >= 4 && spin_mode) {
    // waiting roles as in the conv kernel: spin_mode 1 = all 32 lanes poll, 2 = lane 0 polls then __syncwarp
    if (spin_mode == 1) mbar_wait(&done_bar, 0);
    else { if (lane == 0) mbar_wait(&done_bar, 0); __syncwarp(); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 2000;
  // Second axis (round 1, session 2): the same loop on 1 CTA vs one CTA on every SM.  The per-MMA time of a whole-
  // chip run is what a persistent conv kernel can actually get (power management), not the single-SM figure.
  printf("%5s %8s %6s %4s %5s | %10s %10s\n", "N", "a_rowoff", "sbo", "k4", "grid", "issue/mma", "done/mma");
  cudaFuncSetAttribute(probe_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("pattern probe (k4 blocks, 148 CTAs): distinct A starts / weight tiles / accumulators cycled per block\n");
  printf("%5s %7s %7s %7s | %10s %10s\n", "N", "a_cycle", "b_cycle", "d_cycle", "issue/mma", "done/mma");
  {
    int pn[] = {64, 96, 128, 256};
    int cyc[][3] = {{1, 1, 1}, {9, 1, 1}, {1, 4, 1}, {1, 1, 2}, {1, 1, 4}, {9, 4, 1}, {9, 4, 2}, {4, 4, 4}};
    for (int n : pn)
      for (auto& c : cyc) {
        if (c[2] * n > 512 || c[1] * n * 128 > 150 * 1024) continue;
        probe_pattern<<<148, 128, 220 * 1024>>>(n, 1280, iters * 4, c[0], c[1], c[2], d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double m = (double)iters * 4 * 4;
        printf("%5d %7d %7d %7d | %10.1f %10.1f\n", n, c[0], c[1], c[2], h[0] / m, h[1] / m);
      }
  }
  printf("commit probe (148 CTAs, k4 blocks): tcgen05.commit to an unwaited mbarrier every C blocks\n");
  printf("%5s %8s | %10s %10s\n", "N", "every", "issue/mma", "done/mma");
  {
    int pn[] = {64, 128, 256};
    int ev[] = {0, 16, 4, 2, 1};
    for (int n : pn)
      for (int c : ev) {
        probe_pattern<<<148, 128, 220 * 1024>>>(n, 1280, iters * 4, 9, 4 * 128 * n <= 150 * 1024 ? 4 : 2, 1, d, c);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double m = (double)iters * 4 * 4;
        printf("%5d %8d | %10.1f %10.1f\n", n, c, h[0] / m, h[1] / m);
      }
  }
  int ns[] = {64};
  int grids[] = {1, 148};
  for (int n : ns)
    for (int grid : grids)
      for (int rep = 0; rep < 2; ++rep) {
        const int k4 = 1, off = 11, sbo = 1280;
        probe<<<grid, 128, 200 * 1024>>>(n, off, sbo, iters * (rep ? 8 : 1), k4, 0, d);
        long long h[2];
        cudaError_t e = cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        const double m = (double)iters * (rep ? 8 : 1) * 4;
        printf("%5d %8d %6d %4d %5d | %10.1f %10.1f   (%d MMAs per CTA)\n", n, off, sbo, k4, grid, h[0] / m, h[1] / m, (int)m);
      }
  return 0;
}
