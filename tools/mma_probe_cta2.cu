// Micro-probe (bring-up tool, not product; NOT YET RUN ON A GPU -- written at the end of round 1 for the first GPU call
// of round 2): cycles per tcgen05.mma for a CTA PAIR (cta_group::2, M = 256: 128 rows per SM) next to the
// single-CTA form (M = 128), fp16, SS operands, K = 16, on one pair and on 74 pairs (every SM busy).
//
// Question it answers: with cta_group::2 each SM of the pair reads its own 128 A rows
// and only N/2 rows of the weight tile from its own shared memory (the other half comes from the peer), so (a) the
// per-SM weight footprint halves -- weight residency for dec9b / dec8 / dec7 / dec6b -- and (b) the shared-memory
// operand reads that set the measured 46-48 clk floor at N = 64 shrink.  If issue/mma at (M 256, N 64) on 74 pairs
// is close to the N/2 = 32 clk floor, the pair kernel is worth building; if it stays at ~48, it is not.
//
// Correctness of the result is checked too: A = B = 1.0 everywhere, so every accumulator element must equal
// 16 * (number of MMAs issued into it) -- read back from both CTAs' tensor memory.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe_cta2 mma_probe_cta2.cu
// Run under `timeout 60` (a wrong guess about the pair protocol shows up as a hang, not an error).
#include <cstdio>
#include <cuda_runtime.h>
#include "../digipathai_b200/csrc/ptx.cuh"
using namespace dp;

__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// four K steps of one (A, B) tile pair, issued by the leader CTA for both SMs
__device__ __forceinline__ void umma2_f16_ss_k4(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "add.u64 a1, %1, 2;\n\t"
      "add.u64 b1, %2, 2;\n\t"
      "add.u64 a2, %1, 4;\n\t"
      "add.u64 b2, %2, 4;\n\t"
      "add.u64 a3, %1, 6;\n\t"
      "add.u64 b3, %2, 6;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a1, b1, %3, q;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a2, b2, %3, q;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], a3, b3, %3, q;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this offset in both CTAs once every MMA issued so far has completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
               : "memory");
}
__host__ __device__ inline uint32_t idesc_f16(uint32_t m, uint32_t n) {   // fp16 x fp16 -> fp32, K-major A and B
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// out[0] = issue cycles, out[1] = cycles until completion, out[2] = number of wrong accumulator elements (all CTAs)
template <int PAIR>
__global__ void probe(int n, int iters, int b_cycle, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t ra = smem_u32(raw);
  uint8_t* smem = raw + (((ra + 1023u) & ~1023u) - ra);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (PAIR) { tmem_alloc2(&slot, 512); tmem_relinquish2(); } else { tmem_alloc(&slot, 512); tmem_relinquish(); }
  }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  // weight tile: n rows of 128 bytes (single CTA) or n/2 rows per CTA (pair); b_cycle distinct tiles
  const uint32_t b_rows = PAIR ? n / 2 : n;
  const uint32_t b_tile_u = (b_rows * 128) >> 4;
  const int blocks = iters;               // one block = four K steps into the same accumulator
  if (warp == 1 && rank == 0 && elect_one()) {
    const uint32_t idesc = idesc_f16(PAIR ? 256 : 128, n);
    const uint64_t a0 = (static_cast<uint64_t>(sw128_desc_hi(1024)) << 32) | sw128_desc_lo(smem_u32(smem));
    const uint64_t b0 = (static_cast<uint64_t>(sw128_desc_hi(1024)) << 32) | sw128_desc_lo(smem_u32(smem) + 48 * 1024);
    if (PAIR) umma2_f16_ss_k4(tm, a0, b0, idesc, 0); else umma_f16_ss_k4(tm, a0, b0, idesc, 0);
    if (PAIR) umma2_commit(&bar); else umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t0 = clock64();
    int ib = 0;
    for (int i = 0; i < blocks; ++i) {
      if (PAIR) umma2_f16_ss_k4(tm, a0, b0 + ib * b_tile_u, idesc, 1);
      else umma_f16_ss_k4(tm, a0, b0 + ib * b_tile_u, idesc, 1);
      if (++ib == b_cycle) ib = 0;
    }
    long long t1 = clock64();
    if (PAIR) umma2_commit(&bar); else umma_commit(&bar);
    mbar_wait(&bar, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  // everybody (both CTAs) waits for the second completion, then the four warps 4..7 check the accumulator:
  // rows = 32 lanes x 4 warps (warp w reads TMEM lanes 32 (w % 4) ..), columns 0..n-1, expected 16 * 4 * (blocks + 1)
  if (warp >= 4) {
    mbar_wait(&bar, 0);   // in pair mode the peer's barrier receives both multicast arrives too
    mbar_wait(&bar, 1);   // (phase 1 completes ~1e6 clk after phase 0: no waiter can lag a whole phase behind)
    tc_fence_after();
    const float want = 64.0f * (blocks + 1);
    int bad = 0;
    for (int c = 0; c < n; c += 16) {
      uint32_t v[16];
      tmem_ld16(tm + (static_cast<uint32_t>((warp & 3) * 32) << 16) + c, v);
      tmem_ld_wait();
      for (int k = 0; k < 16; ++k) bad += (__uint_as_float(v[k]) != want);
    }
    if (bad) atomicAdd(reinterpret_cast<unsigned long long*>(out + 2), static_cast<unsigned long long>(bad));
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tm, 512); else tmem_dealloc(tm, 512);
  }
}

template <int PAIR>
static int run(int grid, int n, int iters, int b_cycle, long long* d) {
  cudaMemset(d, 0, 32);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = 202 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = PAIR ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, probe<PAIR>, n, iters, b_cycle, d);
  long long h[3] = {0, 0, 0};
  if (e == cudaSuccess) e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  const double m = static_cast<double>(iters) * 4;
  // exactness bound: accumulators hold 64 * (iters + 1) <= 2^24
  printf("%5s %5d %5d %7d | %10.1f %10.1f | %lld wrong\n", PAIR ? "2-CTA" : "1-CTA", grid, n, b_cycle, h[0] / m,
         h[1] / m, h[2]);
  return 0;
}

int main() {
  long long* d;
  cudaMalloc(&d, 32);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024);
  const int iters = 8000;
  printf("%5s %5s %5s %7s | %10s %10s |\n", "mode", "grid", "N", "b_cycle", "issue/mma", "done/mma");
  int ns[] = {64, 128, 256};
  int grids[] = {2, 148};
  for (int n : ns)
    for (int grid : grids)
      for (int bc : {1, 4}) {
        if (run<0>(grid, n, iters, bc, d)) return 1;
        if (run<1>(grid, n, iters, bc, d)) return 1;
      }
  return 0;
}
