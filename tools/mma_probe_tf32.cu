// Micro-probe (bring-up tool, not product): cycles per tcgen05.mma.kind::tf32 (M = 128, K = 8, SS operands) as a function
// of N, of the operand pattern of the 3xTF32 kernels (hi*hi, lo*hi, hi*lo on the same accumulator) and of the 8-row-group
// stride of the A descriptor (1024 = plain tile, 1280 = halo tile), next to kind::f16 (K = 16) on the same tiles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe_tf32 mma_probe_tf32.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../digipathai_b200/csrc/ptx.cuh"
#include "../digipathai_b200/csrc/precise_tc.cuh"
using namespace dp;

// mode 0: kind::f16, one (A, B) pair repeated;  1: kind::tf32, one pair repeated;  2: kind::tf32, the 3-MMA pattern over
// 4 K-steps (12 MMAs per block) with distinct hi / lo tiles;  3: like 2 but the three products go to three accumulators
__global__ void probe(int mode, int n, int sbo, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t ra = smem_u32(raw);
  uint8_t* smem = raw + (((ra + 1023u) & ~1023u) - ra);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem)[i] = mode == 0 ? 0x3c003c00u : 0x3f800000u;   // 1.0 in fp16 pairs / fp32
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  if (threadIdx.x == 32) { mbar_init(&bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1 && elect_one()) {
    const uint32_t idesc = mode == 0 ? make_idesc_f16(n) : make_idesc_tf32(n);
    const uint32_t base = smem_u32(smem);
    const uint64_t ah = make_sw128_desc(base + 11 * 128, sbo, 0), al = make_sw128_desc(base + 32 * 1024 + 11 * 128, sbo, 0);
    const uint64_t bh = make_sw128_desc(base + 64 * 1024, 1024, 0), bl = make_sw128_desc(base + 96 * 1024, 1024, 0);
    if (mode == 0) umma_f16_ss(tm, ah, bh, idesc, 0); else umma_tf32_ss(tm, ah, bh, idesc, 0);
    umma_commit(&bar); mbar_wait(&bar, 0);
    long long t0 = clock64();
    long long n_mma = 0;
    for (int i = 0; i < iters; ++i) {
      if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tm, ah + 2 * k, bh + 2 * k, idesc, 1);
        n_mma += 4;
      } else if (mode == 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_tf32_ss(tm, ah + 2 * k, bh + 2 * k, idesc, 1);
        n_mma += 4;
      } else {
        const uint32_t d1 = mode == 3 ? tm + n : tm, d2 = mode == 3 ? tm + 2 * n : tm;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_tf32_ss(tm, ah + 2 * k, bh + 2 * k, idesc, 1);
          umma_tf32_ss(d1, al + 2 * k, bh + 2 * k, idesc, 1);
          umma_tf32_ss(d2, ah + 2 * k, bl + 2 * k, idesc, 1);
        }
        n_mma += 12;
      }
    }
    long long t1 = clock64();
    umma_commit(&bar); mbar_wait(&bar, 1);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; out[2] = n_mma; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 32);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%6s %5s %6s %5s | %10s %10s\n", "mode", "N", "sbo", "grid", "issue/mma", "done/mma");
  const char* names[] = {"f16", "tf32", "tf32x3", "x3/3acc"};
  for (int grid : {1, 148})
    for (int mode = 0; mode < 4; ++mode)
      for (int n : {32, 64, 96, 128})
        for (int sbo : {1024, 1280}) {
          if (mode == 3 && 3 * n > 512) continue;
          probe<<<grid, 128, 200 * 1024>>>(mode, n, sbo, 2000, d);
          long long h[3];
          cudaError_t e = cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("%6s %5d %6d %5d | %10.1f %10.1f\n", names[mode], n, sbo, grid, (double)h[0] / h[2], (double)h[1] / h[2]);
        }
  return 0;
}
