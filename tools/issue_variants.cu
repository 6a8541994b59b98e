// SASS experiment: which way of writing the issue loop gives the fewest instructions per UTCHMMA?
#include <cuda_runtime.h>
#include "../digipathai_b200/csrc/ptx.cuh"
using namespace dp;

__device__ __forceinline__ void mma4_desc64(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t flag0) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      ".reg .b64 a1, b1, a2, b2, a3, b3;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.eq.b32 q, 0, 0;\n\t"
      "add.u64 a1, %1, 2;\n\t add.u64 b1, %2, 2;\n\t"
      "add.u64 a2, %1, 4;\n\t add.u64 b2, %2, 4;\n\t"
      "add.u64 a3, %1, 6;\n\t add.u64 b3, %2, 6;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a1, b1, %3, q;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a2, b2, %3, q;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], a3, b3, %3, q;\n\t"
      "}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(flag0) : "memory");
}

// V1: single thread (lane==0 branch), 64-bit descs, 4-MMA asm block
__global__ void v1(uint64_t a0, uint64_t b0, uint32_t idesc, int sub, int n_tile, int groups, uint64_t asub, uint32_t tm, int* out) {
  if ((threadIdx.x & 31) == 0) {
    for (int g = 0; g < groups; ++g) {
      uint64_t a = a0 + g * 64, b = b0 + g * 512;
      uint32_t d = tm;
      for (int s = 0; s < sub; ++s, a += asub, d += n_tile) mma4_desc64(d, a, b, idesc, g > 0);
    }
  }
  if (out) out[0] = 1;
}

// V2: converged warp + elect_one around the 4-MMA block
__global__ void v2(uint64_t a0, uint64_t b0, uint32_t idesc, int sub, int n_tile, int groups, uint64_t asub, uint32_t tm, int* out) {
  for (int g = 0; g < groups; ++g) {
    uint64_t a = a0 + g * 64, b = b0 + g * 512;
    uint32_t d = tm;
    for (int s = 0; s < sub; ++s, a += asub, d += n_tile) {
      if (elect_one()) mma4_desc64(d, a, b, idesc, g > 0);
      __syncwarp();
    }
  }
  if (out) out[0] = 1;
}

// V3: elect once outside, loops inside
__global__ void v3(uint64_t a0, uint64_t b0, uint32_t idesc, int sub, int n_tile, int groups, uint64_t asub, uint32_t tm, int* out) {
  if (elect_one()) {
    for (int g = 0; g < groups; ++g) {
      uint64_t a = a0 + g * 64, b = b0 + g * 512;
      uint32_t d = tm;
      for (int s = 0; s < sub; ++s, a += asub, d += n_tile) mma4_desc64(d, a, b, idesc, g > 0);
    }
  }
  if (out) out[0] = 1;
}
