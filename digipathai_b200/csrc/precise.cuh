// fp32 precision mode (`precision = fp32` in the model container): the same layer program evaluated with fp32
// weights, fp32 activations and fp32 FMA accumulation on the CUDA cores.
//
// Why it exists: BASELINE.json asks for |p - p_reference| <= 1e-3.  With 16-bit storage that bound is a property of
// the network instance, not of the kernels (profiles/r2_parity_conditioning.md: ONE 2^-11 relative rounding of the
// stem output moves the stand-in network's output by 2e-2), so the library offers a mode whose operands carry a
// 24-bit mantissa everywhere.  It is the configuration the 1e-3 parity tests assert, and it is slower than the
// tcgen05 path by the fp32-FMA : fp16-tensor ratio of the machine -- bench.py reports its tiles/s beside the fp16 line.
//
// One kernel covers every conv of the three graphs (1x1, 3x3, up2 sub-pixel phases, generic kh x kw taps, stride 2,
// pre-activation BN(+ReLU) prologue, shift / ReLU / residual epilogue): an implicit GEMM over
// (output pixels) x (Cout) x (taps * Cin) with a 128 x 64 (or 128 x 32) output tile per CTA, 16-channel K slices
// staged through shared memory and an 8 x 4 (8 x 2) register tile per thread.  The operand description is the same
// NaiveConvParams the debug path of the fp16 build uses; pointers are reinterpreted as float.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "conv_tc.cuh"

namespace dp {

constexpr int kPcBM = 128, kPcBK = 16, kPcThreads = 256, kPcAStride = kPcBM + 4;

template <int BN>
__global__ void __launch_bounds__(kPcThreads) conv_f32_kernel(const NaiveConvParams p) {
  constexpr int TN = BN / 16;                 // couts per thread (4 or 2)
  constexpr int BStride = BN + 4;
  __shared__ __align__(16) float As[kPcBK][kPcAStride];
  __shared__ __align__(16) float Bs[kPcBK][BStride];
  const float* __restrict__ in = reinterpret_cast<const float*>(p.in);
  const float* __restrict__ wgt = reinterpret_cast<const float*>(p.w);
  float* __restrict__ out = reinterpret_cast<float*>(p.out);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;     // 16 cout groups x 16 pixel groups of 8
  const long long M = static_cast<long long>(p.n_img) * p.OH * p.OW;
  const long long m0 = static_cast<long long>(blockIdx.x) * kPcBM;
  const int n0 = blockIdx.y * BN;
  const int g = blockIdx.z;                   // accumulator group (up2 sub-pixel phase), else 0

  // the pixel this thread stages into shared memory (two threads per pixel: 8 channels each)
  const int lp = tid >> 1, lc = (tid & 1) * 8;
  const long long lm = m0 + lp;
  const bool lvalid = lm < M;
  int ln = 0, loh = 0, low = 0;
  if (lvalid) {
    low = static_cast<int>(lm % p.OW);
    const long long r = lm / p.OW;
    loh = static_cast<int>(r % p.OH);
    ln = static_cast<int>(r / p.OH);
  }
  // the weight row this thread stages: cout = n0 + (tid >> 2), 4 channels at (tid & 3) * 4  (BN = 64);
  // BN = 32: threads 0..127 only
  const int wr = tid >> 2, wc = (tid & 3) * 4;
  const bool wvalid = wr < BN && (n0 + wr) < p.Cout;

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int e = 0; e < p.n_entries_total; ++e) {
    const TapEntry ent = p.entries[e];
    if (ent.group != g) continue;
    const int ih = loh * p.stride + ent.dy, iw = low * p.stride + ent.dx;
    const bool inb = lvalid && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
    const float* arow = in + ((static_cast<long long>(ln) * p.H + ih) * p.W + iw) * p.in_ctot + p.in_choff;
    const float* wrow = wgt + (static_cast<long long>(e) * p.Cout + (n0 + wr)) * p.Cin;
    for (int c0 = 0; c0 < p.Cin; c0 += kPcBK) {
      // ---- stage A: 128 pixels x 16 channels, pre-activation applied, zero outside the image (the conv's padding
      //      acts on the activated tensor) and beyond Cin
      float a[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) a[t] = 0.f;
      if (inb && c0 + lc < p.Cin) {           // Cin is a multiple of 8
        const float4 v0 = *reinterpret_cast<const float4*>(arow + c0 + lc);
        const float4 v1 = *reinterpret_cast<const float4*>(arow + c0 + lc + 4);
        a[0] = v0.x; a[1] = v0.y; a[2] = v0.z; a[3] = v0.w;
        a[4] = v1.x; a[5] = v1.y; a[6] = v1.z; a[7] = v1.w;
        if (p.pro_mode) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            a[t] = fmaf(a[t], p.pro_scale[c0 + lc + t], p.pro_shift[c0 + lc + t]);
            if (p.pro_mode == 2) a[t] = fmaxf(a[t], 0.f);
          }
        }
      }
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (wvalid && c0 + wc < p.Cin) b = *reinterpret_cast<const float4*>(wrow + c0 + wc);   // Cin % 4 == 0
      __syncthreads();                        // previous slice fully consumed
#pragma unroll
      for (int t = 0; t < 8; ++t) As[lc + t][lp] = a[t];
      if (wr < BN) {
        Bs[wc + 0][wr] = b.x; Bs[wc + 1][wr] = b.y; Bs[wc + 2][wr] = b.z; Bs[wc + 3][wr] = b.w;
      }
      __syncthreads();
      // ---- 16 rank-1 updates of the 8 x TN register tile
#pragma unroll
      for (int k = 0; k < kPcBK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float bv[TN];
        if constexpr (TN == 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
          bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
        } else {
          const float2 b2 = *reinterpret_cast<const float2*>(&Bs[k][tx * 2]);
          bv[0] = b2.x; bv[1] = b2.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }

  // ---- epilogue: BN affine (scale is folded into the weights by the builders; kept for generality), residual, ReLU
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long mm = m0 + ty * 8 + i;
    if (mm >= M) continue;
    const int ow = static_cast<int>(mm % p.OW);
    const long long r = mm / p.OW;
    const int oh = static_cast<int>(r % p.OH);
    const int n = static_cast<int>(r / p.OH);
    long long opix;
    if (p.up2)
      opix = (static_cast<long long>(n) * 2 * p.H + 2 * oh + (g >> 1)) * (2 * p.W) + 2 * ow + (g & 1);
    else
      opix = (static_cast<long long>(n) * p.OH + oh) * p.OW + ow;
    float* orow = out + opix * p.out_ctot + p.out_choff;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int co = n0 + tx * TN + j;
      if (co >= p.Cout) continue;
      float y = fmaf(acc[i][j], p.epi_scale ? p.epi_scale[co] : 1.f, p.epi_shift ? p.epi_shift[co] : 0.f);
      if (p.residual) y += orow[co];
      if (p.relu) y = fmaxf(y, 0.f);
      orow[co] = y;
    }
  }
}

}  // namespace dp
