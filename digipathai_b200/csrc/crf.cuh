// Fully connected CRF refinement of a probability tile: mean-field inference with Gaussian edge potentials, the
// model DigiPathAI/helpers/utils.py:568-603 (post_process_crf) configures through pydensecrf:
//   unary U = -log(clip(p, 1e-5, 1));  pairwise 1: Gaussian on (y, x) / 10, Potts weight 3;
//   pairwise 2: bilateral on ((y, x) / 50, rgb / 20), Potts weight 10;  DIAG_KERNEL, NORMALIZE_SYMMETRIC;
//   10 iterations of  Q <- softmax(-U + sum_k w_k n_k * (K_k (n_k * Q))),  n_k = 1 / sqrt(K_k 1 + 1e-20);  MAP = argmax.
// pydensecrf approximates K Q with a permutohedral lattice; these kernels evaluate the Gaussian filters exactly:
// the spatial kernel is separable (two 1-D passes over the whole tile), the bilateral kernel is a brute-force
// all-pairs sum (N^2 = 4.3e9 pair weights per application on a 256x256 tile; exp on the SFUs, tiles of 256
// source pixels staged in shared memory) -- ~11 applications per tile, tens of milliseconds on a B200.
// Workspace: 19 float planes of h*w per tile (layout below), supplied by the caller.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace dp {

enum : int { CRF_U0 = 0, CRF_U1, CRF_Q0, CRF_Q1, CRF_AG0, CRF_AG1, CRF_AB0, CRF_AB1, CRF_T0, CRF_T1, CRF_G0, CRF_G1,
             CRF_B0, CRF_B1, CRF_N1, CRF_N2, CRF_F0 /* .. CRF_F0 + 4 */, CRF_PLANES = CRF_F0 + 5 };

__device__ __forceinline__ float* crf_plane(float* ws, long long npix, int tile, int plane) {
  return ws + (static_cast<long long>(tile) * CRF_PLANES + plane) * npix;
}

// unary, initial Q = softmax(-U), bilateral features
__global__ void crf_init_kernel(const uint8_t* __restrict__ rgb, const float* __restrict__ p1, int n_tiles, int h, int w,
                                float inv_sb, float inv_cb, float* __restrict__ ws) {
  const long long npix = static_cast<long long>(h) * w;
  const long long total = npix * n_tiles;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = idx / npix;
    const long long i = idx - t * npix;
    const int y = i / w, x = i - static_cast<long long>(y) * w;
    const float p = p1[idx];
    const float c0 = fminf(fmaxf(1.f - p, 1e-5f), 1.f), c1 = fminf(fmaxf(p, 1e-5f), 1.f);
    const float u0 = -logf(c0), u1 = -logf(c1);
    crf_plane(ws, npix, t, CRF_U0)[i] = u0;
    crf_plane(ws, npix, t, CRF_U1)[i] = u1;
    const float m = fmaxf(-u0, -u1);
    const float e0 = expf(-u0 - m), e1 = expf(-u1 - m);
    crf_plane(ws, npix, t, CRF_Q0)[i] = e0 / (e0 + e1);
    crf_plane(ws, npix, t, CRF_Q1)[i] = e1 / (e0 + e1);
    const uint8_t* px = rgb + idx * 3;
    crf_plane(ws, npix, t, CRF_F0 + 0)[i] = y * inv_sb;
    crf_plane(ws, npix, t, CRF_F0 + 1)[i] = x * inv_sb;
    crf_plane(ws, npix, t, CRF_F0 + 2)[i] = px[0] * inv_cb;
    crf_plane(ws, npix, t, CRF_F0 + 3)[i] = px[1] * inv_cb;
    crf_plane(ws, npix, t, CRF_F0 + 4)[i] = px[2] * inv_cb;
  }
}

// a_k = n_k * Q for both kernels (first call: Q := 1, n := 1 gives the inputs of the normalisation pass)
__global__ void crf_scale_kernel(int n_tiles, long long npix, float* __restrict__ ws, int ones) {
  const long long total = npix * n_tiles;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = idx / npix;
    const long long i = idx - t * npix;
    const float q0 = ones ? 1.f : crf_plane(ws, npix, t, CRF_Q0)[i], q1 = ones ? 1.f : crf_plane(ws, npix, t, CRF_Q1)[i];
    const float n1 = ones ? 1.f : crf_plane(ws, npix, t, CRF_N1)[i], n2 = ones ? 1.f : crf_plane(ws, npix, t, CRF_N2)[i];
    crf_plane(ws, npix, t, CRF_AG0)[i] = n1 * q0;
    crf_plane(ws, npix, t, CRF_AG1)[i] = n1 * q1;
    crf_plane(ws, npix, t, CRF_AB0)[i] = n2 * q0;
    crf_plane(ws, npix, t, CRF_AB1)[i] = n2 * q1;
  }
}

// one 1-D pass of the separable spatial Gaussian over two planes: out[p] = sum_q exp(-0.5 ((p - q) inv_s)^2) in[q]
__global__ void crf_gauss_pass_kernel(int n_tiles, int h, int w, int along_x, float inv_s, int src0, int dst0,
                                      float* __restrict__ ws) {
  const long long npix = static_cast<long long>(h) * w;
  const long long total = npix * n_tiles;
  const int len = along_x ? w : h;
  const long long step = along_x ? 1 : w;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = idx / npix;
    const long long i = idx - t * npix;
    const int y = i / w, x = i - static_cast<long long>(y) * w;
    const int pos = along_x ? x : y;
    const long long line0 = along_x ? static_cast<long long>(y) * w : x;
    const float* a0 = crf_plane(ws, npix, t, src0) + line0;
    const float* a1 = crf_plane(ws, npix, t, src0 + 1) + line0;
    float s0 = 0.f, s1 = 0.f;
    for (int q = 0; q < len; ++q) {
      const float d = (pos - q) * inv_s;
      const float wgt = __expf(-0.5f * d * d);
      s0 = fmaf(wgt, a0[q * step], s0);
      s1 = fmaf(wgt, a1[q * step], s1);
    }
    crf_plane(ws, npix, t, dst0)[i] = s0;
    crf_plane(ws, npix, t, dst0 + 1)[i] = s1;
  }
}

// all-pairs bilateral filter of the two planes AB0/AB1 -> B0/B1; grid = (ceil(npix / 256), n_tiles), 256 threads
__global__ void __launch_bounds__(256) crf_bilateral_kernel(int h, int w, float* __restrict__ ws) {
  __shared__ float sf[5][256];
  __shared__ float sv[2][256];
  const long long npix = static_cast<long long>(h) * w;
  const int t = blockIdx.y;
  const long long i = blockIdx.x * 256LL + threadIdx.x;
  const bool live = i < npix;
  float f[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) f[k] = live ? crf_plane(ws, npix, t, CRF_F0 + k)[i] : 0.f;
  float s0 = 0.f, s1 = 0.f;
  for (long long j0 = 0; j0 < npix; j0 += 256) {
    const long long j = j0 + threadIdx.x;
    const bool jl = j < npix;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 5; ++k) sf[k][threadIdx.x] = jl ? crf_plane(ws, npix, t, CRF_F0 + k)[j] : 1e18f;
    sv[0][threadIdx.x] = jl ? crf_plane(ws, npix, t, CRF_AB0)[j] : 0.f;
    sv[1][threadIdx.x] = jl ? crf_plane(ws, npix, t, CRF_AB1)[j] : 0.f;
    __syncthreads();
#pragma unroll 8
    for (int q = 0; q < 256; ++q) {
      const float d0 = f[0] - sf[0][q], d1 = f[1] - sf[1][q], d2 = f[2] - sf[2][q], d3 = f[3] - sf[3][q],
                  d4 = f[4] - sf[4][q];
      const float dd = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, d4 * d4))));
      const float wgt = __expf(-0.5f * dd);
      s0 = fmaf(wgt, sv[0][q], s0);
      s1 = fmaf(wgt, sv[1][q], s1);
    }
  }
  if (live) {
    crf_plane(ws, npix, t, CRF_B0)[i] = s0;
    crf_plane(ws, npix, t, CRF_B1)[i] = s1;
  }
}

// after the normalisation pass (inputs were all ones): n_k = 1 / sqrt(K_k 1 + 1e-20)
__global__ void crf_norm_kernel(int n_tiles, long long npix, float* __restrict__ ws) {
  const long long total = npix * n_tiles;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = idx / npix;
    const long long i = idx - t * npix;
    crf_plane(ws, npix, t, CRF_N1)[i] = rsqrtf(crf_plane(ws, npix, t, CRF_G0)[i] + 1e-20f);
    crf_plane(ws, npix, t, CRF_N2)[i] = rsqrtf(crf_plane(ws, npix, t, CRF_B0)[i] + 1e-20f);
  }
}

// Q <- softmax(-U + w_g n1 G + w_b n2 B); on the last iteration also the MAP label and (optionally) Q1
__global__ void crf_update_kernel(int n_tiles, long long npix, float w_g, float w_b, float* __restrict__ ws,
                                  uint8_t* __restrict__ labels, float* __restrict__ q1_out) {
  const long long total = npix * n_tiles;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = idx / npix;
    const long long i = idx - t * npix;
    const float n1 = crf_plane(ws, npix, t, CRF_N1)[i], n2 = crf_plane(ws, npix, t, CRF_N2)[i];
    const float t0 = -crf_plane(ws, npix, t, CRF_U0)[i] + w_g * n1 * crf_plane(ws, npix, t, CRF_G0)[i] +
                     w_b * n2 * crf_plane(ws, npix, t, CRF_B0)[i];
    const float t1 = -crf_plane(ws, npix, t, CRF_U1)[i] + w_g * n1 * crf_plane(ws, npix, t, CRF_G1)[i] +
                     w_b * n2 * crf_plane(ws, npix, t, CRF_B1)[i];
    const float m = fmaxf(t0, t1);
    const float e0 = expf(t0 - m), e1 = expf(t1 - m);
    const float q1 = e1 / (e0 + e1);
    crf_plane(ws, npix, t, CRF_Q0)[i] = e0 / (e0 + e1);
    crf_plane(ws, npix, t, CRF_Q1)[i] = q1;
    if (labels) labels[idx] = (t1 > t0) ? 1 : 0;   // np.argmax: ties go to label 0
    if (q1_out) q1_out[idx] = q1;
  }
}

}  // namespace dp
