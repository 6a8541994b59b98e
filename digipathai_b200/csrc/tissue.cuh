// Tissue mask of the slide's lowest pyramid level on the device (reference: TissueMaskGenerationOS,
// DigiPathAI/helpers/utils.py:336-354, and BinMorphoProcessMaskOS, utils.py:200-219).  HBM-bound byte work:
//
//   tissue_hist_kernel   one pass over the RGB level image: the three per-channel 256-bin histograms (Otsu on R, G, B)
//                        and the joint 256 x 256 histogram of (max, max - min), from which the host derives the
//                        256-bin histogram of the HSV saturation (max-min)/max exactly (one of 65 536 values per
//                        (max, max-min) pair) and its Otsu threshold.  3 B read per pixel.
//   tissue_mask_kernel   mask = S > t_S  and not (R > t_R and G > t_G and B > t_B)  and  R, G, B > 50, with the
//                        saturation test as a 64 KB lookup table indexed by (max, max-min).  3 B in, 1 B out per pixel.
//   morph_line_kernel    one axis of a rectangular k x k dilate / erode (max / min are separable): window
//                        [i - k/2, i - k/2 + k - 1] (OpenCV's default anchor), cells outside the image ignored
//                        (OpenCV's morphologyDefaultBorderValue).  1 B in, 1 B out per pixel and pass.
// The Otsu arithmetic itself (256 bins, float64) stays on the host: it is a few microseconds of work on 66 k counters.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

constexpr int kTissueHistBins = 768 + 65536;

__global__ void tissue_hist_kernel(const uint8_t* __restrict__ rgb, long long n_pix, unsigned int* __restrict__ hist) {
  __shared__ unsigned int s_ch[768];
  for (int i = threadIdx.x; i < 768; i += blockDim.x) s_ch[i] = 0;
  __syncthreads();
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n_pix;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint8_t* px = rgb + p * 3;
    const unsigned r = px[0], g = px[1], b = px[2];
    atomicAdd(&s_ch[r], 1u);
    atomicAdd(&s_ch[256 + g], 1u);
    atomicAdd(&s_ch[512 + b], 1u);
    const unsigned v = max(max(r, g), b), mn = min(min(r, g), b);
    atomicAdd(&hist[768 + v * 256 + (v - mn)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 768; i += blockDim.x)
    if (s_ch[i]) atomicAdd(&hist[i], s_ch[i]);
}

__global__ void tissue_mask_kernel(const uint8_t* __restrict__ rgb, long long n_pix, int thr_r, int thr_g, int thr_b,
                                   int rgb_min, const uint8_t* __restrict__ sat_lut, uint8_t* __restrict__ mask) {
  for (long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; p < n_pix;
       p += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint8_t* px = rgb + p * 3;
    const int r = px[0], g = px[1], b = px[2];
    const int v = max(max(r, g), b), mn = min(min(r, g), b);
    const bool tissue_s = sat_lut[v * 256 + (v - mn)] != 0;
    const bool bg = r > thr_r && g > thr_g && b > thr_b;
    const bool above = r > rgb_min && g > rgb_min && b > rgb_min;
    mask[p] = (tissue_s && !bg && above) ? 1 : 0;
  }
}

// in / out: uint8 [n0][n1]; axis 0: the window runs over the slow index, axis 1: over the fast index.
// is_max: dilate, else erode.
__global__ void morph_line_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int n0, int n1, int k,
                                  int axis, int is_max) {
  const long long total = static_cast<long long>(n0) * n1;
  const int a = k / 2;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int i1 = static_cast<int>(idx % n1), i0 = static_cast<int>(idx / n1);
    const int c = axis ? i1 : i0, n = axis ? n1 : n0;
    const long long step = axis ? 1 : n1;
    const int lo = max(c - a, 0), hi = min(c - a + k - 1, n - 1);
    const uint8_t* src = in + idx + static_cast<long long>(lo - c) * step;
    int acc = is_max ? 0 : 255;
    for (int j = lo; j <= hi; ++j, src += step) acc = is_max ? max(acc, static_cast<int>(*src)) : min(acc, static_cast<int>(*src));
    out[idx] = static_cast<uint8_t>(acc);
  }
}

}  // namespace dp
