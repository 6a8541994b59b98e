// HBM-bound helper kernels around the tensor-core convs: tile gather + stem im2col, max/avg pooling with the
// pre-activation BN, the naive head, and the slide-plane kernels (stitch / normalise / threshold / pyramid).
// All activations are NHWC fp16 with an explicit channel stride so that concat buffers are written in place.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>
#include "d4.cuh"
#include "ptx.cuh"

namespace dp {

// Activations are stored as fp16 (default) or fp32 (precision mode, precise.cuh); the helper kernels below are
// templates over the storage type and move 8 channels per thread either way (one 16-byte or two 16-byte vectors).
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __half2* hv = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const float2 f = __half22float2(hv[t]);
    v[2 * t] = f.x;
    v[2 * t + 1] = f.y;
  }
}
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  __align__(16) __half2 o[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) o[t] = __floats2half2_rn(v[2 * t], v[2 * t + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<const uint4*>(o);
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// Tile origins are clamped into the raster exactly as the reference clamps them (dataloader.py:351-353:
// x = max(0, min(x, X_slide - P))), so a caller of the C ABI that passes an origin outside the slide reads the border
// tile instead of out-of-bounds memory.  Callers on the path (tissue.TileGrid) pass clamped origins already.
__device__ __forceinline__ long long clamp_origin(long long v, long long extent, int P) {
  const long long hi = extent - P;
  return v < 0 ? 0 : (v > hi ? (hi > 0 ? hi : 0) : v);
}

// ---------------------------------------------------------------------------------------------------------
// Tile gather + (v-128)/128 + forward TTA + ZeroPadding2D(3) + im2col for the 7x7/2 stem conv
// (dataloader.py:357-388 crop/transposed layout/normalise; utils.py:487-501 TTA; densenet.py:116-117 stem).
// slide is uint8 [Wslide][Hslide][3] in the reference's [x][y] orientation; tile b has origin coords[b] = (x,y).
// out is fp16 [B][P/2][P/2][160], column k = (ky*7 + kx)*3 + c for k < 147, zero above.
__global__ void stem_im2col_kernel(const PassDesc* __restrict__ pass, int img0, int B, int P,
                                   __half* __restrict__ out) {
  const uint8_t* __restrict__ slide = pass->slide;
  const long long slide_h = pass->slide_h;
  const int* __restrict__ coords = pass->coords + 2 * img0;
  const int tta_code = pass->tta_in;
  const int OH = P / 2;
  const long long total = static_cast<long long>(B) * OH * OH * 20;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int q = idx % 20;
    long long r = idx / 20;
    const int ow = r % OH; r /= OH;
    const int oh = r % OH;
    const int b = r / OH;
    const long long x0 = clamp_origin(coords[2 * b], pass->slide_w, P), y0 = clamp_origin(coords[2 * b + 1], slide_h, P);
    __align__(16) __half vals[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int k = q * 8 + t;
      float v = 0.f;
      if (k < 147) {
        const int ky = k / 21, rem = k - ky * 21;
        const int kx = rem / 3, c = rem - kx * 3;
        const int i = 2 * oh + ky - 3, j = 2 * ow + kx - 3;
        if (i >= 0 && i < P && j >= 0 && j < P) {
          int ti, tj;
          d4_src(tta_code, i, j, P, ti, tj);
          const uint8_t px = slide[((x0 + ti) * slide_h + (y0 + tj)) * 3 + c];
          v = (static_cast<float>(px) - 128.f) * (1.f / 128.f);  // exact in fp16
        }
      }
      vals[t] = __float2half_rn(v);
    }
    *reinterpret_cast<uint4*>(out + idx * 8) = *reinterpret_cast<const uint4*>(vals);
  }
}

// ---------------------------------------------------------------------------------------------------------
// Stem input, space-to-depth form.  The 7x7 stride-2 conv on 3 channels (densenet.py:116-117) equals a 4x4
// stride-1 conv on the 2x2 space-to-depth image (12 channels).  To keep 128-byte (64-channel) operand rows for
// the tensor-core kernel, the 4 column taps are unrolled into the channel axis here, leaving 4 row taps:
//   out[b][r][q][dq*16 + (a*2+b2)*3 + c] = net_in[2r+a][2(q+dq-2)+b2][c]     (0 outside the tile / ch >= 12)
// with net_in = forward-TTA'd, (v-128)/128-normalised tile cropped from the slide raster.  fp16 [B][P/2][P/2][64].
template <typename T>
__global__ void stem_s2d_kernel(const PassDesc* __restrict__ pass, int img0, int B, int P, T* __restrict__ out) {
  pdl_wait();               // launched with programmatic stream serialization: inputs complete from here
  pdl_launch_dependents();
  const uint8_t* __restrict__ slide = pass->slide;
  const long long slide_h = pass->slide_h;
  const int* __restrict__ coords = pass->coords + 2 * img0;
  const int tta_code = pass->tta_in;
  const int OH = P / 2;
  const unsigned total = static_cast<unsigned>(B) * OH * OH * 4;   // < 2^31 for any supported batch / patch
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int dq = idx & 3;
    unsigned r0 = idx >> 2;
    const int q = r0 % OH; r0 /= OH;
    const int r = r0 % OH;
    const int b = r0 / OH;
    const long long x0 = clamp_origin(coords[2 * b], pass->slide_w, P), y0 = clamp_origin(coords[2 * b + 1], slide_h, P);
    float lo8[8], hi8[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) lo8[t] = hi8[t] = 0.f;
    const int jq = q + dq - 2;
    if (jq >= 0 && jq < OH) {
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
          int ti, tj;
          d4_src(tta_code, 2 * r + a, 2 * jq + b2, P, ti, tj);
          const uint8_t* px = slide + ((x0 + ti) * slide_h + (y0 + tj)) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int k = (a * 2 + b2) * 3 + c;                                          // compile-time after unrolling
            const float v = (static_cast<float>(px[c]) - 128.f) * (1.f / 128.f);        // exact in fp16
            if (k < 8) lo8[k] = v; else hi8[k - 8] = v;
          }
        }
    }
    store8(out + static_cast<size_t>(idx) * 16, lo8);
    store8(out + static_cast<size_t>(idx) * 16 + 8, hi8);
  }
}

// ---------------------------------------------------------------------------------------------------------
// mode 0: ZeroPadding2D(1) + MaxPooling2D(3, strides=2) (densenet.py:122-123): window starts at 2*o - 1 and the
//         explicit zero padding takes part in the max.
// mode 1: MaxPooling2D(3, strides=2, padding='same') (inception.py:178,182,211,231) on an even-sized map:
//         TensorFlow pads 0 in front / 1 behind, window starts at 2*o, padded cells never win.
// 8 channels per thread.
template <typename T>
__global__ void maxpool3s2_kernel(const T* __restrict__ in, int in_ctot, int in_choff,
                                  T* __restrict__ out, int out_ctot, int out_choff, int n_img, int H, int W,
                                  int C, int mode) {
  pdl_wait();               // launched with programmatic stream serialization: inputs complete from here
  pdl_launch_dependents();
  const int OH = H / 2, OW = W / 2, CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * OH * OW * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    unsigned r = idx / CG;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH;
    const int n = r / OH;
    float m[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) m[t] = -INFINITY;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      // one window row at a time: its three loads are in flight together (latency bound otherwise)
      const int ih = 2 * oh - (mode ? 0 : 1) + ky;
      const bool row_in = ih >= 0 && ih < H;
      float f[3][8];
      bool in_map[3];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = 2 * ow - (mode ? 0 : 1) + kx;
        in_map[kx] = row_in && iw >= 0 && iw < W;
        if (in_map[kx]) load8(in + ((static_cast<long long>(n) * H + ih) * W + iw) * in_ctot + in_choff + cg * 8, f[kx]);
      }
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if (in_map[kx]) {
#pragma unroll
          for (int t = 0; t < 8; ++t) m[t] = fmaxf(m[t], f[kx][t]);
        } else if (!mode) {
#pragma unroll
          for (int t = 0; t < 8; ++t) m[t] = fmaxf(m[t], 0.f);  // explicit zero padding takes part in the max
        }
      }
    }
    store8(out + ((static_cast<long long>(n) * OH + oh) * OW + ow) * out_ctot + out_choff + cg * 8, m);
  }
}

// ---------------------------------------------------------------------------------------------------------
// AveragePooling2D(3, strides=1, padding='same') (inception.py:193): mean over the VALID cells of the window.
template <typename T>
__global__ void avgpool3s1_kernel(const T* __restrict__ in, int in_ctot, int in_choff,
                                  T* __restrict__ out, int out_ctot, int out_choff, int n_img, int H, int W,
                                  int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * H * W * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    unsigned r = idx / CG;
    const int ow = r % W; r /= W;
    const int oh = r % H;
    const int n = r / H;
    float acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] = 0.f;
    int cnt = 0;
    for (int ky = -1; ky <= 1; ++ky)
      for (int kx = -1; kx <= 1; ++kx) {
        const int ih = oh + ky, iw = ow + kx;
        if (ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
        ++cnt;
        float f[8];
        load8(in + ((static_cast<long long>(n) * H + ih) * W + iw) * in_ctot + in_choff + cg * 8, f);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc[t] += f[t];
      }
    const float inv = 1.f / static_cast<float>(cnt);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] *= inv;
    store8(out + ((static_cast<long long>(n) * H + oh) * W + ow) * out_ctot + out_choff + cg * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// y = scale*x + shift [, ReLU] [, AveragePooling2D(2,2)]   (transition_block densenet.py:101-107 with the
// pool commuted in front of the 1x1 conv -- both are linear -- and the final `bn` densenet.py:134).
template <typename T>
__global__ void bn_act_pool_kernel(const T* __restrict__ in, int in_ctot, int in_choff,
                                   T* __restrict__ out, int out_ctot, int out_choff, int n_img, int H, int W,
                                   int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                   int relu, int pool) {
  pdl_wait();               // launched with programmatic stream serialization: inputs complete from here
  pdl_launch_dependents();
  const int OH = pool ? H / 2 : H, OW = pool ? W / 2 : W, CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * OH * OW * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    unsigned r = idx / CG;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH;
    const int n = r / OH;
    float sc[8], sh[8], acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { sc[t] = scale[cg * 8 + t]; sh[t] = shift[cg * 8 + t]; acc[t] = 0.f; }
    const T* base = in + ((static_cast<long long>(n) * H + (pool ? 2 * oh : oh)) * W + (pool ? 2 * ow : ow)) * in_ctot +
                    in_choff + cg * 8;
    if (pool) {
      // all four window loads in flight before the first use (the kernel is latency bound otherwise)
      float f[4][8];
      load8(base, f[0]);
      load8(base + in_ctot, f[1]);
      load8(base + static_cast<long long>(W) * in_ctot, f[2]);
      load8(base + static_cast<long long>(W + 1) * in_ctot, f[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          float a = fmaf(f[q][t], sc[t], sh[t]);
          if (relu) a = fmaxf(a, 0.f);
          acc[t] += a;
        }
    } else {
      float f[8];
      load8(base, f);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        float a = fmaf(f[t], sc[t], sh[t]);
        if (relu) a = fmaxf(a, 0.f);
        acc[t] += a;
      }
    }
    const float inv = pool ? 0.25f : 1.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] *= inv;
    store8(out + ((static_cast<long long>(n) * OH + oh) * OW + ow) * out_ctot + out_choff + cg * 8, acc);
  }
}

// ---------------------------------------------------------------------------------------------------------
// DeepLabv3+ helpers (reference: DigiPathAI/models/deeplabv3.py).  All HBM-bound, 8 channels (16 B) per thread.
//
// [ReLU] -> DepthwiseConv2D(3x3, stride 1|2, dilation `rate`) -> + shift (BN scale folded into w) [-> ReLU]
// (SepConv_BN, deeplabv3.py:62-79).  Padding is symmetric `rate` on each side for both strides (stride 1:
// padding='same'; stride 2: explicit ZeroPadding2D((1,1)) + 'valid'), so input pixel = stride * o + (k - 1) * rate.
// Measured (ncu, round 1): instruction-issue bound (80 % issue-active, ~1.3 TB/s), not HBM bound -- fp16 unpack +
// fp32 FMA per tap.  A 4-pixels-per-thread register-window variant (half the loads) was NOT faster (128 registers,
// 22 % occupancy, latency bound) and was dropped; so was an instruction-lean variant (fp32 weights, half2 ReLU,
// template-resolved activations, 48 registers): also not faster -- the issue slots are not what it waits on after all.
template <typename T>
__global__ void dwconv3x3_kernel(const T* __restrict__ in, int in_ctot, int in_choff, T* __restrict__ out,
                                 int out_ctot, int out_choff, int n_img, int H, int W, int C, int stride, int rate,
                                 const T* __restrict__ w, const float* __restrict__ shift, int pre_relu,
                                 int post_relu) {
  pdl_wait();
  pdl_launch_dependents();
  const int OH = H / stride, OW = W / stride, CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * OH * OW * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    unsigned r = idx / CG;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH;
    const int n = r / OH;
    float acc[8];
    {
      const float4 s0 = *reinterpret_cast<const float4*>(shift + cg * 8);
      const float4 s1 = *reinterpret_cast<const float4*>(shift + cg * 8 + 4);
      acc[0] = s0.x; acc[1] = s0.y; acc[2] = s0.z; acc[3] = s0.w;
      acc[4] = s1.x; acc[5] = s1.y; acc[6] = s1.z; acc[7] = s1.w;
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int ih = oh * stride + (ky - 1) * rate;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int iw = ow * stride + (kx - 1) * rate;
        if (iw < 0 || iw >= W) continue;
        float x[8], k[8];
        load8(in + ((static_cast<long long>(n) * H + ih) * W + iw) * in_ctot + in_choff + cg * 8, x);
        load8(w + (ky * 3 + kx) * C + cg * 8, k);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          if (pre_relu) x[t] = fmaxf(x[t], 0.f);
          acc[t] = fmaf(x[t], k[t], acc[t]);
        }
      }
    }
    if (post_relu) {
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] = fmaxf(acc[t], 0.f);
    }
    store8(out + ((static_cast<long long>(n) * OH + oh) * OW + ow) * out_ctot + out_choff + cg * 8, acc);
  }
}

// GlobalAveragePooling2D (deeplabv3.py:378): [n][H][W][C] -> [n][1][1][C], fp32 accumulation.
template <typename T>
__global__ void global_avgpool_kernel(const T* __restrict__ in, int in_ctot, int in_choff,
                                      T* __restrict__ out, int out_ctot, int out_choff, int n_img, int HW, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG, n = idx / CG;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const T* src = in + static_cast<long long>(n) * HW * in_ctot + in_choff + cg * 8;
    for (int p = 0; p < HW; ++p) {
      float f[8];
      load8(src + static_cast<long long>(p) * in_ctot, f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc[t] += f[t];
    }
    const float inv = 1.f / static_cast<float>(HW);
#pragma unroll
    for (int t = 0; t < 8; ++t) acc[t] *= inv;
    store8(out + static_cast<long long>(n) * out_ctot + out_choff + cg * 8, acc);
  }
}

// Bilinear align_corners resize of a 1x1 map == broadcast (deeplabv3.py:385-388): [n][1][1][C] -> [n][H][W][C range].
template <typename T>
__global__ void broadcast_kernel(const T* __restrict__ in, int in_ctot, int in_choff, T* __restrict__ out,
                                 int out_ctot, int out_choff, int n_img, int HW, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * HW * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    const unsigned r = idx / CG;
    const int n = r / HW;
    float v[8];
    load8(in + static_cast<long long>(n) * in_ctot + in_choff + cg * 8, v);
    store8(out + static_cast<long long>(r) * out_ctot + out_choff + cg * 8, v);
  }
}

// tf.compat.v1.image.resize(method='bilinear', align_corners=True) (deeplabv3.py:420): src = dst * (in-1)/(out-1),
// top = floor, bottom = min(top + 1, in - 1), fp32 lerp (rows first, then columns, as TF's kernel does).
__device__ __forceinline__ void bilinear_ac_coord(int o, int in_size, int out_size, int& i0, int& i1, float& f) {
  const float scale = (out_size > 1) ? static_cast<float>(in_size - 1) / static_cast<float>(out_size - 1) : 0.f;
  const float s = static_cast<float>(o) * scale;
  i0 = static_cast<int>(floorf(s));
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = (i0 + 1 < in_size) ? i0 + 1 : in_size - 1;
  f = s - static_cast<float>(i0);
}

template <typename T>
__global__ void resize_bilinear_kernel(const T* __restrict__ in, int in_ctot, int in_choff,
                                       T* __restrict__ out, int out_ctot, int out_choff, int n_img, int H, int W,
                                       int OH, int OW, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int CG = C / 8;
  const unsigned total = static_cast<unsigned>(n_img) * OH * OW * CG;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int cg = idx % CG;
    unsigned r = idx / CG;
    const int ow = r % OW; r /= OW;
    const int oh = r % OH;
    const int n = r / OH;
    int y0, y1, x0, x1;
    float fy, fx;
    bilinear_ac_coord(oh, H, OH, y0, y1, fy);
    bilinear_ac_coord(ow, W, OW, x0, x1, fx);
    const T* base = in + static_cast<long long>(n) * H * W * in_ctot + in_choff + cg * 8;
    float a[8], b[8], c[8], d[8], o[8];
    load8(base + (static_cast<long long>(y0) * W + x0) * in_ctot, a);
    load8(base + (static_cast<long long>(y0) * W + x1) * in_ctot, b);
    load8(base + (static_cast<long long>(y1) * W + x0) * in_ctot, c);
    load8(base + (static_cast<long long>(y1) * W + x1) * in_ctot, d);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const float top = a[t] + (b[t] - a[t]) * fx, bot = c[t] + (d[t] - c[t]) * fx;
      o[t] = top + (bot - top) * fy;
    }
    store8(out + ((static_cast<long long>(n) * OH + oh) * OW + ow) * out_ctot + out_choff + cg * 8, o);
  }
}

// Logits conv collapsed to the class-1-minus-class-0 difference (softmax over 2 classes == sigmoid of it):
// one warp per pixel, lanes split the channels 8 at a time, warp-shuffle reduction, fp32 result.
template <typename T>
__global__ void head_dot_kernel(const T* __restrict__ in, int in_ctot, int in_choff, int C, long long n_pix,
                                const float* __restrict__ w, float bias, float* __restrict__ out, int out_stride_f) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
  const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
  for (long long pix = warp0; pix < n_pix; pix += n_warps) {
    const T* src = in + pix * in_ctot + in_choff;
    float acc = 0.f;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      float f[8];
      load8(src + c0, f);
#pragma unroll
      for (int t = 0; t < 8; ++t) acc = fmaf(f[t], w[c0 + t], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[pix * out_stride_f] = acc + bias;
  }
}

// Final bilinear align_corners resize of the logit difference to PxP + sigmoid + inverse TTA (deeplabv3.py:440-456,
// Segmentation.py:158-167).
__global__ void head_resize_kernel(const float* __restrict__ z, int z_stride_f, int n_img, int h, int w, int P,
                                   const PassDesc* __restrict__ pass, int img0) {
  pdl_wait();
  pdl_launch_dependents();
  const int tta_code = pass->tta_out;
  float* __restrict__ out = pass->probs_out + static_cast<long long>(img0) * P * P;
  const long long total = static_cast<long long>(n_img) * P * P;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = idx % P;
    const int i = (idx / P) % P;
    const int n = idx / (static_cast<long long>(P) * P);
    int y0, y1, x0, x1;
    float fy, fx;
    bilinear_ac_coord(i, h, P, y0, y1, fy);
    bilinear_ac_coord(j, w, P, x0, x1, fx);
    const float* base = z + static_cast<long long>(n) * h * w * z_stride_f;
    const float a = base[(static_cast<long long>(y0) * w + x0) * z_stride_f];
    const float b = base[(static_cast<long long>(y0) * w + x1) * z_stride_f];
    const float c = base[(static_cast<long long>(y1) * w + x0) * z_stride_f];
    const float d = base[(static_cast<long long>(y1) * w + x1) * z_stride_f];
    const float top = a + (b - a) * fx, bot = c + (d - c) * fx;
    const float zz = top + (bot - top) * fy;
    int di, dj;
    d4_src(tta_code, i, j, P, di, dj);
    out[(static_cast<long long>(n) * P + di) * P + dj] = 1.f / (1.f + expf(-zz));
  }
}

// ---------------------------------------------------------------------------------------------------------
// Head for the naive (debug) path: 1x1 conv C->2 + softmax channel 1 + inverse TTA (densenet.py:156).
template <typename T>
__global__ void head_naive_kernel(const T* __restrict__ in, int in_ctot, int in_choff, int C, int n_img,
                                  int P, const float* __restrict__ w, float bias,
                                  const PassDesc* __restrict__ pass, int img0) {
  const int tta_code = pass->tta_out;
  float* __restrict__ out = pass->probs_out + static_cast<long long>(img0) * P * P;
  const long long total = static_cast<long long>(n_img) * P * P;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int wq = idx % P;
    const int h = (idx / P) % P;
    const int n = idx / (static_cast<long long>(P) * P);
    const T* a = in + idx * in_ctot + in_choff;
    float z = 0.f;
    for (int c = 0; c < C; ++c) z = fmaf(static_cast<float>(a[c]), w[c], z);
    z += bias;
    int di, dj;
    d4_src(tta_code, h, wq, P, di, dj);
    out[(static_cast<long long>(n) * P + di) * P + dj] = 1.f / (1.f + expf(-z));
  }
}

// ---------------------------------------------------------------------------------------------------------
// TTA/ensemble statistics + overlap accumulate (Segmentation.py:162-173).
// probs is fp32 [N][B][P][P] (N = |tta| * |models| passes of this batch), planes are [W][H] (x-major).
// For every plane pixel touched by this batch exactly ONE thread (the lowest-index covering tile's thread)
// walks the covering tiles in ascending tile order and does the reference's sequential `+=`, so the result
// is bit-identical to the numpy loop given identical probabilities; no atomics.
// numpy semantics reproduced: mean = (sequential fp32 sum over N) / N; var = sum((p-mean)^2)/N (ddof 0);
// count is uint8 and wraps.
// One covering tile's contribution to one pixel: mean and population variance over the N passes, in numpy's order.
__device__ __forceinline__ void stitch_stats(const float* __restrict__ pp, int N, long long pass_stride, float& mu,
                                             float& vr) {
  float s = pp[0];
  for (int k = 1; k < N; ++k) s = __fadd_rn(s, pp[k * pass_stride]);
  mu = __fdiv_rn(s, static_cast<float>(N));
  const float d0 = __fsub_rn(pp[0], mu);
  float ss = __fmul_rn(d0, d0);
  for (int k = 1; k < N; ++k) {
    const float d = __fsub_rn(pp[k * pass_stride], mu);
    ss = __fadd_rn(ss, __fmul_rn(d, d));
  }
  vr = __fdiv_rn(ss, static_cast<float>(N));
}

// Vector path (every tile origin's y, the patch and the plane height are multiples of 4 -- any stride-128 grid):
// a thread owns 4 consecutive y of one plane row; the four pixels are covered by exactly the same tiles, so the
// ownership test runs once per group and every access is 16 bytes (probabilities, mean, var) or 4 bytes (count).
// Scalar path: one pixel per thread, any geometry.  Both do the same arithmetic in the same order.
__global__ void stitch_kernel(const float* __restrict__ probs, int N, int B, int P,
                              const int* __restrict__ coords, float* __restrict__ mean,
                              float* __restrict__ var, uint8_t* __restrict__ count, long long plane_w,
                              long long plane_h, int x_lo) {
  extern __shared__ int s_xy[];  // [B][2]
  int ok = ((P | static_cast<int>(plane_h & 3)) & 3) == 0;
  for (int i = threadIdx.x; i < 2 * B; i += blockDim.x) {
    const int v = coords[i];
    s_xy[i] = v;
    if ((i & 1) && (v & 3)) ok = 0;
  }
  const bool vec = __syncthreads_and(ok) != 0;
  const int i = blockIdx.y;
  const int xi = s_xy[2 * i], yi = s_xy[2 * i + 1];
  const long long pass_stride = static_cast<long long>(B) * P * P;
  if (vec) {
    const int PQ = P / 4;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < P * PQ; q += gridDim.x * blockDim.x) {
      const int a = q / PQ, b = (q - a * PQ) * 4;
      const int gx = xi + a, gy = yi + b;
      bool owner = true;
      for (int j = 0; j < i; ++j) {
        const int dx = gx - s_xy[2 * j], dy = gy - s_xy[2 * j + 1];
        if (dx >= 0 && dx < P && dy >= 0 && dy < P) { owner = false; break; }
      }
      if (!owner) continue;
      if (gx < x_lo || gx - x_lo >= plane_w || gy < 0 || gy + 3 >= plane_h) continue;
      const long long off = static_cast<long long>(gx - x_lo) * plane_h + gy;
      float4 m4 = *reinterpret_cast<const float4*>(mean + off);
      float4 v4 = *reinterpret_cast<const float4*>(var + off);
      uchar4 c4 = *reinterpret_cast<const uchar4*>(count + off);
      for (int j = i; j < B; ++j) {
        const int dx = gx - s_xy[2 * j], dy = gy - s_xy[2 * j + 1];
        if (dx < 0 || dx >= P || dy < 0 || dy >= P) continue;
        const float* pp = probs + (static_cast<long long>(j) * P + dx) * P + dy;
        // two sweeps over the passes (the second one hits L1): sequential fp32 sum, then sum of squared deviations
        float4 s4 = __ldg(reinterpret_cast<const float4*>(pp));
        for (int k = 1; k < N; ++k) {
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + k * pass_stride));
          s4.x = __fadd_rn(s4.x, p4.x); s4.y = __fadd_rn(s4.y, p4.y);
          s4.z = __fadd_rn(s4.z, p4.z); s4.w = __fadd_rn(s4.w, p4.w);
        }
        const float fn = static_cast<float>(N);
        const float4 mu4 = make_float4(__fdiv_rn(s4.x, fn), __fdiv_rn(s4.y, fn), __fdiv_rn(s4.z, fn), __fdiv_rn(s4.w, fn));
        float4 ss4 = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < N; ++k) {
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(pp + k * pass_stride));
          const float dx0 = __fsub_rn(p4.x, mu4.x), dx1 = __fsub_rn(p4.y, mu4.y);
          const float dx2 = __fsub_rn(p4.z, mu4.z), dx3 = __fsub_rn(p4.w, mu4.w);
          if (k == 0) {
            ss4 = make_float4(__fmul_rn(dx0, dx0), __fmul_rn(dx1, dx1), __fmul_rn(dx2, dx2), __fmul_rn(dx3, dx3));
          } else {
            ss4.x = __fadd_rn(ss4.x, __fmul_rn(dx0, dx0)); ss4.y = __fadd_rn(ss4.y, __fmul_rn(dx1, dx1));
            ss4.z = __fadd_rn(ss4.z, __fmul_rn(dx2, dx2)); ss4.w = __fadd_rn(ss4.w, __fmul_rn(dx3, dx3));
          }
        }
        m4.x = __fadd_rn(m4.x, mu4.x); m4.y = __fadd_rn(m4.y, mu4.y);
        m4.z = __fadd_rn(m4.z, mu4.z); m4.w = __fadd_rn(m4.w, mu4.w);
        v4.x = __fadd_rn(v4.x, __fdiv_rn(ss4.x, fn)); v4.y = __fadd_rn(v4.y, __fdiv_rn(ss4.y, fn));
        v4.z = __fadd_rn(v4.z, __fdiv_rn(ss4.z, fn)); v4.w = __fadd_rn(v4.w, __fdiv_rn(ss4.w, fn));
        c4.x = static_cast<uint8_t>(c4.x + 1); c4.y = static_cast<uint8_t>(c4.y + 1);
        c4.z = static_cast<uint8_t>(c4.z + 1); c4.w = static_cast<uint8_t>(c4.w + 1);
      }
      *reinterpret_cast<float4*>(mean + off) = m4;
      *reinterpret_cast<float4*>(var + off) = v4;
      *reinterpret_cast<uchar4*>(count + off) = c4;
    }
    return;
  }
  for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < P * P; pix += gridDim.x * blockDim.x) {
    const int a = pix / P, b = pix - a * P;
    const int gx = xi + a, gy = yi + b;
    bool owner = true;
    for (int j = 0; j < i; ++j) {
      const int dx = gx - s_xy[2 * j], dy = gy - s_xy[2 * j + 1];
      if (dx >= 0 && dx < P && dy >= 0 && dy < P) { owner = false; break; }
    }
    if (!owner) continue;
    if (gx < x_lo || gx - x_lo >= plane_w || gy < 0 || gy >= plane_h) continue;   // tile hangs over the plane: drop
    const long long off = static_cast<long long>(gx - x_lo) * plane_h + gy;
    float m = mean[off], v = var[off];
    uint8_t cnt = count[off];
    for (int j = i; j < B; ++j) {
      const int dx = gx - s_xy[2 * j], dy = gy - s_xy[2 * j + 1];
      if (dx < 0 || dx >= P || dy < 0 || dy >= P) continue;
      float mu, vr;
      stitch_stats(probs + (static_cast<long long>(j) * P + dx) * P + dy, N, pass_stride, mu, vr);
      m = __fadd_rn(m, mu);
      v = __fadd_rn(v, vr);
      cnt = static_cast<uint8_t>(cnt + 1);
    }
    mean[off] = m;
    var[off] = v;
    count[off] = cnt;
  }
}

// ---------------------------------------------------------------------------------------------------------
// count==0 -> 1; mean /= count; var /= count^2 (Segmentation.py:175-177); label = mean >= thr ? 255 : 0
// (Segmentation.py:336-337).  16 pixels per thread step (count is the narrowest type: 16 B vectors).
__global__ void finalize_kernel(float* __restrict__ mean, float* __restrict__ var, uint8_t* __restrict__ count,
                                long long n, float thr, uint8_t* __restrict__ label) {
  const long long nvec = n / 16;
  for (long long v = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; v < nvec;
       v += static_cast<long long>(gridDim.x) * blockDim.x) {
    uint4 craw = reinterpret_cast<uint4*>(count)[v];
    uint8_t* c = reinterpret_cast<uint8_t*>(&craw);
    __align__(16) uint8_t lab[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 m4 = reinterpret_cast<float4*>(mean)[v * 4 + q];
      float4 v4 = reinterpret_cast<float4*>(var)[v * 4 + q];
      float* mp = reinterpret_cast<float*>(&m4);
      float* vp = reinterpret_cast<float*>(&v4);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint8_t cc = c[q * 4 + t];
        if (cc == 0) cc = 1;
        c[q * 4 + t] = cc;
        const float cf = static_cast<float>(cc);
        mp[t] = __fdiv_rn(mp[t], cf);
        vp[t] = __fdiv_rn(vp[t], __fmul_rn(cf, cf));
        lab[q * 4 + t] = (mp[t] >= thr) ? 255 : 0;
      }
      reinterpret_cast<float4*>(mean)[v * 4 + q] = m4;
      reinterpret_cast<float4*>(var)[v * 4 + q] = v4;
    }
    reinterpret_cast<uint4*>(count)[v] = craw;
    if (label) reinterpret_cast<uint4*>(label)[v] = *reinterpret_cast<const uint4*>(lab);
  }
  // tail
  const long long t0 = nvec * 16 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (t0 < n) {
    uint8_t cc = count[t0];
    if (cc == 0) cc = 1;
    count[t0] = cc;
    const float cf = static_cast<float>(cc);
    const float m = __fdiv_rn(mean[t0], cf);
    mean[t0] = m;
    var[t0] = __fdiv_rn(var[t0], __fmul_rn(cf, cf));
    if (label) label[t0] = (m >= thr) ? 255 : 0;
  }
}

// 2x mean-pool of an [W][H] fp32 plane -> [W/2][H/2] (in-HBM probability pyramid, BASELINE config 3).
__global__ void pyramid_down2_kernel(const float* __restrict__ in, long long W, long long H,
                                     float* __restrict__ out) {
  const long long OW = W / 2, OH = H / 2, total = OW * OH;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long oy = idx % OH, ox = idx / OH;
    const float* p = in + (2 * ox) * H + 2 * oy;
    out[idx] = 0.25f * ((p[0] + p[1]) + (p[H] + p[H + 1]));
  }
}

}  // namespace dp
