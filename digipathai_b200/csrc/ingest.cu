// libdigipath_ingest.so: nvJPEG batch decode of whole-slide-image tiles + scatter into the [x][y][c] raster
// (C ABI in include/digipath_ingest.h).  tests/test_gpu_wsi_ingest.py is its parity test against
// Pillow's (libjpeg) decode of the same streams.  First hardware run (round 2): streams whose three components ARE
// R, G, B (TIFF photometric = RGB, what libtiff writes for RGB input) came back colour-transformed -- nvJPEG applies the
// YCbCr -> RGB matrix to any 3-component stream and ignores the Adobe APP14 "no transform" flag libjpeg honours -- so
// such pages are decoded with NVJPEG_OUTPUT_UNCHANGED (component planes as stored) and interleaved by a kernel here.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>
#include <nvjpeg.h>

#include "../../include/digipath_ingest.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

#define CU_OK(x)                                                                               \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) return fail("%s: %s", #x, cudaGetErrorString(e_));                  \
  } while (0)
#define NJ_OK(x)                                                                               \
  do {                                                                                         \
    nvjpegStatus_t s_ = (x);                                                                   \
    if (s_ != NVJPEG_STATUS_SUCCESS) return fail("%s: nvjpeg status %d", #x, (int)s_);         \
  } while (0)

// One 32 x 32 pixel block of one tile per CTA: rows are read along x (coalesced in the tile's [y][x][c] layout),
// transposed through shared memory, and written along y (coalesced in the raster's [x][y][c] layout).
__global__ void __launch_bounds__(256) scatter_tiles_xy_kernel(const uint8_t* __restrict__ tiles, int tile_w, int tile_h,
                                                               const int32_t* __restrict__ origins,
                                                               uint8_t* __restrict__ raster, long long x_lo,
                                                               long long x_hi, long long height) {
  __shared__ uint8_t sm[32][32 * 3 + 4];      // [y][x * 3 + c]; the pad staggers the column reads across banks
  const int t = blockIdx.z;
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const uint8_t* src = tiles + static_cast<size_t>(t) * tile_h * tile_w * 3;
  const long long ox = origins[2 * t], oy = origins[2 * t + 1];
  // load: 32 rows x 96 bytes
  for (int i = threadIdx.x; i < 32 * 96; i += 256) {
    const int r = i / 96, b = i - r * 96;
    const int y = by + r, xb = bx * 3 + b;
    sm[r][b] = (y < tile_h && xb < tile_w * 3) ? src[(static_cast<size_t>(y) * tile_w) * 3 + xb] : 0;
  }
  __syncthreads();
  // store: for each of the 32 columns (x), 32 pixels along y = 96 contiguous bytes of the raster
  for (int i = threadIdx.x; i < 32 * 96; i += 256) {
    const int cx = i / 96, b = i - cx * 96;
    const int r = b / 3, c = b - r * 3;
    const int x = bx + cx, y = by + r;
    const long long gx = ox + x, gy = oy + y;
    if (x < tile_w && y < tile_h && gx >= x_lo && gx < x_hi && gy >= 0 && gy < height)
      raster[((gx - x_lo) * height + gy) * 3 + c] = sm[r][cx * 3 + c];
  }
}

// planes uint8 [n][3][h][w] -> interleaved uint8 [n][h][w][3]; one thread per pixel.
__global__ void interleave_planes_kernel(const uint8_t* __restrict__ planes, uint8_t* __restrict__ rgb, long long n_pix_tile,
                                         long long total) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / n_pix_tile, p = i - t * n_pix_tile;
    const uint8_t* src = planes + t * 3 * n_pix_tile + p;
    uint8_t* dst = rgb + i * 3;
    dst[0] = src[0];
    dst[1] = src[n_pix_tile];
    dst[2] = src[2 * n_pix_tile];
  }
}

}  // namespace

struct dp_jpeg_decoder {
  int device = 0;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  int batch_ready = 0;   // batch size nvjpegDecodeBatchedInitialize was last called with
  int fmt_ready = -1;    // ... and output format
  uint8_t* planes = nullptr;   // scratch for NVJPEG_OUTPUT_UNCHANGED decodes
  size_t planes_bytes = 0;
};

extern "C" {

int dp_ingest_abi_version(void) { return 1; }
const char* dp_ingest_last_error(void) { return g_err; }

int dp_jpeg_decoder_create(int device, dp_jpeg_decoder** out) {
  if (!out) return fail("null argument");
  *out = nullptr;
  CU_OK(cudaSetDevice(device));
  dp_jpeg_decoder* d = new dp_jpeg_decoder();
  d->device = device;
  nvjpegStatus_t s = nvjpegCreateSimple(&d->handle);
  if (s == NVJPEG_STATUS_SUCCESS) s = nvjpegJpegStateCreate(d->handle, &d->state);
  if (s != NVJPEG_STATUS_SUCCESS) {
    if (d->handle) nvjpegDestroy(d->handle);
    delete d;
    return fail("nvjpeg initialisation failed: status %d", (int)s);
  }
  *out = d;
  return 0;
}

int dp_jpeg_decoder_destroy(dp_jpeg_decoder* d) {
  if (!d) return 0;
  if (d->state) nvjpegJpegStateDestroy(d->state);
  if (d->handle) nvjpegDestroy(d->handle);
  if (d->planes) cudaFree(d->planes);
  delete d;
  return 0;
}

int dp_jpeg_decode_tiles_ex(dp_jpeg_decoder* d, const uint8_t* const* streams, const size_t* lengths, int n, int tile_w,
                            int tile_h, uint8_t* out_rgb, int components_are_rgb, void* stream) {
  if (!d || !streams || !lengths || !out_rgb) return fail("null argument");
  if (n < 1 || tile_w < 1 || tile_h < 1) return fail("bad tile geometry");
  CU_OK(cudaSetDevice(d->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<nvjpegImage_t> dst(n);
  const size_t npix = static_cast<size_t>(tile_w) * tile_h, slot = npix * 3;
  if (components_are_rgb && d->planes_bytes < slot * n) {
    if (d->planes) { CU_OK(cudaStreamSynchronize(st)); CU_OK(cudaFree(d->planes)); d->planes = nullptr; }
    CU_OK(cudaMalloc(&d->planes, slot * n));
    d->planes_bytes = slot * n;
  }
  if (components_are_rgb) CU_OK(cudaMemsetAsync(d->planes, 0, slot * n, st));
  for (int i = 0; i < n; ++i) {
    int comps = 0, w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    nvjpegChromaSubsampling_t ss;
    NJ_OK(nvjpegGetImageInfo(d->handle, streams[i], lengths[i], &comps, &ss, w, h));
    if (comps != 1 && comps != 3) return fail("stream %d has %d components (1 or 3 supported)", i, comps);
    if (w[0] > tile_w || h[0] > tile_h) return fail("stream %d is %d x %d, larger than the %d x %d tile", i, w[0], h[0], tile_w, tile_h);
    memset(&dst[i], 0, sizeof(nvjpegImage_t));
    if (components_are_rgb) {
      if (comps != 3 || ss != NVJPEG_CSS_444)
        return fail("stream %d: RGB-component pages must hold three full-resolution components", i);
      for (int c = 0; c < 3; ++c) {
        dst[i].channel[c] = d->planes + slot * i + npix * c;
        dst[i].pitch[c] = static_cast<size_t>(tile_w);
      }
    } else {
      dst[i].channel[0] = out_rgb + slot * i;
      dst[i].pitch[0] = static_cast<size_t>(tile_w) * 3;
    }
  }
  const nvjpegOutputFormat_t fmt = components_are_rgb ? NVJPEG_OUTPUT_UNCHANGED : NVJPEG_OUTPUT_RGBI;
  if (d->batch_ready != n || d->fmt_ready != (int)fmt) {
    NJ_OK(nvjpegDecodeBatchedInitialize(d->handle, d->state, n, 1, fmt));
    d->batch_ready = n;
    d->fmt_ready = (int)fmt;
  }
  NJ_OK(nvjpegDecodeBatched(d->handle, d->state, streams, lengths, dst.data(), st));
  if (components_are_rgb) {
    const long long total = static_cast<long long>(npix) * n;
    const int grid = static_cast<int>(total / 256 + 1 > 148 * 32 ? 148 * 32 : total / 256 + 1);
    interleave_planes_kernel<<<grid, 256, 0, st>>>(d->planes, out_rgb, static_cast<long long>(npix), total);
    CU_OK(cudaGetLastError());
  }
  return 0;
}

int dp_jpeg_decode_tiles(dp_jpeg_decoder* d, const uint8_t* const* streams, const size_t* lengths, int n, int tile_w,
                         int tile_h, uint8_t* out_rgb, void* stream) {
  return dp_jpeg_decode_tiles_ex(d, streams, lengths, n, tile_w, tile_h, out_rgb, 0, stream);
}

int dp_scatter_tiles_xy(const uint8_t* tiles, int n, int tile_w, int tile_h, const int32_t* origins, uint8_t* raster,
                        int64_t x_lo, int64_t x_hi, int64_t height, void* stream) {
  if (!tiles || !origins || !raster) return fail("null argument");
  if (n < 1 || tile_w < 1 || tile_h < 1 || x_hi <= x_lo || height < 1) return fail("bad scatter geometry");
  if (n > 65535) return fail("at most 65535 tiles per scatter call");
  const dim3 grid((tile_w + 31) / 32, (tile_h + 31) / 32, n);
  scatter_tiles_xy_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(tiles, tile_w, tile_h, origins, raster, x_lo,
                                                                             x_hi, height);
  CU_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
