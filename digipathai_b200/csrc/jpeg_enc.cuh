// Baseline JPEG encoder for the 256 x 256 grayscale tiles of the result pyramids (row N1 of SURVEY.md 8(f)): the
// reference writes each plane with tifffile and then re-encodes it with ImageMagick --
// `convert <p> -compress jpeg -quality 90 -define tiff:tile-geometry=256x256 ptif:<p>` (DigiPathAI/Segmentation.py:333-334,
// 345-346, 351-352) -- which for a 40 000^2 plane is ~33 000 tiles per file and, on host cores, longer than the whole
// B200 segmentation.  Here a tile is one CTA:
//   stage      the tile (border replicated at the plane's edges) into shared memory, min / max for the "constant tile"
//              short cut of the writer;
//   DCT        one thread per 8 x 8 block (4 blocks per thread): level shift, separable float DCT-II, quantisation with
//              the table parsed from the JPEG header the host side uses, zig-zag order, int16 coefficients in shared memory;
//   size       bits of every block's Huffman code (DC difference category + run/size symbols, ZRL, EOB), exclusive scan
//              over the 1024 blocks of the tile -> bit offset of every block;
//   emit       every block writes its codes at its bit offset into a zeroed scratch bit stream (atomicOr on big-endian
//              32-bit words: only the first and last word of a block are shared with its neighbours);
//   stuff      0xFF -> 0xFF 0x00 with a second scan over byte ranges; the final byte is padded with 1 bits.
// The Huffman and quantisation tables are not restated from the standard: the host parses them out of a JPEG produced
// by the same libjpeg the host path uses (tiffio._jpeg_tables), so the scan data emitted here and that header form one
// valid stream by construction.  A tile whose stream would exceed the per-tile capacity is flagged and encoded on the host.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

constexpr int kJpTile = 256;
constexpr int kJpBlocks = 1024;                    // 8 x 8 blocks per tile
constexpr int kJpThreads = 256;

struct JpegTables {
  float inv_q[64];                                 // 1 / quantiser, NATURAL order
  uint16_t dc_code[16]; uint8_t dc_len[16];        // by difference category
  uint16_t ac_code[256]; uint8_t ac_len[256];      // by (run << 4) | size
};

__device__ __forceinline__ int jp_category(int v) { return v == 0 ? 0 : 32 - __clz(v < 0 ? -v : v); }

// appends `len` bits (`code`, right aligned) at bit position `pos` of a big-endian word stream
__device__ __forceinline__ void jp_put(uint32_t* words, long long& pos, uint32_t code, int len) {
  const int sh = static_cast<int>(pos & 31);
  const unsigned long long v = static_cast<unsigned long long>(code) << (64 - len - sh);
  const uint32_t hi = static_cast<uint32_t>(v >> 32), lo = static_cast<uint32_t>(v);
  uint32_t* w = words + (pos >> 5);
  if (hi) atomicOr(w, hi);
  if (lo) atomicOr(w + 1, lo);
  pos += len;
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int jp_scan256(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) s_warp[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = lane < 8 ? s_warp[lane] : 0;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    if (lane < 8) s_warp[lane] = w;                // inclusive warp sums
  }
  __syncthreads();
  const int base = warp ? s_warp[warp - 1] : 0;
  *total = s_warp[7];
  __syncthreads();
  return base + x - v;
}

// plane: uint8 [rows][cols]; one CTA per tile of the tiles_x x tiles_y grid starting at tile index `tile0` (row-major).
// scratch: zeroed words, `scratch_words` per tile; out: `out_cap` bytes per tile; sizes / flags per tile:
// flags bit 0 = constant tile (value in bits 8..15), bit 1 = capacity exceeded (encode on the host).
__global__ void __launch_bounds__(kJpThreads, 1)
jpeg_encode_tiles_kernel(const uint8_t* __restrict__ plane, int rows, int cols, int tiles_x, int tile0, int n_tiles,
                         const JpegTables* __restrict__ tabs, uint32_t* __restrict__ scratch, int scratch_words,
                         uint8_t* __restrict__ out, int out_cap, int* __restrict__ sizes, int* __restrict__ flags) {
  extern __shared__ uint8_t jp_smem[];
  int16_t* coef = reinterpret_cast<int16_t*>(jp_smem);                        // [1024][64], zig-zag order
  uint8_t* pix = jp_smem + kJpBlocks * 64 * 2;                               // [256][256]
  __shared__ int s_warp[8];
  __shared__ int s_minmax[2];
  __shared__ JpegTables s_tab;
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const int tile = tile0 + t, ty = tile / tiles_x, tx = tile - ty * tiles_x;
  const int tid = threadIdx.x;
  for (int i = tid; i < static_cast<int>(sizeof(JpegTables) / 4); i += kJpThreads)
    reinterpret_cast<uint32_t*>(&s_tab)[i] = reinterpret_cast<const uint32_t*>(tabs)[i];
  if (tid == 0) { s_minmax[0] = 255; s_minmax[1] = 0; }
  __syncthreads();
  // ---- stage (rows of 256 bytes, coalesced; clamp = replicate the plane's border into the padding)
  int mn = 255, mx = 0;
  for (int i = tid; i < kJpTile * kJpTile / 4; i += kJpThreads) {
    const int r = i >> 6, c4 = (i & 63) * 4;
    const int gr = min(ty * kJpTile + r, rows - 1);
    uint32_t packed = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int gc = min(tx * kJpTile + c4 + k, cols - 1);
      const int v = plane[static_cast<long long>(gr) * cols + gc];
      packed |= static_cast<uint32_t>(v) << (8 * k);
      mn = min(mn, v); mx = max(mx, v);
    }
    reinterpret_cast<uint32_t*>(pix)[i] = packed;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((tid & 31) == 0) { atomicMin(&s_minmax[0], mn); atomicMax(&s_minmax[1], mx); }
  __syncthreads();
  if (s_minmax[0] == s_minmax[1]) {                // constant tile: the writer shares one stream per value
    if (tid == 0) { sizes[t] = 0; flags[t] = 1 | (s_minmax[0] << 8); }
    return;
  }
  // ---- DCT + quantisation, 4 blocks per thread
  for (int b = tid; b < kJpBlocks; b += kJpThreads) {
    const int by = b >> 5, bx = b & 31;
    float f[64];
#pragma unroll
    for (int y = 0; y < 8; ++y)
#pragma unroll
      for (int x = 0; x < 8; ++x) f[y * 8 + x] = static_cast<float>(pix[(by * 8 + y) * kJpTile + bx * 8 + x]) - 128.f;
    // 8-point DCT-II on rows then columns: F(u) = c(u)/2 * sum_x f(x) cos((2x+1) u pi / 16)
    const float c1 = 0.98078528f, c2 = 0.92387953f, c3 = 0.83146961f, c4 = 0.70710678f, c5 = 0.55557023f,
                c6 = 0.38268343f, c7 = 0.19509032f;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int s = pass ? 8 : 1, o = pass ? i : i * 8;
        const float x0 = f[o], x1 = f[o + s], x2 = f[o + 2 * s], x3 = f[o + 3 * s], x4 = f[o + 4 * s], x5 = f[o + 5 * s],
                    x6 = f[o + 6 * s], x7 = f[o + 7 * s];
        const float s07 = x0 + x7, d07 = x0 - x7, s16 = x1 + x6, d16 = x1 - x6, s25 = x2 + x5, d25 = x2 - x5,
                    s34 = x3 + x4, d34 = x3 - x4;
        const float e0 = s07 + s34, e1 = s16 + s25, e2 = s07 - s34, e3 = s16 - s25;
        f[o] = 0.5f * c4 * (e0 + e1);
        f[o + 4 * s] = 0.5f * c4 * (e0 - e1);
        f[o + 2 * s] = 0.5f * (c2 * e2 + c6 * e3);
        f[o + 6 * s] = 0.5f * (c6 * e2 - c2 * e3);
        f[o + s] = 0.5f * (c1 * d07 + c3 * d16 + c5 * d25 + c7 * d34);
        f[o + 3 * s] = 0.5f * (c3 * d07 - c7 * d16 - c1 * d25 - c5 * d34);
        f[o + 5 * s] = 0.5f * (c5 * d07 - c1 * d16 + c7 * d25 + c3 * d34);
        f[o + 7 * s] = 0.5f * (c7 * d07 - c5 * d16 + c3 * d25 - c1 * d34);
      }
    }
    // natural index n -> zig-zag position (compile-time after unrolling: f stays in registers)
    constexpr int kInvZigzag[64] = {0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
#pragma unroll
    for (int n = 0; n < 64; ++n)
      coef[b * 64 + kInvZigzag[n]] = static_cast<int16_t>(__float2int_rn(f[n] * s_tab.inv_q[n]));
  }
  __syncthreads();
  // ---- bits per block (blocks 4 tid .. 4 tid + 3: consecutive in scan order, so one offset per thread suffices)
  auto block_bits = [&](int b, bool emit, uint32_t* words, long long& pos) -> int {
    const int16_t* c = coef + b * 64;
    const int dc = c[0] - (b ? coef[(b - 1) * 64] : 0);
    int bits = 0;
    {
      const int cat = jp_category(dc);
      bits += s_tab.dc_len[cat] + cat;
      if (emit) {
        jp_put(words, pos, s_tab.dc_code[cat], s_tab.dc_len[cat]);
        if (cat) jp_put(words, pos, static_cast<uint32_t>(dc < 0 ? dc - 1 : dc) & ((1u << cat) - 1), cat);
      }
    }
    int run = 0;
    for (int k = 1; k < 64; ++k) {
      const int v = c[k];
      if (v == 0) { ++run; continue; }
      while (run >= 16) {
        bits += s_tab.ac_len[0xF0];
        if (emit) jp_put(words, pos, s_tab.ac_code[0xF0], s_tab.ac_len[0xF0]);
        run -= 16;
      }
      const int cat = jp_category(v), sym = (run << 4) | cat;
      bits += s_tab.ac_len[sym] + cat;
      if (emit) {
        jp_put(words, pos, s_tab.ac_code[sym], s_tab.ac_len[sym]);
        jp_put(words, pos, static_cast<uint32_t>(v < 0 ? v - 1 : v) & ((1u << cat) - 1), cat);
      }
      run = 0;
    }
    if (run) {
      bits += s_tab.ac_len[0x00];
      if (emit) jp_put(words, pos, s_tab.ac_code[0x00], s_tab.ac_len[0x00]);
    }
    return bits;
  };
  long long dummy = 0;
  int my_bits = 0;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) my_bits += block_bits(4 * tid + i, false, nullptr, dummy);
  int total_bits = 0;
  const int my_off = jp_scan256(my_bits, s_warp, &total_bits);
  const int n_bytes = (total_bits + 7) >> 3;
  uint32_t* words = scratch + static_cast<long long>(t) * scratch_words;
  if (n_bytes + 8 > scratch_words * 4) {           // would not fit the scratch stream
    if (tid == 0) { sizes[t] = 0; flags[t] = 2; }
    return;
  }
  long long pos = my_off;
#pragma unroll 1
  for (int i = 0; i < 4; ++i) block_bits(4 * tid + i, true, words, pos);
  if (tid == kJpThreads - 1 && (total_bits & 7)) jp_put(words, pos, (1u << (8 - (total_bits & 7))) - 1, 8 - (total_bits & 7));
  __threadfence();
  __syncthreads();
  // ---- byte stuffing: thread i owns bytes [i * per, (i + 1) * per)
  const int per = (n_bytes + kJpThreads - 1) / kJpThreads;
  const int b0 = min(tid * per, n_bytes), b1 = min(b0 + per, n_bytes);
  auto byte_at = [&](int k) { return (__ldcg(words + (k >> 2)) >> (24 - 8 * (k & 3))) & 0xFFu; };
  int ff = 0;
  for (int k = b0; k < b1; ++k) ff += byte_at(k) == 0xFFu;
  int total_ff = 0;
  const int ff_off = jp_scan256(ff, s_warp, &total_ff);
  const int out_size = n_bytes + total_ff;
  if (out_size > out_cap) {
    if (tid == 0) { sizes[t] = 0; flags[t] = 2; }
    return;
  }
  uint8_t* o = out + static_cast<long long>(t) * out_cap + b0 + ff_off;
  for (int k = b0; k < b1; ++k) {
    const uint32_t v = byte_at(k);
    *o++ = static_cast<uint8_t>(v);
    if (v == 0xFFu) *o++ = 0;
  }
  if (tid == 0) { sizes[t] = out_size; flags[t] = 0; }
}

// fixed-stride per-tile streams -> one contiguous buffer (offsets = exclusive scan of sizes)
__global__ void jpeg_compact_kernel(const uint8_t* __restrict__ in, int cap, const int* __restrict__ sizes,
                                    const long long* __restrict__ offsets, uint8_t* __restrict__ out, int n_tiles) {
  const int t = blockIdx.x;
  if (t >= n_tiles) return;
  const uint8_t* src = in + static_cast<long long>(t) * cap;
  uint8_t* dst = out + offsets[t];
  for (int i = threadIdx.x; i < sizes[t]; i += blockDim.x) dst[i] = src[i];
}

}  // namespace dp
