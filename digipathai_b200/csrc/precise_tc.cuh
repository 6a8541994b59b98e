// 3xTF32 precision mode (`precision = tf32x3` in the model container): the fp32 layer program of precise.cuh with its
// convs on the tensor cores.
//
// An fp32 operand is split into two TF32 numbers, x = hi + lo (hi = x rounded to a 10-bit mantissa, lo = the rounded
// remainder: 21-22 mantissa bits together), and a product is evaluated as three tcgen05.mma.kind::tf32 instructions
// accumulating in TMEM:      a*w  ~=  a_hi*w_hi + a_lo*w_hi + a_hi*w_lo        (the dropped a_lo*w_lo is 2^-22).
//
// The tensor core's accumulate step TRUNCATES toward zero at fp32 precision (measured, tools/tf32x3_accuracy.py: with
// positive operands the result of a K-long reduction is low by K x 1.0e-8 relative, i.e. about one fp32 ulp per MMA
// instruction; the fp16 path has the same property, invisible below its fp16 output rounding).  Left alone that bias
// is 30x the error of the fp32 FMA mode on the DenseNet.  So a TMEM accumulator only ever holds a CHUNK of two K
// slices (24 MMA instructions); finished chunks are added into fp32 registers with round-to-nearest adds while the
// tensor pipe fills the other of two accumulators.
//
// One kernel covers every conv the fp32 mode runs (same NaiveConvParams description: 1x1, 3x3, up2 sub-pixel phases,
// generic tap tables, stride 2, pre-activation BN(+ReLU) prologue, shift / ReLU / residual epilogue).  Implicit GEMM,
// one CTA per 128 output pixels x N <= 128 couts (x accumulator group); K is walked in slices of 32 channels per tap:
//   cp.async          every thread copies its own 16-byte pieces of the slice (4 of the activation tile, <= 4 of the
//                     weight tile) into a private slot of a 3-deep raw ring, two slices ahead of their use -- global
//                     latency is hidden without holding the data in registers, zero padding is the copy's zero fill;
//   convert           the same thread reads its pieces back, applies the prologue, splits, and writes hi and lo into
//                     the K-major SWIZZLE_128B operand tiles (row = 32 fp32 = 128 B; chunk j of row r at j ^ (r & 7));
//   one elected thread 4 K-steps (8 fp32 = 32 B each) x 3 MMAs, M = 128; tcgen05.commit on the stage's mbarrier;
//   two operand stages the conversion of slice i+1 overlaps the MMAs of slice i;
//   epilogue          8 warps: TMEM lane quadrant = warp & 3, 16-column units alternate between the two warpgroups;
//                     register sums -> scale/shift (+ residual) (+ ReLU) -> fp32 row segments.
#pragma once
#include "conv_tc.cuh"

namespace dp {

constexpr int kTxThreads = 256;
constexpr int kTxSliceK = 32;                       // channels per K slice = one 128-byte swizzled row of fp32
constexpr int kTxATile = 128 * 128;                 // 128 pixel rows x 128 B
constexpr int kTxMaxN = 128;                        // 2 accumulators x N columns in TMEM, N / 2 register sums per thread
constexpr int kTxRaw = 3;                           // raw ring depth: copies run 2 slices ahead

__host__ __device__ inline int tx_b_iters(int n) { return (n + 31) / 32; }
__host__ __device__ inline int tx_op_stage_bytes(int n) { return 2 * kTxATile + 2 * n * 128; }
__host__ __device__ inline int tx_raw_stage_bytes(int n) { return (4 + tx_b_iters(n)) * kTxThreads * 16; }
__host__ __device__ inline int tx_smem_bytes(int n) {
  return 1024 /*align*/ + 2 * tx_op_stage_bytes(n) + kTxRaw * tx_raw_stage_bytes(n) + 64;
}

// kind::tf32 instruction descriptor: TF32 A/B (K-major), fp32 D, M = 128, N = n.
__host__ __device__ __forceinline__ uint32_t make_idesc_tf32(uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// 16-byte asynchronous copy global -> shared; `bytes` = 16, or 0 for a zero fill (nothing is read then).
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// x -> (hi, lo): both representable in TF32 (low 13 mantissa bits zero), hi + lo = x up to 2^-22 |x|.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  uint32_t h, l;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
  hi = __uint_as_float(h);
  const float r = x - hi;                            // exact
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
  lo = __uint_as_float(l);
}

__global__ void __launch_bounds__(kTxThreads, 1) conv_tf32x3_kernel(const NaiveConvParams p, const int n_tile_dbg) {
  const int n_tile = n_tile_dbg & 0xFFFF, dbg = n_tile_dbg >> 16;   // TEMP timing switches
  extern __shared__ uint8_t tx_smem_raw[];
  const uint32_t raw_addr = smem_u32(tx_smem_raw);
  uint8_t* smem = tx_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int op_bytes = tx_op_stage_bytes(n_tile), raw_bytes = tx_raw_stage_bytes(n_tile);
  uint8_t* raw_ring = smem + 2 * op_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(raw_ring + kTxRaw * raw_bytes);   // [0],[1]: operand stage read; [2],[3]: accumulator full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const float* __restrict__ in = reinterpret_cast<const float*>(p.in);
  const float* __restrict__ wgt = reinterpret_cast<const float*>(p.w);
  float* __restrict__ out = reinterpret_cast<float*>(p.out);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long M = static_cast<long long>(p.n_img) * p.OH * p.OW;
  const long long m0 = static_cast<long long>(blockIdx.x) * 128;
  const int n0 = blockIdx.y * n_tile;
  const int g = blockIdx.z;
  const int n_valid = min(n_tile, p.Cout - n0);      // weight rows beyond Cout are staged as zeros

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * n_tile)) tmem_cols <<= 1;
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // ---- this thread's pieces of every slice: 16-byte chunk j (4 channels) of rows (tid >> 3) + 32 i
  const int j = tid & 7, r0 = tid >> 3;
  const uint32_t sw_row = static_cast<uint32_t>((r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));   // same for every i (32 i % 8 == 0)
  long long a_base[4];                               // element offset of (pixel, channel 4 j) for tap (0, 0), slice 0
  int a_hw[4];                                       // (oh * stride) << 16 | (ow * stride); rows past M never pass the bounds test
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long lm = m0 + r0 + 32 * i;
    if (lm < M) {
      const int ow = static_cast<int>(lm % p.OW);
      const long long r = lm / p.OW;
      const int oh = static_cast<int>(r % p.OH);
      const int n = static_cast<int>(r / p.OH);
      a_base[i] = ((static_cast<long long>(n) * p.H + oh * p.stride) * p.W + ow * p.stride) * p.in_ctot + p.in_choff + 4 * j;
      a_hw[i] = ((oh * p.stride) << 16) | (ow * p.stride);
    } else {
      a_base[i] = 0;
      a_hw[i] = 0x40004000;                          // 16384: outside any map
    }
  }
  const int b_iters = tx_b_iters(n_tile);            // <= 4

  const int slices_per_tap = (p.Cin + kTxSliceK - 1) / kTxSliceK;
  int n_taps = 0;
  for (int e = 0; e < p.n_entries_total; ++e) n_taps += (p.entries[e].group == g);
  const int n_slices = n_taps * slices_per_tap;
  // walk (tap, channel slice) incrementally for the copy stream (cs_*) and the convert stream (cv_*)
  auto next_tap = [&](int e) {
    do { ++e; } while (e < p.n_entries_total && p.entries[e].group != g);
    return e;
  };
  int cs_e = next_tap(-1), cs_c = 0, cv_e = cs_e, cv_c = 0;

  auto in_bounds = [&](int hw, const TapEntry& ent) {
    const int ih = (hw >> 16) + ent.dy, iw = (hw & 0xFFFF) + ent.dx;
    return ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
  };
  auto issue = [&](int slot) {                       // cp.async of slice (cs_e, cs_c) into raw slot `slot`; advances
    const TapEntry ent = p.entries[cs_e];
    const uint32_t dst = smem_u32(raw_ring + slot * raw_bytes) + static_cast<uint32_t>(tid) * 16u;
    const int c = cs_c + 4 * j;
    const long long tap_off = (static_cast<long long>(ent.dy) * p.W + ent.dx) * p.in_ctot + cs_c;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool ok = c < p.Cin && in_bounds(a_hw[i], ent);
      cp_async16(dst + i * (kTxThreads * 16), ok ? static_cast<const void*>(in + a_base[i] + tap_off) : static_cast<const void*>(in),
                 ok ? 16u : 0u);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < b_iters) {
        const int r = r0 + 32 * i;
        const bool ok = r < n_valid && c < p.Cin;
        cp_async16(dst + (4 + i) * (kTxThreads * 16),
                   ok ? static_cast<const void*>(wgt + (static_cast<long long>(cs_e) * p.Cout + (n0 + r)) * p.Cin + c)
                      : static_cast<const void*>(wgt), ok ? 16u : 0u);
      }
    }
    cs_c += kTxSliceK;
    if (cs_c >= p.Cin) { cs_c = 0; cs_e = next_tap(cs_e); }
  };
  auto split_store = [&](uint8_t* hi_tile, uint8_t* lo_tile, int i, const float4& v) {
    float4 h, l;
    tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
    const uint32_t off = static_cast<uint32_t>(((r0 >> 3) + 4 * i) * 1024) + sw_row;
    *reinterpret_cast<float4*>(hi_tile + off) = h;
    *reinterpret_cast<float4*>(lo_tile + off) = l;
  };

  // ---- register sums of the finished accumulator chunks: units u = half, half + 2, ... of 16 columns
  const int quad = warp & 3, half = warp >> 2;
  float sum[4][16];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 16; ++b) sum[a][b] = 0.f;
  auto drain = [&](int chunk) {                      // acc[chunk & 1] -> sum (round-to-nearest adds)
    mbar_wait(&bars[2 + (chunk & 1)], (chunk >> 1) & 1);
    tc_fence_after();
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>((chunk & 1) * n_tile);
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const int u = half + 2 * a;
      if (u * 16 < n_tile) {
        uint32_t v[16];
        tmem_ld16(t0 + static_cast<uint32_t>(u * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int b = 0; b < 16; ++b) sum[a][b] += __uint_as_float(v[b]);
      }
    }
    tc_fence_before();
  };

  const uint32_t idesc = make_idesc_tf32(static_cast<uint32_t>(n_tile));
  const int n_chunks = (n_slices + 1) >> 1;
  int drained = 0;
  if (n_slices > 0) issue(0);
  cp_async_commit();
  if (n_slices > 1) issue(1);
  cp_async_commit();
  for (int s = 0; s < n_slices; ++s) {
    const int st = s & 1;
    uint8_t* a_hi = smem + st * op_bytes;
    uint8_t* a_lo = a_hi + kTxATile;
    uint8_t* b_hi = a_lo + kTxATile;
    uint8_t* b_lo = b_hi + n_tile * 128;
    if (s + 2 < n_slices && !(dbg & 8)) issue((s + 2) % kTxRaw);   // that slot was read back by this thread in iteration s - 1
    cp_async_commit();
    cp_async_wait<2>();                              // this thread's pieces of slice s have landed
    if (s >= 2) mbar_wait(&bars[st], ((s >> 1) - 1) & 1);      // the MMAs of slice s-2 have read this operand stage
    if (!(dbg & 4)) {
      const TapEntry ent = p.entries[cv_e];
      const uint8_t* rs = raw_ring + (s % kTxRaw) * raw_bytes + tid * 16;
      const int c = cv_c + 4 * j;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.pro_mode && c < p.Cin) {
        sc = *reinterpret_cast<const float4*>(p.pro_scale + c);
        sh = *reinterpret_cast<const float4*>(p.pro_shift + c);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = *reinterpret_cast<const float4*>(rs + i * (kTxThreads * 16));
        if (p.pro_mode) {
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
          if (p.pro_mode == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
          // the conv's zero padding (and the channel tail) acts on the ACTIVATED tensor
          if (!(c < p.Cin && in_bounds(a_hw[i], ent))) v = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        split_store(a_hi, a_lo, i, v);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i < b_iters && r0 + 32 * i < n_tile)
          split_store(b_hi, b_lo, i, *reinterpret_cast<const float4*>(rs + (4 + i) * (kTxThreads * 16)));
      cv_c += kTxSliceK;
      if (cv_c >= p.Cin) { cv_c = 0; cv_e = next_tap(cv_e); }
    }
    // chunk (s >> 1) - 1 finished issuing one iteration ago: fold it into the register sums before its accumulator
    // is reused by chunk (s >> 1) + 1
    if ((s & 1) && s >= 3 && !(dbg & 16)) drain(drained++);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(((s >> 1) & 1) * n_tile);
      const uint64_t dah = make_sw128_desc(smem_u32(a_hi), 1024, 0), dal = make_sw128_desc(smem_u32(a_lo), 1024, 0);
      const uint64_t dbh = make_sw128_desc(smem_u32(b_hi), 1024, 0), dbl = make_sw128_desc(smem_u32(b_lo), 1024, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                  // 8 fp32 = 32 B = 2 descriptor units per K-step
        if (!(dbg & 2)) umma_tf32_ss(d, dah + 2 * k, dbh + 2 * k, idesc, ((s & 1) | k) ? 1u : 0u);
        if (!(dbg & 3)) umma_tf32_ss(d, dal + 2 * k, dbh + 2 * k, idesc, 1u);
        if (!(dbg & 3)) umma_tf32_ss(d, dah + 2 * k, dbl + 2 * k, idesc, 1u);
      }
      umma_commit(&bars[st]);
      if ((s & 1) || s == n_slices - 1) umma_commit(&bars[2 + ((s >> 1) & 1)]);
    }
    __syncwarp();
  }
  while (drained < n_chunks && !(dbg & 16)) drain(drained++);

  // ---- epilogue
  const long long mm = m0 + quad * 32 + lane;
  const bool mvalid = mm < M;
  long long opix = 0;
  if (mvalid) {
    const int ow = static_cast<int>(mm % p.OW);
    const long long r = mm / p.OW;
    const int oh = static_cast<int>(r % p.OH);
    const int n = static_cast<int>(r / p.OH);
    opix = p.up2 ? (static_cast<long long>(n) * 2 * p.H + 2 * oh + (g >> 1)) * (2 * p.W) + 2 * ow + (g & 1)
                 : (static_cast<long long>(n) * p.OH + oh) * p.OW + ow;
  }
  float* orow = out + opix * p.out_ctot + p.out_choff;
  const bool vec_ok = ((p.out_ctot | p.out_choff) & 3) == 0;
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int u = half + 2 * a;
    if (u * 16 >= n_tile || !mvalid) continue;
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      const int co = n0 + u * 16 + 4 * j4;
      if (co >= p.Cout) continue;
      float y[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int c = co + t;
        float val = sum[a][4 * j4 + t];
        if (c < p.Cout) {
          val = fmaf(val, p.epi_scale ? p.epi_scale[c] : 1.f, p.epi_shift ? p.epi_shift[c] : 0.f);
          if (p.residual) val += orow[c];
          if (p.relu) val = fmaxf(val, 0.f);
        }
        y[t] = val;
      }
      if (co + 3 < p.Cout && vec_ok) {
        *reinterpret_cast<float4*>(orow + co) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
#pragma unroll
        for (int t = 0; t < 4; ++t)
          if (co + t < p.Cout) orow[co + t] = y[t];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

}  // namespace dp
