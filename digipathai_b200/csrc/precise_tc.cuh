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

// x -> (hi, lo): hi = x rounded to TF32 (nearest, ties away: add half an ulp to the magnitude bits and clear the 13 low
// bits -- two integer instructions; cvt.rna.tf32.f32 costs five with its inf / nan handling, and activations are finite),
// lo = x - hi (exact) rounded the same way, so that hi + lo = x up to 2^-22 |x| without the bias the tensor core's own
// truncation of an unrounded lo would add.
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
  lo = __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(kTxThreads, 1) conv_tf32x3_kernel(const NaiveConvParams p, const int n_tile) {
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
    if (s + 2 < n_slices) issue((s + 2) % kTxRaw);   // that slot was read back by this thread in iteration s - 1
    cp_async_commit();
    cp_async_wait<2>();                              // this thread's pieces of slice s have landed
    if (s >= 2) mbar_wait(&bars[st], ((s >> 1) - 1) & 1);      // the MMAs of slice s-2 have read this operand stage
    {
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
    if ((s & 1) && s >= 3) drain(drained++);
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t d = tmem_base + static_cast<uint32_t>(((s >> 1) & 1) * n_tile);
      const uint64_t dah = make_sw128_desc(smem_u32(a_hi), 1024, 0), dal = make_sw128_desc(smem_u32(a_lo), 1024, 0);
      const uint64_t dbh = make_sw128_desc(smem_u32(b_hi), 1024, 0), dbl = make_sw128_desc(smem_u32(b_lo), 1024, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k) {                  // 8 fp32 = 32 B = 2 descriptor units per K-step
        umma_tf32_ss(d, dah + 2 * k, dbh + 2 * k, idesc, ((s & 1) | k) ? 1u : 0u);
        umma_tf32_ss(d, dal + 2 * k, dbh + 2 * k, idesc, 1u);
        umma_tf32_ss(d, dah + 2 * k, dbl + 2 * k, idesc, 1u);
      }
      umma_commit(&bars[st]);
      if ((s & 1) || s == n_slices - 1) umma_commit(&bars[2 + ((s >> 1) & 1)]);
    }
    __syncwarp();
  }
  while (drained < n_chunks) drain(drained++);

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


// ---------------------------------------------------------------------------------------------------------------
// Halo variant for 3x3 / up2 convs without prologue (every decoder conv and the 3x3 of every dense layer: > 90 % of the
// fp32 program's MACs).  Work item = 16 rows x 8 columns of output pixels of one image.  Per 32-channel slice the
// (16+2) x (8+2) input halo arrives as ONE 4-D TMA box (zero fill outside the image = the conv's padding) straight into
// the swizzled operand tile of 180 rows, is split in place (hi stays, lo goes to a second tile) by the worker warps, and
// every tap (dy, dx) is a UMMA descriptor into that tile at row (dy+1) * 10 + (dx+1) with an 8-row-group stride of 10
// rows (the 128-byte swizzle is a function of absolute shared-memory address bits, so unaligned starts are fine:
// tools/mma_probe.cu).  Weights come PRE-SPLIT (hi / lo copies of the container's data section made at model creation)
// through two 3-D TMA boxes per tap into a ring of tap stages.  Warp 9 (one elected thread) issues every TMA, warps 8 and
// 10 the MMAs of alternate accumulator chunks -- one thread doing both spent 1.5 k clk per tap on its serial chain of barrier waits (~200 clk each) and
// commits (~200 clk each) against 0.6-0.8 k clk of MMA time; eight worker warps convert, drain finished accumulator
// chunks (3 taps; 2 for up2) into registers and run the epilogue.  All hand-offs are mbarriers; there is no
// block-wide barrier inside the loop.
constexpr int kThHaloW = 10, kThHaloH = 18, kThRows = 180;
constexpr int kThATile = 23 * 1024;                 // 180 rows x 128 B, rounded up to 1 KB
constexpr int kThItems = 6;                         // ceil(180 * 8 chunks / 256 worker threads)
constexpr int kThBStages = 4;
constexpr int kThThreads = 320;                     // 8 worker warps + MMA warp + TMA producer warp (1x1 kernel)
constexpr int kThHaloThreads = 352;                 // halo kernel: a second MMA warp (chunks alternate between the two)

__host__ __device__ inline int th_smem_bytes(int n, int tps) {
  return 1024 + 2 * 2 * kThATile + (tps > 1 ? 2 * tps : kThBStages) * 2 * n * 128 + 256 + 1024;   // + epilogue constants
}

__global__ void __launch_bounds__(kThHaloThreads, 1)
conv_halo_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_wh,
                        const __grid_constant__ CUtensorMap map_wl, const NaiveConvParams p, const int n_tile, const int tps) {
  extern __shared__ uint8_t tx_smem_raw[];
  const uint32_t raw_addr = smem_u32(tx_smem_raw);
  uint8_t* smem = tx_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // a weight stage holds `tps` taps: 1 (ring of 4 stages), or one accumulator chunk (ring of 2) where that fits -- the
  // per-stage hand-offs (a barrier wait ~200 clk, a commit ~200 clk in the issuing thread) then happen once per chunk
  const int SB = tps > 1 ? 2 : kThBStages;
  const int b_tap_bytes = 2 * n_tile * 128;
  const int b_stage_bytes = tps * b_tap_bytes;
  uint8_t* a_op = smem;                              // [2 stages][hi | lo]
  uint8_t* b_op = smem + 4 * kThATile;               // [SB stages][hi | lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_op + SB * b_stage_bytes);
  uint64_t* b_full = bars;                           // [<= 4] TMA landed
  uint64_t* b_empty = bars + kThBStages;             // [<= 4] MMAs have read the stage
  uint64_t* a_full = bars + 2 * kThBStages;          // [2] raw halo slice landed
  uint64_t* a_ready = a_full + 2;                    // [2] split done (8 worker warps)
  uint64_t* a_empty = a_ready + 2;                   // [2] MMAs have read the stage
  uint64_t* acc_full = a_empty + 2;                  // [2] chunk complete in TMEM
  uint64_t* acc_empty = acc_full + 2;                // [2] chunk drained (8 worker warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  int* s_tap_e = reinterpret_cast<int*>(tmem_slot + 1);                      // [9] entry index of tap i of this group
  uint32_t* s_tap_a = reinterpret_cast<uint32_t*>(s_tap_e + 9);              // [9] byte offset of tap i inside the halo tile
  float* s_epi = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][n_tile] epilogue scale, shift

  float* __restrict__ out = reinterpret_cast<float*>(p.out);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_w = p.W / 8, tiles_h = p.H / 16;
  const int n_items_all = p.n_img * tiles_w * tiles_h;
  // persistent: this CTA walks items blockIdx.x, blockIdx.x + gridDim.x, ... ; stage / chunk counters run on across
  // items, so the producer and the MMA thread start the next item while the workers are still in the epilogue
  const int my_items = (n_items_all - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  auto item_origin = [&](int k, int& img, int& y0, int& x0) {      // k-th item of this CTA
    const int item = blockIdx.x + k * gridDim.x;
    img = item / (tiles_w * tiles_h);
    const int rem = item - img * (tiles_w * tiles_h);
    y0 = (rem / tiles_w) * 16;
    x0 = (rem % tiles_w) * 8;
  };
  const int n0 = blockIdx.y * n_tile, g = blockIdx.z;

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * n_tile)) tmem_cols <<= 1;
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  // two MMA-issuing threads (warps 8 and 10) take alternate accumulator chunks: one thread's serial chain per chunk -- three
  // barrier waits of ~250 clk, 36 MMA issues, two commits -- is longer than the chunk's execution on the narrow layers
  const int n_taps_e = p.n_entries_total / p.n_groups;
  const bool two_issuers = (n_taps_e % 3 == 0 ? n_taps_e / 3 : (n_taps_e % 2 == 0 ? n_taps_e / 2 : n_taps_e)) >= 2;
  if (tid == 256) {
    for (int i = 0; i < SB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1); mbar_init(&a_ready[i], 8); mbar_init(&a_empty[i], two_issuers ? 2 : 1);
      mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8);
    }
    fence_barrier_init();
    int n = 0;
    for (int e = 0; e < p.n_entries_total && n < 9; ++e)
      if (p.entries[e].group == g) {
        s_tap_e[n] = e;
        s_tap_a[n] = static_cast<uint32_t>(((p.entries[e].dy + 1) * kThHaloW + p.entries[e].dx + 1) * 128);
        ++n;
      }
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_wh);
    tma_prefetch_desc(&map_wl);
  }
  for (int i = tid; i < n_tile; i += blockDim.x) {   // epilogue constants of this N tile (identity beyond Cout)
    const int c = n0 + i;
    s_epi[i] = (p.epi_scale && c < p.Cout) ? p.epi_scale[c] : 1.f;
    s_epi[n_tile + i] = (p.epi_shift && c < p.Cout) ? p.epi_shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_taps = p.n_entries_total / p.n_groups;               // the dispatcher guarantees 3..9 taps, all in the halo
  const int chunk_taps = (n_taps % 3 == 0) ? 3 : ((n_taps % 2 == 0) ? 2 : 1);
  const int cph = n_taps / chunk_taps;                             // accumulator chunks per halo slice
  const int n_hs = (p.Cin + kTxSliceK - 1) / kTxSliceK;            // halo slices per item
  const int n_gs = my_items * n_hs;                                // halo slices of this CTA (global slice index gs)

  if (warp == 9) {
    // ================================================= producer: every TMA of this CTA
    if (elect_one()) {
      const uint32_t a_tx = kThRows * 128;
      int l_k = 0, l_hs = 0, l_gs = 0;               // next halo slice to load: item index, slice in item, global index
      auto load_next_a = [&]() {
        if (l_gs < n_gs) {
          int img, y0, x0;
          item_origin(l_k, img, y0, x0);
          uint64_t* bar = &a_full[l_gs & 1];
          mbar_expect_tx(bar, a_tx);
          tma_load_4d(&map_a, bar, a_op + (l_gs & 1) * 2 * kThATile, p.in_choff + l_hs * kTxSliceK, x0 - 1, y0 - 1, img);
          ++l_gs;
          if (++l_hs == n_hs) { l_hs = 0; ++l_k; }
        }
      };
      load_next_a();
      load_next_a();
      int st = 0, use = 0;                           // tap stage being filled and how often it has been filled before
      int hs = 0;
      for (int gs = 0; gs < n_gs; ++gs) {
        for (int tap = 0; tap < n_taps; tap += tps) {
          if (use >= 1) mbar_wait(&b_empty[st], (use - 1) & 1);              // the MMAs that last read this stage are done
          uint64_t* bar = &b_full[st];
          uint8_t* dst = b_op + st * b_stage_bytes;
          mbar_expect_tx(bar, static_cast<uint32_t>(b_stage_bytes));
          for (int i = 0; i < tps; ++i) {
            tma_load_3d(&map_wh, bar, dst + i * b_tap_bytes, hs * kTxSliceK, n0, s_tap_e[tap + i]);
            tma_load_3d(&map_wl, bar, dst + i * b_tap_bytes + n_tile * 128, hs * kTxSliceK, n0, s_tap_e[tap + i]);
          }
          if (++st == SB) { st = 0; ++use; }
          if (tap == 0 && gs >= 1 && gs + 1 < n_gs) {
            // raw halo of slice gs + 1, once the MMAs of slice gs - 1 have left its stage -- AFTER the first weight stage
            // of this slice: waiting here first kept those weights from being requested until slice gs - 1 had finished
            // (1.2-1.5 k clk of exposed TMA latency at every slice start, traced)
            mbar_wait(&a_empty[(gs + 1) & 1], ((gs - 1) >> 1) & 1);
            load_next_a();
          }
        }
        if (++hs == n_hs) hs = 0;
      }
    }
    __syncwarp();
  } else if (warp == 8 || warp == 10) {
    // ================================================= MMA issuers
    const int which = warp == 8 ? 0 : 1;
    if ((which == 0 || two_issuers) && elect_one()) {
      const uint32_t idesc = make_idesc_tf32(static_cast<uint32_t>(n_tile));
      const uint32_t a_op_addr = smem_u32(a_op), b_op_addr = smem_u32(b_op);
      const uint32_t b_desc_hi = sw128_desc_hi(1024), a_desc_hi = sw128_desc_hi(kThHaloW * 128);
      int st = 0, use = 0, spos = 0;                 // weight stage of the current tap, how often it has been used, tap inside it
      int chunk = 0, cpos = 0;
      for (int gs = 0; gs < n_gs; ++gs) {
        mbar_wait_probe(&a_ready[gs & 1], (gs >> 1) & 1);
        tc_fence_after();
        const uint32_t a_hi_addr = a_op_addr + static_cast<uint32_t>((gs & 1) * 2 * kThATile);
        for (int tap = 0; tap < n_taps; ++tap) {
          const bool mine = !two_issuers || (chunk & 1) == which;
          if (mine) {
            if (cpos == 0 && chunk >= 2) {
              mbar_wait_probe(&acc_empty[chunk & 1], ((chunk >> 1) - 1) & 1);       // chunk - 2 has been drained
              tc_fence_after();
            }
            if (spos == 0) mbar_wait_probe(&b_full[st], use & 1);
            const uint32_t d = tmem_base + static_cast<uint32_t>((chunk & 1) * n_tile);
            const uint32_t ta = a_hi_addr + s_tap_a[tap];
            const uint32_t tb = b_op_addr + static_cast<uint32_t>(st * b_stage_bytes + spos * b_tap_bytes);
            const uint64_t dah = (static_cast<uint64_t>(a_desc_hi) << 32) | sw128_desc_lo(ta);
            const uint64_t dal = (static_cast<uint64_t>(a_desc_hi) << 32) | sw128_desc_lo(ta + kThATile);
            const uint64_t dbh = (static_cast<uint64_t>(b_desc_hi) << 32) | sw128_desc_lo(tb);
            const uint64_t dbl = (static_cast<uint64_t>(b_desc_hi) << 32) | sw128_desc_lo(tb + n_tile * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {              // 8 fp32 = 32 B = 2 descriptor units per K-step
              umma_tf32_ss(d, dah + 2 * k, dbh + 2 * k, idesc, (cpos == 0 && k == 0) ? 0u : 1u);
              umma_tf32_ss(d, dal + 2 * k, dbh + 2 * k, idesc, 1u);
              umma_tf32_ss(d, dah + 2 * k, dbl + 2 * k, idesc, 1u);
            }
          }
          if (++spos == tps) {
            if (mine) umma_commit(&b_empty[st]);
            spos = 0;
            if (++st == SB) { st = 0; ++use; }
          }
          if (cpos == chunk_taps - 1) {
            if (mine) {
              umma_commit(&acc_full[chunk & 1]);
              // this thread's last chunk of the slice: its MMAs are out of the A stage once this commit fires
              const int left = cph - 1 - (tap / chunk_taps);             // chunks of this slice after this one
              if (!two_issuers ? left == 0 : left <= 1) umma_commit(&a_empty[gs & 1]);
            }
            cpos = 0; ++chunk;
          } else {
            ++cpos;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================= workers: split, drain, epilogue
    uint32_t a_sw[kThItems];
#pragma unroll
    for (int i = 0; i < kThItems; ++i) {
      const int q = tid + 256 * i;
      const int R = q >> 3, j = q & 7;
      a_sw[i] = static_cast<uint32_t>((R >> 3) * 1024 + (R & 7) * 128 + ((j ^ (R & 7)) << 4));
    }
    auto convert = [&](int gs) {                     // hi tile (raw, as TMA delivered it) -> hi in place, lo beside it
      uint8_t* hi_tile = a_op + (gs & 1) * 2 * kThATile;
      uint8_t* lo_tile = hi_tile + kThATile;
      mbar_wait(&a_full[gs & 1], (gs >> 1) & 1);
#pragma unroll
      for (int i = 0; i < kThItems; ++i) {
        if (tid + 256 * i < kThRows * 8) {
          const float4 v = *reinterpret_cast<const float4*>(hi_tile + a_sw[i]);
          float4 h, l;
          tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
          *reinterpret_cast<float4*>(hi_tile + a_sw[i]) = h;
          *reinterpret_cast<float4*>(lo_tile + a_sw[i]) = l;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(&a_ready[gs & 1]);
    };
    const int quad = warp & 3, half = warp >> 2;
    float sum[4][16];
    auto drain = [&](int chunk) {                    // acc[chunk & 1] -> sum (round-to-nearest adds)
      mbar_wait(&acc_full[chunk & 1], (chunk >> 1) & 1);
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
      mbar_arrive_warp(&acc_empty[chunk & 1]);
    };
    const bool vec_ok = ((p.out_ctot | p.out_choff) & 3) == 0;
    if (n_gs > 0) convert(0);
    int chunk = 0, gs = 0;
    for (int k = 0; k < my_items; ++k) {
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 16; ++b) sum[a][b] = 0.f;
      for (int hs = 0; hs < n_hs; ++hs, ++gs)
        for (int ci = 0; ci < cph; ++ci) {
          drain(chunk++);
          if (ci == 0 && gs + 1 < n_gs) convert(gs + 1);             // while the tensor pipe works on the rest of slice gs
        }
      // ---- epilogue of item k: accumulator row m = 8 r + x  ->  output pixel (y0 + r, x0 + x)
      int img, y0, x0;
      item_origin(k, img, y0, x0);
      const int m = quad * 32 + lane;
      const int oy = y0 + (m >> 3), ox = x0 + (m & 7);
      const long long opix = p.up2 ? (static_cast<long long>(img) * 2 * p.H + 2 * oy + (g >> 1)) * (2 * p.W) + 2 * ox + (g & 1)
                                   : (static_cast<long long>(img) * p.H + oy) * p.W + ox;
      float* orow = out + opix * p.out_ctot + p.out_choff;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int u = half + 2 * a;
        if (u * 16 >= n_tile) continue;
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
              val = fmaf(val, s_epi[c - n0], s_epi[n_tile + c - n0]);
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
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------------
// 1x1 variant (stride 1): the dense layers' bottleneck convs with their BN-ReLU prologue, the transition convs and the
// pointwise convs of DeepLab / Inception.  The activation tensor is the flat [pixels, C] matrix: a 2-D TMA box of
// 128 pixels x 32 channels lands in the swizzled `hi` tile of a ring stage together with the pre-split weight boxes of
// the slice; the worker warps apply the prologue and split in place, the MMA warp issues 12 MMAs per slice, accumulator
// chunks are two slices.  Same roles and barriers as the halo kernel; CTAs are persistent over 128-pixel tiles.
constexpr int kT1Stages = 3;
constexpr int kT1Lag = 2;                           // worker drains run this many slices behind their conversions

__host__ __device__ inline int t1_stage_bytes(int n) { return 2 * kTxATile + 2 * n * 128; }
__host__ __device__ inline int t1_smem_bytes(int n) { return 1024 + kT1Stages * t1_stage_bytes(n) + 256 + 1024; }

__global__ void __launch_bounds__(kThThreads, 1)
conv_1x1_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_wh,
                       const __grid_constant__ CUtensorMap map_wl, const NaiveConvParams p, const int n_tile) {
  extern __shared__ uint8_t tx_smem_raw[];
  const uint32_t raw_addr = smem_u32(tx_smem_raw);
  uint8_t* smem = tx_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  constexpr int S = kT1Stages;
  const int stage_bytes = t1_stage_bytes(n_tile);    // [A hi | A lo | B hi | B lo]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* raw_full = bars;                         // [S] TMA landed (activations + weights)
  uint64_t* ready = bars + S;                        // [S] split done (8 worker warps)
  uint64_t* empty = bars + 2 * S;                    // [S] MMAs have read the stage
  uint64_t* acc_full = bars + 3 * S;                 // [2]
  uint64_t* acc_empty = acc_full + 2;                // [2] (8 worker warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_epi = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);   // [2][n_tile] epilogue scale, shift

  float* __restrict__ out = reinterpret_cast<float*>(p.out);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long M = static_cast<long long>(p.n_img) * p.H * p.W;
  const int n_mtiles = static_cast<int>((M + 127) / 128);
  const int my_tiles = (n_mtiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int n0 = blockIdx.y * n_tile;
  const int n_s = (p.Cin + kTxSliceK - 1) / kTxSliceK;             // slices per tile
  const int n_gs = my_tiles * n_s;

  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(2 * n_tile)) tmem_cols <<= 1;
  if (warp == 0) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  if (tid == 256) {
    for (int i = 0; i < S; ++i) { mbar_init(&raw_full[i], 1); mbar_init(&ready[i], 8); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
    fence_barrier_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_wh);
    tma_prefetch_desc(&map_wl);
  }
  for (int i = tid; i < n_tile; i += blockDim.x) {   // epilogue constants of this N tile (identity beyond Cout)
    const int c = n0 + i;
    s_epi[i] = (p.epi_scale && c < p.Cout) ? p.epi_scale[c] : 1.f;
    s_epi[n_tile + i] = (p.epi_shift && c < p.Cout) ? p.epi_shift[c] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // ================================================= producer
    if (elect_one()) {
      const uint32_t tx = static_cast<uint32_t>(kTxATile + 2 * n_tile * 128);
      int st = 0, use = 0, hs = 0, k = 0;
      for (int gs = 0; gs < n_gs; ++gs) {
        if (use >= 1) mbar_wait(&empty[st], (use - 1) & 1);
        uint8_t* base = smem + st * stage_bytes;
        const int m0 = (static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x)) * 128;
        mbar_expect_tx(&raw_full[st], tx);
        tma_load_2d(&map_a, &raw_full[st], base, p.in_choff + hs * kTxSliceK, m0);
        tma_load_3d(&map_wh, &raw_full[st], base + 2 * kTxATile, hs * kTxSliceK, n0, 0);
        tma_load_3d(&map_wl, &raw_full[st], base + 2 * kTxATile + n_tile * 128, hs * kTxSliceK, n0, 0);
        if (++st == S) { st = 0; ++use; }
        if (++hs == n_s) { hs = 0; ++k; }
      }
    }
    __syncwarp();
  } else if (warp == 8) {
    // ================================================= MMA issuer
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(static_cast<uint32_t>(n_tile));
      const uint32_t smem_addr = smem_u32(smem), desc_hi = sw128_desc_hi(1024);
      int st = 0, use = 0, hs = 0, chunk = 0;
      for (int gs = 0; gs < n_gs; ++gs) {
        const bool chunk_start = (hs & 1) == 0, chunk_end = (hs & 1) == 1 || hs == n_s - 1;
        if (chunk_start && chunk >= 2) {
          mbar_wait_probe(&acc_empty[chunk & 1], ((chunk >> 1) - 1) & 1);
          tc_fence_after();
        }
        mbar_wait_probe(&ready[st], use & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + static_cast<uint32_t>((chunk & 1) * n_tile);
        const uint32_t ta = smem_addr + static_cast<uint32_t>(st * stage_bytes);
        const uint64_t dah = (static_cast<uint64_t>(desc_hi) << 32) | sw128_desc_lo(ta);
        const uint64_t dal = (static_cast<uint64_t>(desc_hi) << 32) | sw128_desc_lo(ta + kTxATile);
        const uint64_t dbh = (static_cast<uint64_t>(desc_hi) << 32) | sw128_desc_lo(ta + 2 * kTxATile);
        const uint64_t dbl = (static_cast<uint64_t>(desc_hi) << 32) | sw128_desc_lo(ta + 2 * kTxATile + n_tile * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_tf32_ss(d, dah + 2 * k, dbh + 2 * k, idesc, (chunk_start && k == 0) ? 0u : 1u);
          umma_tf32_ss(d, dal + 2 * k, dbh + 2 * k, idesc, 1u);
          umma_tf32_ss(d, dah + 2 * k, dbl + 2 * k, idesc, 1u);
        }
        umma_commit(&empty[st]);
        if (chunk_end) { umma_commit(&acc_full[chunk & 1]); ++chunk; }
        if (++st == S) { st = 0; ++use; }
        if (++hs == n_s) hs = 0;
      }
    }
    __syncwarp();
  } else {
    // ================================================= workers: prologue + split, drain, epilogue
    // thread -> 16-byte chunk j (4 channels) of rows (tid >> 3) + 32 i: the prologue terms depend on j only
    const int j = tid & 7, r0 = tid >> 3;
    const uint32_t sw_row = static_cast<uint32_t>((r0 >> 3) * 1024 + (r0 & 7) * 128 + ((j ^ (r0 & 7)) << 4));
    const int quad = warp & 3, half = warp >> 2;
    const bool vec_ok = ((p.out_ctot | p.out_choff) & 3) == 0;
    float sum[4][16];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 16; ++b) sum[a][b] = 0.f;
    auto drain = [&](int chunk) {
      mbar_wait(&acc_full[chunk & 1], (chunk >> 1) & 1);
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
      mbar_arrive_warp(&acc_empty[chunk & 1]);
    };
    auto epilogue = [&](int k) {                     // k-th tile of this CTA; clears the sums
      const long long mm = (static_cast<long long>(blockIdx.x) + static_cast<long long>(k) * gridDim.x) * 128 + quad * 32 + lane;
      float* orow = out + mm * p.out_ctot + p.out_choff;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int u = half + 2 * a;
        if (u * 16 < n_tile && mm < M) {
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
                val = fmaf(val, s_epi[c - n0], s_epi[n_tile + c - n0]);
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
#pragma unroll
        for (int b = 0; b < 16; ++b) sum[a][b] = 0.f;
      }
    };
    // slices are converted in order; the slice `kT1Lag` behind is "retired": its chunk drained, its tile written
    int st = 0, use = 0, hs = 0;                     // conversion cursor
    int r_hs = 0, r_k = 0, r_chunk = 0, retired = 0; // retire cursor
    auto retire_one = [&]() {
      const bool chunk_end = (r_hs & 1) == 1 || r_hs == n_s - 1;
      if (chunk_end) drain(r_chunk++);
      if (++r_hs == n_s) { r_hs = 0; epilogue(r_k++); }
      ++retired;
    };
    for (int gs = 0; gs < n_gs; ++gs) {
      uint8_t* hi_tile = smem + st * stage_bytes;
      uint8_t* lo_tile = hi_tile + kTxATile;
      const int c = hs * kTxSliceK + 4 * j;
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.pro_mode && c < p.Cin) {
        sc = *reinterpret_cast<const float4*>(p.pro_scale + c);
        sh = *reinterpret_cast<const float4*>(p.pro_shift + c);
      }
      mbar_wait(&raw_full[st], use & 1);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t off = static_cast<uint32_t>(i * 4 * 1024) + sw_row;
        float4 v = *reinterpret_cast<const float4*>(hi_tile + off);
        if (p.pro_mode) {
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
          if (p.pro_mode == 2) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        }
        float4 h, l;
        tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi_tile + off) = h;
        *reinterpret_cast<float4*>(lo_tile + off) = l;
      }
      fence_proxy_async_smem();
      mbar_arrive_warp(&ready[st]);
      if (++st == S) { st = 0; ++use; }
      if (++hs == n_s) hs = 0;
      if (gs + 1 - retired > kT1Lag) retire_one();
    }
    while (retired < n_gs) retire_one();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// data section -> hi / lo copies (run once at model creation for precision 2): hi and lo both rounded to TF32
__global__ void tf32_presplit_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    uint32_t b = __float_as_uint(v);
    const bool finite = (b & 0x7F800000u) != 0x7F800000u;
    const float h = finite ? __uint_as_float((b + 0x1000u) & 0xFFFFE000u) : v;
    const float r = finite ? v - h : 0.f;
    hi[i] = h;
    lo[i] = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
  }
}

}  // namespace dp
