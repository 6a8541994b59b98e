// One DenseNet layer per kernel: BN-ReLU -> 1x1 conv (C -> 128) -> BN-ReLU -> 3x3 conv (128 -> 32) -> concat
// (reference: dense_conv_block, DigiPathAI/models/densenet.py:50-75), with the 128-channel bottleneck kept in
// shared memory instead of a round trip through HBM and a second launch.
//
// Work item = one RH x 8 pixel region of one image (RH = 16 or 8).  Per item:
//   phase 1  for every 64-channel chunk of the concat buffer: TMA-load the region WITH its 1-pixel halo
//            ((RH+2) x 10 pixels = 180 / 100 rows of 128 B), pre-activation BN+ReLU in place (8 transform
//            warps), then tcgen05.mma over the halo rows (2 / 1 M-blocks of 128 rows, rows beyond the box are
//            don't-care), N = 128, accumulating in TMEM.  The halo pixels' bottleneck values are recomputed by
//            every region that needs them (1.4-1.6x the 1x1 MACs) -- the price of not synchronising CTAs.
//   mid      epilogue warps: TMEM -> +BN shift -> ReLU -> zero outside the image (the 3x3's `same` padding is
//            applied to the bottleneck, AFTER its BN-ReLU) -> fp16 -> written as two swizzled K-major operand
//            tiles T[2 chunks][rows][64 ch] in shared memory, fence.proxy.async.
//   phase 2  the 3x3 conv as 9 taps x 2 chunks of UMMA descriptors into T (row offset (dy+1)*10 + dx+1,
//            SBO = 10 * 128 B), N = 32 (for RH = 8 the upper 64 accumulator rows are don't-care).
//   final    TMEM -> fp16 -> the layer's 32 new channels in the concat buffer (256-bit stores).
//
// Items are software-pipelined: T and the 3x3 accumulator are double-buffered, so the MMA warp issues
//   ph1(0) | ph1(1) ph2(0) | ph1(2) ph2(1) | ...      and the epilogue warps run   mid(0) | mid(1) fin(0) | ...
// i.e. the tensor pipe works on item k+1's 1x1 while the epilogue warps turn item k's accumulator into T, and
// on item k's 3x3 while they do mid(k+1).
//
// Cross-layer overlap: only the activation producer executes griddepcontrol.wait, and only before the first
// chunk that contains the preceding layer's 32 new channels; older chunks are consumed while that layer runs.
//
// Tried and removed (round 2): a STACK variant that stacked the three dy taps of one dx along N (N = 96, 24 MMAs per
// item instead of 72) and recombined the row-shifted accumulator groups in the final epilogue.  It ran correctly on
// the B200 but did not shorten the step (17 075 vs 17 147 tiles/s with 8-row regions, 16 089 with 14-row regions:
// profiles/r2_bench_stack*.json) -- the item period is set by the ph1 -> mid -> ph1 loop, not by phase 2.
//
// Warp roles: 0 = TMA producer (activation halo chunks), 1 = MMA issuer, 2 = TMEM allocator, 3 = TMA producer
// (W1 chunks / W2 tap groups, one ring, in MMA order), 4-7 and 16-19 = epilogue (mid + final), 8-15 =
// pre-activation transform (BN terms of a thread's 8 channels stay in registers).
#pragma once
#include "conv_tc.cuh"

namespace dp {

constexpr int kDlHaloW = 10;              // 8-pixel-wide regions + 1-pixel halo each side
constexpr int kDlBStage = 128 * 128;      // W1 chunk [128 x 64]; W2 groups (3 taps x [32 x 64] = 12 KB) fit too
constexpr int kDlW2Group = 3;             // taps per W2 stage

__host__ __device__ constexpr int dl_rows(int rh) { return kDlHaloW * (rh + 2); }
// A-stage stride: the halo box rounded up to 1 KB.  The last M-block reads 128 rows regardless, i.e. up to 9 KB
// past the stage end -- into the next stage or the weight ring, always inside the allocation, never used.
__host__ __device__ constexpr int dl_a_stage(int rh) { return (dl_rows(rh) * 128 + 1023) / 1024 * 1024; }
// One 64-channel chunk of the bottleneck tile T (halo rows x 128 B, 1 KB aligned): 23 KB for 16-row regions, 13 KB
// for 8-row regions.  T = 2 chunks (128 channels) x 2 buffers.  The 3x3 MMAs read 128 rows from every tap offset,
// i.e. up to ~7 KB past the tile for 8-row regions: T sits FIRST in shared memory so that this over-read lands in
// the activation ring (don't-care accumulator rows), never outside the allocation.
__host__ __device__ constexpr int dl_t_chunk(int rh) { return dl_a_stage(rh); }
__host__ __device__ constexpr int dl_t_bytes(int rh) { return 4 * dl_t_chunk(rh); }

struct DenseLayerParams {
  int n_img, H, W, C;        // map size, input channels of this layer
  int n_chunks;              // ceil(C / 64)
  int tiles_w, tiles_h, n_items;
  int n_safe_chunks;         // leading channel chunks not written by the immediately preceding kernel: they are
                             // consumed BEFORE griddepcontrol.wait, i.e. while the previous layer still runs
  int rh;                    // region height: 16 (180 halo rows, 2 M-blocks) or 8 (100 halo rows, 1 M-block)
  int a_stages, b_stages;
  int out_ctot, out_choff;   // concat buffer channel stride, offset of the 32 new channels
  const float* pro_scale;    // BN1 [n_chunks * 64]
  const float* pro_shift;
  const float* mid_shift;    // BN2 shift [128] (scale folded into W1)
  __half* out;               // concat buffer base (same tensor the A map reads)
  unsigned long long* trace;
  unsigned long long* gt;    // debug: [0]/[1] receive %globaltimer at entry / exit of CTA 0
  unsigned long long* gt_all;  // debug (option "stamp_ctas"): per CTA [entry, grid-dependency wait returned, exit, SM id]
};

struct DenseLayerSmem {
  static constexpr int kBarBytes = 1024;
  int a_off, b_off, t_off, pro_off, mid_off, total;
};

__host__ __device__ inline DenseLayerSmem dense_layer_smem(const DenseLayerParams& p) {
  DenseLayerSmem L;
  L.t_off = DenseLayerSmem::kBarBytes;
  L.a_off = L.t_off + dl_t_bytes(p.rh);
  L.b_off = L.a_off + p.a_stages * dl_a_stage(p.rh);
  L.pro_off = L.b_off + p.b_stages * kDlBStage;
  L.mid_off = L.pro_off + 2 * p.n_chunks * 64 * 4;
  L.total = L.mid_off + 128 * 4 + 1024;
  return L;
}

struct DlTrace {
  unsigned long long* base = nullptr;
  unsigned n = 0;
};
__device__ __forceinline__ DlTrace dl_trace_open(unsigned long long* trace, unsigned role) {
  DlTrace c;
  if (trace && blockIdx.x == 0) c.base = trace + 8 + 2000 * role;
  return c;
}
__device__ __forceinline__ void dl_trace_ev(DlTrace& c, unsigned ev, unsigned item) {
  if (c.base && c.n < 2000)
    c.base[c.n++] = (static_cast<unsigned long long>(ev) << 48) | (static_cast<unsigned long long>(item & 0xFFFF) << 32) |
                    (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFull);
}
__device__ __forceinline__ void dl_trace_close(unsigned long long* trace, const DlTrace& c, unsigned role) {
  if (c.base) trace[role] = c.n;
}

template <int RH>
__global__ void __launch_bounds__(640, 1)
dense_layer_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w1,
                   const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ DenseLayerParams p) {
  constexpr int kRows = dl_rows(RH), kMBlk = (kRows + 127) / 128, kAStage = dl_a_stage(RH);
  constexpr int kDlTChunk = dl_t_chunk(RH), kDlTBuf = 2 * kDlTChunk;
  constexpr int kAcc2Cols = 32;                    // accumulator columns of one 3x3 buffer
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_ready = a_full + kMaxAStages;
  uint64_t* a_empty = a_ready + kMaxAStages;
  uint64_t* b_full = a_empty + kMaxAStages;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* acc1_full = b_empty + kMaxBStages;
  uint64_t* acc1_empty = acc1_full + 1;
  uint64_t* t_ready = acc1_empty + 1;    // [2]
  uint64_t* t_empty = t_ready + 2;       // [2]
  uint64_t* acc2_full = t_empty + 2;     // [2]
  uint64_t* acc2_empty = acc2_full + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 2);

  const DenseLayerSmem L = dense_layer_smem(p);
  uint8_t* a_base = smem + L.a_off;
  uint8_t* b_base = smem + L.b_off;
  uint8_t* t_base = smem + L.t_off;
  float* s_pro_scale = reinterpret_cast<float*>(smem + L.pro_off);
  float* s_pro_shift = s_pro_scale + p.n_chunks * 64;
  float* s_mid_shift = reinterpret_cast<float*>(smem + L.mid_off);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0 && p.trace && blockIdx.x == 0) p.trace[8 + 2000 * 4 + 1] = (2ull << 48) | (clock64() & 0xFFFFFFFFull);
  if (tid == 0 && p.gt && blockIdx.x == 0) p.gt[0] = globaltimer_ns();
  if (tid == 0 && p.gt_all) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    p.gt_all[4 * blockIdx.x] = globaltimer_ns();
    p.gt_all[4 * blockIdx.x + 3] = smid;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_w2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 8);      // per-warp arrivals (mbar_arrive_warp), 8 transform warps
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 8);       // 8 epilogue warps
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_ready[i], 8);
      mbar_init(&t_empty[i], 1);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  {
    const int npro = p.n_chunks * 64;
    for (int i0 = 0; i0 < npro || i0 < 128; i0 += blockDim.x) {
      const int i = i0 + tid;
      float ps = 0.f, ph = 0.f, ms = 0.f;
      if (i < npro) { ps = p.pro_scale[i]; ph = p.pro_shift[i]; }
      if (i < 128) ms = p.mid_shift[i];
      if (i < npro) { s_pro_scale[i] = ps; s_pro_shift[i] = ph; }
      if (i < 128) s_mid_shift[i] = ms;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // No grid-dependency wait here: only the activation producer needs it (see header).
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0 && p.trace && blockIdx.x == 0) { p.trace[8 + 2000 * 4] = clock64() & 0xFFFFFFFFull; p.trace[4] = 2; }
  const uint32_t acc1_col = tmem_base;        // kMBlk x 128 columns
  const uint32_t acc2_col = tmem_base + 256;  // 2 x 32 columns
  const int first = blockIdx.x;
  const int n_local = (first < p.n_items) ? (p.n_items - first + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;

  auto item_origin = [&](int item, int& n0, int& h0, int& w0) {
    const int tw = item % p.tiles_w;
    const int r = item / p.tiles_w;
    w0 = tw * 8;
    h0 = (r % p.tiles_h) * RH;
    n0 = r / p.tiles_h;
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: activation halo chunks
    if (elect_one()) {
      DlTrace tc = dl_trace_open(p.trace, 0);
      uint32_t sa = 0, pa = 0;
      bool waited = false;
      for (int k = 0; k < n_local; ++k) {
        const int item = first + k * gridDim.x;
        int n0, h0, w0;
        item_origin(item, n0, h0, w0);
        for (int c = 0; c < p.n_chunks; ++c) {
          if (!waited && c >= p.n_safe_chunks) {
            // From here on we need the preceding layer's 32 new channels.  It triggered this launch only after
            // its OWN wait, so everything older was already complete when we started.
            pdl_wait();
            pdl_launch_dependents();
            waited = true;
            if (p.gt_all) p.gt_all[4 * blockIdx.x + 1] = globaltimer_ns();
          }
          mbar_wait(&a_empty[sa], pa ^ 1);
          mbar_expect_tx(&a_full[sa], kRows * 128);
          tma_load_4d(&map_x, &a_full[sa], a_base + sa * kAStage, c * 64, w0 - 1, h0 - 1, n0);
          dl_trace_ev(tc, 1, item);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
        }
      }
      if (!waited) { pdl_wait(); pdl_launch_dependents(); }
      dl_trace_close(p.trace, tc, 0);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ producer: weights, in MMA order
    if (elect_one()) {
      uint32_t sb = 0, pb = 0;
      auto load_w1 = [&]() {
        for (int c = 0; c < p.n_chunks; ++c) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], 128 * 128);
          tma_load_3d(&map_w1, &b_full[sb], b_base + sb * kDlBStage, c * 64, 0, 0);
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
      };
      auto load_w2 = [&]() {
        for (int c = 0; c < 2; ++c)
          for (int g = 0; g < 9 / kDlW2Group; ++g) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_expect_tx(&b_full[sb], kDlW2Group * 32 * 128);
            tma_load_3d(&map_w2, &b_full[sb], b_base + sb * kDlBStage, c * 64, 0, g * kDlW2Group);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
      };
      if (n_local > 0) load_w1();
      for (int k = 0; k < n_local; ++k) {
        if (k + 1 < n_local) load_w1();
        load_w2();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      DlTrace tc = dl_trace_open(p.trace, 1);
      const uint32_t idesc1 = make_idesc_f16(128), idesc2 = make_idesc_f16(kAcc2Cols);
      const uint64_t hi_dense = static_cast<uint64_t>(sw128_desc_hi(1024)) << 32;
      const uint64_t hi_halo = static_cast<uint64_t>(sw128_desc_hi(kDlHaloW * 128)) << 32;
      const uint64_t a_desc0 = hi_dense | sw128_desc_lo(smem_u32(a_base));
      const uint64_t b_desc0 = hi_dense | sw128_desc_lo(smem_u32(b_base));
      const uint64_t t_desc0 = hi_halo | sw128_desc_lo(smem_u32(t_base));
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      // ---- phase 1 of local item k: bottleneck accumulator = W1 * relu(bn(x)) over the halo rows
      auto ph1 = [&](int k) {
        mbar_wait(acc1_empty, (k & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < p.n_chunks; ++c) {
          int ks = (p.C - c * 64 + 15) >> 4;
          ks = ks > 4 ? 4 : ks;
          mbar_wait(&a_ready[sa], pa);
          dl_trace_ev(tc, 1, k);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + sa * (kAStage >> 4);
          const uint64_t b_desc = b_desc0 + sb * (kDlBStage >> 4);
          const uint32_t acc = (c > 0) ? 1u : 0u;
          if (ks == 4) {
            umma_f16_ss_k4(acc1_col, a_desc, b_desc, idesc1, acc);
            if (kMBlk > 1) umma_f16_ss_k4(acc1_col + 128, a_desc + (kATileBytes >> 4), b_desc, idesc1, acc);
          } else if (ks == 2) {
            umma_f16_ss_k2(acc1_col, a_desc, b_desc, idesc1, acc);
            if (kMBlk > 1) umma_f16_ss_k2(acc1_col + 128, a_desc + (kATileBytes >> 4), b_desc, idesc1, acc);
          } else {
            for (int kk = 0; kk < ks; ++kk) {
              umma_f16_ss(acc1_col, a_desc + 2 * kk, b_desc + 2 * kk, idesc1, kk ? 1u : acc);
              if (kMBlk > 1)
                umma_f16_ss(acc1_col + 128, a_desc + (kATileBytes >> 4) + 2 * kk, b_desc + 2 * kk, idesc1, kk ? 1u : acc);
            }
          }
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
          if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
        }
        umma_commit(acc1_full);
        dl_trace_ev(tc, 3, k);
      };
      // ---- phase 2 of local item k: 3x3 conv over the bottleneck tile T[k & 1]
      auto ph2 = [&](int k) {
        const int tb = k & 1;
        const uint32_t u = (k >> 1) & 1;
        mbar_wait(&t_ready[tb], u);
        mbar_wait(&acc2_empty[tb], u ^ 1);
        tc_fence_after();
        dl_trace_ev(tc, 4, k);
        const uint32_t d2 = acc2_col + tb * kAcc2Cols;
        for (int c = 0; c < 2; ++c) {
          const uint64_t t_desc = t_desc0 + ((tb * kDlTBuf + c * kDlTChunk) >> 4);
          for (int g = 0; g < 9 / kDlW2Group; ++g) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            uint64_t b_desc = b_desc0 + sb * (kDlBStage >> 4);
#pragma unroll
            for (int j = 0; j < kDlW2Group; ++j, b_desc += (32 * 128) >> 4) {
              const int tap = g * kDlW2Group + j;
              const int dy = tap / 3, dx = tap - dy * 3;  // (dy+1, dx+1) with dy,dx in -1..1
              umma_f16_ss_k4(d2, t_desc + (dy * kDlHaloW + dx) * 8, b_desc, idesc2, (c | tap) ? 1u : 0u);
            }
            umma_commit(&b_empty[sb]);
            if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit(&t_empty[tb]);
        umma_commit(&acc2_full[tb]);
        dl_trace_ev(tc, 5, k);
      };
      if (n_local > 0) ph1(0);
      for (int k = 0; k < n_local; ++k) {
        if (k + 1 < n_local) ph1(k + 1);
        ph2(k);
      }
      dl_trace_close(p.trace, tc, 1);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 16) {
    // ------------------------------------------------------------------ epilogue warps: mid + final
    // Two groups of four warps (a warp may only touch TMEM lanes 32 (w % 4) .. +31): group 0 (warps 4-7) takes
    // bottleneck channels 0-63 and output channels 0-15, group 1 (warps 16-19) channels 64-127 and 16-31.  The
    // mid step sits on the acc1 -> T -> acc1 critical loop of the item pipeline, hence eight warps.
    const int q = warp & 3;
    const int half = (warp >= 16) ? 1 : 0;
    const int r = q * 32 + lane;
    DlTrace tc;
    if (r == 0 && half == 0) tc = dl_trace_open(p.trace, 2);
    // ---- mid(k): acc1 -> +shift -> ReLU -> zero padding -> fp16 -> swizzled operand tile T[k & 1]
    auto mid = [&](int k) {
      const int item = first + k * gridDim.x;
      int n0, h0, w0;
      item_origin(item, n0, h0, w0);
      const int tb = k & 1;
      const uint32_t u = (k >> 1) & 1;
      mbar_wait(acc1_full, k & 1);
      mbar_wait(&t_empty[tb], u ^ 1);   // the 3x3 MMAs of item k-2 no longer read this buffer
      tc_fence_after();
      dl_trace_ev(tc, 0, k);
      uint8_t* tbuf = t_base + tb * kDlTBuf;
      // 16-column steps, ping-pong register buffers: the tcgen05.ld of step s+1 is in flight while step s is
      // converted and stored (tcgen05.wait::ld waits for ALL outstanding loads, so the next load is issued right
      // after the wait).  Measured: this step takes ~5k clk per 16-row item (trace of conv2_block3) and is the
      // critical resource of the item pipeline, but hiding the load latency this way did NOT shorten it -- the
      // loads are not latency bound; what the step waits on is still open (TMEM port contention with the
      // concurrent N = 32 MMAs is the leading suspect).
      constexpr int kSteps = kMBlk * 4;
      const uint32_t lane_col = acc1_col + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
      auto process = [&](int sidx, const uint32_t (&v)[16]) {
        const int mb = sidx >> 2;
        const int cb = half * 64 + (sidx & 3) * 16;
        const int prow = mb * 128 + r;                 // halo pixel index
        if (prow >= kRows) return;
        const int hh = prow / kDlHaloW, ww = prow - hh * kDlHaloW;
        const int ih = h0 - 1 + hh, iw = w0 - 1 + ww;
        const bool inside = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        float f[16];
        epi_affine16(v, nullptr, s_mid_shift + cb, false, true, f);
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 h2 = inside ? __floats2half2_rn(f[2 * i], f[2 * i + 1]) : __float2half2_rn(0.f);
          pk[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        // channels cb..cb+15 = 16-byte chunks j, j+1 of 64-channel chunk (cb / 64)
        uint8_t* row = tbuf + (cb >> 6) * kDlTChunk + prow * 128;
        const int j = (cb & 63) >> 3;
        *reinterpret_cast<uint4*>(row + (((j) ^ (prow & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(row + (((j + 1) ^ (prow & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      {
        uint32_t v0[16], v1[16];
        tmem_ld16(lane_col, v0);
#pragma unroll 1
        for (int sidx = 0; sidx < kSteps; sidx += 2) {
          tmem_ld_wait();
          tmem_ld16(lane_col + ((sidx + 1) >> 2) * 128 + ((sidx + 1) & 3) * 16, v1);
          process(sidx, v0);
          tmem_ld_wait();
          if (sidx + 2 < kSteps) tmem_ld16(lane_col + ((sidx + 2) >> 2) * 128 + ((sidx + 2) & 3) * 16, v0);
          process(sidx + 1, v1);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive_warp(acc1_empty);
      mbar_arrive_warp(&t_ready[tb]);
      dl_trace_ev(tc, 2, k);
    };
    // ---- fin(k): acc2[k & 1] -> fp16 -> the 32 new channels of the concat buffer
    auto fin = [&](int k) {
      const int item = first + k * gridDim.x;
      int n0, h0, w0;
      item_origin(item, n0, h0, w0);
      const int tb = k & 1;
      const uint32_t u = (k >> 1) & 1;
      mbar_wait(&acc2_full[tb], u);
      tc_fence_after();
      dl_trace_ev(tc, 3, k);
      const int w = w0 + (r & 7), h = h0 + (r >> 3);
      const bool valid = (n0 < p.n_img) && ((r >> 3) < RH) && (h < p.H) && (w < p.W);
      const long long opix = (static_cast<long long>(n0) * p.H + h) * p.W + w;
      __half* orow = p.out + opix * p.out_ctot + p.out_choff;
      uint32_t v[16];
      const uint32_t taddr = acc2_col + tb * kAcc2Cols + half * 16 + (static_cast<uint32_t>(q * 32) << 16);
      tmem_ld16(taddr, v);
      tmem_ld_wait();
      {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 h2 = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          pk[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if (valid) st_global_v8(orow + 16 * half, pk);
      }
      tc_fence_before();
      mbar_arrive_warp(&acc2_empty[tb]);
      dl_trace_ev(tc, 1, k);
    };
    if (n_local > 0) mid(0);
    for (int k = 0; k < n_local; ++k) {
      if (k + 1 < n_local) mid(k + 1);
      fin(k);
    }
    if (r == 0 && half == 0) dl_trace_close(p.trace, tc, 2);
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ pre-activation BN + ReLU on the halo rows
    const int t = tid - 256;
    const __half2 zero2 = __float2half2_rn(0.f);
    DlTrace tc;
    if (t == 0) tc = dl_trace_open(p.trace, 3);
    uint32_t sa = 0, pa = 0;
    for (int k = 0; k < n_local; ++k) {
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&a_full[sa], pa);
        dl_trace_ev(tc, 1, k);
        {
          // Thread t owns logical 16-byte chunk i = t & 7 (8 channels) of rows (t >> 3) + 32 j: its 16 BN terms
          // stay in registers for the whole stage, so shared-memory traffic is the data itself (a quarter-warp
          // covers one full 128-byte row: conflict-free under the 128B swizzle).
          const int i = t & 7;
          const int ch = c * 64 + i * 8;
          const float4 sc0 = *reinterpret_cast<const float4*>(s_pro_scale + ch);
          const float4 sc1 = *reinterpret_cast<const float4*>(s_pro_scale + ch + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_pro_shift + ch);
          const float4 sh1 = *reinterpret_cast<const float4*>(s_pro_shift + ch + 4);
          uint8_t* stage_base = a_base + sa * kAStage;
          constexpr int kIter = (kRows + 31) / 32;
          uint4 raw[kIter];
#pragma unroll
          for (int j = 0; j < kIter; ++j) {
            const int rr = (t >> 3) + 32 * j;
            if (rr < kRows) raw[j] = *reinterpret_cast<const uint4*>(stage_base + rr * 128 + ((i ^ (rr & 7)) << 4));
          }
#pragma unroll
          for (int j = 0; j < kIter; ++j) {
            __half2* hv = reinterpret_cast<__half2*>(&raw[j]);
            float2 x;
            x = __half22float2(hv[0]);
            hv[0] = __hmax2(__floats2half2_rn(fmaf(x.x, sc0.x, sh0.x), fmaf(x.y, sc0.y, sh0.y)), zero2);
            x = __half22float2(hv[1]);
            hv[1] = __hmax2(__floats2half2_rn(fmaf(x.x, sc0.z, sh0.z), fmaf(x.y, sc0.w, sh0.w)), zero2);
            x = __half22float2(hv[2]);
            hv[2] = __hmax2(__floats2half2_rn(fmaf(x.x, sc1.x, sh1.x), fmaf(x.y, sc1.y, sh1.y)), zero2);
            x = __half22float2(hv[3]);
            hv[3] = __hmax2(__floats2half2_rn(fmaf(x.x, sc1.z, sh1.z), fmaf(x.y, sc1.w, sh1.w)), zero2);
          }
#pragma unroll
          for (int j = 0; j < kIter; ++j) {
            const int rr = (t >> 3) + 32 * j;
            if (rr < kRows) *reinterpret_cast<uint4*>(stage_base + rr * 128 + ((i ^ (rr & 7)) << 4)) = raw[j];
          }
        }
        dl_trace_ev(tc, 2, k);
        fence_proxy_async_smem();
        dl_trace_ev(tc, 3, k);
        mbar_arrive_warp(&a_ready[sa]);
        dl_trace_ev(tc, 0, k);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      }
    }
    if (t == 0) dl_trace_close(p.trace, tc, 3);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
    if (lane == 0 && p.gt && blockIdx.x == 0) p.gt[1] = globaltimer_ns();
    if (lane == 0 && p.gt_all) p.gt_all[4 * blockIdx.x + 2] = globaltimer_ns();
    if (lane == 0 && p.trace && blockIdx.x == 0) { p.trace[8 + 2000 * 4 + 2] = (1ull << 48) | (clock64() & 0xFFFFFFFFull); p.trace[4] = 3; }
  }
}

}  // namespace dp
