// A whole DenseNet block on small maps (16x16 / 8x8: conv4, conv5 of densenet.py:126-132) as ONE persistent kernel.
//
// Why.  One launch per layer (dense_layer.cuh) leaves these blocks latency bound: per-CTA %globaltimer stamps
// (profiles/r2_stamp_ctas.txt) show a layer period of 10.3 us on the 16x16 maps although the dependent tail of a
// layer (last chunk -> transform -> MMA -> mid -> 3x3 -> store) takes 4.4 us -- a 128-CTA kernel and its 128-CTA
// successor cannot be resident together (one CTA per SM), so the successor's setup and its "safe" channel chunks
// do not overlap the running layer, and every layer pays ~1.2 us of grid completion -> dependent release on top.
//
// Here CTA i keeps region i (8 x 8 pixels of one image) for all layers of the block.  Per layer the roles do exactly
// what dense_layer_kernel<8> does (same MMA order, same fp32 epilogue arithmetic: outputs are bit-identical), but
//   * the activation / weight rings, TMEM and barriers live across layers: no per-layer setup, and the producer
//     streams layer l+1's safe chunks (channels older than layer l's output) while layer l is still computing;
//   * the only inter-CTA dependency -- the 32 channels layer l adds, needed over the region's 1-pixel halo -- stays
//     inside one image, and the regions of one image form one thread-block CLUSTER (4 CTAs on 16x16 maps, 1 on
//     8x8): after its final stores a CTA arrives (release, cluster scope) on an mbarrier in every CTA of its cluster
//     (mapa + mbarrier.arrive.shared::cluster), and a producer waits (acquire) on its own barrier before the TMA load
//     of the one chunk that holds those channels.  Two barriers alternate by layer parity so that a fast peer's
//     arrival for layer l+1 cannot be counted into phase l.  (First version: per-region flags in global memory
//     polled with ld.acquire -- 5-6k clk from publish to seen; profiles/r2_trace_dense_block.txt.)
//   * per-layer facts (tensor maps of W1 / W2, BN vectors, channel count) come from a table in global memory; the
//     BN vectors are read through L1 (prefetched at layer start) instead of being staged in shared memory.
// Clusters are independent of each other (images are), so there is no grid-wide residency requirement.
//   * the chunk that holds layer l's 32 new channels does not make the round trip through L2 at all: the block's input
//     channel count is a multiple of 64, so that chunk is always one that was OPENED inside this kernel, and it lives
//     as a raw fp16 "tail tile" (100 halo rows x 64 channels, two buffers by chunk parity) in shared memory.  The
//     final epilogue writes a region's new channels into its own tail tile and -- for pixels on a region border --
//     into the halo rows of the neighbouring CTAs' tail tiles through distributed shared memory
//     (st.shared::cluster), then arrives on the cluster barriers; the transform warps wait for that barrier and read
//     the tail tile instead of a TMA-filled stage.  Global memory still receives every new channel (later layers load
//     it as an ordinary "safe" chunk, the decoder reads the block's output); the producer waits for layer l-2's
//     barrier -- one whole layer of slack -- before the TMA load that first touches layer l-2's channels.
//
// MMA issue order: ph1(0) | ph2(0) ph1(1) | ph2(1) ph1(2) | ...   (ph2(l) gates everything downstream, so it goes
// first; the tensor pipe then works through layer l+1's prefetched chunks while layer l is stored and published).
#pragma once
#include "dense_layer.cuh"

namespace dp {

struct __align__(64) DenseBlockLayer {
  CUtensorMap map_w1;        // [1][128][C] fp16, box 64 x 128
  CUtensorMap map_w2;        // [9][32][128] fp16, box 64 x 32 x 3
  const float* pro_scale;    // BN1 [n_chunks * 64]
  const float* pro_shift;
  const float* mid_shift;    // BN2 shift [128]
  int C, n_chunks, out_choff, pad;
};

struct DenseBlockParams {
  int n_img, H, W;
  int tiles_w, tiles_h, n_items;
  int n_layers;
  int a_stages, b_stages;
  int out_ctot;
  __half* out;                      // concat buffer base (the tensor map_x reads)
  const DenseBlockLayer* layers;    // device, [n_layers]
  int cluster_size;                 // regions per image = CTAs per cluster (1, 2, 4 or 8)
  unsigned long long* trace;        // debug (option "trace_block"): role timelines of CTA 0, item = layer (dense_layer.cuh)
  unsigned long long* gt_layers;    // debug (option "stamp"): [n_layers][2] %globaltimer of CTA 0 -- [l][0] = kernel entry
                                    // (l = 0) or the moment layer l-1 was published, [l][1] = layer l published
};

struct DenseBlockSmem {
  static constexpr int kBarBytes = 1024;
  static constexpr int kMidBytes = 2 * 128 * 4;   // BN2 shift of the current and the next layer
  int a_off, b_off, t_off, mid_off, tail_off, total;
};

__host__ __device__ inline DenseBlockSmem dense_block_smem(const DenseBlockParams& p) {
  DenseBlockSmem L;
  L.mid_off = DenseBlockSmem::kBarBytes;
  L.t_off = L.mid_off + DenseBlockSmem::kMidBytes;
  L.tail_off = L.t_off + dl_t_bytes(8) / 2;     // ONE bottleneck tile (2 chunks): see the note at `tbuf`
  L.a_off = L.tail_off + 2 * dl_a_stage(8);
  L.b_off = L.a_off + p.a_stages * dl_a_stage(8);
  L.total = L.b_off + p.b_stages * kDlBStage + 1024;
  return L;
}

// arrive (release, cluster scope) on the mbarrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)), "r"(rank)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// 16-byte store into the shared memory of CTA `rank` of this cluster, at the address `local` has in this CTA
__device__ __forceinline__ void st_cluster_v4(const void* local, uint32_t rank, uint32_t a, uint32_t b, uint32_t c,
                                              uint32_t d) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.v4.b32 [ra], {%2, %3, %4, %5};\n\t"
      "}\n" ::"r"(smem_u32(local)), "r"(rank), "r"(a), "r"(b), "r"(c), "r"(d)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

__global__ void __launch_bounds__(640, 1)
dense_block_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ DenseBlockParams p) {
  constexpr int RH = 8;
  constexpr int kRows = dl_rows(RH), kAStage = dl_a_stage(RH);
  constexpr int kDlTChunk = dl_t_chunk(RH), kDlTBuf = 2 * kDlTChunk;
  static_assert(kRows <= 128, "one M block per region");
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
  uint64_t* t_ready = acc1_empty + 1;    // [2 buffers][2 channel chunks]
  uint64_t* t_empty = t_ready + 4;       // [2]
  uint64_t* acc2_full = t_empty + 2;     // [2]
  uint64_t* acc2_empty = acc2_full + 2;  // [2]
  uint64_t* nb_bar = acc2_empty + 2;     // [2] "every region of my image has published layer l" (l & 1)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nb_bar + 2);

  const DenseBlockSmem L = dense_block_smem(p);
  uint8_t* a_base = smem + L.a_off;
  uint8_t* b_base = smem + L.b_off;
  uint8_t* t_base = smem + L.t_off;
  uint8_t* tail_base = smem + L.tail_off;      // [2][kAStage] raw fp16 tail tiles (see header)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 2 * kAStage / 16; i += blockDim.x) reinterpret_cast<uint4*>(tail_base)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0 && p.gt_layers && blockIdx.x == 0) p.gt_layers[0] = globaltimer_ns();

  if (warp == 0 && lane == 0) tma_prefetch_desc(&map_x);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 8);      // per-warp arrivals, 8 transform warps
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    mbar_init(acc1_full, 1);
    mbar_init(acc1_empty, 8);         // 8 epilogue warps
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_ready[2 * i], 8);
      mbar_init(&t_ready[2 * i + 1], 8);
      mbar_init(&t_empty[i], 1);
      mbar_init(&acc2_full[i], 1);
      mbar_init(&acc2_empty[i], 8);
      mbar_init(&nb_bar[i], static_cast<uint32_t>(p.cluster_size));
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.cluster_size > 1) cluster_sync_all();   // every CTA's barriers are initialised before any remote arrive
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc1_col = tmem_base;        // 128 columns
  const uint32_t acc2_col = tmem_base + 256;  // 2 x 32 columns

  // this CTA's region, the same for every layer
  const int item = blockIdx.x;
  const int tw = item % p.tiles_w;
  const int trow = (item / p.tiles_w) % p.tiles_h;
  const int n0 = item / (p.tiles_w * p.tiles_h);
  const int w0 = tw * 8, h0 = trow * RH;
  const int NL = p.n_layers;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer: activation halo chunks
    uint32_t sa = 0, pa = 0;
    DlTrace tc0;
    if (lane == 0) tc0 = dl_trace_open(p.trace, 0);
    for (int l = 0; l < NL; ++l) {
      const int C = p.layers[l].C, n_chunks = p.layers[l].n_chunks;
      const int n_safe = (l == 0) ? 0 : (C - 32) / 64;     // chunks made only of channels older than layer l-1's output
      if (l == NL - 1 && lane == 0) pdl_launch_dependents();
      for (int c = 0; c < n_chunks; ++c) {
        if (l == 0 && c == 0 && lane == 0) pdl_wait();      // the block's input comes from the preceding kernel
        if (l >= 2 && c == (n_safe > 0 ? n_safe - 1 : 0) && lane == 0) {
          // the last safe chunk may hold layer l-2's channels: they must be visible in global memory (all CTAs of
          // the image published layer l-2 a whole layer ago: this wait is off the critical path)
          uint32_t spins = 0;
          while (!mbar_try_wait_acquire_cluster(&nb_bar[(l - 2) & 1], ((l - 2) >> 1) & 1)) {
            if (++spins > (1u << 26)) {
              printf("dp: dense block barrier timeout (producer) block %d layer %d\n", blockIdx.x, l);
              __trap();
            }
          }
          fence_proxy_async_all();                // the acquired generic-proxy writes are read by TMA (async proxy)
        }
        if (lane == 0) {
          mbar_wait(&a_empty[sa], pa ^ 1);
          if (l > 0 && c == n_safe) {
            mbar_arrive(&a_full[sa]);             // slot is free; its data comes from the tail tile (transform warps)
            dl_trace_ev(tc0, 2, l);
          } else {
            mbar_expect_tx(&a_full[sa], kRows * 128);
            tma_load_4d(&map_x, &a_full[sa], a_base + sa * kAStage, c * 64, w0 - 1, h0 - 1, n0);
            dl_trace_ev(tc0, 1, l);
          }
        }
        if (++sa == static_cast<uint32_t>(p.a_stages)) { sa = 0; pa ^= 1; }
        __syncwarp();
      }
    }
    if (lane == 0) dl_trace_close(p.trace, tc0, 0);
  } else if (warp == 3) {
    // ------------------------------------------------------------------ producer: weights, in MMA order
    if (elect_one()) {
      uint32_t sb = 0, pb = 0;
      for (int l = 0; l < NL; ++l) {
        const DenseBlockLayer* Lr = p.layers + l;
        const int n_chunks = Lr->n_chunks;
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&b_empty[sb], pb ^ 1);
          mbar_expect_tx(&b_full[sb], 128 * 128);
          tma_load_3d(&Lr->map_w1, &b_full[sb], b_base + sb * kDlBStage, c * 64, 0, 0);
          if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
        }
        for (int c = 0; c < 2; ++c)
          for (int g = 0; g < 9 / kDlW2Group; ++g) {
            mbar_wait(&b_empty[sb], pb ^ 1);
            mbar_expect_tx(&b_full[sb], kDlW2Group * 32 * 128);
            tma_load_3d(&Lr->map_w2, &b_full[sb], b_base + sb * kDlBStage, c * 64, 0, g * kDlW2Group);
            if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc_f16(128), idesc2 = make_idesc_f16(32);
      const uint64_t hi_dense = static_cast<uint64_t>(sw128_desc_hi(1024)) << 32;
      const uint64_t hi_halo = static_cast<uint64_t>(sw128_desc_hi(kDlHaloW * 128)) << 32;
      const uint64_t a_desc0 = hi_dense | sw128_desc_lo(smem_u32(a_base));
      const uint64_t b_desc0 = hi_dense | sw128_desc_lo(smem_u32(b_base));
      const uint64_t t_desc0 = hi_halo | sw128_desc_lo(smem_u32(t_base));
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      DlTrace tc = dl_trace_open(p.trace, 1);
      auto ph1 = [&](int l) {
        const int C = p.layers[l].C, n_chunks = p.layers[l].n_chunks;
        mbar_wait(acc1_empty, (l & 1) ^ 1);
        tc_fence_after();
        for (int c = 0; c < n_chunks; ++c) {
          int ks = (C - c * 64 + 15) >> 4;
          ks = ks > 4 ? 4 : ks;
          mbar_wait(&a_ready[sa], pa);
          dl_trace_ev(tc, 1, l);
          mbar_wait(&b_full[sb], pb);
          tc_fence_after();
          const uint64_t a_desc = a_desc0 + sa * (kAStage >> 4);
          const uint64_t b_desc = b_desc0 + sb * (kDlBStage >> 4);
          const uint32_t acc = (c > 0) ? 1u : 0u;
          if (ks == 4) {
            umma_f16_ss_k4(acc1_col, a_desc, b_desc, idesc1, acc);
          } else if (ks == 2) {
            umma_f16_ss_k2(acc1_col, a_desc, b_desc, idesc1, acc);
          } else {
            for (int kk = 0; kk < ks; ++kk) umma_f16_ss(acc1_col, a_desc + 2 * kk, b_desc + 2 * kk, idesc1, kk ? 1u : acc);
          }
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
          if (++sa == static_cast<uint32_t>(p.a_stages)) { sa = 0; pa ^= 1; }
          if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
        }
        umma_commit(acc1_full);
        dl_trace_ev(tc, 3, l);
      };
      auto ph2 = [&](int l) {
        const int tb = l & 1;
        const uint32_t u = (l >> 1) & 1;
        mbar_wait(&acc2_empty[tb], u ^ 1);
        const uint32_t d2 = acc2_col + tb * 32;
        for (int c = 0; c < 2; ++c) {
          mbar_wait(&t_ready[2 * tb + c], u);     // the 64-channel half of T this pass reads (mid publishes them one by one)
          tc_fence_after();
          if (c == 0) dl_trace_ev(tc, 4, l);
          const uint64_t t_desc = t_desc0 + ((c * kDlTChunk) >> 4);
          for (int g = 0; g < 9 / kDlW2Group; ++g) {
            mbar_wait(&b_full[sb], pb);
            tc_fence_after();
            uint64_t b_desc = b_desc0 + sb * (kDlBStage >> 4);
#pragma unroll
            for (int j = 0; j < kDlW2Group; ++j, b_desc += (32 * 128) >> 4) {
              const int tap = g * kDlW2Group + j;
              const int dy = tap / 3, dx = tap - dy * 3;
              umma_f16_ss_k4(d2, t_desc + (dy * kDlHaloW + dx) * 8, b_desc, idesc2, (c | tap) ? 1u : 0u);
            }
            umma_commit(&b_empty[sb]);
            if (++sb == static_cast<uint32_t>(p.b_stages)) { sb = 0; pb ^= 1; }
          }
        }
        umma_commit(&t_empty[tb]);
        umma_commit(&acc2_full[tb]);
        dl_trace_ev(tc, 5, l);
      };
      ph1(0);
      for (int l = 0; l < NL; ++l) {
        ph2(l);
        if (l + 1 < NL) ph1(l + 1);
      }
      dl_trace_close(p.trace, tc, 1);
    }
  } else if ((warp >= 4 && warp < 8) || warp >= 16) {
    // ------------------------------------------------------------------ epilogue warps: mid + final + publish
    const int q = warp & 3;
    const int half = (warp >= 16) ? 1 : 0;
    const int r = q * 32 + lane;
    DlTrace tc;
    if (r == 0 && half == 0) tc = dl_trace_open(p.trace, 2);
    // BN2 shift vectors are staged in shared memory one layer ahead: buffer (l & 1) is rewritten for layer l + 2 only
    // after the publish barrier of layer l, i.e. after every epilogue warp has finished mid(l)
    float* s_mid = reinterpret_cast<float*>(smem + L.mid_off);
    if (half == 0) {
      s_mid[r] = p.layers[0].mid_shift[r];
      if (NL > 1) s_mid[128 + r] = p.layers[1].mid_shift[r];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int l = 0; l < NL; ++l) {
      const DenseBlockLayer* Lr = p.layers + l;
      const float* mid_shift = s_mid + (l & 1) * 128;
      const int out_choff = Lr->out_choff;
      const int tb = l & 1;
      const uint32_t u = (l >> 1) & 1;
      // ---- mid(l): acc1 -> +shift -> ReLU -> zero padding -> fp16 -> swizzled operand tile T[l & 1]
      mbar_wait(acc1_full, l & 1);
      mbar_wait(&t_empty[tb], u ^ 1);   // the 3x3 MMAs of layer l-2 no longer read this buffer
      tc_fence_after();
      dl_trace_ev(tc, 0, l);
      {
        // One T tile is enough here: mid(l+1) starts on acc1_full(l+1), a tcgen05.commit issued after ph2(l)'s MMAs by
        // the same thread, so every 3x3 read of layer l's tile has completed before layer l+1's tile is written.
        uint8_t* tbuf = t_base;
        // step s covers bottleneck channels (s >> 1) * 64 + half * 32 + (s & 1) * 16 .. + 15: both warp groups finish
        // T chunk 0 (channels 0-63) after two steps, so the 3x3 MMAs on chunk 0 overlap the conversion of chunk 1
        const uint32_t lane_col = acc1_col + half * 32 + (static_cast<uint32_t>(q * 32) << 16);
        const int prow = r;
        const int hh = prow / kDlHaloW, ww = prow - hh * kDlHaloW;
        const int ih = h0 - 1 + hh, iw = w0 - 1 + ww;
        const bool inside = ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
        auto process = [&](int sidx, const uint32_t (&v)[16]) {
          const int cb = (sidx >> 1) * 64 + half * 32 + (sidx & 1) * 16;
          if (prow >= kRows) return;
          float f[16];
          epi_affine16(v, nullptr, mid_shift + cb, false, true, f);
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            __half2 h2 = inside ? __floats2half2_rn(f[2 * i], f[2 * i + 1]) : __float2half2_rn(0.f);
            pk[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
          uint8_t* row = tbuf + (cb >> 6) * kDlTChunk + prow * 128;
          const int j = (cb & 63) >> 3;
          *reinterpret_cast<uint4*>(row + (((j) ^ (prow & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(row + (((j + 1) ^ (prow & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        };
        uint32_t v0[16], v1[16];
        tmem_ld16(lane_col, v0);
        tmem_ld_wait();
        tmem_ld16(lane_col + 16, v1);
        process(0, v0);
        tmem_ld_wait();
        tmem_ld16(lane_col + 64, v0);
        process(1, v1);
        fence_proxy_async_smem();
        mbar_arrive_warp(&t_ready[2 * tb]);
        tmem_ld_wait();
        tmem_ld16(lane_col + 80, v1);
        process(2, v0);
        tmem_ld_wait();
        process(3, v1);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive_warp(acc1_empty);
      mbar_arrive_warp(&t_ready[2 * tb + 1]);
      dl_trace_ev(tc, 2, l);
      // ---- fin(l): acc2[l & 1] -> fp16 -> the 32 new channels of the concat buffer
      mbar_wait(&acc2_full[tb], u);
      tc_fence_after();
      dl_trace_ev(tc, 3, l);
      {
        const int w = w0 + (r & 7), h = h0 + (r >> 3);
        const bool valid = ((r >> 3) < RH) && (h < p.H) && (w < p.W);
        const long long opix = (static_cast<long long>(n0) * p.H + h) * p.W + w;
        __half* orow = p.out + opix * p.out_ctot + out_choff;
        uint32_t v[16];
        tmem_ld16(acc2_col + tb * 32 + half * 16 + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          __half2 h2 = __floats2half2_rn(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
          pk[i] = *reinterpret_cast<uint32_t*>(&h2);
        }
        if (valid) {
          st_global_v8(orow + 16 * half, pk);
          // the same 16 channels into the tail tiles: chunk C / 64, 16-byte chunks (C % 64) / 8 + 2 half, +1
          const int C = Lr->C;
          uint8_t* tail = tail_base + ((C >> 6) & 1) * kAStage;
          const int j = ((C & 63) >> 3) + 2 * half;
          const int y = r >> 3, x = r & 7;
          {
            const int row = (y + 1) * kDlHaloW + (x + 1);
            *reinterpret_cast<uint4*>(tail + row * 128 + ((j ^ (row & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(tail + row * 128 + (((j + 1) ^ (row & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
          if (p.cluster_size > 1) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
              for (int dx = -1; dx <= 1; ++dx) {
                if (dy == 0 && dx == 0) continue;
                if ((dy < 0 && y != 0) || (dy > 0 && y != RH - 1) || (dx < 0 && x != 0) || (dx > 0 && x != 7)) continue;
                const int r2 = trow + dy, c2 = tw + dx;
                if (r2 < 0 || r2 >= p.tiles_h || c2 < 0 || c2 >= p.tiles_w) continue;
                const uint32_t peer = static_cast<uint32_t>(r2 * p.tiles_w + c2);
                const int row = (y + 1 - RH * dy) * kDlHaloW + (x + 1 - 8 * dx);     // my pixel in the peer's halo frame
                st_cluster_v4(tail + row * 128 + ((j ^ (row & 7)) << 4), peer, pk[0], pk[1], pk[2], pk[3]);
                st_cluster_v4(tail + row * 128 + (((j + 1) ^ (row & 7)) << 4), peer, pk[4], pk[5], pk[6], pk[7]);
              }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_warp(&acc2_empty[tb]);
      dl_trace_ev(tc, 1, l);
      // ---- publish: CTA barrier over the epilogue warps, then one release-arrive per cluster CTA (thread r -> CTA r):
      //      the barrier orders every warp's global, local and remote stores before the arrive, the release makes them
      //      visible at cluster scope (the pattern of a cooperative-groups grid sync)
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (r < p.cluster_size && half == 0) mbar_arrive_remote_release(&nb_bar[l & 1], static_cast<uint32_t>(r));
      if (r == 0 && half == 0) {
        dl_trace_ev(tc, 4, l);
        if (p.gt_layers && blockIdx.x == 0) {
          const unsigned long long now = globaltimer_ns();
          p.gt_layers[2 * l + 1] = now;
          if (l + 1 < NL) p.gt_layers[2 * l + 2] = now;
        }
      }
      if (half == 0 && l + 2 < NL) s_mid[(l & 1) * 128 + r] = p.layers[l + 2].mid_shift[r];   // off the critical path
    }
    if (r == 0 && half == 0) dl_trace_close(p.trace, tc, 2);
  } else if (warp >= 8) {
    // ------------------------------------------------------------------ pre-activation BN + ReLU on the halo rows
    const int t = tid - 256;
    const __half2 zero2 = __float2half2_rn(0.f);
    uint32_t sa = 0, pa = 0;
    DlTrace tc;
    if (t == 0) tc = dl_trace_open(p.trace, 3);
    for (int l = 0; l < NL; ++l) {
      const DenseBlockLayer* Lr = p.layers + l;
      const float* pro_scale = Lr->pro_scale;
      const float* pro_shift = Lr->pro_shift;
      const int n_chunks = Lr->n_chunks;
      // this layer's BN vectors into L1 (2 x n_chunks x 256 B = 4 n_chunks lines of 128 B)
      for (int i = t; i < 2 * n_chunks; i += 256) {
        prefetch_l1(pro_scale + 32 * i);
        prefetch_l1(pro_shift + 32 * i);
      }
      for (int c = 0; c < n_chunks; ++c) {
        const int i = t & 7;
        const int ch = c * 64 + i * 8;
        const float4 sc0 = __ldg(reinterpret_cast<const float4*>(pro_scale + ch));
        const float4 sc1 = __ldg(reinterpret_cast<const float4*>(pro_scale + ch + 4));
        const float4 sh0 = __ldg(reinterpret_cast<const float4*>(pro_shift + ch));
        const float4 sh1 = __ldg(reinterpret_cast<const float4*>(pro_shift + ch + 4));
        mbar_wait(&a_full[sa], pa);
        uint8_t* stage_base = a_base + sa * kAStage;
        const uint8_t* src_base = stage_base;
        if (l > 0 && c == (Lr->C - 32) / 64) {
          // the chunk with layer l-1's channels: raw data sits in the tail tile once every CTA of the image has
          // published layer l-1 (their border pixels are written straight into this CTA's tile)
          uint32_t spins = 0;
          while (!mbar_try_wait_acquire_cluster(&nb_bar[(l - 1) & 1], ((l - 1) >> 1) & 1)) {
            if (++spins > (1u << 26)) {
              printf("dp: dense block barrier timeout (transform) block %d layer %d\n", blockIdx.x, l);
              __trap();
            }
          }
          src_base = tail_base + (c & 1) * kAStage;
        }
        dl_trace_ev(tc, 1, l);
        {
          constexpr int kIter = (kRows + 31) / 32;
          uint4 raw[kIter];
#pragma unroll
          for (int j = 0; j < kIter; ++j) {
            const int rr = (t >> 3) + 32 * j;
            if (rr < kRows) raw[j] = *reinterpret_cast<const uint4*>(src_base + rr * 128 + ((i ^ (rr & 7)) << 4));
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
        fence_proxy_async_smem();
        mbar_arrive_warp(&a_ready[sa]);
        dl_trace_ev(tc, 0, l);
        if (++sa == static_cast<uint32_t>(p.a_stages)) { sa = 0; pa ^= 1; }
      }
    }
    if (t == 0) dl_trace_close(p.trace, tc, 3);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.cluster_size > 1) cluster_sync_all();   // no CTA leaves while a peer may still arrive on its barriers
}

}  // namespace dp
