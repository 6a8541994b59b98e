// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05.mma, accumulators in TMEM), NHWC fp16
// activations, fp32 accumulate.  One kernel covers every conv of the DenseNet-121 U-Net
// (reference graph: DigiPathAI/models/densenet.py:37-159):
//
//   GEMM view   D[pixels, Cout] = sum over (tap entry e, channel chunk c)  A_e,c[pixels, 64] * W_e,c[Cout, 64]^T
//
//   MODE_D  1x1 conv: A tiles are flat [128 pixels x 64 ch] boxes of a 2-D [pixels, C] view.
//   MODE_T  3x3 conv on maps too small for a halo region (8x8): one shifted 4-D TMA box per tap;
//           out-of-bounds pixels are zero-filled by TMA == Keras padding='same'.
//   MODE_H  3x3 conv with a shared-memory resident halo: ONE 4-D TMA box [(16+2) x (8*SUB+2) pixels x 64 ch]
//           per channel chunk; the 9 taps are 9 UMMA descriptors into that box at different row offsets
//           (an 8-pixel row segment is one 8-row swizzle group; consecutive image rows are SBO apart).
//           This is what keeps the L2->SMEM traffic at ~1.3x the activation bytes instead of 9x.
//   up2     UpSampling2D()+conv3x3 (densenet.py:138-155) is evaluated as 4 sub-pixel phases, each a 2x2 conv
//           on the LOW-resolution map with pre-summed weights (2.25x fewer MACs, no upsampled tensor in HBM).
//
// Warp roles (1 CTA / SM, persistent over work items):
//   warp 0 : TMA producer, activations (A ring)        warp 1 : tcgen05.mma issuer (one elected lane each)
//   warp 2 : TMEM allocator                             warp 3 : TMA producer, weights (B ring)
//   warps 4-7 : epilogue (tcgen05.ld -> BN shift/ReLU -> staged, coalesced stores to HBM)
//   warps 8-11    : (PROLOGUE only) pre-activation BN+ReLU applied to the A tile in shared memory
//                   (dense-layer `_0_bn/_0_relu`, densenet.py:59-63) before the MMA reads it.
#pragma once
#include "ptx.cuh"
#include "d4.cuh"

namespace dp {

enum : int { MODE_D = 0, MODE_T = 1, MODE_H = 2 };
enum : int { EPI_STORE = 0, EPI_HEAD = 1 };

constexpr int kMaxEntries = 32;   // 5x5 taps; tap_first_mask is 32 bits
constexpr int kMaxAStages = 8;
constexpr int kMaxBStages = 16;
constexpr int kTmemCols = 512;
constexpr int kATileBytes = 128 * 128;  // 128 pixel rows x 64 fp16
constexpr int kEpiRowBytes = 144;        // 64 fp16 + 16 B pad: conflict-free 16-byte row writes
constexpr int kEpiStageBytes = 32 * kEpiRowBytes;  // per epilogue warp

struct TapEntry {
  int8_t dy, dx;   // input offset relative to the (low-res) output pixel
  int8_t group;    // accumulator group (sub-pixel phase in MODE_H up2, else 0)
  int8_t pad;
};

struct ConvParams {
  int mode, sub, n_tile, n_ntiles;
  int n_img, H, W, Cin;  // input grid per image and channels visible through the A tensor map
  int OH, OW;            // grid the work items tile: the output grid (== input grid for stride 1; for up2 the
                         // low-resolution grid, the output being 2x that)
  int stride;            // MODE_T only: input pixel = stride * output pixel + tap offset (TMA element strides)
  int cout;              // true Cout; n_tile * n_ntiles may exceed it (weight rows beyond it are TMA zero fill,
                         // the epilogue clips the store)
  int residual;          // out = act(out_old + acc + shift): Inception-ResNet block tail (inception.py:152-160)
  int halo_left;         // MODE_H: columns of halo left of the region
  int n_chunks;          // ceil(Cin / 64)
  int n_entries;         // tap entries per work item
  int n_groups;          // accumulator groups per work item
  int n_phase_items;     // MODE_T up2: 4 (phase is part of the work item), else 1
  int up2;               // output grid is (2H, 2W)
  int box_w, box_h, box_n;  // MODE_T: TMA box; MODE_H: halo box (WW, HH, 1)
  int tiles_w, tiles_h;     // tile grid per image (MODE_T: per box_n images)
  int m_total;              // MODE_D: total pixels
  int n_mtiles, n_items;
  int a_stage_bytes, b_stage_bytes, a_stages, b_stages, acc_stages, a_tx_bytes;
  int b_group;              // tap entries carried by one B stage (TMA box depth)
  int halo_top;             // MODE_H: rows of halo above the region (1 for 3x3 / up2, 2 for the 4x4 stem)
  int b_pair;               // 1: launched as 2-CTA clusters; each CTA fetches half of every weight tile and TMA
                            //    multicasts it to both (halves the L2->SM weight traffic that bounds the
                            //    high-resolution decoder layers)
  int b_resident;           // 1: all weight tiles of the layer are loaded once per CTA and stay in shared memory
                            //    (stage c holds every tap of channel chunk c); no re-streaming from L2 per item
  int epi_direct;           // 1: epilogue stores 32-byte vectors straight from registers (no smem staging)
  int epi_mode, relu;
  int out_ctot, out_choff;  // output NHWC buffer: channel stride and channel offset (elements)
  int desc_base_mode;       // 0: descriptor base_offset field = 0; 1: (addr >> 7) & 7
  int phase_fixed;          // MODE_H up2: item = m_tile * n_phase_items + phase group, n_groups phases per item,
                            // the CTA's phase group (blockIdx % n_phase_items) never changes -> resident weights
  int fast_id;              // MODE_H: 0 = generic tap loop, else (kind, sub, taps per B stage) of a fully unrolled
                            // issue sequence: fast_id = kind * 100 + sub * 10 + b_group (see issue_chunk_h)
  int dbg_skip;             // TIMING EXPERIMENTS ONLY (wrong results): bit 0 = weight tiles are fetched only on the
                            // first pass through the B ring, bit 1 = same for activation tiles, bit 2 = epilogue
                            // does not store, bit 3 = epilogue does nothing but release the accumulator
  int pro_relu;
  int img0, P;              // EPI_HEAD: first image of this sub-batch within the call; tile side
  float head_b;
  TapEntry entries[kMaxEntries];
  // host-precomputed per tap entry (kernel parameters live in the constant bank, so the issue loop reads these
  // straight into uniform registers): A-descriptor offset in 16-byte units, accumulator column offset, and a
  // bitmask of entries that are the first of their accumulator group
  uint32_t tap_a[kMaxEntries];
  uint32_t tap_d[kMaxEntries];
  uint32_t tap_first_mask;
  const float* epi_scale;
  const float* epi_shift;
  const float* pro_scale;
  const float* pro_shift;
  __half* out;
  const float* head_w;
  const PassDesc* pass;     // EPI_HEAD: tta_out + probability output of the current call
  unsigned long long* gt;     // debug: [0]/[1] receive %globaltimer at entry / exit of CTA 0
  unsigned long long* trace;  // debug: [0] = entry counter, then (role<<56 | event<<48 | item<<32 | clock32)
};

struct ConvSmemLayout {
  // offsets from the 1024-aligned base
  static constexpr int kBarBytes = 1024;
  int a_off, b_off, epi_off, pro_off, head_off, stage_off, total;
};

__host__ __device__ inline ConvSmemLayout conv_smem_layout(const ConvParams& p) {
  ConvSmemLayout L;
  L.a_off = ConvSmemLayout::kBarBytes;
  L.b_off = L.a_off + p.a_stages * p.a_stage_bytes;
  L.epi_off = L.b_off + p.b_stages * p.b_stage_bytes;
  int cout = p.n_tile * p.n_ntiles;
  L.pro_off = L.epi_off + 2 * cout * 4;
  L.head_off = L.pro_off + 2 * p.n_chunks * 64 * 4;
  L.stage_off = L.head_off + 256 * 4;
  // per-warp staging rows exist only for the staged (non-direct, non-head) store path
  const bool staged = !(p.epi_direct || p.epi_mode == EPI_HEAD);
  L.total = L.stage_off + (staged ? 4 * kEpiStageBytes : 0) + 1024;  // + alignment slack
  return L;
}

struct WorkItem {
  int nt, ph;
  int n0, h0, w0;  // T/H: tile origin on the input grid
  int m0;          // D: first pixel
};

__device__ __forceinline__ WorkItem decode_item(const ConvParams& p, int item) {
  WorkItem wi;
  int mt = item % p.n_mtiles;
  int rest = item / p.n_mtiles;
  wi.nt = rest % p.n_ntiles;
  wi.ph = rest / p.n_ntiles;
  if (p.phase_fixed) {
    wi.ph = item % p.n_phase_items;
    mt = item / p.n_phase_items;
    wi.nt = 0;
  }
  wi.m0 = 0; wi.n0 = 0; wi.h0 = 0; wi.w0 = 0;
  if (p.mode == MODE_D) {
    wi.m0 = mt * p.sub * 128;
  } else if (p.mode == MODE_T) {
    int tw = mt % p.tiles_w;
    int r = mt / p.tiles_w;
    int th = r % p.tiles_h;
    int tn = r / p.tiles_h;
    wi.w0 = tw * p.box_w; wi.h0 = th * p.box_h; wi.n0 = tn * p.box_n;
  } else {
    int tw = mt % p.tiles_w;
    int r = mt / p.tiles_w;
    int th = r % p.tiles_h;
    wi.n0 = r / p.tiles_h;
    wi.w0 = tw * 8 * p.sub; wi.h0 = th * 16;
  }
  return wi;
}

// Epilogue arithmetic for 16 accumulator columns: y = acc [* scale] + shift, optional ReLU.  BatchNorm scales
// are normally folded into the fp16 weights by the packer (program.py), leaving one vectorised shift load per
// four channels; the scale path remains for containers that keep it separate.
__device__ __forceinline__ void epi_affine16(const uint32_t (&v)[16], const float* sc, const float* sh, bool has_scale,
                                             bool relu, float (&f)[16]) {
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    const float4 b = *reinterpret_cast<const float4*>(sh + 4 * i4);
    float x0 = __uint_as_float(v[4 * i4]), x1 = __uint_as_float(v[4 * i4 + 1]);
    float x2 = __uint_as_float(v[4 * i4 + 2]), x3 = __uint_as_float(v[4 * i4 + 3]);
    if (has_scale) {
      const float4 a = *reinterpret_cast<const float4*>(sc + 4 * i4);
      x0 = fmaf(x0, a.x, b.x); x1 = fmaf(x1, a.y, b.y); x2 = fmaf(x2, a.z, b.z); x3 = fmaf(x3, a.w, b.w);
    } else {
      x0 += b.x; x1 += b.y; x2 += b.z; x3 += b.w;
    }
    if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); x2 = fmaxf(x2, 0.f); x3 = fmaxf(x3, 0.f); }
    f[4 * i4] = x0; f[4 * i4 + 1] = x1; f[4 * i4 + 2] = x2; f[4 * i4 + 3] = x3;
  }
}

// Debug timeline: each role appends to its own 2000-entry region (no atomics, fire-and-forget stores), so the
// probe costs a few clocks.  Entry = event<<48 | item<<32 | clock32; region r starts at trace[8 + 2000 r];
// trace[r] receives the final entry count of role r.
struct TraceCursor {
  unsigned long long* base = nullptr;
  unsigned n = 0;
};
__device__ __forceinline__ TraceCursor trace_open(const ConvParams& p, unsigned role) {
  TraceCursor c;
  if (p.trace && blockIdx.x == 0) c.base = p.trace + 8 + 2000 * role;
  return c;
}
__device__ __forceinline__ void trace_ev(TraceCursor& c, unsigned ev, unsigned item) {
  if (c.base && c.n < 2000) {
    c.base[c.n++] = (static_cast<unsigned long long>(ev) << 48) | (static_cast<unsigned long long>(item & 0xFFFF) << 32) |
                    (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFull);
  }
}
__device__ __forceinline__ void trace_close(const ConvParams& p, const TraceCursor& c, unsigned role) {
  if (c.base) p.trace[role] = c.n;
}

// ---------------------------------------------------------------------------------------------------------
// Fully unrolled MMA issue of one 64-channel chunk in halo mode.  The generic tap loop spends ~250 clocks of
// scalar work per tap in the single issuing thread (indexed constant loads of the tap table, R2UR moves, 64-bit
// descriptor adds, loop control) -- more than the 192 clocks four N = 64 MMAs take to execute -- so the
// high-resolution layers were issue bound, not tensor bound (measured with the tap table stubbed out, TMA and
// epilogue disabled: same time).  Here every tap offset, accumulator group and first-use flag is a compile-time
// constant: KIND 3 = 3x3 taps (dy, dx in -1..1), KIND 4 = up2 sub-pixel entries [phase][2x2 tap]; SUB sub-tiles
// of 8 pixels (box width 8 SUB + 2); BG taps per weight stage (BG == all taps when the weights are resident).
struct IssueRing {
  uint64_t* b_full;
  uint64_t* b_empty;
  uint32_t sb, pb;
};

// E0 / NE select a sub-range of the up2 entries (phase-split items: NE = 4 G entries starting at phase group
// E0 / NE; the resident weight stage then starts at entry E0).
template <int KS, int KIND, int SUB, int BG, int E0 = 0, int NE = ((KIND == 3) ? 9 : 16)>
__device__ __forceinline__ void issue_chunk_h(const ConvParams& p, IssueRing& r, uint64_t a_stage_desc, uint64_t b_desc0,
                                              uint32_t b_stage_u, uint32_t b_ent_u, uint32_t d_stage, uint32_t n_tile,
                                              uint32_t idesc, uint32_t acc_c, int c, bool b_loaded, TraceCursor& tc,
                                              int item) {
  constexpr int WW = 8 * SUB + 2;
  constexpr int G = (KIND == 4) ? ((NE >= 16) ? 4 : NE / 4) : 1;   // accumulator groups (phases per item)
  static_assert(NE % BG == 0, "taps per weight stage must divide the tap count");
#pragma unroll
  for (int g = 0; g < NE / BG; ++g) {
    if (p.b_resident) {
      r.sb = c;
      if (!b_loaded) mbar_wait(&r.b_full[r.sb], 0);
    } else {
      mbar_wait(&r.b_full[r.sb], r.pb);
    }
    // no tcgen05.fence here: the operands were written by TMA and published through the mbarrier's transaction
    // count; the fence after the accumulator hand-off (acc_empty) is the one that orders against other threads'
    // tcgen05 operations
    trace_ev(tc, 2, item);
    const uint64_t b_desc = b_desc0 + r.sb * b_stage_u;
#pragma unroll
    for (int j = 0; j < BG; ++j) {
      const int e = E0 + g * BG + j;
      int dy, dx, grp;
      if (KIND == 3) {
        dy = e / 3 - 1; dx = e % 3 - 1; grp = 0;
      } else {
        const int ph = e >> 2, t = e & 3;
        dy = (ph >> 1) - 1 + (t >> 1); dx = (ph & 1) - 1 + (t & 1); grp = ph % G;
      }
      const uint32_t a_off = static_cast<uint32_t>(((dy + 1) * WW + dx + 1) * 8);
      const bool first = (KIND == 3) ? (e == 0) : ((e & 3) == 0);
      const uint32_t flag = first ? acc_c : 1u;
#pragma unroll
      for (int s = 0; s < SUB; ++s) {
        if (KS == 4)
          umma_f16_ss_k4(d_stage + static_cast<uint32_t>(grp * SUB + s) * n_tile, a_stage_desc + a_off + s * 64u,
                         b_desc + static_cast<uint32_t>(j) * b_ent_u, idesc, flag);
        else
          umma_f16_ss_k2(d_stage + static_cast<uint32_t>(grp * SUB + s) * n_tile, a_stage_desc + a_off + s * 64u,
                         b_desc + static_cast<uint32_t>(j) * b_ent_u, idesc, flag);
      }
      trace_ev(tc, 4, item);
    }
    if (!p.b_resident) {
      if (p.b_pair) umma_commit_mcast(&r.b_empty[r.sb], 3);
      else umma_commit(&r.b_empty[r.sb]);
      trace_ev(tc, 5, item);
      if (++r.sb == static_cast<uint32_t>(p.b_stages)) { r.sb = 0; r.pb ^= 1; }
    }
  }
}

// Epilogue warp groups of a kernel variant: the plain variants run TWO groups of four warps (warps 4-7 and 8-11; a
// warp can only read TMEM lanes 32 (w % 4) .. +31, so the groups split an item's (accumulator, 32-column step)
// units between them).  With one epilogue warp per SM sub-partition the epilogue is issue-latency bound
// (~0.16 IPC per warp, tensor pipe 25-45 % busy on the high-resolution decoder layers); a second warp per
// sub-partition hides that latency.  The PROLOGUE variant uses warps 8-11 for the pre-activation transform and
// the RESIDUAL variant holds a 128-register prefetch per thread: both keep one group.
#ifndef DP_EPI_GROUPS
#define DP_EPI_GROUPS 2
#endif
template <bool PROLOGUE, bool RESIDUAL>
struct ConvKernelShape {
  static constexpr int kEpiGroups = (PROLOGUE || RESIDUAL) ? 1 : DP_EPI_GROUPS;
  static constexpr int kThreads = (PROLOGUE || kEpiGroups == 2) ? 384 : 256;
};

// FAST != 0 compiles exactly one unrolled issue sequence into the kernel (issue_chunk_h): FAST = kind * 100 +
// sub * 10 + taps-per-weight-stage, or 4000 + 100 G + 10 sub for phase-split up-convs.  One kernel per variant
// keeps each of them lean: inlining all variants into one kernel cost registers (spills) and instruction-cache
// footprint and slowed EVERY layer down by 5-10 %.
template <int MODE, bool PROLOGUE, bool RESIDUAL = false, int FAST = 0>
__global__ void __launch_bounds__((ConvKernelShape<PROLOGUE, RESIDUAL>::kThreads), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ ConvParams p) {
  constexpr int kEpiGroups = ConvKernelShape<PROLOGUE, RESIDUAL>::kEpiGroups;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* a_ready = a_full + kMaxAStages;
  uint64_t* a_empty = a_ready + kMaxAStages;
  uint64_t* b_full = a_empty + kMaxAStages;
  uint64_t* b_empty = b_full + kMaxBStages;
  uint64_t* acc_full = b_empty + kMaxBStages;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const ConvSmemLayout L = conv_smem_layout(p);
  uint8_t* a_base = smem + L.a_off;
  uint8_t* b_base = smem + L.b_off;
  float* s_epi_scale = reinterpret_cast<float*>(smem + L.epi_off);
  float* s_epi_shift = s_epi_scale + p.n_tile * p.n_ntiles;
  float* s_pro_scale = reinterpret_cast<float*>(smem + L.pro_off);
  float* s_pro_shift = s_pro_scale + p.n_chunks * 64;
  float* s_head_w = reinterpret_cast<float*>(smem + L.head_off);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  if (tid == 0 && p.trace && blockIdx.x == 0) p.trace[8 + 2000 * 4 + 1] = (2ull << 48) | (clock64() & 0xFFFFFFFFull);
  if (tid == 0 && p.gt && blockIdx.x == 0) p.gt[0] = globaltimer_ns();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], p.b_pair ? 2 : 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 128 * kEpiGroups);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  {
    // Per-layer constants -> smem. They are weights-side data (never produced by a preceding kernel), so this
    // runs BEFORE griddepcontrol.wait and overlaps the previous layer's tail under programmatic dependent
    // launch.  All global loads are issued before the first dependent store.
    const int cout = p.n_tile * p.n_ntiles;
    const int npro = PROLOGUE ? p.n_chunks * 64 : 0;
    const bool head = p.epi_mode == EPI_HEAD;
    for (int i0 = 0; i0 < cout || i0 < npro; i0 += blockDim.x) {
      const int i = i0 + tid;
      float es = 1.f, eh = 0.f, ps = 0.f, ph = 0.f, hw = 0.f;
      if (i < p.cout) {
        if (p.epi_scale) es = p.epi_scale[i];
        if (p.epi_shift) eh = p.epi_shift[i];
        if (head) hw = p.head_w[i];
      }
      if (i < npro) { ps = p.pro_scale[i]; ph = p.pro_shift[i]; }
      if (i < cout) {
        s_epi_scale[i] = es; s_epi_shift[i] = eh;
        if (head) s_head_w[i] = hw;
      }
      if (i < npro) { s_pro_scale[i] = ps; s_pro_shift[i] = ph; }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (p.b_pair) cluster_sync_all();   // peer's barriers are initialised before anything remote can arrive on them
  pdl_wait();               // activations written by the previous layer are complete and visible from here on
  // Trigger only AFTER our own wait: a dependent that starts now can rely on every kernel before this one
  // being complete (dense_layer_kernel reads older channels of the concat buffer ahead of its own wait).
  pdl_launch_dependents();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0 && p.trace && blockIdx.x == 0) { p.trace[8 + 2000 * 4] = clock64() & 0xFFFFFFFFull; p.trace[4] = 2; }
  const int n_bgroups = p.n_entries / p.b_group;  // B stages per (work item, channel chunk)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer: activations (A ring)
    if (elect_one()) {
      TraceCursor tc = trace_open(p, 0);
      uint32_t sa = 0, pa = 0;  // ring slot + phase parity
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const WorkItem wi = decode_item(p, item);
        const int ebase = wi.ph * p.n_entries;
        for (int c = 0; c < p.n_chunks; ++c) {
          const int c0 = c * 64;
          if (MODE != MODE_T) {
            mbar_wait(&a_empty[sa], pa ^ 1);
            if ((p.dbg_skip & 2) && pa) {
              mbar_expect_tx(&a_full[sa], 0);
              if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
              continue;
            }
            mbar_expect_tx(&a_full[sa], p.a_tx_bytes);
            uint8_t* dst = a_base + sa * p.a_stage_bytes;
            if (MODE == MODE_D) {
              for (int s = 0; s < p.sub; ++s)
                tma_load_2d(&map_a, &a_full[sa], dst + s * kATileBytes, c0, wi.m0 + s * 128);
            } else {
              tma_load_4d(&map_a, &a_full[sa], dst, c0, wi.w0 - p.halo_left, wi.h0 - p.halo_top, wi.n0);
            }
            trace_ev(tc, 1, item);
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          } else {
            for (int g = 0; g < n_bgroups; ++g) {  // one shifted A box per tap
              const TapEntry ent = p.entries[ebase + g];
              mbar_wait(&a_empty[sa], pa ^ 1);
              mbar_expect_tx(&a_full[sa], p.a_tx_bytes);
              tma_load_4d(&map_a, &a_full[sa], a_base + sa * p.a_stage_bytes, c0, wi.w0 * p.stride + ent.dx,
                          wi.h0 * p.stride + ent.dy, wi.n0);
              trace_ev(tc, 1, item);
              if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            }
          }
        }
      }
      trace_close(p, tc, 0);
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ TMA producer: weights (B ring)
    if (elect_one()) {
      uint32_t sb = 0, pb = 0, b_pass = 0;
      if (p.b_resident) {
        // Weight-stationary: the high-resolution decoder layers are L2->SM bandwidth bound when every CTA
        // re-streams the layer's weights for every 128/256-pixel item; when they fit, load them exactly once.
        if (blockIdx.x < p.n_items)
          for (int c = 0; c < p.n_chunks; ++c) {
            mbar_expect_tx(&b_full[c], p.b_stage_bytes);
            tma_load_3d(&map_b, &b_full[c], b_base + c * p.b_stage_bytes, c * 64, 0,
                        p.phase_fixed ? static_cast<int>(blockIdx.x % p.n_phase_items) * p.n_entries : 0);
          }
      } else if (p.b_pair) {
        // CTA pair: both CTAs walk the same (n-tile, phase, chunk, tap) sequence on neighbouring M tiles.  Each
        // fetches rows [rank * N/2, +N/2) of every tap's weight tile and multicasts them into both CTAs, so the
        // pair pulls each weight byte out of L2 once.  A slot is refilled only after BOTH MMA warps released it
        // (b_empty counts 2, arrived by multicast tcgen05.commit).
        const uint32_t rank = cluster_ctarank();
        const int half_rows = p.n_tile >> 1;
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          const int rest = item / p.n_mtiles;
          const int n0 = (rest % p.n_ntiles) * p.n_tile;
          const int ebase = (rest / p.n_ntiles) * p.n_entries;
          for (int c = 0; c < p.n_chunks; ++c) {
            for (int g = 0; g < n_bgroups; ++g) {
              mbar_wait(&b_empty[sb], pb ^ 1);
              mbar_expect_tx(&b_full[sb], p.b_stage_bytes);
              uint8_t* dst = b_base + sb * p.b_stage_bytes + rank * half_rows * 128;
              for (int j = 0; j < p.b_group; ++j)
                tma_load_3d_mcast(&map_b, &b_full[sb], dst + j * p.n_tile * 128, c * 64, n0 + rank * half_rows,
                                  ebase + g * p.b_group + j, 3);
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; }
            }
          }
        }
      } else {
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
          const int rest = item / p.n_mtiles;
          const int n0 = (rest % p.n_ntiles) * p.n_tile;
          const int ebase = (rest / p.n_ntiles) * p.n_entries;
          for (int c = 0; c < p.n_chunks; ++c) {
            for (int g = 0; g < n_bgroups; ++g) {
              if ((p.dbg_skip & 16) && b_pass) { if (++sb == p.b_stages) sb = 0; continue; }
              mbar_wait(&b_empty[sb], pb ^ 1);
              if ((p.dbg_skip & 1) && pb) {
                mbar_expect_tx(&b_full[sb], 0);
              } else {
                mbar_expect_tx(&b_full[sb], p.b_stage_bytes);
                tma_load_3d(&map_b, &b_full[sb], b_base + sb * p.b_stage_bytes, c * 64, n0, ebase + g * p.b_group);
              }
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; ++b_pass; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // One elected thread feeds the tensor core.  At N <= 64 an MMA retires every ~48 clocks (measured,
    // tools/mma_probe.cu), so the loop body must stay at a few uniform-datapath instructions per tcgen05.mma:
    // 64-bit descriptors advanced by adds, four K-steps per asm block, and `elect_one()` (not `lane == 0`) so
    // that ptxas emits bare UTCHMMA instead of a per-thread ELECT/BRA.U.ANY loop around each one.
    if (elect_one()) {
      TraceCursor tc = trace_open(p, 1);
      const uint32_t idesc = make_idesc_f16(p.n_tile);
      const uint64_t a_desc0 = (static_cast<uint64_t>(sw128_desc_hi((MODE == MODE_H) ? p.box_w * 128 : 1024)) << 32) |
                               sw128_desc_lo(smem_u32(a_base));
      const uint64_t b_desc0 = (static_cast<uint64_t>(sw128_desc_hi(1024)) << 32) | sw128_desc_lo(smem_u32(b_base));
      const uint32_t a_stage_u = p.a_stage_bytes >> 4, b_stage_u = p.b_stage_bytes >> 4;
      const uint32_t b_ent_u = (p.n_tile * 128) >> 4;
      const uint32_t a_sub_u = (MODE == MODE_D) ? (kATileBytes >> 4) : 64u;  // H: 8 pixels = 8 rows of 128 B
      const uint32_t n_tile = p.n_tile, sub = p.sub, bgroup = p.b_group;
      const uint32_t d_group_stride = sub * n_tile;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0, as = 0, ap = 0, b_pass_m = 0;
      bool b_loaded = false;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        mbar_wait(&acc_empty[as], ap ^ 1);
        tc_fence_after();
        trace_ev(tc, 0, item);
        const uint32_t d_stage = tmem_base + as * p.n_groups * d_group_stride;
        const int ebase = p.phase_fixed ? (item % p.n_phase_items) * p.n_entries
                                        : ((MODE == MODE_T) ? (item / (p.n_mtiles * p.n_ntiles)) * p.n_entries : 0);
        for (int c = 0; c < p.n_chunks; ++c) {
          int ks = (p.Cin - c * 64 + 15) >> 4;
          ks = ks > 4 ? 4 : ks;
          uint64_t a_stage_desc = 0;
          if (MODE != MODE_T) {
            mbar_wait(PROLOGUE ? &a_ready[sa] : &a_full[sa], pa);
            trace_ev(tc, 1, item);
            a_stage_desc = a_desc0 + sa * a_stage_u;
          }
          int e = ebase;
          const uint32_t first_mask = (c == 0) ? p.tap_first_mask : 0u;
          bool fast_done = false;
          if constexpr (FAST != 0 && MODE == MODE_H) {
            if ((ks == 4 || ks == 2) && !(p.dbg_skip & ~3)) {
              IssueRing ring{b_full, b_empty, sb, pb};
              const uint32_t acc_c = (c > 0) ? 1u : 0u;
#define DP_ISSUE(...)                                                                                                 \
  do {                                                                                                                \
    if (ks == 4)                                                                                                      \
      issue_chunk_h<4, __VA_ARGS__>(p, ring, a_stage_desc, b_desc0, b_stage_u, b_ent_u, d_stage, n_tile, idesc, acc_c, \
                                    c, b_loaded, tc, item);                                                           \
    else                                                                                                              \
      issue_chunk_h<2, __VA_ARGS__>(p, ring, a_stage_desc, b_desc0, b_stage_u, b_ent_u, d_stage, n_tile, idesc, acc_c, \
                                    c, b_loaded, tc, item);                                                           \
  } while (0)
              if constexpr (FAST < 4000) {
                DP_ISSUE(FAST / 100, (FAST / 10) % 10, FAST % 10);
              } else {
                constexpr int G = (FAST - 4000) / 100, S = ((FAST - 4000) / 10) % 10;
                const int pg = ebase / p.n_entries;   // this CTA's phase group (constant over its items)
                if constexpr (G == 2) {
                  if (pg == 0) DP_ISSUE(4, S, 8, 0, 8); else DP_ISSUE(4, S, 8, 8, 8);
                } else {
                  if (pg == 0) DP_ISSUE(4, S, 4, 0, 4);
                  else if (pg == 1) DP_ISSUE(4, S, 4, 4, 4);
                  else if (pg == 2) DP_ISSUE(4, S, 4, 8, 4);
                  else DP_ISSUE(4, S, 4, 12, 4);
                }
              }
#undef DP_ISSUE
              fast_done = true;
              sb = ring.sb; pb = ring.pb;
            }
          }
          for (int g = 0; g < (fast_done ? 0 : n_bgroups); ++g) {
            if (MODE == MODE_T) {
              mbar_wait(&a_full[sa], pa);
              a_stage_desc = a_desc0 + sa * a_stage_u;
            }
            if (p.b_resident) {
              sb = c;                                   // stage c == channel chunk c, filled once
              if (!b_loaded) mbar_wait(&b_full[sb], 0);
            } else if (!((p.dbg_skip & 16) && b_pass_m)) {
              mbar_wait(&b_full[sb], pb);
            }
            if (p.dbg_skip & 32) tc_fence_after();   // not needed (see issue_chunk_h); kept switchable for experiments
            trace_ev(tc, 2, item);
            uint64_t b_desc = b_desc0 + sb * b_stage_u;
            if (ks == 4) {
              for (uint32_t j = 0; j < bgroup; ++j, ++e, b_desc += b_ent_u) {
                const uint32_t flag0 = ((first_mask >> e) & 1u) ^ 1u;
                uint64_t a_desc = a_stage_desc + ((p.dbg_skip & 64) ? 0u : p.tap_a[e]);
                uint32_t d = d_stage + p.tap_d[e];
                for (uint32_t s = 0; s < sub; ++s, a_desc += a_sub_u, d += n_tile)
                  umma_f16_ss_k4(d, a_desc, b_desc, idesc, flag0);
                trace_ev(tc, 4, item);
              }
            } else if (ks == 2) {  // 32-channel tail chunk
              for (uint32_t j = 0; j < bgroup; ++j, ++e, b_desc += b_ent_u) {
                const uint32_t flag0 = ((first_mask >> e) & 1u) ^ 1u;
                uint64_t a_desc = a_stage_desc + ((p.dbg_skip & 64) ? 0u : p.tap_a[e]);
                uint32_t d = d_stage + p.tap_d[e];
                for (uint32_t s = 0; s < sub; ++s, a_desc += a_sub_u, d += n_tile)
                  umma_f16_ss_k2(d, a_desc, b_desc, idesc, flag0);
              }
            } else {  // other channel tails (Cin % 64 in {16, 48})
              for (uint32_t j = 0; j < bgroup; ++j, ++e, b_desc += b_ent_u) {
                const uint32_t flag0 = ((first_mask >> e) & 1u) ^ 1u;
                uint64_t a_desc = a_stage_desc + ((p.dbg_skip & 64) ? 0u : p.tap_a[e]);
                uint32_t d = d_stage + p.tap_d[e];
                for (uint32_t s = 0; s < sub; ++s, a_desc += a_sub_u, d += n_tile)
                  for (int k = 0; k < ks; ++k)
                    umma_f16_ss(d, a_desc + 2 * k, b_desc + 2 * k, idesc, k ? 1u : flag0);
              }
            }
            if (!p.b_resident) {
              if ((p.dbg_skip & 16) && b_pass_m) { /* timing experiment: no ring protocol after the first pass */ }
              else if (p.b_pair) umma_commit_mcast(&b_empty[sb], 3);
              else umma_commit(&b_empty[sb]);
              trace_ev(tc, 5, item);
              if (++sb == p.b_stages) { sb = 0; pb ^= 1; ++b_pass_m; }
            }
            if (MODE == MODE_T) {
              umma_commit(&a_empty[sa]);
              if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
            }
          }
          if (MODE != MODE_T) {
            umma_commit(&a_empty[sa]);
            if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
          }
        }
        b_loaded = true;
        umma_commit(&acc_full[as]);
        trace_ev(tc, 3, item);
        if (++as == p.acc_stages) { as = 0; ap ^= 1; }
      }
      trace_close(p, tc, 1);
    }
  } else if ((warp >= 4 && warp < 8) || (kEpiGroups == 2 && warp >= 8)) {
    // ------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int eg = (warp >= 8) ? 1 : 0;   // epilogue group: takes the work units u with u % kEpiGroups == eg
    const int r = q * 32 + lane;  // accumulator row == TMEM lane == pixel within the sub-tile
    const bool has_scale = p.epi_scale != nullptr;
    uint32_t as = 0, ap = 0;
    TraceCursor tc;
    if (r == 0 && eg == 0) tc = trace_open(p, 2);
    if constexpr (RESIDUAL) {
      // Inception-ResNet block tail (MODE_D, sub == 1, one accumulator group): out = act(out_old + acc + shift),
      // in place.  The old values of this thread's pixel row (<= 256 channels = 16 x 32 B) are fetched BEFORE the
      // accumulator wait, so their HBM/L2 latency hides behind the item's MMAs instead of serialising the
      // epilogue (one dependent ~1 us load per 32 columns otherwise).
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const WorkItem wi = decode_item(p, item);
        const int m = wi.m0 + r;
        const bool valid = m < p.m_total;
        const int ch0 = wi.nt * p.n_tile;
        __half* orow = p.out + static_cast<long long>(m) * p.out_ctot + p.out_choff + ch0;
        uint32_t old[16][8];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
          for (int i = 0; i < 8; ++i) old[j][i] = 0u;
          if (valid && j * 16 < p.n_tile && ch0 + j * 16 < p.cout) ld_global_v8(orow + j * 16, old[j]);
        }
        mbar_wait(&acc_full[as], ap);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * p.n_tile;
#pragma unroll
        for (int j2 = 0; j2 < 8; ++j2) {
          const int cc = 32 * j2;
          if (cc >= p.n_tile || ch0 + cc >= p.cout) break;
          uint32_t v[2][16];
          const bool two = cc + 16 < p.n_tile && ch0 + cc + 16 < p.cout;
          tmem_ld16(taddr + cc, v[0]);
          if (two) tmem_ld16(taddr + cc + 16, v[1]);
          tmem_ld_wait();
#pragma unroll
          for (int hsel = 0; hsel < 2; ++hsel) {
            if (hsel == 1 && !two) break;
            const int cb = ch0 + cc + 16 * hsel;
            float f[16];
            epi_affine16(v[hsel], s_epi_scale + cb, s_epi_shift + cb, has_scale, false, f);
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float2 o2 = __half22float2(*reinterpret_cast<const __half2*>(&old[2 * j2 + hsel][i]));
              float a = f[2 * i] + o2.x, b = f[2 * i + 1] + o2.y;
              if (p.relu) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
              __half2 h2 = __floats2half2_rn(a, b);
              pk[i] = *reinterpret_cast<uint32_t*>(&h2);
            }
            if (valid) st_global_v8(orow + cc + 16 * hsel, pk);
          }
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[as]);
        if (++as == p.acc_stages) { as = 0; ap ^= 1; }
      }
    } else
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const WorkItem wi = decode_item(p, item);
      mbar_wait(&acc_full[as], ap);
      tc_fence_after();
      trace_ev(tc, 0, item);
      const int ch0 = wi.nt * p.n_tile;
      int unit = 0;   // (accumulator [, 32-column step]) counter, identical in both groups
      for (int g = 0; g < ((p.dbg_skip & 8) ? 0 : p.n_groups); ++g) {
        const int ph = (MODE == MODE_H) ? (p.phase_fixed ? wi.ph * p.n_groups + g : g) : wi.ph;
        for (int s = 0; s < p.sub; ++s) {
          if (kEpiGroups == 2 && (p.epi_mode == EPI_HEAD || !p.epi_direct)) {
            // whole accumulators are the unit: the head needs a pixel's full channel dot product in one thread, the
            // staged store path owns per-warp staging rows
            const bool mine = (p.epi_mode == EPI_HEAD && (p.n_groups * p.sub) % 2 == 0) ? ((unit & 1) == eg) : (eg == 0);
            ++unit;
            if (!mine) continue;
          }
          // ---- where does this accumulator row land?
          bool valid;
          long long opix;  // output pixel index (flat over n, oh, ow)
          int n = 0, h = 0, w = 0;
          if (MODE == MODE_D) {
            const int m = wi.m0 + s * 128 + r;
            valid = m < p.m_total;
            opix = m;
            if (p.epi_mode == EPI_HEAD) {
              n = m / (p.H * p.W);
              const int rem = m - n * p.H * p.W;
              h = rem / p.W; w = rem - h * p.W;
            }
          } else {
            if (MODE == MODE_T) {
              w = wi.w0 + r % p.box_w;
              const int t = r / p.box_w;
              h = wi.h0 + t % p.box_h;
              n = wi.n0 + t / p.box_h;
            } else {
              w = wi.w0 + 8 * s + (r & 7);
              h = wi.h0 + (r >> 3);
              n = wi.n0;
            }
            valid = (n < p.n_img) && (h < p.OH) && (w < p.OW);
            if (p.up2) {
              h = 2 * h + (ph >> 1);
              w = 2 * w + (ph & 1);
              opix = (static_cast<long long>(n) * (2 * p.OH) + h) * (2 * p.OW) + w;
            } else {
              opix = (static_cast<long long>(n) * p.OH + h) * p.OW + w;
            }
          }
          const uint32_t taddr =
              tmem_base + (static_cast<uint32_t>(q * 32) << 16) + ((as * p.n_groups + g) * p.sub + s) * p.n_tile;
          float head_acc = 0.f;
          __half* orow = p.out ? p.out + opix * p.out_ctot + p.out_choff + ch0 : nullptr;
          if (p.epi_mode == EPI_HEAD) {
            for (int cc = 0; cc < p.n_tile; cc += 32) {
              uint32_t v[2][16];
              const bool two = cc + 16 < p.n_tile;
              tmem_ld16(taddr + cc, v[0]);
              if (two) tmem_ld16(taddr + cc + 16, v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int hsel = 0; hsel < 2; ++hsel) {
                if (hsel == 1 && !two) break;
                const int cb = ch0 + cc + 16 * hsel;
                float f[16];
                epi_affine16(v[hsel], s_epi_scale + cb, s_epi_shift + cb, has_scale, p.relu != 0, f);
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  const float4 hw = *reinterpret_cast<const float4*>(s_head_w + cb + 4 * i4);
                  head_acc = fmaf(f[4 * i4], hw.x, head_acc);
                  head_acc = fmaf(f[4 * i4 + 1], hw.y, head_acc);
                  head_acc = fmaf(f[4 * i4 + 2], hw.z, head_acc);
                  head_acc = fmaf(f[4 * i4 + 3], hw.w, head_acc);
                }
              }
            }
          } else if (p.epi_direct) {
            // 32 columns per step: TMEM -> registers -> BN shift/ReLU -> fp16 -> two 256-bit stores per thread
            // (each a full 32-byte sector of this pixel's channel run); no shared-memory round trip.
            for (int cc = 0; cc < p.n_tile && ch0 + cc < p.cout; cc += 32) {
              if (kEpiGroups == 2 && ((unit++ & 1) != eg)) continue;
              uint32_t v[2][16];
              const bool two = cc + 16 < p.n_tile && ch0 + cc + 16 < p.cout;
              tmem_ld16(taddr + cc, v[0]);
              if (two) tmem_ld16(taddr + cc + 16, v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int hsel = 0; hsel < 2; ++hsel) {
                if (hsel == 1 && !two) break;
                const int cb = ch0 + cc + 16 * hsel;
                float f[16];
                epi_affine16(v[hsel], s_epi_scale + cb, s_epi_shift + cb, has_scale, p.relu != 0, f);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __half2 h2 = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                  pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                }
                if (valid && !(p.dbg_skip & 4)) st_global_v8(orow + cc + 16 * hsel, pk);
              }
            }
          } else {
            // Panels of <= 64 channels: TMEM -> registers -> BN/ReLU -> fp16 -> this warp's padded staging rows
            // -> coalesced 16-byte stores (8 lanes cover one pixel's 128 contiguous bytes, 4 pixels per
            // instruction) instead of 32 scattered half-sector writes per instruction.
            uint8_t* stage = smem + L.stage_off + q * kEpiStageBytes;
            const unsigned long long my_row = reinterpret_cast<unsigned long long>(orow);
            for (int cc = 0; cc < p.n_tile; cc += 64) {
              const int pw = (p.n_tile - cc) < 64 ? (p.n_tile - cc) : 64;  // panel width (multiple of 16)
              uint32_t v[4][16];
              tmem_ld16(taddr + cc, v[0]);
              if (pw > 16) tmem_ld16(taddr + cc + 16, v[1]);
              if (pw > 32) tmem_ld16(taddr + cc + 32, v[2]);
              if (pw > 48) tmem_ld16(taddr + cc + 48, v[3]);
              tmem_ld_wait();
#pragma unroll
              for (int hsel = 0; hsel < 4; ++hsel) {
                if (hsel * 16 >= pw) break;
                const int cb = ch0 + cc + 16 * hsel;
                float f[16];
                epi_affine16(v[hsel], s_epi_scale + cb, s_epi_shift + cb, has_scale, p.relu != 0, f);
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  __half2 h2 = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
                  pk[i] = *reinterpret_cast<uint32_t*>(&h2);
                }
                uint4* dst = reinterpret_cast<uint4*>(stage + lane * kEpiRowBytes + hsel * 32);
                dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              }
              __syncwarp();
              const int k = pw >> 3;            // 16-byte chunks per row: 8 (64-channel panel) or 4 (32)
              const int ksh = (k == 8) ? 3 : 2;
              for (int it0 = 0; it0 < k; it0 += 4) {   // 4 independent row-pointer / load / store chains at a time
                unsigned long long rp[4];
                int rv[4];
                uint4 val[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int row = ((it0 + u) * 32 + lane) >> ksh;
                  rp[u] = __shfl_sync(0xffffffffu, my_row, row);
                  rv[u] = __shfl_sync(0xffffffffu, static_cast<int>(valid), row);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const int idx = (it0 + u) * 32 + lane;
                  val[u] = *reinterpret_cast<const uint4*>(stage + (idx >> ksh) * kEpiRowBytes + (idx & (k - 1)) * 16);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  if (rv[u]) {
                    const int j = ((it0 + u) * 32 + lane) & (k - 1);
                    *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(rp[u]) + cc + j * 8) = val[u];
                  }
                }
              }
              __syncwarp();
            }
          }
          if (p.epi_mode == EPI_HEAD && valid) {
            // softmax over 2 logits, channel 1 == sigmoid(z1 - z0) (Segmentation.py:167 uses [..., 1] only)
            const float z = head_acc + p.head_b;
            const float prob = 1.f / (1.f + expf(-z));
            int di, dj;
            d4_src(p.pass->tta_out, h, w, p.P, di, dj);
            p.pass->probs_out[(static_cast<long long>(n + p.img0) * p.P + di) * p.P + dj] = prob;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty[as]);
      trace_ev(tc, 1, item);
      if (++as == p.acc_stages) { as = 0; ap ^= 1; }
    }
    if (r == 0 && eg == 0) trace_close(p, tc, 2);
  } else if (PROLOGUE && warp >= 8) {
    // ------------------------------------------------------------------ A-tile pre-activation (MODE_D only)
    const int t = tid - 256;
    TraceCursor tc;
    if (t == 0) tc = trace_open(p, 3);
    uint32_t sa = 0, pa = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      for (int c = 0; c < p.n_chunks; ++c) {
        mbar_wait(&a_full[sa], pa);
        const __half2 zero2 = __float2half2_rn(0.f);
        {
          // Thread t owns logical 16-byte chunk i = t & 7 (8 channels) of rows (t >> 3) + 16 k of every sub-tile:
          // its 16 BN terms stay in registers for the whole stage, so shared-memory traffic is the data itself
          // (a quarter-warp covers one full 128-byte row: conflict-free under the 128B swizzle).  Loads of a
          // batch are issued before its stores: the in-place stores may alias as far as the compiler knows.
          const int i = t & 7;
          const int ch = c * 64 + i * 8;
          const float4 sc0 = *reinterpret_cast<const float4*>(s_pro_scale + ch);
          const float4 sc1 = *reinterpret_cast<const float4*>(s_pro_scale + ch + 4);
          const float4 sh0 = *reinterpret_cast<const float4*>(s_pro_shift + ch);
          const float4 sh1 = *reinterpret_cast<const float4*>(s_pro_shift + ch + 4);
          for (int s = 0; s < p.sub; ++s) {
            uint8_t* tile = a_base + sa * p.a_stage_bytes + s * kATileBytes;
            uint4 raw[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = (t >> 3) + 16 * k;
              raw[k] = *reinterpret_cast<const uint4*>(tile + rr * 128 + ((i ^ (rr & 7)) << 4));
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              __half2* hv = reinterpret_cast<__half2*>(&raw[k]);
              float2 x;
              x = __half22float2(hv[0]);
              hv[0] = __floats2half2_rn(fmaf(x.x, sc0.x, sh0.x), fmaf(x.y, sc0.y, sh0.y));
              x = __half22float2(hv[1]);
              hv[1] = __floats2half2_rn(fmaf(x.x, sc0.z, sh0.z), fmaf(x.y, sc0.w, sh0.w));
              x = __half22float2(hv[2]);
              hv[2] = __floats2half2_rn(fmaf(x.x, sc1.x, sh1.x), fmaf(x.y, sc1.y, sh1.y));
              x = __half22float2(hv[3]);
              hv[3] = __floats2half2_rn(fmaf(x.x, sc1.z, sh1.z), fmaf(x.y, sc1.w, sh1.w));
              if (p.pro_relu) {
                hv[0] = __hmax2(hv[0], zero2); hv[1] = __hmax2(hv[1], zero2);
                hv[2] = __hmax2(hv[2], zero2); hv[3] = __hmax2(hv[3], zero2);
              }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int rr = (t >> 3) + 16 * k;
              *reinterpret_cast<uint4*>(tile + rr * 128 + ((i ^ (rr & 7)) << 4)) = raw[k];
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(&a_ready[sa]);
        trace_ev(tc, 0, item);
        if (++sa == p.a_stages) { sa = 0; pa ^= 1; }
      }
    }
    if (t == 0) trace_close(p, tc, 3);
  }

  tc_fence_before();
  __syncthreads();
  if (p.b_pair) cluster_sync_all();   // do not exit while the peer may still multicast into this CTA
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
    if (lane == 0 && p.gt && blockIdx.x == 0) p.gt[1] = globaltimer_ns();
    if (lane == 0 && p.trace && blockIdx.x == 0) { p.trace[8 + 2000 * 4 + 2] = (1ull << 48) | (clock64() & 0xFFFFFFFFull); p.trace[4] = 3; }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Naive reference of the SAME packed problem (same tap table, same packed weights, same buffers), one thread
// per output value on CUDA cores.  Debug/bisect tool only (DP_NAIVE_CONV=1): lets a GPU run compare the
// tcgen05 path against an obviously-correct evaluation layer by layer without shipping tensors back.
struct NaiveConvParams {
  int n_img, H, W, Cin, in_ctot, in_choff;
  int OH, OW, stride, residual;
  int Cout, out_ctot, out_choff;
  int n_entries_total, n_groups, entries_per_group, up2;
  int relu, pro_mode;  // pro_mode: 0 none, 1 affine, 2 affine+relu
  TapEntry entries[kMaxEntries];
  // host-precomputed per tap entry (kernel parameters live in the constant bank, so the issue loop reads these
  // straight into uniform registers): A-descriptor offset in 16-byte units, accumulator column offset, and a
  // bitmask of entries that are the first of their accumulator group
  uint32_t tap_a[kMaxEntries];
  uint32_t tap_d[kMaxEntries];
  uint32_t tap_first_mask;
  const __half* in;
  const __half* w;  // [entries][Cout][Cin]
  const float* epi_scale;
  const float* epi_shift;
  const float* pro_scale;
  const float* pro_shift;
  __half* out;
};

__global__ void conv_naive_kernel(const NaiveConvParams p) {
  const long long total = static_cast<long long>(p.n_img) * p.OH * p.OW * p.n_groups * p.Cout;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int co = idx % p.Cout;
    long long r = idx / p.Cout;
    const int g = r % p.n_groups; r /= p.n_groups;
    const int w = r % p.OW; r /= p.OW;
    const int h = r % p.OH;
    const int n = r / p.OH;
    float acc = 0.f;
    for (int e = 0; e < p.n_entries_total; ++e) {
      const TapEntry ent = p.entries[e];
      if (ent.group != g) continue;
      const int ih = h * p.stride + ent.dy, iw = w * p.stride + ent.dx;
      if (ih < 0 || ih >= p.H || iw < 0 || iw >= p.W) continue;
      const __half* a = p.in + ((static_cast<long long>(n) * p.H + ih) * p.W + iw) * p.in_ctot + p.in_choff;
      const __half* wt = p.w + (static_cast<long long>(e) * p.Cout + co) * p.Cin;
      for (int ci = 0; ci < p.Cin; ++ci) {
        float x = __half2float(a[ci]);
        if (p.pro_mode) {
          x = fmaf(x, p.pro_scale[ci], p.pro_shift[ci]);
          if (p.pro_mode == 2) x = fmaxf(x, 0.f);
          x = __half2float(__float2half_rn(x));  // the tensor-core path rounds the activated tile to fp16
        }
        acc = fmaf(x, __half2float(wt[ci]), acc);
      }
    }
    float y = fmaf(acc, p.epi_scale ? p.epi_scale[co] : 1.f, p.epi_shift ? p.epi_shift[co] : 0.f);
    long long opix;
    if (p.up2)
      opix = (static_cast<long long>(n) * 2 * p.H + 2 * h + (g >> 1)) * (2 * p.W) + 2 * w + (g & 1);
    else
      opix = (static_cast<long long>(n) * p.OH + h) * p.OW + w;
    if (p.residual) y += __half2float(p.out[opix * p.out_ctot + p.out_choff + co]);
    if (p.relu) y = fmaxf(y, 0.f);
    p.out[opix * p.out_ctot + p.out_choff + co] = __float2half_rn(y);
  }
}

}  // namespace dp
