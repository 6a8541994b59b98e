// Runtime behind include/digipath_b200.h: parses the flat "DPB1" model container, owns the NHWC fp16
// activation buffers, turns every op of the layer program into a launch plan (TMA tensor maps + tile shapes
// chosen per layer geometry) and replays the plan on the caller's stream.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/digipath_b200.h"
#include "aux.cuh"
#include "conv_tc.cuh"
#include "crf.cuh"
#include "crf_lattice.cuh"
#include "dense_layer.cuh"
#include "dense_block.cuh"
#include "precise.cuh"
#include "precise_tc.cuh"
#include "jpeg_enc.cuh"
#include "tissue.cuh"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
thread_local uint64_t g_captured = 0;   // kernel launches recorded into the graph being captured on this thread

int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CU_OK(expr)                                                                           \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_OK()                                                                                        \
  do {                                                                                                     \
    cudaError_t e__ = cudaGetLastError();                                                                  \
    if (e__ != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    {                                                                                                      \
      cudaStreamCaptureStatus cs__ = cudaStreamCaptureStatusNone;                                          \
      cudaStreamIsCapturing(st, &cs__);                                                                    \
      if (cs__ == cudaStreamCaptureStatusNone) g_launches.fetch_add(1, std::memory_order_relaxed);         \
      else ++g_captured;                                                                                   \
    }                                                                                                      \
  } while (0)

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// fp16 tensor map, 128-byte swizzle, zero fill out of bounds. dims/strides innermost first; strides in bytes
// for dims 1..rank-1.
// `pix_stride` > 1 sets the traversal (element) stride of dims 1 and 2 (W, H of an NHWC map): the box then
// covers box[i] input elements and delivers ceil(box[i] / pix_stride) of them -- a strided conv's operand tile.
int make_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
             const uint32_t* box, int pix_stride = 1, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_FLOAT16) {
  auto enc = get_encode();
  if (!enc) return fail("cuTensorMapEncodeTiled entry point not available (driver too old?)");
  uint32_t estr[5] = {1, 1, 1, 1, 1};
  if (pix_stride > 1 && rank >= 3) estr[1] = estr[2] = (uint32_t)pix_stride;
  CUresult r = enc(map, dtype, rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r, rank,
                (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  return 0;
}

// ---------------------------------------------------------------- container format (see weights.py)
struct BlobHeader {
  char magic[4];
  uint32_t version, n_bufs, n_ops, patch, precision;   // precision 0: fp16 weights + activations, 1: fp32
  uint64_t data_off, total_bytes;
  uint64_t reserved1[4];
};
static_assert(sizeof(BlobHeader) == 72, "header layout");
struct BlobBuf {
  int32_t H, W, C, pad;
};
enum : int { OP_STEM_IM2COL = 1, OP_MAXPOOL = 2, OP_CONV = 3, OP_BNPOOL = 4, OP_STEM_S2D = 5, OP_DENSE_LAYER = 6,
              OP_AVGPOOL3 = 7, OP_DWCONV = 8, OP_GAP = 9, OP_BCAST = 10, OP_RESIZE = 11, OP_HEAD_DOT = 12,
              OP_HEAD_RESIZE = 13 };
struct BlobOp {
  int32_t type, in_buf, in_choff, cin, out_buf, out_choff, cout, kind, relu, pro, head, pool;
  float head_b;
  int32_t rsv[3];
  int64_t w_off, epi_scale_off, epi_shift_off, pro_scale_off, pro_shift_off, head_w_off, rsv64[2];
};
static_assert(sizeof(BlobOp) == 128, "op layout");

struct Launch {
  int type = 0;
  // conv
  bool prologue = false;
  CUtensorMap map_a, map_b;
  dp::ConvParams cp;
  dp::NaiveConvParams np;
  int grid = 0, smem = 0;
  uint64_t macs = 0;
  // head (naive path)
  int head_C = 0;
  // fused dense layer
  CUtensorMap map_w2;
  dp::DenseLayerParams dl;
  dp::NaiveConvParams np2;  // debug path: the 3x3 half (np holds the 1x1 half)
  // 3xTF32 mode (precision 2): TMA maps of the halo kernel for np (tx[0]) / np2 (tx[1]); n_tile 0 = generic kernel
  struct TxPlan { CUtensorMap a, wh, wl; int n_tile = 0, n_ntiles = 0, kind = 0; } tx[2];   // kind 1: halo kernel, 2: 1x1 kernel
  // persistent dense-block kernel (dense_block.cuh): set on the FIRST layer of a run of dense layers on small maps;
  // the other layers of the run carry block_member and launch nothing when the whole program is executed
  int block_len = 0;
  bool block_member = false;
  CUtensorMap map_x_block;
  dp::DenseBlockParams db;
  int block_smem = 0;
};

struct SubPlan {
  int img0 = 0, n_img = 0;       // slice of the call's tile batch this sub-plan covers
  std::vector<Launch> launches;  // one per op
  std::vector<std::shared_ptr<void>> dev_allocs;   // layer tables of the dense-block kernels (cudaFree'd with the plan)
};

// A plan for one tile count: the batch is cut into `subs.size()` independent sub-batches whose op chains are
// captured as parallel branches of one CUDA graph -- the ~100 latency-bound small-map kernels of the encoder
// (16x16 / 8x8 maps) then overlap across branches instead of running back to back.
struct Plan {
  std::vector<SubPlan> subs;
  cudaGraphExec_t exec = nullptr;
  cudaGraph_t graph = nullptr;
  uint64_t n_launches = 0;   // kernel nodes of the captured graph
};

}  // namespace

struct dp_model {
  int device = 0, max_batch = 0, patch = 0, num_sms = 148;
  int precision = 0, esize = 2;  // 0 = fp16 tensor-core path; 1 = fp32 storage + fp32 FMA kernels (precise.cuh), esize 4;
                                 // 2 = fp32 storage, convs as 3xTF32 on the tensor cores (precise_tc.cuh)
  std::vector<BlobBuf> bufs;
  std::vector<BlobOp> ops;
  uint8_t* data_dev = nullptr;             // container data section (weights, BN vectors) on the device
  std::shared_ptr<void> data_owner;        // ... shared between a model and its lanes (dp_model_clone)
  float* data_hi_dev = nullptr;            // precision 2: TF32 hi / lo copies of the data section (precise_tc.cuh)
  float* data_lo_dev = nullptr;
  std::shared_ptr<void> data_hi_owner, data_lo_owner;
  size_t data_bytes = 0;
  std::vector<__half*> buf_dev;    // activation buffers (raw storage: __half or float elements, see esize)
  __half* scratch_head = nullptr;  // naive / fp32 paths: output of the head-fused conv
  uint64_t device_bytes = 0;
  int naive_conv = 0, desc_base_mode = 0, halo_pad8 = 0, profile = 0;
  int use_graph = 1, split = 1, use_pdl = 1, epi_direct = 1, use_overlap = 1, b_resident = 1, b_pair = 0;
  int dense_block = 1;   // runs of dense layers on small maps as one persistent kernel (dense_block.cuh); 2 = wherever
                         // the kernel applies, 1 = only where per-layer kernels cannot overlap (plan_dense_blocks)
  unsigned long long* gt_dev = nullptr;     // debug: per-op %globaltimer stamps (option "stamp")
  int stamp = 0;
  unsigned long long* gt_all_dev = nullptr; // debug: per-CTA stamps of the dense layers [n_ops][256][4] (option "stamp_ctas")
  int stamp_ctas = 0;
  unsigned long long* trace_dev = nullptr;  // debug timeline buffer (option "trace_op")
  int trace_op = -1, trace_block = -1;
  dp::PassDesc* pass_dev = nullptr;          // per-call arguments read by the stem and head kernels
  std::vector<cudaStream_t> branch_streams;  // capture-time fork/join streams
  std::vector<cudaEvent_t> branch_events;
  cudaStream_t cap_stream = nullptr;
  cudaEvent_t fork_event = nullptr;
  std::vector<cudaEvent_t> ev;  // profile option: 2 events per op of the last run
  int ev_begin = 0, ev_end = 0;
  std::map<std::pair<int, int>, Plan> plans;  // key: (n_tiles, split)
  std::mutex mu;
};

namespace {

template <typename T>
const T* dptr(const dp_model* m, int64_t off) {
  return off < 0 ? nullptr : reinterpret_cast<const T*>(m->data_dev + off);
}

int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Start of image `img0` inside activation buffer `buf` (element size follows the model's precision).
__half* buf_ptr(const dp_model* m, int buf, int img0) {
  const BlobBuf& bb = m->bufs[buf];
  return reinterpret_cast<__half*>(reinterpret_cast<char*>(m->buf_dev[buf]) +
                                   (size_t)img0 * bb.H * bb.W * bb.C * m->esize);
}

// MODE_H kernel variants by unrolled issue sequence (ConvParams::fast_id); 0 = generic tap loop.
typedef void (*ConvKernelH)(const CUtensorMap, const CUtensorMap, const dp::ConvParams);
struct FastKernel {
  int id;
  ConvKernelH fn;
};
#define DP_FK(ID) {ID, dp::conv_tc_kernel<dp::MODE_H, false, false, ID>}
const FastKernel kFastKernels[] = {DP_FK(0),   DP_FK(311), DP_FK(313), DP_FK(319), DP_FK(321),  DP_FK(323),  DP_FK(329),
                                   DP_FK(411), DP_FK(412), DP_FK(414), DP_FK(4210), DP_FK(4220), DP_FK(4110), DP_FK(4120)};
#undef DP_FK
ConvKernelH fast_kernel(int id) {
  for (const FastKernel& k : kFastKernels)
    if (k.id == id) return k.fn;
  return nullptr;
}

// TensorFlow padding='same': leading pad along one axis (the odd cell goes behind).
int same_pad_before(int size, int k, int stride) {
  const int out = (size + stride - 1) / stride;
  int total = (out - 1) * stride + k - size;
  if (total < 0) total = 0;
  return total / 2;
}

struct Halo {
  int top = 1, bottom = 1, left = 1, right = 1;
};

// Tap table of a conv kind (kind 6 = generic kh x kw 'same' conv, stride 1 or 2, taps ky-major) and the halo the
// taps reach around a stride-1 region.
void fill_entries(int kind, int H, int W, int kh, int kw, int stride, dp::TapEntry* e, int* n, Halo* halo) {
  Halo hl;
  if (kind == 1) {
    e[0] = {0, 0, 0, 0};
    *n = 1;
    hl = {0, 0, 0, 0};
  } else if (kind == 3) {
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) e[ky * 3 + kx] = {(int8_t)(ky - 1), (int8_t)(kx - 1), 0, 0};
    *n = 9;
  } else if (kind == 5) {  // stem as a 4-row-tap conv on the width-unrolled space-to-depth image
    for (int t = 0; t < 4; ++t) e[t] = {(int8_t)(t - 2), 0, 0, 0};
    *n = 4;
    hl.top = 2;
  } else if (kind == 6) {
    const int ph = same_pad_before(H, kh, stride), pw = same_pad_before(W, kw, stride);
    hl = {0, 0, 0, 0};
    for (int ky = 0; ky < kh; ++ky)
      for (int kx = 0; kx < kw; ++kx) {
        const int dy = ky - ph, dx = kx - pw;
        e[ky * kw + kx] = {(int8_t)dy, (int8_t)dx, 0, 0};
        if (-dy > hl.top) hl.top = -dy;
        if (dy > hl.bottom) hl.bottom = dy;
        if (-dx > hl.left) hl.left = -dx;
        if (dx > hl.right) hl.right = dx;
      }
    *n = kh * kw;
  } else {  // up2: phase (a,b), tap (ty,tx): input offset (a-1+ty, b-1+tx)
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 4; ++t) {
        const int a = ph >> 1, b = ph & 1, ty = t >> 1, tx = t & 1;
        e[ph * 4 + t] = {(int8_t)(a - 1 + ty), (int8_t)(b - 1 + tx), (int8_t)ph, 0};
      }
    *n = 16;
  }
  if (halo) *halo = hl;
}

// 3xTF32 mode: N tiling of a conv, and -- for 3x3 / up2 convs without prologue on maps the 16 x 8 regions tile -- the
// TMA maps of conv_halo_tf32x3_kernel (activation halo box, pre-split weight boxes).
int plan_tx(dp_model* m, const dp::NaiveConvParams& q, Launch::TxPlan& t) {
  t.n_ntiles = (q.Cout + dp::kTxMaxN - 1) / dp::kTxMaxN;
  const int n_tile = round_up((q.Cout + t.n_ntiles - 1) / t.n_ntiles, 16);
  t.n_tile = 0; t.kind = 0;
  const int epg = q.n_groups ? q.n_entries_total / q.n_groups : 0;
  bool halo = m->precision == 2 && q.stride == 1 && !q.pro_mode && !q.residual && q.H % 16 == 0 && q.W % 8 == 0 &&
              q.Cin % 4 == 0 && (q.up2 || (q.OH == q.H && q.OW == q.W)) && epg >= 3 && epg <= 9 && m->data_hi_dev &&
              !getenv("DP_TX_NO_HALO");
  for (int e = 0; e < q.n_entries_total && halo; ++e)
    halo = q.entries[e].dy >= -1 && q.entries[e].dy <= 1 && q.entries[e].dx >= -1 && q.entries[e].dx <= 1;
  const bool one = m->precision == 2 && !halo && q.n_entries_total == 1 && q.n_groups == 1 && q.stride == 1 && !q.up2 &&
                   q.entries[0].dy == 0 && q.entries[0].dx == 0 && q.OH == q.H && q.OW == q.W && q.Cin % 4 == 0 &&
                   m->data_hi_dev && !getenv("DP_TX_NO_1X1");
  if (!halo && !one) return 0;
  if (halo) {
    uint64_t dims[4] = {(uint64_t)(q.in_choff + q.Cin), (uint64_t)q.W, (uint64_t)q.H, (uint64_t)q.n_img};
    const uint64_t cs = (uint64_t)q.in_ctot * 4;
    uint64_t str[3] = {cs, cs * q.W, cs * q.W * q.H};
    uint32_t box[4] = {32, (uint32_t)dp::kThHaloW, (uint32_t)dp::kThHaloH, 1};
    if (make_map(&t.a, q.in, 4, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
  } else {   // flat [pixels, C] view
    uint64_t dims[2] = {(uint64_t)(q.in_choff + q.Cin), (uint64_t)q.n_img * q.H * q.W};
    uint64_t str[1] = {(uint64_t)q.in_ctot * 4};
    uint32_t box[2] = {32, 128};
    if (make_map(&t.a, q.in, 2, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
  }
  t.kind = halo ? 1 : 2;
  const size_t woff = reinterpret_cast<const uint8_t*>(q.w) - m->data_dev;
  uint64_t dims[3] = {(uint64_t)q.Cin, (uint64_t)q.Cout, (uint64_t)q.n_entries_total};
  uint64_t str[2] = {(uint64_t)q.Cin * 4, (uint64_t)q.Cin * q.Cout * 4};
  uint32_t box[3] = {32, (uint32_t)n_tile, 1};
  if (make_map(&t.wh, reinterpret_cast<const uint8_t*>(m->data_hi_dev) + woff, 3, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
  if (make_map(&t.wl, reinterpret_cast<const uint8_t*>(m->data_lo_dev) + woff, 3, dims, str, box, 1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return 1;
  t.n_tile = n_tile;
  return 0;
}

int plan_conv(dp_model* m, const BlobOp& op, int img0, int B, Launch& L) {
  using namespace dp;
  const BlobBuf& ib = m->bufs[op.in_buf];
  const int H = ib.H, W = ib.W;
  const bool up2 = op.kind == 4;
  ConvParams& p = L.cp;
  memset(&p, 0, sizeof p);
  int n_entries_total = 0;
  TapEntry table[kMaxEntries];
  memset(table, 0, sizeof table);
  const int kh = op.rsv[0] & 0xff, kw = (op.rsv[0] >> 8) & 0xff;
  const int stride = (op.kind == 6) ? ((op.rsv[0] >> 16) & 0xff) : 1;
  if (op.kind == 6 && (kh < 1 || kw < 1 || kh * kw > kMaxEntries || (stride != 1 && stride != 2)))
    return fail("conv: generic tap kernel %dx%d stride %d unsupported", kh, kw, stride);
  if (op.kind != 1 && op.kind != 3 && op.kind != 4 && op.kind != 5 && op.kind != 6) return fail("conv: unknown kind %d", op.kind);
  Halo halo;
  fill_entries(op.kind, H, W, kh, kw, stride, table, &n_entries_total, &halo);
  const int OH = (H + stride - 1) / stride, OW = (W + stride - 1) / stride;  // work grid (up2: the low-res grid)
  const int residual = (op.rsv[1] != 0) ? 1 : 0;
  if (residual && (op.kind != 1 || op.head || op.pro)) return fail("conv: residual epilogue needs a plain 1x1 conv");

  // ---- naive description (always built; used when the naive_conv option is on)
  NaiveConvParams& q = L.np;
  memset(&q, 0, sizeof q);
  q.n_img = B; q.H = H; q.W = W; q.Cin = op.cin; q.in_ctot = ib.C; q.in_choff = op.in_choff;
  q.OH = OH; q.OW = OW; q.stride = stride; q.residual = residual;
  q.Cout = op.cout;
  q.n_entries_total = n_entries_total; q.n_groups = up2 ? 4 : 1; q.up2 = up2;
  q.relu = op.relu; q.pro_mode = op.pro;
  memcpy(q.entries, table, sizeof table);
  auto buf_at = [&](int buf) { return buf_ptr(m, buf, img0); };
  q.in = buf_at(op.in_buf);
  q.w = dptr<__half>(m, op.w_off);
  q.epi_scale = dptr<float>(m, op.epi_scale_off);
  q.epi_shift = dptr<float>(m, op.epi_shift_off);
  q.pro_scale = dptr<float>(m, op.pro_scale_off);
  q.pro_shift = dptr<float>(m, op.pro_shift_off);
  if (op.head) {
    q.out = reinterpret_cast<__half*>(reinterpret_cast<char*>(m->scratch_head) +
                                      (size_t)img0 * m->patch * m->patch * op.cout * m->esize);
    q.out_ctot = op.cout; q.out_choff = 0;
  } else {
    q.out = buf_at(op.out_buf); q.out_ctot = m->bufs[op.out_buf].C; q.out_choff = op.out_choff;
  }
  L.head_C = op.cout;
  if (m->precision) {   // fp32 modes execute the description above with conv_f32_kernel / the 3xTF32 kernels
    L.macs = (uint64_t)B * OH * OW * n_entries_total * op.cin * op.cout;
    return plan_tx(m, q, L.tx[0]);
  }

  // ---- tensor-core plan
  p.n_img = B; p.H = H; p.W = W; p.Cin = op.cin;
  p.OH = OH; p.OW = OW; p.stride = stride; p.cout = op.cout; p.residual = residual;
  p.n_chunks = (op.cin + 63) / 64;
  p.up2 = up2;
  p.relu = op.relu;
  p.pro_relu = op.pro == 2;
  p.desc_base_mode = m->desc_base_mode;
  p.epi_mode = op.head ? EPI_HEAD : EPI_STORE;
  // store path is fixed at plan time (the shared-memory layout depends on it): 256-bit direct stores whenever the
  // output placement is 16-channel aligned
  p.epi_direct = (!op.head && (m->epi_direct || residual) && (op.out_choff % 16) == 0 &&
                  (m->bufs[op.out_buf].C % 16) == 0) ? 1 : 0;
  p.epi_scale = q.epi_scale; p.epi_shift = q.epi_shift;
  p.pro_scale = q.pro_scale; p.pro_shift = q.pro_shift;
  p.head_w = dptr<float>(m, op.head_w_off);
  p.head_b = op.head_b;
  p.P = m->patch;
  p.img0 = img0;
  p.pass = m->pass_dev;
  if (!op.head) {
    p.out = buf_at(op.out_buf);
    p.out_ctot = m->bufs[op.out_buf].C;
    p.out_choff = op.out_choff;
  }
  L.prologue = op.pro != 0;
  if (op.cout % 16) return fail("conv: Cout %d not a multiple of 16", op.cout);
  if (op.cin % 8) return fail("conv: Cin %d not a multiple of 8", op.cin);

  const int groups_if_h = up2 ? 4 : 1;
  if (op.kind == 1) {
    p.mode = MODE_D;
  } else if (stride == 1 && H >= 16 && H % 16 == 0 && W % 8 == 0) {
    p.mode = MODE_H;
  } else {
    // one (element-strided) TMA box per tap; a work item is 128 output pixels: whole images of a small map or
    // a band of rows of a larger one
    if (OH * OW <= 128 ? (128 % (OH * OW) != 0) : (OW > 128 || 128 % OW != 0 || OH % (128 / OW) != 0))
      return fail("conv: output map %dx%d unsupported by the per-tap mode", OH, OW);
    if (stride == 2 && (H % 2 || W % 2)) return fail("conv: stride 2 needs an even map, got %dx%d", H, W);
    p.mode = MODE_T;
  }
  if (L.prologue && p.mode != MODE_D) return fail("conv: pre-activation prologue only supported for 1x1 convs");
  if (op.head && (up2 || op.cout > 256)) return fail("conv: fused head needs a plain conv with Cout <= 256");

  p.n_groups = (p.mode == MODE_H) ? groups_if_h : 1;
  p.n_phase_items = (p.mode == MODE_T && up2) ? 4 : 1;
  p.n_entries = (p.mode == MODE_T && up2) ? 4 : n_entries_total;
  memcpy(p.entries, table, sizeof table);
  if (p.mode == MODE_T)
    for (int i = 0; i < kMaxEntries; ++i) p.entries[i].group = 0;
  // Phase-split up-conv: a work item covers G of the 4 sub-pixel phases and every CTA keeps ONE phase group for
  // all of its items, so that group's weights (4G taps x all chunks) stay resident in shared memory.  Streaming
  // the 16 tap tiles per 128-pixel item through a 4-stage ring is what bounded dec10a / dec9a (ring shallower
  // than MMA-completion lag + TMA latency); the price is loading the (small) activation halo 4/G
  // times.  Chosen when the resident weights fit beside a two-stage activation ring.
  p.phase_fixed = 0;
  if (p.mode == MODE_H && up2 && m->b_resident && op.cout <= 256 && !getenv("DP_NO_PHASE_SPLIT")) {
    const long long room = 227 * 1024 - 8 * 1024 - (p.epi_direct ? 0 : 4 * kEpiStageBytes);   // barriers, constants, slack
    for (int G = 2; G >= 1; --G) {
      const long long wbytes = 4LL * G * ((op.cin + 63) / 64) * op.cout * 128;
      const long long a2 = 2LL * round_up(18 * 10 * 128, 1024);   // two sub = 1 halo stages
      if ((op.cin + 63) / 64 <= kMaxBStages && wbytes + a2 <= room) {
        p.phase_fixed = 1;
        p.n_groups = G;
        p.n_phase_items = 4 / G;
        p.n_entries = 4 * G;
        for (int e = 0; e < 16; ++e) p.entries[e].group = (int8_t)((e / 4) % G);
        break;
      }
    }
  }

  // N tile: whole Cout if the accumulators fit; otherwise k equal tiles (multiples of 16) chosen for the least
  // padding, ties to fewer tiles.  A padded last tile (Cout 1088 -> 5 x 224) reads TMA-zero-filled weight rows
  // and its store is clipped in the epilogue.
  int max_cols = kTmemCols / p.n_groups;
  if (max_cols > 256) max_cols = 256;
  int n_tile = 0, n_ntiles = 0;
  {
    const int kmin = (op.cout + max_cols - 1) / max_cols;
    long best = -1;
    for (int k = kmin; k <= kmin + 8; ++k) {
      const int nt = round_up((op.cout + k - 1) / k, 16);
      if (nt > max_cols || nt < 16) continue;
      const int kk = (op.cout + nt - 1) / nt;
      const long cost = (long)kk * nt;
      if (best < 0 || cost < best) { best = cost; n_tile = nt; n_ntiles = kk; }
    }
    if (best < 0) return fail("conv: no valid N tile for Cout %d", op.cout);
  }
  if (op.head && n_ntiles != 1) return fail("conv: fused head needs a single N tile");
  p.n_tile = n_tile;
  p.n_ntiles = n_ntiles;
  const bool out_aligned = op.head || ((op.out_choff % 16) == 0 && (m->bufs[op.out_buf].C % 16) == 0);
  if ((residual || n_tile * n_ntiles != op.cout) && !out_aligned)
    return fail("conv: residual / clipped-N epilogue needs 16-channel aligned output placement");
  if (n_tile * n_ntiles != op.cout && !op.head) p.epi_direct = 1;   // the clipped store exists only in the direct path

  // sub-tiles per CTA tile
  const long long m_total = (long long)B * OH * OW;
  if (p.mode == MODE_D) {
    p.m_total = (int)m_total;
    // Sub-tiles of 128 pixels per item: minimise  waves x time-per-item  with the measured per-chunk costs: MMA time
    // sub x 4 x max(64, N/2) clk (N/2 only above the 48-64 clk floor) plus ~700 clk of fixed per-chunk hand-off
    // (two barrier waits, two commits, issue).  Fewer, larger items win whenever they do not add a wave.
    int sub = 1;
    if (!getenv("DP_D_SUB_BY_ITEMS")) {
      double best = 0;
      for (int cand = 1; cand <= 4; cand <<= 1) {
        if (cand * n_tile > kTmemCols) break;
        const long long items = ((m_total + cand * 128 - 1) / (cand * 128)) * p.n_ntiles;
        const long long waves = (items + m->num_sms - 1) / m->num_sms;
        const double per_item = (double)p.n_chunks * (cand * 4.0 * (n_tile / 2 > 64 ? n_tile / 2 : 64) + 700.0) + 3000.0;
        const double cost = waves * per_item;
        if (best == 0 || cost < best) { best = cost; sub = cand; }
      }
    } else {
      sub = 4;
      while (sub > 1 && (sub * n_tile > kTmemCols ||
                         ((m_total + sub * 128 - 1) / (sub * 128)) * p.n_ntiles < m->num_sms))
        sub >>= 1;
    }
    if (residual) sub = 1;   // the residual epilogue prefetches one pixel row (<= 256 channels) per thread
    p.sub = sub;
    p.n_mtiles = (int)((m_total + sub * 128 - 1) / (sub * 128));
    p.a_stage_bytes = sub * kATileBytes;
    p.a_tx_bytes = p.a_stage_bytes;
  } else if (p.mode == MODE_T) {
    p.sub = 1;
    if (OH * OW <= 128) { p.box_w = OW; p.box_h = OH; p.box_n = 128 / (OH * OW); }
    else { p.box_w = OW; p.box_h = 128 / OW; p.box_n = 1; }
    p.tiles_w = 1; p.tiles_h = OH / p.box_h;
    p.n_mtiles = ((B + p.box_n - 1) / p.box_n) * p.tiles_h;
    p.a_stage_bytes = kATileBytes;
    p.a_tx_bytes = kATileBytes;
  } else {
    int sub = (W % 16 == 0) ? 2 : 1;
    while (sub > 1 && (p.n_groups * sub * n_tile > kTmemCols)) sub >>= 1;
    // prefer two accumulator stages (epilogue of item i overlaps the MMAs of item i+1) over a wider region
    if (sub == 2 && 2 * p.n_groups * sub * n_tile > kTmemCols && 2 * p.n_groups * n_tile <= kTmemCols &&
        !getenv("DP_PREFER_SUB2"))
      sub = 1;
    // prefer more CTAs over wider regions when the layer cannot fill the GPU
    if (sub == 2 && (long long)B * (H / 16) * (W / 16) * p.n_ntiles < m->num_sms) sub = 1;
    // resident phase-split weights must leave room for two activation stages
    if (sub == 2 && p.phase_fixed &&
        (long long)p.n_entries * p.n_chunks * n_tile * 128 + 2LL * round_up(18 * 18 * 128, 1024) >
            227 * 1024 - 8 * 1024 - (p.epi_direct ? 0 : 4 * kEpiStageBytes))
      sub = 1;
    p.sub = sub;
    p.box_w = m->halo_pad8 ? round_up(8 * sub + halo.left + halo.right, 8) : 8 * sub + halo.left + halo.right;
    p.halo_top = halo.top;
    p.halo_left = halo.left;
    p.box_h = 16 + halo.top + halo.bottom; p.box_n = 1;
    p.tiles_w = W / (8 * sub); p.tiles_h = H / 16;
    p.n_mtiles = B * p.tiles_w * p.tiles_h;
    p.a_tx_bytes = p.box_w * p.box_h * 128;
    p.a_stage_bytes = round_up(p.a_tx_bytes, 1024);
  }
  // per-tap constants for the issue loop (see ConvParams::tap_a)
  p.tap_first_mask = 0;
  for (int e = 0; e < n_entries_total; ++e) {
    const TapEntry& te = p.entries[e];
    const int base = (e / p.n_entries) * p.n_entries;
    bool first = true;
    for (int f = base; f < e; ++f)
      if (p.entries[f].group == te.group) first = false;
    if (first) p.tap_first_mask |= 1u << e;
    p.tap_a[e] = (p.mode == MODE_H) ? (uint32_t)(((te.dy + p.halo_top) * p.box_w + (te.dx + p.halo_left)) * 8) : 0u;
    p.tap_d[e] = (p.mode == MODE_H) ? (uint32_t)(te.group * p.sub * n_tile) : 0u;
  }
  p.n_items = p.n_mtiles * p.n_ntiles * p.n_phase_items;
  // several taps per B stage for small N: keeps the single producer / MMA-issue threads off the critical path
  p.b_group = 1;
  if (p.mode != MODE_T)
    for (int g = p.n_entries; g >= 1; --g)
      if (p.n_entries % g == 0 && g * n_tile * 128 <= 36 * 1024) { p.b_group = g; break; }
  p.b_stage_bytes = p.b_group * n_tile * 128;
  // weight-stationary when the whole layer fits beside the activation ring
  p.b_resident = 0;
  if (p.phase_fixed && p.n_ntiles != 1) return fail("conv: phase-split up-conv needs a single N tile");
  if (p.phase_fixed ||
      (m->b_resident && p.mode != MODE_T && p.n_ntiles == 1 && p.n_phase_items == 1 && p.n_chunks <= kMaxBStages &&
       (long long)p.n_chunks * p.n_entries * n_tile * 128 <= 96 * 1024)) {
    p.b_resident = 1;
    p.b_group = p.n_entries;
    p.b_stage_bytes = p.n_entries * n_tile * 128;
  }
  p.acc_stages = (2 * p.n_groups * p.sub * n_tile <= kTmemCols) ? 2 : 1;
  // fully unrolled issue sequence (conv_tc.cuh:issue_chunk_h) where one is instantiated
  p.fast_id = 0;
  if (p.mode == MODE_H && !m->halo_pad8 && !getenv("DP_NO_FAST_ISSUE") && (op.kind == 3 || op.kind == 4) &&
      p.box_w == 8 * p.sub + 2 && !p.phase_fixed) {
    const int id = op.kind * 100 + p.sub * 10 + p.b_group;
    if (fast_kernel(id)) p.fast_id = id;
  }
  if (p.phase_fixed && p.mode == MODE_H && !m->halo_pad8 && !getenv("DP_NO_FAST_ISSUE") && p.box_w == 8 * p.sub + 2 &&
      fast_kernel(4000 + p.n_groups * 100 + p.sub * 10))
    p.fast_id = 4000 + p.n_groups * 100 + p.sub * 10;

  // shared-memory ring depths
  const int budget = 227 * 1024 - ConvSmemLayout::kBarBytes - 2 * n_tile * n_ntiles * 4 - 2 * p.n_chunks * 64 * 4 - 256 * 4 -
                     ((p.epi_direct || op.head) ? 0 : 4 * kEpiStageBytes) - 1024;
  int a_stages = (p.mode == MODE_H) ? 2 : 4;
  const int b_min = p.b_resident ? p.n_chunks : 2;
  while (a_stages > 1 && a_stages * p.a_stage_bytes + b_min * p.b_stage_bytes > budget) --a_stages;
  int b_stages = (budget - a_stages * p.a_stage_bytes) / p.b_stage_bytes;
  if (b_stages > kMaxBStages) b_stages = kMaxBStages;
  if (p.mode != MODE_H && b_stages > 6) b_stages = 6;
  if (p.b_resident) {
    if (b_stages < p.n_chunks) return fail("conv: resident weights do not fit (%d stages of %d B)", p.n_chunks, p.b_stage_bytes);
    b_stages = p.n_chunks;
  }
  if (b_stages < 2 && !p.b_resident) return fail("conv: shared memory budget exceeded (A %d B %d)", p.a_stage_bytes, p.b_stage_bytes);
  // spend what is left on deeper A rings: the flat modes, and halo mode when the weights are resident (the halo
  // box is ~180 separate 128-byte segments, ~3.5k clk of TMA latency under load: two stages do not cover it)
  if (p.mode == MODE_H && p.b_resident) {
    while (a_stages < 4 && (a_stages + 1) * p.a_stage_bytes + b_stages * p.b_stage_bytes <= budget) ++a_stages;
  }
  if (p.mode != MODE_H) {
    while (a_stages < kMaxAStages && (a_stages + 1) * p.a_stage_bytes + b_stages * p.b_stage_bytes <= budget &&
           a_stages < 6)
      ++a_stages;
  }
  p.a_stages = a_stages;
  p.b_stages = b_stages;
  L.smem = conv_smem_layout(p).total;
  if (L.smem > 227 * 1024) return fail("conv: smem %d too large", L.smem);
  const int rounds = (p.n_items + m->num_sms - 1) / m->num_sms;
  L.grid = (p.n_items + rounds - 1) / rounds;
  if (p.phase_fixed) {
    // item = m_tile * n_phase_items + phase group: a grid that is a multiple of n_phase_items keeps the phase
    // group of a CTA fixed across its grid-stride loop
    L.grid = (m->num_sms / p.n_phase_items) * p.n_phase_items;
    if (L.grid > p.n_items) L.grid = p.n_items;
  }
  // CTA pairs with multicast weights: worthwhile where weights are re-streamed per item (not resident) and there
  // is more than one item per CTA; needs lock-step pairs (even items / M tiles / grid) and 1 KB-aligned halves.
  p.b_pair = 0;
  if (m->b_pair && !p.b_resident && p.mode != MODE_T && p.n_items % 2 == 0 && p.n_mtiles % 2 == 0 && n_tile % 16 == 0 &&
      p.n_items >= 2 * m->num_sms) {
    p.b_pair = 1;
    if (L.grid % 2) L.grid -= 1;
  }

  // executed MACs (tensor-core work actually issued, 16-wide K steps)
  {
    uint64_t ksteps = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      int ks = (op.cin - c * 64 + 15) / 16;
      ksteps += ks > 4 ? 4 : ks;
    }
    L.macs = (uint64_t)m_total * op.cout * ksteps * 16 * (up2 ? 16 : n_entries_total);
  }

  // ---- tensor maps
  const __half* in_base = buf_at(op.in_buf) + op.in_choff;
  const uint64_t cstride = (uint64_t)ib.C * 2;
  if (p.mode == MODE_D) {
    uint64_t dims[2] = {(uint64_t)op.cin, (uint64_t)m_total};
    uint64_t str[1] = {cstride};
    uint32_t box[2] = {64, 128};
    if (make_map(&L.map_a, in_base, 2, dims, str, box)) return 1;
  } else {
    uint64_t dims[4] = {(uint64_t)op.cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {cstride, cstride * W, cstride * W * H};
    uint32_t box[4] = {64, (uint32_t)(p.box_w * stride), (uint32_t)(p.box_h * stride), (uint32_t)p.box_n};
    if (make_map(&L.map_a, in_base, 4, dims, str, box, stride)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)op.cin, (uint64_t)op.cout, (uint64_t)n_entries_total};
    uint64_t str[2] = {(uint64_t)op.cin * 2, (uint64_t)op.cin * 2 * op.cout};
    uint32_t box[3] = {64, (uint32_t)n_tile, (uint32_t)p.b_group};
    if (p.b_pair) { box[1] = (uint32_t)n_tile / 2; box[2] = 1; }
    if (make_map(&L.map_b, q.w, 3, dims, str, box)) return 1;
  }
  return 0;
}

int plan_dense_layer(dp_model* m, const BlobOp& op, int img0, int B, Launch& L) {
  using namespace dp;
  const BlobBuf& ib = m->bufs[op.in_buf];
  const int H = ib.H, W = ib.W;
  if (H % 8 || W % 8) return fail("dense layer: map %dx%d needs H %% 8 == 0 and W %% 8 == 0", H, W);
  if (op.cout != 32 || op.cin % 8) return fail("dense layer: growth must be 32 and Cin a multiple of 8");
  const int mid_buf = op.rsv[0];
  if (mid_buf < 0 || mid_buf >= (int)m->bufs.size() || m->bufs[mid_buf].C != 128 || m->bufs[mid_buf].H != H)
    return fail("dense layer: bad bottleneck buffer");
  auto buf_at = [&](int buf) { return buf_ptr(m, buf, img0); };
  // ---- debug path: the same layer as two naive convs through the bottleneck buffer
  TapEntry t1[kMaxEntries], t3[kMaxEntries];
  int n1 = 0, n3 = 0;
  fill_entries(1, H, W, 1, 1, 1, t1, &n1, nullptr);
  fill_entries(3, H, W, 3, 3, 1, t3, &n3, nullptr);
  NaiveConvParams& a = L.np;
  memset(&a, 0, sizeof a);
  a.n_img = B; a.H = H; a.W = W; a.Cin = op.cin; a.in_ctot = ib.C; a.in_choff = op.in_choff;
  a.OH = H; a.OW = W; a.stride = 1;
  a.Cout = 128; a.out_ctot = 128; a.out_choff = 0; a.n_entries_total = 1; a.n_groups = 1;
  a.relu = 1; a.pro_mode = 2;
  memcpy(a.entries, t1, sizeof t1);
  a.in = buf_at(op.in_buf); a.w = dptr<__half>(m, op.w_off);
  a.epi_shift = dptr<float>(m, op.epi_shift_off);
  a.pro_scale = dptr<float>(m, op.pro_scale_off); a.pro_shift = dptr<float>(m, op.pro_shift_off);
  a.out = buf_at(mid_buf);
  NaiveConvParams& b = L.np2;
  memset(&b, 0, sizeof b);
  b.n_img = B; b.H = H; b.W = W; b.Cin = 128; b.in_ctot = 128; b.in_choff = 0;
  b.OH = H; b.OW = W; b.stride = 1;
  b.Cout = 32; b.out_ctot = ib.C; b.out_choff = op.out_choff; b.n_entries_total = 9; b.n_groups = 1;
  memcpy(b.entries, t3, sizeof t3);
  b.in = buf_at(mid_buf); b.w = dptr<__half>(m, op.rsv64[0]); b.out = buf_at(op.in_buf);
  if (m->precision) {   // fp32 modes: the two convs above through the bottleneck buffer
    L.macs = (uint64_t)B * H * W * ((uint64_t)op.cin * 128 + 9ull * 128 * 32);
    return plan_tx(m, L.np, L.tx[0]) || plan_tx(m, L.np2, L.tx[1]);
  }

  // ---- fused tensor-core plan
  DenseLayerParams& p = L.dl;
  memset(&p, 0, sizeof p);
  p.n_img = B; p.H = H; p.W = W; p.C = op.cin;
  p.n_chunks = (op.cin + 63) / 64;
  p.rh = 16;
  {
    // 8-row regions when 16-row regions would leave most SMs idle.  With cross-layer overlap a half-empty GPU is
    // not wasted (the next layer's early chunks run on the free SMs), so the threshold is a tunable.
    const char* env = getenv("DP_DL_MIN_ITEMS16");
    const long long min_items = env ? atoll(env) : m->num_sms;
    if (H % 16 || (long long)B * (W / 8) * (H / 16) < min_items) p.rh = 8;
  }
  p.tiles_w = W / 8; p.tiles_h = (H + p.rh - 1) / p.rh;
  p.n_items = B * p.tiles_w * p.tiles_h;
  p.out_ctot = ib.C; p.out_choff = op.out_choff;
  p.n_safe_chunks = m->use_overlap ? op.rsv[1] / 64 : 0;   // rsv[1] = channels older than the preceding kernel's output
  if (p.n_safe_chunks > p.n_chunks) p.n_safe_chunks = p.n_chunks;
  p.pro_scale = a.pro_scale; p.pro_shift = a.pro_shift; p.mid_shift = a.epi_shift;
  p.out = buf_at(op.in_buf);
  const int budget = 227 * 1024 - DenseLayerSmem::kBarBytes - dl_t_bytes(p.rh) - 2 * p.n_chunks * 64 * 4 - 128 * 4 - 1024 -
                     10 * 1024;   // the last A stage's M-block over-read must stay inside the allocation
  const int a_stage = dl_a_stage(p.rh);
  // Activation ring as deep as shared memory allows (up to 8): a halo chunk is 100-180 separate 128-byte segments,
  // ~3.5k clk of TMA latency under load, and phase 1 of the small-map layers consumed one chunk per ~950 clk with
  // 4 stages in flight -- TMA-latency bound, not transform or MMA bound (halving the transform math or dropping
  // its proxy fence changed nothing).
  p.a_stages = kMaxAStages; p.b_stages = 4;
  {
    const char* env = getenv("DP_DL_A_STAGES");
    if (env && atoi(env) >= 2 && atoi(env) <= kMaxAStages) p.a_stages = atoi(env);
    env = getenv("DP_DL_B_STAGES");
    if (env && atoi(env) >= 2 && atoi(env) <= kMaxBStages) p.b_stages = atoi(env);
  }
  while (p.a_stages * a_stage + p.b_stages * kDlBStage > budget && p.a_stages > 2) --p.a_stages;
  while (p.a_stages * a_stage + p.b_stages * kDlBStage > budget && p.b_stages > 2) --p.b_stages;
  if (p.a_stages * a_stage + p.b_stages * kDlBStage > budget) return fail("dense layer: shared memory budget exceeded");
  L.smem = dense_layer_smem(p).total;
  const int rounds = (p.n_items + m->num_sms - 1) / m->num_sms;
  L.grid = (p.n_items + rounds - 1) / rounds;
  {
    uint64_t ksteps = 0;
    for (int c = 0; c < p.n_chunks; ++c) {
      int ks = (op.cin - c * 64 + 15) / 16;
      ksteps += ks > 4 ? 4 : ks;
    }
    // executed: 1x1 on 256 rows per 128-pixel region (halo recompute + padding rows), 3x3 on 128 rows
    L.macs = (uint64_t)p.n_items * ((dl_rows(p.rh) > 128 ? 256ull : 128ull) * 128 * ksteps * 16 +
                                    128ull * 32 * 9 * 128);
  }
  {
    uint64_t dims[4] = {(uint64_t)op.cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint64_t cs = (uint64_t)ib.C * 2;
    uint64_t str[3] = {cs, cs * W, cs * W * H};
    uint32_t box[4] = {64, (uint32_t)kDlHaloW, (uint32_t)(p.rh + 2), 1};
    if (make_map(&L.map_a, buf_at(op.in_buf) + op.in_choff, 4, dims, str, box)) return 1;
  }
  {
    uint64_t dims[3] = {(uint64_t)op.cin, 128, 1};
    uint64_t str[2] = {(uint64_t)op.cin * 2, (uint64_t)op.cin * 2 * 128};
    uint32_t box[3] = {64, 128, 1};
    if (make_map(&L.map_b, a.w, 3, dims, str, box)) return 1;
  }
  {
    uint64_t dims[3] = {128, 32, 9};
    uint64_t str[2] = {128 * 2, 128 * 2 * 32};
    uint32_t box[3] = {64, 32, (uint32_t)kDlW2Group};
    if (make_map(&L.map_w2, b.w, 3, dims, str, box)) return 1;
  }
  return 0;
}

// Groups runs of consecutive fused dense layers that (a) extend the same concat buffer layer by layer, (b) use 8-row
// regions, (c) have 1, 2, 4 or 8 regions per image (one thread-block cluster) and (d) start at a channel count that is
// a multiple of 64 (the kernel's tail tiles) into one persistent kernel launch each.
int plan_dense_blocks(dp_model* m, SubPlan& sp) {
  using namespace dp;
  const int n = (int)m->ops.size();
  struct Run { int first, len; };
  std::vector<Run> runs;
  for (int i = 0; i < n;) {
    const int per_img = (m->ops[i].type == OP_DENSE_LAYER) ? sp.launches[i].dl.tiles_w * sp.launches[i].dl.tiles_h : 0;
    // (e) only where two consecutive per-layer kernels cannot be resident together (2 x regions > SMs): there the
    //     per-layer path loses its cross-layer overlap (measured: 16x16 maps at batch 32, 268 -> 189 us); where they
    //     can (8x8 maps: 32 regions) the per-layer kernels are as fast or faster (138 vs 150 us) and stay.
    const bool crowded = (m->ops[i].type == OP_DENSE_LAYER) && (2 * sp.launches[i].dl.n_items > m->num_sms || m->dense_block > 1);
    if (m->ops[i].type != OP_DENSE_LAYER || sp.launches[i].dl.rh != 8 || (m->ops[i].cin % 64) != 0 || !crowded ||
        !(per_img == 1 || per_img == 2 || per_img == 4 || per_img == 8)) { ++i; continue; }
    int j = i + 1;
    while (j < n && m->ops[j].type == OP_DENSE_LAYER && m->ops[j].in_buf == m->ops[i].in_buf &&
           m->ops[j].in_choff == m->ops[i].in_choff && m->ops[j].cin == m->ops[j - 1].cin + 32 &&
           m->ops[j].out_choff == m->ops[j].in_choff + m->ops[j].cin && sp.launches[j].dl.rh == 8 &&
           sp.launches[j].dl.n_items == sp.launches[i].dl.n_items)
      ++j;
    if (j - i >= 2) runs.push_back({i, j - i});
    i = j;
  }
  if (runs.empty()) return 0;
  for (const Run& r : runs) {
    Launch& L0 = sp.launches[r.first];
    const BlobOp& op0 = m->ops[r.first];
    const BlobOp& opl = m->ops[r.first + r.len - 1];
    const BlobBuf& ib = m->bufs[op0.in_buf];
    std::vector<DenseBlockLayer> tab(r.len);
    int a_stages = kMaxAStages, b_stages = kMaxBStages;
    for (int l = 0; l < r.len; ++l) {
      const Launch& Ll = sp.launches[r.first + l];
      memset(&tab[l], 0, sizeof(DenseBlockLayer));
      tab[l].map_w1 = Ll.map_b;
      tab[l].map_w2 = Ll.map_w2;
      tab[l].pro_scale = Ll.dl.pro_scale;
      tab[l].pro_shift = Ll.dl.pro_shift;
      tab[l].mid_shift = Ll.dl.mid_shift;
      tab[l].C = Ll.dl.C;
      tab[l].n_chunks = Ll.dl.n_chunks;
      tab[l].out_choff = Ll.dl.out_choff;
      if (Ll.dl.a_stages < a_stages) a_stages = Ll.dl.a_stages;
      if (Ll.dl.b_stages < b_stages) b_stages = Ll.dl.b_stages;
    }
    DenseBlockLayer* tab_dev = nullptr;
    CU_OK(cudaMalloc(&tab_dev, tab.size() * sizeof(DenseBlockLayer)));
    sp.dev_allocs.emplace_back(tab_dev, [](void* q) { cudaFree(q); });
    CU_OK(cudaMemcpy(tab_dev, tab.data(), tab.size() * sizeof(DenseBlockLayer), cudaMemcpyHostToDevice));
    DenseBlockParams& p = L0.db;
    memset(&p, 0, sizeof p);
    p.n_img = L0.dl.n_img; p.H = L0.dl.H; p.W = L0.dl.W;
    p.tiles_w = L0.dl.tiles_w; p.tiles_h = L0.dl.tiles_h; p.n_items = L0.dl.n_items;
    p.n_layers = r.len;
    {
      // shared memory: barriers + BN2 staging + T (2 buffers) + 2 tail tiles + rings; the last A stage's M-block
      // over-read must stay inside the allocation (10 KB of slack, as in the per-layer kernel)
      const int fixed = DenseBlockSmem::kBarBytes + DenseBlockSmem::kMidBytes + dl_t_bytes(8) / 2 + 2 * dl_a_stage(8) + 1024 + 10 * 1024;
      while (a_stages > 2 && fixed + a_stages * dl_a_stage(8) + b_stages * kDlBStage > 227 * 1024) --a_stages;
      while (b_stages > 2 && fixed + a_stages * dl_a_stage(8) + b_stages * kDlBStage > 227 * 1024) --b_stages;
      if (fixed + a_stages * dl_a_stage(8) + b_stages * kDlBStage > 227 * 1024) return fail("dense block: shared memory budget exceeded");
    }
    p.a_stages = a_stages; p.b_stages = b_stages;
    p.out_ctot = L0.dl.out_ctot;
    p.out = L0.dl.out;
    p.layers = tab_dev;
    p.cluster_size = p.tiles_w * p.tiles_h;
    {
      // one activation map for the whole block: channels beyond a layer's C (later layers' slots) may hold values of
      // an earlier forward, but the K-steps issued per layer stop at C, so they never reach the tensor cores
      uint64_t dims[4] = {(uint64_t)opl.cin, (uint64_t)ib.W, (uint64_t)ib.H, (uint64_t)sp.n_img};
      const uint64_t cs = (uint64_t)ib.C * 2;
      uint64_t str[3] = {cs, cs * ib.W, cs * ib.W * ib.H};
      uint32_t box[4] = {64, (uint32_t)kDlHaloW, 10, 1};
      if (make_map(&L0.map_x_block, buf_ptr(m, op0.in_buf, sp.img0) + op0.in_choff, 4, dims, str, box)) return 1;
    }
    L0.block_smem = dense_block_smem(p).total;
    L0.block_len = r.len;
    for (int l = 1; l < r.len; ++l) sp.launches[r.first + l].block_member = true;
  }
  return 0;
}

int get_plan(dp_model* m, int B, int split, Plan** out) {
  std::lock_guard<std::mutex> lk(m->mu);
  if (split < 1 || B % split || B / split < 1) split = 1;
  auto key = std::make_pair(B, split);
  auto it = m->plans.find(key);
  if (it != m->plans.end()) {
    *out = &it->second;
    return 0;
  }
  Plan plan;
  plan.subs.resize(split);
  const int bs = B / split;
  for (int s = 0; s < split; ++s) {
    SubPlan& sp = plan.subs[s];
    sp.img0 = s * bs;
    sp.n_img = bs;
    sp.launches.resize(m->ops.size());
    for (size_t i = 0; i < m->ops.size(); ++i) {
      sp.launches[i].type = m->ops[i].type;
      int prc = 0;
      if (m->ops[i].type == OP_CONV) prc = plan_conv(m, m->ops[i], sp.img0, bs, sp.launches[i]);
      if (m->ops[i].type == OP_DENSE_LAYER) prc = plan_dense_layer(m, m->ops[i], sp.img0, bs, sp.launches[i]);
      if (prc) {
        g_err = "op " + std::to_string(i) + ": " + g_err;
        return 1;
      }
    }
  }
  if (m->dense_block && !m->precision)
    for (SubPlan& sp : plan.subs)
      if (plan_dense_blocks(m, sp)) return 1;
  auto res = m->plans.emplace(key, std::move(plan));
  *out = &res.first->second;
  return 0;
}

int grid_for(long long total, int threads) {
  long long g = (total + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// fp32 precision mode: every op of the program on fp32 buffers (aux.cuh templates instantiated for float,
// conv_f32_kernel for the convs; a fused dense layer runs as its two convs through the bottleneck buffer).
int run_op_f32(dp_model* m, SubPlan* sp, int i, cudaStream_t st) {
  const BlobOp& op = m->ops[i];
  Launch& L = sp->launches[i];
  const int B = sp->n_img, img0 = sp->img0;
  auto fb = [&](int buf) { return reinterpret_cast<float*>(buf_ptr(m, buf, img0)); };
  const BlobBuf& ib = m->bufs[op.in_buf];
  const BlobBuf& ob = m->bufs[op.out_buf];
  const bool pdl = m->use_pdl != 0;
  auto conv = [&](const dp::NaiveConvParams& q, const Launch::TxPlan& t) {
    const long long M = (long long)q.n_img * q.OH * q.OW;
    if (m->precision == 2) {   // 3xTF32 on the tensor cores (precise_tc.cuh): N tiles of <= 128 couts, multiples of 16
      if (t.n_tile && t.kind == 2) {   // 1x1, stride 1: flat [pixels, C] TMA tiles, prologue + split in place
        const int n_mtiles = (int)((M + 127) / 128);
        int gx = (m->num_sms + t.n_ntiles - 1) / t.n_ntiles;
        if (gx > n_mtiles) gx = n_mtiles;
        dim3 ogrid((unsigned)gx, (unsigned)t.n_ntiles, 1);
        dp::conv_1x1_tf32x3_kernel<<<ogrid, dp::kThThreads, dp::t1_smem_bytes(t.n_tile), st>>>(t.a, t.wh, t.wl, q, t.n_tile);
        return;
      }
      if (t.n_tile) {          // 3x3 / up2 without prologue: TMA halo tile per channel slice, taps as descriptor offsets
        // persistent CTAs: about one per SM in total, each walking its share of the 16 x 8 regions
        const int n_items = q.n_img * (q.H / 16) * (q.W / 8), per_x = t.n_ntiles * q.n_groups;
        int gx = (m->num_sms + per_x - 1) / per_x;
        if (gx > n_items) gx = n_items;
        if (getenv("DP_TX_NO_PERSIST")) gx = n_items;
        dim3 hgrid((unsigned)gx, (unsigned)t.n_ntiles, (unsigned)q.n_groups);
        // weight stages of one accumulator chunk (3 taps; 2 for up2) where two of them fit beside the halo tiles
        const int epg = q.n_entries_total / q.n_groups;
        const int chunk_taps = epg % 3 == 0 ? 3 : (epg % 2 == 0 ? 2 : 1);
        const int tps = (t.n_tile <= 64 && !getenv("DP_TX_TPS1")) ? chunk_taps : 1;
        dp::conv_halo_tf32x3_kernel<<<hgrid, dp::kThHaloThreads, dp::th_smem_bytes(t.n_tile, tps), st>>>(t.a, t.wh, t.wl, q, t.n_tile, tps);
        return;
      }
      const int n_tile = round_up((q.Cout + t.n_ntiles - 1) / t.n_ntiles, 16);
      dim3 grid((unsigned)((M + 127) / 128), (unsigned)t.n_ntiles, (unsigned)q.n_groups);
      dp::conv_tf32x3_kernel<<<grid, dp::kTxThreads, dp::tx_smem_bytes(n_tile), st>>>(q, n_tile);
      return;
    }
    if (q.Cout <= 32) {
      dim3 grid((unsigned)((M + dp::kPcBM - 1) / dp::kPcBM), (unsigned)((q.Cout + 31) / 32), (unsigned)q.n_groups);
      dp::conv_f32_kernel<32><<<grid, dp::kPcThreads, 0, st>>>(q);
    } else {
      dim3 grid((unsigned)((M + dp::kPcBM - 1) / dp::kPcBM), (unsigned)((q.Cout + 63) / 64), (unsigned)q.n_groups);
      dp::conv_f32_kernel<64><<<grid, dp::kPcThreads, 0, st>>>(q);
    }
  };
  cudaError_t le = cudaSuccess;
  switch (op.type) {
    case OP_STEM_S2D: {
      if (ob.C != 64 || ob.H != m->patch / 2) return fail("stem s2d buffer must be [P/2][P/2][64]");
      const long long total = (long long)B * ob.H * ob.W * 4;
      le = launch_pdl(dp::stem_s2d_kernel<float>, grid_for(total, 256), 256, st, pdl, m->pass_dev, img0, B, m->patch,
                      fb(op.out_buf));
      break;
    }
    case OP_MAXPOOL: {
      const long long total = (long long)B * (ib.H / 2) * (ib.W / 2) * (op.cin / 8);
      le = launch_pdl(dp::maxpool3s2_kernel<float>, grid_for(total, 256), 256, st, pdl, fb(op.in_buf), ib.C, op.in_choff,
                      fb(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, op.cin, op.pool);
      break;
    }
    case OP_DWCONV: {
      const int stride = op.rsv[0] & 0xff, rate = (op.rsv[0] >> 8) & 0xff;
      if ((stride != 1 && stride != 2) || rate < 1 || ib.H % stride || ob.H != ib.H / stride || op.cin % 8)
        return fail("depthwise conv: bad geometry (stride %d rate %d map %d -> %d)", stride, rate, ib.H, ob.H);
      const long long total = (long long)B * ob.H * ob.W * (op.cin / 8);
      le = launch_pdl(dp::dwconv3x3_kernel<float>, grid_for(total, 256), 256, st, pdl, fb(op.in_buf), ib.C, op.in_choff,
                      fb(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, op.cin, stride, rate, dptr<float>(m, op.w_off),
                      dptr<float>(m, op.epi_shift_off), op.pro, op.relu);
      break;
    }
    case OP_GAP:
    case OP_BCAST: {
      const BlobBuf& small = (op.type == OP_GAP) ? ob : ib;
      const BlobBuf& big = (op.type == OP_GAP) ? ib : ob;
      if (small.H != 1 || small.W != 1) return fail("global pool / broadcast needs a 1x1 map on one side");
      if (op.type == OP_GAP)
        le = launch_pdl(dp::global_avgpool_kernel<float>, grid_for((long long)B * (op.cin / 8), 128), 128, st, pdl,
                        fb(op.in_buf), ib.C, op.in_choff, fb(op.out_buf), ob.C, op.out_choff, B, big.H * big.W, op.cin);
      else
        le = launch_pdl(dp::broadcast_kernel<float>, grid_for((long long)B * big.H * big.W * (op.cin / 8), 256), 256, st,
                        pdl, fb(op.in_buf), ib.C, op.in_choff, fb(op.out_buf), ob.C, op.out_choff, B, big.H * big.W, op.cin);
      break;
    }
    case OP_RESIZE: {
      const long long total = (long long)B * ob.H * ob.W * (op.cin / 8);
      le = launch_pdl(dp::resize_bilinear_kernel<float>, grid_for(total, 256), 256, st, pdl, fb(op.in_buf), ib.C,
                      op.in_choff, fb(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, ob.H, ob.W, op.cin);
      break;
    }
    case OP_HEAD_DOT: {
      if (ob.C != 8 || ob.H != ib.H || op.cin % 8) return fail("head dot: output buffer must be [H][W][8]");
      const long long n_pix = (long long)B * ib.H * ib.W;
      le = launch_pdl(dp::head_dot_kernel<float>, grid_for(n_pix * 32, 256), 256, st, pdl, fb(op.in_buf), ib.C, op.in_choff,
                      op.cin, n_pix, dptr<float>(m, op.head_w_off), op.head_b, fb(op.out_buf), 8);
      break;
    }
    case OP_HEAD_RESIZE: {
      if (ib.C != 8) return fail("head resize: input buffer must be [H][W][8]");
      const long long total = (long long)B * m->patch * m->patch;
      le = launch_pdl(dp::head_resize_kernel, grid_for(total, 256), 256, st, pdl, (const float*)fb(op.in_buf), 8, B, ib.H,
                      ib.W, m->patch, m->pass_dev, img0);
      break;
    }
    case OP_AVGPOOL3: {
      const long long total = (long long)B * ib.H * ib.W * (op.cin / 8);
      le = launch_pdl(dp::avgpool3s1_kernel<float>, grid_for(total, 256), 256, st, pdl, fb(op.in_buf), ib.C, op.in_choff,
                      fb(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, op.cin);
      break;
    }
    case OP_BNPOOL: {
      const int OH = op.pool ? ib.H / 2 : ib.H, OW = op.pool ? ib.W / 2 : ib.W;
      const long long total = (long long)B * OH * OW * (op.cin / 8);
      le = launch_pdl(dp::bn_act_pool_kernel<float>, grid_for(total, 256), 256, st, pdl, fb(op.in_buf), ib.C, op.in_choff,
                      fb(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, op.cin, dptr<float>(m, op.epi_scale_off),
                      dptr<float>(m, op.epi_shift_off), op.relu, op.pool);
      break;
    }
    case OP_CONV: {
      conv(L.np, L.tx[0]);
      LAUNCH_OK();
      if (op.head) {
        const long long tot = (long long)B * m->patch * m->patch;
        dp::head_naive_kernel<float><<<grid_for(tot, 256), 256, 0, st>>>(
            reinterpret_cast<const float*>(L.np.out), op.cout, 0, op.cout, B, m->patch, dptr<float>(m, op.head_w_off),
            op.head_b, m->pass_dev, img0);
      }
      break;
    }
    case OP_DENSE_LAYER: {
      conv(L.np, L.tx[0]);
      LAUNCH_OK();
      conv(L.np2, L.tx[1]);
      break;
    }
    default:
      return fail("op type %d has no fp32 kernel", op.type);
  }
  if (le != cudaSuccess) return fail("fp32 op launch failed: %s", cudaGetErrorString(le));
  LAUNCH_OK();
  return 0;
}

int run_op(dp_model* m, SubPlan* sp, int i, cudaStream_t st, bool whole_program = false) {
  if (m->precision) return run_op_f32(m, sp, i, st);
  const BlobOp& op = m->ops[i];
  Launch& L = sp->launches[i];
  const int B = sp->n_img, img0 = sp->img0;
  auto buf_at = [&](int buf) { return buf_ptr(m, buf, img0); };
  switch (op.type) {
    case OP_STEM_IM2COL: {
      const BlobBuf& ob = m->bufs[op.out_buf];
      if (ob.C != 160 || ob.H != m->patch / 2) return fail("stem im2col buffer must be [P/2][P/2][160]");
      const long long total = (long long)B * ob.H * ob.W * 20;
      dp::stem_im2col_kernel<<<grid_for(total, 256), 256, 0, st>>>(m->pass_dev, img0, B, m->patch,
                                                                   buf_at(op.out_buf));
      LAUNCH_OK();
      return 0;
    }
    case OP_STEM_S2D: {
      const BlobBuf& ob = m->bufs[op.out_buf];
      if (ob.C != 64 || ob.H != m->patch / 2) return fail("stem s2d buffer must be [P/2][P/2][64]");
      const long long total = (long long)B * ob.H * ob.W * 4;
      cudaError_t le = launch_pdl(dp::stem_s2d_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0, m->pass_dev, img0, B,
                                  m->patch, buf_at(op.out_buf));
      if (le != cudaSuccess) return fail("stem launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_MAXPOOL: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const long long total = (long long)B * (ib.H / 2) * (ib.W / 2) * (op.cin / 8);
      cudaError_t le = launch_pdl(dp::maxpool3s2_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0,
                                  buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, ib.H,
                                  ib.W, op.cin, op.pool);
      if (le != cudaSuccess) return fail("maxpool launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_DWCONV: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const int stride = op.rsv[0] & 0xff, rate = (op.rsv[0] >> 8) & 0xff;
      if ((stride != 1 && stride != 2) || rate < 1 || ib.H % stride || ob.H != ib.H / stride || op.cin % 8)
        return fail("depthwise conv: bad geometry (stride %d rate %d map %d -> %d)", stride, rate, ib.H, ob.H);
      const long long total = (long long)B * ob.H * ob.W * (op.cin / 8);
      cudaError_t le = launch_pdl(dp::dwconv3x3_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0, buf_at(op.in_buf),
                                  ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W, op.cin, stride,
                                  rate, dptr<__half>(m, op.w_off), dptr<float>(m, op.epi_shift_off), op.pro, op.relu);
      if (le != cudaSuccess) return fail("depthwise conv launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_GAP:
    case OP_BCAST: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const BlobBuf& small = (op.type == OP_GAP) ? ob : ib;
      const BlobBuf& big = (op.type == OP_GAP) ? ib : ob;
      if (small.H != 1 || small.W != 1) return fail("global pool / broadcast needs a 1x1 map on one side");
      cudaError_t le;
      if (op.type == OP_GAP)
        le = launch_pdl(dp::global_avgpool_kernel<__half>, grid_for((long long)B * (op.cin / 8), 128), 128, st, m->use_pdl != 0,
                        buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, big.H * big.W, op.cin);
      else
        le = launch_pdl(dp::broadcast_kernel<__half>, grid_for((long long)B * big.H * big.W * (op.cin / 8), 256), 256, st,
                        m->use_pdl != 0, buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B,
                        big.H * big.W, op.cin);
      if (le != cudaSuccess) return fail("pool/broadcast launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_RESIZE: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const long long total = (long long)B * ob.H * ob.W * (op.cin / 8);
      cudaError_t le = launch_pdl(dp::resize_bilinear_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0,
                                  buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, ib.H, ib.W,
                                  ob.H, ob.W, op.cin);
      if (le != cudaSuccess) return fail("resize launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_HEAD_DOT: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      if (ob.C != 8 || ob.H != ib.H || op.cin % 8) return fail("head dot: output buffer must be [H][W][8] (one fp32 per pixel)");
      const long long n_pix = (long long)B * ib.H * ib.W;
      cudaError_t le = launch_pdl(dp::head_dot_kernel<__half>, grid_for(n_pix * 32, 256), 256, st, m->use_pdl != 0, buf_at(op.in_buf),
                                  ib.C, op.in_choff, op.cin, n_pix, dptr<float>(m, op.head_w_off), op.head_b,
                                  reinterpret_cast<float*>(buf_at(op.out_buf)), 4);
      if (le != cudaSuccess) return fail("head dot launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_HEAD_RESIZE: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      if (ib.C != 8) return fail("head resize: input buffer must be [H][W][8] (one fp32 per pixel)");
      const long long total = (long long)B * m->patch * m->patch;
      cudaError_t le = launch_pdl(dp::head_resize_kernel, grid_for(total, 256), 256, st, m->use_pdl != 0,
                                  reinterpret_cast<const float*>(buf_at(op.in_buf)), 4, B, ib.H, ib.W, m->patch, m->pass_dev,
                                  img0);
      if (le != cudaSuccess) return fail("head resize launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_AVGPOOL3: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const long long total = (long long)B * ib.H * ib.W * (op.cin / 8);
      cudaError_t le = launch_pdl(dp::avgpool3s1_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0,
                                  buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, ib.H,
                                  ib.W, op.cin);
      if (le != cudaSuccess) return fail("avgpool launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_BNPOOL: {
      const BlobBuf& ib = m->bufs[op.in_buf];
      const BlobBuf& ob = m->bufs[op.out_buf];
      const int OH = op.pool ? ib.H / 2 : ib.H, OW = op.pool ? ib.W / 2 : ib.W;
      const long long total = (long long)B * OH * OW * (op.cin / 8);
      cudaError_t le = launch_pdl(dp::bn_act_pool_kernel<__half>, grid_for(total, 256), 256, st, m->use_pdl != 0,
                                  buf_at(op.in_buf), ib.C, op.in_choff, buf_at(op.out_buf), ob.C, op.out_choff, B, ib.H,
                                  ib.W, op.cin, dptr<float>(m, op.epi_scale_off), dptr<float>(m, op.epi_shift_off),
                                  op.relu, op.pool);
      if (le != cudaSuccess) return fail("bn/pool launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_CONV: {
      if (m->naive_conv) {
        const long long total = (long long)B * L.np.OH * L.np.OW * L.np.n_groups * L.np.Cout;
        dp::conv_naive_kernel<<<grid_for(total, 256), 256, 0, st>>>(L.np);
        LAUNCH_OK();
        if (op.head) {
          const long long tot = (long long)B * m->patch * m->patch;
          dp::head_naive_kernel<__half><<<grid_for(tot, 256), 256, 0, st>>>(L.np.out, op.cout, 0, op.cout, B, m->patch,
                                                                    dptr<float>(m, op.head_w_off), op.head_b,
                                                                    m->pass_dev, img0);
          LAUNCH_OK();
        }
        return 0;
      }
      dp::ConvParams cp = L.cp;
      cp.desc_base_mode = m->desc_base_mode;

      cp.trace = (m->trace_op == i) ? m->trace_dev : nullptr;
#ifdef DP_EXPERIMENTS   // work-skipping timing switches (wrong results): experiment builds only, never in the shipped library
      { static const int dbg = getenv("DP_DBG_SKIP") ? atoi(getenv("DP_DBG_SKIP")) : 0; cp.dbg_skip = dbg; }
#else
      cp.dbg_skip = 0;
#endif
      cp.gt = (m->stamp && m->gt_dev) ? m->gt_dev + 2 * i : nullptr;
      // Programmatic dependent launch: the kernel's setup (barrier init, TMEM alloc, BN constants -> smem)
      // runs before its griddepcontrol.wait and so overlaps the tail of the preceding kernel in the stream.
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(L.grid);
      cfg.blockDim = dim3(cp.residual ? dp::ConvKernelShape<false, true>::kThreads
                                      : (L.prologue ? dp::ConvKernelShape<true, false>::kThreads
                                                    : dp::ConvKernelShape<false, false>::kThreads));
      cfg.dynamicSmemBytes = L.smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = m->use_pdl ? 1 : 0;
      if (cp.b_pair) {
        attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
        attr[cfg.numAttrs].val.clusterDim.x = 2;
        attr[cfg.numAttrs].val.clusterDim.y = 1;
        attr[cfg.numAttrs].val.clusterDim.z = 1;
        ++cfg.numAttrs;
      }
      cudaError_t le;
      switch (cp.mode) {
        case dp::MODE_D:
          if (cp.residual) le = cudaLaunchKernelEx(&cfg, dp::conv_tc_kernel<dp::MODE_D, false, true>, L.map_a, L.map_b, cp);
          else if (L.prologue) le = cudaLaunchKernelEx(&cfg, dp::conv_tc_kernel<dp::MODE_D, true>, L.map_a, L.map_b, cp);
          else le = cudaLaunchKernelEx(&cfg, dp::conv_tc_kernel<dp::MODE_D, false>, L.map_a, L.map_b, cp);
          break;
        case dp::MODE_T:
          le = cudaLaunchKernelEx(&cfg, dp::conv_tc_kernel<dp::MODE_T, false>, L.map_a, L.map_b, cp);
          break;
        default: {
          ConvKernelH fn = fast_kernel(cp.fast_id);
          if (!fn) { fn = fast_kernel(0); cp.fast_id = 0; }
          le = cudaLaunchKernelEx(&cfg, fn, L.map_a, L.map_b, cp);
          break;
        }
      }
      if (le != cudaSuccess) return fail("conv launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    case OP_DENSE_LAYER: {
      if (m->naive_conv) {
        long long total = (long long)B * L.np.H * L.np.W * L.np.Cout;
        dp::conv_naive_kernel<<<grid_for(total, 256), 256, 0, st>>>(L.np);
        LAUNCH_OK();
        total = (long long)B * L.np2.H * L.np2.W * L.np2.Cout;
        dp::conv_naive_kernel<<<grid_for(total, 256), 256, 0, st>>>(L.np2);
        LAUNCH_OK();
        return 0;
      }
      // whole-program runs execute a run of small-map dense layers as one persistent kernel (debug options that
      // inspect single layers keep the per-layer kernels)
      const bool use_block = whole_program && m->dense_block && m->trace_op < 0 && !m->stamp_ctas;
      if (use_block && L.block_member) return 0;
      if (use_block && L.block_len > 0) {
        dp::DenseBlockParams db = L.db;
        db.gt_layers = (m->stamp && m->gt_dev) ? m->gt_dev + 2 * i : nullptr;
        db.trace = (m->trace_block == i) ? m->trace_dev : nullptr;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof cfg);
        cfg.gridDim = dim3(db.n_items);
        cfg.blockDim = dim3(640);
        cfg.dynamicSmemBytes = L.block_smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = m->use_pdl ? 1 : 0;
        if (db.cluster_size > 1) {
          attr[cfg.numAttrs].id = cudaLaunchAttributeClusterDimension;
          attr[cfg.numAttrs].val.clusterDim.x = (unsigned)db.cluster_size;
          attr[cfg.numAttrs].val.clusterDim.y = 1;
          attr[cfg.numAttrs].val.clusterDim.z = 1;
          ++cfg.numAttrs;
        }
        cudaError_t le = cudaLaunchKernelEx(&cfg, dp::dense_block_kernel, L.map_x_block, db);
        if (le != cudaSuccess) return fail("dense block launch failed: %s", cudaGetErrorString(le));
        LAUNCH_OK();
        return 0;
      }
      dp::DenseLayerParams dl = L.dl;
      dl.trace = (m->trace_op == i) ? m->trace_dev : nullptr;
      dl.gt = (m->stamp && m->gt_dev) ? m->gt_dev + 2 * i : nullptr;
      dl.gt_all = (m->stamp_ctas && m->gt_all_dev && L.grid <= 256) ? m->gt_all_dev + (size_t)i * 1024 : nullptr;
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof cfg);
      cfg.gridDim = dim3(L.grid);
      cfg.blockDim = dim3(640);
      cfg.dynamicSmemBytes = L.smem;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = m->use_pdl ? 1 : 0;
      cudaError_t le;
      le = (dl.rh == 16) ? cudaLaunchKernelEx(&cfg, dp::dense_layer_kernel<16>, L.map_a, L.map_b, L.map_w2, dl)
                         : cudaLaunchKernelEx(&cfg, dp::dense_layer_kernel<8>, L.map_a, L.map_b, L.map_w2, dl);
      if (le != cudaSuccess) return fail("dense layer launch failed: %s", cudaGetErrorString(le));
      LAUNCH_OK();
      return 0;
    }
    default:
      return fail("unknown op type %d", op.type);
  }
}

int check_model(const dp_model* m) {
  if (!m) return fail("null model");
  return 0;
}

}  // namespace

// Everything a model owns besides the container's data section: activation buffers, the head scratch, the per-call
// argument record, kernel attributes and the environment overrides.  Shared by dp_model_create and dp_model_clone.
static int alloc_lane_state(dp_model* m) {
  cudaError_t e;
  const uint32_t n_bufs = (uint32_t)m->bufs.size();
  const int max_batch = m->max_batch;
  const char* env = nullptr;
  m->buf_dev.assign(n_bufs, nullptr);
  for (uint32_t i = 0; i < n_bufs; ++i) {
    const BlobBuf& b = m->bufs[i];
    if (b.C % 8) { return fail("buffer %u: channel count %d not a multiple of 8", i, b.C); }
    const size_t sz = (size_t)max_batch * b.H * b.W * b.C * m->esize;
    e = cudaMalloc(&m->buf_dev[i], sz);
    if (e != cudaSuccess) { return fail("cudaMalloc buffer %u (%zu B): %s", i, sz, cudaGetErrorString(e)); }
    cudaMemset(m->buf_dev[i], 0, sz);
    m->device_bytes += sz;
  }
  {
    size_t sz = 0;
    for (const BlobOp& op : m->ops)
      if (op.type == OP_CONV && op.head) sz = (size_t)max_batch * m->patch * m->patch * op.cout * m->esize;
    if (sz) {
      e = cudaMalloc(&m->scratch_head, sz);
      if (e != cudaSuccess) { return fail("cudaMalloc head scratch: %s", cudaGetErrorString(e)); }
      m->device_bytes += sz;
    }
  }
  e = cudaMalloc(&m->pass_dev, sizeof(dp::PassDesc));
  if (e != cudaSuccess) { return fail("cudaMalloc pass descriptor: %s", cudaGetErrorString(e)); }
  cudaMemset(m->pass_dev, 0, sizeof(dp::PassDesc));
  env = getenv("DP_SPLIT");
  if (env && atoi(env) > 0) m->split = atoi(env);
  {
    const int kMaxSmem = 227 * 1024;
    cudaError_t e1 = cudaFuncSetAttribute(dp::conv_tc_kernel<dp::MODE_D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaError_t e2 = cudaFuncSetAttribute(dp::conv_tc_kernel<dp::MODE_D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaError_t e3 = cudaFuncSetAttribute(dp::conv_tc_kernel<dp::MODE_T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaError_t e4 = cudaSuccess;
    for (const FastKernel& k : kFastKernels) {
      cudaError_t ek = cudaFuncSetAttribute(k.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
      if (ek != cudaSuccess) e4 = ek;
    }
    cudaError_t e7 = cudaFuncSetAttribute(dp::conv_tc_kernel<dp::MODE_D, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaError_t e5 = cudaFuncSetAttribute(dp::dense_layer_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e5 == cudaSuccess) e5 = cudaFuncSetAttribute(dp::dense_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    cudaError_t e6 = cudaFuncSetAttribute(dp::dense_layer_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e6 == cudaSuccess) e6 = cudaFuncSetAttribute(dp::conv_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e6 == cudaSuccess) e6 = cudaFuncSetAttribute(dp::conv_halo_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e6 == cudaSuccess) e6 = cudaFuncSetAttribute(dp::conv_1x1_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess || e5 != cudaSuccess || e6 != cudaSuccess || e7 != cudaSuccess) {
      return fail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
  }
  env = getenv("DP_B_PAIR");
  if (env) m->b_pair = atoi(env);
  env = getenv("DP_DENSE_BLOCK");
  if (env) m->dense_block = atoi(env);
  env = getenv("DP_NAIVE_CONV");
  if (env && atoi(env)) m->naive_conv = 1;
  env = getenv("DP_DESC_BASE_MODE");
  if (env) m->desc_base_mode = atoi(env);
  return 0;
}

extern "C" {

int dp_abi_version(void) { return 1; }

const char* dp_last_error(void) { return g_err.c_str(); }

uint64_t dp_kernel_launch_count(void) { return g_launches.load(); }

void dp_d4_src(int code, int i, int j, int P, int* a, int* b) { dp::d4_src(code, i, j, P, *a, *b); }

int dp_model_create(const void* blob, size_t nbytes, int device, int max_batch, dp_model** out) {
  if (!blob || !out) return fail("null argument");
  if (nbytes < sizeof(BlobHeader)) return fail("container too small");
  const uint8_t* bytes = static_cast<const uint8_t*>(blob);
  BlobHeader h;
  memcpy(&h, bytes, sizeof h);
  if (memcmp(h.magic, "DPB1", 4) || h.version != 1) return fail("not a DPB1 v1 container");
  if (h.total_bytes != nbytes) return fail("container size mismatch: header says %llu, got %zu",
                                           (unsigned long long)h.total_bytes, nbytes);
  const size_t tab = sizeof h + (size_t)h.n_bufs * sizeof(BlobBuf) + (size_t)h.n_ops * sizeof(BlobOp);
  if (tab > h.data_off || h.data_off > nbytes) return fail("container tables out of range");
  if (max_batch < 1) return fail("max_batch must be >= 1");
  int ndev = 0;
  CU_OK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail("CUDA device %d not available (%d devices)", device, ndev);
  CU_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail("this library is built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);

  dp_model* m = new dp_model;
  m->device = device;
  m->max_batch = max_batch;
  m->patch = (int)h.patch;
  if (h.precision > 2) { delete m; return fail("container asks for unknown precision %u", h.precision); }
  m->precision = (int)h.precision;
  m->esize = m->precision ? 4 : 2;
  m->num_sms = prop.multiProcessorCount;
  m->bufs.resize(h.n_bufs);
  m->ops.resize(h.n_ops);
  memcpy(m->bufs.data(), bytes + sizeof h, h.n_bufs * sizeof(BlobBuf));
  memcpy(m->ops.data(), bytes + sizeof h + h.n_bufs * sizeof(BlobBuf), h.n_ops * sizeof(BlobOp));
  for (const BlobOp& op : m->ops) {
    if (op.in_buf >= (int)h.n_bufs || op.out_buf >= (int)h.n_bufs) {
      delete m;
      return fail("op references a buffer out of range");
    }
  }
  m->data_bytes = nbytes - h.data_off;
  auto cleanup = [&]() { dp_model_destroy(m); };
  cudaError_t e = cudaMalloc(&m->data_dev, m->data_bytes ? m->data_bytes : 256);
  if (e != cudaSuccess) { cleanup(); return fail("cudaMalloc weights: %s", cudaGetErrorString(e)); }
  m->data_owner = std::shared_ptr<void>(m->data_dev, [](void* p) { cudaFree(p); });
  m->device_bytes += m->data_bytes;
  e = cudaMemcpy(m->data_dev, bytes + h.data_off, m->data_bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { cleanup(); return fail("cudaMemcpy weights: %s", cudaGetErrorString(e)); }
  if (m->precision == 2 && m->data_bytes >= 4) {   // TF32 hi / lo copies of the weights for the pre-split operand path
    e = cudaMalloc(&m->data_hi_dev, m->data_bytes);
    if (e == cudaSuccess) m->data_hi_owner = std::shared_ptr<void>(m->data_hi_dev, [](void* q) { cudaFree(q); });
    if (e == cudaSuccess) e = cudaMalloc(&m->data_lo_dev, m->data_bytes);
    if (e == cudaSuccess) m->data_lo_owner = std::shared_ptr<void>(m->data_lo_dev, [](void* q) { cudaFree(q); });
    if (e != cudaSuccess) { cleanup(); return fail("cudaMalloc TF32 weight copies: %s", cudaGetErrorString(e)); }
    dp::tf32_presplit_kernel<<<1024, 256>>>(reinterpret_cast<const float*>(m->data_dev), m->data_hi_dev, m->data_lo_dev,
                                            m->data_bytes / 4);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cleanup(); return fail("TF32 weight split: %s", cudaGetErrorString(e)); }
    m->device_bytes += 2 * m->data_bytes;
  }
  if (alloc_lane_state(m)) { cleanup(); return 1; }
  *out = m;
  return 0;
}

int dp_model_clone(const dp_model* src, dp_model** out) {
  if (check_model(src) || !out) return fail("null argument");
  CU_OK(cudaSetDevice(src->device));
  dp_model* m = new dp_model;
  m->device = src->device; m->max_batch = src->max_batch; m->patch = src->patch; m->num_sms = src->num_sms;
  m->precision = src->precision; m->esize = src->esize;
  m->bufs = src->bufs; m->ops = src->ops;
  m->data_dev = src->data_dev; m->data_owner = src->data_owner; m->data_bytes = src->data_bytes;
  m->data_hi_dev = src->data_hi_dev; m->data_hi_owner = src->data_hi_owner;
  m->data_lo_dev = src->data_lo_dev; m->data_lo_owner = src->data_lo_owner;
  if (alloc_lane_state(m)) { dp_model_destroy(m); return 1; }
  m->use_graph = src->use_graph; m->split = src->split; m->use_pdl = src->use_pdl; m->epi_direct = src->epi_direct;
  m->use_overlap = src->use_overlap; m->b_resident = src->b_resident; m->b_pair = src->b_pair;
  m->dense_block = src->dense_block; m->naive_conv = src->naive_conv; m->desc_base_mode = src->desc_base_mode;
  m->halo_pad8 = src->halo_pad8;
  *out = m;
  return 0;
}

int dp_model_destroy(dp_model* m) {
  if (!m) return 0;
  cudaSetDevice(m->device);
  cudaDeviceSynchronize();
  for (cudaEvent_t e : m->ev) cudaEventDestroy(e);
  for (auto& kv : m->plans) {
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
  }
  for (cudaStream_t bs : m->branch_streams) cudaStreamDestroy(bs);
  for (cudaEvent_t be : m->branch_events) cudaEventDestroy(be);
  if (m->cap_stream) cudaStreamDestroy(m->cap_stream);
  if (m->fork_event) cudaEventDestroy(m->fork_event);
  if (m->pass_dev) cudaFree(m->pass_dev);
  if (m->trace_dev) cudaFree(m->trace_dev);
  if (m->gt_dev) cudaFree(m->gt_dev);
  if (m->gt_all_dev) cudaFree(m->gt_all_dev);
  for (__half* p : m->buf_dev)
    if (p) cudaFree(p);
  if (m->scratch_head) cudaFree(m->scratch_head);
  m->data_owner.reset();   // frees the data section with its last user
  m->data_hi_owner.reset();
  m->data_lo_owner.reset();
  delete m;
  return 0;
}

int dp_model_info(const dp_model* m, int* patch, int* max_batch, uint64_t* device_bytes) {
  if (check_model(m)) return 1;
  if (patch) *patch = m->patch;
  if (max_batch) *max_batch = m->max_batch;
  if (device_bytes) *device_bytes = m->device_bytes;
  return 0;
}

int dp_model_precision(const dp_model* m) { return m ? m->precision : -1; }

int dp_model_set_option(dp_model* m, const char* key, int value) {
  if (check_model(m) || !key) return fail("null argument");
  if (!strcmp(key, "naive_conv")) m->naive_conv = value;
  else if (!strcmp(key, "desc_base_mode")) m->desc_base_mode = value;
  else if (!strcmp(key, "profile")) m->profile = value;
  else if (!strcmp(key, "use_graph")) m->use_graph = value;
  else if (!strcmp(key, "stamp")) {
    if (!m->gt_dev) CU_OK(cudaMalloc(&m->gt_dev, 2 * m->ops.size() * sizeof(unsigned long long)));
    CU_OK(cudaMemset(m->gt_dev, 0, 2 * m->ops.size() * sizeof(unsigned long long)));
    m->stamp = value;
  }
  else if (!strcmp(key, "stamp_ctas")) {
    const size_t nb = m->ops.size() * 1024 * sizeof(unsigned long long);
    if (!m->gt_all_dev) CU_OK(cudaMalloc(&m->gt_all_dev, nb));
    CU_OK(cudaMemset(m->gt_all_dev, 0, nb));
    m->stamp_ctas = value;
  }
  else if (!strcmp(key, "b_resident") || !strcmp(key, "b_pair")) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (!strcmp(key, "b_pair")) m->b_pair = value; else m->b_resident = value;
    for (auto& kv : m->plans) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    m->plans.clear();
  }
  else if (!strcmp(key, "use_overlap") || !strcmp(key, "dense_block")) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (!strcmp(key, "dense_block")) m->dense_block = value; else m->use_overlap = value;
    for (auto& kv : m->plans) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    m->plans.clear();
  }
  else if (!strcmp(key, "epi_direct")) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->epi_direct = value;
    for (auto& kv : m->plans) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    m->plans.clear();   // the store path is part of the plan (shared-memory layout)
  }
  else if (!strcmp(key, "use_pdl")) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->use_pdl = value;
    for (auto& kv : m->plans) {   // captured graphs bake the launch attribute
      if (kv.second.exec) { cudaGraphExecDestroy(kv.second.exec); kv.second.exec = nullptr; }
      if (kv.second.graph) { cudaGraphDestroy(kv.second.graph); kv.second.graph = nullptr; }
    }
  }
  else if (!strcmp(key, "trace_op") || !strcmp(key, "trace_block")) {
    if (!m->trace_dev) CU_OK(cudaMalloc(&m->trace_dev, 10016 * sizeof(unsigned long long)));
    CU_OK(cudaMemset(m->trace_dev, 0, 10016 * sizeof(unsigned long long)));
    if (!strcmp(key, "trace_op")) m->trace_op = value; else m->trace_block = value;
  }
  else if (!strcmp(key, "split")) m->split = value < 1 ? 1 : value;
  else if (!strcmp(key, "halo_pad8")) {
    std::lock_guard<std::mutex> lk(m->mu);
    m->halo_pad8 = value;
    for (auto& kv : m->plans) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
    }
    m->plans.clear();  // tile shapes depend on it
  }
  else return fail("unknown option '%s'", key);
  return 0;
}

int dp_model_program_size(const dp_model* m, int* n_ops, int* n_bufs) {
  if (check_model(m)) return 1;
  if (n_ops) *n_ops = (int)m->ops.size();
  if (n_bufs) *n_bufs = (int)m->bufs.size();
  return 0;
}

int dp_model_buffer_shape(const dp_model* m, int buf, int* h, int* w, int* c) {
  if (check_model(m)) return 1;
  if (buf < 0 || buf >= (int)m->bufs.size()) return fail("buffer %d out of range", buf);
  if (h) *h = m->bufs[buf].H;
  if (w) *w = m->bufs[buf].W;
  if (c) *c = m->bufs[buf].C;
  return 0;
}

static int upload_pass(dp_model* m, const dp::PassDesc& d, cudaStream_t st) {
  // pageable-host source: the runtime stages the 48 bytes before returning, so `d` may live on the stack
  CU_OK(cudaMemcpyAsync(m->pass_dev, &d, sizeof d, cudaMemcpyHostToDevice, st));
  return 0;
}

// Direct (un-captured) execution of ops [op_begin, op_end) on the whole batch: profiling, naive and debug paths.
static int run_range(dp_model* m, int B, int op_begin, int op_end, const dp::PassDesc& d, void* stream) {
  if (check_model(m)) return 1;
  if (B < 1 || B > m->max_batch) return fail("n_tiles %d outside [1, %d]", B, m->max_batch);
  if (op_begin < 0 || op_end > (int)m->ops.size() || op_begin > op_end) return fail("op range out of bounds");
  CU_OK(cudaSetDevice(m->device));
  Plan* plan = nullptr;
  if (get_plan(m, B, 1, &plan)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool whole = op_begin == 0 && op_end == (int)m->ops.size();   // persistent block kernels need the whole block
  if (upload_pass(m, d, st)) return 1;
  if (m->profile && m->ev.empty()) {
    m->ev.resize(2 * m->ops.size());
    for (auto& e : m->ev) CU_OK(cudaEventCreate(&e));
  }
  for (int i = op_begin; i < op_end; ++i) {
    const BlobOp& op = m->ops[i];
    if ((op.type == OP_STEM_IM2COL || op.type == OP_STEM_S2D) && (!d.slide || !d.coords)) return fail("op %d: stem needs a slide and coordinates", i);
    if (((op.type == OP_CONV && op.head) || op.type == OP_HEAD_RESIZE) && !d.probs_out)
      return fail("op %d: head needs a probability output buffer", i);
    if (m->profile) CU_OK(cudaEventRecord(m->ev[2 * i], st));
    if (run_op(m, &plan->subs[0], i, st, whole)) {
      g_err = "op " + std::to_string(i) + ": " + g_err;
      return 1;
    }
    if (m->profile) CU_OK(cudaEventRecord(m->ev[2 * i + 1], st));
  }
  if (m->profile) { m->ev_begin = op_begin; m->ev_end = op_end; }
  return 0;
}

// Whole forward as one CUDA-graph launch: `split` parallel branches, one per sub-batch.
static int run_graph(dp_model* m, int B, const dp::PassDesc& d, void* stream) {
  CU_OK(cudaSetDevice(m->device));
  Plan* plan = nullptr;
  if (get_plan(m, B, m->split, &plan)) return 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!plan->exec) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (!plan->exec) {
      const int ns = (int)plan->subs.size();
      if (!m->cap_stream) {
        CU_OK(cudaStreamCreateWithFlags(&m->cap_stream, cudaStreamNonBlocking));
        CU_OK(cudaEventCreateWithFlags(&m->fork_event, cudaEventDisableTiming));
      }
      while ((int)m->branch_streams.size() < ns) {
        cudaStream_t bs; cudaEvent_t be;
        CU_OK(cudaStreamCreateWithFlags(&bs, cudaStreamNonBlocking));
        CU_OK(cudaEventCreateWithFlags(&be, cudaEventDisableTiming));
        m->branch_streams.push_back(bs);
        m->branch_events.push_back(be);
      }
      CU_OK(cudaStreamBeginCapture(m->cap_stream, cudaStreamCaptureModeThreadLocal));
      g_captured = 0;
      int rc = 0;
      cudaError_t ce = cudaEventRecord(m->fork_event, m->cap_stream);
      for (int s = 0; s < ns && !rc && ce == cudaSuccess; ++s) {
        cudaStream_t bs = (s == 0) ? m->cap_stream : m->branch_streams[s];
        if (s) ce = cudaStreamWaitEvent(bs, m->fork_event, 0);
        for (int i = 0; i < (int)m->ops.size() && !rc && ce == cudaSuccess; ++i) rc = run_op(m, &plan->subs[s], i, bs, true);
        if (s && !rc && ce == cudaSuccess) {
          ce = cudaEventRecord(m->branch_events[s], bs);
          if (ce == cudaSuccess) ce = cudaStreamWaitEvent(m->cap_stream, m->branch_events[s], 0);
        }
      }
      cudaGraph_t g = nullptr;
      cudaError_t ee = cudaStreamEndCapture(m->cap_stream, &g);
      if (rc) return 1;
      if (ce != cudaSuccess) return fail("graph capture failed: %s", cudaGetErrorString(ce));
      if (ee != cudaSuccess) return fail("cudaStreamEndCapture failed: %s", cudaGetErrorString(ee));
      cudaGraphExec_t ex = nullptr;
      CU_OK(cudaGraphInstantiate(&ex, g, 0));
      plan->graph = g;
      plan->n_launches = g_captured;
      plan->exec = ex;
    }
  }
  if (upload_pass(m, d, st)) return 1;
  CU_OK(cudaGraphLaunch(plan->exec, st));
  g_launches.fetch_add(plan->n_launches, std::memory_order_relaxed);
  return 0;
}

int dp_forward_tiles(dp_model* m, const uint8_t* slide, int64_t slide_w, int64_t slide_h, const int32_t* coords,
                     int n_tiles, int tta_in, int tta_out, float* probs_out, void* stream) {
  if (check_model(m)) return 1;
  if (!slide || !coords || !probs_out) return fail("null argument");
  if (slide_w < 1 || slide_h < 1) return fail("bad slide extent");
  if ((tta_in | tta_out) & ~7) return fail("D4 codes must be in [0, 8)");
  if (n_tiles < 1 || n_tiles > m->max_batch) return fail("n_tiles %d outside [1, %d]", n_tiles, m->max_batch);
  dp::PassDesc d;
  if (slide_w < m->patch || slide_h < m->patch)
    return fail("raster %lld x %lld is smaller than the %d-pixel patch", (long long)slide_w, (long long)slide_h, m->patch);
  d.slide = slide; d.slide_h = slide_h; d.slide_w = slide_w; d.coords = coords; d.probs_out = probs_out;
  d.tta_in = tta_in; d.tta_out = tta_out;
  if (m->use_graph && !m->profile && !m->naive_conv) return run_graph(m, n_tiles, d, stream);
  return run_range(m, n_tiles, 0, (int)m->ops.size(), d, stream);
}

int dp_debug_run_ops(dp_model* m, int n_tiles, int op_begin, int op_end, int tta_out, float* probs_out,
                     void* stream) {
  dp::PassDesc d;
  memset(&d, 0, sizeof d);
  d.tta_out = tta_out; d.probs_out = probs_out;
  return run_range(m, n_tiles, op_begin, op_end, d, stream);
}

int dp_debug_read_buffer(dp_model* m, int buf, int n_tiles, void* host, size_t nbytes) {
  if (check_model(m) || !host) return fail("null argument");
  if (buf < 0 || buf >= (int)m->bufs.size()) return fail("buffer %d out of range", buf);
  const BlobBuf& b = m->bufs[buf];
  const size_t need = (size_t)n_tiles * b.H * b.W * b.C * m->esize;
  if (n_tiles > m->max_batch || nbytes != need) return fail("size mismatch: need %zu bytes, got %zu", need, nbytes);
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(host, m->buf_dev[buf], need, cudaMemcpyDeviceToHost));
  return 0;
}

int dp_debug_write_buffer(dp_model* m, int buf, int n_tiles, const void* host, size_t nbytes) {
  if (check_model(m) || !host) return fail("null argument");
  if (buf < 0 || buf >= (int)m->bufs.size()) return fail("buffer %d out of range", buf);
  const BlobBuf& b = m->bufs[buf];
  const size_t need = (size_t)n_tiles * b.H * b.W * b.C * m->esize;
  if (n_tiles > m->max_batch || nbytes != need) return fail("size mismatch: need %zu bytes, got %zu", need, nbytes);
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(m->buf_dev[buf], host, need, cudaMemcpyHostToDevice));
  return 0;
}

int dp_model_op_times(dp_model* m, float* ms, int n) {
  if (check_model(m) || !ms) return fail("null argument");
  if (n != (int)m->ops.size()) return fail("expected room for %zu ops", m->ops.size());
  if (m->ev.empty()) return fail("no profiled run yet (set option 'profile' and run a forward)");
  CU_OK(cudaSetDevice(m->device));
  for (int i = 0; i < n; ++i) ms[i] = 0.f;
  for (int i = m->ev_begin; i < m->ev_end; ++i) {
    CU_OK(cudaEventSynchronize(m->ev[2 * i + 1]));
    CU_OK(cudaEventElapsedTime(&ms[i], m->ev[2 * i], m->ev[2 * i + 1]));
  }
  return 0;
}

int dp_debug_read_stamps(dp_model* m, unsigned long long* out, int n) {
  if (check_model(m) || !out) return fail("null argument");
  if (!m->gt_dev) return fail("no stamps (set option 'stamp')");
  if (n != 2 * (int)m->ops.size()) return fail("expected room for %zu values", 2 * m->ops.size());
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(out, m->gt_dev, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int dp_debug_read_cta_stamps(dp_model* m, int op, unsigned long long* out, int n) {
  if (check_model(m) || !out) return fail("null argument");
  if (!m->gt_all_dev) return fail("no per-CTA stamps (set option 'stamp_ctas')");
  if (op < 0 || op >= (int)m->ops.size() || n != 1024) return fail("expected op in range and room for 1024 values");
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(out, m->gt_all_dev + (size_t)op * 1024, 1024 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int dp_debug_read_trace(dp_model* m, unsigned long long* out, int n) {
  if (check_model(m) || !out) return fail("null argument");
  if (!m->trace_dev) return fail("no trace buffer (set option 'trace_op')");
  if (n > 10016) n = 10016;
  CU_OK(cudaSetDevice(m->device));
  CU_OK(cudaDeviceSynchronize());
  CU_OK(cudaMemcpy(out, m->trace_dev, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return 0;
}

int dp_model_op_info(const dp_model* m, int op, int* type, int* kind, int* cin, int* cout, int* h, int* w,
                     uint64_t* macs_per_tile) {
  if (check_model(m)) return 1;
  if (op < 0 || op >= (int)m->ops.size()) return fail("op %d out of range", op);
  const BlobOp& o = m->ops[op];
  if (type) *type = o.type;
  if (kind) *kind = o.kind;
  if (cin) *cin = o.cin;
  if (cout) *cout = o.cout;
  if (h) *h = m->bufs[o.in_buf].H;
  if (w) *w = m->bufs[o.in_buf].W;
  if (macs_per_tile) {
    *macs_per_tile = 0;
    if (o.type == OP_CONV || o.type == OP_DENSE_LAYER) {
      Plan* plan = nullptr;
      if (get_plan(const_cast<dp_model*>(m), 1, 1, &plan)) return 1;
      *macs_per_tile = plan->subs[0].launches[op].macs;
    }
  }
  return 0;
}

int dp_model_executed_macs(const dp_model* m, int n_tiles, uint64_t* macs) {
  if (check_model(m) || !macs) return fail("null argument");
  Plan* plan = nullptr;
  if (get_plan(const_cast<dp_model*>(m), n_tiles, 1, &plan)) return 1;
  uint64_t t = 0;
  for (const Launch& L : plan->subs[0].launches) t += L.macs;
  *macs = t;
  return 0;
}

int dp_stitch(const float* probs, int n_pass, int n_tiles, int patch, const int32_t* coords, float* mean,
              float* var, uint8_t* count, int64_t plane_w, int64_t plane_h, int64_t x_lo, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!probs || !coords || !mean || !var || !count) return fail("null argument");
  if (n_pass < 1 || n_tiles < 1 || patch < 1 || plane_w < patch || plane_h < patch) return fail("bad stitch geometry");
  if (n_tiles > 4096) return fail("at most 4096 tiles per stitch call");
  dim3 grid((patch * patch + 255) / 256, n_tiles);
  dp::stitch_kernel<<<grid, 256, 2 * n_tiles * sizeof(int), st>>>(
      probs, n_pass, n_tiles, patch, coords, mean, var, count, plane_w, plane_h, (int)x_lo);
  LAUNCH_OK();
  return 0;
}

int dp_finalize(float* mean, float* var, uint8_t* count, int64_t n, float threshold, uint8_t* label, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!mean || !var || !count) return fail("null argument");
  if (n < 1) return fail("empty plane");
  if ((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(var) | reinterpret_cast<uintptr_t>(count) |
       reinterpret_cast<uintptr_t>(label)) & 15)
    return fail("plane pointers must be 16-byte aligned");
  dp::finalize_kernel<<<grid_for((n + 15) / 16, 256), 256, 0, st>>>(
      mean, var, count, n, threshold, label);
  LAUNCH_OK();
  return 0;
}

int dp_pyramid_down2(const float* in, int64_t w, int64_t h, float* out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!in || !out) return fail("null argument");
  if (w < 2 || h < 2) return fail("plane too small");
  dp::pyramid_down2_kernel<<<grid_for((w / 2) * (h / 2), 256), 256, 0, st>>>(in, w, h,
                                                                                                          out);
  LAUNCH_OK();
  return 0;
}

int dp_tissue_hist(const uint8_t* rgb, int64_t n_pix, uint32_t* hist, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!rgb || !hist) return fail("null argument");
  if (n_pix < 1) return fail("empty image");
  CU_OK(cudaMemsetAsync(hist, 0, dp::kTissueHistBins * sizeof(uint32_t), st));
  dp::tissue_hist_kernel<<<grid_for(n_pix, 256), 256, 0, st>>>(rgb, n_pix, hist);
  LAUNCH_OK();
  return 0;
}

int dp_tissue_mask(const uint8_t* rgb, int64_t n_pix, int thr_r, int thr_g, int thr_b, int rgb_min,
                   const uint8_t* sat_lut, uint8_t* mask, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!rgb || !sat_lut || !mask) return fail("null argument");
  if (n_pix < 1) return fail("empty image");
  dp::tissue_mask_kernel<<<grid_for(n_pix, 256), 256, 0, st>>>(rgb, n_pix, thr_r, thr_g, thr_b, rgb_min, sat_lut, mask);
  LAUNCH_OK();
  return 0;
}

int dp_morph_rect(const uint8_t* in, uint8_t* out, uint8_t* tmp, int n0, int n1, int k, int dilate, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!in || !out || !tmp) return fail("null argument");
  if (n0 < 1 || n1 < 1 || k < 1) return fail("bad morphology geometry");
  if (tmp == in || tmp == out) return fail("tmp must not alias in / out");
  const long long total = (long long)n0 * n1;
  dp::morph_line_kernel<<<grid_for(total, 256), 256, 0, st>>>(in, tmp, n0, n1, k, 1, dilate ? 1 : 0);
  LAUNCH_OK();
  dp::morph_line_kernel<<<grid_for(total, 256), 256, 0, st>>>(tmp, out, n0, n1, k, 0, dilate ? 1 : 0);
  LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ lattice CRF
namespace {
struct LatGeom {
  int D, log2cap, cap, m_max;
  size_t bytes;   // per tile
};
LatGeom lat_geom(int D, int N) {
  LatGeom g;
  g.D = D;
  g.m_max = N * (D + 1);
  g.log2cap = 1;
  while ((1 << g.log2cap) < 2 * g.m_max) ++g.log2cap;
  g.cap = 1 << g.log2cap;
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  g.bytes = al(8ull * g.cap) + al(4ull * g.cap) + al(8ull * g.m_max) + 2 * al(4ull * g.m_max) +
            al(8ull * (D + 1) * g.m_max) + al(16ull * g.m_max) + 2 * al(8ull * g.m_max) + 256;
  return g;
}
size_t lat_pixel_bytes(int N) { return (size_t)N * (8 + 4 + 4 + 4 + 8 + 8 + 8) + 8 * 256; }
// tiles per pass through the lattice kernels: every launch then covers chunk x 65 536 pixels / lattice points (the
// kernels are small and latency bound: ~200 launches per chunk), ~75 MB of workspace per tile of the chunk
static int crf_chunk() {
  static const int v = [] { const char* e = getenv("DP_CRF_CHUNK"); const int c = e ? atoi(e) : 32; return c < 1 ? 1 : (c > 256 ? 256 : c); }();
  return v;
}
#define kCrfChunk crf_chunk()
}  // namespace

size_t dp_crf_lattice_workspace_bytes(int n_tiles, int h, int w) {
  if (n_tiles < 1 || h < 1 || w < 1) return 0;
  const int N = h * w, chunk = n_tiles < kCrfChunk ? n_tiles : kCrfChunk;
  return (size_t)chunk * (lat_geom(2, N).bytes + lat_geom(5, N).bytes + lat_pixel_bytes(N)) +
         2 * (size_t)chunk * sizeof(dp::LatticeTile) + 1024;
}

int dp_crf_tiles_lattice(const uint8_t* rgb, const float* p1, int n_tiles, int h, int w, int n_iter, float sdims_gauss,
                         float compat_gauss, float sdims_bilateral, float schan_bilateral, float compat_bilateral,
                         void* workspace, size_t workspace_bytes, uint8_t* labels, float* q1_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!rgb || !p1 || !workspace || (!labels && !q1_out)) return fail("null argument");
  if (n_tiles < 1 || h < 1 || w < 1 || n_iter < 0) return fail("bad CRF geometry");
  if (!(sdims_gauss > 0) || !(sdims_bilateral > 0) || !(schan_bilateral > 0)) return fail("CRF kernel widths must be positive");
  if (workspace_bytes < dp_crf_lattice_workspace_bytes(n_tiles, h, w))
    return fail("CRF workspace too small: need %zu bytes", dp_crf_lattice_workspace_bytes(n_tiles, h, w));
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail("CRF workspace must be 256-byte aligned");
  const int N = h * w;
  // packed keys hold 12 bits per coordinate: |feature| * scale stays far below 2048 for tile-sized inputs
  if ((float)(h > w ? h : w) / sdims_gauss > 300.f || 255.f / schan_bilateral > 100.f)
    return fail("CRF kernel widths too small for the lattice key range");
  const bool use_bil = compat_bilateral != 0.f;
  const LatGeom g2 = lat_geom(2, N), g5 = lat_geom(5, N);
  const int chunk = n_tiles < kCrfChunk ? n_tiles : kCrfChunk;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  auto al = [](size_t b) { return (b + 255) / 256 * 256; };
  // carve: tile structs | per-chunk pixel arrays | lattices
  dp::LatticeTile* t2_dev = reinterpret_cast<dp::LatticeTile*>(base);
  dp::LatticeTile* t5_dev = t2_dev + chunk;
  size_t off = al(2 * (size_t)chunk * sizeof(dp::LatticeTile));
  auto take = [&](size_t bytes) { uint8_t* q = base + off; off += al(bytes); return q; };
  float* u = reinterpret_cast<float*>(take((size_t)chunk * N * 8));
  float* q1 = reinterpret_cast<float*>(take((size_t)chunk * N * 4));
  float* norm_g = reinterpret_cast<float*>(take((size_t)chunk * N * 4));
  float* norm_b = reinterpret_cast<float*>(take((size_t)chunk * N * 4));
  float* fin = reinterpret_cast<float*>(take((size_t)chunk * N * 8));
  float* fg = reinterpret_cast<float*>(take((size_t)chunk * N * 8));
  float* fb = reinterpret_cast<float*>(take((size_t)chunk * N * 8));
  const size_t lat_off = off;
  std::vector<dp::LatticeTile> t2(chunk), t5(chunk);
  auto carve_lat = [&](const LatGeom& g, dp::LatticeTile& L) {
    L.keys = reinterpret_cast<unsigned long long*>(take(8ull * g.cap));
    L.ids = reinterpret_cast<int*>(take(4ull * g.cap));
    L.ckeys = reinterpret_cast<unsigned long long*>(take(8ull * g.m_max));
    L.offset = reinterpret_cast<int*>(take(4ull * g.m_max));
    L.bary = reinterpret_cast<float*>(take(4ull * g.m_max));
    L.nb = reinterpret_cast<int*>(take(8ull * (g.D + 1) * g.m_max));
    L.fix = reinterpret_cast<long long*>(take(16ull * g.m_max));
    L.val_a = reinterpret_cast<float*>(take(8ull * g.m_max));
    L.val_b = reinterpret_cast<float*>(take(8ull * g.m_max));
    L.count = reinterpret_cast<int*>(take(256));
  };
  for (int t = 0; t < chunk; ++t) { carve_lat(g2, t2[t]); carve_lat(g5, t5[t]); }
  const size_t lat_bytes = off - lat_off;
  if (off > workspace_bytes) return fail("internal: CRF workspace carve exceeds the buffer");
  CU_OK(cudaMemcpyAsync(t2_dev, t2.data(), chunk * sizeof(dp::LatticeTile), cudaMemcpyHostToDevice, st));
  CU_OK(cudaMemcpyAsync(t5_dev, t5.data(), chunk * sizeof(dp::LatticeTile), cudaMemcpyHostToDevice, st));

  auto blocks = [](long long n) { return (unsigned)((n + 255) / 256); };
  for (int t0 = 0; t0 < n_tiles; t0 += chunk) {
    const int nt = (n_tiles - t0 < chunk) ? n_tiles - t0 : chunk;
    const long long total = (long long)nt * N;
    const int ew = grid_for(total, 256);
    const uint8_t* rgb_c = rgb + (size_t)t0 * N * 3;
    // keys = empty (all ones), everything else of the lattices (counts, accumulators) = 0
    CU_OK(cudaMemsetAsync(base + lat_off, 0, lat_bytes, st));
    for (int t = 0; t < nt; ++t) {
      CU_OK(cudaMemsetAsync(t2[t].keys, 0xFF, 8ull * g2.cap, st));
      if (use_bil) CU_OK(cudaMemsetAsync(t5[t].keys, 0xFF, 8ull * g5.cap, st));
    }
    auto build = [&](auto Dtag, const LatGeom& g, dp::LatticeTile* tiles, float inv_sd, float inv_sc) {
      constexpr int D = decltype(Dtag)::value;
      dp::lat_build_kernel<D><<<dim3(blocks(N), nt), 256, 0, st>>>(rgb_c, h, w, inv_sd, inv_sc, tiles, g.log2cap);
      dp::lat_compact_kernel<<<dim3(blocks(g.cap), nt), 256, 0, st>>>(tiles, g.cap);
      dp::lat_remap_kernel<<<dim3(blocks(g.m_max), nt), 256, 0, st>>>(tiles, g.m_max);
      dp::lat_neighbors_kernel<D><<<dim3(blocks(g.m_max), nt), 256, 0, st>>>(tiles, g.log2cap, g.m_max);
      g_launches.fetch_add(4, std::memory_order_relaxed);
    };
    auto filter = [&](auto Dtag, const LatGeom& g, dp::LatticeTile* tiles, const float* in, float* out) {
      constexpr int D = decltype(Dtag)::value;
      dp::lat_splat_kernel<D><<<dim3(blocks(g.m_max), nt), 256, 0, st>>>(tiles, in, N);
      dp::lat_fix2f_kernel<<<dim3(blocks(g.m_max), nt), 256, 0, st>>>(tiles);
      for (int j = 0; j <= D; ++j) dp::lat_blur_kernel<<<dim3(blocks(g.m_max), nt), 256, 0, st>>>(tiles, j, g.m_max);
      dp::lat_slice_kernel<D><<<dim3(blocks(N), nt), 256, 0, st>>>(tiles, out, N);
      g_launches.fetch_add(D + 4, std::memory_order_relaxed);
    };
    using D2 = std::integral_constant<int, 2>;
    using D5 = std::integral_constant<int, 5>;
    build(D2{}, g2, t2_dev, 1.f / sdims_gauss, 0.f);
    if (use_bil) build(D5{}, g5, t5_dev, 1.f / sdims_bilateral, 1.f / schan_bilateral);
    dp::mf_init_kernel<<<ew, 256, 0, st>>>(p1 + (size_t)t0 * N, total, u, q1, fin);
    filter(D2{}, g2, t2_dev, fin, fg);
    dp::mf_norm_kernel<<<ew, 256, 0, st>>>(fg, total, norm_g);
    if (use_bil) {
      filter(D5{}, g5, t5_dev, fin, fb);
      dp::mf_norm_kernel<<<ew, 256, 0, st>>>(fb, total, norm_b);
    }
    uint8_t* lab_c = labels ? labels + (size_t)t0 * N : nullptr;
    float* q_c = q1_out ? q1_out + (size_t)t0 * N : nullptr;
    if (n_iter == 0)
      dp::mf_update_kernel<<<ew, 256, 0, st>>>(u, nullptr, nullptr, 0.f, nullptr, nullptr, 0.f, total, q1, lab_c, q_c);
    for (int it = 0; it < n_iter; ++it) {
      dp::mf_scale_kernel<<<ew, 256, 0, st>>>(q1, norm_g, total, fin);
      filter(D2{}, g2, t2_dev, fin, fg);
      if (use_bil) {
        dp::mf_scale_kernel<<<ew, 256, 0, st>>>(q1, norm_b, total, fin);
        filter(D5{}, g5, t5_dev, fin, fb);
      }
      const bool last = it + 1 == n_iter;
      dp::mf_update_kernel<<<ew, 256, 0, st>>>(u, fg, norm_g, compat_gauss, use_bil ? fb : nullptr, norm_b, compat_bilateral,
                                               total, q1, last ? lab_c : nullptr, last ? q_c : nullptr);
    }
    LAUNCH_OK();
  }
  return 0;
}

size_t dp_crf_workspace_bytes(int n_tiles, int h, int w) {
  if (n_tiles < 1 || h < 1 || w < 1) return 0;
  return (size_t)n_tiles * dp::CRF_PLANES * (size_t)h * w * sizeof(float);
}

int dp_crf_tiles(const uint8_t* rgb, const float* p1, int n_tiles, int h, int w, int n_iter, float sdims_gauss,
                 float compat_gauss, float sdims_bilateral, float schan_bilateral, float compat_bilateral,
                 void* workspace, size_t workspace_bytes, uint8_t* labels, float* q1_out, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!rgb || !p1 || !workspace || (!labels && !q1_out)) return fail("null argument");
  if (n_tiles < 1 || h < 1 || w < 1 || n_iter < 0) return fail("bad CRF geometry");
  if (!(sdims_gauss > 0) || !(sdims_bilateral > 0) || !(schan_bilateral > 0)) return fail("CRF kernel widths must be positive");
  if (workspace_bytes < dp_crf_workspace_bytes(n_tiles, h, w))
    return fail("CRF workspace too small: need %zu bytes", dp_crf_workspace_bytes(n_tiles, h, w));
  float* ws = static_cast<float*>(workspace);
  const long long npix = (long long)h * w;
  const int ew = grid_for(npix * n_tiles, 256);
  const dim3 bgrid((unsigned)((npix + 255) / 256), (unsigned)n_tiles);
  auto filters = [&]() {   // AG -> G (separable spatial Gaussian), AB -> B (all-pairs bilateral)
    dp::crf_gauss_pass_kernel<<<ew, 256, 0, st>>>(n_tiles, h, w, 1, 1.f / sdims_gauss, dp::CRF_AG0, dp::CRF_T0, ws);
    dp::crf_gauss_pass_kernel<<<ew, 256, 0, st>>>(n_tiles, h, w, 0, 1.f / sdims_gauss, dp::CRF_T0, dp::CRF_G0, ws);
    dp::crf_bilateral_kernel<<<bgrid, 256, 0, st>>>(h, w, ws);
    g_launches.fetch_add(3, std::memory_order_relaxed);
    return cudaGetLastError();
  };
  dp::crf_init_kernel<<<ew, 256, 0, st>>>(rgb, p1, n_tiles, h, w, 1.f / sdims_bilateral, 1.f / schan_bilateral, ws);
  LAUNCH_OK();
  dp::crf_scale_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, ws, 1);
  LAUNCH_OK();
  CU_OK(filters());
  dp::crf_norm_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, ws);
  LAUNCH_OK();
  if (n_iter == 0) {   // MAP of the unary alone
    dp::crf_scale_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, ws, 0);
    LAUNCH_OK();
  }
  for (int it = 0; it < n_iter; ++it) {
    dp::crf_scale_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, ws, 0);
    LAUNCH_OK();
    CU_OK(filters());
    const bool last = it + 1 == n_iter;
    dp::crf_update_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, compat_gauss, compat_bilateral, ws, last ? labels : nullptr,
                                              last ? q1_out : nullptr);
    LAUNCH_OK();
  }
  if (n_iter == 0) {
    // Q is still softmax(-U): labels from it
    dp::crf_update_kernel<<<ew, 256, 0, st>>>(n_tiles, npix, 0.f, 0.f, ws, labels, q1_out);
    LAUNCH_OK();
  }
  return 0;
}

// ---------------------------------------------------------------- JPEG tile encoder of the pyramidal writer (jpeg_enc.cuh)
size_t dp_jpeg_encode_workspace_bytes(int n_tiles, int scratch_bytes_per_tile) {
  if (n_tiles < 1 || scratch_bytes_per_tile < 1024) return 0;
  return 4096 + (size_t)n_tiles * (size_t)((scratch_bytes_per_tile + 3) / 4 * 4);
}

int dp_jpeg_encode_gray_tiles(const uint8_t* plane, int64_t rows, int64_t cols, int tile0, int n_tiles, const void* tables_host,
                              size_t tables_bytes, void* workspace, size_t workspace_bytes, int scratch_bytes_per_tile,
                              uint8_t* out, int out_cap, int32_t* sizes, int32_t* flags, void* stream) {
  if (!plane || !tables_host || !workspace || !out || !sizes || !flags) return fail("null argument");
  if (rows < 1 || cols < 1 || rows > (1 << 30) || cols > (1 << 30)) return fail("bad plane extent");
  if (tables_bytes != sizeof(dp::JpegTables)) return fail("JPEG tables: expected %zu bytes, got %zu", sizeof(dp::JpegTables), tables_bytes);
  const int tiles_x = (int)((cols + dp::kJpTile - 1) / dp::kJpTile), tiles_y = (int)((rows + dp::kJpTile - 1) / dp::kJpTile);
  if (n_tiles < 1 || tile0 < 0 || (long long)tile0 + n_tiles > (long long)tiles_x * tiles_y) return fail("tile range outside the plane");
  if (out_cap < 1024) return fail("out_cap too small");
  if (workspace_bytes < dp_jpeg_encode_workspace_bytes(n_tiles, scratch_bytes_per_tile) || scratch_bytes_per_tile < 1024)
    return fail("JPEG workspace too small: need %zu bytes", dp_jpeg_encode_workspace_bytes(n_tiles, scratch_bytes_per_tile));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int smem = dp::kJpBlocks * 64 * 2 + dp::kJpTile * dp::kJpTile;
  // per call, not once per process: the attribute belongs to the current device
  CU_OK(cudaFuncSetAttribute(dp::jpeg_encode_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  const int scratch_words = (scratch_bytes_per_tile + 3) / 4;
  // the tables are 1.3 KB: a synchronous copy keeps the caller's host buffer free to go away after the call
  CU_OK(cudaStreamSynchronize(st));
  CU_OK(cudaMemcpy(ws, tables_host, sizeof(dp::JpegTables), cudaMemcpyHostToDevice));
  CU_OK(cudaMemsetAsync(ws + 4096, 0, (size_t)n_tiles * scratch_words * 4, st));
  dp::jpeg_encode_tiles_kernel<<<n_tiles, dp::kJpThreads, smem, st>>>(plane, (int)rows, (int)cols, tiles_x, tile0, n_tiles,
                                                                     reinterpret_cast<const dp::JpegTables*>(ws),
                                                                     reinterpret_cast<uint32_t*>(ws + 4096), scratch_words, out,
                                                                     out_cap, sizes, flags);
  LAUNCH_OK();
  return 0;
}

int dp_jpeg_compact(const uint8_t* in, int cap, const int32_t* sizes, const int64_t* offsets, uint8_t* out, int n_tiles,
                    void* stream) {
  if (!in || !sizes || !offsets || !out || n_tiles < 1) return fail("null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dp::jpeg_compact_kernel<<<n_tiles, 256, 0, st>>>(
      in, cap, sizes, reinterpret_cast<const long long*>(offsets), out, n_tiles);
  LAUNCH_OK();
  return 0;
}

}  // extern "C"
