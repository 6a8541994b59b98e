// DenseCRF mean-field inference on the permutohedral lattice -- what pydensecrf (Kraehenbuehl's densecrf) computes for
// post_process_crf (DigiPathAI/helpers/utils.py:568-603) and do_crf (utils.py:548-566): the Gaussian filters K Q are
// evaluated by splat / blur / slice on the lattice of Adams, Baek & Davis 2010 instead of exactly (crf.cuh).
// Oracle: oracle/lattice_ref.py (same arithmetic in numpy; parity against a pydensecrf binary is unpinned).
//
// One lattice per tile and kernel (d = 2: positions; d = 5: positions + colour), built once and used for the
// normalisation vector and every mean-field iteration.  All kernels are HBM / latency bound index work:
//   lat_build      one thread per pixel: elevate, nearest remainder-0 point, rank, barycentric weights, the d+1 simplex
//                  vertices inserted into a per-tile open-addressing hash table (64-bit packed keys, atomicCAS)
//   lat_compact    occupied slots -> dense lattice-point ids (the ids are arbitrary; no result depends on them)
//   lat_remap      per-pixel vertex slots -> ids
//   lat_neighbors  per lattice point and axis: the two blur neighbours (hash lookups)
//   lat_splat      val[vertex] += b * v, accumulated in 2^-30 fixed point (integer atomics: order independent, so the
//                  filter is bit-reproducible), lat_fix2f converts to float
//   lat_blur       d+1 passes  val' = val + (val[n1] + val[n2]) / 2
//   lat_slice      out = alpha * sum_r b_r val[vertex_r]
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dp {

constexpr unsigned long long kLatEmpty = ~0ull;
constexpr float kLatFix = 1073741824.f;   // 2^30

struct LatticeTile {       // device pointers of one (tile, kernel) lattice inside the caller's workspace
  unsigned long long* keys;   // [cap] packed keys, kLatEmpty = free
  int* ids;                   // [cap] slot -> lattice point id
  unsigned long long* ckeys;  // [m_max] id -> packed key
  int* offset;                // [N][d+1] vertex id of every pixel
  float* bary;                // [N][d+1]
  int* nb;                    // [d+1][m_max][2] blur neighbours (-1 = none)
  long long* fix;             // [m_max][2] splat accumulators
  float* val_a;               // [m_max][2]
  float* val_b;               // [m_max][2]
  int* count;                 // [1] number of lattice points
};

__device__ __forceinline__ unsigned lat_hash(unsigned long long key, int log2cap) {
  return static_cast<unsigned>((key * 0x9E3779B97F4A7C15ull) >> (64 - log2cap));
}

template <int D>
__device__ __forceinline__ unsigned long long lat_pack(const int (&k)[D]) {
  unsigned long long p = 0;
#pragma unroll
  for (int i = 0; i < D; ++i) p |= static_cast<unsigned long long>((k[i] + 2048) & 0xFFF) << (12 * i);
  return p;
}
template <int D>
__device__ __forceinline__ void lat_unpack(unsigned long long p, int (&k)[D]) {
#pragma unroll
  for (int i = 0; i < D; ++i) k[i] = static_cast<int>((p >> (12 * i)) & 0xFFF) - 2048;
}

__device__ __forceinline__ int lat_insert(unsigned long long* keys, int log2cap, unsigned long long key) {
  const unsigned mask = (1u << log2cap) - 1;
  unsigned s = lat_hash(key, log2cap);
  for (;;) {
    const unsigned long long prev = atomicCAS(&keys[s], kLatEmpty, key);
    if (prev == kLatEmpty || prev == key) return static_cast<int>(s);
    s = (s + 1) & mask;
  }
}
__device__ __forceinline__ int lat_find(const unsigned long long* keys, int log2cap, unsigned long long key) {
  const unsigned mask = (1u << log2cap) - 1;
  unsigned s = lat_hash(key, log2cap);
  for (;;) {
    const unsigned long long cur = keys[s];
    if (cur == key) return static_cast<int>(s);
    if (cur == kLatEmpty) return -1;
    s = (s + 1) & mask;
  }
}

// features: (y, x) / sdims [, (r, g, b) / schan]; rgb is uint8 [n_tiles][h][w][3]
template <int D>
__global__ void lat_build_kernel(const uint8_t* __restrict__ rgb, int h, int w, float inv_sdims, float inv_schan,
                                 const LatticeTile* __restrict__ tiles, int log2cap) {
  const LatticeTile L = tiles[blockIdx.y];
  const int N = h * w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  constexpr int D1 = D + 1;
  float f[D];
  f[0] = static_cast<float>(i / w) * inv_sdims;
  f[1] = static_cast<float>(i % w) * inv_sdims;
  if (D > 2) {
    const uint8_t* px = rgb + (static_cast<size_t>(blockIdx.y) * N + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (2 + c < D) f[2 + c] = static_cast<float>(px[c]) * inv_schan;
  }
  // elevate
  float E[D1];
  {
    const float inv_std = sqrtf(2.f / 3.f) * static_cast<float>(D1);
    float sm = 0.f;
#pragma unroll
    for (int j = D; j > 0; --j) {
      const float scale = inv_std / sqrtf(static_cast<float>((j + 1) * j));   // 1 / sqrt((j-1+2)(j-1+1))
      const float cf = f[j - 1] * scale;
      E[j] = sm - static_cast<float>(j) * cf;
      sm += cf;
    }
    E[0] = sm;
  }
  const float down = 1.f / static_cast<float>(D1);
  float rem0[D1];
  int rank[D1];
  int ssum = 0;
#pragma unroll
  for (int k = 0; k < D1; ++k) {
    const float rd = roundf(down * E[k]);          // half away from zero, like C's round()
    rem0[k] = rd * static_cast<float>(D1);
    ssum += static_cast<int>(rd);
    rank[k] = 0;
  }
#pragma unroll
  for (int a = 0; a < D; ++a)
#pragma unroll
    for (int b = a + 1; b < D1; ++b) {
      if (E[a] - rem0[a] < E[b] - rem0[b]) ++rank[a]; else ++rank[b];
    }
  if (ssum > 0) {
#pragma unroll
    for (int k = 0; k < D1; ++k) {
      if (rank[k] >= D1 - ssum) { rem0[k] -= static_cast<float>(D1); rank[k] += ssum - D1; }
      else rank[k] += ssum;
    }
  } else if (ssum < 0) {
#pragma unroll
    for (int k = 0; k < D1; ++k) {
      if (rank[k] < -ssum) { rem0[k] += static_cast<float>(D1); rank[k] += D1 + ssum; }
      else rank[k] += ssum;
    }
  }
  float bary[D1 + 1];
#pragma unroll
  for (int k = 0; k < D1 + 1; ++k) bary[k] = 0.f;
#pragma unroll
  for (int k = 0; k < D1; ++k) {
    const float v = (E[k] - rem0[k]) * down;
    // bary[D - rank] += v; bary[D + 1 - rank] -= v   (dynamic index resolved by a compare chain: D1 <= 6)
#pragma unroll
    for (int q = 0; q < D1 + 1; ++q) {
      if (q == D - rank[k]) bary[q] += v;
      if (q == D1 - rank[k]) bary[q] -= v;
    }
  }
  bary[0] += 1.f + bary[D1];
#pragma unroll
  for (int r = 0; r < D1; ++r) {
    int key[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const int canon = (rank[k] <= D - r) ? r : r - D1;    // canonical[r][rank]
      key[k] = static_cast<int>(rem0[k]) + canon;
    }
    L.offset[i * D1 + r] = lat_insert(L.keys, log2cap, lat_pack<D>(key));
    L.bary[i * D1 + r] = bary[r];
  }
}

__global__ void lat_compact_kernel(const LatticeTile* __restrict__ tiles, int cap) {
  const LatticeTile L = tiles[blockIdx.y];
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= cap) return;
  const unsigned long long k = L.keys[s];
  if (k == kLatEmpty) return;
  const int id = atomicAdd(L.count, 1);
  L.ids[s] = id;
  L.ckeys[id] = k;
}

__global__ void lat_remap_kernel(const LatticeTile* __restrict__ tiles, int n_entries) {
  const LatticeTile L = tiles[blockIdx.y];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_entries) L.offset[e] = L.ids[L.offset[e]];
}

template <int D>
__global__ void lat_neighbors_kernel(const LatticeTile* __restrict__ tiles, int log2cap, int m_max) {
  const LatticeTile L = tiles[blockIdx.y];
  const int M = *L.count;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= M) return;
  int key[D];
  lat_unpack<D>(L.ckeys[id], key);
#pragma unroll
  for (int j = 0; j <= D; ++j) {
    int n1[D], n2[D];
#pragma unroll
    for (int k = 0; k < D; ++k) { n1[k] = key[k] - 1; n2[k] = key[k] + 1; }
#pragma unroll
    for (int k = 0; k < D; ++k)
      if (k == j) { n1[k] = key[k] + D; n2[k] = key[k] - D; }
    const int s1 = lat_find(L.keys, log2cap, lat_pack<D>(n1));
    const int s2 = lat_find(L.keys, log2cap, lat_pack<D>(n2));
    int* dst = L.nb + (static_cast<size_t>(j) * m_max + id) * 2;
    dst[0] = s1 < 0 ? -1 : L.ids[s1];
    dst[1] = s2 < 0 ? -1 : L.ids[s2];
  }
}

// in: [n_tiles][N][2] channel pairs (ch0, ch1).  fix accumulators must be zero on entry.
template <int D>
__global__ void lat_splat_kernel(const LatticeTile* __restrict__ tiles, const float* __restrict__ in, int N) {
  const LatticeTile L = tiles[blockIdx.y];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * (D + 1)) return;
  const int i = e / (D + 1);
  const float b = L.bary[e];
  const float2 v = reinterpret_cast<const float2*>(in)[static_cast<size_t>(blockIdx.y) * N + i];
  long long* dst = L.fix + 2 * static_cast<size_t>(L.offset[e]);
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(__float2ll_rn(b * v.x * kLatFix)));
  atomicAdd(reinterpret_cast<unsigned long long*>(dst + 1), static_cast<unsigned long long>(__float2ll_rn(b * v.y * kLatFix)));
}

__global__ void lat_fix2f_kernel(const LatticeTile* __restrict__ tiles) {
  const LatticeTile L = tiles[blockIdx.y];
  const int M = *L.count;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= M) return;
  L.val_a[2 * id] = static_cast<float>(L.fix[2 * id]) * (1.f / kLatFix);
  L.val_a[2 * id + 1] = static_cast<float>(L.fix[2 * id + 1]) * (1.f / kLatFix);
  L.fix[2 * id] = 0;
  L.fix[2 * id + 1] = 0;
}

// one blur pass along axis j: src -> dst (src = val_a for even j, val_b for odd j)
__global__ void lat_blur_kernel(const LatticeTile* __restrict__ tiles, int j, int m_max) {
  const LatticeTile L = tiles[blockIdx.y];
  const int M = *L.count;
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= M) return;
  const float2* src = reinterpret_cast<const float2*>((j & 1) ? L.val_b : L.val_a);
  float2* dst = reinterpret_cast<float2*>((j & 1) ? L.val_a : L.val_b);
  const int* nbp = L.nb + (static_cast<size_t>(j) * m_max + id) * 2;
  const int a = nbp[0], b = nbp[1];
  const float2 c = src[id];
  const float2 va = a >= 0 ? src[a] : make_float2(0.f, 0.f);
  const float2 vb = b >= 0 ? src[b] : make_float2(0.f, 0.f);
  dst[id] = make_float2(c.x + 0.5f * (va.x + vb.x), c.y + 0.5f * (va.y + vb.y));
}

// out: [n_tiles][N][2]
template <int D>
__global__ void lat_slice_kernel(const LatticeTile* __restrict__ tiles, float* __restrict__ out, int N) {
  const LatticeTile L = tiles[blockIdx.y];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  // after D + 1 passes the result sits in val_b when D + 1 is odd, in val_a when it is even
  const float2* val = reinterpret_cast<const float2*>(((D + 1) & 1) ? L.val_b : L.val_a);
  const float alpha = 1.f / (1.f + exp2f(-static_cast<float>(D)));
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int r = 0; r <= D; ++r) {
    const float b = L.bary[i * (D + 1) + r];
    const float2 v = val[L.offset[i * (D + 1) + r]];
    acc.x += b * v.x * alpha;
    acc.y += b * v.y * alpha;
  }
  reinterpret_cast<float2*>(out)[static_cast<size_t>(blockIdx.y) * N + i] = acc;
}

// ---- mean field on two labels.  Per-pixel state: q1 (marginal of label 1), u (2 energies), norm_g, norm_b.
__global__ void mf_init_kernel(const float* __restrict__ p1, long long total, float* __restrict__ u, float* __restrict__ q1,
                               float* __restrict__ ones) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float p = p1[i];
    const float u0 = -logf(fminf(fmaxf(1.f - p, 1e-5f), 1.f)), u1 = -logf(fminf(fmaxf(p, 1e-5f), 1.f));
    u[2 * i] = u0;
    u[2 * i + 1] = u1;
    const float m = fmaxf(-u0, -u1);
    const float e0 = expf(-u0 - m), e1 = expf(-u1 - m);
    q1[i] = e1 / (e0 + e1);
    ones[2 * i] = 1.f;
    ones[2 * i + 1] = 1.f;
  }
}
__global__ void mf_norm_kernel(const float* __restrict__ filtered, long long total, float* __restrict__ norm) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    norm[i] = 1.f / sqrtf(filtered[2 * i] + 1e-20f);
}
// in = norm * (1 - q1, q1)
__global__ void mf_scale_kernel(const float* __restrict__ q1, const float* __restrict__ norm, long long total,
                                float* __restrict__ in) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float n = norm[i], q = q1[i];
    in[2 * i] = n * (1.f - q);
    in[2 * i + 1] = n * q;
  }
}
// t = -u + w_g norm_g F_g + w_b norm_b F_b; q = softmax(t)
__global__ void mf_update_kernel(const float* __restrict__ u, const float* __restrict__ fg, const float* __restrict__ ng,
                                 float wg, const float* __restrict__ fb, const float* __restrict__ nbn, float wb,
                                 long long total, float* __restrict__ q1, uint8_t* __restrict__ labels,
                                 float* __restrict__ q1_out) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float t0 = -u[2 * i], t1 = -u[2 * i + 1];
    if (fg) { const float n = ng[i]; t0 += wg * n * fg[2 * i]; t1 += wg * n * fg[2 * i + 1]; }
    if (fb) { const float n = nbn[i]; t0 += wb * n * fb[2 * i]; t1 += wb * n * fb[2 * i + 1]; }
    const float m = fmaxf(t0, t1);
    const float e0 = expf(t0 - m), e1 = expf(t1 - m);
    const float q = e1 / (e0 + e1);
    q1[i] = q;
    if (labels) labels[i] = (q > 1.f - q) ? 1 : 0;       // argmax, ties to label 0 like np.argmax
    if (q1_out) q1_out[i] = q;
  }
}

}  // namespace dp
