// Dihedral-group (D4) index maps for test-time augmentation.
// The reference applies np.fliplr / np.rot90(k) per tile on axes (0,1) (DigiPathAI/helpers/utils.py:487-522).
// A transform T is represented by its SOURCE map  out[i,j] = in[src_T(i,j)], encoded in 3 bits:
//   bit0 = swap (take (j,i)), bit1 = mirror first coordinate, bit2 = mirror second coordinate.
//   identity 0 | fliplr 4 | rot90 5 | rot180 6 | rot270 3          (table in digipathai_b200/tta.py)
// Gather use (stem):  net_in[i,j] = tile[src_G(i,j)]   with G the cumulative forward transform.
// Scatter use (head): out[src_T(h,w)] = pred[h,w]      which equals out = T^-1(pred).
#pragma once

namespace dp {

__host__ __device__ __forceinline__ void d4_src(int code, int i, int j, int P, int& a, int& b) {
  int u = (code & 1) ? j : i;
  int v = (code & 1) ? i : j;
  a = (code & 2) ? (P - 1 - u) : u;
  b = (code & 4) ? (P - 1 - v) : v;
}

// Per-call arguments of one forward pass, resident in device memory so that a captured CUDA graph can be
// replayed for any slide / tile batch / TTA pass: only this 48-byte record is rewritten (stream-ordered copy).
struct PassDesc {
  const unsigned char* slide;  // uint8 [x][y][3]
  long long slide_h;
  long long slide_w;           // extent along x: tile origins are clamped to [0, slide_w - P] x [0, slide_h - P] by the gather
  const int* coords;           // int32 [n_tiles][2]
  float* probs_out;            // float32 [n_tiles][P][P]
  int tta_in, tta_out;
};

}  // namespace dp
