/* C ABI of libdigipath_ingest.so: JPEG tile streams -> the HBM-resident [x][y][c] slide raster.
 *
 * Replaces, for slides stored as JPEG-compressed tiled TIFF / SVS, the host-side decode the reference does per patch
 * in its DataLoader workers -- openslide read_region -> PIL -> numpy -> transpose
 * (DigiPathAI/loaders/dataloader.py:239,357-358; 8 CPU processes, Segmentation.py:92) -- SURVEY.md 8(f) row N2.
 * The decode itself is nvJPEG (a library call, like cuBLAS for a plain GEMM); the scatter / transpose into the
 * raster layout the stem gather reads is this library's own kernel.
 *
 * A separate shared object on purpose: libdigipath_b200.so (the forward path) keeps no dependency besides the CUDA
 * driver, and a box without nvJPEG loses only this ingest path.  Plain pointers and sizes only; every function
 * returns 0 on success and a non-zero code with dp_ingest_last_error() set otherwise.
 */
#ifndef DIGIPATH_INGEST_H
#define DIGIPATH_INGEST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_jpeg_decoder dp_jpeg_decoder;

int dp_ingest_abi_version(void);
const char* dp_ingest_last_error(void);

/* One decoder per device / host thread. */
int dp_jpeg_decoder_create(int device, dp_jpeg_decoder** out);
int dp_jpeg_decoder_destroy(dp_jpeg_decoder* dec);

/* Decodes n self-contained JPEG streams (host pointers) to interleaved RGB.
 *   out_rgb : DEVICE uint8 [n][tile_h][tile_w][3]; a stream smaller than the tile (last strip of a page) fills the
 *             top-left corner of its slot, the rest of the slot is left untouched.
 * Streams larger than tile_w x tile_h, or with a component count other than 1 or 3, are rejected.
 * Work is enqueued on `stream` (a cudaStream_t); the host buffers may be released when the call returns. */
int dp_jpeg_decode_tiles(dp_jpeg_decoder* dec, const uint8_t* const* streams, const size_t* lengths, int n,
                         int tile_w, int tile_h, uint8_t* out_rgb, void* stream);

/* Same, for pages whose TIFF PhotometricInterpretation is RGB (2): `components_are_rgb` != 0 says the three stored
 * components ARE R, G, B (no YCbCr transform; what libtiff writes for RGB input and what OpenSlide / libjpeg decode
 * through the Adobe APP14 marker or the 'R','G','B' component ids).  nvJPEG would colour-transform them, so they are
 * decoded as stored component planes and interleaved on the device.  All three components must be full resolution. */
int dp_jpeg_decode_tiles_ex(dp_jpeg_decoder* dec, const uint8_t* const* streams, const size_t* lengths, int n,
                            int tile_w, int tile_h, uint8_t* out_rgb, int components_are_rgb, void* stream);

/* Scatters decoded tiles into the raster stripe the forward path reads.
 *   tiles   : DEVICE uint8 [n][tile_h][tile_w][3]   (image layout: row = y)
 *   origins : DEVICE int32 [n][2], level-0 (x, y) of each tile's top-left pixel
 *   raster  : DEVICE uint8 [x_hi - x_lo][height][3] (the reference's tile orientation [x][y][c],
 *             dataloader.py:357-358), covering slide columns [x_lo, x_hi)
 * Pixels outside [x_lo, x_hi) x [0, height) -- tile padding beyond the slide edge, or columns another GPU owns --
 * are dropped. */
int dp_scatter_tiles_xy(const uint8_t* tiles, int n, int tile_w, int tile_h, const int32_t* origins, uint8_t* raster,
                        int64_t x_lo, int64_t x_hi, int64_t height, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIGIPATH_INGEST_H */
