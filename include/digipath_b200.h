/*
 * digipath_b200.h -- C ABI of the B200-native replacement for DigiPathAI's tile-segmentation hot path.
 *
 * The reference (haranrk/DigiPathAI, pure Python on TensorFlow 1.x) has exactly one operator seam on this path:
 *
 *     prediction = models[model_name].predict(image_patches, batch_size=batch_size, ...)
 *                                                            DigiPathAI/Segmentation.py:154-156
 *
 * surrounded by numpy glue for TTA (DigiPathAI/helpers/utils.py:487-522), the overlap accumulate
 * (Segmentation.py:164-173), normalise (:175-177) and threshold (:336-337).  Each entry point below replaces
 * one of those call sites; the citation on each says which.  The reference-side binding a maintainer would
 * add (a ctypes stub inside DigiPathAI/Segmentation.py) is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; every pointer documented as "device" is a CUDA device pointer owned by the caller
 *     (e.g. a torch tensor's data_ptr()); "host" pointers are ordinary host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Calls are asynchronous on
 *     that stream unless stated otherwise. No hidden allocation happens after dp_model_create/dp_model_reserve.
 *   - every function returns 0 on success, non-zero on failure; dp_last_error() returns a thread-local,
 *     human-readable message for the last failure on the calling thread.
 *   - orientation follows the reference: planes and tiles are indexed [x][y] (width-major), see
 *     DigiPathAI/loaders/dataloader.py:357-358 and Segmentation.py:116-129.
 */
#ifndef DIGIPATH_B200_H
#define DIGIPATH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_model dp_model;

/* Library/ABI version (bumped when a signature changes). */
int dp_abi_version(void);

/* Thread-local message of the last failed call on this thread ("" if none). */
const char* dp_last_error(void);

/*
 * Replaces load_trained_models(model, path, patch_size)      DigiPathAI/helpers/utils.py:427-448
 * `blob` is a host pointer to a flat "DPB1" model container (layer program + packed fp16 weights + folded
 * BatchNorm affine terms) produced by digipathai_b200.weights.pack_*; see DESIGN.md "Model container".
 * Allocates all activation buffers for up to `max_batch` tiles of `patch` x `patch` on CUDA device `device`.
 */
int dp_model_create(const void* blob, size_t nbytes, int device, int max_batch, dp_model** out);

/*
 * A second execution lane of `src`: a model that SHARES the container's data section (weights, BatchNorm vectors) with
 * `src` on the device and owns its own activation buffers, per-call argument record and captured graphs.  Lanes of one
 * model may run dp_forward_tiles concurrently on different streams: consecutive tile batches (or the TTA passes of one
 * batch, the `for transform_index in range(...)` loop of Segmentation.py:150-160) are independent, and the kernels of
 * one batch fill the SMs another batch's small-map layers leave idle.  The reference has no counterpart (it runs one
 * `Model.predict` at a time, Segmentation.py:154).  Destroy every lane with dp_model_destroy; the data section is
 * freed with its last user.  Options set on `src` so far are copied.
 */
int dp_model_clone(const dp_model* src, dp_model** out);

/* Frees everything owned by the model (synchronises the device first). */
int dp_model_destroy(dp_model* m);

/* Model facts: patch side the container was built for, max batch, bytes of HBM held. */
int dp_model_info(const dp_model* m, int* patch, int* max_batch, uint64_t* device_bytes);

/*
 * Arithmetic the container was packed for: 0 = fp16 weights / activations with fp32 accumulation on the tensor
 * cores (the configuration BASELINE.json names), 1 = fp32 weights, activations and accumulation (the mode that
 * reproduces the reference's fp32 Model.predict, Segmentation.py:154-156, within 1e-3; slower). -1 for a null model.
 */
int dp_model_precision(const dp_model* m);

/*
 * Replaces  apply_tta(image_patches, tta_) -> models[name].predict(image_patches) -> transform_prob(pred, tta_)
 *                                        Segmentation.py:150-158, utils.py:487-522, dataloader.py:340-390
 * for ONE pass over ONE batch, including the tile crop + (v-128)/128 normalisation of __getitem__.
 *
 *   slide      device, uint8 [slide_w][slide_h][3] raster in [x][y][c] order (the reference transposes every
 *              tile to this orientation, dataloader.py:357-358). For a pre-gathered batch of tiles pass the
 *              tiles as [n_tiles*P][P][3] with slide_h = P and coords[b] = (b*P, 0).
 *   coords     device, int32 [n_tiles][2] = (x, y) level-0 tile origins (already clamped, dataloader.py:348-353)
 *   tta_in     D4 code (see dp_d4_*) of the CUMULATIVE forward transform the network input has undergone
 *              (the reference applies TTAs in place, so pass k sees T_k o ... o T_1; Segmentation.py:151)
 *   tta_out    D4 code of the transform whose inverse is applied to the prediction (transform_prob)
 *   probs_out  device, float32 [n_tiles][P][P]: softmax channel 1 (the only channel the reference consumes,
 *              Segmentation.py:167), already inverse-transformed.
 * A model serves one stream at a time: the per-call arguments live in one device-side record that is rewritten
 * (stream-ordered) before each captured graph replay.
 */
int dp_forward_tiles(dp_model* m, const uint8_t* slide, int64_t slide_w, int64_t slide_h, const int32_t* coords,
                     int n_tiles, int tta_in, int tta_out, float* probs_out, void* stream);

/*
 * Replaces the statistics + overlap accumulate of one batch          Segmentation.py:162-173
 *   probs      device, float32 [n_pass][n_tiles][P][P] (n_pass = |tta_list| * |models| results of dp_forward_tiles)
 *   mean,var   device, float32 planes [x_hi - x_lo][plane_h]; count device uint8, same shape.
 *              x_lo is the level-0 x of plane row 0 (0 for a whole-slide plane; >0 for a rank's stripe).
 * Adds np.mean / np.var (ddof 0) over the pass axis into the planes and 1 into count (uint8, wraps), tile by
 * tile in ascending tile order -- bit-identical to the reference's sequential loop for identical probs.
 */
int dp_stitch(const float* probs, int n_pass, int n_tiles, int patch, const int32_t* coords, float* mean,
              float* var, uint8_t* count, int64_t plane_w, int64_t plane_h, int64_t x_lo, void* stream);

/*
 * Replaces normalise + threshold                              Segmentation.py:175-177 and :336-337
 *   count==0 -> 1; mean /= count; var /= count^2 (in place);  label[i] = mean[i] >= threshold ? 255 : 0
 * `label` (device uint8, n elements) may be NULL to skip the threshold. All pointers 16-byte aligned.
 */
int dp_finalize(float* mean, float* var, uint8_t* count, int64_t n, float threshold, uint8_t* label, void* stream);

/* One 2x mean-pool level of an [w][h] float32 plane into [w/2][h/2] (in-HBM probability pyramid). */
int dp_pyramid_down2(const float* in, int64_t w, int64_t h, float* out, void* stream);

/*
 * JPEG tile encoder of the pyramidal result files.  Replaces the ImageMagick pass of the reference --
 * `convert <p> -compress jpeg -quality 90 -define tiff:tile-geometry=256x256 ptif:<p>`    DigiPathAI/Segmentation.py:333-334,
 * 345-346, 351-352 (helpers/convert_to_pyramidal.py:32-37) -- for one pyramid level resident in HBM:
 * `plane` device uint8 [rows][cols]; tiles of 256 x 256 (border replicated at the plane's edges), row-major tile indices
 * tile0 .. tile0 + n_tiles - 1.  `tables_host`: the packed table block (quantiser reciprocals in natural order + DC / AC
 * Huffman codes) the host side builds from the JPEG header it writes in front of every tile (tiffio._jpeg_tables).
 * Per tile t: `out + t * out_cap` receives the entropy-coded scan (byte-stuffed, padded; no markers), sizes[t] its length;
 * flags[t] bit 0 = constant tile (value in bits 8..15; nothing written), bit 1 = stream larger than the capacities (nothing
 * written: encode it on the host).  `workspace`: device, dp_jpeg_encode_workspace_bytes(n_tiles, scratch_bytes_per_tile).
 * dp_jpeg_compact gathers the fixed-stride streams into one buffer at `offsets` (device int64, exclusive scan of sizes).
 */
size_t dp_jpeg_encode_workspace_bytes(int n_tiles, int scratch_bytes_per_tile);
int dp_jpeg_encode_gray_tiles(const uint8_t* plane, int64_t rows, int64_t cols, int tile0, int n_tiles, const void* tables_host,
                              size_t tables_bytes, void* workspace, size_t workspace_bytes, int scratch_bytes_per_tile,
                              uint8_t* out, int out_cap, int32_t* sizes, int32_t* flags, void* stream);
int dp_jpeg_compact(const uint8_t* in, int cap, const int32_t* sizes, const int64_t* offsets, uint8_t* out, int n_tiles,
                    void* stream);


/* D4 helpers shared with the host code: source map of transform `code` on a P x P tile. */
void dp_d4_src(int code, int i, int j, int P, int* a, int* b);

/* ---- instrumentation / debugging (not on the hot path) ------------------------------------------------ */

/* Number of kernels this library has launched since load (all models, all threads). */
uint64_t dp_kernel_launch_count(void);

/* Options: "naive_conv" (0/1: evaluate convs with the CUDA-core reference kernel),
 *          "desc_base_mode" (0/1: UMMA descriptor base_offset policy, bring-up only), "halo_pad8" (0/1: pad halo pitch to 8 pixels, bring-up only),
 *          "profile" (0/1: record CUDA events around every op for dp_model_op_times; implies direct launches),
 *          "use_graph" (0/1, default 1: replay dp_forward_tiles as one captured CUDA graph),
 *          "use_pdl" (0/1, default 1: conv kernels use programmatic dependent launch),
 *          "use_overlap" (0/1, default 1: a dense layer consumes the channels older than its predecessor's output
 *                         before its grid-dependency wait, overlapping consecutive layers),
 *          "b_pair" (0/1, default 0: 2-CTA clusters with TMA-multicast weight tiles; measured no gain on B200, where
 *                    L2 already de-duplicates unicast requests of up to 4 neighbouring SMs -- kept for experiments),
 *          "b_resident" (0/1, default 1: layers whose weights fit keep them in shared memory for the whole kernel),
 *          "epi_direct" (0/1, default 1: epilogue writes 256-bit vectors from registers instead of staging in smem),
 *          "split" (default 1: number of sub-batches captured as parallel graph branches). */
int dp_model_set_option(dp_model* m, const char* key, int value);

/* Number of ops / buffers in the layer program; buffer geometry (per-image H, W, C; fp16 NHWC). */
int dp_model_program_size(const dp_model* m, int* n_ops, int* n_bufs);
int dp_model_buffer_shape(const dp_model* m, int buf, int* h, int* w, int* c);

/* Copy an activation buffer (first n_tiles images) to / from host fp16 memory; synchronous. */
int dp_debug_read_buffer(dp_model* m, int buf, int n_tiles, void* host_fp16, size_t nbytes);
int dp_debug_write_buffer(dp_model* m, int buf, int n_tiles, const void* host_fp16, size_t nbytes);

/* Run ops [op_begin, op_end) of the layer program on n_tiles images already resident in the buffers.
 * Head ops write to probs_out (may be NULL if the range has no head). */
int dp_debug_run_ops(dp_model* m, int n_tiles, int op_begin, int op_end, int tta_out, float* probs_out,
                     void* stream);

/* Per-op device time (ms, CUDA events around each op's launches) of the last run made with option
 * "profile" = 1; `n` must equal the number of ops.  dp_model_op_info describes op `op` of the layer program
 * (type: 1 stem gather/im2col, 2 maxpool, 3 conv, 4 bn/pool; conv kind: 1 = 1x1, 3 = 3x3, 4 = upsample+3x3). */
int dp_model_op_times(dp_model* m, float* ms, int n);
int dp_model_op_info(const dp_model* m, int op, int* type, int* kind, int* cin, int* cout, int* h, int* w,
                     uint64_t* macs_per_tile);

/* Debug timeline of CTA 0 of the conv op selected with option "trace_op" (direct-launch path only):
 * out[r] = entries of role r (0 producer, 1 MMA, 2 epilogue, 3 transform, 4 setup); role r's entries start at
 * out[8 + 2000 r], each event<<48 | item<<32 | clock32. */
int dp_debug_read_cta_stamps(dp_model* m, int op, unsigned long long* out, int n);   /* option "stamp_ctas": [256][4] */
int dp_debug_read_trace(dp_model* m, unsigned long long* out, int n);

/* %globaltimer (ns) at entry / exit of CTA 0 of every tensor-core op of the last direct-launch run made with
 * option "stamp" = 1: out[2*op], out[2*op+1]; n must be 2 * number of ops. */
int dp_debug_read_stamps(dp_model* m, unsigned long long* out, int n);

/* Executed tensor-core MACs of one forward pass over n_tiles tiles (after the sub-pixel rewrite). */
int dp_model_executed_macs(const dp_model* m, int n_tiles, uint64_t* macs);

/*
 * Replaces TissueMaskGenerationOS                                  DigiPathAI/helpers/utils.py:336-354
 * (the part that touches pixels).  `rgb` is the slide's lowest pyramid level, device uint8 [n_pix][3].
 * dp_tissue_hist fills `hist` (device uint32 [768 + 65536]): the 256-bin histograms of R, G and B followed by the
 * joint 256 x 256 histogram of (max(R,G,B), max - min), row = max.  The caller derives the four Otsu thresholds from
 * them on the host (256-bin float64 arithmetic, skimage.filters.threshold_otsu) and calls dp_tissue_mask with the
 * three channel thresholds, the `> 50` floor and a 65 536-entry table `sat_lut[max * 256 + (max - min)]` = 1 where the
 * HSV saturation (max-min)/max exceeds its threshold.  mask (device uint8 [n_pix]) = 1 for tissue, else 0.
 */
int dp_tissue_hist(const uint8_t* rgb, int64_t n_pix, uint32_t* hist, void* stream);
int dp_tissue_mask(const uint8_t* rgb, int64_t n_pix, int thr_r, int thr_g, int thr_b, int rgb_min,
                   const uint8_t* sat_lut, uint8_t* mask, void* stream);

/*
 * Replaces cv2.dilate / cv2.erode with a k x k rectangular kernel (np.ones((k, k))), the building block of
 * BinMorphoProcessMaskOS                                           DigiPathAI/helpers/utils.py:200-219
 * (close 20 = dilate, erode; open 5 = erode, dilate; dilate 60 / 35 / 10).  in / out / tmp: device uint8 [n0][n1];
 * OpenCV's default anchor (k/2) and border rule (cells outside the image are ignored); out may alias in.
 */
int dp_morph_rect(const uint8_t* in, uint8_t* out, uint8_t* tmp, int n0, int n1, int k, int dilate, void* stream);

/*
 * Fully connected CRF refinement of probability tiles -- replaces `post_process_crf(image, probs, 2)`
 * (DigiPathAI/helpers/utils.py:568-603: pydensecrf DenseCRF, unary_from_softmax(clip=1e-5), Gaussian pairwise
 * sdims (10,10) compat 3, bilateral sdims (50,50) schan (20,20,20) compat 10, DIAG_KERNEL, NORMALIZE_SYMMETRIC,
 * inference(10), argmax; its call site Segmentation.py:327-331 is commented out in the reference).  Mean-field
 * inference with the Gaussian filters evaluated exactly (pydensecrf approximates them on a permutohedral lattice).
 *   rgb   : device uint8 [n_tiles][h][w][3]       p1 : device float32 [n_tiles][h][w] (probability of label 1)
 *   labels: device uint8 [n_tiles][h][w] in {0,1} (may be NULL)   q1_out: device float32 marginal of label 1 (may be NULL)
 *   workspace: device scratch of dp_crf_workspace_bytes(n_tiles, h, w) bytes, owned by the caller.
 */
size_t dp_crf_workspace_bytes(int n_tiles, int h, int w);
int dp_crf_tiles(const uint8_t* rgb, const float* p1, int n_tiles, int h, int w, int n_iter, float sdims_gauss,
                 float compat_gauss, float sdims_bilateral, float schan_bilateral, float compat_bilateral,
                 void* workspace, size_t workspace_bytes, uint8_t* labels, float* q1_out, void* stream);

/*
 * The same inference with the Gaussian filters evaluated on the permutohedral lattice (splat / blur / slice), i.e.
 * the approximation pydensecrf itself uses -- what post_process_crf / do_crf of DigiPathAI/helpers/utils.py:548-603
 * compute through `d.inference(n)`.  Same arguments as dp_crf_tiles; the workspace is larger (hash tables, blur
 * neighbours: dp_crf_lattice_workspace_bytes, ~75 MB per tile for up to 8 tiles processed at a time) and must be
 * 256-byte aligned.  compat_bilateral == 0 leaves the bilateral term (and its lattice) out.  Bit-reproducible.
 */
size_t dp_crf_lattice_workspace_bytes(int n_tiles, int h, int w);
int dp_crf_tiles_lattice(const uint8_t* rgb, const float* p1, int n_tiles, int h, int w, int n_iter, float sdims_gauss,
                         float compat_gauss, float sdims_bilateral, float schan_bilateral, float compat_bilateral,
                         void* workspace, size_t workspace_bytes, uint8_t* labels, float* q1_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DIGIPATH_B200_H */
