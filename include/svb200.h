/*
 * svb200.h — C ABI of libsvb200.so: the B200-native (sm_100a) batched scan path of sudoku-vision.
 *
 * The reference has no FFI on this path: its boundary is a set of Python functions that call
 * OpenCV / PyTorch.  Each entry point below replaces the body of one of those
 * functions (cited as reference file:line, relative to the reference tree) for a BATCH of n
 * frames / cells.  The Python drop-in modules (sudoku-vision_b200/dropin/) bind these with ctypes
 * and keep the reference's signatures; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *  - Plain C types only.  Unless a parameter is named host_*, every pointer is a DEVICE pointer
 *    to memory owned by the caller; the library never frees or retains caller memory.
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the call returns
 *    without synchronising (the *_host convenience calls synchronise before returning).
 *  - Return value: 0 = SVB_OK, negative = error (svb_last_error() gives the text).  Per-frame
 *    outcomes such as "no quadrilateral found" (cv/grid.py:71 returns None) are DATA (found[n]),
 *    not errors.
 *  - Images are uint8, row-major, tightly packed: frames [n][h][w][3] in BGR order (cv2.imread
 *    layout), masks / gray [n][h][w].  Cells are [n][81][28][28], row-major by (row, col).
 *  - There is no CPU fallback: every entry point launches CUDA kernels or fails.
 */
#ifndef SVB200_H
#define SVB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_OK 0
#define SVB_ERR_INVALID (-1)      /* bad argument (null pointer, non-positive size, ...)        */
#define SVB_ERR_UNSUPPORTED (-2)  /* valid in the reference, not implemented here (non-default
                                     ksize / block_size ...): the shim raises NotImplementedError */
#define SVB_ERR_CUDA (-3)         /* a CUDA runtime call failed                                   */
#define SVB_ERR_NOT_LOADED (-4)   /* classifier weights were not loaded                           */

typedef struct svb_ctx svb_ctx;

/* ---- context ------------------------------------------------------------------------------ */
/* One context per (host thread, device).  Owns scratch buffers and packed classifier weights.   */
int svb_create(int device, svb_ctx **out);
void svb_destroy(svb_ctx *ctx);
const char *svb_last_error(void);
int svb_abi_version(void);
/* Number of kernels this library has launched through `ctx` since creation (bench.py's
 * gpu_launches claim is read from here). */
long long svb_launch_count(const svb_ctx *ctx);

/* Per-stage device timing of svb_scan_batch_v1 (bench.py's roofline numbers are taken INSIDE the
 * timed region with these): when enabled, CUDA events are recorded on the launching stream between
 * the stages of every scan.  svb_last_stage_ms waits for the last scan's final event and returns the
 * elapsed milliseconds of its stages: [0] K1 fused preprocess, [1] K2 contour (reset+probe+select),
 * [2] K3+K4 homography + cells, [3] K5 convolution stack (conv1 + conv2 + pooling), [4] K5 fc1 + fc2 + softmax
 * (+ not-found masking). */
/* Context options.  SVB_OPT_OVERLAP (default 0 = all stages in order on the caller's stream): a value p >= 2 makes
 * svb_scan_batch_v1 cut a batch of >= 64 frames into p parts that alternate between two internal streams, so that the
 * latency-bound contour stage of one part can run under the other part's kernels.  Results are bit-identical either way.
 * Measured on B200 (1024 x 1080p, p = 4): 8.42 ms against 7.51 ms in order — the bandwidth- and tensor-bound kernels
 * each fill the machine on their own, so interleaving them costs more than hiding K2 gains; hence off by default.
 * Stage timing (below) implies the in-order form. */
#define SVB_OPT_OVERLAP 1
int svb_set_option(svb_ctx *ctx, int option, int value);

#define SVB_NUM_STAGES 5
int svb_stage_timing(svb_ctx *ctx, int enable);
int svb_last_stage_ms(svb_ctx *ctx, float *ms);

/* ---- P1..P4: cv/preprocess.py ---------------------------------------------------------------- */
/* grayscale(image)  cv/preprocess.py:15-19  (cv2.cvtColor BGR2GRAY, 15-bit fixed point) */
int svb_grayscale(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *gray, void *stream);
/* blur(image, ksize=5)  cv/preprocess.py:22-29  (cv2.GaussianBlur (5,5), sigma 0, REFLECT_101).
 * ksize != 5 -> SVB_ERR_UNSUPPORTED */
int svb_blur(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, int ksize, uint8_t *out, void *stream);
/* threshold(image, block_size=11, c=2)  cv/preprocess.py:32-54  (adaptiveThreshold GAUSSIAN_C).
 * inverted != 0: THRESH_BINARY_INV (cv/preprocess.py:51); 0: THRESH_BINARY (pipeline/run.py:87-88).
 * block_size != 11 or c != 2 -> SVB_ERR_UNSUPPORTED */
int svb_adaptive_threshold(svb_ctx *ctx, const uint8_t *gray, int n, int h, int w, int block_size, int c,
                           int inverted, uint8_t *out, void *stream);
/* preprocess_for_grid_detection(image)  cv/preprocess.py:57-65: one fused kernel, BGR -> mask */
int svb_preprocess_v1(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *mask, void *stream);

/* ---- V1: cv/preprocess_v2.py ------------------------------------------------------------------------- */
/* Frame sides: at least 32 px, at most 3990 px; sides that do not divide by 8 get OpenCV's REFLECT_101-extended CLAHE tile grid.  `channels` = 3 (BGR frames) or 1 (already gray: the reference's grayscale()
 * passes 2-D input through, cv/preprocess_v2.py:33-37).
 * info (optional): uint8 [n][4] = {has_glare, has_shadow, method (0 adaptive, 1 otsu, 2 sauvola), Otsu level}.
 *
 * preprocess_for_grid_detection(image, use_illumination_norm=True, use_shadow_removal=True)  cv/preprocess_v2.py:205-244:
 * detect_glare + detect_shadow, remove_shadow on the frames that have one, normalize_illumination (elliptical close with
 * k = max(h,w)//10 | 1, >= 51), CLAHE 8x8, GaussianBlur 5, adaptive threshold 11/2, 3x3 close + 2x2 open. */
int svb_preprocess_v2(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, int use_illumination_norm,
                      int use_shadow_removal, uint8_t *mask, uint8_t *info, void *stream);
/* preprocess_multi_strategy(image)  cv/preprocess_v2.py:247-308 -> PreprocessResult: binary = the best-scoring of the
 * adaptive / Otsu / Sauvola candidates (required); gray, enhanced, illumination_normalized: uint8 [n][h][w], each optional. */
int svb_preprocess_multi_v2(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, uint8_t *binary,
                            uint8_t *gray, uint8_t *enhanced, uint8_t *illumination_normalized, uint8_t *info,
                            void *stream);
/* The individual functions of cv/preprocess_v2.py on gray / binary images uint8 [n][h][w] (any size the operation
 * itself allows).  `arg` is the operation's parameter where it has one; `info` as above (only the bytes the op sets). */
#define SVB_V2_NORMALIZE_ILLUMINATION 1 /* normalize_illumination(gray)           :40-60                          */
#define SVB_V2_DETECT_GLARE 2           /* detect_glare(gray, threshold=arg|250)   :63-81   dst = glare mask (opt) */
#define SVB_V2_DETECT_SHADOW 3          /* detect_shadow(gray)                     :84-102  dst = shadow mask (opt)*/
#define SVB_V2_REMOVE_SHADOW 4          /* remove_shadow(gray)                     :105-119                        */
#define SVB_V2_CLAHE8 5                 /* apply_clahe(gray, 2.0, 8)               :122-129                        */
#define SVB_V2_OTSU 6                   /* threshold_otsu(gray)                    :146-149 info[3] = level        */
#define SVB_V2_SAUVOLA 7                /* threshold_sauvola(gray, 25, 0.2)        :152-175                        */
#define SVB_V2_CLEANUP 8                /* morphological_cleanup(binary, 3, 2)     :178-202                        */
#define SVB_V2_DILATE_ELLIPSE 9         /* cv2.dilate(gray, MORPH_ELLIPSE (arg,arg)), arg odd <= 399               */
#define SVB_V2_ERODE_ELLIPSE 10         /* cv2.erode(...)                                                          */
#define SVB_V2_BOX_BLUR 11              /* cv2.blur(gray, (arg,arg))               :89                             */
#define SVB_V2_GAUSS21 12               /* cv2.GaussianBlur(gray, (21,21), 0)      :112                            */
int svb_v2_stage(svb_ctx *ctx, int op, const uint8_t *src, int n, int h, int w, int arg, uint8_t *dst, uint8_t *info,
                 void *stream);

/* ---- G1..G4: cv/grid.py ---------------------------------------------------------------------- */
/* find_grid_contour(binary, min_area_ratio=0.1)  cv/grid.py:37-71 with approximate_polygon's
 * epsilon_ratio (cv/grid.py:24-34; the reference always uses 0.02).
 * corners: int32 [n][4][2] (x, y) in approxPolyDP output order; found: uint8 [n] (0 = None). */
int svb_find_grid_contour(svb_ctx *ctx, const uint8_t *mask, int n, int h, int w, double min_area_ratio,
                          double eps_ratio, int32_t *corners, uint8_t *found, void *stream);

/* find_contours(binary)  cv/grid.py:16-21  (cv2.findContours RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) for ONE mask [h][w]:
 * EVERY top-level contour in cv2's order (reverse raster order of the start pixels), each starting at its component's
 * raster-first pixel.  A debug / tooling helper (cv/test_pipeline.py:21-23); the scan path never materialises this list.
 * Two calls: _count floods the outer background, finds the start pixels and walks every border once; it SYNCHRONISES
 * `stream` and returns the totals through the two host pointers.  _fetch (same ctx, same mask, right after) writes
 * points int32 [n_points][2] = (x, y) and offsets int64 [n_contours + 1] (device pointers; contour i is
 * points[offsets[i] .. offsets[i+1])) asynchronously. */
int svb_find_contours_count(svb_ctx *ctx, const uint8_t *mask, int h, int w, long long *host_n_contours,
                            long long *host_n_points, void *stream);
int svb_find_contours_fetch(svb_ctx *ctx, const uint8_t *mask, int h, int w, int32_t *points, long long *offsets, void *stream);
/* approximate_polygon(contour, epsilon_ratio=0.02)  cv/grid.py:24-34  (cv2.arcLength(closed) and the closed-curve
 * cv2.approxPolyDP of OpenCV 4.13: point-to-segment distances, first-maximum ties, final clean-up pass).
 * contour int32 [n_points][2] with coordinates in [0, 65535]; out int32 [n_points][2] (capacity); n_out (device int32):
 * vertex count, -1 = a coordinate outside [0, 65535], -2 = scratch exhausted. */
int svb_approx_poly_dp(svb_ctx *ctx, const int32_t *contour, int n_points, double epsilon_ratio, int32_t *out, int32_t *n_out,
                       void *stream);
/* v2 contour method: detect_grid_contour(binary, min_area_ratio=0.1)  cv/grid_v2.py:102-128 — as above, but a
 * 4-gon must also pass is_valid_quadrilateral (cv/grid_v2.py:64-95: angles in [45,135] deg, longest side <= 2x
 * shortest) or the search continues with the next contour; corners come back ORDERED (order_points,
 * cv/grid_v2.py:49-61: TL, TR, BR, BL) as int32 [n][4][2] (the reference returns the same values as float32).
 * This is method 1 of detect_grid (cv/grid_v2.py:427-437); the Hough / rotation / Harris fallbacks are not built. */
int svb_detect_grid_contour_v2(svb_ctx *ctx, const uint8_t *mask, int n, int h, int w, double min_area_ratio,
                               int32_t *corners, uint8_t *found, void *stream);
/* order_points + getPerspectiveTransform + warpPerspective  cv/grid.py:74-133 (inset_ratio 0).
 * board: uint8 [n][out_size][out_size][3]; frames with found == 0 produce an all-zero board.
 * `found` may be NULL (all frames valid). */
int svb_warp_perspective(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                         const uint8_t *found, int out_size, uint8_t *board, void *stream);

/* ---- E1: cv/extract.py ----------------------------------------------------------------------- */
/* extract_cells(grid_image, cell_size=28, margin_ratio=0.1)  cv/extract.py:13-56.
 * board: [n][size][size][3] BGR (size must be a multiple of 9); cells: uint8 [n][81][28][28]. */
int svb_extract_cells(svb_ctx *ctx, const uint8_t *board, int n, int size, uint8_t *cells, void *stream);

/* ---- C1/C2: pipeline/run.py ------------------------------------------------------------------ */
/* preprocess_cell (pipeline/run.py:73-95: CLAHE(2.0,(4,4)) + adaptiveThreshold BINARY 11,2)
 * followed by predict_cells' invert + /255 + (x-0.5)/0.5 (run.py:129-135).
 * cells: uint8 [n_cells][28][28].  Outputs (each may be NULL):
 *   thresh  uint8 [n_cells][28][28]  = preprocess_cell's return value (0/255)
 *   pm1     float [n_cells][1][28][28] = the tensor fed to the model (exactly -1 / +1) */
int svb_cell_prep(svb_ctx *ctx, const uint8_t *cells, long long n_cells, uint8_t *thresh, float *pm1,
                  void *stream);
/* is_cell_empty(cell, threshold=0.02)  cv/extract.py:59-79: cv2.threshold(THRESH_BINARY_INV + THRESH_OTSU), countNonZero,
 * (non_zero / total) < threshold.  cells uint8 [n_cells][cell_h][cell_w]; empty uint8 [n_cells] (1 = empty);
 * info (optional) int32 [n_cells][2] = {Otsu level, non-zero count}. */
int svb_is_cell_empty(svb_ctx *ctx, const uint8_t *cells, int n_cells, int cell_h, int cell_w, double threshold, uint8_t *empty,
                      int32_t *info, void *stream);

/* Fused G3+G4+E1+C1+C2 for the batched path: frames + corners -> cells, no 450x450 board in HBM.
 * cells_u8 (optional): uint8 [n][81][28][28] = extract_cells output.
 * cells_pm1 (required): float [n][81][28][28] = model input.  Frames with found == 0 are zero-filled. */
int svb_cells_from_frames(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                          const uint8_t *found, uint8_t *cells_u8, float *cells_pm1, void *stream);
/* The same without the float tensor: the classifier input of the batched path, cells_bits uint32 [n][81][28] — one word
 * per cell row, bit x = 1 <=> pixel (x, y) is +1 (ink), 0 <=> -1; bits 28..31 are 0.  (pipeline/run.py:129-135 maps the
 * thresholded cell to exactly {-1, +1}, so 28 bits per row carry the whole tensor: 112 B per cell instead of 3136.) */
int svb_cells_from_frames_bits(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, const int32_t *corners,
                               const uint8_t *found, uint32_t *cells_bits, void *stream);
/* +-1 float cells [n_cells][28][28] -> bit rows [n_cells][28] (x > 0 <=> bit set) */
int svb_pack_cells_bits(svb_ctx *ctx, const float *cells_pm1, long long n_cells, uint32_t *cells_bits, void *stream);

/* ---- M1/M2: ml/model.py, pipeline/run.py:139-150 --------------------------------------------- */
/* DigitCNN parameters (ml/model.py:22-32), PyTorch layouts, fp32, DEVICE pointers:
 * conv1.weight (32,1,3,3) conv1.bias (32) conv2.weight (64,32,3,3) conv2.bias (64)
 * fc1.weight (128,3136) fc1.bias (128) fc2.weight (10,128) fc2.bias (10).  Packs them into the
 * kernels' layouts inside ctx (replaces load_state_dict at pipeline/run.py:108). */
int svb_digitcnn_load(svb_ctx *ctx, const float *conv1_w, const float *conv1_b, const float *conv2_w,
                      const float *conv2_b, const float *fc1_w, const float *fc1_b, const float *fc2_w,
                      const float *fc2_b, void *stream);
/* DigitCNN.forward (ml/model.py:34-42), eval mode.  x: float [n][1][28][28]; logits: float [n][10].
 * Optional epilogue (pipeline/run.py:141-143): digits uint8 [n] = argmax, conf float [n] =
 * softmax(logits)[argmax]; either may be NULL. */
int svb_digitcnn_forward(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits,
                         float *conf, void *stream);
/* DigitCNN.forward on +-1 cells given as bit rows (svb_cells_from_frames_bits / svb_pack_cells_bits): conv1 + bias of a
 * pixel is looked up by its 9-bit neighbourhood pattern (512 x 32 table built by svb_digitcnn_load) instead of computed;
 * everything downstream is the same tcgen05 path.  Same outputs as svb_digitcnn_forward. */
int svb_digitcnn_forward_bits(svb_ctx *ctx, const uint32_t *cells_bits, long long n, float *logits, uint8_t *digits, float *conf,
                              void *stream);

/* ---- M3: ml/model_v3.py ------------------------------------------------------------------------------ */
/* DigitCNNv3 (ml/model_v3.py:95-184), eval mode.  `folded` is a HOST array of 38 DEVICE pointers to fp32 tensors
 * with BatchNorm already folded into the preceding convolution (w' = w * gamma / sqrt(var + 1e-5),
 * b' = beta - mean * gamma / sqrt(var + 1e-5); the Python shim does this from the reference's state_dict):
 *   0 stem.w (32,1,3,3)  1 stem.b (32)
 *   per layer L = 1..5 in order: conv1.w, conv1.b, conv2.w, conv2.b (PyTorch OIHW), se.excite.0.weight (c/4,c),
 *   se.excite.2.weight (c,c/4), and for L = 2, 4 additionally shortcut.w (cout,cin,1,1), shortcut.b
 *   then fc.weight (10,128), fc.bias (10). */
int svb_digitcnn_v3_load(svb_ctx *ctx, const float *const *folded, int count, void *stream);
/* forward(x, return_features) ml/model_v3.py:163-184: x float [n][1][28][28] -> logits float [n][10];
 * optional digits uint8 [n] / conf float [n] (run_v2.py:166-170 softmax-max / argmax) and features float [n][128]
 * (the return_features=True branch). */
int svb_digitcnn_v3_forward(svb_ctx *ctx, const float *x, long long n, float *logits, uint8_t *digits, float *conf,
                            float *features, void *stream);

/* Classifier implementation used by svb_digitcnn_forward and svb_scan_batch_v1:
 * 0 (default) = tcgen05/TMEM implicit GEMM with fp16 hi/lo split operands (fp32-grade logits),
 * 1 = plain fp32 on the CUDA cores (kept as the on-device cross-check). */
int svb_set_classifier_mode(svb_ctx *ctx, int mode);

/* ---- whole path: pipeline/run.py:257-318 (CV + ML sections) for n frames ------------------------ */
/* Outputs: digits uint8 [n][81], conf float [n][81], logits float [n][81][10] (optional),
 * corners int32 [n][4][2], found uint8 [n].  Frames with found == 0 get digits 0 / conf 0. */
int svb_scan_batch_v1(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                      float *logits, int32_t *corners, uint8_t *found, void *stream);
/* Same with HOST buffers (pinned or pageable): copies frames H2D, runs the path, copies results
 * D2H and synchronises.  This is the end-to-end call bench.py times as `e2e`. */
int svb_scan_batch_v1_host(svb_ctx *ctx, const uint8_t *host_bgr, int n, int h, int w, uint8_t *host_digits,
                           float *host_conf, int32_t *host_corners, uint8_t *host_found);

/* ---- frame ingest: cv2.imread's decode step, pipeline/run.py:250 ------------------------------------------ */
/* n JPEG files back to back in HOST memory: file i is host_jpeg[host_offsets[i] .. host_offsets[i+1]) (n + 1 offsets).
 * Headers are parsed on the host; the compressed bytes are copied to the device and decoded there (Huffman decode, the
 * libjpeg "islow" integer IDCT, h2v2 fancy chroma upsampling, fixed-point YCbCr -> BGR: the arithmetic of the libjpeg-turbo
 * inside cv2, so the frames equal cv2.imdecode's bit for bit) into bgr, DEVICE [n][h][w][3].
 * Supported: baseline / extended-sequential 8-bit Huffman JPEG, YCbCr 4:2:0 or 4:4:4 or gray, one scan, all files h x w.
 * Anything else -> SVB_ERR_UNSUPPORTED, malformed headers -> SVB_ERR_INVALID; nothing is decoded approximately.
 * The parallelism is (files x restart intervals): files written without DRI decode with one thread each.
 * status (optional, device uint8 [n]): bit 0 = the file holds fewer / more restart markers than its DRI header implies. */
int svb_jpeg_decode_host(svb_ctx *ctx, const uint8_t *host_jpeg, const long long *host_offsets, int n, int h, int w, uint8_t *bgr,
                         uint8_t *status, void *stream);
/* svb_scan_batch_v1_host fed compressed frames: copy (compressed) + decode + whole path + boards back, chunked over two
 * internal streams; synchronous.  ~10-20x fewer PCIe bytes per frame than raw BGR. */
int svb_scan_batch_v1_jpeg_host(svb_ctx *ctx, const uint8_t *host_jpeg, const long long *host_offsets, int n, int h, int w,
                                uint8_t *host_digits, float *host_conf, int32_t *host_corners, uint8_t *host_found);

/* ---- cv/grid_quality.py ---------------------------------------------------------------------------------- */
/* assess_grid_quality(image, binary, corners)  cv/grid_quality.py:228-306.  frames: BGR (channels 3) or gray (1);
 * binary: the preprocessed mask uint8 [n][h][w]; corners: int32 [n][4][2] in any order (the contour method's corners are
 * integral); found (optional): frames with found != 1 get zeros.
 * scores: double [n][6] = overall, sharpness (Laplacian variance), contrast (95 % histogram range), completeness (mask
 * coverage of the 20 grid-line bands after the 450x450 warp), geometry, size.  The issue / recommendation strings of
 * QualityScore are host logic on these numbers (drop-in module). */
int svb_assess_grid_quality(svb_ctx *ctx, const uint8_t *frames, int n, int h, int w, int channels, const uint8_t *binary,
                            const int32_t *corners, const uint8_t *found, double *scores, void *stream);

/* ---- whole v2 path: pipeline/run_v2.py:276-330 for n frames ------------------------------------------ */
/* preprocess_multi_strategy -> detect_grid method 1 (contour + is_valid_quadrilateral; the Hough / rotation / Harris
 * fallbacks of cv/grid_v2.py:446-508 are not built) -> assess_grid_quality and the min_quality_score gate
 * (run_v2.py:300-308; min_quality_score < 0 = `--no-quality-check`, the reference's default is 40) -> warp + 81 cells +
 * preprocess_cell -> DigitCNNv3 -> softmax top-3 (run_v2.py:149-190).  Requires svb_digitcnn_v3_load.
 * quality (optional): double [n][6] as svb_assess_grid_quality.  found: 0 no grid, 1 scanned, 3 grid found but below
 * min_quality_score (status 'quality_failed').
 * Outputs: digits uint8 [n][81] / conf float [n][81] (best class), alt_digits uint8 [n][81][2] / alt_conf float
 * [n][81][2] (2nd and 3rd, optional), logits float [n][81][10] (optional), corners int32 [n][4][2] ordered TL,TR,BR,BL,
 * found uint8 [n], info uint8 [n][4] as svb_preprocess_multi_v2 (optional).  Frames with found == 0 get zeros. */
int svb_scan_batch_v2(svb_ctx *ctx, const uint8_t *bgr, int n, int h, int w, uint8_t *digits, float *conf,
                      uint8_t *alt_digits, float *alt_conf, float *logits, int32_t *corners, uint8_t *found, uint8_t *info,
                      double *quality, double min_quality_score, void *stream);

/* ---- after the path: solve_sudoku (solver/src/sudoku.c:72-87), batched -------------------------------------- */
/* Replaces run_solver's one-subprocess-per-image call (pipeline/run.py:163-202) for n recognised boards at once.
 * grids: uint8 [n][81], 0 = empty (the digits output of svb_scan_batch_*); solutions: uint8 [n][81] (the input grid
 * when the puzzle is not solved, as run_solver returns); status: int8 [n] = 1 SOLVE_SUCCESS, 0 SOLVE_NOSOLUTION,
 * -1 SOLVE_INVALID (solver/include/sudoku.h:12-15).  Same search order as the reference: identical results for
 * every input, including grids with several solutions. */
int svb_solve_batch(svb_ctx *ctx, const uint8_t *grids, int n, uint8_t *solutions, int8_t *status, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SVB200_H */
