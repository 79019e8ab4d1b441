/*
 * instageo_b200.h -- C ABI of libinstageo_b200.so (hand-written sm_100a CUDA).
 *
 * Drop-in boundary for InstaGeo's chip-inference hot path.  The reference has no FFI
 * layer (it is pure Python over PyTorch); every entry point below cites the reference
 * Python interface it replaces (paths relative to the reference repo root).  The Python
 * host mirror that binds these with ctypes lives in
 * instageo-e2e-geospatial-ml_b200/_lib.py; INTEGRATION.md shows the stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every data pointer is a DEVICE pointer owned by the caller (PyTorch allocates);
 *     the library owns only the opaque ig_model (packed bf16 weights).
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing executes
 *     on any other stream, no device-wide synchronisation in the forward path.  ig_model_forward
 *     is CUDA-graph capturable by the caller; left alone it replays its own captured graph
 *     (see "Forward schedule cache" below).
 *   - return 0 on success, negative IG_E* on failure; ig_last_error() gives a
 *     thread-local message.  Never aborts, never prints.
 *   - no CPU fallback: on a device that is not compute capability 10.x every compute
 *     entry point returns IG_EARCH.
 */
#ifndef INSTAGEO_B200_H
#define INSTAGEO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IG_OK 0
#define IG_EINVAL (-1)  /* bad argument */
#define IG_ESHAPE (-2)  /* unsupported / inconsistent shape */
#define IG_ECUDA (-3)   /* CUDA runtime or driver error */
#define IG_ENOMEM (-4)  /* allocation failed / workspace too small */
#define IG_EARCH (-5)   /* device is not sm_100 */
#define IG_ESTATE (-6)  /* call order violated (e.g. forward before finalize) */

/* element types of caller buffers */
#define IG_F32 0
#define IG_BF16 1
#define IG_I16 2
#define IG_U16 3
#define IG_F64 4 /* ig_preprocess raw input only: already-scaled float64 host arrays */
#define IG_I64 5 /* label element types of the metric kernels */
#define IG_I32 6
#define IG_U8 7
#define IG_I8 8

/* masking strategies, instageo/data/data_pipeline.py:254-267 */
#define IG_MASK_EACH 0
#define IG_MASK_ANY 1

int ig_version(void);
const char* ig_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Kernel 1: fused raster -> chip preprocessing.
 * Replaces, per chip / window:
 *   get_raster_data band gather            instageo/model/dataloader.py:700-703
 *   arr_x * constant_multiplier (float64)  instageo/model/dataloader.py:741
 *   arr_x == no_data_value                 instageo/model/dataloader.py:899
 *   process_and_augment(crop=False) ->
 *   normalize_and_convert_to_tensor        instageo/model/dataloader.py:495-524, 527-585
 *   process_test window crops              instageo/model/dataloader.py:618-669
 *   apply_mask / decode_fmask_value        instageo/data/data_pipeline.py:229-267,
 *                                          instageo/data/hls_utils.py:77-86
 *
 * raw        [n_img, n_src_bands, H, W] int16|uint16 (or f32|f64, the reference's post-multiply
 *            arrays) selected by raw_dtype, strides in ELEMENTS:
 *            img_stride, band_stride, row_stride (pixel stride is 1).
 * band_idx   [T*C] int32 (device) : source band of output (t, c) = band_idx[t*C + c].
 * win_yx     [n_win, 3] int32 (device): (image index, top, left) of each win x win window;
 *            NULL => n_win = n_img whole images with H == W == win.
 * out_f32    [n_win, C, T, win, win] float32 or NULL          (reference layout)
 * out_patch  [n_win*T*(win/16)^2, C*256] bf16 or NULL        (tubelet rows for kernel 2)
 * mask_elem  [n_win, T*C, win, win] uint8 or NULL  : f64(raw)*cm == no_data_value
 * mask_px    [n_win, win, win] uint8 or NULL       : OR of mask_elem over T*C
 * fmask      [n_img, T, H, W] uint8 or NULL (same H/W geometry, dense); bits in fmask_bits
 *            (bit p set => decode position p) mark pixels that are replaced by
 *            no_data_value BEFORE scaling (strategy each|any).
 * mean/std   [C] float32 (device).
 */
int ig_preprocess(const void* raw, int raw_dtype, int n_img, int n_src_bands, int H, int W,
                  int64_t img_stride, int64_t band_stride, int64_t row_stride,
                  const int32_t* band_idx, int T, int C, const int32_t* win_yx, int n_win, int win,
                  double constant_multiplier, const float* mean, const float* std, int has_nodata,
                  double no_data_value, const uint8_t* fmask, uint32_t fmask_bits,
                  int masking_strategy, float* out_f32, void* out_patch, uint8_t* mask_elem,
                  uint8_t* mask_px, void* stream);

/* Tile-level nodata map for the sliding-window path: out[y - y0, x] = 1 when ANY selected band /
 * timestep of tile pixel (y, x) is nodata by the element test of ig_preprocess (Fmask-flagged
 * pixels replaced by no_data_value before scaling, float64(raw) * constant_multiplier ==
 * no_data_value; instageo/model/dataloader.py:741, :899), for rows [y0, y1) of ONE image
 * raw [n_src_bands, H, W] (strides in elements, pixel stride 1; W need not be a multiple of 8).
 * fmask [T, H, W] uint8 or NULL.  out [y1 - y0, W] uint8.  This is the nodata_px input of ig_stitch. */
int ig_nodata_map(const void* raw, int raw_dtype, int n_src_bands, int H, int W, int64_t band_stride,
                  int64_t row_stride, const int32_t* band_idx, int T, int C,
                  double constant_multiplier, int has_nodata, double no_data_value,
                  const uint8_t* fmask, uint32_t fmask_bits, int masking_strategy, int y0, int y1,
                  uint8_t* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Kernel 5: overlap-averaging stitch of sliding-window logits (specification: SURVEY.md
 * Appendix A.6; the reference snapshot only has the window grid, dataloader.py:655-664, the
 * per-chip argmax, infer_utils.py:96-101, and a non-overlapping gdal_merge mosaic,
 * instageo/new_apps/backend/app/cog_converter.py:122-134).
 *
 * win_logits [n_win, nc, win, win] float32; ys [ny], xs [nx] int32 (device) sorted window
 * origins, window (iy, ix) is entry iy*nx + ix - win_base of win_logits (windows outside
 * [win_base, win_base + n_win) must not cover rows [y0, y1)).
 * Output rows [y0, y1) of the H x W tile: class_map [y1-y0, W] int8, optional avg
 * [nc, y1-y0, W] float32, optional class histogram hist [nc+1] uint64 (last = nodata).
 * nodata_px [y1-y0, W] uint8 or NULL: the SAME rows [y0, y1) as the outputs (ig_nodata_map writes
 * exactly this); non-zero pixels become nodata_class.
 */
int ig_stitch(const float* win_logits, int n_win, int win_base, int nc, int win,
              const int32_t* ys, int ny, const int32_t* xs, int nx, int H, int W, int y0, int y1,
              const uint8_t* nodata_px, int nodata_class, float* avg, int8_t* class_map,
              unsigned long long* hist, void* stream);

/* ---------------------------------------------------------------------------------------
 * Kernels 2-4: PrithviSeg forward (instageo/model/model.py:292-419,
 * instageo/model/pritvhi.py:370-530, timm==1.0.20 Block).
 */
typedef struct ig_model ig_model;

typedef struct ig_model_cfg {
  int embed_dim;    /* D: 768 | 1024 (256 for prithvi_eo_tiny) */
  int depth;        /* encoder blocks */
  int num_heads;    /* D / 64 */
  int temporal;     /* T (num_frames) */
  int num_classes;  /* 1 = regression head */
  int img_size;     /* 224 */
  int patch_size;   /* 16 */
  int in_chans;     /* 6 */
  int head_dims[5]; /* embed_dims of the segmentation head, model.py:380-383 */
} ig_model_cfg;

int ig_model_create(const ig_model_cfg* cfg, ig_model** out);
/* key = state_dict key of the reference PrithviSeg ("prithvi_encoder.blocks.0.attn.qkv.weight",
 * "segmentation_head.0.0.weight", ...); data = float32 DEVICE pointer, contiguous. Unknown
 * keys that the reference forward never reads (e.g. *_embed_enc.scale, num_batches_tracked)
 * return IG_OK and are ignored. */
int ig_model_load_weight(ig_model* m, const char* key, const float* data, const int64_t* shape,
                         int ndim, void* stream);
/* bf16 repack, BatchNorm fold, ConvTranspose phase split, channel permutation. */
int ig_model_finalize(ig_model* m, void* stream);
size_t ig_model_workspace_bytes(const ig_model* m, int batch);
/* x: [B, C, T, 224, 224] float32 (x_dtype IG_F32), or tubelet rows written by
 * ig_preprocess(out_patch) (x_dtype IG_BF16).  logits [B, nc, 224, 224] float32 or NULL;
 * argmax [B, 224, 224] int8 or NULL = torch.argmax(logits, dim=1) fused into the head epilogue
 * (instageo/model/infer_utils.py:96-101); feats [B, D*T, 14, 14] float32 or NULL. */
int ig_model_forward(ig_model* m, const void* x, int x_dtype, int batch, float* logits,
                     int8_t* argmax, float* feats, void* workspace, size_t workspace_bytes,
                     void* stream);
/* PrithviSegmentationModule.predict_step (instageo/model/segmentation.py:202-213):
 * prob_pos [B, 224, 224] float32 = softmax(logits, dim=1)[:, 1], computed in the head epilogue
 * (logits never stored).  num_classes >= 2. */
int ig_model_predict_proba(ig_model* m, const void* x, int x_dtype, int batch, float* prob_pos,
                           void* workspace, size_t workspace_bytes, void* stream);
/* number of kernels one ig_model_forward runs (f32 entry; the bf16 tubelet-row entry runs one
 * less: no patchify), for bench accounting */
int ig_model_launches_per_forward(const ig_model* m);
/* Forward schedule cache.  The model keeps, per (batch, workspace address), the GEMM plans with
 * their TMA descriptors, the fact that the zero borders of the head buffers inside that workspace
 * have been cleared, and -- for the bf16 tubelet-row entry -- an instantiated CUDA graph of the
 * whole forward (one cudaGraphLaunch per call; x / logits / argmax / prob pointers that change
 * between calls are patched into two kernel nodes).  The workspace contents between forwards
 * therefore belong to the library: if anything else writes into it, or it is freed and the
 * address reused, call ig_model_reset_cache.  Set IG_NO_GRAPH=1 in the environment to force the
 * eager path (identical kernels, launched one by one). */
int ig_model_reset_cache(ig_model* m);
/* returns 1 while graph replay is enabled, 0 once a CUDA graph API call failed (the forward then
 * stays on the eager path; `note` receives the reason).  *last_forward_was_graph: 1 if the last
 * forward was a graph launch; *kernels_in_graph: kernel nodes of the last captured graph. */
int ig_model_graph_status(const ig_model* m, int* last_forward_was_graph, int* kernels_in_graph,
                          char* note, size_t note_bytes);
/* debug / parity taps: copy an intermediate activation of the LAST forward as float32.
 *   head maps  "feat" (pre-head features), "convt<i>", "stage<i>": [B, C, H, W], read from the
 *              workspace;
 *   encoder    "x" (residual stream after the last block, from the workspace), and -- only while
 *              a tap buffer is attached -- "embed" (after patch embed + pos + cls), "block<i>"
 *              (after block i), "tokens" (after the final LayerNorm): [B, N, D].
 * ig_model_set_tap_buffer attaches a caller-owned float32 device buffer of at least
 * (depth + 2) * B * N * D elements that every following forward fills (eager path, extra
 * device-to-device copies: test use only); NULL detaches it. */
int ig_model_set_tap_buffer(ig_model* m, float* buf, size_t elems);
int ig_model_debug_tap(ig_model* m, const char* name, int batch, void* workspace, float* dst,
                       size_t dst_elems, void* stream);
int ig_model_destroy(ig_model* m);

/* ---------------------------------------------------------------------------------------
 * Building blocks exposed for parity tests and benchmarks (same kernels the model uses).
 */
/* out = epilogue(A[M,K] * W[N,K]^T + bias); A, W bf16 row-major; bias f32 or NULL.
 * act: 0 none, 1 GELU(erf).  out_dtype IG_BF16 | IG_F32.  If resid != NULL (f32 [M,N]) it is
 * added (out_dtype must be IG_F32; out may alias resid). */
int ig_linear(const void* A, const void* W, const float* bias, const float* resid, void* out,
              int out_dtype, int M, int N, int K, int act, void* stream);
/* x f32 [M, D] -> bf16 [M, D], eps 1e-5 */
int ig_layernorm(const float* x, const float* gamma, const float* beta, void* out, int M, int D,
                 void* stream);
/* qkv bf16 [B*N, 3*D] (timm layout: q | k | v, head-major inside) -> out bf16 [B*N, D] */
int ig_attention(const void* qkv, void* out, int B, int N, int heads, void* stream);

/* ---------------------------------------------------------------------------------------
 * Chip-creation masking at tile scale, one pass instead of the reference's
 * xarray chain in HLSRasterPipeline (instageo/data/hls_utils.py:359-403):
 *   apply_mask + decode_fmask_value   instageo/data/data_pipeline.py:229-267, hls_utils.py:77-86
 *   chip.clip(0, 10000), astype(uint16)                    hls_utils.py:371-373, :386, :401
 *   mask_segmentation_map, astype(int8)                    data_pipeline.py:66-98, hls_utils.py:392-400
 *   the two "skip if nothing is left" counts               hls_utils.py:389, :395
 *
 * chip      [n_bands, H, W] int16 | uint16, dense.   fmask [n_mask_steps, H, W] uint8 or NULL; band b
 *           uses mask step b / (n_bands / n_mask_steps) (strategy each) or the OR of all steps (any);
 *           a pixel is masked when any bit of fmask_bits is set in its Fmask byte, and becomes
 *           no_data_value.  Then values are clipped to [clip_min, clip_max] (clip_min > clip_max: no
 *           clipping) and written to out [n_bands, H, W] (uint16; the input's type when not clipping).
 * seg_map   [H, W] int8 or NULL: pixels whose chip is no_data_value in every band (strategy each) or
 *           in at least one band (any) become seg_no_data_value in seg_out [H, W].
 * counts    [2] uint64 or NULL, accumulated: [0] += chip elements != no_data_value after clipping,
 *           [1] += label pixels != seg_no_data_value.
 */
int ig_chip_mask(const void* chip, int chip_dtype, int n_bands, int64_t height, int64_t width,
                 const uint8_t* fmask, int n_mask_steps, uint32_t fmask_bits, int masking_strategy,
                 int no_data_value, int clip_min, int clip_max, uint16_t* out, const int8_t* seg_map,
                 int seg_masking_strategy, int seg_no_data_value, int8_t* seg_out,
                 unsigned long long* counts, void* stream);

/* ---------------------------------------------------------------------------------------
 * Raw-tile ingestion and prediction writing, device side.  The reference
 * reads rasters with rasterio / GDAL (`src.read()`, instageo/model/dataloader.py:672-704) and writes
 * predictions band by band (instageo/model/infer_utils.py:37-54).  Inflate stays on host threads
 * (zlib into pinned memory); what GDAL does after it runs here:
 *
 * ig_tiff_unpack16: blocks = the inflated TIFF blocks (strips: block_w = W; or tiles) of a 16-bit
 *   raster, block-major -- block ((plane * nby + by) * nbx + bx) at a fixed stride of
 *   block_h * block_w * (planar ? 1 : spp) samples, still horizontally differenced (predictor 2) /
 *   byte-swapped / pixel-interleaved as stored in the file.  Undoes the predictor (wrapping prefix
 *   sum along x per sample, TIFF 6.0 section 14), swaps bytes, de-interleaves, crops partial edge
 *   blocks: dst [spp, H, W] uint16 / int16 (same bits) -- the raster ig_preprocess reads in place.
 * ig_tiff_predict: forward horizontal differencing of `rows` rows of W 1- or 2-byte samples (one band
 *   plane, e.g. int8 class maps [n * H, W]) ahead of the host's Deflate. */
int ig_tiff_unpack16(const void* blocks, void* dst, int W, int H, int spp, int block_w, int block_h,
                     int planar, int predictor, int byteswap, void* stream);
int ig_tiff_predict(const void* src, void* dst, int sample_bytes, int64_t rows, int W, void* stream);

/* ---------------------------------------------------------------------------------------
 * Eval-mode streaming metrics accumulated on the device.
 * Replaces the per-step host round trip of
 *   RunningConfusionMatrix.update   instageo/model/metrics.py:86-108
 *   RunningAUC.update / _bin        instageo/model/metrics.py:209-244
 *   RunningRegressionMetrics.update instageo/model/metrics.py:330-356
 * as driven by PrithviSegmentationModule._shared_step, instageo/model/segmentation.py:117-156
 * (argmax, softmax, ignore_index gather, three D2H copies, np.bincount / np.add.at).
 * All outputs are ACCUMULATED (+=) into caller-owned, zero-initialised device buffers.
 *
 * labels      [n] int64 | int32 | uint8 | int8 (label_dtype IG_I64 ...); has_ignore != 0 drops
 *             elements equal to ignore_index (before any range check, like the reference's mask).
 * matrix      [k*k] uint64, row = truth, column = prediction.
 * counters    [2] uint64: [0] += valid samples (RunningConfusionMatrix.total), [1] += samples whose
 *             label or prediction is outside [0, k) -- the reference raises ValueError there
 *             (np.bincount / reshape); the host mirror raises when it reads a non-zero counter.
 */
int ig_confusion_update(const int8_t* pred, const void* labels, int label_dtype, int64_t n,
                        int num_classes, int has_ignore, int64_t ignore_index,
                        unsigned long long* matrix, unsigned long long* counters, void* stream);
/* Fused eval step from logits f32 [n_img, nc, hw] (NCHW; hw % 4 == 0) and labels [n_img, hw]:
 * first-max argmax -> matrix (may be NULL), float32 softmax over classes -> one-vs-rest score
 * histograms pos_hist / neg_hist [nc, n_bins] uint64 (both NULL to skip), binned with the
 * reference's float32 arithmetic trunc((s - min) / (max - min) * (n_bins - 1)) after clamping.
 * Labels outside [0, nc) that are not ignored count as negatives of every class (as in
 * RunningAUC.update) and in counters[1]. */
int ig_seg_metrics_update(const float* logits, int64_t n_img, int num_classes, int64_t hw,
                          const void* labels, int label_dtype, int has_ignore, int64_t ignore_index,
                          unsigned long long* matrix, unsigned long long* counters, int n_bins,
                          float min_score, float max_score, unsigned long long* pos_hist,
                          unsigned long long* neg_hist, void* stream);
/* RunningAUC.update on given probabilities: scores [n, k] row-major float32 | float64 (binned in the
 * scores' own precision, as numpy >= 2 scalar arithmetic does in RunningAUC._bin), labels [n]. */
int ig_auc_update(const void* scores, int score_dtype, int64_t n, int num_classes, const void* labels,
                  int label_dtype, int n_bins, double min_score, double max_score,
                  unsigned long long* pos_hist, unsigned long long* neg_hist, void* stream);
/* sums [7] float64 += (x, y, xy, x^2, y^2, |y-x|, (y-x)^2) with x = y_true, y = y_pred;
 * counts [2] uint64 += (n, |y-x| <= ee_bias + ee_coef*x evaluated in float32 like the reference).
 * has_ignore drops elements whose y_true == ignore_value (regression.py:154). */
int ig_regression_update(const float* y_true, const float* y_pred, int64_t n, int has_ignore,
                         float ignore_value, float ee_bias, float ee_coef, double* sums,
                         unsigned long long* counts, void* stream);

/* ---------------------------------------------------------------------------------------
 * Measurement aid (bench.py roofline): when enabled, every launch is bracketed by CUDA events on
 * its own stream.  Families: 0 preprocess, 1 stitch, 2 encoder GEMMs, 3 head implicit-GEMM convs,
 * 4 attention, 5 layernorm, 6 other.  ig_profile_report sums and clears (blocks on the events). */
#define IG_PROF_FAMILIES 7
int ig_profile_enable(int on);
int ig_profile_report(double* ms, int* launches, int ncat);
/* the same records one by one, in launch order: ms[i] / family[i] of the first `cap` launches since the
 * last report; returns the number of launches recorded (clears them). */
int ig_profile_report_launches(double* ms, int* family, int cap);

#ifdef __cplusplus
}
#endif
#endif /* INSTAGEO_B200_H */
