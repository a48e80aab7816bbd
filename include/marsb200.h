/* marsb200 - C ABI of the B200-native MARS proposal scoring / ranking / merging stage.
 *
 * The reference (a pure-Python/PyTorch repo) has no FFI; the surface a
 * maintainer would bind is the arithmetic inside its Python classes.  Every
 * entry point below names the reference code it replaces (paths relative to
 * the reference checkout).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host;
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library
 *    never allocates, frees or synchronises, and enqueues work only on
 *    `stream` (a cudaStream_t passed as void*), so a whole episode batch can
 *    be captured into a CUDA graph;
 *  - all arrays are dense row-major; a leading E is the episode batch;
 *  - return value: 0 on success, a negative MARSB200_ERR_* otherwise, with a
 *    thread-local message available from marsb200_last_error();
 *  - there is no CPU fallback: every call needs an sm_100a device.
 */
#ifndef MARSB200_H
#define MARSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MARSB200_OK 0
#define MARSB200_ERR_ARG (-1)
#define MARSB200_ERR_CUDA (-2)
#define MARSB200_ERR_UNSUPPORTED (-3)

/* mask element types accepted by the ingest kernels */
#define MARSB200_MASK_F32 0 /* reference-native float32 0/1 (main_MARS.py:62, Matcher.py:728-729) */
#define MARSB200_MASK_U8 1  /* uint8 / bool, 1 byte per pixel */

/* Back ends.  There is ONE product path per operation: MARSB200_GEMM_TCGEN05 for the contractions and
 * MARSB200_PAIR_AUTO for the intersections (callers pass exactly these; the Python wrappers default to them).  The other
 * values select second implementations of the same arithmetic that exist to pin the tensor-core kernels in the test
 * suite (tests/test_gpu_parity.py compares them bit for bit / within 3e-6): TEST-ONLY, not a dispatch a deployment
 * chooses from - they are all sm_100a CUDA, there is no other target and no CPU path. */
#define MARSB200_GEMM_TCGEN05 0 /* product: TMA + tcgen05 3xTF32 (error-compensated), fp32 accumulate in TMEM */
#define MARSB200_GEMM_SIMT 1    /* test-only: fp32 FFMA tiles, pins the tensor-core kernel */

#define MARSB200_PAIR_AUTO 3 /* product: kind::mxf4 when P <= 256 and HW < 2^24, kind::i8 otherwise (same counts, bit for bit) */
#define MARSB200_PAIR_FP4 2  /* the P <= 256 half of AUTO: bits expanded to e2m1 nibbles, tcgen05 kind::mxf4 with unit scale factors */
#define MARSB200_PAIR_MMA 1  /* the any-P half of AUTO: bits expanded to int8 in shared memory, tcgen05 kind::i8 into TMEM */
#define MARSB200_PAIR_POPC 0 /* test-only: shared-memory tiled AND + popcount, pins the tensor-core kernels */

int marsb200_version(void);
const char* marsb200_last_error(void);
/* SMs a launch on `stream` can use: the partition of the stream's CUDA green context, else the device's SM
 * count.  The persistent tensor-core kernels size their grids with it, so an episode engine can give the HBM-bound
 * mask ingest and the contractions disjoint SM partitions (no reference counterpart: the reference is single-stream,
 * main_MARS.py:54-94).  count_host is a HOST pointer. */
int marsb200_stream_sm_count(void* stream, int* count_host);
/* Caps what marsb200_stream_sm_count reports for `stream` at `cap` SMs (0 removes the cap; process-wide table keyed by
 * the stream handle, thread-safe; remove the cap before destroying the stream).  A persistent kernel launched on a
 * capped stream starts at most `cap` CTAs and leaves the other SMs to kernels of other streams: the single-episode
 * schedule uses it so that the contractions do not displace the mask ingest running beside them (no reference
 * counterpart, main_MARS.py:54-94 is single-stream). */
int marsb200_stream_set_sm_cap(void* stream, int cap);
/* words (uint32) per packed mask row for an H*W mask: ceil(HW/32) rounded up to 32 words (128 B) */
int64_t marsb200_words_per_mask(int64_t hw);
/* padded extents used by the contraction operands */
int64_t marsb200_pad_rows(int64_t rows);  /* -> multiple of 128 */
int64_t marsb200_pad_k(int64_t k);        /* -> multiple of 32  */

/* ---- A1: row L2-normalise into the contraction's operand layout ---------------------------
 * Replaces F.normalize(feats, p=2, dim=1): VisualVisualAlignmentModule.py:124-125,
 * matcher/Matcher.py:297-298.  x [E, rows, k] fp32 -> out [E, pad_rows(rows), pad_k(k)] fp32, rows divided by
 * max(||row||, 1e-12), and out_lo of the same shape = out - (out with its low 13 mantissa bits cleared): the
 * residual the error-compensated tensor-core product needs (kind::tf32 reads only the top 19 bits of `out`).
 * Padding is zero-filled.  If `normalize` is 0 the rows are only copied into the layout. */
int marsb200_normalize_rows(const float* x, int64_t ld_x, int E, int64_t rows, int64_t k, int normalize, float* out,
                            float* out_lo, void* stream);

/* ---- A3/A6: adaptive max pool of masks to the patch grid ---------------------------------
 * Replaces F.adaptive_max_pool2d(mask, (g, g)) > 0: VisualVisualAlignmentModule.py:72-76,
 * FilteringMergingModule.py:73-76.  masks [n, H, W] -> out [n, g*g] uint8 0/1. */
int marsb200_pool_mask(const void* masks, int mask_dtype, int64_t n, int H, int W, int g, uint8_t* out, void* stream);

/* ---- A2 + A3: S = Fs Fq^T with fused masked column statistics -----------------------------
 * Replaces torch.matmul at VisualVisualAlignmentModule.py:69 and the two recomputed products
 * + max/mean reductions at :78-102, and matcher/Matcher.py:437-440.
 * a, a_lo [E, pad_rows(M), pad_k(K)], b, b_lo [E, pad_rows(N), pad_k(K)] fp32 from marsb200_normalize_rows.
 * sim_out / cost_out [E, M, N] are optional (NULL to skip); cost = (1 - S) / 2.
 * row_fg [E, M] uint8 (pooled support mask, shot-major) and colstats [E, tiles_m, 4, N] are optional
 * (both or neither): per 128-row tile, the per-column {max over fg rows, sum over fg rows,
 * max over bg rows, sum over bg rows}; -inf / 0 when a tile has no such row. tiles_m = pad_rows(M)/128.
 * The tcgen05 back end is an error-compensated 3xTF32 product (fp32-grade accuracy, fp32 accumulate in TMEM):
 * a*b ~ a*b_ + a_lo*b + a*b_lo with the tensor core truncating a, b to tf32; the SIMT back end ignores *_lo. */
int marsb200_sim_contract(const float* a, const float* a_lo, const float* b, const float* b_lo, int E, int64_t M,
                          int64_t N, int64_t K, float* sim_out, float* cost_out, const uint8_t* row_fg,
                          float* colstats, int backend, void* stream);

/* Row top-k (k <= 8, per support patch over the query patches) and column arg-max (per query patch over the
 * support rows selected by row_mask, or all rows when NULL) of S with warp-shuffle reductions: the
 * mutual-nearest-neighbour candidates of bidirectional matching (north-star kernel 1).  The reference's Matcher
 * matches with two exact assignments (matcher/Matcher.py:443-477); these are its GPU-side candidates, SURVEY.md D1.
 * sim [E, M, N]; row_vals/row_idx [E, M, k]; col_vals/col_idx [E, N]; either output pair may be NULL.
 * Ties resolve to the lowest index. */
int marsb200_match_argmax(const float* sim, const uint8_t* row_mask, int E, int M, int N, int k, float* row_vals,
                          int32_t* row_idx, float* col_vals, int32_t* col_idx, void* stream);

/* Exact rectangular linear-sum assignment: scipy.optimize.linear_sum_assignment(S, maximize=True) of
 * Matcher's forward / reverse patch matching, matcher/Matcher.py:449-450, 471-472.  sim [E, R, C]; row_sel [E, R] and
 * col_sel [E, C] pick the participating rows / columns (NULL = all).  Every element of the smaller side is assigned.
 * t_cap / m_cap bound the number of selected elements on the smaller / larger side the shared-memory state is sized
 * for (<= 0: min(R, C) / max(R, C)); 28 * t_cap + 23 * m_cap bytes must fit in 220 KB (1369 x 6845, the 5-shot
 * reverse matching, does).  row_to_col [E, R]: assigned column per selected row or -1; objective [E]: sum of
 * assigned similarities (fp64); *status: 0 or the size that exceeded the caps.  Shortest augmenting paths
 * (Jonker-Volgenant), one CTA per problem.  Near-square problems (m_cap - t_cap <= max(8, m_cap / 16), e.g. the 5-shot
 * forward matching 1374 x 1369) are padded with zero-cost dummy nodes to a square and solved by multi-source phases
 * with equal-distance waves instead of one search per row (40 bytes of shared-memory state per node, n <= 2048): 3.5x scipy at 1374 x 1369. */
int marsb200_lsap(const float* sim, const uint8_t* row_sel, const uint8_t* col_sel, int E, int R, int C, int maximize,
                  int t_cap, int m_cap, int32_t* row_to_col, double* objective, int32_t* status, void* stream);

/* vva = mean_fg*max_fg - mean_bg*max_bg (bg skipped when there is no bg row), then min-max with eps 1e-7.
 * Replaces VisualVisualAlignmentModule.py:82-102.  out [E, N].  Returns an error when an episode has no fg row
 * only at the Python level (the kernel writes NaN for such an episode). */
int marsb200_vva_finalize(const float* colstats, const uint8_t* row_fg, int E, int64_t M, int64_t N, float* out,
                          void* stream);

/* ---- A4: prior-information refinement ------------------------------------------------------
 * Mean over layers and heads of the patch-to-patch attention, dropping `skip` leading tokens
 * (cls + registers).  Replaces PriorInformationRefinementModule.py:31-45.
 * maps_host: host array of n_maps device pointers, each [heads, T, T] of fp32 (dtype 0) or fp16 (dtype 1);
 * out [N, ld_out] fp32 with N = T - skip. */
int marsb200_attn_mean(const void* const* maps_host, int n_maps, int dtype, int heads, int T, int skip,
                       float* out, int64_t ld_out, void* stream);

/* Bytes of workspace marsb200_pir_refine needs for E episodes of an N x N attention matrix. */
int64_t marsb200_pir_workspace_bytes(int E, int64_t N);

/* Full PIR: box mask from the prior (uint8 quantise, threshold, 8-connected components, clipped
 * boxes), D = A/colsum, D = D/rowsum, R = max(D, D D^T), out = R (R (B*prior)) [== ((R R)*B) prior],
 * optional min-max.  Replaces PriorInformationRefinementModule.py:47-89 and :91-122.
 * prior [E, g*g]; attn [E, N, ld_attn] with N = g*g; out [E, N]; box_out [E, N] uint8 optional. */
int marsb200_pir_refine(const float* prior, const float* attn, int64_t ld_attn, int E, int g, double box_threshold,
                        int apply_minmax, float* out, uint8_t* box_out, void* workspace, int64_t workspace_bytes,
                        int backend, void* stream);

/* The same in stages, so that a scheduler can hoist what does not depend on the prior: MARSB200_PIR_NORMALISE (column /
 * row normalisation of the attention into the operand arrays of the workspace; needs only `attn`), MARSB200_PIR_CONTRACT
 * (G = D D^T; needs only the workspace) and MARSB200_PIR_APPLY (box mask from the prior, the two mat-vecs, optional
 * min-max).  Stages of one refinement must run in this order on the same workspace; marsb200_pir_refine = all three. */
#define MARSB200_PIR_NORMALISE 1
#define MARSB200_PIR_CONTRACT 2
#define MARSB200_PIR_APPLY 4
#define MARSB200_PIR_ALL 7
int marsb200_pir_stages(const float* prior, const float* attn, int64_t ld_attn, int E, int g, double box_threshold,
                        int apply_minmax, float* out, uint8_t* box_out, void* workspace, int64_t workspace_bytes,
                        int backend, int stages, void* stream);

/* The box list behind the box mask: `_scoremap2bbox(scoremap, multi_contour_eval=True)` of
 * PriorInformationRefinementModule.py:91-122.  prior [E, g*g]; boxes_out [E, g*g, 4] int32 (x0, y0, x1, y1) with the
 * reference's clip x1 = min(x + w, g - 1), one per 8-connected component in raster order of its first pixel;
 * count_out [E].  Hole contours (OpenCV RETR_TREE lists them too) are nested in their component's box and never change
 * the mask, so they are not listed. */
int marsb200_scoremap_boxes(const float* prior, int E, int g, double box_threshold, int32_t* boxes_out,
                            int32_t* count_out, void* stream);

/* Nearest-neighbour resize of [E, gs, gs] maps to [E, gd, gd] followed by min-max (eps 1e-7).
 * Replaces mars/MARS.py:77-82. */
int marsb200_resize_minmax(const float* src, int E, int gs, int gd, int apply_minmax, float* out, void* stream);

/* ---- A6/A9 ingest: bit-pack the proposal masks ---------------------------------------------
 * masks [n, HW] (float32 or uint8, pixel set iff value > 0) -> bits [n, words_per_mask(HW)] uint32,
 * bit k of word w = pixel 32*w + k; padding words are zero.  Replaces the per-proposal `m_p > 0`
 * views of FilteringMergingModule.py:77,104-108 and feeds every later mask kernel. */
int marsb200_pack_masks(const void* masks, int mask_dtype, int64_t n, int64_t HW, uint32_t* bits, void* stream);
/* The same for ONE pixel slice of every mask: packed words [word_begin, word_begin + word_count) of bits [n, wpm] from pixels
 * [32 * word_begin, 32 * (word_begin + word_count)) of masks [n, HW] - whole 512-word blocks inside the mask's pixels,
 * 16-byte aligned masks with HW % 4 == 0 (float32) / HW % 16 == 0 (uint8).  Slices that tile [0, wpm) produce exactly the bits
 * of marsb200_pack_masks.  With marsb200_pairwise_inter_slice it lets a single episode's intersections start while most of
 * its masks are still being read (no reference counterpart; the reference holds every mask as a dense tensor). */
int marsb200_pack_masks_slice(const void* masks, int mask_dtype, int64_t n, int64_t HW, int64_t word_begin, int64_t word_count,
                              uint32_t* bits, void* stream);

/* Pooled patch bitmap, pixel area and pooled count of every packed mask.
 * Replaces F.adaptive_max_pool2d(m_p, (g, g)) > 0 in the hot loop, FilteringMergingModule.py:104-108.
 * bits [n, wpm] -> pooled [n, ceil(g*g/32)] uint32, area [n] int32, pooled_count [n] int32. */
int marsb200_pool_packed(const uint32_t* bits, int64_t n, int H, int W, int g, uint32_t* pooled, int32_t* area,
                         int32_t* pooled_count, void* stream);

/* Ingest + pooling in ONE pass over the masks: bits = marsb200_pack_masks(masks) and (pooled, area, pooled_count) =
 * marsb200_pool_packed(bits) - the pooling runs on the packed words while they are still in registers, so the packed
 * bits are never read back (W % 32 == 0 and at most 2048 patches; any other geometry runs the two kernels internally, so
 * the call is always valid and the results are always identical).  Replaces the per-proposal adaptive_max_pool2d loop of
 * FilteringMergingModule.py:103-110 together with the ingest of the proposal tensors (main_MARS.py:62).
 * masks [n, H, W] float32 / uint8; bits [n, wpm]; pooled [n, ceil(g*g/32)]; area, pooled_count [n]. */
int marsb200_pack_pool_masks(const void* masks, int mask_dtype, int64_t n, int H, int W, int g, uint32_t* bits,
                             uint32_t* pooled, int32_t* area, int32_t* pooled_count, void* stream);

/* Per-proposal sums of the vva / vta maps under the pooled bitmap and the pooled size of the
 * proposal union.  Replaces FilteringMergingModule.py:77-81,108-110.
 * pooled [E, P, npw]; vva, vta [E, N]; sum_vva, sum_vta [E, P] fp32; union_count [E] int32. */
int marsb200_region_sums(const uint32_t* pooled, int E, int P, int N, const float* vva, const float* vta,
                         float* sum_vva, float* sum_vta, int32_t* union_count, void* stream);

/* ---- A9: pairwise intersection matrix (builder-defined; no reference code) -----------------
 * inter[e,i,j] = popcount(bits[e,i] & bits[e,j]) as int32 [E, P, P]; the diagonal is the area. */
int marsb200_pairwise_inter(const uint32_t* bits, int E, int P, int64_t words_per_mask, int32_t* inter,
                            int backend, void* stream);
/* The contribution of one pixel slice (packed words [word_begin, word_begin + word_count) of every mask, whole 256-pixel
 * blocks): inter = (accumulate ? inter : 0) + popcount over the slice.  Integer additions: slices that tile the mask sum to
 * exactly marsb200_pairwise_inter's counts in any order.  Tensor-core back ends only (AUTO / FP4 / MMA). */
int marsb200_pairwise_inter_slice(const uint32_t* bits, int E, int P, int64_t words_per_mask, int64_t word_begin,
                                  int64_t word_count, int accumulate, int32_t* inter, int backend, void* stream);

/* Fused ingest: bits = pack(masks) AND inter = pairwise intersections in ONE pass over the masks
 * (float32 masks, P <= 256, tensor-core back end: the masks are read once at HBM rate and the int8 MMAs run
 * underneath the stream).  Other shapes / dtypes / back ends run marsb200_pack_masks followed by
 * marsb200_pairwise_inter internally, so the call is always valid.  masks [E, P, HW]; bits [E, P, wpm];
 * inter [E, P, P]. */
int marsb200_pack_pairwise(const void* masks, int mask_dtype, int E, int P, int64_t HW, uint32_t* bits,
                           int32_t* inter, int pair_backend, void* stream);

/* ---- A7: exact EMD scores on the device ---------------------------------------------
 * score[e,p] = 1 - EMD(uniform 1/T over the fg support rows, uniform 1/M_p over the proposal's pooled patches,
 * cost = C[fg rows][patches] in float64): the transportation LP of ot.emd2 at FilteringMergingModule.py:160-167 and
 * matcher/Matcher.py:1187-1194, solved exactly on integer flows (primal-dual method: multi-source shortest-path
 * phases, all sinks at the minimum distance settled per step), one CTA per proposal, largest problems first.
 * cost [E, m_rows, N] fp32 ((1 - S) / 2 from marsb200_sim_contract); row_fg [E, m_rows] uint8; pooled [E, P, ceil(N/32)].
 * t_cap / m_cap size the FAST PATH (solver state in shared memory): the number of fg rows / of pooled patches per proposal
 * it is laid out for (<= 0: all rows / N; clamped to what fits 200 KB; a smaller m_cap lets more problems share an SM).
 * A problem beyond those caps is not an error: it is queued for a second launch of the same solver whose state lives in a
 * per-CTA slab of the workspace (global memory), sized for ANY problem of the episode shape - so, like ot.emd2, there is no
 * capacity limit (5-shot episodes with large support masks included) up to the 16-bit index range (about 8000 fg rows;
 * beyond it the score is NaN and *status receives the row count, or (1 << 24) + the patch count).  *status: 0 = every
 * problem solved, -1 = internal capacity fault (score NaN).  An empty proposal or empty support scores 1.0 (zero
 * transport).  workspace: 256-byte aligned, marsb200_emd_workspace_bytes(E, P, N, m_rows, t_cap, m_cap) bytes. */
int64_t marsb200_emd_workspace_bytes(int E, int P, int N, int64_t m_rows, int t_cap, int m_cap);
int marsb200_emd_scores(const float* cost, const uint8_t* row_fg, const uint32_t* pooled, int E, int P, int64_t m_rows,
                        int N, int t_cap, int m_cap, void* workspace, int64_t workspace_bytes, double* out,
                        int32_t* status, void* stream);

/* ---- host side of the ingest ------------------------------------------------------------------------------------------
 * The same packing for masks that are still in HOST memory (the reference's proposals are CPU tensors when they reach
 * MARS.predict, main_MARS.py:62-69): a team of host threads (AVX2 when the CPU has it) writes the bit layout of
 * marsb200_pack_masks into a host buffer, so that 1/32 of the bytes cross PCIe.  masks [n, HW] float32 / uint8 and bits
 * [n, marsb200_words_per_mask(HW)] are HOST pointers (pinned or not); threads <= 0 = one per hardware thread.  Blocking; no
 * CUDA call.  A format conversion in front of the copy - nothing is scored on the host. */
int marsb200_host_pack_masks(const void* masks_host, int mask_dtype, int64_t n, int64_t HW, uint32_t* bits_host, int threads);

/* ---- A8: AlphaCLIP cosine scores -----------------------------------------------------------
 * clip[e,p] = img[e,p,:] . txt[e,:].  Replaces img_feats @ text_feats.T, FilteringMergingModule.py:97. */
int marsb200_clip_scores(const float* img, const float* txt, int E, int P, int D, float* out, void* stream);
/* The same with float16 features (AlphaCLIP runs in half precision on a GPU, :189,195): fp32 accumulation, ONE rounding to
 * float16 like a half-precision matmul; out holds the float16 values as float32. */
int marsb200_clip_scores_f16(const void* img_f16, const void* txt_f16, int E, int P, int D, float* out, void* stream);

/* ---- A8 + A10 + A11: fuse, rank, NMS, select ------------------------------------------------
 * score = (minmax(emd) + minmax(clip) + a*pvv+(1-a)*cov + a*pvt+(1-a)*cov) / 4, stable descending
 * rank, optional greedy IoU-NMS over `inter` (skipped when inter is NULL or nms_iou_threshold < 0),
 * static/dynamic threshold selection.  Replaces FilteringMergingModule.py:118-138 and :213-217; the NMS is
 * builder-defined with the semantics of torchvision nms (segment_anything/automatic_mask_generator.py:370-376).
 * emd is fp64 (the reference's 1 - ot.emd2 is a float64) and alpha / thresholds are doubles (Python floats in
 * the reference); the AlphaCLIP min-max is evaluated in the feature dtype exactly as the reference's numpy does:
 * clip_f16 = 0: float32 scores; clip_f16 = 1: `clip` holds float16 values (marsb200_clip_scores_f16) and the min-max and
 * the first addition of the fusion are float16 operations with NumPy's promotion rules (FilteringMergingModule.py:97,
 * 126-136 as run on a GPU, where AlphaCLIP is half precision).
 * Outputs: scores [E,P] fp64 by proposal index; order [E,P] int32 (rank -> index);
 * flags [E,P] uint8 by proposal index (bit0 = kept by NMS, bit1 = selected for the merge);
 * summary [E,4] int32 = {n_kept, n_selected, top_index, n_nonfinite}: n_nonfinite counts proposals whose fused score is
 * NaN / inf (a NaN input score); they rank last and the caller should treat the episode as failed.
 * record: optional [E, record_stride] bytes, see marsb200_record_bytes.
 * nms_bits: optional [E, P, ceil(P/32)] uint32 from marsb200_nms_bitmask (same inter, same threshold): the pairwise
 * suppression relation does not depend on the ranking, so it can be computed by a many-CTA launch as soon as the
 * intersections exist, off the critical path of the one-CTA-per-episode ranking kernel; NULL = built inside (same result). */
int marsb200_fuse_rank(const double* emd, const float* clip, const int32_t* pooled_count, const float* sum_vva,
                       const float* sum_vta, const int32_t* union_count, const int32_t* inter, const uint32_t* nms_bits,
                       int E, int P, double alpha, double static_threshold, double dynamic_threshold,
                       float nms_iou_threshold, int clip_f16, double* scores, int32_t* order, uint8_t* flags,
                       int32_t* summary, uint8_t* record, int64_t record_stride, void* stream);
/* nms_bits[e, i] bit j = IoU(i, j) > nms_iou_threshold for j != i, IoU = inter[i][j] / (inter[i][i] + inter[j][j] -
 * inter[i][j]) in float32 (builder-defined; torchvision-nms comparison semantics). */
int marsb200_nms_bitmask(const int32_t* inter, int E, int P, float nms_iou_threshold, uint32_t* nms_bits, void* stream);
/* Bytes of one episode's result record: order int32[P] | score float32[P] | flags uint8[P rounded up to 4] | summary
 * int32[4].  When `record` is given (row e at record + e * record_stride, 4-byte aligned) marsb200_fuse_rank writes the
 * record itself, so the rows can BE a rank's slice of the all-gather table (the only collective of the path
 * gathers these records; nothing is copied or concatenated before it). */
int64_t marsb200_record_bytes(int P);

/* OR of the selected packed masks and expansion to the float32 [H,W] map the reference returns.
 * Replaces (torch.sum(torch.stack(ranked_masks), 0) > 0).float(), FilteringMergingModule.py:219, and the
 * merges at matcher/Matcher.py:776,823.  merged_bits [E, wpm] and merged_f32 [E, HW] are each optional. */
int marsb200_merge_masks(const uint32_t* bits, const uint8_t* flags, int E, int P, int64_t HW,
                         uint32_t* merged_bits, float* merged_f32, void* stream);

/* ---- A12: Matcher-derived scoring ---------------------------------------------------------
 * Number of matched points falling inside every packed mask (is_in_mask, matcher/Matcher.py:1166-1173,
 * 1198-1199).  points [K,2] int32 (x, y), clipped to the image; out [n] int32. */
int marsb200_points_in_masks(const uint32_t* bits, int64_t n, int H, int W, const int32_t* points, int K,
                             int32_t* out, void* stream);

/* purity = in/max(pooled,1) + 1e-6, coverage = in/K + 1e-6, score = alpha*emd + beta*purity*coverage^exp.
 * Replaces matcher/Matcher.py:1203-1209 and :719-720.  All arrays [n] fp32. */
int marsb200_matcher_scores(const int32_t* points_in, const int32_t* pooled_count, const float* emd, int64_t n, int K,
                            float alpha, float beta, float exp, float* purity, float* coverage, float* scores,
                            void* stream);

/* ---- A13: mask-pooled features and masked similarity statistics (diagnostics of the Matcher ancestry) -------
 * Prototypes: out[e, p, :] = mean over the patches of pooled mask p of feats[e, patch, :] (unnormalised features),
 * matcher/Matcher.py:1076-1079 (`feats[mask].mean(dim=0)`), computed for all masks at once as one
 * masks-by-features contraction [P, N] x [N, C] on the tensor cores (0/1 is exact in tf32; the features use the
 * error-compensated split).  pooled [E, P, ceil(N/32)]; feats [E, N, C]; out [E, P, C]; an empty mask gives NaN.
 * workspace: 256-byte aligned, marsb200_masked_feature_means_workspace_bytes. */
int64_t marsb200_masked_feature_means_workspace_bytes(int E, int P, int N, int C);
int marsb200_masked_feature_means(const uint32_t* pooled, const float* feats, int E, int P, int N, int C, float* out,
                                  void* workspace, int64_t workspace_bytes, int backend, void* stream);

/* mean, max, unbiased std and element count of sim[e][row_mask][:, col_mask]: get_aposteriori_statistics,
 * matcher/Matcher.py:1069-1085.  sim [E, M, N]; masks uint8; out [E, 4] float64 = {mean, max (0 if empty), std, n}. */
int marsb200_masked_sim_stats(const float* sim, const uint8_t* row_mask, const uint8_t* col_mask, int E, int64_t M,
                              int64_t N, double* out, void* stream);

/* out[e, c] = mean over the selected rows of sim[e, :, c]: get_ref_to_target_similarity, matcher/Matcher.py:593-611
 * (mean over the masked support patches of Fq Fs_masked^T).  out [E, N] fp32. */
int marsb200_masked_row_mean(const float* sim, const uint8_t* row_mask, int E, int64_t M, int64_t N, float* out,
                             void* stream);

/* ---- 8f-3: evaluator -------------------------------------------------------------------------
 * Per-sample foreground/background intersection and union areas of a prediction against the
 * ground truth with an optional ignore mask.  Replaces Evaluator.classify_prediction,
 * mars/utils/evaluation.py:12-38.  pred, gt, ignore: [n, HW] fp32; out [n, 4] int32 =
 * {inter_bg, inter_fg, union_bg, union_fg}. */
int marsb200_eval_areas(const float* pred, const float* gt, const float* ignore, int64_t n, int64_t HW, int32_t* out,
                        void* stream);

/* Per-class accumulation of those areas: AverageMeter.update, mars/utils/logger.py:61-66
 * (intersection_buf / union_buf .index_add_(1, class_id, ...)).  areas [n, 4] int32 from marsb200_eval_areas;
 * class_id [n] int64; inter_buf, union_buf [2, nclass] int64 pixel counts (exact; the reference keeps float32
 * buffers).  *status is set to 1 if a class id lies outside [0, nclass).  Ranks combine their buffers with one
 * all-reduce (sum). */
int marsb200_eval_accumulate(const int32_t* areas, const int64_t* class_id, int64_t n, int nclass, int64_t* inter_buf,
                             int64_t* union_buf, int32_t* status, void* stream);

/* AverageMeter.compute_iou, mars/utils/logger.py:69-78: interest [k] int64 class ids;
 * out [2 + k] float64 = {mIoU, FB-IoU, fg IoU of every class of interest}; IoU = inter / max(union, 1). */
int marsb200_eval_iou(const int64_t* inter_buf, const int64_t* union_buf, int nclass, const int64_t* interest, int k,
                      double* out, void* stream);

/* ---- 8f-4: proposal wire format and SAM-AMG post-processing ------------------------------------------
 * Uncompressed COCO RLE (column-major runs; counts of mask m are counts[offsets[m] .. offsets[m+1]), starting
 * with a run of zeros - the output of mask_to_rle_pytorch, segment_anything/utils/amg.py:107-135) decoded into
 * the packed row-major bits of marsb200_pack_masks: the inverse of rle_to_mask (amg.py:138-149) without ever
 * materialising a byte or float mask.  H and W must be multiples of 32 (MARSB200_ERR_UNSUPPORTED otherwise).
 * bits [n, words_per_mask(H*W)]; *status = 1 if some mask's counts do not sum to H*W. */
int64_t marsb200_rle_workspace_bytes(int64_t n, int H, int W);
int marsb200_rle_decode(const int32_t* counts, const int64_t* offsets, int64_t n, int H, int W, uint32_t* bits,
                        void* workspace, int64_t workspace_bytes, int32_t* status, void* stream);

/* XYXY boxes of packed masks, [0,0,0,0] for an empty mask: batched_mask_to_box, amg.py:310-353.  boxes [n, 4] int32. */
int marsb200_mask_boxes(const uint32_t* bits, int64_t n, int H, int W, int32_t* boxes, void* stream);

/* Stability score |logits > t + o| / |logits > t - o|: calculate_stability_score, amg.py:156-176.
 * logits [n, HW] fp32; out [n] fp32; counts [n, 2] int32 receives the two pixel counts. */
int marsb200_stability_score(const float* logits, int64_t n, int64_t HW, float mask_threshold, float threshold_offset,
                             float* out, int32_t* counts, void* stream);

/* Greedy box NMS with torchvision.ops.nms semantics (batched_nms with a single category at
 * segment_anything/automatic_mask_generator.py:284-289, 370-375): boxes [n, 4] fp32 XYXY, scores [n];
 * order [n] int32 = box indices by descending score (ties: lower index first); keep [n] uint8 by box index;
 * *n_keep = number kept.  n <= 16384. */
int64_t marsb200_box_nms_workspace_bytes(int n);
int marsb200_box_nms(const float* boxes, const float* scores, int n, float iou_threshold, int32_t* order, uint8_t* keep,
                     int32_t* n_keep, void* workspace, int64_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MARSB200_H */
