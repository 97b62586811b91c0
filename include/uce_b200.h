/*
 * uce_b200.h — C ABI of the B200-native UCE hot path (libuce_b200.so).
 *
 * The reference (rohitgandikota/unified-concept-editing) has no FFI of its own: its hot
 * path is Python calling torch ops.  This header is the tensor-level seam that sits exactly
 * where trainscripts/uce_sd_erase.py:45-82 (and the identical block in
 * trainscripts/uce_sd_debias.py:114-140) do their arithmetic.  Every entry point names the
 * reference lines it replaces.
 *
 * Conventions
 *   - plain C: pointers, sizes, a cudaStream_t passed as void*; no torch types.
 *   - all matrices are row-major, contiguous, fp32 (reference: torch_dtype=float32,
 *     uce_sd_erase.py:117).
 *   - return value: 0 = ok, <0 = bad argument / state (UCE_E_*), >0 = cudaError_t.
 *     uce_last_error() returns a thread-local description of the last failure.
 *   - the caller owns every buffer; the library owns only the opaque workspace.
 *   - *_dev entry points enqueue on the caller's stream and never synchronise;
 *     *_host entry points take HOST buffers, do their own H2D/D2H and return when the
 *     outputs are valid.
 *   - a workspace is single-threaded; distinct workspaces are independent.
 *
 * Algebra: with C = [edit rows; preserve rows], G = [guide rows; preserve rows],
 * S = diag(scales), the reference's per-projection result is
 *       W_new = (lamb W + sum_i s_i (W g_i) c_i^T) (lamb I + sum_i s_i c_i c_i^T)^-1
 *             = W + (W E^T) Q,        E = G_e - C_e  [n_edit,K],
 *                                     Q = S_e C_e (lamb I + C^T S C)^-1  [n_edit,K]
 * Q depends on neither the projection nor W, so it is computed once (uce_factor_*), in fp64,
 * through whichever of the two equivalent systems is smaller (n x n "dual" when n <= K,
 * K x K "primal" otherwise); the per-projection work is two skinny GEMMs (uce_apply_*).
 */
#ifndef UCE_B200_H
#define UCE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UCE_B200_ABI_VERSION 1

/* negative status codes */
#define UCE_E_ARG        (-1)   /* null pointer, non-positive size, n_edit > n_rows ...            */
#define UCE_E_STATE      (-2)   /* apply before factor, workspace too small ...                    */
#define UCE_E_NOT_SPD    (-3)   /* lamb I + C^T S C is not positive definite (negative scales)     */
#define UCE_E_NO_DEVICE  (-4)   /* no CUDA device / not an sm_100 device                           */

typedef struct uce_ws uce_ws;   /* opaque workspace */

int          uce_abi_version(void);
const char  *uce_last_error(void);

/* Create / destroy a workspace on CUDA device `device` for text dimension K (768 for SD-1.x,
 * 2048 for SDXL: the in_features of attn2.to_k/to_v, uce_sd_erase.py:63) and at most
 * `max_rows` concept rows per solve. */
int uce_ws_create(int device, int K, int max_rows, uce_ws **out);
int uce_ws_destroy(uce_ws *ws);

/* Select the apply kernel: 0 = auto (rank pad <= 64: the K-split two-kernel tcgen05 apply, 7; any other low-rank edit: the
 * two-GEMM tcgen05 apply, 5; dense K x K factor: SIMT), 1 = SIMT fp32 (validation twin),
 * 4 = tcgen05, fused one-kernel form, two row blocks per CTA in one planned wave (rank pad <= 64, K % 128 == 0),
 * 5 = tcgen05 two-GEMM apply, P through an HBM scratch (any rank pad, K % 32 == 0),
 * 7 = tcgen05 K-split apply: kernel A (partial products W E^T per K slice) + kernel B (update), rank pad <= 64, K % 32 == 0.
 * Returns the previous value. All are hand-written CUDA (3xTF32, fp32 fidelity); there is no CPU path. */
int uce_ws_set_apply_impl(uce_ws *ws, int impl);

/* Host-only helper (no GPU needed): the row-block plan the two-block apply (impl 4) uses for `n_layers` projections of
 * d[l] rows (out_features of attn2.to_k / to_v, uce_sd_erase.py:15-22) on a GPU with `sm_count` SMs.  A CTA owns two
 * blocks of block_rows[l] rows of projection l; first_cta[l] is the index of its first CTA.  Returns the number of CTAs
 * (> 0) or a negative UCE_E_* code.  UCE_TC3_BLOCK_ROWS in the environment overrides the search (debugging aid). */
int uce_plan_row_blocks(int sm_count, const int *d, int n_layers, int *block_rows, int *first_cta);

/* Select the factor path: 0 = auto (single-CTA low-latency kernel when n <= 160 rows and the dual system
 * applies, general blocked path otherwise), 1 = always the general blocked path. Returns the previous value. */
int uce_ws_set_factor_impl(uce_ws *ws, int impl);

/* Debug mode keeps a copy of the assembled system matrix for uce_ws_debug_read(…, 0, …).
 * Returns the previous value. */
int uce_ws_set_debug(uce_ws *ws, int on);

/* Profiling mode brackets the factor and each apply kernel with CUDA events on the launching
 * stream; uce_ws_timings() returns the last durations in milliseconds (synchronises the events):
 * ms[0] = factor, ms[1] = apply stage 1 (P = W E^T, or the whole apply when fused/dense),
 * ms[2] = apply stage 2 (W_new = W + P Q; 0 when there is no second kernel). */
int uce_ws_set_profile(uce_ws *ws, int on);
int uce_ws_timings(uce_ws *ws, float ms[3]);

/* Phase 1 — shared factor.  Replaces the `mat2` accumulation (uce_sd_erase.py:63,71,79) and
 * its inversion (:82), done ONCE instead of once per projection, plus the concept-row part
 * of the `mat1` accumulation (:70,78).
 *   C      [n_rows,K]  concept rows; rows [0,n_edit) are the edit concepts (zip order of
 *                      edit_concepts, :66), rows [n_edit,n_rows) the preserve concepts (:74)
 *   G      [n_edit,K]  guide rows paired with the edit rows (:66-68); preserve rows guide
 *                      themselves (:75-76)
 *   scales [n_rows]    erase_scale for edit rows, preserve_scale for preserve rows (:70-79);
 *                      HOST pointer (n_rows floats), rows with scale 0 are ignored
 *   lamb               regularisation (:61,63)
 * Device pointers for C and G; enqueued on `stream`. */
int uce_factor_dev_f32(uce_ws *ws, const float *C, const float *G, const float *scales_host,
                       int n_rows, int n_edit, float lamb, void *stream);

/* Phase 2 — per-projection apply, batched over projections.  Replaces the guide outputs
 * v* = W_old c (uce_sd_erase.py:45-53), mat1 = lamb W_old + sum s v* c^T (:61,70,78) and
 * mat1 @ inverse(mat2) (:82) for `n_layers` projections.
 *   W_old[l]  [d[l],K] device, read-only (the reference's `original_modules`, :21)
 *   W_new[l]  [d[l],K] device, written    (the reference's `uce_modules[...].weight`, :82)
 * W_new[l] may equal W_old[l] (in place). Pointer arrays and d[] are HOST arrays. */
int uce_apply_dev_f32(uce_ws *ws, const float *const *W_old, float *const *W_new, const int *d,
                      int n_layers, void *stream);

/* Factor + apply in ONE call on device buffers (same arguments as the two calls above, same results).  Knowing both halves lets the
 * library take work off the critical path: the first kernel of the K-split apply (the partial products W_old E^T — the reference's
 * guide outputs v* = W_old c, uce_sd_erase.py:45-53 — which need only E = G_e - C_e) runs on an internal stream WHILE the factor
 * computes Q; only the update kernel waits for the factor.  Everything is ordered after the work already enqueued on `stream`
 * and `stream` waits for all of it (safe under CUDA-graph capture: the internal stream joins and leaves the capture). */
int uce_edit_dev_f32(uce_ws *ws, const float *C, const float *G, const float *scales_host, int n_rows, int n_edit, float lamb,
                     const float *const *W_old, float *const *W_new, const int *d, int n_layers, void *stream);

/* Whole edit with HOST buffers (the call a host-side integration makes): H2D of C, G and each
 * W_old[l], factor, apply, D2H of each W_new[l]; copies are pipelined against the kernels on
 * internal streams; returns after the last W_new byte has landed.  Same arguments as above but
 * every pointer is a host pointer (pinned memory gives full PCIe bandwidth).  Projections whose host
 * buffers follow each other in memory (W_old[l + 1] == W_old[l] + d[l] * K, likewise W_new: a caller
 * that keeps its weights in one pinned arena) travel as one copy per pipeline group and direction. */
int uce_edit_host_f32(uce_ws *ws, const float *C, const float *G, const float *scales,
                      int n_rows, int n_edit, float lamb,
                      const float *const *W_old, float *const *W_new, const int *d, int n_layers);

/* Blocks until everything enqueued on `stream` through this workspace is done and reports
 * deferred numerical failures (UCE_E_NOT_SPD). */
int uce_ws_check(uce_ws *ws, void *stream);

/* Introspection (tests, bench): mode 0 = none, 1 = dual (n x n), 2 = primal (K x K);
 * rank = rows of E/Q actually used (0 with dense != 0 means the single dense K x K factor
 * D = E^T Q is applied instead); sys_n = padded system size; launches = kernels enqueued by
 * the last factor + apply. */
int uce_ws_info(uce_ws *ws, int *mode, int *rank, int *dense, int *sys_n, int *launches_factor,
                int *launches_apply);

/* Debug read-back of intermediate device state into host memory (tests only).
 * which: 0 = H/B system matrix as assembled [sys_n*sys_n f64], 1 = Cholesky factor L (lower)
 * [sys_n*sys_n f64], 2 = Q [rank*K f32], 3 = E [rank*K f32], 4 = D^T [K*K f32] (dense only).
 * Synchronises the device. `cap_bytes` is the size of `out`. */
int uce_ws_debug_read(uce_ws *ws, int which, void *out, size_t cap_bytes);

/* ---- the UCE artifact: host-only, no GPU needed ----------------------------------------------------------------------
 * The reference stores ONLY the edited attn2.to_k / to_v weights, fp32, key = module path + ".weight", with
 * safetensors.torch.save_file (trainscripts/uce_sd_erase.py:85-88; uce_sd_debias.py writes the same way) and reads them back with
 * load_file + load_state_dict(strict=False) (evalscripts/generate-images-sd.py:17-19).  uce_artifact_write_f32 produces the
 * byte-identical file for the same dictionary (n two-dimensional fp32 tensors data[i] of rows[i] x cols[i], HOST pointers — pinned
 * staging buffers can be written without another copy); the reader opens any safetensors file and reads its F32 tensors. */
typedef struct uce_artifact uce_artifact;
int uce_artifact_write_f32(const char *path, int n, const char *const *names, const float *const *data, const long *rows, const long *cols);
int uce_artifact_open(const char *path, uce_artifact **out);
int uce_artifact_count(const uce_artifact *a);
/* name / dtype point into the handle (valid until close); shape receives ndim <= 8 extents */
int uce_artifact_entry(const uce_artifact *a, int i, const char **name, const char **dtype, int *ndim, long shape[8]);
int uce_artifact_read_f32(uce_artifact *a, int i, float *dst, size_t cap_elems);
int uce_artifact_close(uce_artifact *a);

/* ---- PNG writer for the generated images (host-only) ------------------------------------------------------------------
 * The reference saves every image of a CSV row as {case_number}_{num}.png with PIL (evalscripts/generate-images-sd.py:45-46).
 * rgb: H x W x 3 bytes, row-major.  level 0..9 (zlib; other values = 6).  threads <= 0 = all host cores: the image is deflated in
 * stripes of rows in parallel and stitched into one zlib stream.  Lossless: any PNG reader returns rgb exactly. */
int uce_png_write_rgb8(const char *path, const unsigned char *rgb, int H, int W, int level, int threads);

#ifdef __cplusplus
}
#endif
#endif /* UCE_B200_H */
