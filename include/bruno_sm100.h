/* libbruno_sm100.so -- C ABI of the B200-native BRUNO hot path.
 *
 * The reference (IraKorshunova/bruno, TF-1.7, pure Python) has no FFI of its own: its boundary is the Python
 * layer protocol of nn_extra_nvp.py / nn_extra_student.py / nn_extra_gauss.py, every method of which expands into
 * stock TensorFlow ops.  Each entry point below replaces the TF-op expansion of the reference method(s) cited next
 * to it (paths relative to the reference tree).  The host-side mirror of the layer classes (bruno_b200/*.py) binds
 * these symbols with ctypes; see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 (BRUNO_OK) or a negative BRUNO_E* code; bruno_last_error() gives the message
 *   - all data pointers are DEVICE pointers owned by the caller.  The library never allocates or frees device (or
 *     pinned host) memory: every buffer, including scratch, is an argument sized by a *_bytes / *_floats query; the one
 *     buffer that is not an argument of each call -- the split-K / column-reduction workspace of the dense-path GEMMs --
 *     is caller-owned too and bound per stream with bruno_workspace_bind (below)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - float tensors are contiguous fp32; images are NHWC like the reference
 *   - entry points may be called concurrently from several host threads on different streams / devices.  Process-wide
 *     state, all of it listed here: (1) the launch counter and the optional CUDA-event profile (bruno_profile_*);
 *     (2) the table of caller-bound workspaces (mutex-protected); (3) one-time per-device kernel attributes
 *     (dynamic shared-memory opt-in, mutex-protected); (4) the A/B switches of the DEBUG section at the end of each
 *     group (bruno_debug_*, bruno_process_set_*_variant): not thread-safe, never touched by the product path, compiled
 *     out with -DBRUNO_NO_DEBUG.
 */
#ifndef BRUNO_SM100_H_
#define BRUNO_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRUNO_OK 0
#define BRUNO_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define BRUNO_ECUDA (-2)   /* CUDA launch / runtime error */
#define BRUNO_EARCH (-3)   /* device is not sm_100 */

const char* bruno_last_error(void);
int bruno_abi_version(void);
/* number of CUDA kernels this library has launched in this process (evidence for bench.py's gpu_launches) */
long bruno_launch_count(void);

/* Workspace of the dense-path GEMMs (bruno_sgemm, bruno_coupling_dense_*, bruno_dense_layer_*): column-reduction partials,
 * the split-K partial tiles of the fallback GEMM paths (the default tcgen05 GEMM reduces its split-K partials inside a
 * thread-block cluster through peer shared memory and does not touch it) and their ticket counters.  The caller allocates
 * bruno_workspace_bytes() bytes of device memory, zero-initialises it once, and binds it to the stream it will launch on.
 * One workspace per stream: two streams that run GEMMs concurrently need two.  Launches on a stream without a bound
 * workspace still work: the cluster GEMM as usual, the fallback GEMMs and the column reductions un-split (same results
 * up to summation order, slower).  Unbind before freeing the memory. */
size_t bruno_workspace_bytes(void);
int bruno_workspace_bind(void* stream, void* workspace, size_t bytes);
int bruno_workspace_unbind(void* stream);

/* Optional CUDA-event timing of every launch of a kernel family on its own stream (off by default; bench.py turns it
 * on for a second pass over the timed region to get the live per-launch duration of the dominant kernel). */
#define BRUNO_PROF_CONV 0         /* conv3x3 forward / dgrad (tcgen05 implicit GEMM) */
#define BRUNO_PROF_WGRAD 1        /* conv3x3 wgrad */
#define BRUNO_PROF_PROCESS_FWD 2  /* t-process / GP scan forward */
#define BRUNO_PROF_PROCESS_BWD 3
#define BRUNO_PROF_KINDS 4
int bruno_profile_enable(int on);
int bruno_profile_read(int kind, double* total_ms, long* launches);

/* ------------------------------------------------------------------------------------------------------------
 * Exchangeable Student-t / Gaussian process over each latent dimension.
 *   kind: BRUNO_TP = nn_extra_student.py:19-149, BRUNO_GP = nn_extra_gauss.py:18-116.
 * ---------------------------------------------------------------------------------------------------------- */
#define BRUNO_TP 0
#define BRUNO_GP 1

/* Number of floats of the per-(step, dim) coefficient table built by bruno_process_table for `T` steps of a
 * `D`-dimensional process (forward part; the backward part with parameter derivatives is 4x that). */
size_t bruno_process_table_floats(int kind, int T, int D, int with_grad);

/* Builds the coefficient table of the closed-form posterior (nn_extra_student.py:85-106 unrolled: after n
 * observations mu_n, k_n, beta_n, nu_n are functions of the prefix sums of (x - mu0) and (x - mu0)^2, see
 * tests_gp_tp/utils_stats.py:87-95) for steps n0 .. n0+T-1, from the RAW variables of the layer:
 *   var = softplus(var_vbl)^2 (student:38), nu = exp(nu_vbl) + min_nu (:46), corr = sigmoid(corr_vbl) (:60).
 * lgamma / softplus / derivatives are evaluated in fp64.  `mask_dim` (may be NULL) is the optional per-dimension
 * multiplier of get_log_likelihood(observation, mask_dim) (student:115-116).  nu_vbl is ignored for BRUNO_GP.
 * with_grad != 0 additionally stores d(coefficient)/d(raw variable) for bruno_process_bwd. */
int bruno_process_table(int kind, const float* var_vbl, const float* nu_vbl, const float* corr_vbl,
                        const float* mu0, const float* mask_dim, float min_nu, int n0, int T, int D,
                        int with_grad, float* table, void* stream);

/* The unrolled loop of config_rnn/bn2_omniglot_tp.py:103-124 in one launch: for t = 0..T-1
 *   lp[b,t]       = get_log_likelihood(z[b,t,:])              (student:108-118 / gauss:89-96)
 *   lp_prior[b,t] = get_log_likelihood_under_prior(z[b,t,:])  (student:120-130 / gauss:98-105)   [NULL: skipped]
 *   update_distribution(z[b,t,:])                             (student:85-106 / gauss:75-87)
 * z [B,T,D].  s1/s2 [B,D] (may both be NULL = fresh reset(), student:81-83) hold the running prefix sums
 * sum(x - mu0), sum((x - mu0)^2) after n0 observations and are updated in place, which is what makes the
 * reference's step-by-step API (update_distribution / get_log_likelihood) expressible on the same kernel.
 * `table` comes from bruno_process_table(kind, ..., n0, T, D).  `scratch` holds >= bruno_process_scratch_floats
 * floats (only used when D is split across blocks). */
size_t bruno_process_scratch_floats(int B, int T, int D);
#ifndef BRUNO_NO_DEBUG
/* DEBUG (process-wide, not thread-safe).  Picks the forward-scan kernel for A/B measurements and the variant-equivalence
 * test: 0 = default (persistent bulk-copy ring kernel when D % 4 == 0 and all pointers are 16-byte aligned,
 * register-staged kernel otherwise), 9 = always the register-staged kernel, 10-14 = ring shapes.  Same results for every value. */
void bruno_process_set_fwd_variant(int variant);
/* Same switch for bruno_process_bwd: 0 = default (ring kernel when eligible), 9 = register-staged, 10-12 = ring shapes. */
void bruno_process_set_bwd_variant(int variant);
#endif
int bruno_process_fwd(int kind, const float* z, const float* mu0, const float* table, float* s1, float* s2,
                      float* lp, float* lp_prior, float* scratch, int B, int T, int D, void* stream);

/* Hand-written backward of bruno_process_fwd (fresh state, n0 = 0): given dlp [B,T] (and dlp_prior, may be NULL)
 * computes dz [B,T,D] and the gradients of the three raw variables dvar_vbl, dnu_vbl, dcorr_vbl [D]
 * (dnu_vbl untouched for BRUNO_GP).  `table` must have been built with with_grad = 1.
 * `scratch` holds >= bruno_process_bwd_scratch_floats floats. */
size_t bruno_process_bwd_scratch_floats(int B, int T, int D);
int bruno_process_bwd(int kind, const float* z, const float* mu0, const float* table, const float* dlp,
                      const float* dlp_prior, float* dz, float* dvar_vbl, float* dnu_vbl, float* dcorr_vbl,
                      float* scratch, int B, int T, int D, void* stream);

/* current_distribution after each prefix (student:85-106): for t = 0..T (T+1 entries, t = number of observations)
 * mu_out/var_out [B,T+1,D], nu_out [T+1,D] (NULL for GP).  Used by the sampling branch (bn2:106-128). */
int bruno_process_posterior(int kind, const float* z, const float* var_vbl, const float* nu_vbl,
                            const float* corr_vbl, const float* mu0, float min_nu, float* mu_out, float* var_out,
                            float* nu_out, int B, int T, int D, void* stream);

/* StudentRecurrentLayer.sample (student:132-149, Bailey polar method) with the two uniform draws supplied:
 * rvs [2,n,B,D]; mu/var [B,D] (row stride `stride_bd` floats between b), nu [D]; out [n,B,D].
 * GP (gauss:107-116): rvs [n,B,D] are standard normals, out = mu + sqrt(var) * eps. */
int bruno_process_sample(int kind, const float* rvs, const float* mu, const float* var, const float* nu,
                         float* out, int n, int B, int D, long stride_b, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Real NVP coupling stack (nn_extra_nvp.py).
 * ---------------------------------------------------------------------------------------------------------- */
/* mask types: CouplingLayerConv.get_mask (nvp:143-160), CouplingLayerDense.get_mask (nvp:199-221); b = 1 marks the
 * pass-through (conditioning) half. */
#define BRUNO_MASK_CHECKERBOARD0 0
#define BRUNO_MASK_CHECKERBOARD1 1
#define BRUNO_MASK_CHANNEL0 2
#define BRUNO_MASK_CHANNEL1 3
#define BRUNO_MASK_EVEN 4
#define BRUNO_MASK_ODD 5

/* Pre-processing.  op: 0 ScaleLayer.forward_and_jacobian (nvp:45-49), 1 LogitLayer.forward_and_jacobian (:26-31),
 * 2 LogitLayer.backward (:33-38), 3 ScaleLayer.backward (:51-55), 4 = 0 then 1 fused, 5 = 2 then 3 fused.
 * x,y [N,D]; ld_in (NULL = zeros), ld_out [N].  conditional_variant != 0 drops the ln(1-alpha) term
 * (nn_extra_nvp_conditional.py:13). */
int bruno_pre(int op, const float* x, const float* ld_in, float* y, float* ld_out, int N, int D, float scale, float alpha,
              int conditional_variant, void* stream);

/* SqueezingLayer (nvp:277-298): tf.space_to_depth / depth_to_space with block 2.  H, W = the LARGE resolution;
 * inverse = 0: in [N,H,W,C] -> out [N,H/2,W/2,4C]; inverse = 1: in [N,H/2,W/2,4C] -> out [N,H,W,C]. */
int bruno_space_to_depth(const float* in, float* out, int N, int H, int W, int C, int inverse, void* stream);

/* FactorOutLayer slices / concats (nvp:313-319, :336-344): dst[r, dst_off + c] = src[r, src_off + c], c < nch. */
int bruno_channel_copy(const float* src, float* dst, long rows, int src_C, int src_off, int dst_C, int dst_off, int nch,
                       void* stream);

/* --- conv coupling layer, CouplingLayerConv (nvp:58-187) with function_s_t_wn (:108-141) -------------------------
 * Hidden activations (64 channels) live in bf16 "slabs" (layout: bruno_b200/csrc/slab.cuh); bruno_slab_bytes gives
 * the size of one, which the caller allocates ZERO-INITIALISED once (kernels keep the padding rows zero). */
size_t bruno_slab_bytes(int N, int H, int W);
size_t bruno_conv_wpack_bytes(void);
#ifndef BRUNO_NO_DEBUG
/* DEBUG (process-wide, not thread-safe): CTA 0 of bruno_conv3x3 records clock64() of its pipeline events into
 * device_buffer [64 tiles][16] (NULL = off) */
int bruno_debug_conv_trace(long long* device_buffer);
/* DEBUG (process-wide): 0 = automatic (line-tile kernel for >= 3 output lines per SM, else per-tap), 1 = per-tap kernel
 * (also the fallback for a row pitch > 128), 2 = line-tile kernel. */
int bruno_debug_conv_variant(int variant);
#endif

/* conv2d_wn weight normalisation (nvp:389) of L stacked 3x3 64->64 kernels V [L,3,3,64,64] (HWIO), g [L,64] into the
 * packed bf16 tensor-core operands: wfwd [L] (forward) and wbwd [L] (dgrad: transposed + flipped; may be NULL). */
int bruno_conv_wn_pack(const float* V, const float* g, void* wfwd, void* wbwd, int L, void* stream);
/* backward of the normalisation: dW [L,K,64] -> dV [L,K,64], dg [L,64]  (K = kh*kw*cin rows per output channel) */
int bruno_conv_wn_bwd(const float* V, const float* g, const float* dW, float* dV, float* dg, int L, int K, void* stream);

/* One 3x3, 64->64, stride-1 SAME convolution (tf.nn.conv2d at nvp:392) as a tcgen05 implicit GEMM, fused epilogue:
 *   mode 0: out = ELU(conv + bias)                      c2 of a residual block (nvp:120)
 *   mode 1: out = ELU(conv + bias + aux1)               c3 + skip (nvp:121-124)
 *   mode 2: out = conv * ELU'(aux1)                     dgrad through c3, aux1 = saved output of c2
 *   mode 3: out = (conv + aux2) * ELU'(aux1)            dgrad through c2 + skip gradient, aux1 = saved block input
 *   mode 4: out = conv + bias                           (data-dependent init, nvp:394-398)
 * ELU' is evaluated from the saved OUTPUT o of the ELU: 1 if o > 0 else o + 1.
 * bias_img_stride = 0: bias [64] shared by all images; >= 64: per-image bias table bias[n * stride + c] -- the
 * label-conditioned layers add dense_wn(y_label)[n, :] to the conv output (nn_extra_nvp_conditional.py:422-433). */
int bruno_conv3x3(int mode, const void* in_slab, const void* wpack, const float* bias, int bias_img_stride,
                  const void* aux1, const void* aux2, void* out_slab, int N, int H, int W, void* stream);
/* dW [3,3,64,64] (HWIO) += sum_pos x[pos + tap] (x) g[pos];  db [64] += sum_pos g[pos]   (fp32, atomically added) */
int bruno_conv3x3_wgrad(const void* x_slab, const void* g_slab, float* dW, float* db, int N, int H, int W, void* stream);
/* the same for L layers in ONE launch: layer l uses slab l of x_slab / g_slab (consecutive slabs of bruno_slab_bytes),
 * dW + l*3*3*64*64, db + l*64; the grid is divided among the layers. */
int bruno_conv3x3_wgrad_batched(const void* x_slab, const void* g_slab, float* dW, float* db, int L, int N, int H, int W,
                                void* stream);

/* c1 (nvp:113-115): out_slab = ELU(conv1x1_wn(x * mask)), x [N,H,W,C] fp32, V1 [C,64], g1, b1 [64]. */
/* f32_slab = 1: the slab is the fp32 verification layout ([row][64] floats, same padded row indexing, no swizzle). */
int bruno_c1_fwd(const float* x, const float* V1, const float* g1, const float* b1, int bias_img_stride, void* out_slab,
                 int f32_slab, int N, int H, int W, int C, int mask_type, int ks, void* stream);
/* g_slab = dL/d(pre-activation of c1): dx += mask * (g W1^T); dW1 [ks*ks*C,64] += (x mask)^T g; db1 [64] += sum g.
 * ks = 1: the 1x1 c1 of nn_extra_nvp.py:113; ks = 3: the c1 of a label-conditioned coupling, which the reference
 * evaluates as a 3x3 convolution (nn_extra_nvp_conditional.py:427 does not forward filter_size); V1 [ks*ks*C, 64] rows
 * ordered (kh, kw, cin) = the HWIO variable flattened. */
int bruno_c1_bwd(const float* x, const float* V1, const float* g1, const void* g_slab, float* dx, float* dW1, float* db1,
                 int N, int H, int W, int C, int mask_type, int ks, void* stream);

/* conv_out_scale / conv_out_translation heads (nvp:129-139) + affine coupling + log-det (forward :162-174, inverse
 * :176-187).  Ks, Kt [64,C]; bs, bt [C]; x, y [N,H,W,C]; ld [N]. */
/* label_s / label_t (both NULL or both [N,C]): per-image additive terms of the label-conditioned heads
 * (nn_extra_nvp_conditional.py:382-388). */
int bruno_heads_coupling(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                         const float* bt, const float* label_s, const float* label_t, const float* ld_in, float* y,
                         float* ld_out, int f32_slab, int N, int H, int W, int C, int mask_type, int inverse, void* stream);
/* fp32 verification path (CouplingLayerConv precision 'fp32'): the nine taps of a 3x3 convolution are accumulated with
 * bruno_sgemm on fp32 slabs ([rows][64] floats; row = position as in slab.cuh); this applies bias (+ skip) (+ ELU) on
 * pixel rows and zeroes the padding rows. */
int bruno_conv_epilogue_f32(float* acc, const float* bias, const float* skip, int elu, int N, int H, int W, void* stream);
/* backward of the forward coupling: dx = direct part (the s/t-network part is added by bruno_c1_bwd), g_slab =
 * dL/d(pre-activation of the last residual block); dKs, dbs, dKt, dbt are overwritten (deterministic two-level sum:
 * per-block partials in `scratch` of bruno_heads_bwd_scratch_floats(C) floats, then a fixed-order reduction). */
size_t bruno_heads_bwd_scratch_floats(int C);
int bruno_heads_coupling_bwd(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                             const float* bt, const float* dy, const float* dld, float* dx, void* g_slab, float* dKs,
                             float* dbs, float* dKt, float* dbt, float* scratch, int N, int H, int W, int C, int mask_type,
                             void* stream);
/* label-conditioned heads (nn_extra_nvp_conditional.py:372-392): label_s [N,C] is the per-image term of the scale head's
 * pre-activation; dlabel_s, dlabel_t [N,C] receive the per-image sums of d(s_pre), d(t_pre).  NULLs = the call above. */
int bruno_heads_coupling_bwd_cond(const float* x, const void* h_slab, const float* Ks, const float* bs, const float* Kt,
                                  const float* bt, const float* dy, const float* dld, float* dx, void* g_slab, float* dKs,
                                  float* dbs, float* dKt, float* dbt, float* scratch, int N, int H, int W, int C,
                                  int mask_type, const float* label_s, float* dlabel_s, float* dlabel_t, void* stream);
/* out[n, l*64 + c] = sum over the positions of image n of slab l (L stacked slabs): gradient of per-image bias tables */
int bruno_slab_image_colsum(const void* slabs, int L, float* out, int out_stride, int N, int H, int W, void* stream);

/* The whole CouplingLayerConv in one call.  forward_and_jacobian (nvp:162-174) / backward (:176-187, inverse = 1).
 * Stacked parameters: V1 [C,64], g1, b1 [64] (c1); packed residual weights wfwd [2R] (order c2_0, c3_0, c2_1, ...,
 * from bruno_conv_wn_pack), br [2R,64]; Ks, Kt [64,C], bs, bt [C].  `slabs`: n_slabs zero-initialised slabs of
 * bruno_slab_bytes each; n_slabs >= 2R+1 keeps every activation for the backward (slab 0 = c1 output, 2r+1 = c2_r
 * output, 2r+2 = block-r output), n_slabs = 3 ping-pongs (inference).
 * Label-conditioned variant (nn_extra_nvp_conditional.py:422-433): bias_img_stride != 0, b1 / br point into ONE per-image
 * table (image n of conv layer l at br + l*64 + n*bias_img_stride); label_s / label_t [N,C] as in bruno_heads_coupling. */
size_t bruno_coupling_conv_slabs(int R, int save);
int bruno_coupling_conv_fwd(const float* x, const float* ld_in, float* y, float* ld_out, int N, int H, int W, int C, int R,
                            int mask_type, int inverse, const float* V1, const float* g1, const float* b1,
                            const void* wfwd, const float* br, int bias_img_stride, const float* Ks, const float* bs,
                            const float* Kt, const float* bt, const float* label_s, const float* label_t, void* slabs,
                            int n_slabs, int c1_ks, void* stream);
/* Hand-written backward of the forward coupling: dx [N,H,W,C] and the gradients of every variable of the layer
 * (weight-norm backward included: dVr, dgr are w.r.t. V and g).  saved_slabs = the 2R+1 slabs of the forward;
 * gslabs = 2R+1 zero-initialised scratch slabs (slab 0: gradient w.r.t. the c1 pre-activation, slab l+1: gradient
 * w.r.t. the pre-activation of conv layer l -- all kept so that the 2R weight gradients are ONE batched launch);
 * wbwd = dgrad operands from bruno_conv_wn_pack; scratch = bruno_heads_bwd_scratch_floats(C) floats. */
int bruno_coupling_conv_bwd(const float* x, const float* dy, const float* dld, int N, int H, int W, int C, int R,
                            int mask_type, const float* V1, const float* g1, const float* Vr, const float* gr,
                            const void* wbwd, const float* Ks, const float* bs, const float* Kt, const float* bt,
                            const void* saved_slabs, void* gslabs, float* scratch, float* dx, float* dV1, float* dg1, float* db1,
                            float* dVr, float* dgr, float* dbr, float* dKs, float* dbs, float* dKt, float* dbt, void* stream);
/* the same for the label-conditioned coupling (nn_extra_nvp_conditional.py:34-159): label_s as in the forward call;
 * dtab [N, dtab_stride >= (2R+1)*64] receives the gradient of the per-image bias tables (column block l = conv layer l,
 * block 0 = c1), dlabel_s / dlabel_t [N,C] the gradients of the per-image head terms. */
int bruno_coupling_conv_bwd_cond(const float* x, const float* dy, const float* dld, int N, int H, int W, int C, int R,
                                 int mask_type, const float* V1, const float* g1, const float* Vr, const float* gr,
                                 const void* wbwd, const float* Ks, const float* bs, const float* Kt, const float* bt,
                                 const void* saved_slabs, void* gslabs, float* scratch, float* dx, float* dV1, float* dg1,
                                 float* db1, float* dVr, float* dgr, float* dbr, float* dKs, float* dbs, float* dKt, float* dbt,
                                 const float* label_s, float* dtab, int dtab_stride, float* dlabel_s, float* dlabel_t,
                                 int c1_ks, void* stream);

/* --- dense coupling layer, CouplingLayerDense (nvp:190-274) with dense_wn (:408-432), fp32 ---------------------- */
size_t bruno_coupling_dense_ws_floats(int N, int D, int U);
size_t bruno_coupling_dense_bwd_ws_floats(int N, int D, int U);
/* x, y [N,D]; V1 [D,U], V2 [U,U], Ks, Kt [U,D]; `ws` keeps what the backward needs.  inverse = 1: CouplingLayer.backward. */
int bruno_coupling_dense_fwd(const float* x, const float* ld_in, float* y, float* ld_out, int N, int D, int U, int mask_type,
                             int inverse, const float* V1, const float* g1, const float* b1, const float* V2,
                             const float* g2, const float* b2, const float* Ks, const float* bs, const float* Kt,
                             const float* bt, float* ws, void* stream);
int bruno_coupling_dense_bwd(const float* x, const float* dy, const float* dld, int N, int D, int U, int mask_type,
                             const float* V1, const float* g1, const float* b1, const float* V2, const float* g2,
                             const float* b2, const float* Ks, const float* Kt, const float* ws_fwd, float* ws_bwd,
                             float* dx, float* dV1, float* dg1, float* db1, float* dV2, float* dg2, float* db2, float* dKs,
                             float* dbs, float* dKt, float* dbt, void* stream);
/* sc[j] = g[j] / ||V[:, j]||, inv_nrm[j] = 1 / ||V[:, j]|| (may be NULL): the scaler of dense_wn (nvp:419) */
int bruno_dense_wn_scale(const float* V, const float* g, float* sc, float* inv_nrm, int K, int N, void* stream);
/* the affine coupling + log-det of CouplingLayerDense given the two head outputs s_pre, t_pre [N,D] (nvp:168-186) */
int bruno_dense_coupling(const float* x, const float* s_pre, const float* t_pre, const float* ld_in, float* y, float* ld_out,
                         int N, int D, int mask_type, int inverse, void* stream);
/* backward of the forward coupling above: dx = dy (b + (1-b) e^s) (direct part only; the s/t network adds its own through
 * ds_pre, dt_pre), ds_pre = (dy (1-b) x e^s + dld[n]) (1 - tanh^2), dt_pre = dy (1-b). */
int bruno_dense_coupling_bwd(const float* x, const float* s_pre, const float* dy, const float* dld, float* dx, float* ds_pre,
                             float* dt_pre, int N, int D, int mask_type, void* stream);
#ifndef BRUNO_NO_DEBUG
/* DEBUG (process-wide, not thread-safe) / A-B timing: which kernel the dense-path GEMMs (bruno_sgemm and the dense couplings)
 * use when the operands are 16-byte aligned: 0 = CUDA-core SGEMM (strict fp32 FFMA), 1 = tcgen05 3xTF32 split
 * (fp32-accurate, default; the split-K CTAs of a tile are one thread-block cluster and reduce through peer shared memory),
 * 2 = the same kernel with the round-1 split-K fix-up through the bound workspace (bit-identical results: both add the
 * partials in K order).  variant < 0 queries. */
int bruno_debug_gemm_variant(int variant);
/* DEBUG / host logic (no GPU needed): launch geometry of the tcgen05 GEMM for a shape on a device with `sms` SMs.  pair_mode 0 =
 * one product, 1 = two products side by side, 2 = two products K-concatenated (the dense coupling's s|t heads, dKs|dKt, dh).
 * out6 = {tile columns, gridDim.x, gridDim.y, gridDim.z (0: the pair needs two launches), K tiles of 32 per split-K slice,
 * thread-block cluster size}. */
int bruno_debug_gemm_plan(int M, int N, int K, int pair_mode, int sms, int* out6);
/* DEBUG: CTA (0,0,0) of the tcgen05 GEMM records clock64() of its pipeline events into device_buffer [64][8] (NULL = off) */
int bruno_debug_gemm_trace(long long* device_buffer);
#endif

/* C[M,N] (+)= act(op(A) op(B) * col_scale + bias), row-major fp32 (tf.matmul at nvp:418, tf.layers.dense :262,:268);
 * act: 0 none, 1 ELU, 2 leaky_relu(0.2), 3 tanh */
int bruno_sgemm(int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                int ldc, const float* col_scale, const float* bias, int act, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser steps with TF-1.7 semantics (config_rnn/train.py:121-131) over a flat fp32 buffer; grad_scale folds the
 * 1/nr_gpu tower average (train.py:103-105) and the student-grad schedule (train.py:108-111).
 * ---------------------------------------------------------------------------------------------------------- */
/* Label-conditioned dense coupling (nn_extra_nvp_conditional.py:162-266): every weight-normalised dense layer is
 * dense_wn(concat(leaky_relu(dense_wn(y_label)), input)) (:461-471), every head dense(concat(leaky_relu(dense(y_label)), h))
 * (:350-369).  yl [N, L] = the label embedding; `vars` = 20 device pointers in the reference's construction order:
 *   d1/label_h/{V [L,D/2], g, b}, d1/d1/{V [D/2+D,U], g, b}, d2/label_h/{V [L,U/2], g, b}, d2/d2/{V [U/2+U,U], g, b},
 *   d_scale/label_W1/{kernel [L,U/2], bias}, d_scale/d_scale/{kernel [U/2+U,D], bias}, d_translate/... (same shapes).
 * fwd: y, ld_out as bruno_coupling_dense_fwd; `ws` (bruno_coupling_dense_cond_ws_floats floats) keeps the activations.
 * bwd: dx [N,D], dyl [N,L] (gradient w.r.t. the label embedding), grads = 20 device pointers shaped like vars (overwritten;
 * weight-norm backward included); ws_bwd: bruno_coupling_dense_cond_bwd_ws_floats floats. */
size_t bruno_coupling_dense_cond_ws_floats(int N, int D, int U, int L);
size_t bruno_coupling_dense_cond_bwd_ws_floats(int N, int D, int U, int L);
int bruno_coupling_dense_cond_fwd(const float* x, const float* yl, const float* ld_in, float* y, float* ld_out, int N, int D,
                                  int U, int L, int mask_type, int inverse, const float* const* vars, float* ws, void* stream);
int bruno_coupling_dense_cond_bwd(const float* x, const float* yl, const float* dy, const float* dld, int N, int D, int U, int L,
                                  int mask_type, const float* const* vars, const float* ws_fwd, float* ws_bwd, float* dx,
                                  float* dyl, float* const* grads, void* stream);

/* One dense layer with its hand-written backward (the label paths of the conditional model: labels_layer m1:66-67,
 * label_h of every weight-normalised conv cond:422-433, label_h of the conv heads cond:382-388).
 *   g != NULL: dense_wn -- out = act((x W) g / ||W||_col + b + extra_bias)   (cond:437-459);   g == NULL: plain dense.
 *   act: 0 none, 1 ELU, 2 leaky_relu(0.2), 3 tanh (forward only).  fwd outputs: out [N,U], pre [N,U] (may be NULL when
 *   act == 0 or 2 and g == NULL), sc / inv_nrm [U] (weight norm), btot [U] = b + extra_bias.
 *   bwd: dW [K,U] (w.r.t. V for weight norm), dg [U], db [U] (= gradient of b AND of extra_bias), dx [N,K] (may be NULL);
 *   ws: bruno_dense_layer_bwd_ws_floats(N, U) floats. */
size_t bruno_dense_layer_bwd_ws_floats(int N, int U);
int bruno_dense_layer_fwd(const float* x, const float* W, const float* g, const float* b, const float* extra_bias, int N, int K,
                          int U, int act, float* out, float* pre, float* sc, float* inv_nrm, float* btot, void* stream);
int bruno_dense_layer_bwd(const float* x, const float* W, const float* g, const float* btot, const float* sc,
                          const float* inv_nrm, const float* pre, const float* out, const float* dout, int N, int K, int U,
                          int act, float* ws, float* dx, float* dW, float* dg, float* db, void* stream);

/* Episodic fine-tuning loss (config_rnn/bn2_omniglot_tp_ft_1s_20w.py:222-226): log_probs [M*B, T] (M episodes of B candidate
 * sequences, candidate 0 = the true class); loss[0] = -mean_m log softmax_b(log_probs[m, b, T-1])[0]; dlog_probs (optional,
 * [M*B, T]) receives d loss / d log_probs. */
int bruno_episode_softmax_loss(const float* log_probs, int M, int B, int T, float* loss, float* dlog_probs, void* stream);

int bruno_rmsprop_tf17(float* p, const float* g, float* ms, size_t n, float lr, float decay, float eps, float grad_scale,
                       void* stream);
int bruno_adam_tf17(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2, float eps,
                    int t, float grad_scale, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BRUNO_SM100_H_ */
