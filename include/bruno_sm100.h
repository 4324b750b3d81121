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
 *   - all data pointers are DEVICE pointers owned by the caller; nothing here allocates device memory
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *   - float tensors are contiguous fp32; images are NHWC like the reference
 *   - entry points are re-entrant across streams and devices (no mutable global state)
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

#ifdef __cplusplus
}
#endif
#endif /* BRUNO_SM100_H_ */
