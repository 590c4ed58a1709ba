/*
 * gpmpc.h -- C ABI of libgpmpc.so, the B200 (sm_100a) implementation of the GP-MPC rollout hot path.
 *
 * The reference (Thiagodcv/gaussian-process-mpc) has no FFI; its hot path sits behind Python classes.
 * Each entry point below names the reference interface it replaces (file:line relative to the
 * reference root).  The Python classes in gaussian-process-mpc_b200/ bind these with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 (GPMPC_OK) or a negative error code and never throws;
 *     gpmpc_last_error(h) gives the message of the last failure on that handle;
 *   - all arrays are fp64, row-major, dense; every data pointer may be a HOST or a DEVICE pointer
 *     (detected per pointer with cudaPointerGetAttributes); host buffers are staged through the
 *     handle's stream;
 *   - a handle owns one GP bundle (E outputs sharing the training inputs X) on one CUDA device; calls on
 *     a handle are serialised on its stream; results written to device buffers are ordered on that
 *     stream, results written to host buffers are complete on return;
 *   - non-finite results (e.g. the NaN cost of log(det) < 0, src/mpc.py:183) are data, not errors.
 *
 * Symbols: n training points, D = E + m input dimension, E outputs (state_dim), m actions,
 *          B independent rollouts, H horizon.
 */
#ifndef GPMPC_H
#define GPMPC_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpmpc_ctx *gpmpc_handle;

#define GPMPC_OK               0
#define GPMPC_ERR_INVALID     -1   /* bad argument / shape                                   */
#define GPMPC_ERR_CUDA        -2   /* CUDA runtime failure                                   */
#define GPMPC_ERR_NOT_FIT     -3   /* no training data / gpmpc_fit not called                */
#define GPMPC_ERR_UNSUPPORTED -4   /* D, E outside the compiled range                        */
#define GPMPC_ERR_NOT_PD      -5   /* Cholesky met a non-positive pivot (Ky not positive definite) */

#define GPMPC_MAX_D 8              /* compiled input dimensions: 1..8                         */
#define GPMPC_MAX_E 8              /* compiled output counts:    1..8 (E < D unless m == 0)   */

/* gpmpc_get_matrix selectors */
#define GPMPC_MAT_KF     0         /* [n,n]  sigma_f^2 exp(-1/2 d^2)          src/gpr.py:169  */
#define GPMPC_MAT_KY     1         /* [n,n]  Kf + noise_var I                 src/gpr.py:170  */
#define GPMPC_MAT_KY_INV 2         /* [n,n]  Ky^-1                            src/gpr.py:171  */
#define GPMPC_MAT_BETA   3         /* [n]    Ky^-1 y          src/tools/uncertainty_prop.py:327 */

int gpmpc_version(void);

/* Handle lifetime.  Replaces Dynamics.__init__ / GaussianProcessRegression.__init__
 * (src/dynamics.py:10-37, src/gpr.py:11-49): one handle per bundle of E GPs with D inputs.        */
int gpmpc_create(int device, int D, int E, gpmpc_handle *out);
int gpmpc_destroy(gpmpc_handle h);
const char *gpmpc_last_error(gpmpc_handle h);   /* h may be NULL: message of the last failed create */
int gpmpc_set_stream(gpmpc_handle h, void *cuda_stream);   /* cudaStream_t; NULL = legacy default  */
int gpmpc_synchronize(gpmpc_handle h);
/* Tuning switches (all default to the fastest path).
 *   "persistent_single" (0): 1 = a single rollout (B = 1, one IPOPT callback, src/mpc.py:202-255) runs its whole horizon
 *                            in one persistent cooperative launch; 0 = one fused launch per horizon step (measured faster
 *                            on one GPU).  A rollout split over several GPUs (gpmpc_split_*) always uses the persistent kernel.
 *   "split_timeline" (0):    stamp the inter-GPU exchange of every step (see gpmpc_split_last_exchange_us).
 *   "l2_persist" (0):        1 = few-rollouts kernels launch with an L2 access-policy window over the weights so that the
 *                            part of Wt that fits the persisting L2 carve-out stays resident from one horizon step to the
 *                            next (measured: no effect on B200, the kernel is not bound by the stream alone).
 *   "single_big_share" (660): per mille (500..900) of the tiles of a single rollout that go to the CTA that arrived first on
 *                            each SM (the warp scheduler favours it); 500 = equal static slices.  The result does not
 *                            depend on which CTA claimed which slice.                                                    */
int gpmpc_set_option(gpmpc_handle h, const char *name, int value);
int gpmpc_num_train(gpmpc_handle h);

/* Fit: Gram matrix, blocked Cholesky, Ky^-1, beta and the moment-matching weight matrices.
 * Replaces GaussianProcessRegression.build_Ky_inv_mat for all E outputs (src/gpr.py:159-171) as driven by
 * Dynamics.append_train_data (src/dynamics.py:39-60).
 *   X[n,D], Y[n,E], lambdas[E,D] (SQUARED length-scales), sigma_f[E], noise_var[E] (= sigma_n^2 as the
 *   caller wants it added to the diagonal; the reference adds an fp32-rounded value, src/gpr.py:170).   */
int gpmpc_fit(gpmpc_handle h, int n, const double *X, const double *Y, const double *lambdas,
              const double *sigma_f, const double *noise_var);

/* Refit one output only (hyper-parameters of output `a` changed; X unchanged).  Replaces a single
 * GaussianProcessRegression.build_Ky_inv_mat call (src/gpr.py:159-171).  y may be NULL (keep targets). */
int gpmpc_refit_output(gpmpc_handle h, int a, const double *y, const double *lambdas_a, double sigma_f_a,
                       double noise_var_a);

/* Append ONE observation (x[D], y[E]) to a fitted bundle with a bordered (rank-1) update of every Ky^-1, beta,
 * log det Ky and weight matrix: O(E n^2) instead of the O(E n^3) rebuild the reference performs on every
 * closed-loop step (src/simulator.py:55 -> src/gpr.py:122,171; its own attempt at this, src/gpr.py:137-157, is
 * marked "don't use").  Returns GPMPC_OK, or GPMPC_REFIT_NEEDED (> 0, not an error) when the padded layout
 * is full or a Schur complement s = kappa - k^T Ky^-1 k is not safely positive (s <= 1e3 eps n kappa^2 / noise, the
 * rounding error an explicit inverse of condition n sf^2 / noise leaves in it): the caller then calls gpmpc_fit
 * with all the data, which also bounds the drift of repeated updates to at most 63 of them.              */
#define GPMPC_REFIT_NEEDED 1
int gpmpc_append_point(gpmpc_handle h, const double *x, const double *y);

/* Change the length-scales / amplitudes used by the MOMENT-MATCHING formulas without refitting Ky^-1.
 * The reference reads log_lambdas / sigma_f at rollout time (src/dynamics.py:171,173) but Ky_inv from the
 * last build (src/dynamics.py:170), so setters without a rebuild affect only the propagation.           */
int gpmpc_set_propagation_hypers(gpmpc_handle h, const double *lambdas, const double *sigma_f);

/* Copy a fitted matrix of output a into out (leading dimension n for matrices).                      */
int gpmpc_get_matrix(gpmpc_handle h, int which, int a, double *out);

/* K(X*, X_train) for output a: out[p,n].  Replaces compute_pred_train_covariance (src/gpr.py:253-283). */
int gpmpc_kernel_matrix(gpmpc_handle h, int a, int p, const double *Xs, double *out);

/* Posterior at p test inputs for output a: mean[p]; cov[p,p] if cov != NULL (+ noise_var I if
 * add_noise).  Replaces predict_latent_vars with f_nom = None (src/gpr.py:285-332).                  */
int gpmpc_predict(gpmpc_handle h, int a, int p, const double *Xs, double *mean, double *cov, int add_noise);
/* Extended forms.  The reference evaluates K(X*, X) and K(X*, X*) with the hyper-parameters held by the object at
 * CALL time (src/gpr.py:268-276,317-324) while Ky^-1 stays from the last build, and adds the fp64 sigma_n^2 to the
 * target covariance (src/gpr.py:329): hyp = [lambda_1..D, sigma_f, noise_var] overrides the fit-time values for these
 * kernel evaluations (NULL = fit-time values).  resid[n] = y - f_nom(X) replaces the targets in the mean,
 * mean[p] = K(X*,X) Ky^-1 resid; the caller adds f_nom(X*) (src/gpr.py:309).  resid == NULL: the targets.        */
int gpmpc_kernel_matrix_ex(gpmpc_handle h, int a, int p, const double *Xs, const double *hyp, double *out);
int gpmpc_predict_ex(gpmpc_handle h, int a, int p, const double *Xs, const double *resid, const double *hyp,
                     double *mean, double *cov, int add_noise);

/* Log marginal likelihood of output a for the hyper-parameters of the last fit,
 *   ml = -1/2 r^T Ky^-1 r - 1/2 log det Ky - n/2 log 2 pi,   r = resid (or the training targets if resid == NULL),
 * with log det Ky = 2 sum log diag(L) from the Cholesky factor, and (if grad != NULL) its gradient w.r.t.
 * [log lambda_1..D, log sigma_f, log sigma_n] = 1/2 tr((alpha alpha^T - Ky^-1) dKy/dtheta), alpha = Ky^-1 r.
 * Replaces compute_marginal_likelihood + the autograd backward used by update_hyperparams
 * (src/gpr.py:240-251,334-370).                                                                       */
int gpmpc_marginal_likelihood(gpmpc_handle h, int a, const double *resid, double *ml, double *grad);

/* Exact moments of all E GP outputs for B Gaussian inputs N(U_b, S_b).  S is [B,D] (diagonal variances,
 * s_is_full = 0) or [B,D,D] (full covariance, s_is_full = 1).  mean[B,E], var[B,E] (latent variance, no
 * noise term).  Replaces mean_prop_torch + variance_prop_torch called on a fitted bundle
 * (src/tools/uncertainty_prop.py:296-338,341-399 as used at src/dynamics.py:175-181).                 */
int gpmpc_moment_match(gpmpc_handle h, int B, const double *U, const double *S, int s_is_full,
                       double *mean, double *var);

/* Same for a FULL input covariance S[B,D,D], returning the full E x E output covariance cov[B,E,E]: variances on
 * the diagonal and cross-covariances beta_a^T Qt beta_b - m_a m_b off it (src/tools/uncertainty_prop.py:187-236,
 * the formula-correct NumPy form).  This is the moment-matching step of a full-covariance rollout, which the
 * reference leaves as a TODO (src/dynamics.py:184); all B inputs are evaluated by one batched launch sequence.  */
int gpmpc_moment_match_cov(gpmpc_handle h, int B, const double *U, const double *S, double *mean, double *cov);

/* FULL-covariance moment-matched rollout of B control sequences: Sigma_t keeps the cross-covariances between the
 * outputs; the input covariance of step t is blockdiag(Sigma_{t-1}, fp32(1e-3) I) (the wiring the reference left
 * commented out, src/dynamics.py:104-121,184, with the cross term of src/tools/uncertainty_prop.py:187-236).
 *   x0[B,E], U[B,H,m]  ->  means[B,H+1,E], covs[B,H+1,E,E]
 * gpmpc_rollout_full_vjp: vector-Jacobian product of the last gpmpc_rollout_full (reverse mode with recomputation):
 *   gmeans[B,H+1,E], gcovs[B,H+1,E,E] (either may be NULL)  ->  gU[B,H,m], gx0[B,E] (may be NULL).               */
int gpmpc_rollout_full(gpmpc_handle h, int B, int H, const double *x0, const double *U, double *means, double *covs);
int gpmpc_rollout_full_vjp(gpmpc_handle h, int B, int H, const double *gmeans, const double *gcovs, double *gU,
                           double *gx0);

/* Fused objective + gradient under the full-covariance rollout: same arguments as gpmpc_rollout_cost_grad, the cost
 * is src/mpc.py:156-200 evaluated on the full Sigma_t (which that function already accepts, :182-185);
 * covs[B,H+1,E,E] (may be NULL).                                                                               */
int gpmpc_rollout_cost_grad_full(gpmpc_handle h, int B, int H, const double *x0, const double *U,
                                 const double *gamma, const double *Q, const double *R, const double *Rdelta,
                                 const double *last_u, const double *xref, const double *uref, double *cost,
                                 double *grad, double *means, double *covs);

/* Stateless form of the same two functions for caller-supplied matrices (no handle state is used except
 * the device/stream): Kinv[n,n], lambdas[D], u[D], S[D,D] full, X[n,D], y[n].
 * Outputs: mean[1], var[1] (may be NULL), beta[n] (may be NULL), l[n] (may be NULL).
 * If y == NULL the call has variance_prop_torch's signature: beta[n] and mean[1] are INPUTS and only var
 * is written.  Replaces mean_prop_torch / variance_prop_torch as free functions
 * (src/tools/uncertainty_prop.py:296-338,341-399).                                                  */
int gpmpc_moment_match_raw(gpmpc_handle h, int n, int D, const double *Kinv, const double *lambdas,
                           const double *u, const double *S, const double *X, const double *y,
                           double sigma_f, double *mean, double *var, double *beta, double *l);

/* Cross-covariance of outputs a and b for one Gaussian input, caller-supplied vectors:
 * cov = beta1^T Qt beta2 - mean1 mean2 (src/tools/uncertainty_prop.py:187-236 / 402-465).
 * bugcompat != 0 reproduces the transposed cross term of the torch function (:446).                  */
int gpmpc_covariance_raw(gpmpc_handle h, int n, int D, const double *lambdas1, const double *lambdas2,
                         const double *u, const double *S, const double *X, double mean1, double mean2,
                         const double *beta1, const double *beta2, double sigma_f1, double sigma_f2,
                         int bugcompat, double *cov);

/* Variance-only moment-matched rollout of B control sequences over H steps.
 *   x0[B,E], U[B,H,m]  ->  means[B,H+1,E], vars[B,H+1,E]
 * Step-0 variance is 1e-3 (exact fp64) and the action variance is fp32(1e-3) as in
 * src/dynamics.py:148,162.  The per-step partial derivatives are kept in the handle for gpmpc_rollout_vjp.
 * Replaces Dynamics.forward_propagate_torch (src/dynamics.py:126-191).                                */
int gpmpc_rollout(gpmpc_handle h, int B, int H, const double *x0, const double *U, double *means,
                  double *vars);

/* Vector-Jacobian product of the last gpmpc_rollout: given d loss / d means and d loss / d vars
 * ([B,H+1,E] each, either may be NULL) return d loss / d U [B,H,m] and, if gx0 != NULL, d loss / d x0 [B,E].
 * Replaces the autograd replay through forward_propagate_torch (src/mpc.py:251).                      */
int gpmpc_rollout_vjp(gpmpc_handle h, int B, int H, const double *gmeans, const double *gvars, double *gU,
                      double *gx0);

/* Fused objective + gradient for B independent control sequences: rollout, risk-sensitive cost
 * (src/mpc.py:156-200) and its exact gradient w.r.t. U (src/mpc.py:231-255).
 *   x0[B,E], U[B,H,m], gamma[B], Q[E,E], R[m,m], Rdelta[m,m] or NULL, last_u[B,m] (used iff Rdelta),
 *   xref[E], uref[m]  ->  cost[B], grad[B,H,m] (NULL = cost only), means/vars [B,H+1,E] (may be NULL).
 * Replaces RiskSensitiveMPC.objective + gradient (src/mpc.py:202-255).                                */
int gpmpc_rollout_cost_grad(gpmpc_handle h, int B, int H, const double *x0, const double *U,
                            const double *gamma, const double *Q, const double *R, const double *Rdelta,
                            const double *last_u, const double *xref, const double *uref, double *cost,
                            double *grad, double *means, double *vars);

/* ONE rollout split over several GPUs of a node (the only split that shortens a single IPOPT callback,
 * src/mpc.py:202-255).  Every rank fits the same GP (replica), connects once, and from then on a B = 1 call of
 * gpmpc_rollout_cost_grad / gpmpc_rollout made by ALL ranks with the same arguments (same calls, same order) sweeps 1/world
 * of every step's pair space per GPU; the per-step sums are exchanged by the kernels themselves through peer-mapped
 * mailboxes (NVLink P2P stores + release/acquire flags) and every rank returns the same result.
 *   gpmpc_split_export        creates this handle's mailbox and returns its CUDA IPC handle (GPMPC_IPC_HANDLE_BYTES bytes)
 *   gpmpc_split_connect       all_handles = the `world` exported handles in rank order (other processes' mailboxes)
 *   gpmpc_split_connect_local same for handles living in THIS process (peers[r] for r != rank, one device each)
 *   gpmpc_split_disconnect    back to single-GPU evaluation                                                       */
#define GPMPC_IPC_HANDLE_BYTES 64
int gpmpc_split_export(gpmpc_handle h, void *ipc_handle_out);
int gpmpc_split_connect(gpmpc_handle h, int rank, int world, const void *all_handles);
int gpmpc_split_connect_local(gpmpc_handle h, int rank, int world, gpmpc_handle *peers);
int gpmpc_split_disconnect(gpmpc_handle h);
/* With gpmpc_set_option(h, "split_timeline", 1): per-step time (us, %globaltimer) this rank spent between "local sums
 * ready" and "all ranks' sums read" in the last split evaluation: mean and maximum over the horizon steps.         */
int gpmpc_split_last_exchange_us(gpmpc_handle h, double *mean_us, double *max_us);

/* Introspection for benchmarks: kernels launched by this handle since creation, and the device time
 * (ms, CUDA events on the handle's stream) of the last pair-sum kernel sequence.                      */
long long gpmpc_launch_count(gpmpc_handle h);
int gpmpc_last_pair_kernel_ms(gpmpc_handle h, double *ms, long long *pair_evals);
/* The timers above synchronise the stream after every horizon step while armed (the first call of
 * gpmpc_last_pair_kernel_ms arms them); gpmpc_set_pair_timing(h, 0) disarms them again.              */
int gpmpc_set_pair_timing(gpmpc_handle h, int on);

/* Measured ceilings for the roofline: sustained fp64 FMA rate (TFLOP/s, 2 flop per FMA) and the rate of
 * the library's own exp() (Gexp/s) on this device.                                                    */
int gpmpc_measure_fp64_peak(gpmpc_handle h, double *fma_tflops, double *exp_gops);

#ifdef __cplusplus
}
#endif
#endif /* GPMPC_H */
