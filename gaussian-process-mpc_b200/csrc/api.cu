// C-ABI entry points of libgpmpc.so that are not part of the rollout (see include/gpmpc.h):
// handle lifetime, fit, matrix export, posterior prediction, full-covariance moment matching and the
// stateless ("raw") forms of the reference's free functions.
#include "common.cuh"
#include "mm_pairs.cuh"
#include <cmath>

using namespace gpmpc;

extern "C" int gpmpc_moment_match_diag_internal(gpmpc_handle h, int B, const double *U_dev, const double *S_dev,
                                                double *mean_out, double *var_out);

static std::string g_create_error;

__global__ void gpmpc_rowdot_kernel(const double *__restrict__ A, int ld, int n, const double *__restrict__ x,
                                    double *__restrict__ y, int rows);

extern "C" int gpmpc_version(void) { return 100; }

extern "C" const char *gpmpc_last_error(gpmpc_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int gpmpc_create(int device, int D, int E, gpmpc_handle *out)
{
    if (!out) return GPMPC_ERR_INVALID;
    *out = nullptr;
    if (D < 1 || D > kMaxD || E < 1 || E > kMaxE || E > D) {
        g_create_error = "gpmpc_create: need 1 <= E <= D <= 8";
        return GPMPC_ERR_UNSUPPORTED;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        g_create_error = std::string("gpmpc_create: no such CUDA device (") + cudaGetErrorString(e) + ")";
        cudaGetLastError();
        return GPMPC_ERR_CUDA;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return GPMPC_ERR_CUDA;
    }
    gpmpc_ctx *h = new gpmpc_ctx();
    h->device = device; h->D = D; h->E = E; h->m = D - E;
    for (int a = 0; a < kMaxE; ++a) {
        h->sf_fit[a] = h->sf_prop[a] = 1.0; h->noise[a] = 1.0;
        for (int k = 0; k < kMaxD; ++k) h->lam_fit[a][k] = h->lam_prop[a][k] = 1.0;
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    *out = h;
    return GPMPC_OK;
}

extern "C" int gpmpc_destroy(gpmpc_handle h)
{
    if (!h) return GPMPC_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    for (DevBuf *b : {&h->X, &h->Y, &h->Kinv, &h->Wt, &h->beta, &h->chol, &h->zt, &h->tt, &h->linv, &h->info, &h->hyp,
                      &h->mu, &h->var, &h->tape, &h->cst, &h->part, &h->mpart, &h->stage_in, &h->stage_out, &h->gbuf,
                      &h->tickets, &h->dbg, &h->Wx, &h->fc_plan, &h->fc_mu, &h->fc_cov, &h->fc_cst, &h->fc_raw, &h->fc_part,
                      &h->fc_red, &h->fc_gbar, &h->fc_seed, &h->fc_carry, &h->fc_io})
        b->release();
    for (gpmpc::LookAhead &l : h->la) {
        if (l.side) cudaStreamDestroy(l.side);
        for (cudaEvent_t e : l.ev) cudaEventDestroy(e);
    }
    for (cudaStream_t st : h->aux_streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : h->aux_events) cudaEventDestroy(ev);
    h->pin_in.release(); h->pin_out.release(); h->claim.release();
    gpmpc_split_disconnect(h);
    h->split_buf.release();
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return GPMPC_OK;
}

extern "C" int gpmpc_set_stream(gpmpc_handle h, void *cuda_stream)
{
    if (!h) return GPMPC_ERR_INVALID;
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return GPMPC_OK;
}

extern "C" int gpmpc_synchronize(gpmpc_handle h)
{
    if (!h) return GPMPC_ERR_INVALID;
    GP_CUDA(h, cudaSetDevice(h->device));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

extern "C" int gpmpc_set_option(gpmpc_handle h, const char *name, int value)
{
    if (!h || !name) return GPMPC_ERR_INVALID;
    if (std::strcmp(name, "persistent_single") == 0) { h->opt_persistent = value != 0; return GPMPC_OK; }
    if (std::strcmp(name, "single_big_share") == 0) {       // per mille, 500..900
        if (value < 500 || value > 900) return fail(h, GPMPC_ERR_INVALID, "single_big_share: 500..900 per mille");
        h->opt_single_big = value; return GPMPC_OK;
    }
    if (std::strcmp(name, "split_timeline") == 0) { h->opt_split_timeline = value != 0; return GPMPC_OK; }
    if (std::strcmp(name, "l2_persist") == 0) {
        h->opt_l2_persist = value != 0;
        if (!h->opt_l2_persist) { cudaSetDevice(h->device); cudaCtxResetPersistingL2Cache(); }
        return GPMPC_OK;
    }
    return fail(h, GPMPC_ERR_INVALID, std::string("gpmpc_set_option: unknown option ") + name);
}

extern "C" int gpmpc_split_last_exchange_us(gpmpc_handle h, double *mean_us, double *max_us)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (mean_us) *mean_us = h->split_exchange_mean_us;
    if (max_us) *max_us = h->split_exchange_max_us;
    return GPMPC_OK;
}

extern "C" int gpmpc_num_train(gpmpc_handle h) { return h ? h->n : GPMPC_ERR_INVALID; }

extern "C" long long gpmpc_launch_count(gpmpc_handle h) { return h ? h->launches : 0; }

extern "C" int gpmpc_last_pair_kernel_ms(gpmpc_handle h, double *ms, long long *pair_evals)
{
    if (!h) return GPMPC_ERR_INVALID;
    h->time_pairs = true;                 // first call arms the timers; later calls read them
    if (ms) *ms = h->last_pair_ms;
    if (pair_evals) *pair_evals = h->last_pair_evals;
    return GPMPC_OK;
}

extern "C" int gpmpc_set_pair_timing(gpmpc_handle h, int on)
{
    if (!h) return GPMPC_ERR_INVALID;
    h->time_pairs = on != 0;
    return GPMPC_OK;
}

// fetch a small host-or-device array into host memory
static int fetch_host(gpmpc_ctx *h, const double *src, double *dst, size_t cnt)
{
    if (is_device_ptr(src)) {
        GP_CUDA(h, cudaMemcpyAsync(dst, src, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GP_CUDA(h, cudaStreamSynchronize(h->stream));
    } else {
        std::memcpy(dst, src, cnt * sizeof(double));
    }
    return GPMPC_OK;
}

namespace gpmpc {
struct KsArg { double inv_lam[kMaxD]; double sf2; };     // fit-time kernel hyper-parameters of one output
static KsArg ks_arg(const gpmpc_ctx *h, int a)
{
    KsArg k;
    for (int i = 0; i < kMaxD; ++i) k.inv_lam[i] = i < h->D ? 1.0 / h->lam_fit[a][i] : 0.0;
    k.sf2 = h->sf_fit[a] * h->sf_fit[a];
    return k;
}
// kernel hyper-parameters supplied by the caller: hyp = [lambda_1..D, sigma_f, noise_var]
static KsArg ks_arg_from(int D, const double *hyp)
{
    KsArg k;
    for (int i = 0; i < kMaxD; ++i) k.inv_lam[i] = i < D ? 1.0 / hyp[i] : 0.0;
    k.sf2 = hyp[D] * hyp[D];
    return k;
}
__global__ void transpose_y_kernel(const double *__restrict__ Y, int n, int E, int ld, double *__restrict__ Yt)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (j < ld) Yt[(size_t)a * ld + j] = (j < n) ? Y[(size_t)j * E + a] : 0.0;
}
}  // namespace gpmpc

extern "C" int gpmpc_fit(gpmpc_handle h, int n, const double *X, const double *Y, const double *lambdas,
                         const double *sigma_f, const double *noise_var)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (n <= 0 || !X || !Y || !lambdas || !sigma_f || !noise_var) return fail(h, GPMPC_ERR_INVALID, "gpmpc_fit: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int D = h->D, E = h->E;
    double lam[kMaxE * kMaxD], sf[kMaxE], nv[kMaxE];
    int rc;
    if ((rc = fetch_host(h, lambdas, lam, (size_t)E * D))) return rc;
    if ((rc = fetch_host(h, sigma_f, sf, E))) return rc;
    if ((rc = fetch_host(h, noise_var, nv, E))) return rc;
    for (int a = 0; a < E; ++a) {
        for (int k = 0; k < D; ++k) {
            if (!(lam[a * D + k] > 0.0)) return fail(h, GPMPC_ERR_INVALID, "gpmpc_fit: lambdas must be positive");
            h->lam_fit[a][k] = h->lam_prop[a][k] = lam[a * D + k];
        }
        h->sf_fit[a] = h->sf_prop[a] = sf[a];
        h->noise[a] = nv[a];
    }
    h->n = n;
    h->ld = round_up(n, kTile);
    h->fitted = false;
    h->tape_B = h->tape_H = 0; h->fc_B = h->fc_H = 0;
    const int ld = h->ld;
    GP_CUDA(h, h->X.reserve((size_t)ld * D * sizeof(double)));
    GP_CUDA(h, h->Y.reserve((size_t)ld * E * sizeof(double)));
    GP_CUDA(h, cudaMemsetAsync(h->X.p, 0, (size_t)ld * D * sizeof(double), h->stream));
    GP_CUDA(h, to_device(h, h->X.p, X, (size_t)n * D * sizeof(double)));
    // Y arrives [n,E]; keep it as [E,ld]
    {
        const double *Yd = Y;
        if (!is_device_ptr(Y)) {
            GP_CUDA(h, h->stage_in.reserve((size_t)n * E * sizeof(double)));
            GP_CUDA(h, cudaMemcpyAsync(h->stage_in.p, Y, (size_t)n * E * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            Yd = h->stage_in.as<double>();
        }
        transpose_y_kernel<<<dim3((ld + 127) / 128, E), 128, 0, h->stream>>>(Yd, n, E, ld, h->Y.as<double>());
        GP_LAUNCH_CHECK(h);
    }
    if ((rc = upload_prop_hypers(h))) return rc;
    bool which[kMaxE];
    for (int a = 0; a < kMaxE; ++a) which[a] = a < E;
    return fit_all(h, which);
}

extern "C" int gpmpc_refit_output(gpmpc_handle h, int a, const double *y, const double *lambdas_a, double sigma_f_a,
                                  double noise_var_a)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (h->n <= 0) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_refit_output: no training data");
    if (a < 0 || a >= h->E || !lambdas_a) return fail(h, GPMPC_ERR_INVALID, "gpmpc_refit_output: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    double lam[kMaxD];
    int rc;
    if ((rc = fetch_host(h, lambdas_a, lam, h->D))) return rc;
    for (int k = 0; k < h->D; ++k) {
        if (!(lam[k] > 0.0)) return fail(h, GPMPC_ERR_INVALID, "gpmpc_refit_output: lambdas must be positive");
        h->lam_fit[a][k] = h->lam_prop[a][k] = lam[k];
    }
    h->sf_fit[a] = h->sf_prop[a] = sigma_f_a;
    h->noise[a] = noise_var_a;
    if (y) {
        GP_CUDA(h, cudaMemsetAsync(h->Y.as<double>() + (size_t)a * h->ld, 0, (size_t)h->ld * sizeof(double), h->stream));
        GP_CUDA(h, to_device(h, h->Y.as<double>() + (size_t)a * h->ld, y, (size_t)h->n * sizeof(double)));
    }
    if ((rc = upload_prop_hypers(h))) return rc;
    bool which[kMaxE];
    for (int i = 0; i < kMaxE; ++i) which[i] = i == a;
    h->tape_B = h->tape_H = 0; h->fc_B = h->fc_H = 0;
    // a failed factorisation leaves Kinv / beta / Wt of this output half-written: the handle must not keep
    // reporting "fitted" (fit_all sets the flag again on success)
    h->fitted = false;
    return fit_all(h, which);
}

// ---------------------------------------------------------------------------------------------
// Incremental refit: bordered update of Ky^-1 for one new training point
// ---------------------------------------------------------------------------------------------
namespace gpmpc {
// kv[i] = sf^2 exp(-1/2 sum_k (x_ik - xs_k)^2 / lam_k), i < n
__global__ void knew_kernel(const double *__restrict__ X, int n, int D, KsArg hp, const double *__restrict__ xs,
                            double *__restrict__ kv)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double q = 0.0;
    for (int k = 0; k < D; ++k) { const double d = X[(size_t)i * D + k] - xs[k]; q = fma(d * d, hp.inv_lam[k], q); }
    kv[i] = hp.sf2 * exp(-0.5 * q);
}
// scal[0] = kappa - sum_i kv_i v_i  (Schur complement), scal[1] = 1 / scal[0]      (single block, fixed order)
__global__ void __launch_bounds__(1024) schur_kernel(const double *__restrict__ kv, const double *__restrict__ v, int n,
                                                      double kappa, double *__restrict__ scal)
{
    __shared__ double sm[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s = fma(kv[i], v[i], s);
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) { if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) { const double sc = kappa - sm[0]; scal[0] = sc; scal[1] = 1.0 / sc; }
}
// Kinv[0:n,0:n] += v v^T / s;  Kinv[i][n] = Kinv[n][i] = -v_i / s;  Kinv[n][n] = 1 / s
__global__ void border_update_kernel(double *__restrict__ Kinv, int ld, int n, const double *__restrict__ v,
                                     const double *__restrict__ scal)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i > n || j > n) return;
    const double is = scal[1];
    double *p = Kinv + (size_t)i * ld + j;
    if (i < n && j < n) *p = fma(v[i] * is, v[j], *p);
    else if (i == n && j == n) *p = is;
    else *p = -(i == n ? v[j] : v[i]) * is;
}
}  // namespace gpmpc

extern "C" int gpmpc_append_point(gpmpc_handle h, const double *x, const double *y)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!x || !y) return fail(h, GPMPC_ERR_INVALID, "gpmpc_append_point: null");
    if (!h->fitted || h->n <= 0) return GPMPC_REFIT_NEEDED;
    if (h->n + 1 > h->ld) return GPMPC_REFIT_NEEDED;
    GP_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n, D = h->D, E = h->E, ld = h->ld;
    const size_t mat = (size_t)ld * ld;
    double xh[kMaxD], yh[kMaxE];
    int rc;
    if ((rc = fetch_host(h, x, xh, D))) return rc;
    if ((rc = fetch_host(h, y, yh, E))) return rc;
    // workspace: xs [D] | kv [ld] | v [ld] | scal [2 * E]
    GP_CUDA(h, h->gbuf.reserve(((size_t)2 * ld + kMaxD + 2 * kMaxE + 8) * sizeof(double)));
    double *w = h->gbuf.as<double>();
    double *xs = w; w += kMaxD;
    double *kv = w; w += ld;
    double *v = w; w += ld;
    double *scal = w;
    GP_CUDA(h, cudaMemcpyAsync(xs, xh, D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    // first pass: Schur complements of all outputs (nothing is modified until every one is known to be positive)
    std::vector<double> hs(2 * E);
    for (int a = 0; a < E; ++a) {
        knew_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->X.as<double>(), n, D, ks_arg(h, a), xs, kv);
        GP_LAUNCH_CHECK(h);
        gpmpc_rowdot_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->Kinv.as<double>() + a * mat, ld, n, kv, v, n);
        GP_LAUNCH_CHECK(h);
        schur_kernel<<<1, 1024, 0, h->stream>>>(kv, v, n, h->sf_fit[a] * h->sf_fit[a] + h->noise[a], scal + 2 * a);
        GP_LAUNCH_CHECK(h);
    }
    GP_CUDA(h, cudaMemcpyAsync(hs.data(), scal, 2 * E * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    // The Schur complement s = kappa - k^T Ky^-1 k is a difference computed from an EXPLICIT inverse, so it carries an
    // absolute error of about cond(Ky) eps kappa with cond(Ky) <= n sf^2 / noise + 1.  The true value is >= noise; when
    // the error bound is not far below it (the reference's experiments use sigma_n = 1e-5, i.e. noise 1e-10) 1/s and
    // the whole rank-1 update would be garbage without any visible sign, so the caller is told to refit instead.
    for (int a = 0; a < E; ++a) {
        const double kappa = h->sf_fit[a] * h->sf_fit[a] + h->noise[a];
        const double tol = 1e3 * 2.220446049250313e-16 * (double)n * kappa / h->noise[a];
        if (!(hs[2 * a] > tol * kappa) || !std::isfinite(hs[2 * a])) return GPMPC_REFIT_NEEDED;
    }
    // second pass: apply
    GP_CUDA(h, cudaMemcpyAsync(h->X.as<double>() + (size_t)n * D, xh, D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    for (int a = 0; a < E; ++a) {
        double *Kinv = h->Kinv.as<double>() + a * mat;
        knew_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(h->X.as<double>(), n, D, ks_arg(h, a), xs, kv);
        GP_LAUNCH_CHECK(h);
        gpmpc_rowdot_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(Kinv, ld, n, kv, v, n);
        GP_LAUNCH_CHECK(h);
        dim3 blk(32, 8), grid((n + 1 + 31) / 32, (n + 1 + 7) / 8);
        border_update_kernel<<<grid, blk, 0, h->stream>>>(Kinv, ld, n, v, scal + 2 * a);
        GP_LAUNCH_CHECK(h);
        GP_CUDA(h, cudaMemcpyAsync(h->Y.as<double>() + (size_t)a * ld + n, &yh[a], sizeof(double), cudaMemcpyHostToDevice, h->stream));
        h->logdet[a] += std::log(hs[2 * a]);
    }
    GP_CUDA(h, cudaStreamSynchronize(h->stream));     // xh / yh are stack buffers
    h->n = n + 1;
    for (int a = 0; a < E; ++a) {
        // beta = Ky^-1 y over the n+1 points, then the weight matrix
        gpmpc_rowdot_kernel<<<(n + 1 + 7) / 8, 256, 0, h->stream>>>(h->Kinv.as<double>() + a * mat, ld, n + 1,
                                                                   h->Y.as<double>() + (size_t)a * ld,
                                                                   h->beta.as<double>() + (size_t)a * ld, n + 1);
        GP_LAUNCH_CHECK(h);
        if ((rc = derive_weights(h, a))) return rc;
    }
    h->tape_B = h->tape_H = 0; h->fc_B = h->fc_H = 0;
    return GPMPC_OK;
}

extern "C" int gpmpc_set_propagation_hypers(gpmpc_handle h, const double *lambdas, const double *sigma_f)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!lambdas || !sigma_f) return fail(h, GPMPC_ERR_INVALID, "gpmpc_set_propagation_hypers: null");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int D = h->D, E = h->E;
    double lam[kMaxE * kMaxD], sf[kMaxE];
    int rc;
    if ((rc = fetch_host(h, lambdas, lam, (size_t)E * D))) return rc;
    if ((rc = fetch_host(h, sigma_f, sf, E))) return rc;
    bool changed[kMaxE];
    for (int a = 0; a < E; ++a) {
        changed[a] = false;
        for (int k = 0; k < D; ++k) {
            if (!(lam[a * D + k] > 0.0)) return fail(h, GPMPC_ERR_INVALID, "lambdas must be positive");
            if (h->lam_prop[a][k] != lam[a * D + k]) changed[a] = true;
            h->lam_prop[a][k] = lam[a * D + k];
        }
        h->sf_prop[a] = sf[a];
    }
    if ((rc = upload_prop_hypers(h))) return rc;
    if (h->fitted)
        for (int a = 0; a < E; ++a)
            if (changed[a] && (rc = derive_weights(h, a))) return rc;
    h->tape_B = h->tape_H = 0; h->fc_B = h->fc_H = 0;
    return GPMPC_OK;
}

namespace gpmpc {
__global__ void copy_sub_kernel(const double *__restrict__ src, int lds, int n, double *__restrict__ dst, int ldd)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n) dst[(size_t)i * ldd + j] = src[(size_t)i * lds + j];
}
}  // namespace gpmpc

extern "C" int gpmpc_get_matrix(gpmpc_handle h, int which, int a, double *out)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_get_matrix: not fitted");
    if (a < 0 || a >= h->E || !out) return fail(h, GPMPC_ERR_INVALID, "gpmpc_get_matrix: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n, ld = h->ld;
    const size_t mat = (size_t)ld * ld;
    const bool host = !is_device_ptr(out);
    if (which == GPMPC_MAT_BETA) {
        GP_CUDA(h, cudaMemcpyAsync(out, h->beta.as<double>() + (size_t)a * ld, (size_t)n * sizeof(double), cudaMemcpyDefault, h->stream));
        if (host) GP_CUDA(h, cudaStreamSynchronize(h->stream));
        return GPMPC_OK;
    }
    double *dev = out;
    if (host) { GP_CUDA(h, h->stage_out.reserve((size_t)n * n * sizeof(double))); dev = h->stage_out.as<double>(); }
    dim3 blk(32, 8), grid((n + 31) / 32, (n + 7) / 8);
    int rc = GPMPC_OK;
    switch (which) {
        case GPMPC_MAT_KF: rc = gram_into(h, a, dev, n, false); break;
        case GPMPC_MAT_KY: rc = gram_into(h, a, dev, n, true); break;
        case GPMPC_MAT_KY_INV:
            copy_sub_kernel<<<grid, blk, 0, h->stream>>>(h->Kinv.as<double>() + a * mat, ld, n, dev, n);
            GP_LAUNCH_CHECK(h);
            break;
        default: return fail(h, GPMPC_ERR_INVALID, "gpmpc_get_matrix: unknown selector");
    }
    if (rc) return rc;
    if (host) {
        GP_CUDA(h, cudaMemcpyAsync(out, dev, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return GPMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// K(X*, X) and the posterior (src/gpr.py:253-332)
// ---------------------------------------------------------------------------------------------
namespace gpmpc {
__global__ void kstar_kernel(const double *__restrict__ Xs, int p, const double *__restrict__ X, int n, int D, KsArg hp,
                             double *__restrict__ out, int ldo, int pp, int np)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= pp || j >= np) return;
    double v = 0.0;
    if (i < p && j < n) {
        double q = 0.0;
        for (int k = 0; k < D; ++k) {
            const double d = Xs[(size_t)i * D + k] - X[(size_t)j * D + k];
            q = fma(d * d, hp.inv_lam[k], q);
        }
        v = hp.sf2 * exp(-0.5 * q);
    }
    out[(size_t)i * ldo + j] = v;
}
// cov[i][j] = K**[i][j] - sum_k T[i][k] Ks[j][k] (+ noise on the diagonal)
__global__ void post_cov_kernel(const double *__restrict__ Xs, int p, int D, KsArg hp, const double *__restrict__ TK, int ldt,
                                double noise, double *__restrict__ cov)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= p || j >= p) return;
    double q = 0.0;
    for (int k = 0; k < D; ++k) {
        const double d = Xs[(size_t)i * D + k] - Xs[(size_t)j * D + k];
        q = fma(d * d, hp.inv_lam[k], q);
    }
    double v = hp.sf2 * exp(-0.5 * q) - TK[(size_t)i * ldt + j];
    if (i == j) v += noise;
    cov[(size_t)i * p + j] = v;
}
}  // namespace gpmpc

extern "C" int gpmpc_kernel_matrix(gpmpc_handle h, int a, int p, const double *Xs, double *out)
{
    return gpmpc_kernel_matrix_ex(h, a, p, Xs, nullptr, out);
}

extern "C" int gpmpc_kernel_matrix_ex(gpmpc_handle h, int a, int p, const double *Xs, const double *hyp, double *out)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_kernel_matrix: not fitted");
    if (a < 0 || a >= h->E || p <= 0 || !Xs || !out) return fail(h, GPMPC_ERR_INVALID, "gpmpc_kernel_matrix: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n, D = h->D;
    const double *Xd = Xs;
    size_t off = 0;
    GP_CUDA(h, h->stage_in.reserve((size_t)p * D * sizeof(double) + 256));
    if (!is_device_ptr(Xs)) {
        GP_CUDA(h, cudaMemcpyAsync(h->stage_in.p, Xs, (size_t)p * D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
        Xd = h->stage_in.as<double>();
    }
    (void)off;
    const bool host = !is_device_ptr(out);
    double *dev = out;
    if (host) { GP_CUDA(h, h->stage_out.reserve((size_t)p * n * sizeof(double))); dev = h->stage_out.as<double>(); }
    dim3 blk(32, 8), grid((n + 31) / 32, (p + 7) / 8);
    KsArg ka = ks_arg(h, a);
    if (hyp) {
        double hh[kMaxD + 2];
        int rc = fetch_host(h, hyp, hh, (size_t)D + 2);
        if (rc) return rc;
        ka = ks_arg_from(D, hh);
    }
    kstar_kernel<<<grid, blk, 0, h->stream>>>(Xd, p, h->X.as<double>(), n, D, ka, dev, n, p, n);
    GP_LAUNCH_CHECK(h);
    if (host) {
        GP_CUDA(h, cudaMemcpyAsync(out, dev, (size_t)p * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        GP_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return GPMPC_OK;
}

extern "C" int gpmpc_predict(gpmpc_handle h, int a, int p, const double *Xs, double *mean, double *cov, int add_noise)
{
    return gpmpc_predict_ex(h, a, p, Xs, nullptr, nullptr, mean, cov, add_noise);
}

extern "C" int gpmpc_predict_ex(gpmpc_handle h, int a, int p, const double *Xs, const double *resid, const double *hyp,
                                double *mean, double *cov, int add_noise)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_predict: not fitted");
    if (a < 0 || a >= h->E || p <= 0 || !Xs || !mean) return fail(h, GPMPC_ERR_INVALID, "gpmpc_predict: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n, D = h->D, ld = h->ld;
    const int pp = round_up(p, kTile);
    const size_t mat = (size_t)ld * ld;
    // workspace (gbuf): Xs [pp*D] | Ks [pp, ld] | T [pp, ld] | TK [pp, pp] | mean [pp] | cov [p*p] | resid [ld]
    const size_t cnt = (size_t)pp * D + 2 * (size_t)pp * ld + (size_t)pp * pp + pp + (size_t)p * p + ld + 64;
    GP_CUDA(h, h->gbuf.reserve(cnt * sizeof(double)));
    double *w = h->gbuf.as<double>();
    double *Xd = w; w += (size_t)pp * D;
    double *Ks = w; w += (size_t)pp * ld;
    double *T = w; w += (size_t)pp * ld;
    double *TK = w; w += (size_t)pp * pp;
    double *md = w; w += pp;
    double *cd = w; w += (size_t)p * p;
    double *rd = w;
    const double *yv = h->Y.as<double>() + (size_t)a * ld;      // targets, or the caller's residual y - f_nom(X)
    if (resid) { GP_CUDA(h, to_device(h, rd, resid, (size_t)n * sizeof(double))); yv = rd; }
    KsArg ka = ks_arg(h, a);
    double noise = h->noise[a];
    if (hyp) {
        double hh[kMaxD + 2];
        int rch = fetch_host(h, hyp, hh, (size_t)D + 2);
        if (rch) return rch;
        ka = ks_arg_from(D, hh);
        noise = hh[D + 1];
    }
    GP_CUDA(h, cudaMemsetAsync(Xd, 0, (size_t)pp * D * sizeof(double), h->stream));
    GP_CUDA(h, to_device(h, Xd, Xs, (size_t)p * D * sizeof(double)));
    dim3 blk(32, 8), grid((ld + 31) / 32, (pp + 7) / 8);
    kstar_kernel<<<grid, blk, 0, h->stream>>>(Xd, p, h->X.as<double>(), n, D, ka, Ks, ld, pp, ld);
    GP_LAUNCH_CHECK(h);
    // T = Ks Ky^-1 (Ky^-1 symmetric -> NT form), the order the reference multiplies in (src/gpr.py:306)
    int rc = dgemm_nt(h, pp, ld, ld, 1.0, Ks, ld, h->Kinv.as<double>() + a * mat, ld, 0.0, T, ld, false, 0);
    if (rc) return rc;
    // mean = T y   (rows of the padded region of Ky^-1 are identity, but Ks is zero there)
    gpmpc_rowdot_kernel<<<(pp + 7) / 8, 256, 0, h->stream>>>(T, ld, n, yv, md, pp);
    GP_LAUNCH_CHECK(h);
    GP_CUDA(h, cudaMemcpyAsync(mean, md, (size_t)p * sizeof(double), cudaMemcpyDefault, h->stream));
    if (cov) {
        rc = dgemm_nt(h, pp, pp, ld, 1.0, T, ld, Ks, ld, 0.0, TK, pp, false, 0);
        if (rc) return rc;
        dim3 cgrid((p + 31) / 32, (p + 7) / 8);
        const bool host = !is_device_ptr(cov);
        double *dev = host ? cd : cov;
        post_cov_kernel<<<cgrid, blk, 0, h->stream>>>(Xd, p, D, ka, TK, pp, add_noise ? noise : 0.0, dev);
        GP_LAUNCH_CHECK(h);
        if (host) GP_CUDA(h, cudaMemcpyAsync(cov, dev, (size_t)p * p * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    }
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

__global__ void gpmpc_rowdot_kernel(const double *__restrict__ A, int ld, int n, const double *__restrict__ x,
                                    double *__restrict__ y, int rows)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    double s = 0.0;
    for (int j = lane; j < n; j += 32) s = fma(A[(size_t)row * ld + j], x[j], s);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) y[row] = s;
}

// ---------------------------------------------------------------------------------------------
// Full-covariance / stateless moment matching (generic kernels; not the hot path)
// ---------------------------------------------------------------------------------------------
namespace gpmpc {

// host-side LU helpers for D x D matrices
static double host_lu(int n, const double *Ain, double *inv)
{
    double A[kMaxD * kMaxD];
    int piv[kMaxD];
    double det = 1.0;
    std::memcpy(A, Ain, sizeof(double) * n * n);
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(A[r * n + c]) > std::fabs(A[p * n + c])) p = r;
        piv[c] = p;
        if (p != c) { for (int k = 0; k < n; ++k) std::swap(A[c * n + k], A[p * n + k]); det = -det; }
        det *= A[c * n + c];
        const double dinv = 1.0 / A[c * n + c];
        for (int r = c + 1; r < n; ++r) {
            const double f = A[r * n + c] * dinv;
            A[r * n + c] = f;
            for (int k = c + 1; k < n; ++k) A[r * n + k] -= f * A[c * n + k];
        }
    }
    if (inv) {
        for (int col = 0; col < n; ++col) {
            double x[kMaxD];
            for (int r = 0; r < n; ++r) x[r] = (r == col) ? 1.0 : 0.0;
            for (int c = 0; c < n; ++c) if (piv[c] != c) std::swap(x[c], x[piv[c]]);
            for (int r = 0; r < n; ++r) for (int k = 0; k < r; ++k) x[r] -= A[r * n + k] * x[k];
            for (int r = n - 1; r >= 0; --r) {
                for (int k = r + 1; k < n; ++k) x[r] -= A[r * n + k] * x[k];
                x[r] /= A[r * n + r];
            }
            for (int r = 0; r < n; ++r) inv[r * n + col] = x[r];
        }
    }
    return det;
}

struct FullArg {
    double A[kMaxD * kMaxD];     // symmetrised quadratic form matrix
    double u[kMaxD];
    double inv_lam[kMaxD];
    double scale;                // factor in front of the quadratic form inside exp
    int D;
};

// l_j = pref * exp(-1/2 v_j^T Bm v_j), v_j = u - x_j; bl_j = beta_j l_j
__global__ void mean_full_kernel(const double *__restrict__ X, int n, FullArg fa, double pref, const double *__restrict__ beta,
                                 double *__restrict__ l, double *__restrict__ bl)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double v[kMaxD];
    for (int k = 0; k < fa.D; ++k) v[k] = fa.u[k] - X[(size_t)j * fa.D + k];
    double q = 0.0;
    for (int r = 0; r < fa.D; ++r) {
        double s = 0.0;
        for (int k = 0; k < fa.D; ++k) s = fma(fa.A[r * fa.D + k], v[k], s);
        q = fma(v[r], s, q);
    }
    const double lj = pref * exp(fa.scale * q);
    if (l) l[j] = lj;
    bl[j] = beta[j] * lj;
}

// rowsum[i] = sum_j w_ij exp(scale * s_ij^T A s_ij), s_ij = v_i + v_j
//   RAW : w_ij = (Kinv_ij - beta_i beta_j) exp(-1/4 (x_i-x_j)^T Lam^-1 (x_i-x_j)), all j
//   !RAW: w_ij = Wt_ij (upper-triangular weights), j >= i
template <bool RAW>
__global__ void __launch_bounds__(256) pairs_full_kernel(const double *__restrict__ X, int n, FullArg fa,
                                                          const double *__restrict__ Wm, int ldw,
                                                          const double *__restrict__ beta, double *__restrict__ rowsum)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int D = fa.D;
    double vi[kMaxD], xi[kMaxD];
    for (int k = 0; k < D; ++k) { xi[k] = X[(size_t)i * D + k]; vi[k] = fa.u[k] - xi[k]; }
    double acc = 0.0;
    const int j0 = RAW ? 0 : (i / 32) * 32;
    for (int j = j0 + lane; j < n; j += 32) {
        if (!RAW && j < i) continue;
        double s[kMaxD], el = 0.0;
        for (int k = 0; k < D; ++k) {
            const double xj = X[(size_t)j * D + k];
            s[k] = vi[k] + (fa.u[k] - xj);
            const double d = xi[k] - xj;
            el = fma(d * d, fa.inv_lam[k], el);
        }
        double q = 0.0;
        for (int r = 0; r < D; ++r) {
            double t = 0.0;
            for (int k = 0; k < D; ++k) t = fma(fa.A[r * D + k], s[k], t);
            q = fma(s[r], t, q);
        }
        // RAW: Wm is a row-major [.., ldw] matrix; else the tile-major upper-triangular Wt (ldw = padded n)
        double w = RAW ? Wm[(size_t)i * ldw + j]
                       : Wm[wt_tile_index(i / kPairTile, j / kPairTile, ldw / kPairTile) * kPairTile * kPairTile +
                            (i % kPairTile) * kPairTile + (j % kPairTile)];
        if (RAW) w = (w - beta[i] * beta[j]) * exp(-0.25 * el);
        acc = fma(w, exp(fa.scale * q), acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rowsum[i] = acc;
}

// cross-covariance rows: rowsum[i] = beta1_i k1_i sum_j beta2_j k2_j exp(1/2 z_ij^T T z_ij)
struct CovArg {
    double T[kMaxD * kMaxD];     // symmetrised T = R^-1 S
    double u[kMaxD], il1[kMaxD], il2[kMaxD];
    int D, bugcompat;
};
__global__ void __launch_bounds__(256) cov_rows_kernel(const double *__restrict__ X, int n, CovArg ca,
                                                        const double *__restrict__ beta1, const double *__restrict__ beta2,
                                                        double *__restrict__ rowsum)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int D = ca.D;
    // reference torch form (bugcompat): z = Lam2^-1 (x_i-u) + Lam1^-1 (x_j-u) in the cross term, i.e. the roles of
    // the two length-scale sets are swapped between i and j inside the exponent only (uncertainty_prop.py:446).
    double zi[kMaxD], ci[kMaxD], ki = 0.0;
    for (int k = 0; k < D; ++k) {
        ci[k] = X[(size_t)i * D + k] - ca.u[k];
        zi[k] = ci[k] * ca.il1[k];
        ki = fma(ci[k] * ci[k], ca.il1[k], ki);
    }
    double acc = 0.0;
    for (int j = lane; j < n; j += 32) {
        double z[kMaxD], kj = 0.0, cj[kMaxD];
        for (int k = 0; k < D; ++k) {
            cj[k] = X[(size_t)j * D + k] - ca.u[k];
            kj = fma(cj[k] * cj[k], ca.il2[k], kj);
            z[k] = zi[k] + cj[k] * ca.il2[k];
        }
        double q = 0.0;
        if (!ca.bugcompat) {
            for (int r = 0; r < D; ++r) {
                double t = 0.0;
                for (int k = 0; k < D; ++k) t = fma(ca.T[r * D + k], z[k], t);
                q = fma(z[r], t, q);
            }
        } else {
            // A_z1[i] + 2 z2_i^T T z1_j + A_z2[j]   with z1 = Lam1^-1 c, z2 = Lam2^-1 c
            double z1i[kMaxD], z2i[kMaxD], z1j[kMaxD], z2j[kMaxD];
            for (int k = 0; k < D; ++k) { z1i[k] = ci[k] * ca.il1[k]; z2i[k] = ci[k] * ca.il2[k]; z1j[k] = cj[k] * ca.il1[k]; z2j[k] = cj[k] * ca.il2[k]; }
            double a1 = 0.0, a2 = 0.0, cr = 0.0;
            for (int r = 0; r < D; ++r) {
                double t1 = 0.0, t2 = 0.0, t3 = 0.0;
                for (int k = 0; k < D; ++k) {
                    t1 = fma(ca.T[r * D + k], z1i[k], t1);
                    t2 = fma(ca.T[r * D + k], z2j[k], t2);
                    t3 = fma(ca.T[r * D + k], z1j[k], t3);
                }
                a1 = fma(z1i[r], t1, a1); a2 = fma(z2j[r], t2, a2); cr = fma(z2i[r], t3, cr);
            }
            q = a1 + 2.0 * cr + a2;
        }
        acc = fma(beta2[j], exp(-0.5 * kj + 0.5 * q), acc);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) rowsum[i] = beta1[i] * exp(-0.5 * ki) * acc;
}

// deterministic sum of a vector (single block, fixed tree)
__global__ void __launch_bounds__(1024) sum_kernel(const double *__restrict__ v, int n, double *__restrict__ out)
{
    __shared__ double sm[1024];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 1024) s += v[i];
    sm[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sm[0];
}

struct FullSetup { FullArg mean_arg, var_arg; double mean_pref, var_pref; };

// Host-side D x D algebra of src/tools/uncertainty_prop.py:329-335,374-377 for a full S.
static void full_setup(int D, const double *lam, const double *u, const double *S, double sf, FullSetup &fs)
{
    double M[kMaxD * kMaxD], Inv[kMaxD * kMaxD];
    // mean: (S + Lam)^-1, det(Lam^-1 S + I)
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) M[r * D + k] = S[r * D + k] + (r == k ? lam[r] : 0.0);
    host_lu(D, M, Inv);
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) fs.mean_arg.A[r * D + k] = 0.5 * (Inv[r * D + k] + Inv[k * D + r]);
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) M[r * D + k] = S[r * D + k] / lam[r] + (r == k ? 1.0 : 0.0);
    fs.mean_pref = sf * sf * std::pow(host_lu(D, M, nullptr), -0.5);
    fs.mean_arg.scale = -0.5;
    // variance: (Lam/2 + S)^-1, det(2 Lam^-1 S + I)
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) M[r * D + k] = S[r * D + k] + (r == k ? 0.5 * lam[r] : 0.0);
    host_lu(D, M, Inv);
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) fs.var_arg.A[r * D + k] = 0.5 * (Inv[r * D + k] + Inv[k * D + r]);
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) M[r * D + k] = 2.0 * S[r * D + k] / lam[r] + (r == k ? 1.0 : 0.0);
    fs.var_pref = sf * sf * sf * sf * std::pow(host_lu(D, M, nullptr), -0.5);
    fs.var_arg.scale = -0.125;
    for (FullArg *fa : {&fs.mean_arg, &fs.var_arg}) {
        fa->D = D;
        for (int k = 0; k < D; ++k) { fa->u[k] = u[k]; fa->inv_lam[k] = 1.0 / lam[k]; }
    }
}

}  // namespace gpmpc

static int moment_match_impl(gpmpc_handle h, int B, const double *U, const double *S, int s_is_full,
                             double *mean, double *var);
static int cov_core(gpmpc_ctx *h, int n, int D, const double *l1, const double *l2, const double *uh, const double *Sh,
                    const double *Xd, const double *b1, const double *b2, double mean1, double mean2, double sf1,
                    double sf2, int bugcompat, double *rows, double *scal, double *out_host);

extern "C" int gpmpc_moment_match(gpmpc_handle h, int B, const double *U, const double *S, int s_is_full,
                                  double *mean, double *var)
{
    return moment_match_impl(h, B, U, S, s_is_full, mean, var);
}

static int moment_match_impl(gpmpc_handle h, int B, const double *U, const double *S, int s_is_full,
                             double *mean, double *var)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_moment_match: not fitted");
    if (B <= 0 || !U || !S || !mean || !var) return fail(h, GPMPC_ERR_INVALID, "gpmpc_moment_match: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int D = h->D, E = h->E, n = h->n, ld = h->ld;
    if (!s_is_full) {
        const size_t cnt = 2 * (size_t)B * D;
        GP_CUDA(h, h->stage_in.reserve(cnt * sizeof(double) + 512));
        const double *Ud = U, *Sd = S;
        double *st = h->stage_in.as<double>();
        if (!is_device_ptr(U)) { GP_CUDA(h, cudaMemcpyAsync(st, U, (size_t)B * D * sizeof(double), cudaMemcpyHostToDevice, h->stream)); Ud = st; }
        if (!is_device_ptr(S)) { GP_CUDA(h, cudaMemcpyAsync(st + (size_t)B * D, S, (size_t)B * D * sizeof(double), cudaMemcpyHostToDevice, h->stream)); Sd = st + (size_t)B * D; }
        return gpmpc_moment_match_diag_internal(h, B, Ud, Sd, mean, var);
    }
    // full input covariance: the batched full-covariance step (fullcov.cu); the variances are the diagonal of its result
    std::vector<double> ch((size_t)B * E * E), vh((size_t)B * E);
    int rc = gpmpc_moment_match_cov(h, B, U, S, mean, ch.data());
    if (rc) return rc;
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < E; ++a) vh[(size_t)b * E + a] = ch[((size_t)b * E + a) * E + a];
    GP_CUDA(h, cudaMemcpyAsync(var, vh.data(), vh.size() * sizeof(double), cudaMemcpyDefault, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

extern "C" int gpmpc_moment_match_raw(gpmpc_handle h, int n, int D, const double *Kinv, const double *lambdas,
                                      const double *u, const double *S, const double *X, const double *y,
                                      double sigma_f, double *mean, double *var, double *beta, double *l)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (n <= 0 || D < 1 || D > kMaxD || !Kinv || !lambdas || !u || !S || !X || !mean || (!y && !beta))
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_moment_match_raw: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    double lam[kMaxD], uh[kMaxD], Sh[kMaxD * kMaxD];
    int rc;
    if ((rc = fetch_host(h, lambdas, lam, D))) return rc;
    if ((rc = fetch_host(h, u, uh, D))) return rc;
    if ((rc = fetch_host(h, S, Sh, (size_t)D * D))) return rc;
    // workspace: Kinv [n*n] | X [n*D] | y [n] | beta [n] | l [n] | bl [n] | rows [n] | scal [4]
    const size_t cnt = (size_t)n * n + (size_t)n * D + 5 * (size_t)n + 16;
    GP_CUDA(h, h->gbuf.reserve(cnt * sizeof(double)));
    double *w = h->gbuf.as<double>();
    const double *Kd = Kinv, *Xd = X, *yd = y;
    if (!is_device_ptr(Kinv)) { GP_CUDA(h, cudaMemcpyAsync(w, Kinv, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, h->stream)); Kd = w; }
    w += (size_t)n * n;
    if (!is_device_ptr(X)) { GP_CUDA(h, cudaMemcpyAsync(w, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, h->stream)); Xd = w; }
    w += (size_t)n * D;
    if (y && !is_device_ptr(y)) { GP_CUDA(h, cudaMemcpyAsync(w, y, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream)); yd = w; }
    w += n;
    double *bd = w; w += n;
    double *ld_ = w; w += n;
    double *bl = w; w += n;
    double *rows = w; w += n;
    double *scal = w;
    double m_in = 0.0;
    if (y) {
        // beta = Kinv y   (src/tools/uncertainty_prop.py:327)
        gpmpc_rowdot_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(Kd, n, n, yd, bd, n);
        GP_LAUNCH_CHECK(h);
    } else {
        // caller supplies beta and the mean (variance_prop_torch signature, uncertainty_prop.py:341)
        GP_CUDA(h, cudaMemcpyAsync(bd, beta, (size_t)n * sizeof(double), cudaMemcpyDefault, h->stream));
        if ((rc = fetch_host(h, mean, &m_in, 1))) return rc;
    }
    FullSetup fs;
    full_setup(D, lam, uh, Sh, sigma_f, fs);
    if (y) {
        mean_full_kernel<<<(n + 127) / 128, 128, 0, h->stream>>>(Xd, n, fs.mean_arg, fs.mean_pref, bd, ld_, bl);
        GP_LAUNCH_CHECK(h);
        sum_kernel<<<1, 1024, 0, h->stream>>>(bl, n, scal);
        GP_LAUNCH_CHECK(h);
    }
    if (var) {
        pairs_full_kernel<true><<<(n + 7) / 8, 256, 0, h->stream>>>(Xd, n, fs.var_arg, Kd, n, bd, rows);
        GP_LAUNCH_CHECK(h);
        sum_kernel<<<1, 1024, 0, h->stream>>>(rows, n, scal + 1);
        GP_LAUNCH_CHECK(h);
    }
    double r2[2] = {0.0, 0.0};
    GP_CUDA(h, cudaMemcpyAsync(r2, scal, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (y && beta) GP_CUDA(h, cudaMemcpyAsync(beta, bd, (size_t)n * sizeof(double), cudaMemcpyDefault, h->stream));
    if (y && l) GP_CUDA(h, cudaMemcpyAsync(l, ld_, (size_t)n * sizeof(double), cudaMemcpyDefault, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    const double m = y ? r2[0] : m_in;
    const double v = sigma_f * sigma_f - fs.var_pref * r2[1] - m * m;
    if (y) GP_CUDA(h, cudaMemcpyAsync(mean, &m, sizeof(double), cudaMemcpyDefault, h->stream));
    if (var) GP_CUDA(h, cudaMemcpyAsync(var, &v, sizeof(double), cudaMemcpyDefault, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// cov = sf1^2 sf2^2 |R|^-1/2 sum_ij beta1_i beta2_j k1_i k2_j exp(1/2 z_ij^T T z_ij) - mean1 mean2 on device vectors
// (src/tools/uncertainty_prop.py:212-236 / 433-465); rows [n] and scal [1] are device scratch.
static int cov_core(gpmpc_ctx *h, int n, int D, const double *l1, const double *l2, const double *uh, const double *Sh,
                    const double *Xd, const double *b1, const double *b2, double mean1, double mean2, double sf1,
                    double sf2, int bugcompat, double *rows, double *scal, double *out_host)
{
    // R = S (Lam1^-1 + Lam2^-1) + I, T = R^-1 S  (:439-443)
    double R[kMaxD * kMaxD], Ri[kMaxD * kMaxD], T[kMaxD * kMaxD];
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) R[r * D + k] = Sh[r * D + k] * (1.0 / l1[k] + 1.0 / l2[k]) + (r == k ? 1.0 : 0.0);
    const double det = host_lu(D, R, Ri);
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k) {
        double acc = 0.0;
        for (int q = 0; q < D; ++q) acc += Ri[r * D + q] * Sh[q * D + k];
        T[r * D + k] = acc;
    }
    CovArg ca;
    ca.D = D; ca.bugcompat = bugcompat ? 1 : 0;
    for (int r = 0; r < D; ++r) for (int k = 0; k < D; ++k)
        ca.T[r * D + k] = bugcompat ? T[r * D + k] : 0.5 * (T[r * D + k] + T[k * D + r]);
    for (int k = 0; k < D; ++k) { ca.u[k] = uh[k]; ca.il1[k] = 1.0 / l1[k]; ca.il2[k] = 1.0 / l2[k]; }
    cov_rows_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(Xd, n, ca, b1, b2, rows);
    GP_LAUNCH_CHECK(h);
    sum_kernel<<<1, 1024, 0, h->stream>>>(rows, n, scal);
    GP_LAUNCH_CHECK(h);
    double r = 0.0;
    GP_CUDA(h, cudaMemcpyAsync(&r, scal, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    *out_host = sf1 * sf1 * sf2 * sf2 * std::pow(det, -0.5) * r - mean1 * mean2;
    return GPMPC_OK;
}

extern "C" int gpmpc_covariance_raw(gpmpc_handle h, int n, int D, const double *lambdas1, const double *lambdas2,
                                    const double *u, const double *S, const double *X, double mean1, double mean2,
                                    const double *beta1, const double *beta2, double sigma_f1, double sigma_f2,
                                    int bugcompat, double *cov)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (n <= 0 || D < 1 || D > kMaxD || !lambdas1 || !lambdas2 || !u || !S || !X || !beta1 || !beta2 || !cov)
        return fail(h, GPMPC_ERR_INVALID, "gpmpc_covariance_raw: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    double l1[kMaxD], l2[kMaxD], uh[kMaxD], Sh[kMaxD * kMaxD];
    int rc;
    if ((rc = fetch_host(h, lambdas1, l1, D))) return rc;
    if ((rc = fetch_host(h, lambdas2, l2, D))) return rc;
    if ((rc = fetch_host(h, u, uh, D))) return rc;
    if ((rc = fetch_host(h, S, Sh, (size_t)D * D))) return rc;
    const size_t cnt = (size_t)n * D + 3 * (size_t)n + 16;
    GP_CUDA(h, h->gbuf.reserve(cnt * sizeof(double)));
    double *w = h->gbuf.as<double>();
    const double *Xd = X, *b1 = beta1, *b2 = beta2;
    if (!is_device_ptr(X)) { GP_CUDA(h, cudaMemcpyAsync(w, X, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice, h->stream)); Xd = w; }
    w += (size_t)n * D;
    if (!is_device_ptr(beta1)) { GP_CUDA(h, cudaMemcpyAsync(w, beta1, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream)); b1 = w; }
    w += n;
    if (!is_device_ptr(beta2)) { GP_CUDA(h, cudaMemcpyAsync(w, beta2, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream)); b2 = w; }
    w += n;
    double c = 0.0;
    if ((rc = cov_core(h, n, D, l1, l2, uh, Sh, Xd, b1, b2, mean1, mean2, sigma_f1, sigma_f2, bugcompat, w, w + n, &c))) return rc;
    GP_CUDA(h, cudaMemcpyAsync(cov, &c, sizeof(double), cudaMemcpyDefault, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// Marginal likelihood and its gradient (src/gpr.py:240-251; the reference differentiates it with autograd)
// ---------------------------------------------------------------------------------------------
namespace gpmpc {
struct MlArg { double inv_lam[kMaxD]; double sf2; int D; };
// rows[(k)*n + i] = sum_j B_ij Kf_ij d_ijk^2 / lam_k (k < D),  rows[D*n + i] = sum_j B_ij Kf_ij,
// rows[(D+1)*n + i] = B_ii,   B = alpha alpha^T - Ky^-1
__global__ void __launch_bounds__(256) ml_grad_rows_kernel(const double *__restrict__ X, int n, MlArg ma,
                                                            const double *__restrict__ Kinv, int ldk,
                                                            const double *__restrict__ alpha, double *__restrict__ rows)
{
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int D = ma.D;
    double xi[kMaxD], acc[kMaxD + 1];
    for (int k = 0; k < D; ++k) { xi[k] = X[(size_t)i * D + k]; acc[k] = 0.0; }
    acc[D] = 0.0;
    const double ai = alpha[i];
    for (int j = lane; j < n; j += 32) {
        double dk[kMaxD], q = 0.0;
        for (int k = 0; k < D; ++k) {
            const double d = xi[k] - X[(size_t)j * D + k];
            dk[k] = d * d * ma.inv_lam[k];
            q += dk[k];
        }
        const double bk = (ai * alpha[j] - Kinv[(size_t)i * ldk + j]) * ma.sf2 * exp(-0.5 * q);
        for (int k = 0; k < D; ++k) acc[k] = fma(bk, dk[k], acc[k]);
        acc[D] += bk;
    }
    for (int k = 0; k <= D; ++k) {
        double v = acc[k];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) rows[(size_t)k * n + i] = v;
    }
    if (lane == 0) rows[(size_t)(D + 1) * n + i] = ai * ai - Kinv[(size_t)i * ldk + i];
}
__global__ void dot_rows_kernel(const double *__restrict__ a, const double *__restrict__ b, int n, double *__restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}
}  // namespace gpmpc

extern "C" int gpmpc_marginal_likelihood(gpmpc_handle h, int a, const double *resid, double *ml, double *grad)
{
    if (!h) return GPMPC_ERR_INVALID;
    if (!h->fitted) return fail(h, GPMPC_ERR_NOT_FIT, "gpmpc_marginal_likelihood: not fitted");
    if (a < 0 || a >= h->E || !ml) return fail(h, GPMPC_ERR_INVALID, "gpmpc_marginal_likelihood: bad argument");
    GP_CUDA(h, cudaSetDevice(h->device));
    const int n = h->n, D = h->D, ld = h->ld;
    const size_t mat = (size_t)ld * ld;
    // workspace: r [n] | alpha [n] | prod [n] | rows [(D+2) n] | scal [D+4]
    const size_t cnt = (size_t)(D + 5) * n + D + 8;
    GP_CUDA(h, h->gbuf.reserve(cnt * sizeof(double)));
    double *w = h->gbuf.as<double>();
    double *r = w; w += n;
    double *alpha = w; w += n;
    double *prod = w; w += n;
    double *rows = w; w += (size_t)(D + 2) * n;
    double *scal = w;
    const double *Kinv = h->Kinv.as<double>() + a * mat;
    const double *rd, *ad;
    if (resid) {
        GP_CUDA(h, cudaMemcpyAsync(r, resid, (size_t)n * sizeof(double), cudaMemcpyDefault, h->stream));
        gpmpc_rowdot_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(Kinv, ld, n, r, alpha, n);
        GP_LAUNCH_CHECK(h);
        rd = r; ad = alpha;
    } else {
        rd = h->Y.as<double>() + (size_t)a * ld;
        ad = h->beta.as<double>() + (size_t)a * ld;
    }
    dot_rows_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(rd, ad, n, prod);
    GP_LAUNCH_CHECK(h);
    sum_kernel<<<1, 1024, 0, h->stream>>>(prod, n, scal);
    GP_LAUNCH_CHECK(h);
    if (grad) {
        MlArg ma;
        ma.D = D; ma.sf2 = h->sf_fit[a] * h->sf_fit[a];
        for (int k = 0; k < kMaxD; ++k) ma.inv_lam[k] = k < D ? 1.0 / h->lam_fit[a][k] : 0.0;
        ml_grad_rows_kernel<<<(n + 7) / 8, 256, 0, h->stream>>>(h->X.as<double>(), n, ma, Kinv, ld, ad, rows);
        GP_LAUNCH_CHECK(h);
        for (int k = 0; k < D + 2; ++k) {
            sum_kernel<<<1, 1024, 0, h->stream>>>(rows + (size_t)k * n, n, scal + 1 + k);
            GP_LAUNCH_CHECK(h);
        }
    }
    double hs[kMaxD + 4];
    GP_CUDA(h, cudaMemcpyAsync(hs, scal, (size_t)(grad ? D + 3 : 1) * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    const double value = -0.5 * hs[0] - 0.5 * h->logdet[a] - 0.5 * n * 1.8378770664093453;   // log(2 pi)
    GP_CUDA(h, cudaMemcpyAsync(ml, &value, sizeof(double), cudaMemcpyDefault, h->stream));
    double g[kMaxD + 2];
    if (grad) {
        for (int k = 0; k < D; ++k) g[k] = 0.25 * hs[1 + k];          // 1/2 tr(B dK/dlog lam_k), dK = Kf d_k^2 / (2 lam_k)
        g[D] = hs[1 + D];                                             // 1/2 tr(B 2 Kf)
        g[D + 1] = h->noise[a] * hs[2 + D];                           // 1/2 tr(B 2 sigma_n^2 I)
        GP_CUDA(h, cudaMemcpyAsync(grad, g, (size_t)(D + 2) * sizeof(double), cudaMemcpyDefault, h->stream));
    }
    GP_CUDA(h, cudaStreamSynchronize(h->stream));
    return GPMPC_OK;
}

// ---------------------------------------------------------------------------------------------
// Measured ceilings of the FP64 pipe (roofline denominators for the pair kernels)
// ---------------------------------------------------------------------------------------------
namespace gpmpc {
__global__ void __launch_bounds__(256) fma_peak_kernel(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
__global__ void __launch_bounds__(256) exp_peak_kernel(double *out, int iters, double seed)
{
    __shared__ double etab[16];
    if (threadIdx.x < 16) etab[threadIdx.x] = kExp2Tab[threadIdx.x];
    __syncthreads();
    double s0 = seed + 1e-3 * threadIdx.x, s1 = s0 + 0.1, s2 = s0 + 0.2, s3 = s0 + 0.3;
    double acc = 0.0;
    for (int i = 0; i < iters; ++i) {
        const double e0 = exp_neg(s0, etab), e1 = exp_neg(s1, etab), e2 = exp_neg(s2, etab), e3 = exp_neg(s3, etab);
        acc += (e0 + e1) + (e2 + e3);
        s0 = e0 + 0.5; s1 = e1 + 0.6; s2 = e2 + 0.7; s3 = e3 + 0.8;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
}  // namespace gpmpc

extern "C" int gpmpc_measure_fp64_peak(gpmpc_handle h, double *fma_tflops, double *exp_gops)
{
    if (!h) return GPMPC_ERR_INVALID;
    GP_CUDA(h, cudaSetDevice(h->device));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    const int blocks = sms * 8, threads = 256;
    GP_CUDA(h, h->gbuf.reserve((size_t)blocks * threads * sizeof(double)));
    float ms = 0.f;
    const int it_f = 1 << 15, it_e = 1 << 12;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(h->ev0, h->stream);
        fma_peak_kernel<<<blocks, threads, 0, h->stream>>>(h->gbuf.as<double>(), it_f, 1.0);
        cudaEventRecord(h->ev1, h->stream);
        GP_LAUNCH_CHECK(h);
        GP_CUDA(h, cudaEventSynchronize(h->ev1));
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    }
    if (fma_tflops) *fma_tflops = 2.0 * 8.0 * it_f * (double)blocks * threads / (ms * 1e-3) / 1e12;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(h->ev0, h->stream);
        exp_peak_kernel<<<blocks, threads, 0, h->stream>>>(h->gbuf.as<double>(), it_e, 0.25);
        cudaEventRecord(h->ev1, h->stream);
        GP_LAUNCH_CHECK(h);
        GP_CUDA(h, cudaEventSynchronize(h->ev1));
        cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    }
    if (exp_gops) *exp_gops = 4.0 * it_e * (double)blocks * threads / (ms * 1e-3) / 1e9;
    return GPMPC_OK;
}
