// Pieces of one moment-matching step shared by rollout.cu and the fused few-rollouts kernel
// (mm_step_single.cuh): the step dimensions and the tail that turns the reduced sums of one (rollout, output)
// into mean_t, var_t and the tape entry (closed-form partial derivatives).
#pragma once
#include "common.cuh"

namespace gpmpc {

struct StepDims { int B, Bpad, D, E, m, G, n, ld; };

//   tape[((t-1)*E + a) * (2+4D) + e][Bpad]:  e = 0 mean, 1 var, 2.. dm/du, dm/ds, dv/du, dv/ds
__device__ __forceinline__ void finalize_math(const StepDims &d, int t, int a, int b, const double *accN,
                                              const double *accM, const double *__restrict__ us,
                                              const double *__restrict__ hyp, double *__restrict__ mu,
                                              double *__restrict__ var, double *__restrict__ tape, int want_grad)
{
    const int D = d.D;
    const double *lam = hyp + (size_t)a * D;
    const double sf = hyp[(size_t)d.E * D + a];
    double detm = 1.0, detv = 1.0;
    double s[kMaxD];
    for (int k = 0; k < D; ++k) {
        s[k] = us[(size_t)(D + k) * d.Bpad + b];
        detm *= 1.0 + s[k] / lam[k];            // |Lam^-1 S + I|      uncertainty_prop.py:335
        detv *= 1.0 + 2.0 * s[k] / lam[k];      // |2 Lam^-1 S + I|    uncertainty_prop.py:377
    }
    const double sf2 = sf * sf;
    const double cmf = sf2 / sqrt(detm);
    const double cvf = sf2 * sf2 / sqrt(detv);
    const double M0 = cmf * accM[0];
    const double N0 = cvf * accN[0];
    const double mean = M0;
    const double v = sf2 - N0 - mean * mean;     // latent variance, uncertainty_prop.py:399
    mu[((size_t)t * d.E + a) * d.Bpad + b] = mean;
    var[((size_t)t * d.E + a) * d.Bpad + b] = v;
    if (!want_grad) return;
    const int NT = 2 + 4 * D;
    double *tp = tape + (((size_t)(t - 1) * d.E + a) * NT) * d.Bpad + b;
    tp[0] = mean;
    tp[(size_t)d.Bpad] = v;
    for (int k = 0; k < D; ++k) {
        const double ak = 1.0 / (0.5 * lam[k] + s[k]);
        const double bk = 1.0 / (s[k] + lam[k]);
        const double c = sqrt(0.125 * ak), cm = sqrt(0.5 * bk);
        const double M1 = cmf * accM[1 + k] / cm;               // sum beta l v_k
        const double M2 = cmf * accM[1 + D + k] / (cm * cm);    // sum beta l v_k^2
        const double N1 = cvf * accN[1 + k] / c;                // sum w (v_ik + v_jk)
        const double N2 = cvf * accN[1 + D + k] / (c * c);      // sum w (v_ik + v_jk)^2
        const double dmu = -bk * M1;
        const double dms = 0.5 * bk * bk * M2 - 0.5 * bk * M0;
        const double dTu = -0.5 * ak * N1;
        const double dTs = 0.125 * ak * ak * N2 - N0 / (lam[k] + 2.0 * s[k]);
        tp[(size_t)(2 + k) * d.Bpad] = dmu;
        tp[(size_t)(2 + D + k) * d.Bpad] = dms;
        tp[(size_t)(2 + 2 * D + k) * d.Bpad] = -dTu - 2.0 * mean * dmu;
        tp[(size_t)(2 + 3 * D + k) * d.Bpad] = -dTs - 2.0 * mean * dms;
    }
}

// The same tail spread over the lanes of ONE warp (lane k < D owns input dimension k; all 32 lanes must call):
// the serial version is a chain of ~12 FP64 divisions / square roots per dimension, which is what a single
// thread of the fused few-rollouts kernel would otherwise spend ~10 us on.  Same operations in the same order
// per element, hence bit-identical to finalize_math.
__device__ __forceinline__ void finalize_math_lanes(const StepDims &d, int t, int a, int b, int lane, const double *accN,
                                                    const double *accM, const double *__restrict__ us,
                                                    const double *__restrict__ hyp, double *__restrict__ mu,
                                                    double *__restrict__ var, double *__restrict__ tape, int want_grad,
                                                    const double *s_local = nullptr)
{
    // s_local: the D input variances of this rollout in shared memory; the persistent rollout kernel passes them because
    // `us` changes while it runs (a __restrict__ const global load may take the non-coherent path)
    const int D = d.D;
    const int k = lane < D ? lane : 0;               // idle lanes mirror lane 0
    const double lamk = hyp[(size_t)a * D + k];
    const double sf = hyp[(size_t)d.E * D + a];
    const double sk = s_local ? s_local[k] : us[(size_t)(D + k) * d.Bpad + b];
    const double fm = 1.0 + sk / lamk;               // |Lam^-1 S + I| factor     uncertainty_prop.py:335
    const double fv = 1.0 + 2.0 * sk / lamk;         // |2 Lam^-1 S + I| factor   uncertainty_prop.py:377
    double detm = 1.0, detv = 1.0;
    for (int j = 0; j < D; ++j) {
        detm *= __shfl_sync(0xffffffffu, fm, j);
        detv *= __shfl_sync(0xffffffffu, fv, j);
    }
    const double sf2 = sf * sf;
    const double cmf = sf2 / sqrt(detm);
    const double cvf = sf2 * sf2 / sqrt(detv);
    const double M0 = cmf * accM[0];
    const double N0 = cvf * accN[0];
    const double mean = M0;
    const double v = sf2 - N0 - mean * mean;         // latent variance, uncertainty_prop.py:399
    const int NT = 2 + 4 * D;
    double *tp = tape + (((size_t)(t - 1) * d.E + a) * NT) * d.Bpad + b;
    if (lane == 0) {
        mu[((size_t)t * d.E + a) * d.Bpad + b] = mean;
        var[((size_t)t * d.E + a) * d.Bpad + b] = v;
        if (want_grad) { tp[0] = mean; tp[(size_t)d.Bpad] = v; }
    }
    if (!want_grad || lane >= D) return;
    const double ak = 1.0 / (0.5 * lamk + sk);
    const double bk = 1.0 / (sk + lamk);
    const double c = sqrt(0.125 * ak), cm = sqrt(0.5 * bk);
    const double M1 = cmf * accM[1 + k] / cm;               // sum beta l v_k
    const double M2 = cmf * accM[1 + D + k] / (cm * cm);    // sum beta l v_k^2
    const double N1 = cvf * accN[1 + k] / c;                // sum w (v_ik + v_jk)
    const double N2 = cvf * accN[1 + D + k] / (c * c);      // sum w (v_ik + v_jk)^2
    const double dmu = -bk * M1;
    const double dms = 0.5 * bk * bk * M2 - 0.5 * bk * M0;
    const double dTu = -0.5 * ak * N1;
    const double dTs = 0.125 * ak * ak * N2 - N0 / (lamk + 2.0 * sk);
    tp[(size_t)(2 + k) * d.Bpad] = dmu;
    tp[(size_t)(2 + D + k) * d.Bpad] = dms;
    tp[(size_t)(2 + 2 * D + k) * d.Bpad] = -dTu - 2.0 * mean * dmu;
    tp[(size_t)(2 + 3 * D + k) * d.Bpad] = -dTs - 2.0 * mean * dms;
}

// Scaled constants of one input dimension for the pair / mean kernels (prep_step, src/dynamics.py:154-163):
//   c = sqrt(a/8), a = 1/(lam/2 + s)  (uncertainty_prop.py:376);  cm = sqrt(b/2), b = 1/(s + lam)  (:331)
__device__ __forceinline__ void step_constants(double u, double s, double lam, double &c, double &cu, double &cm,
                                               double &cmu)
{
    const double a = 1.0 / (0.5 * lam + s);
    const double bb = 1.0 / (s + lam);
    c = sqrt(0.125 * a);
    cm = sqrt(0.5 * bb);
    cu = c * u;
    cmu = cm * u;
}

}  // namespace gpmpc
