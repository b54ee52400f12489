// Device-side point reconstructions for the three stencil kernels.
//
// Behaviour follows the reference's numba functions (weno.py:21-43 linear,
// :75-100 weno3z, :103-134 cweno3z, :167-197 weno5z, :255-300 cweno5z_v0,
// selectors :303-335, method table :338-343).
//
// Two builds:
//   F2D_EXACT   (libf2d_exact.so, compiled with -fmad=false): every operation
//               in the reference's order -> bit-identical to numba/LLVM.
//   default     (libf2d.so): FMA contraction on, and the WENO-Z weights are
//               evaluated with ONE division instead of 4 (3-point: 1 instead
//               of 3): w_k = g_k (1 + tau/a_k) is scaled by prod(a_m), which
//               leaves sum(w_k q_k)/sum(w_k) unchanged algebraically.  fp64
//               division is ~30 instructions on sm_100, so this is what keeps
//               the advection kernel on the HBM side of the ridge.
#pragma once
#include <cuda_runtime.h>

namespace f2d {

enum Method { WENO = 0, UPWIND = 1, CENTERED = 2, CWENO = 3 };

__device__ __forceinline__ double sq(double x) { return x * x; }

// a / b.  Exact build: IEEE division (~25 instructions on sm_100).  Production
// build: hardware reciprocal seed (20 bits) + two Newton steps + one residual
// correction of the quotient, 8 instructions, result within 1 ulp.  Only used
// where b is a sum of positive WENO weights (never 0, inf or subnormal).
__device__ __forceinline__ double wdiv(double a, double b) {
#ifdef F2D_EXACT
    return a / b;
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    r = fma(fma(-b, r, 1.0), r, r);
    r = fma(fma(-b, r, 1.0), r, r);
    double q = a * r;
    return fma(fma(-b, q, a), r, q);
#endif
}

__device__ __forceinline__ double ce2(double qm, double qp) { return (qm + qp) * 0.5; }
__device__ __forceinline__ double ce4(double qmm, double qm, double qp, double qpp) {
    return ((-qmm + 7 * (qm + qp)) - qpp) / 12;
}
__device__ __forceinline__ double ce6(double a, double b, double c, double d, double e, double f) {
    return (((a - 8 * (b + e)) + 37 * (c + d)) + f) / 60;
}
__device__ __forceinline__ double up3(double qm, double q0, double qp) {
    return ((5 * q0 + 2 * qp) - qm) / 6;
}
__device__ __forceinline__ double up5(double a, double b, double c, double d, double e) {
    return ((((2 * a - 13 * b) + 47 * c) + 27 * d) - 3 * e) / 60;
}

__device__ __forceinline__ double weno3z(double qm, double q0, double qp) {
    const double eps = 1e-14;
    double qi1 = -1. / 2. * qm + 3. / 2. * q0;
    double qi2 = 1. / 2. * (q0 + qp);
    double beta1 = sq(q0 - qm);
    double beta2 = sq(qp - q0);
    double tau = fabs(beta2 - beta1);
    const double g1 = 1. / 3., g2 = 2. / 3.;
#ifdef F2D_EXACT
    double w1 = g1 * (1. + tau / (beta1 + eps));
    double w2 = g2 * (1. + tau / (beta2 + eps));
#else
    double a1 = beta1 + eps, a2 = beta2 + eps;
    double w1 = g1 * ((a1 + tau) * a2);
    double w2 = g2 * ((a2 + tau) * a1);
#endif
    return wdiv(w1 * qi1 + w2 * qi2, w1 + w2);
}

__device__ __forceinline__ double cweno3z(double U, double qmm, double qm, double qp, double qpp) {
    const double eps = 1e-14;
    double qi1 = -1. / 2. * qmm + 3. / 2. * qm;
    double qi2 = 1. / 2. * (qm + qp);
    double qi3 = -1. / 2. * qpp + 3. / 2. * qp;
    double beta1 = (U > 0) ? sq(qm - qmm) : sq(qp - qpp);
    double beta2 = sq(qp - qm);
    double tau = fabs(beta2 - beta1);
    const double g1 = 1. / 3., g2 = 2. / 3.;
#ifdef F2D_EXACT
    double w1 = g1 * (1. + tau / (beta1 + eps));
    double w2 = g2 * (1. + tau / (beta2 + eps));
#else
    double a1 = beta1 + eps, a2 = beta2 + eps;
    double w1 = g1 * ((a1 + tau) * a2);
    double w2 = g2 * ((a2 + tau) * a1);
#endif
    return wdiv(w1 * (qi1 + qi3) * 0.5 + w2 * qi2, w1 + w2);
}

struct W5 { double w1, w2, w3; };

__device__ __forceinline__ W5 weno5z_weights(double qmm, double qm, double q0, double qp, double qpp) {
    const double eps = 1e-16;
    const double k1 = 13. / 12., k2 = 0.25;
    double beta1 = k1 * sq(qmm - 2 * qm + q0) + k2 * sq(qmm - 4 * qm + 3 * q0);
    double beta2 = k1 * sq(qm - 2 * q0 + qp) + k2 * sq(qm - qp);
    double beta3 = k1 * sq(q0 - 2 * qp + qpp) + k2 * sq(3 * q0 - 4 * qp + qpp);
    double tau5 = fabs(beta1 - beta3);
    const double g1 = 0.1, g2 = 0.6, g3 = 0.3;
    W5 w;
#ifdef F2D_EXACT
    w.w1 = g1 * (1 + tau5 / (beta1 + eps));
    w.w2 = g2 * (1 + tau5 / (beta2 + eps));
    w.w3 = g3 * (1 + tau5 / (beta3 + eps));
#else
    double a1 = beta1 + eps, a2 = beta2 + eps, a3 = beta3 + eps;
    w.w1 = g1 * ((a1 + tau5) * (a2 * a3));
    w.w2 = g2 * ((a2 + tau5) * (a1 * a3));
    w.w3 = g3 * ((a3 + tau5) * (a1 * a2));
#endif
    return w;
}

__device__ __forceinline__ double weno5z(double qmm, double qm, double q0, double qp, double qpp) {
    double qi1 = 1. / 3. * qmm - 7. / 6. * qm + 11. / 6. * q0;
    double qi2 = -1. / 6. * qm + 5. / 6. * q0 + 1. / 3. * qp;
    double qi3 = 1. / 3. * q0 + 5. / 6. * qp - 1. / 6. * qpp;
    W5 w = weno5z_weights(qmm, qm, q0, qp, qpp);
    return wdiv(w.w1 * qi1 + w.w2 * qi2 + w.w3 * qi3, w.w1 + w.w2 + w.w3);
}

__device__ __forceinline__ double cweno5z_v0(double qmmm, double qmm, double qm, double qp,
                                             double qpp, double qppp) {
    double qi1 = 1. / 3. * qmmm - 7. / 6. * qmm + 11. / 6. * qm;
    double qi2 = -1. / 6. * qmm + 5. / 6. * qm + 1. / 3. * qp;
    double qi3 = 1. / 3. * qm + 5. / 6. * qp - 1. / 6. * qpp;
    double qi4 = 1. / 3. * qp + 5. / 6. * qm - 1. / 6. * qmm;
    double qi5 = -1. / 6. * qpp + 5. / 6. * qp + 1. / 3. * qm;
    double qi6 = 1. / 3. * qppp - 7. / 6. * qpp + 11. / 6. * qp;
    W5 w = weno5z_weights(qmmm, qmm, qm, qp, qpp);
    return wdiv(w.w1 * (qi1 + qi6) + w.w2 * (qi2 + qi5) + w.w3 * (qi3 + qi4), 2 * (w.w1 + w.w2 + w.w3));
}

// ---- method table (weno.py:338-343): f1 / f3 / f5 of each method ----------
template <int M>
__device__ __forceinline__ double f1(double U, double qm, double qp) {
    if (M == WENO || M == UPWIND) return U > 0 ? qm : qp;
    return ce2(qm, qp);
}

template <int M>
__device__ __forceinline__ double f3(double U, double qmm, double qm, double qp, double qpp) {
    if (M == WENO) return U > 0 ? weno3z(qmm, qm, qp) : weno3z(qpp, qp, qm);
    if (M == UPWIND) return U > 0 ? up3(qmm, qm, qp) : up3(qpp, qp, qm);
    if (M == CENTERED) return ce4(qmm, qm, qp, qpp);
    return cweno3z(U, qmm, qm, qp, qpp);
}

template <int M>
__device__ __forceinline__ double f5(double U, double qmmm, double qmm, double qm, double qp,
                                     double qpp, double qppp) {
    if (M == WENO) {
        // select the upwind 5-point window first, then one reconstruction
        bool pos = U > 0;
        double a = pos ? qmmm : qppp, b = pos ? qmm : qpp, c = pos ? qm : qp, d = pos ? qp : qm,
               e = pos ? qpp : qmm;
        return weno5z(a, b, c, d, e);
    }
    if (M == UPWIND) return U > 0 ? up5(qmmm, qmm, qm, qp, qpp) : up5(qppp, qpp, qp, qm, qmm);
    if (M == CENTERED) return ce6(qmmm, qmm, qm, qp, qpp, qppp);
    return cweno5z_v0(qmmm, qmm, qm, qp, qpp, qppp);
}

// Variable-order reconstruction at one half point: order o in {0,2,4,6},
// window w[0..5] = q[i-2s .. i+3s] in the vortexforce/innerproduct convention
// (weno.py:372-384, :393-403).
template <int M>
__device__ __forceinline__ double recon(int o, double U, double w0, double w1, double w2,
                                        double w3, double w4, double w5) {
    if (o > 4) return f5<M>(U, w0, w1, w2, w3, w4, w5);
    if (o > 2) return f3<M>(U, w1, w2, w3, w4);
    return f1<M>(U, w2, w3);
}

}  // namespace f2d
