/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code.
 *
 * Plain-C restatement of the reference's numba kernels
 *   /root/reference/src/fluids2d/weno.py
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may load this library.  The product path
 * (fluids2d_b200/) never does.
 *
 * Build:  gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC
 * (no FMA contraction: numba/LLVM does not contract a*b+c either, so every
 * operation below rounds exactly as the reference's does; the expression
 * trees follow the Python operator precedence of the cited lines).
 *
 * Parity pin: tests/test_oracle_vs_golden.py checks every function here
 * bit-for-bit against vectors produced by the live reference
 * (tests/golden/make_golden.py).
 */
#include <math.h>
#include <stdint.h>

typedef double f8;

/* ---- linear reconstructions: weno.py:21-43 ---------------------------- */
static inline f8 ce2(f8 U, f8 qm, f8 qp) { (void)U; return (qm + qp) * 0.5; }
static inline f8 ce4(f8 U, f8 qmm, f8 qm, f8 qp, f8 qpp) {
    (void)U;
    return ((-qmm + 7 * (qm + qp)) - qpp) / 12;
}
static inline f8 ce6(f8 U, f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp, f8 qppp) {
    (void)U;
    return (((qmmm - 8 * (qmm + qpp)) + 37 * (qm + qp)) + qppp) / 60;
}
static inline f8 up3(f8 qm, f8 q0, f8 qp) { return ((5 * q0 + 2 * qp) - qm) / 6; }
static inline f8 up5(f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp) {
    return ((((2 * qmmm - 13 * qmm) + 47 * qm) + 27 * qp) - 3 * qpp) / 60;
}

/* ---- WENO-Z, 3 points: weno.py:75-100 ---------------------------------- */
static inline f8 weno3z(f8 qm, f8 q0, f8 qp) {
    const f8 eps = 1e-14;
    f8 qi1 = -1. / 2. * qm + 3. / 2. * q0;
    f8 qi2 = 1. / 2. * (q0 + qp);
    f8 d1 = q0 - qm, d2 = qp - q0;
    f8 beta1 = d1 * d1;
    f8 beta2 = d2 * d2;
    f8 tau = fabs(beta2 - beta1);
    const f8 g1 = 1. / 3., g2 = 2. / 3.;
    f8 w1 = g1 * (1. + tau / (beta1 + eps));
    f8 w2 = g2 * (1. + tau / (beta2 + eps));
    return (w1 * qi1 + w2 * qi2) / (w1 + w2);
}

/* ---- centred WENO-Z, 3 points: weno.py:103-134 ------------------------- */
static inline f8 cweno3z(f8 U, f8 qmm, f8 qm, f8 qp, f8 qpp) {
    const f8 eps = 1e-14;
    f8 qi1 = -1. / 2. * qmm + 3. / 2. * qm;
    f8 qi2 = 1. / 2. * (qm + qp);
    f8 qi3 = -1. / 2. * qpp + 3. / 2. * qp;
    f8 beta1, beta2, t;
    if (U > 0) {
        t = qm - qmm; beta1 = t * t;
        t = qp - qm;  beta2 = t * t;
    } else {
        t = qp - qpp; beta1 = t * t;
        t = qp - qm;  beta2 = t * t;
    }
    f8 tau = fabs(beta2 - beta1);
    const f8 g1 = 1. / 3., g2 = 2. / 3.;
    f8 w1 = g1 * (1. + tau / (beta1 + eps));
    f8 w2 = g2 * (1. + tau / (beta2 + eps));
    return (w1 * (qi1 + qi3) * 0.5 + w2 * qi2) / (w1 + w2);
}

/* ---- WENO-Z, 5 points: weno.py:167-197 --------------------------------- */
static inline f8 sq(f8 x) { return x * x; }

static inline f8 weno5z(f8 qmm, f8 qm, f8 q0, f8 qp, f8 qpp) {
    const f8 eps = 1e-16;
    f8 qi1 = 1. / 3. * qmm - 7. / 6. * qm + 11. / 6. * q0;
    f8 qi2 = -1. / 6. * qm + 5. / 6. * q0 + 1. / 3. * qp;
    f8 qi3 = 1. / 3. * q0 + 5. / 6. * qp - 1. / 6. * qpp;
    const f8 k1 = 13. / 12., k2 = 0.25;
    f8 beta1 = k1 * sq(qmm - 2 * qm + q0) + k2 * sq(qmm - 4 * qm + 3 * q0);
    f8 beta2 = k1 * sq(qm - 2 * q0 + qp) + k2 * sq(qm - qp);
    f8 beta3 = k1 * sq(q0 - 2 * qp + qpp) + k2 * sq(3 * q0 - 4 * qp + qpp);
    f8 tau5 = fabs(beta1 - beta3);
    const f8 g1 = 0.1, g2 = 0.6, g3 = 0.3;
    f8 w1 = g1 * (1 + tau5 / (beta1 + eps));
    f8 w2 = g2 * (1 + tau5 / (beta2 + eps));
    f8 w3 = g3 * (1 + tau5 / (beta3 + eps));
    return (w1 * qi1 + w2 * qi2 + w3 * qi3) / (w1 + w2 + w3);
}

/* ---- centred WENO-Z, 5 points (the _v0 variant in the method table):
 *      weno.py:255-300.  Both branches of the `if U > 0` at :272-279 are
 *      textually identical in the reference, so U has no effect. ---------- */
static inline f8 cweno5z_v0(f8 U, f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp, f8 qppp) {
    (void)U;
    const f8 eps = 1e-16;
    f8 qi1 = 1. / 3. * qmmm - 7. / 6. * qmm + 11. / 6. * qm;
    f8 qi2 = -1. / 6. * qmm + 5. / 6. * qm + 1. / 3. * qp;
    f8 qi3 = 1. / 3. * qm + 5. / 6. * qp - 1. / 6. * qpp;
    f8 qi4 = 1. / 3. * qp + 5. / 6. * qm - 1. / 6. * qmm;
    f8 qi5 = -1. / 6. * qpp + 5. / 6. * qp + 1. / 3. * qm;
    f8 qi6 = 1. / 3. * qppp - 7. / 6. * qpp + 11. / 6. * qp;
    const f8 k1 = 13. / 12., k2 = 0.25;
    f8 beta1 = k1 * sq(qmmm - 2 * qmm + qm) + k2 * sq(qmmm - 4 * qmm + 3 * qm);
    f8 beta2 = k1 * sq(qmm - 2 * qm + qp) + k2 * sq(qmm - qp);
    f8 beta3 = k1 * sq(qm - 2 * qp + qpp) + k2 * sq(3 * qm - 4 * qp + qpp);
    f8 tau5 = fabs(beta1 - beta3);
    const f8 g1 = 0.1, g2 = 0.6, g3 = 0.3;
    f8 w1 = g1 * (1 + tau5 / (beta1 + eps));
    f8 w2 = g2 * (1 + tau5 / (beta2 + eps));
    f8 w3 = g3 * (1 + tau5 / (beta3 + eps));
    return (w1 * (qi1 + qi6) + w2 * (qi2 + qi5) + w3 * (qi3 + qi4)) / (2 * (w1 + w2 + w3));
}

/* ---- upwind selectors: weno.py:303-335 --------------------------------- */
static inline f8 flx1(f8 U, f8 qm, f8 qp) { return U > 0 ? qm : qp; }
static inline f8 flx3(f8 U, f8 qmm, f8 qm, f8 qp, f8 qpp) {
    return U > 0 ? weno3z(qmm, qm, qp) : weno3z(qpp, qp, qm);
}
static inline f8 cflx3(f8 U, f8 qmm, f8 qm, f8 qp, f8 qpp) { return cweno3z(U, qmm, qm, qp, qpp); }
static inline f8 flxup3(f8 U, f8 qmm, f8 qm, f8 qp, f8 qpp) {
    return U > 0 ? up3(qmm, qm, qp) : up3(qpp, qp, qm);
}
static inline f8 flx5(f8 U, f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp, f8 qppp) {
    return U > 0 ? weno5z(qmmm, qmm, qm, qp, qpp) : weno5z(qppp, qpp, qp, qm, qmm);
}
static inline f8 cflx5(f8 U, f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp, f8 qppp) {
    return cweno5z_v0(U, qmmm, qmm, qm, qp, qpp, qppp);
}
static inline f8 flxup5(f8 U, f8 qmmm, f8 qmm, f8 qm, f8 qp, f8 qpp, f8 qppp) {
    return U > 0 ? up5(qmmm, qmm, qm, qp, qpp) : up5(qppp, qpp, qp, qm, qmm);
}

/* method table: weno.py:338-343.  0 weno, 1 upwind, 2 centered, 3 cweno */
static inline f8 F1(int m, f8 U, f8 a, f8 b) {
    return (m == 0 || m == 1) ? flx1(U, a, b) : ce2(U, a, b);
}
static inline f8 F3(int m, f8 U, f8 a, f8 b, f8 c, f8 d) {
    switch (m) {
    case 0: return flx3(U, a, b, c, d);
    case 1: return flxup3(U, a, b, c, d);
    case 2: return ce4(U, a, b, c, d);
    default: return cflx3(U, a, b, c, d);
    }
}
static inline f8 F5(int m, f8 U, f8 a, f8 b, f8 c, f8 d, f8 e, f8 f) {
    switch (m) {
    case 0: return flx5(U, a, b, c, d, e, f);
    case 1: return flxup5(U, a, b, c, d, e, f);
    case 2: return ce6(U, a, b, c, d, e, f);
    default: return cflx5(U, a, b, c, d, e, f);
    }
}

/* numba wraps negative indices like Python; out-of-range high reads are
 * undefined in the reference and never reached for mesh-derived orders
 * (meshes.py:146-186 guards both ends).  We clamp them to index 0 so the
 * oracle never faults. */
static inline int64_t wrapi(int64_t k, int64_t n) {
    if (k < 0) k += n;
    if (k < 0 || k >= n) k = 0;
    return k;
}
#define AT(a, k) (a)[wrapi((k), n)]

/* scalar probes, used by the per-function parity tests */
f8 oracle_weno3z(f8 a, f8 b, f8 c) { return weno3z(a, b, c); }
f8 oracle_weno5z(f8 a, f8 b, f8 c, f8 d, f8 e) { return weno5z(a, b, c, d, e); }
f8 oracle_f1(int m, f8 U, f8 a, f8 b) { return F1(m, U, a, b); }
f8 oracle_f3(int m, f8 U, f8 a, f8 b, f8 c, f8 d) { return F3(m, U, a, b, c, d); }
f8 oracle_f5(int m, f8 U, f8 a, f8 b, f8 c, f8 d, f8 e, f8 f) { return F5(m, U, a, b, c, d, e, f); }

/* weno.py:346-364 */
void oracle_compflux(f8 *flx, const f8 *U, const f8 *q, const int8_t *o,
                     int64_t n, int64_t s, int method, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
        if (o[i] > 4)
            flx[i] = F5(method, U[i], AT(q, i - 3 * s), AT(q, i - 2 * s), AT(q, i - s), q[i],
                        AT(q, i + s), AT(q, i + 2 * s)) * U[i];
        else if (o[i] > 2)
            flx[i] = F3(method, U[i], AT(q, i - 2 * s), AT(q, i - s), q[i], AT(q, i + s)) * U[i];
        else if (o[i] > 0)
            flx[i] = F1(method, U[i], AT(q, i - s), q[i]) * U[i];
        else
            flx[i] = 0;
    }
}

/* weno.py:367-385 */
void oracle_vortexforce(f8 *du, const f8 *V, const f8 *q, const int8_t *o,
                        int64_t n, int64_t s, int64_t s2, int sign, int method,
                        int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
        if (o[i] > 0) {
            f8 Vm = 0.25 * (((V[i] + AT(V, i + s)) + AT(V, i - s2)) + AT(V, i + s - s2));
            f8 r;
            if (o[i] > 4)
                r = F5(method, Vm, AT(q, i - 2 * s), AT(q, i - s), q[i], AT(q, i + s),
                       AT(q, i + 2 * s), AT(q, i + 3 * s));
            else if (o[i] > 2)
                r = F3(method, Vm, AT(q, i - s), q[i], AT(q, i + s), AT(q, i + 2 * s));
            else
                r = F1(method, Vm, q[i], AT(q, i + s));
            du[i] = (sign * r) * Vm;
        } else
            du[i] = 0;
    }
}

/* weno.py:388-405 */
void oracle_innerproduct(f8 *ke, const f8 *U, const f8 *q, const int8_t *o,
                         int64_t n, int64_t s, int method, int64_t i0, int64_t i1) {
    for (int64_t i = i0; i < i1; i++) {
        if (o[i] > 0) {
            f8 Um = 0.5 * (U[i] + AT(U, i + s));
            f8 r;
            if (o[i] > 4)
                r = F5(method, Um, AT(q, i - 2 * s), AT(q, i - s), q[i], AT(q, i + s),
                       AT(q, i + 2 * s), AT(q, i + 3 * s));
            else if (o[i] > 2)
                r = F3(method, Um, AT(q, i - s), q[i], AT(q, i + s), AT(q, i + 2 * s));
            else
                r = F1(method, Um, q[i], AT(q, i + s));
            ke[i] += r * Um;
        }
    }
}

/* meshes.py:146-186 -- stencil order from a 0/1 mask in flat index space.
 * `shift` carries the sign the reference passes. */
void oracle_set_order(const int8_t *m, int64_t n, int64_t shift, int8_t *o, int maxorder) {
    for (int64_t i = 0; i < n; i++) {
        int s2, s4, s6;
        if (shift > 0) {
            s2 = (i - shift >= 0) ? m[i - shift] + m[i] : 0;
            s4 = ((i - shift * 2 >= 0) && (i + shift < n)) ? m[i - shift * 2] + m[i + shift] + s2 : 0;
            s6 = ((i - shift * 3 >= 0) && (i + 2 * shift < n)) ? m[i - shift * 3] + m[i + shift * 2] + s4 : 0;
        } else {
            s2 = (i - shift < n) ? m[i - shift] + m[i] : 0;
            s4 = ((i - shift * 2 < n) && (i + shift >= 0)) ? m[i - shift * 2] + m[i + shift] + s2 : 0;
            s6 = ((i - shift * 3 < n) && (i + 2 * shift >= 0)) ? m[i - shift * 3] + m[i + shift * 2] + s4 : 0;
        }
        int ord = (s6 == 6) ? 6 : ((s4 == 4) ? 4 : ((s2 == 2) ? 2 : 0));
        o[i] = (int8_t)(ord < maxorder ? ord : maxorder);
    }
}
