// Model right-hand sides, Runge-Kutta updates and diagnostics on the device.
//
// Each kernel fuses a run of the reference's whole-array numpy passes
// (operators.py / equations.py / integrators.py) into one sweep; the
// x-periodic halo copy `mesh.fill` (meshes.py:135-143) is folded in by
// evaluating halo columns at their periodic image column.  Index arithmetic
// is the reference's flat-index arithmetic (weno.py:346-405), guarded by the
// same stencil-order arrays, so interior results depend on the same operands.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.cuh"
#include "reduce.cuh"
#include "weno.cuh"
#include "tma.cuh"

namespace f2d {

struct Grid {
    int n2, n1, nh, nx, xper;
    double idx2, idy2;   // 1/dx**2, 1/dy**2   (operators.py:59-64)
};

static Grid grid_of(const f2d_ctx *c) {
    Grid g;
    g.n2 = c->n2; g.n1 = c->n1; g.nh = c->nh; g.nx = c->cfg.nx; g.xper = c->cfg.xperiodic;
    g.idx2 = c->idx2; g.idy2 = c->idy2;
    return g;
}

// column at which a (possibly halo) column is evaluated: its periodic image
__device__ __forceinline__ int image_col(const Grid &g, int i) {
    if (g.xper) {
        if (i < g.nh) return i + g.nx;
        if (i >= g.n1 - g.nh) return i - g.nx;
    }
    return i;
}

#define THREAD_2D(g)                                         \
    int i_out = blockIdx.x * blockDim.x + threadIdx.x;       \
    int j = blockIdx.y * blockDim.y + threadIdx.y;           \
    if (i_out >= (g).n1 || j >= (g).n2) return;              \
    int i = image_col((g), i_out);                           \
    long k = (long)j * (g).n1 + i;                           \
    long k_out = (long)j * (g).n1 + i_out;                   \
    const long s1 = (g).n1;                                  \
    (void)k_out; (void)s1;

static dim3 blk2d() {
    static int bx = 0, by = 0;
    if (!bx) {
        bx = 64; by = 4;
        if (const char *e = getenv("F2D_BLK")) { int x, y; if (sscanf(e, "%dx%d", &x, &y) == 2 && x * y <= 256) { bx = x; by = y; } }
    }
    return dim3(bx, by);
}
static dim3 grd2d(const f2d_ctx *c) { dim3 b = blk2d(); return dim3((c->n1 + b.x - 1) / b.x, (c->n2 + b.y - 1) / b.y); }

enum { M_EULER = 0, M_BOUSS = 1, M_RSW = 2, M_QGRSW = 3, M_VADV = 4 };

// ---------------------------------------------------------------------------
// momentum tendency:  addvortexforce (operators.py:6-13, weno.py:367-385)
//   + addcoriolis (:32-39) + addgrad(ke) / addgrad(p) (:49-53) + addbuoyancy
//   (:152-153), then fill.          equations.py:11-15, 29-35, 121-127, 141-148
// ---------------------------------------------------------------------------
// NC > 0 fuses the Runge-Kutta update of the velocity (integrators.py:154-174):
//   ub = u + ((c0 ds_0) + c1 ds_1) + c2 ds_2, the last ds being this tendency;
//   ub is a second buffer because neighbouring threads still read u.
struct RkFuse {
    double c[3];
    const double *dx[2], *dy[2];   // earlier tendencies
    double *ubx, *uby;             // updated velocity
    int write_ds;                  // later stages need this tendency
};

template <int MV, int MODEL, int NC>
__global__ void __launch_bounds__(256, 8)
k_rhs_mom(Grid g, const double *__restrict__ ux, const double *__restrict__ uy,
          const double *__restrict__ omega, const double *__restrict__ ke,
          const double *__restrict__ p, const double *__restrict__ b,
          const int8_t *__restrict__ ovx, const int8_t *__restrict__ ovy,
          const int8_t *__restrict__ mskx, const int8_t *__restrict__ msky,
          double fcor, double halfdy, double *__restrict__ dux, double *__restrict__ duy, RkFuse rk) {
    THREAD_2D(g);
    double rx = 0, ry = 0;
    int oy = ovy[k];
    if (oy > 0) {   // du.x: V = U.y, s = yshift, s2 = xshift, sign +1
        double Vm = 0.25 * (((uy[k] * g.idy2 + uy[k + s1] * g.idy2) + uy[k - 1] * g.idy2) +
                            uy[k + s1 - 1] * g.idy2);
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = omega[k], w3 = omega[k + s1];
        if (oy > 2) { w1 = omega[k - s1]; w4 = omega[k + 2 * s1]; }
        if (oy > 4) { w0 = omega[k - 2 * s1]; w5 = omega[k + 3 * s1]; }
        rx = recon<MV>(oy, Vm, w0, w1, w2, w3, w4, w5) * Vm;
    }
    int ox = ovx[k];
    if (ox > 0) {   // du.y: V = U.x, s = xshift, s2 = yshift, sign -1
        double Vm = 0.25 * (((ux[k] * g.idx2 + ux[k + 1] * g.idx2) + ux[k - s1] * g.idx2) +
                            ux[k + 1 - s1] * g.idx2);
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = omega[k], w3 = omega[k + 1];
        if (ox > 2) { w1 = omega[k - 1]; w4 = omega[k + 2]; }
        if (ox > 4) { w0 = omega[k - 2]; w5 = omega[k + 3]; }
        ry = (-recon<MV>(ox, Vm, w0, w1, w2, w3, w4, w5)) * Vm;
    }
    if (MODEL == M_RSW || MODEL == M_QGRSW) {   // Coriolis, not masked in the reference
        if (j <= g.n2 - 2 && i >= 1 && i <= g.n1 - 2)
            rx += fcor * (((uy[k - 1] * g.idy2 + uy[k + s1 - 1] * g.idy2) + uy[k] * g.idy2) +
                          uy[k + s1] * g.idy2);
        if (j >= 1 && j <= g.n2 - 2 && i <= g.n1 - 2)
            ry -= fcor * (((ux[k - s1] * g.idx2 + ux[k - s1 + 1] * g.idx2) + ux[k] * g.idx2) +
                          ux[k + 1] * g.idx2);
    }
    if (MODEL != M_QGRSW) {
        if (i >= 1) rx -= (ke[k] - ke[k - 1]) * (double)mskx[k];
        if (j >= 1) ry -= (ke[k] - ke[k - s1]) * (double)msky[k];
    }
    if (MODEL == M_RSW) {
        if (i >= 1) rx -= (p[k] - p[k - 1]) * (double)mskx[k];
        if (j >= 1) ry -= (p[k] - p[k - s1]) * (double)msky[k];
    }
    if (MODEL == M_BOUSS) {
        if (j >= 1) ry += (halfdy * (b[k] + b[k - s1])) * (double)msky[k];
    }
    if (NC == 0 || rk.write_ds) {
        dux[k_out] = rx;
        duy[k_out] = ry;
    }
    if (NC > 0) {
        double ax, ay;
        if (NC == 1) { ax = rk.c[0] * rx; ay = rk.c[0] * ry; }
        else {
            ax = rk.c[0] * rk.dx[0][k_out]; ay = rk.c[0] * rk.dy[0][k_out];
            if (NC == 2) { ax = ax + rk.c[1] * rx; ay = ay + rk.c[1] * ry; }
            else {
                ax = ax + rk.c[1] * rk.dx[1][k_out]; ay = ay + rk.c[1] * rk.dy[1][k_out];
                ax = ax + rk.c[2] * rx; ay = ay + rk.c[2] * ry;
            }
        }
        rk.ubx[k_out] = ux[k_out] + ax;
        rk.uby[k_out] = uy[k_out] + ay;
    }
}

// ---------------------------------------------------------------------------
// divflux (operators.py:21-29): face fluxes (weno.py:346-364) ...
// ---------------------------------------------------------------------------
template <int MC>
__global__ void __launch_bounds__(256)
k_flux(Grid g, const double *__restrict__ ux, const double *__restrict__ uy,
       const double *__restrict__ q, const int8_t *__restrict__ ocx,
       const int8_t *__restrict__ ocy, double *__restrict__ fx, double *__restrict__ fy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.n1 || j >= g.n2) return;
    long k = (long)j * g.n1 + i;
    const long s1 = g.n1;
    double rx = 0, ry = 0;
    int ox = ocx[k];
    if (ox > 0) {
        double U = ux[k] * g.idx2;
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = q[k - 1], w3 = q[k];
        if (ox > 2) { w1 = q[k - 2]; w4 = q[k + 1]; }
        if (ox > 4) { w0 = q[k - 3]; w5 = q[k + 2]; }
        rx = recon<MC>(ox, U, w0, w1, w2, w3, w4, w5) * U;
    }
    int oy = ocy[k];
    if (oy > 0) {
        double U = uy[k] * g.idy2;
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = q[k - s1], w3 = q[k];
        if (oy > 2) { w1 = q[k - 2 * s1]; w4 = q[k + s1]; }
        if (oy > 4) { w0 = q[k - 3 * s1]; w5 = q[k + 2 * s1]; }
        ry = recon<MC>(oy, U, w0, w1, w2, w3, w4, w5) * U;
    }
    fx[k] = rx;
    fy[k] = ry;
}

// ... and their divergence `div` (operators.py:104-107), then fill.
__global__ void __launch_bounds__(256)
k_divflux(Grid g, const double *__restrict__ fx, const double *__restrict__ fy,
          const int8_t *__restrict__ msk, double *__restrict__ dq) {
    THREAD_2D(g);
    double d = 0;
    if (i <= g.n1 - 2) d = -(fx[k + 1] - fx[k]);
    if (j <= g.n2 - 2) d -= fy[k + s1] - fy[k];
    dq[k_out] = d * (double)msk[k];
}

// ((c0 x0) + c1 x1) + c2 x2 of addto_list (integrators.py:154-174) with the contraction into FMAs
// spelled out: left to the compiler, WHICH product of c0 x0 + c1 x1 is fused differs from kernel to
// kernel (measured: 1-ulp differences of b in 1 % of the points between k_addto and the fused
// transport kernel), and the scalar update must not depend on the kernel that performs it.
// libf2d_exact.so (-fmad=false) keeps the reference's separately rounded products.
template <int NC>
__device__ __forceinline__ double rk_sum(double c0, double c1, double c2, double x0, double x1, double x2) {
#ifdef F2D_EXACT
    double acc = c0 * x0;
    if (NC > 1) acc = acc + c1 * x1;
    if (NC > 2) acc = acc + c2 * x2;
    return acc;
#else
    double acc = __dmul_rn(c0, x0);
    if (NC > 1) acc = __fma_rn(c1, x1, acc);
    if (NC > 2) acc = __fma_rn(c2, x2, acc);
    return acc;
#endif
}
__device__ __forceinline__ double rk_add(double y, double acc) {
#ifdef F2D_EXACT
    return y + acc;
#else
    return __dadd_rn(y, acc);
#endif
}

// ... with the Runge-Kutta update of the advected scalar fused in (rsw stage, fused_stage_rsw):
// y += ((c0 ds_0) + c1 ds_1) + c2 ds_2 as addto_list does (integrators.py:154-174), the last
// ds being this tendency; the update is in place because the kernel only reads the fluxes.
struct RkScalar {
    double c[3];
    const double *d[2];     // earlier tendencies
    double *y;
    int write_ds;
};
template <int NC>
__global__ void __launch_bounds__(256)
k_divflux_upd(Grid g, const double *__restrict__ fx, const double *__restrict__ fy,
              const int8_t *__restrict__ msk, double *__restrict__ dq, RkScalar rk) {
    THREAD_2D(g);
    double d = 0;
    if (i <= g.n1 - 2) d = -(fx[k + 1] - fx[k]);
    if (j <= g.n2 - 2) d -= fy[k + s1] - fy[k];
    d = d * (double)msk[k];
    if (rk.write_ds) dq[k_out] = d;
    double acc;
    if (NC == 1) acc = rk_sum<1>(rk.c[0], 0, 0, d, 0, 0);
    else if (NC == 2) acc = rk_sum<2>(rk.c[0], rk.c[1], 0, rk.d[0][k_out], d, 0);
    else acc = rk_sum<3>(rk.c[0], rk.c[1], rk.c[2], rk.d[0][k_out], rk.d[1][k_out], d);
    rk.y[k_out] = rk_add(rk.y[k_out], acc);
}

// The tracer tendency (equations.py:217-222) is the same `div` WITHOUT the fill
// that follows every model-owned scalar: halo columns hold what `div` itself
// leaves there.  operators.py:105 does not assign the last column, so there the
// y-difference accumulates onto the previous content of dq (visible only when
// the halo mask is 1, i.e. xperiodic).
__global__ void __launch_bounds__(256)
k_divflux_nofill(Grid g, const double *__restrict__ fx, const double *__restrict__ fy,
                 const int8_t *__restrict__ msk, double *__restrict__ dq) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.n1 || j >= g.n2) return;
    long k = (long)j * g.n1 + i;
    int m = msk[k];
    double d = (i <= g.n1 - 2) ? -(fx[k + 1] - fx[k]) : (m ? dq[k] : 0.0);
    if (j <= g.n2 - 2) d -= fy[k + g.n1] - fy[k];
    dq[k] = d * (double)m;
}

// ---------------------------------------------------------------------------
// addto_list (integrators.py:154-174): y += ((0 + c0 x0) + c1 x1) + c2 x2
// ---------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256)
k_addto(long n, double *__restrict__ y, const double *__restrict__ x0,
        const double *__restrict__ x1, const double *__restrict__ x2, double c0, double c1,
        double c2) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double acc = rk_sum<NC>(c0, c1, c2, x0[k], NC > 1 ? x1[k] : 0.0, NC > 2 ? x2[k] : 0.0);
    y[k] = rk_add(y[k], acc);
}

// ---------------------------------------------------------------------------
// pressure_projection, first half (operators.py:114-116): delta = div(sharp(u))
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_div_u(Grid g, const double *__restrict__ ux, const double *__restrict__ uy,
        const int8_t *__restrict__ msk, double *__restrict__ delta) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.n1 || j >= g.n2) return;
    long k = (long)j * g.n1 + i;
    double d = 0;
    if (i <= g.n1 - 2) d = -(ux[k + 1] * g.idx2 - ux[k] * g.idx2);
    if (j <= g.n2 - 2) d -= uy[k + g.n1] * g.idy2 - uy[k] * g.idy2;
    delta[k] = d * (double)msk[k];
}

// ---------------------------------------------------------------------------
// diag:  [addgrad(p, u); fill(u)]  sharp; compute_vorticity (operators.py:67-77);
//        compute_kinetic_energy (:80-93, weno.py:388-405); [compute_pressure
//        (:110-111)]; fill(omega, ke).    equations.py:17-22, 37-43, 129-134,
//        150-155.   `uxin/uyin` hold the un-projected velocity when PROJECT.
//        MK: innerproduct method (4 = "classic"), -1 = no kinetic energy.
// ---------------------------------------------------------------------------
template <int MK, bool PROJECT, int MODEL>
__global__ void __launch_bounds__(256)
k_diag(Grid g, const double *__restrict__ uxin, const double *__restrict__ uyin,
       const double *__restrict__ p, const double *__restrict__ h, const double *__restrict__ hb,
       const int8_t *__restrict__ msk, const int8_t *__restrict__ mskx,
       const int8_t *__restrict__ msky, const int8_t *__restrict__ slip,
       const int8_t *__restrict__ okx, const int8_t *__restrict__ oky, double g_over_area,
       double *__restrict__ uxo, double *__restrict__ uyo, double *__restrict__ Ux,
       double *__restrict__ Uy, double *__restrict__ omega, double *__restrict__ ke,
       double *__restrict__ pout) {
    THREAD_2D(g);
    // velocity after the projection, at flat offset m from k (column i + di)
    auto UX = [&](long m, int di) -> double {
        double v = uxin[k + m];
        if (PROJECT && (i + di) >= 1) v -= (p[k + m] - p[k + m - 1]) * (double)mskx[k + m];
        return v;
    };
    auto UY = [&](long m, int dj) -> double {
        double v = uyin[k + m];
        if (PROJECT && (j + dj) >= 1) v -= (p[k + m] - p[k + m - s1]) * (double)msky[k + m];
        return v;
    };
    double ux0 = UX(0, 0), uy0 = UY(0, 0);
    if (PROJECT) { uxo[k_out] = ux0; uyo[k_out] = uy0; }
    if (Ux) {           // null: U is formed on demand (ensure_U)
        Ux[k_out] = ux0 * g.idx2;
        Uy[k_out] = uy0 * g.idy2;
    }
    double om = 0;
    if (j >= 1) om = -(ux0 - UX(-s1, 0));
    if (i >= 1) om += uy0 - UY(-1, 0);
    omega[k_out] = om * (double)slip[k];
    if (MK >= 0) {
        double e = 0;
        if (MK == F2D_METHOD_CLASSIC) {   // operators.py:86-89
            if (i <= g.n1 - 2) {
                double a = UX(1, 1);
                e = a * (a * g.idx2) + ux0 * (ux0 * g.idx2);
            }
            if (j <= g.n2 - 2) {
                double a = UY(s1, 1);
                e += a * (a * g.idy2) + uy0 * (uy0 * g.idy2);
            }
            e *= (double)msk[k] * 0.25;
        } else {
            constexpr int M = MK < 0 ? 0 : (MK > 3 ? 0 : MK);
            int ox = okx[k];
            if (ox > 0) {
                double w3 = UX(1, 1);
                double Um = 0.5 * (ux0 * g.idx2 + w3 * g.idx2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (ox > 2) { w1 = UX(-1, -1); w4 = UX(2, 2); }
                if (ox > 4) { w0 = UX(-2, -2); w5 = UX(3, 3); }
                e += recon<M>(ox, Um, w0, w1, ux0, w3, w4, w5) * Um;
            }
            int oy = oky[k];
            if (oy > 0) {
                double w3 = UY(s1, 1);
                double Um = 0.5 * (uy0 * g.idy2 + w3 * g.idy2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (oy > 2) { w1 = UY(-s1, -1); w4 = UY(2 * s1, 2); }
                if (oy > 4) { w0 = UY(-2 * s1, -2); w5 = UY(3 * s1, 3); }
                e += recon<M>(oy, Um, w0, w1, uy0, w3, w4, w5) * Um;
            }
            e *= (double)msk[k] * 0.5;
        }
        ke[k_out] = e;
    }
    if (MODEL == M_RSW) pout[k_out] = g_over_area * (h[k] + hb[k]);
}

// ---------------------------------------------------------------------------
// The same diagnostics for the projecting models, tiled through shared memory:
// the projected velocity of a (64+6) x (16+6) window is formed ONCE per point
// (1 velocity + 2 pressure loads) and the vorticity / kinetic-energy stencils
// read it from shared memory, instead of re-projecting ~12 neighbours per
// point from global memory.  Halo columns are computed in place like the
// reference does and then overwritten by the periodic fill (k_fill_many).
// ---------------------------------------------------------------------------
constexpr int DTX = 64, DTY = 16, DH = 3, DWX = DTX + 2 * DH, DWY = DTY + 2 * DH;

template <int MK>
__global__ void __launch_bounds__(256)
k_diag_tiled(Grid g, const double *__restrict__ uxin, const double *__restrict__ uyin,
             const double *__restrict__ p, const int8_t *__restrict__ msk,
             const int8_t *__restrict__ mskx, const int8_t *__restrict__ msky,
             const int8_t *__restrict__ slip, const int8_t *__restrict__ okx,
             const int8_t *__restrict__ oky, double *__restrict__ uxo, double *__restrict__ uyo,
             double *__restrict__ Ux, double *__restrict__ Uy, double *__restrict__ omega,
             double *__restrict__ ke) {
    __shared__ double sux[DWY][DWX], suy[DWY][DWX];
    const int i0 = blockIdx.x * DTX, j0 = blockIdx.y * DTY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const long s1 = g.n1;
    for (int t = tid; t < DWY * DWX; t += 256) {
        int a = t / DWX, b = t - a * DWX;
        int j = j0 - DH + a, i = i0 - DH + b;
        double vx = 0.0, vy = 0.0;
        if (j >= 0 && j < g.n2 && i >= 0 && i < g.n1) {
            long k = (long)j * s1 + i;
            double pc = p[k];
            vx = uxin[k];
            vy = uyin[k];
            if (i >= 1) vx -= (pc - p[k - 1]) * (double)mskx[k];      // addgrad, operators.py:49-53
            if (j >= 1) vy -= (pc - p[k - s1]) * (double)msky[k];
        }
        sux[a][b] = vx;
        suy[a][b] = vy;
    }
    __syncthreads();
    const int b = DH + threadIdx.x, i = i0 + threadIdx.x;
    if (i >= g.n1) return;
#pragma unroll
    for (int r = 0; r < DTY / 4; r++) {
        const int a = DH + threadIdx.y + 4 * r, j = j0 + threadIdx.y + 4 * r;
        if (j >= g.n2) break;
        const long k = (long)j * s1 + i;
        const double ux0 = sux[a][b], uy0 = suy[a][b];
        uxo[k] = ux0;
        uyo[k] = uy0;
        if (Ux) {       // null: U is formed on demand (ensure_U)
            Ux[k] = ux0 * g.idx2;
            Uy[k] = uy0 * g.idy2;
        }
        double om = 0.0;
        if (j >= 1) om = -(ux0 - sux[a - 1][b]);
        if (i >= 1) om += uy0 - suy[a][b - 1];
        omega[k] = om * (double)slip[k];
        double e = 0.0;
        if (MK == F2D_METHOD_CLASSIC) {
            if (i <= g.n1 - 2) { double w = sux[a][b + 1]; e = w * (w * g.idx2) + ux0 * (ux0 * g.idx2); }
            if (j <= g.n2 - 2) { double w = suy[a + 1][b]; e += w * (w * g.idy2) + uy0 * (uy0 * g.idy2); }
            e *= (double)msk[k] * 0.25;
        } else {
            constexpr int M = MK > 3 ? 0 : MK;
            int ox = okx[k];
            if (ox > 0) {
                double w3 = sux[a][b + 1];
                double Um = 0.5 * (ux0 * g.idx2 + w3 * g.idx2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (ox > 2) { w1 = sux[a][b - 1]; w4 = sux[a][b + 2]; }
                if (ox > 4) { w0 = sux[a][b - 2]; w5 = sux[a][b + 3]; }
                e += recon<M>(ox, Um, w0, w1, ux0, w3, w4, w5) * Um;
            }
            int oy = oky[k];
            if (oy > 0) {
                double w3 = suy[a + 1][b];
                double Um = 0.5 * (uy0 * g.idy2 + w3 * g.idy2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (oy > 2) { w1 = suy[a - 1][b]; w4 = suy[a + 2][b]; }
                if (oy > 4) { w0 = suy[a - 2][b]; w5 = suy[a + 3][b]; }
                e += recon<M>(oy, Um, w0, w1, uy0, w3, w4, w5) * Um;
            }
            e *= (double)msk[k] * 0.5;
        }
        ke[k] = e;
    }
}

// ---------------------------------------------------------------------------
// The fused tendency + Runge-Kutta update of the projecting models, tiled the
// same way: omega, u.x, u.y, ke of a (64+6) x (16+6) window are staged in shared
// memory once; the two variable-order reconstructions, the 4-point velocity
// averages and the kinetic-energy gradient read them from there.  Same
// arithmetic as k_rhs_mom, so the two agree bit for bit away from the periodic
// halo columns (which k_fill_many overwrites, as mesh.fill does).
// ---------------------------------------------------------------------------
template <int MV, int MODEL, int NC>
__global__ void __launch_bounds__(256)
k_stage_tiled(Grid g, const double *__restrict__ ux, const double *__restrict__ uy,
              const double *__restrict__ omega, const double *__restrict__ ke,
              const double *__restrict__ bb, const int8_t *__restrict__ ovx,
              const int8_t *__restrict__ ovy, const int8_t *__restrict__ mskx,
              const int8_t *__restrict__ msky, double halfdy, double *__restrict__ dux,
              double *__restrict__ duy, RkFuse rk) {
    // omega needs the full 3-point halo; u.x, u.y, ke only one point
    constexpr int H1 = 1, W1X = DTX + 2 * H1, W1Y = DTY + 2 * H1;
    __shared__ double som[DWY][DWX], sux[W1Y][W1X], suy[W1Y][W1X], ske[W1Y][W1X];
    const int i0 = blockIdx.x * DTX, j0 = blockIdx.y * DTY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const long s1 = g.n1;
    for (int t = tid; t < DWY * DWX; t += 256) {
        int a = t / DWX, b = t - a * DWX;
        int j = j0 - DH + a, i = i0 - DH + b;
        som[a][b] = (j >= 0 && j < g.n2 && i >= 0 && i < g.n1) ? omega[(long)j * s1 + i] : 0.0;
    }
    for (int t = tid; t < W1Y * W1X; t += 256) {
        int a = t / W1X, b = t - a * W1X;
        int j = j0 - H1 + a, i = i0 - H1 + b;
        double vx = 0.0, vy = 0.0, vk = 0.0;
        if (j >= 0 && j < g.n2 && i >= 0 && i < g.n1) {
            long k = (long)j * s1 + i;
            vx = ux[k]; vy = uy[k]; vk = ke[k];
        }
        sux[a][b] = vx; suy[a][b] = vy; ske[a][b] = vk;
    }
    __syncthreads();
    const int b = DH + threadIdx.x, b1 = H1 + threadIdx.x, i = i0 + threadIdx.x;
    if (i >= g.n1) return;
#pragma unroll
    for (int r = 0; r < DTY / 4; r++) {
        const int a = DH + threadIdx.y + 4 * r, a1 = H1 + threadIdx.y + 4 * r, j = j0 + threadIdx.y + 4 * r;
        if (j >= g.n2) break;
        const long k = (long)j * s1 + i;
        double rx = 0, ry = 0;
        int oy = ovy[k];
        if (oy > 0) {   // du.x: V = U.y, s = yshift, s2 = xshift, sign +1
            double Vm = 0.25 * (((suy[a1][b1] * g.idy2 + suy[a1 + 1][b1] * g.idy2) + suy[a1][b1 - 1] * g.idy2) +
                                suy[a1 + 1][b1 - 1] * g.idy2);
            double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = som[a][b], w3 = som[a + 1][b];
            if (oy > 2) { w1 = som[a - 1][b]; w4 = som[a + 2][b]; }
            if (oy > 4) { w0 = som[a - 2][b]; w5 = som[a + 3][b]; }
            rx = recon<MV>(oy, Vm, w0, w1, w2, w3, w4, w5) * Vm;
        }
        int ox = ovx[k];
        if (ox > 0) {   // du.y: V = U.x, s = xshift, s2 = yshift, sign -1
            double Vm = 0.25 * (((sux[a1][b1] * g.idx2 + sux[a1][b1 + 1] * g.idx2) + sux[a1 - 1][b1] * g.idx2) +
                                sux[a1 - 1][b1 + 1] * g.idx2);
            double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = som[a][b], w3 = som[a][b + 1];
            if (ox > 2) { w1 = som[a][b - 1]; w4 = som[a][b + 2]; }
            if (ox > 4) { w0 = som[a][b - 2]; w5 = som[a][b + 3]; }
            ry = (-recon<MV>(ox, Vm, w0, w1, w2, w3, w4, w5)) * Vm;
        }
        if (i >= 1) rx -= (ske[a1][b1] - ske[a1][b1 - 1]) * (double)mskx[k];
        if (j >= 1) ry -= (ske[a1][b1] - ske[a1 - 1][b1]) * (double)msky[k];
        if (MODEL == M_BOUSS) {
            if (j >= 1) ry += (halfdy * (bb[k] + bb[k - s1])) * (double)msky[k];
        }
        if (rk.write_ds) { dux[k] = rx; duy[k] = ry; }
        double ax, ay;
        if (NC == 1) { ax = rk.c[0] * rx; ay = rk.c[0] * ry; }
        else {
            ax = rk.c[0] * rk.dx[0][k]; ay = rk.c[0] * rk.dy[0][k];
            if (NC == 2) { ax = ax + rk.c[1] * rx; ay = ay + rk.c[1] * ry; }
            else {
                ax = ax + rk.c[1] * rk.dx[1][k]; ay = ay + rk.c[1] * rk.dy[1][k];
                ax = ax + rk.c[2] * rx; ay = ay + rk.c[2] * ry;
            }
        }
        rk.ubx[k] = sux[a1][b1] + ax;
        rk.uby[k] = suy[a1][b1] + ay;
    }
}

// ---------------------------------------------------------------------------
// The fused tendency + Runge-Kutta update of the projecting models, TMA-fed.
// One CTA = one 64 x 16 tile of output points.  One elected thread issues four
// (boussinesq: five) cp.async.bulk.tensor.2d box loads -- omega with its 3-point
// halo, u.x, u.y, ke (, b) with a 1-point halo -- that land in shared memory and
// complete on an mbarrier; meanwhile every thread fetches what has no reuse (the
// packed stencil-order / mask byte and the earlier tendencies of its own points) with
// plain loads.  Box coordinates outside the array are zero-filled by the TMA unit,
// so the window needs no bounds test.  A thread owns 4 consecutive rows of one
// column: the y-windows of its points overlap and every neighbour is a
// shared-memory load at a compile-time offset (no 64-bit address arithmetic per
// operand, which was a quarter of the instructions of the per-point kernel).
// 41 KB of shared memory per CTA, 5 CTAs per SM: the loads of the next tiles are
// in flight while this one computes.  Same expressions, in the same order, as
// k_rhs_mom: the two agree bit for bit (halo columns are filled afterwards by
// k_fill_many, as mesh.fill does).
// ---------------------------------------------------------------------------
// (the inner box coordinate must be a multiple of 16 bytes -- an even column for fp64,
//  scripts/tma_probe.cu -- so the boxes start one column further left than the halo needs)
constexpr int STX = 64, STY = 16, SHO = 3;
constexpr int SOX = 4, S1X = 2;                              // columns left of the tile in the omega / 1-halo boxes
constexpr int SOW = STX + 2 * SOX, SOH = STY + 2 * SHO;      // omega box
constexpr int S1W = STX + 2 * S1X, S1H = STY + 2;            // u.x, u.y, ke, b boxes
constexpr size_t pad128(size_t n) { return (n + 127) & ~size_t(127); }
constexpr size_t SOM_BYTES = pad128((size_t)SOW * SOH * 8), S1_BYTES = pad128((size_t)S1W * S1H * 8);
struct StageMaps { CUtensorMap om, ux, uy, ke, b; };

template <int MV, int MODEL, int NC>
__global__ void __launch_bounds__(256, 4)
k_stage_tma(const __grid_constant__ StageMaps M, Grid g, const uint8_t *__restrict__ smask, double halfdy, double fcor,
            double *__restrict__ dux, double *__restrict__ duy, RkFuse rk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    double(*som)[SOW] = reinterpret_cast<double(*)[SOW]>(smem_raw);
    double(*sux)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + SOM_BYTES);
    double(*suy)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + SOM_BYTES + S1_BYTES);
    double(*ske)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + SOM_BYTES + 2 * S1_BYTES);
    double(*sbb)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + SOM_BYTES + 3 * S1_BYTES);
    const int i0 = blockIdx.x * STX, j0 = blockIdx.y * STY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        constexpr unsigned bytes = (unsigned)(SOW * SOH * 8 + ((MODEL == M_BOUSS || MODEL == M_RSW) ? 4 : 3) * S1W * S1H * 8);
        mbar_expect_tx(&bar, bytes);
        tma_load_2d(&som[0][0], &M.om, i0 - SOX, j0 - SHO, &bar);
        tma_load_2d(&sux[0][0], &M.ux, i0 - S1X, j0 - 1, &bar);
        tma_load_2d(&suy[0][0], &M.uy, i0 - S1X, j0 - 1, &bar);
        tma_load_2d(&ske[0][0], &M.ke, i0 - S1X, j0 - 1, &bar);
        if (MODEL == M_BOUSS || MODEL == M_RSW) tma_load_2d(&sbb[0][0], &M.b, i0 - S1X, j0 - 1, &bar);     // b, or the rsw pressure
    }
    // ---- per-point operands without reuse: plain loads, in flight with the boxes
    constexpr int R = STY / 4;
    const int i = i0 + threadIdx.x;
    const int jb = j0 + R * threadIdx.y;
    const long s1 = g.n1;
    const bool col_ok = i < g.n1;
    int oxv[R], oyv[R];
    double mx[R], my[R], d0x[R], d0y[R], d1x[R], d1y[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = jb + r;
        const bool ok = col_ok && j < g.n2;
        const long k = ok ? (long)j * s1 + i : 0;
        const unsigned m = ok ? smask[k] : 0u;           // ov.x/2 | ov.y/2 << 2 | mskx << 4 | msky << 5
        oxv[r] = (m & 3u) << 1;
        oyv[r] = ((m >> 2) & 3u) << 1;
        mx[r] = (double)((m >> 4) & 1u);
        my[r] = (double)((m >> 5) & 1u);
        d0x[r] = d0y[r] = d1x[r] = d1y[r] = 0.0;
        if (NC >= 2 && ok) { d0x[r] = rk.dx[0][k]; d0y[r] = rk.dy[0][k]; }
        if (NC >= 3 && ok) { d1x[r] = rk.dx[1][k]; d1y[r] = rk.dy[1][k]; }
    }
    mbar_wait(&bar, 0);
    if (!col_ok) return;
    const int b = SOX + threadIdx.x, b1 = S1X + threadIdx.x;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = jb + r;
        if (j >= g.n2) break;
        const int a = SHO + R * threadIdx.y + r, a1 = 1 + R * threadIdx.y + r;
        const long k = (long)j * s1 + i;
        double rx = 0, ry = 0;
        const int oy = oyv[r];
        if (oy > 0) {   // du.x: V = U.y, s = yshift, s2 = xshift, sign +1
            double Vm = 0.25 * (((suy[a1][b1] * g.idy2 + suy[a1 + 1][b1] * g.idy2) + suy[a1][b1 - 1] * g.idy2) +
                                suy[a1 + 1][b1 - 1] * g.idy2);
            rx = recon<MV>(oy, Vm, som[a - 2][b], som[a - 1][b], som[a][b], som[a + 1][b], som[a + 2][b], som[a + 3][b]) * Vm;
        }
        const int ox = oxv[r];
        if (ox > 0) {   // du.y: V = U.x, s = xshift, s2 = yshift, sign -1
            double Vm = 0.25 * (((sux[a1][b1] * g.idx2 + sux[a1][b1 + 1] * g.idx2) + sux[a1 - 1][b1] * g.idx2) +
                                sux[a1 - 1][b1 + 1] * g.idx2);
            ry = (-recon<MV>(ox, Vm, som[a][b - 2], som[a][b - 1], som[a][b], som[a][b + 1], som[a][b + 2], som[a][b + 3])) * Vm;
        }
        if (MODEL == M_RSW) {   // Coriolis, not masked in the reference (operators.py:32-39)
            if (j <= g.n2 - 2 && i >= 1)
                rx += fcor * (((suy[a1][b1 - 1] * g.idy2 + suy[a1 + 1][b1 - 1] * g.idy2) + suy[a1][b1] * g.idy2) +
                              suy[a1 + 1][b1] * g.idy2);
            if (j >= 1 && j <= g.n2 - 2 && i <= g.n1 - 2)
                ry -= fcor * (((sux[a1 - 1][b1] * g.idx2 + sux[a1 - 1][b1 + 1] * g.idx2) + sux[a1][b1] * g.idx2) +
                              sux[a1][b1 + 1] * g.idx2);
        }
        if (i >= 1) rx -= (ske[a1][b1] - ske[a1][b1 - 1]) * mx[r];
        if (j >= 1) ry -= (ske[a1][b1] - ske[a1 - 1][b1]) * my[r];
        if (MODEL == M_RSW) {   // grad p, p = g (h + hb) / area from the last diag
            if (i >= 1) rx -= (sbb[a1][b1] - sbb[a1][b1 - 1]) * mx[r];
            if (j >= 1) ry -= (sbb[a1][b1] - sbb[a1 - 1][b1]) * my[r];
        }
        if (MODEL == M_BOUSS) {
            if (j >= 1) ry += (halfdy * (sbb[a1][b1] + sbb[a1 - 1][b1])) * my[r];
        }
        if (rk.write_ds) { dux[k] = rx; duy[k] = ry; }
        double ax, ay;
        if (NC == 1) { ax = rk.c[0] * rx; ay = rk.c[0] * ry; }
        else {
            ax = rk.c[0] * d0x[r]; ay = rk.c[0] * d0y[r];
            if (NC == 2) { ax = ax + rk.c[1] * rx; ay = ay + rk.c[1] * ry; }
            else {
                ax = ax + rk.c[1] * d1x[r]; ay = ay + rk.c[1] * d1y[r];
                ax = ax + rk.c[2] * rx; ay = ay + rk.c[2] * ry;
            }
        }
        rk.ubx[k] = sux[a1][b1] + ax;
        rk.uby[k] = suy[a1][b1] + ay;
    }
}

// ---------------------------------------------------------------------------
// Transport of a flux-form scalar (rsw thickness, Boussinesq buoyancy), one kernel per RK
// stage: face fluxes (k_flux, weno.py:346-364), their divergence (k_divflux,
// operators.py:104-107) and the Runge-Kutta update y* = y + sum c_i ds_i (k_addto /
// k_divflux_upd), TMA-fed.  The two flux arrays never reach HBM: a thread holds the west and
// south face fluxes of its 4 consecutive rows in registers, takes the east one from its
// neighbour lane (SHFL; the last lane of a warp gets it through shared memory from the next warp,
// or from the 16 threads that evaluate the tile's east edge) and the north one of its top row
// from the thread above through 2 KB of shared memory (the top quarter of the tile evaluates
// it): 67 warp-level reconstructions per tile of 64, where letting every warp's last lane
// evaluate its own east face costs 98.  y* goes to a separate array (neighbouring tiles still read y), the caller swaps the
// two pointers.  Boxes: q 72 x 22 (columns i0-4 .., rows j0-3 ..), u.x / u.y 68 x 18; the
// orders and the centre mask come packed in one byte (engine.cuh: tmask).  Algorithmic
// traffic 6 x 8 + 1 = 49 B/point against (8 x 8 + 3) + 5 x 8 = 107 of the three kernels.
// ---------------------------------------------------------------------------
constexpr int TQW = STX + 2 * SOX, TQH = STY + 2 * SHO;
constexpr size_t TQ_BYTES = pad128((size_t)TQW * TQH * 8);
struct TransportMaps { CUtensorMap q, ux, uy; };
struct RkScalarOut {
    double c[3];
    const double *d[2];     // earlier tendencies
    double *ynew;           // y + sum c_i ds_i
    int write_ds;
};

template <int MC>
__device__ __forceinline__ double face_flux_x(const double (*sq)[TQW], const double (*sux)[S1W], int a, int a1, int b,
                                              int b1, int ox, double idx2) {
    if (ox <= 0) return 0.0;
    const double U = sux[a1][b1] * idx2;
    return recon<MC>(ox, U, sq[a][b - 3], sq[a][b - 2], sq[a][b - 1], sq[a][b], sq[a][b + 1], sq[a][b + 2]) * U;
}
template <int MC>
__device__ __forceinline__ double face_flux_y(const double (*sq)[TQW], const double (*suy)[S1W], int a, int a1, int b,
                                              int b1, int oy, double idy2) {
    if (oy <= 0) return 0.0;
    const double U = suy[a1][b1] * idy2;
    return recon<MC>(oy, U, sq[a - 3][b], sq[a - 2][b], sq[a - 1][b], sq[a][b], sq[a + 1][b], sq[a + 2][b]) * U;
}

template <int MC, int NC>
__global__ void __launch_bounds__(256, 4)
k_transport_tma(const __grid_constant__ TransportMaps M, Grid g, const uint8_t *__restrict__ tmask,
                double *__restrict__ dq, RkScalarOut rk) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    double(*sq)[TQW] = reinterpret_cast<double(*)[TQW]>(smem_raw);
    double(*sux)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + TQ_BYTES);
    double(*suy)[S1W] = reinterpret_cast<double(*)[S1W]>(smem_raw + TQ_BYTES + S1_BYTES);
    double(*sfy)[STX] = reinterpret_cast<double(*)[STX]>(smem_raw + TQ_BYTES + 2 * S1_BYTES);
    double(*sfe)[2] = reinterpret_cast<double(*)[2]>(smem_raw + TQ_BYTES + 2 * S1_BYTES + 4 * STX * 8);
    const int i0 = blockIdx.x * STX, j0 = blockIdx.y * STY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (unsigned)(TQW * TQH * 8 + 2 * S1W * S1H * 8));
        tma_load_2d(&sq[0][0], &M.q, i0 - SOX, j0 - SHO, &bar);
        tma_load_2d(&sux[0][0], &M.ux, i0 - S1X, j0 - 1, &bar);
        tma_load_2d(&suy[0][0], &M.uy, i0 - S1X, j0 - 1, &bar);
    }
    // ---- per-point operands: plain loads, in flight with the boxes
    constexpr int R = STY / 4;
    const int i = i0 + threadIdx.x;
    const int jb = j0 + R * threadIdx.y;
    const long s1 = g.n1;
    const bool col_ok = i < g.n1;
    const bool east_lane = (threadIdx.x & 31) == 31;          // its east face belongs to the next warp / tile
    const bool edge_thread = tid < STY;                        // evaluates the east face of the tile's last column, row tid
    unsigned mk[R], mn = 0, me = 0;
    double d0[R], d1[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = jb + r;
        const bool ok = col_ok && j < g.n2;
        const long k = ok ? (long)j * s1 + i : 0;
        mk[r] = ok ? tmask[k] : 0u;                            // oc.x/2 | oc.y/2 << 2 | msk << 4
        d0[r] = d1[r] = 0.0;
        if (NC >= 2 && ok) d0[r] = rk.d[0][k];
        if (NC >= 3 && ok) d1[r] = rk.d[1][k];
    }
    if (threadIdx.y == 3 && col_ok && jb + R < g.n2) mn = tmask[(long)(jb + R) * s1 + i];
    if (edge_thread && j0 + tid < g.n2 && i0 + STX < g.n1) me = tmask[(long)(j0 + tid) * s1 + i0 + STX];
    mbar_wait(&bar, 0);
    const int b = SOX + threadIdx.x, b1 = S1X + threadIdx.x;
    double fxw[R], fys[R + 1];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int a = SHO + R * threadIdx.y + r, a1 = 1 + R * threadIdx.y + r;
        fxw[r] = face_flux_x<MC>(sq, sux, a, a1, b, b1, (int)((mk[r] & 3u) << 1), g.idx2);
        fys[r] = face_flux_y<MC>(sq, suy, a, a1, b, b1, (int)(((mk[r] >> 2) & 3u) << 1), g.idy2);
    }
    sfy[threadIdx.y][threadIdx.x] = fys[0];
    if (threadIdx.x == 32) {
#pragma unroll
        for (int r = 0; r < R; r++) sfe[R * threadIdx.y + r][0] = fxw[r];
    }
    if (edge_thread)
        sfe[tid][1] = face_flux_x<MC>(sq, sux, SHO + tid, 1 + tid, SOX + STX, S1X + STX, (int)((me & 3u) << 1), g.idx2);
    __syncthreads();
    if (threadIdx.y < 3) fys[R] = sfy[threadIdx.y + 1][threadIdx.x];
    else fys[R] = face_flux_y<MC>(sq, suy, SHO + STY, 1 + STY, b, b1, (int)(((mn >> 2) & 3u) << 1), g.idy2);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int j = jb + r;
        const int a = SHO + R * threadIdx.y + r;
        double fxe = __shfl_down_sync(0xffffffffu, fxw[r], 1);
        if (east_lane) fxe = sfe[R * threadIdx.y + r][threadIdx.x >> 5];
        if (!col_ok || j >= g.n2) continue;
        const long k = (long)j * s1 + i;
        double d = 0;
        if (i <= g.n1 - 2) d = -(fxe - fxw[r]);
        if (j <= g.n2 - 2) d -= fys[r + 1] - fys[r];
        d = d * (double)((mk[r] >> 4) & 1u);
        if (rk.write_ds) dq[k] = d;
        double acc;
        if (NC == 1) acc = rk_sum<1>(rk.c[0], 0, 0, d, 0, 0);
        else if (NC == 2) acc = rk_sum<2>(rk.c[0], rk.c[1], 0, d0[r], d, 0);
        else acc = rk_sum<3>(rk.c[0], rk.c[1], rk.c[2], d0[r], d1[r], d);
        rk.ynew[k] = rk_add(sq[a][b], acc);
    }
}

// ---------------------------------------------------------------------------
// The projection + diagnostics of the projecting models, TMA-fed (k_diag_tiled's
// arithmetic, bit for bit).  Boxes: the un-projected u.x, u.y (70 x 21), p (72 x 22)
// and the packed mask byte (96 x 21, engine.cuh: dmask) of a 64 x 16 tile with the
// halos the WENO inner product needs.  Phase 1 projects the whole velocity window in
// place in shared memory (u -= grad p * mask, operators.py:49-53; zero-filled boxes
// give 0 outside the array, as the guarded loads of k_diag_tiled do); phase 2 forms
// u, omega, ke of the thread's 4 consecutive rows from it.  38 KB per CTA.
// ---------------------------------------------------------------------------
constexpr int DUW = STX + 6, DUH = STY + 5;       // velocity boxes: columns i0-2 .. i0+67, rows j0-2 .. j0+18
constexpr int DPW = STX + 8, DPH = STY + 6;       // pressure box:   columns i0-4 .. i0+67, rows j0-3 .. j0+18
constexpr int DMW = STX + 32, DMH = STY + 5;      // mask box:       columns i0-16 .. i0+79, rows j0-2 .. j0+18
constexpr size_t DU_BYTES = pad128((size_t)DUW * DUH * 8), DP_BYTES = pad128((size_t)DPW * DPH * 8),
                 DM_BYTES = pad128((size_t)DMW * DMH);
struct DiagMaps { CUtensorMap ux, uy, p, m; };

// PROJECT = false (rotating shallow water): the velocity boxes are the state's u itself, there is
// no pressure box and no phase 1; the kernel also writes p = g (h + hb) / area (operators.py:110-111).
struct DiagRsw { const double *h, *hb; double *p; double g_over_area; };

#ifndef F2D_DIAG_CTAS
#define F2D_DIAG_CTAS 4
#endif
template <int MK, bool PROJECT>
__global__ void __launch_bounds__(256, F2D_DIAG_CTAS)
k_diag_tma(const __grid_constant__ DiagMaps M, Grid g, double *__restrict__ uxo, double *__restrict__ uyo,
           double *__restrict__ omega, double *__restrict__ ke, DiagRsw rsw) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    double(*sux)[DUW] = reinterpret_cast<double(*)[DUW]>(smem_raw);
    double(*suy)[DUW] = reinterpret_cast<double(*)[DUW]>(smem_raw + DU_BYTES);
    double(*spp)[DPW] = reinterpret_cast<double(*)[DPW]>(smem_raw + 2 * DU_BYTES);
    uint8_t(*smk)[DMW] = reinterpret_cast<uint8_t(*)[DMW]>(smem_raw + 2 * DU_BYTES + DP_BYTES);
    const int i0 = blockIdx.x * STX, j0 = blockIdx.y * STY;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, (unsigned)(2 * DUW * DUH * 8 + (PROJECT ? DPW * DPH * 8 : 0) + DMW * DMH));
        tma_load_2d(&sux[0][0], &M.ux, i0 - 2, j0 - 2, &bar);
        tma_load_2d(&suy[0][0], &M.uy, i0 - 2, j0 - 2, &bar);
        if (PROJECT) tma_load_2d(&spp[0][0], &M.p, i0 - 4, j0 - 3, &bar);
        tma_load_2d(&smk[0][0], &M.m, i0 - 16, j0 - 2, &bar);
    }
    mbar_wait(&bar, 0);
    if (PROJECT) {
        // ---- phase 1: addgrad(p) on the window (operators.py:49-53)
        for (int t = tid; t < DUW * DUH; t += 256) {
            const int a = t / DUW, b = t - a * DUW;
            const int j = j0 - 2 + a, i = i0 - 2 + b;
            const unsigned m = smk[a][b + 14];
            const double pc = spp[a + 1][b + 2];
            if (i >= 1) sux[a][b] -= (pc - spp[a + 1][b + 1]) * (double)(m & 1u);
            if (j >= 1) suy[a][b] -= (pc - spp[a][b + 2]) * (double)((m >> 1) & 1u);
        }
        __syncthreads();
    }
    // ---- phase 2
    constexpr int R = STY / 4;
    const int b = 2 + threadIdx.x, i = i0 + threadIdx.x;
    if (i >= g.n1) return;
    const long s1 = g.n1;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int a = 2 + R * threadIdx.y + r, j = j0 + R * threadIdx.y + r;
        if (j >= g.n2) break;
        const long k = (long)j * s1 + i;
        const unsigned m = smk[a][b + 14];
        const double ux0 = sux[a][b], uy0 = suy[a][b];
        if (PROJECT) {
            uxo[k] = ux0;
            uyo[k] = uy0;
        } else {
            rsw.p[k] = rsw.g_over_area * (rsw.h[k] + rsw.hb[k]);
        }
        double om = 0.0;
        if (j >= 1) om = -(ux0 - sux[a - 1][b]);
        if (i >= 1) om += uy0 - suy[a][b - 1];
        omega[k] = om * (double)((m >> 3) & 1u);
        const double mskd = (double)((m >> 2) & 1u);
        double e = 0.0;
        if (MK == F2D_METHOD_CLASSIC) {
            if (i <= g.n1 - 2) { double w = sux[a][b + 1]; e = w * (w * g.idx2) + ux0 * (ux0 * g.idx2); }
            if (j <= g.n2 - 2) { double w = suy[a + 1][b]; e += w * (w * g.idy2) + uy0 * (uy0 * g.idy2); }
            e *= mskd * 0.25;
        } else {
            constexpr int MM = MK > 3 ? 0 : MK;
            const int ox = (int)((m >> 4) & 3u) << 1;
            if (ox > 0) {
                double w3 = sux[a][b + 1];
                double Um = 0.5 * (ux0 * g.idx2 + w3 * g.idx2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (ox > 2) { w1 = sux[a][b - 1]; w4 = sux[a][b + 2]; }
                if (ox > 4) { w0 = sux[a][b - 2]; w5 = sux[a][b + 3]; }
                e += recon<MM>(ox, Um, w0, w1, ux0, w3, w4, w5) * Um;
            }
            const int oy = (int)((m >> 6) & 3u) << 1;
            if (oy > 0) {
                double w3 = suy[a + 1][b];
                double Um = 0.5 * (uy0 * g.idy2 + w3 * g.idy2);
                double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
                if (oy > 2) { w1 = suy[a - 1][b]; w4 = suy[a + 2][b]; }
                if (oy > 4) { w0 = suy[a - 2][b]; w5 = suy[a + 3][b]; }
                e += recon<MM>(oy, Um, w0, w1, uy0, w3, w4, w5) * Um;
            }
            e *= mskd * 0.5;
        }
        ke[k] = e;
    }
}

// meshes.py:135-143 for up to six arrays in one launch
struct FillMany { double *a[6]; int n; };
__global__ void k_fill_many(FillMany f, int n2, int n1, int nh) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n2 * 2 * nh) return;
    int j = t / (2 * nh), kk = t % (2 * nh);
    for (int q = 0; q < f.n; q++) {
        double *row = f.a[q] + (size_t)j * n1;
        if (kk < nh) row[kk] = row[n1 - 2 * nh + kk];
        else row[n1 - nh + (kk - nh)] = row[nh + (kk - nh)];
    }
}

// y copy of the same arrays for a truly periodic y direction (param.ywrap; after the x copy)
__global__ void k_fill_many_y(FillMany f, int n2, int n1, int nh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;
    if (i >= n1) return;
    const int ny = n2 - 2 * nh;
    int j = k < nh ? k : n2 - nh + (k - nh);
    int src = k < nh ? j + ny : j - ny;
    for (int q = 0; q < f.n; q++) f.a[q][(size_t)j * n1 + i] = f.a[q][(size_t)src * n1 + i];
}

// ---------------------------------------------------------------------------
// qg_projection of the tendency (operators.py:176-211), anomaly form
// ---------------------------------------------------------------------------
// pv = curl(du) * slip ; pv[1:,1:] += 1/4 sum4(dh * (-f0/H)) ; pv *= mskv
__global__ void __launch_bounds__(256)
k_qg_pv(Grid g, const double *__restrict__ dux, const double *__restrict__ duy,
        const double *__restrict__ dh, const int8_t *__restrict__ slip,
        const int8_t *__restrict__ mskv, double mf0H, double *__restrict__ pv) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.n1 || j >= g.n2) return;
    long k = (long)j * g.n1 + i;
    const long s1 = g.n1;
    double om = 0;
    if (j >= 1) om = -(dux[k] - dux[k - s1]);
    if (i >= 1) om += duy[k] - duy[k - 1];
    om *= (double)slip[k];
    if (i >= 1 && j >= 1)
        om += 0.25 * (((dh[k - s1 - 1] * mf0H + dh[k - 1] * mf0H) + dh[k - s1] * mf0H) + dh[k] * mf0H);
    pv[k] = om * (double)mskv[k];
}

// dh = verticestocenters(psi * (f0*area/g)) * msk * msk ; du = perpgrad(psi) ; fill
__global__ void __launch_bounds__(256)
k_qg_back(Grid g, const double *__restrict__ psi, const int8_t *__restrict__ msk,
          const int8_t *__restrict__ mskx, const int8_t *__restrict__ msky,
          const int8_t *__restrict__ mskv, double f0ag, double *__restrict__ dux,
          double *__restrict__ duy, double *__restrict__ dh) {
    THREAD_2D(g);
    if (j <= g.n2 - 2 && i <= g.n1 - 2) {
        int coef = mskv[k] + mskv[k + s1] + mskv[k + 1] + mskv[k + s1 + 1];
        double sum = ((psi[k] * f0ag + psi[k + s1] * f0ag) + psi[k + 1] * f0ag) + psi[k + s1 + 1] * f0ag;
        // the reference yields inf*0 = NaN in cells with no fluid vertex; those
        // cells are masked, we store 0 there instead.
        double v = coef > 0 ? (1.0 / (double)coef) * sum : 0.0;
        double m = (double)msk[k];
        dh[k_out] = (v * m) * m;
    } else if (i_out == i) {
        dh[k_out] = dh[k_out] * (double)msk[k] * (double)msk[k];
    }
    if (j <= g.n2 - 2) dux[k_out] = -(psi[k + s1] - psi[k]) * (double)mskx[k];
    if (i <= g.n1 - 2) duy[k_out] = (psi[k + 1] - psi[k]) * (double)msky[k];
}

// ---------------------------------------------------------------------------
// glue of the stream-function models (eulerpsi, qg) and of vectoradv
// ---------------------------------------------------------------------------
// centerstovertices (operators.py:126-133): v[1:,1:] = 1/4 sum4((a - sub*hb)) ; v *= mskv
__global__ void __launch_bounds__(256)
k_c2v(Grid g, const double *__restrict__ a, const double *__restrict__ hb, double sub,
      const int8_t *__restrict__ mskv, double *__restrict__ v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= g.n1 || j >= g.n2) return;
    long k = (long)j * g.n1 + i;
    const long s1 = g.n1;
    double r = v[k];                      // row 0 / column 0 keep their value (times mskv)
    if (i >= 1 && j >= 1) {
        auto A = [&](long m) { return a[m] - hb[m] * sub; };
        r = 0.25 * (((A(k - s1 - 1) + A(k - 1)) + A(k - s1)) + A(k));
    }
    v[k] = r * (double)mskv[k];
}

// perpgrad (operators.py:144-149) with the contravariant scaling, then fill
__global__ void __launch_bounds__(256)
k_perpgrad(Grid g, const double *__restrict__ psi, const int8_t *__restrict__ mskx,
           const int8_t *__restrict__ msky, double sx, double sy, double *__restrict__ ux,
           double *__restrict__ uy) {
    THREAD_2D(g);
    double vx = ux[k], vy = uy[k];
    if (j <= g.n2 - 2) vx = -(psi[k + s1] - psi[k]) * (double)mskx[k];
    if (i <= g.n1 - 2) vy = (psi[k + 1] - psi[k]) * (double)msky[k];
    ux[k_out] = vx * sx;
    uy[k_out] = vy * sy;
}

// vectoradv diag (equations.py:181-185): omega = curl(v) * slip ; q = 2 * ke(v, U) ; fill
template <int MK>
__global__ void __launch_bounds__(256)
k_vadv_diag(Grid g, const double *__restrict__ vx, const double *__restrict__ vy,
            const double *__restrict__ Ux, const double *__restrict__ Uy,
            const int8_t *__restrict__ msk, const int8_t *__restrict__ slip,
            const int8_t *__restrict__ okx, const int8_t *__restrict__ oky,
            double *__restrict__ omega, double *__restrict__ q) {
    THREAD_2D(g);
    double om = 0;
    if (j >= 1) om = -(vx[k] - vx[k - s1]);
    if (i >= 1) om += vy[k] - vy[k - 1];
    omega[k_out] = om * (double)slip[k];
    double e = 0;
    if (MK == F2D_METHOD_CLASSIC) {
        if (i <= g.n1 - 2) e = vx[k + 1] * Ux[k + 1] + vx[k] * Ux[k];
        if (j <= g.n2 - 2) e += vy[k + s1] * Uy[k + s1] + vy[k] * Uy[k];
        e *= (double)msk[k] * 0.25;
    } else {
        constexpr int M = MK > 3 ? 0 : MK;
        int ox = okx[k];
        if (ox > 0) {
            double Um = 0.5 * (Ux[k] + Ux[k + 1]);
            double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = vx[k], w3 = vx[k + 1];
            if (ox > 2) { w1 = vx[k - 1]; w4 = vx[k + 2]; }
            if (ox > 4) { w0 = vx[k - 2]; w5 = vx[k + 3]; }
            e += recon<M>(ox, Um, w0, w1, w2, w3, w4, w5) * Um;
        }
        int oy = oky[k];
        if (oy > 0) {
            double Um = 0.5 * (Uy[k] + Uy[k + s1]);
            double w0 = 0, w1 = 0, w4 = 0, w5 = 0, w2 = vy[k], w3 = vy[k + s1];
            if (oy > 2) { w1 = vy[k - s1]; w4 = vy[k + 2 * s1]; }
            if (oy > 4) { w0 = vy[k - 2 * s1]; w5 = vy[k + 3 * s1]; }
            e += recon<M>(oy, Um, w0, w1, w2, w3, w4, w5) * Um;
        }
        e *= (double)msk[k] * 0.5;
    }
    q[k_out] = e * 2;
}

// Model.set_dt (model.py:85): max|U.x|, max|U.y|
__global__ void __launch_bounds__(256)
k_maxabs(long n, const double *__restrict__ ux, const double *__restrict__ uy, double idx2,
         double idy2, double *part, unsigned int *count, double *out) {
    double v[2] = {0.0, 0.0};
    for (long k = (long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long)gridDim.x * blockDim.x) {
        v[0] = fmax(v[0], fabs(ux[k] * idx2));
        v[1] = fmax(v[1], fabs(uy[k] * idy2));
    }
    grid_reduce<OpMax, 2>(v, part, count, out);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define LAUNCH_CHECK(c)                 \
    do {                                \
        (c)->launches++;                \
        F2D_CUDA(cudaGetLastError());   \
    } while (0)

// mesh.fill of up to six arrays: the x-periodic halo copy (skipped when the producing kernel
// already evaluated the halo columns at their periodic image) and, for param.ywrap, the y copy
static int fill_many(f2d_ctx *c, const FillMany &f, bool x_copy = true) {
    if (f.n == 0) return F2D_OK;
    if (c->cfg.xperiodic && x_copy) {
        int tot = c->n2 * 2 * c->nh;
        k_fill_many<<<(tot + 127) / 128, 128, 0, c->stream>>>(f, c->n2, c->n1, c->nh);
        LAUNCH_CHECK(c);
    }
    if (c->cfg.yperiodic == 2) {
        k_fill_many_y<<<dim3((c->n1 + 127) / 128, 2 * c->nh), 128, 0, c->stream>>>(f, c->n2, c->n1, c->nh);
        LAUNCH_CHECK(c);
    }
    return F2D_OK;
}
static int fill_leaves(f2d_ctx *c, const std::vector<std::string> &names, const std::string &prefix = "") {
    if (c->cfg.yperiodic != 2) return F2D_OK;       // (the element-wise updates keep x-periodic halos consistent)
    FillMany f;
    f.n = 0;
    for (const std::string &nm : names) {
        f.a[f.n++] = c->f(prefix + nm);
        if (f.n == 6) { F2D_TRY(fill_many(c, f)); f.n = 0; }
    }
    return fill_many(c, f);
}

template <int MODEL, int NC = 0>
static int launch_rhs_mom(f2d_ctx *c, double *dux, double *duy, RkFuse rk = RkFuse()) {
    Grid g = grid_of(c);
    const double *p = c->has("p") ? c->f("p") : nullptr;
    const double *b = c->has("b") ? c->f("b") : nullptr;
    // vectoradv (equations.py:175-179): the advecting velocity is state.U itself, `q` plays ke
    const bool vadv = MODEL == M_VADV;
    if (vadv) g.idx2 = g.idy2 = 1.0;
    const double *ke = vadv ? c->f("q") : c->f("ke");
    const double *ax = vadv ? c->f("U.x") : c->f("u.x"), *ay = vadv ? c->f("U.y") : c->f("u.y");
    double fcor = c->cfg.f0 * c->area * 0.25;
    double halfdy = 0.5 * c->dy;
#define RHS_ARGS g, ax, ay, c->f("omega"), ke, p, b, c->m("ov.x"), c->m("ov.y"), \
                 c->m("mskx"), c->m("msky"), fcor, halfdy, dux, duy, rk
    switch (c->cfg.vortexforce) {
    case F2D_METHOD_WENO: k_rhs_mom<WENO, MODEL, NC><<<grd2d(c), blk2d(), 0, c->stream>>>(RHS_ARGS); break;
    case F2D_METHOD_UPWIND: k_rhs_mom<UPWIND, MODEL, NC><<<grd2d(c), blk2d(), 0, c->stream>>>(RHS_ARGS); break;
    case F2D_METHOD_CENTERED: k_rhs_mom<CENTERED, MODEL, NC><<<grd2d(c), blk2d(), 0, c->stream>>>(RHS_ARGS); break;
    case F2D_METHOD_CWENO: k_rhs_mom<CWENO, MODEL, NC><<<grd2d(c), blk2d(), 0, c->stream>>>(RHS_ARGS); break;
    default: set_error("bad vortexforce method"); return F2D_ERR_ARG;
    }
#undef RHS_ARGS
    LAUNCH_CHECK(c);
    return F2D_OK;
}

// direct_U: the model's transport velocity IS the contravariant state.U
// (eulerpsi, qg, advection), not sharp(u)
static int launch_divflux(f2d_ctx *c, const double *q, double *dq, bool direct_U = false, bool fill = true) {
    Grid g = grid_of(c);
    if (direct_U) g.idx2 = g.idy2 = 1.0;
    const double *vx = direct_U ? c->f("U.x") : c->f("u.x"), *vy = direct_U ? c->f("U.y") : c->f("u.y");
    double *fx = c->f("flx.x"), *fy = c->f("flx.y");
#define FLX_ARGS g, vx, vy, q, c->m("oc.x"), c->m("oc.y"), fx, fy
    switch (c->cfg.compflux) {
    case F2D_METHOD_WENO: k_flux<WENO><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_UPWIND: k_flux<UPWIND><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_CENTERED: k_flux<CENTERED><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_CWENO: k_flux<CWENO><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    default: set_error("bad compflux method"); return F2D_ERR_ARG;
    }
#undef FLX_ARGS
    LAUNCH_CHECK(c);
    if (fill) k_divflux<<<grd2d(c), blk2d(), 0, c->stream>>>(g, fx, fy, c->m("msk"), dq);
    else k_divflux_nofill<<<grd2d(c), blk2d(), 0, c->stream>>>(g, fx, fy, c->m("msk"), dq);
    LAUNCH_CHECK(c);
    return F2D_OK;
}

// addtracerequation (equations.py:217-226): after the model's own tendency,
// ds.tracer = -div(flux(tracer, s.U)), not filled
static int tracer_rhs(f2d_ctx *c, int k) {
    const int m = c->cfg.model;
    const bool direct = m == F2D_MODEL_EULERPSI || m == F2D_MODEL_QG || m == F2D_MODEL_ADVECTION ||
                        m == F2D_MODEL_VECTORADV;
    return launch_divflux(c, c->f("tracer"), c->f("ds" + std::to_string(k) + ".tracer"), direct, false);
}

// ---------------------------------------------------------------------------
// First guess of the elliptic solves inside f2d_step.  The reference's direct
// solve ignores what x holds; an iterative solve does not.  The solution of RK
// stage k changes slowly from one time step to the next, while it differs by
// O(1) between stages (the incremental RK form scales each stage's pressure by
// other coefficients) -- so the guess is extrapolated from the SAME stage of the
// previous steps by polynomial extrapolation (order = param.solver_guess, as far
// as the history allows; 0 keeps what x holds).
// ---------------------------------------------------------------------------
// Polynomial extrapolation in model time through the last `order` solutions of
// this stage: Lagrange weights for the nodes t_k evaluated at the current step's
// time (equal steps: the alternating binomials 4, -6, 4, -1 ...).  The pressure
// of the incremental RK form is proportional to the step length (u is already
// divergence free, only the dt-scaled increment is projected), so with
// scale_dt the smooth quantity p / dt is what gets extrapolated -- an adaptive
// dt (model.py:71-87) then costs no accuracy.
//
// The solve works IN the history: guess_begin hands out the oldest slot of the
// stage (one deeper than the extrapolation reads) as the array to solve in, makes
// the field name point at it, and describes the first guess x0 = sum w_k g_k for the
// solver's initial-residual kernel to form on the fly (mg.cu: k_cg_resid_guess);
// guess_end files the slot as the newest entry.  No guess array is written and read
// back, no solution is copied.
static int guess_begin(f2d_ctx *c, int stage, const char *field, bool scale_dt, GuessSpec *spec, double **x) {
    *spec = GuessSpec();
    *x = c->f(field);
    if (stage < 0 || stage >= 3 || c->guess_order <= 0) return F2D_OK;
    GuessHistory &G = c->guess[stage];
    const int depth = std::min(c->guess_order, 6) + 1;
    for (int k = 0; k < depth; k++)
        if (!G.g[k]) {
            F2D_CUDA(cudaMalloc(&G.g[k], c->n * sizeof(double)));
            F2D_CUDA(cudaMemsetAsync(G.g[k], 0, c->n * sizeof(double), c->stream));
        }
    double *prev = c->f(field);                 // the latest solution of this field, wherever it lives
    double *slot = G.g[depth - 1];              // oldest entry: not read by the extrapolation below
    if (slot == prev) {
        // (only if a caller shrank the history) never solve into the array the guess reads
        return F2D_OK;
    }
    int order = std::min(G.valid, std::min(c->guess_order, 6));
    const double tn = c->sim_t;
    if (order > 0 && !(tn > G.t[0])) order = 0;      // not a later time (dt <= 0)
    for (int k = 1; k < order; k++)                 // nodes must be strictly ordered in time
        if (!(G.t[k - 1] > G.t[k])) { order = k; break; }
    if (order > 0) {
        spec->n = order;
        for (int k = 0; k < order; k++) {
            double w = 1.0;
            for (int m = 0; m < order; m++)
                if (m != k) w *= (tn - G.t[m]) / (G.t[k] - G.t[m]);
            if (scale_dt) w *= c->sim_dt / G.dt[k];
            spec->g[k] = G.g[k];
            spec->w[k] = w;
        }
    } else {                                        // no usable history yet: start from what the field holds
        spec->n = 1;
        spec->g[0] = prev;
        spec->w[0] = 1.0;
    }
    if (!c->field_home.count(field)) c->field_home[field] = prev;    // the field's own allocation
    c->fields[field] = slot;
    *x = slot;
    return F2D_OK;
}

static int guess_end(f2d_ctx *c, int stage, const double *x) {
    if (stage < 0 || stage >= 3 || c->guess_order <= 0) return F2D_OK;
    GuessHistory &G = c->guess[stage];
    const int depth = std::min(c->guess_order, 6) + 1;
    if (x != G.g[depth - 1]) return F2D_OK;         // guess_begin did not rotate
    // rotate: the slot just solved in becomes entry 0
    double *last = G.g[depth - 1];
    for (int k = depth - 1; k > 0; k--) { G.g[k] = G.g[k - 1]; G.t[k] = G.t[k - 1]; G.dt[k] = G.dt[k - 1]; }
    G.g[0] = last;
    G.t[0] = c->sim_t;
    G.dt[0] = c->sim_dt;
    G.valid = std::min(G.valid + 1, depth - 1);
    return F2D_OK;
}

// U = sharp(u) (operators.py:59-64) for the models that carry the covariant u: the
// diagnostic kernels no longer store it every stage (16 B per point per stage that
// nothing on the device reads -- every kernel scales u itself); it is formed when
// somebody asks for it: a download, f2d_field_ptr, the bulk sums.
__global__ void __launch_bounds__(256)
k_sharp(long n, const double *__restrict__ ux, const double *__restrict__ uy, double idx2, double idy2,
        double *__restrict__ Ux, double *__restrict__ Uy) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Ux[k] = ux[k] * idx2;
    Uy[k] = uy[k] * idy2;
}

int ensure_U(f2d_ctx *c) {
    if (!c->U_stale) return F2D_OK;
    c->U_stale = false;
    if (!(c->has("u.x") && c->has("U.x"))) return F2D_OK;
    long n = (long)c->n;
    k_sharp<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, c->f("u.x"), c->f("u.y"), c->idx2, c->idy2,
                                                               c->f("U.x"), c->f("U.y"));
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

static std::string dsname(int k, const char *leaf) { return "ds" + std::to_string(k) + "." + leaf; }

static int model_rhs_core(f2d_ctx *c, int k);

// ---------------------------------------------------------------------------
// Forcing terms of the form  ds.<leaf> += amplitude * pattern  kept on the
// device (model.add_forcing, model.py:121-123; addforcingterm,
// equations.py:229-238: applied after the model's tendency, not filled).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_force(long n, double *__restrict__ ds, double *__restrict__ ustar, const double *__restrict__ F,
        double amp, double cu) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double f = amp * F[k];
    if (ds) ds[k] += f;
    if (ustar) ustar[k] += cu * f;
}

// ds may be null (last fused stage: the tendency is not stored); ustar != null
// adds cu * forcing to an already updated field (fused momentum stage)
static int apply_forcing(f2d_ctx *c, const std::string &leaf, double *ds, double *ustar = nullptr, double cu = 0.0) {
    auto it = c->forcing.find(leaf);
    if (it == c->forcing.end() || it->second.amplitude == 0.0) return F2D_OK;
    long n = (long)c->n;
    k_force<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, ds, ustar, it->second.pattern,
                                                               it->second.amplitude, cu);
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

int set_forcing(f2d_ctx *c, const std::string &leaf, const double *h_pattern, double amplitude) {
    if (std::find(c->prognostic.begin(), c->prognostic.end(), leaf) == c->prognostic.end()) {
        set_error("forcing: '%s' is not a prognostic field of this model", leaf.c_str());
        return F2D_ERR_ARG;
    }
    auto it = c->forcing.find(leaf);
    if (!h_pattern && it == c->forcing.end()) {
        set_error("forcing: no pattern set for '%s'", leaf.c_str());
        return F2D_ERR_STATE;
    }
    if (h_pattern) {
        f2d_ctx::Forcing &f = c->forcing[leaf];
        if (!f.pattern) F2D_CUDA(cudaMalloc(&f.pattern, c->n * sizeof(double)));
        F2D_CUDA(cudaMemcpyAsync(f.pattern, h_pattern, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        F2D_CUDA(cudaStreamSynchronize(c->stream));    // the host array may be pageable
        f.amplitude = amplitude;
    } else {
        it->second.amplitude = amplitude;
    }
    return F2D_OK;
}

int model_rhs(f2d_ctx *c, int k) {
    if (!c->mesh_ready) { set_error("f2d_rhs before f2d_set_mask"); return F2D_ERR_STATE; }
    if (k < 0 || k >= c->nstages) { set_error("stage %d out of range", k); return F2D_ERR_ARG; }
    F2D_TRY(model_rhs_core(c, k));
    if (c->tracer) F2D_TRY(tracer_rhs(c, k));
    for (const std::string &leaf : c->prognostic)
        F2D_TRY(apply_forcing(c, leaf, c->f("ds" + std::to_string(k) + "." + leaf)));
    return F2D_OK;
}

static int model_rhs_core(f2d_ctx *c, int k) {
    Grid g = grid_of(c);
    switch (c->cfg.model) {     // the scalar-transport models have no momentum tendency
    case F2D_MODEL_EULERPSI: return launch_divflux(c, c->f("omega"), c->f(dsname(k, "omega")), true);
    case F2D_MODEL_QG: return launch_divflux(c, c->f("pv"), c->f(dsname(k, "pv")), true);
    case F2D_MODEL_ADVECTION: return launch_divflux(c, c->f("q"), c->f(dsname(k, "q")), true);
    case F2D_MODEL_VECTORADV:
        return launch_rhs_mom<M_VADV>(c, c->f(dsname(k, "v.x")), c->f(dsname(k, "v.y")));
    default: break;
    }
    double *dux = c->f(dsname(k, "u.x")), *duy = c->f(dsname(k, "u.y"));
    switch (c->cfg.model) {
    case F2D_MODEL_EULER:
        return launch_rhs_mom<M_EULER>(c, dux, duy);
    case F2D_MODEL_BOUSSINESQ:
        F2D_TRY(launch_rhs_mom<M_BOUSS>(c, dux, duy));
        return launch_divflux(c, c->f("b"), c->f(dsname(k, "b")));
    case F2D_MODEL_RSW:
        F2D_TRY(launch_rhs_mom<M_RSW>(c, dux, duy));
        return launch_divflux(c, c->f("h"), c->f(dsname(k, "h")));
    case F2D_MODEL_QGRSW: {
        double *dh = c->f(dsname(k, "h"));
        // un-filled tendencies first (the projection reads them before fill)
        F2D_TRY(launch_rhs_mom<M_QGRSW>(c, dux, duy));
        F2D_TRY(launch_divflux(c, c->f("h"), dh));
        double mf0H = -c->cfg.f0 / c->cfg.H;
        k_qg_pv<<<grd2d(c), blk2d(), 0, c->stream>>>(g, dux, duy, dh, c->m("slip"), c->m("mskv"),
                                                      mf0H, c->f("pv"));
        LAUNCH_CHECK(c);
        GuessSpec gs;
        double *psi;
        F2D_TRY(guess_begin(c, c->stage_hint, "psi", false, &gs, &psi));
        F2D_TRY(mg_solve(c, F2D_SOLVER_HELMHOLTZ, c->f("pv"), 1.0, psi, nullptr, nullptr, &gs));
        F2D_TRY(guess_end(c, c->stage_hint, psi));
        double f0ag = c->cfg.f0 * c->area / c->cfg.g;
        k_qg_back<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f("psi"), c->m("msk"), c->m("mskx"),
                                                        c->m("msky"), c->m("mskv"), f0ag, dux, duy, dh);
        LAUNCH_CHECK(c);
        return F2D_OK;
    }
    }
    set_error("unknown model %d", c->cfg.model);
    return F2D_ERR_ARG;
}

int model_addto(f2d_ctx *c, int ncoef, const double *coefs) {
    if (ncoef < 1 || ncoef > 3 || ncoef > c->nstages) { set_error("addto: ncoef=%d", ncoef); return F2D_ERR_ARG; }
    long n = (long)c->n;
    unsigned grd = (unsigned)((n + 255) / 256);
    for (const std::string &leaf : c->prognostic) {
        double *y = c->f(leaf);
        const double *x0 = c->f(dsname(0, leaf.c_str()));
        const double *x1 = ncoef > 1 ? c->f(dsname(1, leaf.c_str())) : nullptr;
        const double *x2 = ncoef > 2 ? c->f(dsname(2, leaf.c_str())) : nullptr;
        double c0 = coefs[0], c1 = ncoef > 1 ? coefs[1] : 0, c2 = ncoef > 2 ? coefs[2] : 0;
        if (ncoef == 1) k_addto<1><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, c0, c1, c2);
        else if (ncoef == 2) k_addto<2><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, c0, c1, c2);
        else k_addto<3><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, c0, c1, c2);
        LAUNCH_CHECK(c);
    }
    return fill_leaves(c, c->prognostic);
}

template <bool PROJECT, int MODEL>
static int launch_diag(f2d_ctx *c, const double *uxin, const double *uyin, int mk) {
    Grid g = grid_of(c);
    const double *p = c->has("p") ? c->f("p") : nullptr;
    const double *h = c->has("h") ? c->f("h") : nullptr;
    double goa = c->cfg.g / c->area;
#define DIAG_ARGS g, uxin, uyin, p, h, c->hb, c->m("msk"), c->m("mskx"), c->m("msky"), c->m("slip"), \
                  c->m("ok.x"), c->m("ok.y"), goa, c->f("u.x"), c->f("u.y"), (double *)nullptr, (double *)nullptr, \
                  c->f("omega"), c->f("ke"), c->has("p") ? c->f("p") : nullptr
    switch (mk) {
    case -1: k_diag<-1, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    case F2D_METHOD_WENO: k_diag<0, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    case F2D_METHOD_UPWIND: k_diag<1, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    case F2D_METHOD_CENTERED: k_diag<2, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    case F2D_METHOD_CWENO: k_diag<3, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    case F2D_METHOD_CLASSIC: k_diag<4, PROJECT, MODEL><<<grd2d(c), blk2d(), 0, c->stream>>>(DIAG_ARGS); break;
    default: set_error("bad innerproduct method"); return F2D_ERR_ARG;
    }
#undef DIAG_ARGS
    LAUNCH_CHECK(c);
    c->U_stale = true;          // U = sharp(u) is formed on demand (ensure_U)
    return F2D_OK;
}

static int fill_diag_outputs(f2d_ctx *c) {
    c->U_stale = true;          // U = sharp(u) is formed on demand (ensure_U)
    FillMany f;
    f.n = 4;
    const char *names[4] = {"u.x", "u.y", "omega", "ke"};
    for (int q = 0; q < 4; q++) f.a[q] = c->f(names[q]);
    return fill_many(c, f);
}

static bool byte_map(f2d_ctx *c, const uint8_t *base, long rows, long pitch, int bh, int bw, CUtensorMap *out) {
    auto key = std::make_pair((const void *)base, (long)bh * 1024 + bw);
    auto it = c->tma_cache.find(key);
    if (it == c->tma_cache.end()) {
        f2d_ctx::TmaBlob blob;
        CUtensorMap m;
        blob.ok = tma_make_2d(&m, base, 1, rows, pitch, pitch, bh, bw);
        memcpy(blob.b, &m, sizeof(m));
        it = c->tma_cache.emplace(key, blob).first;
    }
    if (!it->second.ok) return false;
    memcpy(out, it->second.b, sizeof(CUtensorMap));
    return true;
}

static int stage_variant();
static bool field_map(f2d_ctx *c, const double *base, int bh, int bw, CUtensorMap *out);

static int launch_diag_tiled(f2d_ctx *c, const double *uxin, const double *uyin) {
    static const bool untiled = getenv("F2D_UNTILED_DIAG") != nullptr;
    if (untiled) return launch_diag<true, M_EULER>(c, uxin, uyin, c->cfg.innerproduct);
    Grid g = grid_of(c);
    if (stage_variant() == 0) {       // TMA-fed (default where the arrays qualify)
        DiagMaps M;
        if (field_map(c, uxin, DUH, DUW, &M.ux) && field_map(c, uyin, DUH, DUW, &M.uy) && field_map(c, c->f("p"), DPH, DPW, &M.p) &&
            byte_map(c, c->dmask, c->n2, c->dpitch, DMH, DMW, &M.m)) {
            constexpr size_t smem = 2 * DU_BYTES + DP_BYTES + DM_BYTES;
            dim3 grd((c->n1 + STX - 1) / STX, (c->n2 + STY - 1) / STY), blk(STX, 4);
#define DT_LAUNCH(MKV)                                                                                                \
    {                                                                                                                 \
        static bool once = false;                                                                                     \
        if (!once) {                                                                                                  \
            F2D_CUDA(cudaFuncSetAttribute(k_diag_tma<MKV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            once = true;                                                                                              \
        }                                                                                                             \
        k_diag_tma<MKV, true><<<grd, blk, smem, c->stream>>>(M, g, c->f("u.x"), c->f("u.y"), c->f("omega"), c->f("ke"), DiagRsw()); \
    }
            switch (c->cfg.innerproduct) {
            case F2D_METHOD_WENO: DT_LAUNCH(0) break;
            case F2D_METHOD_UPWIND: DT_LAUNCH(1) break;
            case F2D_METHOD_CENTERED: DT_LAUNCH(2) break;
            case F2D_METHOD_CWENO: DT_LAUNCH(3) break;
            case F2D_METHOD_CLASSIC: DT_LAUNCH(4) break;
            default: set_error("bad innerproduct method"); return F2D_ERR_ARG;
            }
#undef DT_LAUNCH
            LAUNCH_CHECK(c);
            return fill_diag_outputs(c);
        }
    }
    dim3 grd((c->n1 + DTX - 1) / DTX, (c->n2 + DTY - 1) / DTY), blk(DTX, 4);
#define TD_ARGS g, uxin, uyin, c->f("p"), c->m("msk"), c->m("mskx"), c->m("msky"), c->m("slip"), c->m("ok.x"), \
                c->m("ok.y"), c->f("u.x"), c->f("u.y"), (double *)nullptr, (double *)nullptr, c->f("omega"), c->f("ke")
    switch (c->cfg.innerproduct) {
    case F2D_METHOD_WENO: k_diag_tiled<0><<<grd, blk, 0, c->stream>>>(TD_ARGS); break;
    case F2D_METHOD_UPWIND: k_diag_tiled<1><<<grd, blk, 0, c->stream>>>(TD_ARGS); break;
    case F2D_METHOD_CENTERED: k_diag_tiled<2><<<grd, blk, 0, c->stream>>>(TD_ARGS); break;
    case F2D_METHOD_CWENO: k_diag_tiled<3><<<grd, blk, 0, c->stream>>>(TD_ARGS); break;
    case F2D_METHOD_CLASSIC: k_diag_tiled<4><<<grd, blk, 0, c->stream>>>(TD_ARGS); break;
    default: set_error("bad innerproduct method"); return F2D_ERR_ARG;
    }
#undef TD_ARGS
    LAUNCH_CHECK(c);
    return fill_diag_outputs(c);
}

// rotating shallow water: TMA-fed where the arrays qualify (no projection phase), else the per-point kernel.
// x-periodic halos: the per-point kernel evaluates them at the periodic image, the tiled one fills afterwards.
static int launch_diag_rsw(f2d_ctx *c) {
    Grid g = grid_of(c);
    if (stage_variant() == 0) {
        DiagMaps M;
        if (field_map(c, c->f("u.x"), DUH, DUW, &M.ux) && field_map(c, c->f("u.y"), DUH, DUW, &M.uy) &&
            byte_map(c, c->dmask, c->n2, c->dpitch, DMH, DMW, &M.m)) {
            M.p = M.ux;
            constexpr size_t smem = 2 * DU_BYTES + DP_BYTES + DM_BYTES;
            dim3 grd((c->n1 + STX - 1) / STX, (c->n2 + STY - 1) / STY), blk(STX, 4);
            DiagRsw R{c->f("h"), c->hb, c->f("p"), c->cfg.g / c->area};
#define DR_LAUNCH(MKV)                                                                                                 \
    {                                                                                                                  \
        static bool once = false;                                                                                      \
        if (!once) {                                                                                                   \
            F2D_CUDA(cudaFuncSetAttribute(k_diag_tma<MKV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            once = true;                                                                                               \
        }                                                                                                              \
        k_diag_tma<MKV, false><<<grd, blk, smem, c->stream>>>(M, g, nullptr, nullptr, c->f("omega"), c->f("ke"), R);   \
    }
            switch (c->cfg.innerproduct) {
            case F2D_METHOD_WENO: DR_LAUNCH(0) break;
            case F2D_METHOD_UPWIND: DR_LAUNCH(1) break;
            case F2D_METHOD_CENTERED: DR_LAUNCH(2) break;
            case F2D_METHOD_CWENO: DR_LAUNCH(3) break;
            case F2D_METHOD_CLASSIC: DR_LAUNCH(4) break;
            default: set_error("bad innerproduct method"); return F2D_ERR_ARG;
            }
#undef DR_LAUNCH
            LAUNCH_CHECK(c);
            c->U_stale = true;
            FillMany f;
            f.n = 3;
            f.a[0] = c->f("omega"); f.a[1] = c->f("ke"); f.a[2] = c->f("p");
            return fill_many(c, f);
        }
    }
    return launch_diag<false, M_RSW>(c, c->f("u.x"), c->f("u.y"), c->cfg.innerproduct);
}

// `pre`: the un-projected velocity already sits in tmp[0..1] (fused stage kernel)
static int model_diag_impl(f2d_ctx *c, bool pre) {
    if (!c->mesh_ready) { set_error("f2d_diag before f2d_set_mask"); return F2D_ERR_STATE; }
    Grid g = grid_of(c);
    switch (c->cfg.model) {
    case F2D_MODEL_EULER:
    case F2D_MODEL_BOUSSINESQ: {
        double *ux = c->f("u.x"), *uy = c->f("u.y");
        if (!pre) {
            // the projection reads neighbours of u, so it cannot run in place:
            // copy the un-projected velocity aside first
            size_t bytes = c->n * sizeof(double);
            F2D_CUDA(cudaMemcpyAsync(c->tmp[0], ux, bytes, cudaMemcpyDeviceToDevice, c->stream));
            F2D_CUDA(cudaMemcpyAsync(c->tmp[1], uy, bytes, cudaMemcpyDeviceToDevice, c->stream));
        }
        k_div_u<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->tmp[0], c->tmp[1], c->m("msk"), c->f("div"));
        LAUNCH_CHECK(c);
        // A p = -delta * area       (operators.py:117)
        GuessSpec gs;
        double *pp;
        F2D_TRY(guess_begin(c, c->stage_hint, "p", true, &gs, &pp));
        F2D_TRY(mg_solve(c, F2D_SOLVER_CENTERS, c->f("div"), -c->area, pp, nullptr, nullptr, &gs));
        F2D_TRY(guess_end(c, c->stage_hint, pp));
        F2D_TRY(launch_diag_tiled(c, c->tmp[0], c->tmp[1]));
        if (c->dist.on) {   // the next tendency reads omega +-3 rows, u and ke +-1
            void *a[4] = {ux, uy, c->f("omega"), c->f("ke")};
            F2D_TRY(dist_exchange(c, 4, a, (size_t)c->n1 * sizeof(double), c->n2, 0));
        }
        return F2D_OK;
    }
    case F2D_MODEL_RSW:
        return launch_diag_rsw(c);
    case F2D_MODEL_QGRSW:
        return launch_diag<false, M_QGRSW>(c, c->f("u.x"), c->f("u.y"), -1);
    case F2D_MODEL_ADVECTION:
        return F2D_OK;                                   // equations.py:167-168
    case F2D_MODEL_EULERPSI:                             // equations.py:81-85
    case F2D_MODEL_QG: {                                 // equations.py:98-101, operators.py:186-191
        const bool qg = c->cfg.model == F2D_MODEL_QG;
        if (qg && c->cfg.reserved[4]) { set_error("qg with beta != 0 (mesh.f) is not on the device path"); return F2D_ERR_UNSUPPORTED; }
        double *rhs = c->f(qg ? "work" : "vomega");
        k_c2v<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f(qg ? "pv" : "omega"), c->hb, qg ? c->cfg.f0 / +c->cfg.H : 0.0,
                                                     c->m("mskv"), rhs);
        LAUNCH_CHECK(c);
        GuessSpec gs;
        double *psi;
        F2D_TRY(guess_begin(c, c->stage_hint, "psi", false, &gs, &psi));
        F2D_TRY(mg_solve(c, qg ? F2D_SOLVER_HELMHOLTZ : F2D_SOLVER_VERTICES, rhs, 1.0, psi, nullptr, nullptr, &gs));
        F2D_TRY(guess_end(c, c->stage_hint, psi));
        // perpgrad(..., contravariant=True): u.x *= 1/dy**2, u.y *= 1/dx**2
        k_perpgrad<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f("psi"), c->m("mskx"), c->m("msky"), c->idy2, c->idx2,
                                                          c->f("U.x"), c->f("U.y"));
        LAUNCH_CHECK(c);
        return F2D_OK;
    }
    case F2D_MODEL_VECTORADV: {                          // equations.py:181-185
#define VD_ARGS g, c->f("v.x"), c->f("v.y"), c->f("U.x"), c->f("U.y"), c->m("msk"), c->m("slip"), c->m("ok.x"), c->m("ok.y"), \
                c->f("omega"), c->f("q")
        switch (c->cfg.innerproduct) {
        case F2D_METHOD_WENO: k_vadv_diag<0><<<grd2d(c), blk2d(), 0, c->stream>>>(VD_ARGS); break;
        case F2D_METHOD_UPWIND: k_vadv_diag<1><<<grd2d(c), blk2d(), 0, c->stream>>>(VD_ARGS); break;
        case F2D_METHOD_CENTERED: k_vadv_diag<2><<<grd2d(c), blk2d(), 0, c->stream>>>(VD_ARGS); break;
        case F2D_METHOD_CWENO: k_vadv_diag<3><<<grd2d(c), blk2d(), 0, c->stream>>>(VD_ARGS); break;
        case F2D_METHOD_CLASSIC: k_vadv_diag<4><<<grd2d(c), blk2d(), 0, c->stream>>>(VD_ARGS); break;
        default: set_error("bad innerproduct method"); return F2D_ERR_ARG;
        }
#undef VD_ARGS
        LAUNCH_CHECK(c);
        return F2D_OK;
    }
    }
    set_error("unknown model %d", c->cfg.model);
    return F2D_ERR_ARG;
}

int model_diag(f2d_ctx *c) { return model_diag_impl(c, false); }

// integrators.py:82-124, incremental form
static int rk_coefs(int integ, double dt, int stage, double *co) {
    switch (integ) {
    case F2D_INT_EF: co[0] = dt; return 1;
    case F2D_INT_RK3:
        if (stage == 0) { co[0] = dt; return 1; }
        if (stage == 1) { co[0] = -3 * dt / 4; co[1] = dt / 4; return 2; }
        co[0] = -dt / 12; co[1] = -dt / 12; co[2] = 2 * dt / 3; return 3;
    case F2D_INT_ENRK3:
        if (stage == 0) { co[0] = dt / 3; return 1; }
        if (stage == 1) { co[0] = -dt / 3 - 5 * dt / 48; co[1] = 15 * dt / 16; return 2; }
        co[0] = 5 * dt / 48 + dt / 10; co[1] = -7 * dt / 16; co[2] = 2 * dt / 5; return 3;
    }
    return 0;
}

// cached tensor map of an (n2,n1) fp64 field for boxes of bh x bw; false if TMA cannot be used for it
static bool field_map(f2d_ctx *c, const double *base, int bh, int bw, CUtensorMap *out) {
    auto key = std::make_pair((const void *)base, (long)bh * 1024 + bw);
    auto it = c->tma_cache.find(key);
    if (it == c->tma_cache.end()) {
        f2d_ctx::TmaBlob blob;
        static_assert(sizeof(CUtensorMap) == sizeof(blob.b), "CUtensorMap is a 128-byte blob");
        CUtensorMap m;
        blob.ok = tma_make_2d(&m, base, 8, c->n2, c->n1, c->n1, bh, bw);
        memcpy(blob.b, &m, sizeof(m));
        it = c->tma_cache.emplace(key, blob).first;
    }
    if (!it->second.ok) return false;
    memcpy(out, it->second.b, sizeof(CUtensorMap));
    return true;
}

static int fill_stage_outputs(f2d_ctx *c, double *dux, double *duy, const RkFuse &rk, bool x_copy = true) {
    FillMany f;
    f.n = 0;
    if (rk.write_ds) { f.a[f.n++] = dux; f.a[f.n++] = duy; }
    f.a[f.n++] = rk.ubx; f.a[f.n++] = rk.uby;
    return fill_many(c, f, x_copy);
}

// F2D_STAGE=tma (default where the arrays qualify: even n1) | point (one thread per point,
// L1-cached loads: the round-1 kernel) | tiled (its shared-memory variant, kept for study)
static int stage_variant() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("F2D_STAGE");
        v = !e ? 0 : (!strcmp(e, "point") ? 1 : (!strcmp(e, "tiled") ? 2 : 0));
        if (getenv("F2D_TILED_STAGE")) v = 2;
    }
    return v;
}

template <int MODEL, int NC>
static int launch_stage_tiled(f2d_ctx *c, double *dux, double *duy, const RkFuse &rk) {
    const int variant = stage_variant();
    Grid g = grid_of(c);
    const double *b = c->has("b") ? c->f("b") : nullptr;
    if (variant == 0) {
        StageMaps M;
        bool ok = field_map(c, c->f("omega"), SOH, SOW, &M.om) && field_map(c, c->f("u.x"), S1H, S1W, &M.ux) &&
                  field_map(c, c->f("u.y"), S1H, S1W, &M.uy) && field_map(c, c->f("ke"), S1H, S1W, &M.ke);
        if (ok && MODEL == M_BOUSS) ok = field_map(c, b, S1H, S1W, &M.b);
        else if (ok && MODEL == M_RSW) ok = field_map(c, c->f("p"), S1H, S1W, &M.b);
        else if (ok) M.b = M.ke;
        if (ok) {
            constexpr size_t smem = SOM_BYTES + ((MODEL == M_BOUSS || MODEL == M_RSW) ? 4 : 3) * S1_BYTES;
            dim3 grd((c->n1 + STX - 1) / STX, (c->n2 + STY - 1) / STY), blk(STX, 4);
#define TMA_ARGS M, g, c->smask, 0.5 * c->dy, c->cfg.f0 * c->area * 0.25, dux, duy, rk
#define TMA_LAUNCH(MV)                                                                                        \
    {                                                                                                         \
        static bool once = false;                                                                             \
        if (!once) {                                                                                          \
            F2D_CUDA(cudaFuncSetAttribute(k_stage_tma<MV, MODEL, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            once = true;                                                                                      \
        }                                                                                                     \
        k_stage_tma<MV, MODEL, NC><<<grd, blk, smem, c->stream>>>(TMA_ARGS);                                  \
    }
            switch (c->cfg.vortexforce) {
            case F2D_METHOD_WENO: TMA_LAUNCH(WENO) break;
            case F2D_METHOD_UPWIND: TMA_LAUNCH(UPWIND) break;
            case F2D_METHOD_CENTERED: TMA_LAUNCH(CENTERED) break;
            case F2D_METHOD_CWENO: TMA_LAUNCH(CWENO) break;
            default: set_error("bad vortexforce method"); return F2D_ERR_ARG;
            }
#undef TMA_LAUNCH
#undef TMA_ARGS
            LAUNCH_CHECK(c);
            return fill_stage_outputs(c, dux, duy, rk);
        }
    }
    if (variant != 2 || MODEL == M_RSW) {
        F2D_TRY((launch_rhs_mom<MODEL, NC>(c, dux, duy, rk)));
        return fill_stage_outputs(c, dux, duy, rk, false);     // the kernel evaluated the x halo at its periodic image
    }
    dim3 grd((c->n1 + DTX - 1) / DTX, (c->n2 + DTY - 1) / DTY), blk(DTX, 4);
#define ST_ARGS g, c->f("u.x"), c->f("u.y"), c->f("omega"), c->f("ke"), b, c->m("ov.x"), c->m("ov.y"), \
                c->m("mskx"), c->m("msky"), 0.5 * c->dy, dux, duy, rk
    switch (c->cfg.vortexforce) {
    case F2D_METHOD_WENO: k_stage_tiled<WENO, MODEL, NC><<<grd, blk, 0, c->stream>>>(ST_ARGS); break;
    case F2D_METHOD_UPWIND: k_stage_tiled<UPWIND, MODEL, NC><<<grd, blk, 0, c->stream>>>(ST_ARGS); break;
    case F2D_METHOD_CENTERED: k_stage_tiled<CENTERED, MODEL, NC><<<grd, blk, 0, c->stream>>>(ST_ARGS); break;
    case F2D_METHOD_CWENO: k_stage_tiled<CWENO, MODEL, NC><<<grd, blk, 0, c->stream>>>(ST_ARGS); break;
    default: set_error("bad vortexforce method"); return F2D_ERR_ARG;
    }
#undef ST_ARGS
    LAUNCH_CHECK(c);
    return fill_stage_outputs(c, dux, duy, rk);
}

// One-kernel transport of the model's flux-form scalar `leaf` for RK stage s (k_transport_tma):
// ds_s.leaf and y* = leaf + sum c_i ds_i.leaf, the latter in tmp[2]; *done = false (nothing
// launched) where TMA does not apply (odd n1, F2D_STAGE=point), the caller then takes the
// flux / divergence / update kernels.  The caller swaps fields[leaf] and tmp[2] once nothing
// reads the old scalar any more.
template <int NC>
static int launch_transport(f2d_ctx *c, const std::string &leaf, int s, const double *co, bool *done) {
    *done = false;
    if (stage_variant() != 0 || !c->tmp[2] || !c->tmask) return F2D_OK;
    TransportMaps M;
    if (!(field_map(c, c->f(leaf), TQH, TQW, &M.q) && field_map(c, c->f("u.x"), S1H, S1W, &M.ux) &&
          field_map(c, c->f("u.y"), S1H, S1W, &M.uy)))
        return F2D_OK;
    RkScalarOut rk;
    for (int k = 0; k < 3; k++) rk.c[k] = k < NC ? co[k] : 0.0;
    for (int k = 0; k < 2; k++) rk.d[k] = k < NC - 1 ? c->f(dsname(k, leaf.c_str())) : nullptr;
    rk.ynew = c->tmp[2];
    rk.write_ds = s < c->nstages - 1;
    double *dq = c->f(dsname(s, leaf.c_str()));
    Grid g = grid_of(c);
    constexpr size_t smem = TQ_BYTES + 2 * S1_BYTES + 4 * STX * 8 + STY * 2 * 8;
    dim3 grd((c->n1 + STX - 1) / STX, (c->n2 + STY - 1) / STY), blk(STX, 4);
#define TR_LAUNCH(MC)                                                                                         \
    {                                                                                                         \
        static bool once = false;                                                                             \
        if (!once) {                                                                                          \
            F2D_CUDA(cudaFuncSetAttribute(k_transport_tma<MC, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            once = true;                                                                                      \
        }                                                                                                     \
        k_transport_tma<MC, NC><<<grd, blk, smem, c->stream>>>(M, g, c->tmask, dq, rk);                      \
    }
    switch (c->cfg.compflux) {
    case F2D_METHOD_WENO: TR_LAUNCH(WENO) break;
    case F2D_METHOD_UPWIND: TR_LAUNCH(UPWIND) break;
    case F2D_METHOD_CENTERED: TR_LAUNCH(CENTERED) break;
    case F2D_METHOD_CWENO: TR_LAUNCH(CWENO) break;
    default: set_error("bad compflux method"); return F2D_ERR_ARG;
    }
#undef TR_LAUNCH
    LAUNCH_CHECK(c);
    // mesh.fill of the tendency (the reference fills every model-owned ds) and of y*
    FillMany f;
    f.n = 0;
    if (rk.write_ds) f.a[f.n++] = dq;
    f.a[f.n++] = rk.ynew;
    F2D_TRY(fill_many(c, f));
    // forcing: into the stored tendency and, scaled by this stage's own coefficient, into y*
    F2D_TRY(apply_forcing(c, leaf.c_str(), rk.write_ds ? dq : nullptr, rk.ynew, co[NC - 1]));
    *done = true;
    return F2D_OK;
}
static int launch_transport(f2d_ctx *c, const std::string &leaf, int s, int nc, const double *co, bool *done) {
    if (nc == 1) return launch_transport<1>(c, leaf, s, co, done);
    if (nc == 2) return launch_transport<2>(c, leaf, s, co, done);
    return launch_transport<3>(c, leaf, s, co, done);
}

// Euler / Boussinesq stage with the velocity update fused into the tendency
// kernel: u* = u + sum c_i ds_i lands in tmp[0..1], the projection writes u.
static int fused_stage(f2d_ctx *c, int s, int nc, const double *co) {
    RkFuse rk;
    for (int k = 0; k < 3; k++) rk.c[k] = k < nc ? co[k] : 0.0;
    for (int k = 0; k < 2; k++) {
        rk.dx[k] = k < nc - 1 ? c->f(dsname(k, "u.x")) : nullptr;
        rk.dy[k] = k < nc - 1 ? c->f(dsname(k, "u.y")) : nullptr;
    }
    rk.ubx = c->tmp[0]; rk.uby = c->tmp[1];
    rk.write_ds = s < c->nstages - 1;
    double *dux = c->f(dsname(s, "u.x")), *duy = c->f(dsname(s, "u.y"));
    const bool bouss = c->cfg.model == F2D_MODEL_BOUSSINESQ;
#define STAGE(NCV)                                                                     \
    (bouss ? launch_stage_tiled<M_BOUSS, NCV>(c, dux, duy, rk) : launch_stage_tiled<M_EULER, NCV>(c, dux, duy, rk))
    if (nc == 1) F2D_TRY(STAGE(1));
    else if (nc == 2) F2D_TRY(STAGE(2));
    else F2D_TRY(STAGE(3));
#undef STAGE
    // momentum forcing: into the stored tendency and, scaled by this stage's own RK
    // coefficient, into u* which the kernel above has already formed
    F2D_TRY(apply_forcing(c, "u.x", rk.write_ds ? dux : nullptr, c->tmp[0], co[nc - 1]));
    F2D_TRY(apply_forcing(c, "u.y", rk.write_ds ? duy : nullptr, c->tmp[1], co[nc - 1]));
    if (c->dist.on) {   // ghost rows of the updated velocity
        void *a[2] = {c->tmp[0], c->tmp[1]};
        F2D_TRY(dist_exchange(c, 2, a, (size_t)c->n1 * sizeof(double), c->n2, 0));
    }
    // advected scalars: tendency from the old u (u.x/u.y are still the old velocity,
    // the momentum kernel wrote u* aside), then their RK update
    auto update_scalar = [&](const char *leaf) -> int {
        long n = (long)c->n;
        unsigned grd = (unsigned)((n + 255) / 256);
        double *y = c->f(leaf);
        const double *x0 = c->f(dsname(0, leaf)), *x1 = nc > 1 ? c->f(dsname(1, leaf)) : nullptr,
                     *x2 = nc > 2 ? c->f(dsname(2, leaf)) : nullptr;
        if (nc == 1) k_addto<1><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, co[0], 0, 0);
        else if (nc == 2) k_addto<2><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, co[0], co[1], 0);
        else k_addto<3><<<grd, 256, 0, c->stream>>>(n, y, x0, x1, x2, co[0], co[1], co[2]);
        LAUNCH_CHECK(c);
        if (c->dist.on) F2D_TRY(dist_exchange1(c, y, (size_t)c->n1 * sizeof(double), c->n2, 0));
        return fill_leaves(c, {leaf});
    };
    bool b_fused = false;        // buoyancy: flux, divergence and update in one kernel where TMA applies
    if (bouss) F2D_TRY(launch_transport(c, "b", s, nc, co, &b_fused));
    if (bouss && !b_fused) F2D_TRY(launch_divflux(c, c->f("b"), c->f(dsname(s, "b"))));
    if (c->tracer) F2D_TRY(tracer_rhs(c, s));
    if (bouss && !b_fused) F2D_TRY(apply_forcing(c, "b", c->f(dsname(s, "b"))));
    if (c->tracer) F2D_TRY(apply_forcing(c, "tracer", c->f(dsname(s, "tracer"))));
    if (bouss && !b_fused) F2D_TRY(update_scalar("b"));
    if (b_fused) {
        std::swap(c->fields["b"], c->tmp[2]);          // b* becomes b
        if (c->dist.on) F2D_TRY(dist_exchange1(c, c->f("b"), (size_t)c->n1 * sizeof(double), c->n2, 0));
    }
    if (c->tracer) F2D_TRY(update_scalar("tracer"));
    return model_diag_impl(c, true);
}

// Rotating shallow water stage with the RK update fused the same way: the momentum
// kernel leaves u* = u + sum c_i ds_i in tmp[0..1] (neighbouring threads still read u),
// the flux-divergence kernel updates h in place, then u and u* swap roles (a pointer
// swap, no copy).  Three k_addto passes per stage (25 % of the 8192^2 step) disappear.
template <int NC>
static int rsw_thickness_unfused(f2d_ctx *c, int s, const double *co, int write_ds, double *dh);
template <int NC>
static int fused_stage_rsw_nc(f2d_ctx *c, int s, const double *co) {
    RkFuse rk;
    for (int k = 0; k < 3; k++) rk.c[k] = k < NC ? co[k] : 0.0;
    for (int k = 0; k < 2; k++) {
        rk.dx[k] = k < NC - 1 ? c->f(dsname(k, "u.x")) : nullptr;
        rk.dy[k] = k < NC - 1 ? c->f(dsname(k, "u.y")) : nullptr;
    }
    rk.ubx = c->tmp[0]; rk.uby = c->tmp[1];
    rk.write_ds = s < c->nstages - 1;
    double *dux = c->f(dsname(s, "u.x")), *duy = c->f(dsname(s, "u.y")), *dh = c->f(dsname(s, "h"));
    F2D_TRY((launch_stage_tiled<M_RSW, NC>(c, dux, duy, rk)));      // TMA-fed where the arrays qualify
    F2D_TRY(apply_forcing(c, "u.x", rk.write_ds ? dux : nullptr, c->tmp[0], co[NC - 1]));
    F2D_TRY(apply_forcing(c, "u.y", rk.write_ds ? duy : nullptr, c->tmp[1], co[NC - 1]));
    // thickness: fluxes from the OLD u (u.x / u.y still are); one kernel where TMA applies (h* in tmp[2]) ...
    bool h_fused = false;
    F2D_TRY(launch_transport<NC>(c, "h", s, co, &h_fused));
    if (!h_fused) F2D_TRY((rsw_thickness_unfused<NC>(c, s, co, rk.write_ds, dh)));
    if (c->tracer) {
        F2D_TRY(tracer_rhs(c, s));
        F2D_TRY(apply_forcing(c, "tracer", c->f(dsname(s, "tracer"))));
        long n = (long)c->n;
        unsigned grd = (unsigned)((n + 255) / 256);
        const double *x0 = c->f(dsname(0, "tracer")), *x1 = NC > 1 ? c->f(dsname(1, "tracer")) : nullptr,
                     *x2 = NC > 2 ? c->f(dsname(2, "tracer")) : nullptr;
        k_addto<NC><<<grd, 256, 0, c->stream>>>(n, c->f("tracer"), x0, x1, x2, co[0], NC > 1 ? co[1] : 0.0, NC > 2 ? co[2] : 0.0);
        LAUNCH_CHECK(c);
    }
    if (c->dist.on) {
        // slabs: ghost rows of u* and h* from their owners.  The diagnostics below then run on the whole
        // local array: omega (u +-2 rows), ke and p are exact on every ghost row the next stage reads
        // (omega +-3, the others +-1 around the owned rows; G = 8 ghost rows), no exchange of their own.
        void *a[3] = {c->tmp[0], c->tmp[1], h_fused ? c->tmp[2] : c->f("h")};
        F2D_TRY(dist_exchange(c, 3, a, (size_t)c->n1 * sizeof(double), c->n2, 0));
        if (c->tracer) F2D_TRY(dist_exchange1(c, c->f("tracer"), (size_t)c->n1 * sizeof(double), c->n2, 0));
    }
    // u* becomes u, h* becomes h
    std::swap(c->fields["u.x"], c->tmp[0]);
    std::swap(c->fields["u.y"], c->tmp[1]);
    if (h_fused) std::swap(c->fields["h"], c->tmp[2]);
    return model_diag_impl(c, false);
}

// ... else the flux kernel, then divergence + update of h in place
template <int NC>
static int rsw_thickness_unfused(f2d_ctx *c, int s, const double *co, int write_ds, double *dh) {
    Grid g = grid_of(c);
    double *fx = c->f("flx.x"), *fy = c->f("flx.y");
#define FLX_ARGS g, c->f("u.x"), c->f("u.y"), c->f("h"), c->m("oc.x"), c->m("oc.y"), fx, fy
    switch (c->cfg.compflux) {
    case F2D_METHOD_WENO: k_flux<WENO><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_UPWIND: k_flux<UPWIND><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_CENTERED: k_flux<CENTERED><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    case F2D_METHOD_CWENO: k_flux<CWENO><<<grd2d(c), blk2d(), 0, c->stream>>>(FLX_ARGS); break;
    default: set_error("bad compflux method"); return F2D_ERR_ARG;
    }
#undef FLX_ARGS
    LAUNCH_CHECK(c);
    RkScalar rs;
    for (int k = 0; k < 3; k++) rs.c[k] = k < NC ? co[k] : 0.0;
    for (int k = 0; k < 2; k++) rs.d[k] = k < NC - 1 ? c->f(dsname(k, "h")) : nullptr;
    rs.y = c->f("h");
    rs.write_ds = write_ds;
    k_divflux_upd<NC><<<grd2d(c), blk2d(), 0, c->stream>>>(g, fx, fy, c->m("msk"), dh, rs);
    LAUNCH_CHECK(c);
    (void)s;
    return apply_forcing(c, "h", write_ds ? dh : nullptr, c->f("h"), co[NC - 1]);
}

static int fused_stage_rsw(f2d_ctx *c, int s, int nc, const double *co) {
    if (nc == 1) return fused_stage_rsw_nc<1>(c, s, co);
    if (nc == 2) return fused_stage_rsw_nc<2>(c, s, co);
    return fused_stage_rsw_nc<3>(c, s, co);
}

// ---------------------------------------------------------------------------
// Leap-frog + Robert-Asselin filter (integrators.py:20-53).  scratch = [sb, sa, ds]
//   first:  sb = s ; sa = s ; s += dt ds
//   else:   sa += 2 dt ds ; s += (gamma sa + gamma sb) - 2 gamma s ; (sa, s, sb) <- (s, sa, s)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_lfra(long n, double *__restrict__ s, double *__restrict__ sb, double *__restrict__ sa,
       const double *__restrict__ ds, double dt, double gamma, int first) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    double sv = s[k], d = ds[k];
    if (first) {
        sb[k] = sv;
        sa[k] = sv;
        s[k] = sv + dt * d;
    } else {
        double san = sa[k] + (2 * dt) * d;
        double acc = gamma * san;
        acc = acc + gamma * sb[k];
        acc = acc + (-2 * gamma) * sv;
        double sf = sv + acc;
        sb[k] = sf;        // rightpermute(sa, s, sb): sb <- s, s <- sa, sa <- sb
        s[k] = san;
        sa[k] = sf;
    }
}

int model_step_lfra(f2d_ctx *c, double dt, int first, double gamma) {
    if (!c->mesh_ready) { set_error("f2d_step before f2d_set_mask"); return F2D_ERR_STATE; }
    if (c->cfg.integrator != F2D_INT_LFRA) { set_error("context was not created with the LFRA integrator"); return F2D_ERR_STATE; }
    if (c->dist.on) { set_error("LFRA is not available in slab mode"); return F2D_ERR_UNSUPPORTED; }
    F2D_TRY(model_rhs(c, 2));
    long n = (long)c->n;
    unsigned grd = (unsigned)((n + 255) / 256);
    for (const std::string &leaf : c->prognostic) {
        k_lfra<<<grd, 256, 0, c->stream>>>(n, c->f(leaf), c->f(dsname(0, leaf.c_str())), c->f(dsname(1, leaf.c_str())),
                                           c->f(dsname(2, leaf.c_str())), dt, gamma, first);
        LAUNCH_CHECK(c);
    }
    F2D_TRY(fill_leaves(c, c->prognostic));
    return model_diag(c);
}

static bool fuse_rsw_off() {
    static const bool off = getenv("F2D_RSW_UNFUSED") != nullptr;      // A/B against the addto_list passes
    return off;
}

int model_step(f2d_ctx *c, double dt, int nsteps) {
    if (!c->mesh_ready) { set_error("f2d_step before f2d_set_mask"); return F2D_ERR_STATE; }
    if (c->cfg.integrator == F2D_INT_LFRA) { set_error("LFRA steps go through f2d_step_lfra"); return F2D_ERR_STATE; }
    for (int it = 0; it < nsteps; it++) {
        c->sim_dt = dt;
        for (int s = 0; s < c->nstages; s++) {
            double co[3];
            int nc = rk_coefs(c->cfg.integrator, dt, s, co);
            c->stage_hint = s;      // the solves of this stage may use its history
            int st;
            if (c->cfg.model == F2D_MODEL_EULER || c->cfg.model == F2D_MODEL_BOUSSINESQ)
                st = fused_stage(c, s, nc, co);
            else if (c->cfg.model == F2D_MODEL_RSW && c->tmp[0] && !fuse_rsw_off())
                st = fused_stage_rsw(c, s, nc, co);
            else if (c->dist.on && c->cfg.model != F2D_MODEL_QGRSW) {
                set_error("slab mode steps through the fused stage kernels (euler, boussinesq, rsw) or the qgrsw stage");
                st = F2D_ERR_UNSUPPORTED;
            } else {
                // qgrsw on slabs needs no message of its own: the vertex Helmholtz solve returns psi with
                // current ghost rows (mg_solve), k_qg_back REPLACES the tendencies by functions of psi +-1
                // row (exact 7 ghost rows deep), the point-wise update keeps u and h exact that deep, and the
                // diagnostics (u +-2) and the raw tendencies (omega +-3, h +-3) then are exact on the owned
                // rows, which is all the right-hand side of the next solve needs (its residual is exchanged).
                st = model_rhs(c, s);
                if (st == F2D_OK) st = model_addto(c, nc, co);
                if (st == F2D_OK) st = model_diag(c);
            }
            c->stage_hint = -1;
            F2D_TRY(st);
        }
        c->sim_t += dt;
    }
    return F2D_OK;
}

int max_abs_U(f2d_ctx *c, double *out) {
    int nb = c->nsm * 8;
    // models that carry the covariant u derive U = sharp(u); the others hold U itself
    const bool cov = c->has("u.x");
    k_maxabs<<<nb, 256, 0, c->stream>>>((long)c->n, c->f(cov ? "u.x" : "U.x"), c->f(cov ? "u.y" : "U.y"),
                                        cov ? c->idx2 : 1.0, cov ? c->idy2 : 1.0, c->d_part, c->d_count, c->d_scal + 12);
    F2D_CUDA(cudaMemcpyAsync(c->d_scal + 8, c->d_scal + 12, 2 * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    LAUNCH_CHECK(c);
    F2D_TRY(dist_allreduce(c, c->d_scal + 8, 2, true));
    F2D_CUDA(cudaMemcpyAsync(c->h_scal, c->d_scal + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    *out = c->h_scal[0] + c->h_scal[1];
    return F2D_OK;
}


// ---------------------------------------------------------------------------
// Bulk diagnostics (diagnostics.py:39-62): the six whole-array sums behind
// ke / enstrophy / vorticity / angular momentum, in ONE pass over the state:
//   out = [ sum ke, sum omega^2, sum omega, sum U.y*xv, sum U.x*yu, sum msk ]
// xv = (i - nh + 0.5) dx, yu = (j + row0 - nh + 0.5) dy  (meshes.py:56-65).
// The reference sums the whole haloed array; slabs sum the rows they own.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bulk(int n1, int j0, int j1, int nh, int row0, double dx, double dy,
       const double *__restrict__ ke, const double *__restrict__ om, const double *__restrict__ Ux,
       const double *__restrict__ Uy, const int8_t *__restrict__ msk, double *part,
       unsigned int *count, double *out) {
    double v[6] = {0, 0, 0, 0, 0, 0};
    long n = (long)(j1 - j0) * n1;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        int j = j0 + (int)(t / n1), i = (int)(t % n1);
        long k = (long)j * n1 + i;
        double w = om[k];
        v[0] += ke[k];
        v[1] += w * w;
        v[2] += w;
        v[3] += Uy[k] * (((double)(i - nh) + 0.5) * dx);
        v[4] += Ux[k] * (((double)(j + row0 - nh) + 0.5) * dy);
        v[5] += (double)msk[k];
    }
    grid_reduce<OpSum, 6>(v, part, count, out);
}

int bulk_sums(f2d_ctx *c, int row0, double *out) {
    if (!c->mesh_ready) { set_error("f2d_bulk_sums before f2d_set_mask"); return F2D_ERR_STATE; }
    if (!(c->has("ke") && c->has("omega") && c->has("U.x"))) {
        set_error("bulk diagnostics need ke, omega and U (euler, boussinesq, rsw, qgrsw)");
        return F2D_ERR_UNSUPPORTED;
    }
    F2D_TRY(ensure_U(c));
    int gs = c->cfg.reserved[1], gn = c->cfg.reserved[2];
    // at an interface the local array is [G ghost rows | owned rows | G ghost rows] with no wall halo on that side
    int j0 = gs > 0 ? gs : 0, j1 = gn > 0 ? c->n2 - gn : c->n2;
    int nb = c->nsm * 4;
    k_bulk<<<nb, 256, 0, c->stream>>>(c->n1, j0, j1, c->nh, row0, c->dx, c->dy, c->f("ke"), c->f("omega"),
                                      c->f("U.x"), c->f("U.y"), c->m("msk"), c->d_part, c->d_count, c->d_scal + 16);
    LAUNCH_CHECK(c);
    F2D_TRY(dist_allreduce(c, c->d_scal + 16, 6, false));
    F2D_CUDA(cudaMemcpyAsync(c->h_scal + 16, c->d_scal + 16, 6 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 6; k++) out[k] = c->h_scal[16 + k];
    return F2D_OK;
}

// ---------------------------------------------------------------------------
// History output (io.py:12-32 stores float32): convert on the device into a
// per-field float32 staging buffer (step stream), copy it out on a separate
// stream so the step loop keeps running; f2d_io_sync waits for the copies.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_to_f32(long n, const double *__restrict__ a, float *__restrict__ b) {
    long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) b[k] = (float)a[k];
}

int download_f32(f2d_ctx *c, const double *src, const std::string &key, float *h_dst) {
    if (!c->io_stream) F2D_CUDA(cudaStreamCreateWithFlags(&c->io_stream, cudaStreamNonBlocking));
    f2d_ctx::IoStage &st = c->io_stage[key];
    if (!st.d) {
        F2D_CUDA(cudaMalloc(&st.d, c->n * sizeof(float)));
        F2D_CUDA(cudaEventCreateWithFlags(&st.filled, cudaEventDisableTiming));
        F2D_CUDA(cudaEventCreateWithFlags(&st.drained, cudaEventDisableTiming));
    } else {
        F2D_CUDA(cudaStreamWaitEvent(c->stream, st.drained, 0));   // previous copy of this buffer
    }
    long n = (long)c->n;
    k_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(n, src, st.d);
    LAUNCH_CHECK(c);
    F2D_CUDA(cudaEventRecord(st.filled, c->stream));
    F2D_CUDA(cudaStreamWaitEvent(c->io_stream, st.filled, 0));
    F2D_CUDA(cudaMemcpyAsync(h_dst, st.d, c->n * sizeof(float), cudaMemcpyDeviceToHost, c->io_stream));
    F2D_CUDA(cudaEventRecord(st.drained, c->io_stream));
    return F2D_OK;
}

// ---------------------------------------------------------------------------
// bench.py hook: time one kernel alone, `reps` launches between two events on
// the context's stream.  *bytes = algorithmic bytes of ONE launch (every
// distinct array read once / written once, fp64 = 8 B, masks = 1 B).
// Clobbers the scratch tendencies and the diagnostics: call it last.
// ---------------------------------------------------------------------------
int bench_step_kernel(f2d_ctx *c, const char *name, int reps, float *ms, double *bytes) {
    if (!c->mesh_ready) { set_error("bench before f2d_set_mask"); return F2D_ERR_STATE; }
    const int model = c->cfg.model;
    const bool proj = model == F2D_MODEL_EULER || model == F2D_MODEL_BOUSSINESQ;       // projecting models: fused stage
    const bool sw = model == F2D_MODEL_RSW || model == F2D_MODEL_QGRSW;
    if (!(proj || sw) || c->nstages != 3) {
        set_error("kernel benches are defined for euler, boussinesq, rsw and qgrsw with a 3-stage integrator");
        return F2D_ERR_UNSUPPORTED;
    }
    std::string k(name);
    Grid g = grid_of(c);
    double npts = (double)c->n;
    if (k == "project_diag") {
        if (!proj) { set_error("'%s' belongs to the projecting models", name); return F2D_ERR_ARG; }
        F2D_CUDA(cudaMemcpyAsync(c->tmp[0], c->f("u.x"), c->n * 8, cudaMemcpyDeviceToDevice, c->stream));
        F2D_CUDA(cudaMemcpyAsync(c->tmp[1], c->f("u.y"), c->n * 8, cudaMemcpyDeviceToDevice, c->stream));
    }
    const char *scalar = model == F2D_MODEL_BOUSSINESQ ? "b" : "h";       // the flux-form scalar of the model
    for (int pass = 0; pass < 2; pass++) {   // pass 0 = warm-up
        int n = pass == 0 ? 2 : reps;
        if (pass == 1) F2D_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int r = 0; r < n; r++) {
            if (k == "advection" && (proj || (model == F2D_MODEL_RSW && c->tmp[0]))) {
                // stage 2 of rk3, fused with the RK update:
                // R u.x u.y omega ke ds0.x ds0.y [b], W ds1.x ds1.y ub.x ub.y, packed masks (ov.x ov.y mskx msky in 1 byte)
                RkFuse rk;
                rk.c[0] = rk.c[1] = rk.c[2] = 0.0;
                rk.dx[0] = c->f("ds0.u.x"); rk.dy[0] = c->f("ds0.u.y"); rk.dx[1] = rk.dy[1] = nullptr;
                rk.ubx = c->tmp[0]; rk.uby = c->tmp[1]; rk.write_ds = 1;
                if (model == F2D_MODEL_EULER) F2D_TRY((launch_stage_tiled<M_EULER, 2>(c, c->f("ds1.u.x"), c->f("ds1.u.y"), rk)));
                else if (model == F2D_MODEL_RSW) F2D_TRY((launch_stage_tiled<M_RSW, 2>(c, c->f("ds1.u.x"), c->f("ds1.u.y"), rk)));     // + p
                else F2D_TRY((launch_stage_tiled<M_BOUSS, 2>(c, c->f("ds1.u.x"), c->f("ds1.u.y"), rk)));
                *bytes = npts * ((model == F2D_MODEL_EULER ? 10 : 11) * 8 + 1);     // one packed mask byte (smask)
            } else if (k == "advection") {
                // rsw / qgrsw momentum tendency (vortex force + Coriolis [+ grad(ke + p)]), no fused update:
                // R u.x u.y omega [ke p], W ds.x ds.y, masks ov.x ov.y mskx msky
                if (model == F2D_MODEL_RSW) F2D_TRY((launch_rhs_mom<M_RSW>(c, c->f("ds1.u.x"), c->f("ds1.u.y"))));
                else F2D_TRY((launch_rhs_mom<M_QGRSW>(c, c->f("ds1.u.x"), c->f("ds1.u.y"))));
                *bytes = npts * ((model == F2D_MODEL_RSW ? 7 : 5) * 8 + 4);
            } else if (k == "flux_div" && (sw || model == F2D_MODEL_BOUSSINESQ)) {
                // divflux of the advected scalar, two kernels: R u.x u.y q, W flx.x flx.y (oc.x oc.y);
                // R flx.x flx.y, W dq (msk)
                // where TMA applies (rsw, boussinesq): the one-kernel transport of stage 2 with zero
                // coefficients, R u.x u.y q ds0, W ds1 q* (tmask) -- fill kernels included in the time
                bool fused = false;
                if (model != F2D_MODEL_QGRSW) {
                    const double zero[3] = {0.0, 0.0, 0.0};
                    F2D_TRY(launch_transport<2>(c, scalar, 1, zero, &fused));
                }
                if (fused) *bytes = npts * (6 * 8 + 1);
                else {
                    F2D_TRY(launch_divflux(c, c->f(scalar), c->f(dsname(1, scalar))));
                    *bytes = npts * (8 * 8 + 3);
                }
            } else if (k == "rk_update") {
                // R u ds0 ds1 ds2, W u (one component; coefficients 0 keep u intact)
                long n1 = (long)c->n;
                k_addto<3><<<(unsigned)((n1 + 255) / 256), 256, 0, c->stream>>>(
                    n1, c->f("u.x"), c->f("ds0.u.x"), c->f("ds1.u.x"), c->f("ds2.u.x"), 0.0, 0.0, 0.0);
                LAUNCH_CHECK(c);
                *bytes = npts * (5 * 8);
            } else if (k == "divergence" && proj) {
                // R u.x u.y, W div, mask msk
                k_div_u<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f("u.x"), c->f("u.y"), c->m("msk"), c->f("div"));
                LAUNCH_CHECK(c);
                *bytes = npts * (3 * 8 + 1);
            } else if (k == "project_diag") {
                // R p u.x u.y, W u.x u.y omega ke, one packed mask byte (dmask)
                F2D_TRY(launch_diag_tiled(c, c->tmp[0], c->tmp[1]));
                *bytes = npts * (7 * 8 + 1);
            } else if (k == "diag" && sw) {
                // rsw: R u.x u.y h hb, W omega ke p (msk slip ok.x ok.y); qgrsw: R u.x u.y, W omega (slip)
                if (model == F2D_MODEL_RSW) F2D_TRY(launch_diag_rsw(c));      // TMA-fed where the arrays qualify
                else F2D_TRY((launch_diag<false, M_QGRSW>(c, c->f("u.x"), c->f("u.y"), -1)));
                *bytes = npts * (model == F2D_MODEL_RSW ? 7 * 8 + 1 : 3 * 8 + 1);
            } else if (k == "qg_pv" && model == F2D_MODEL_QGRSW) {
                // R du.x du.y dh, W pv (slip mskv)
                k_qg_pv<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f("ds1.u.x"), c->f("ds1.u.y"), c->f("ds1.h"), c->m("slip"),
                                                              c->m("mskv"), -c->cfg.f0 / c->cfg.H, c->f("pv"));
                LAUNCH_CHECK(c);
                *bytes = npts * (4 * 8 + 2);
            } else if (k == "qg_back" && model == F2D_MODEL_QGRSW) {
                // R psi, W du.x du.y dh (msk mskx msky mskv)
                k_qg_back<<<grd2d(c), blk2d(), 0, c->stream>>>(g, c->f("psi"), c->m("msk"), c->m("mskx"), c->m("msky"), c->m("mskv"),
                                                                c->cfg.f0 * c->area / c->cfg.g, c->f("ds1.u.x"), c->f("ds1.u.y"), c->f("ds1.h"));
                LAUNCH_CHECK(c);
                *bytes = npts * (4 * 8 + 4);
            } else { set_error("unknown kernel '%s' for this model", name); return F2D_ERR_ARG; }
        }
    }
    F2D_CUDA(cudaEventRecord(c->ev1, c->stream));
    F2D_CUDA(cudaEventSynchronize(c->ev1));
    F2D_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    *ms /= reps;
    return F2D_OK;
}

}  // namespace f2d
