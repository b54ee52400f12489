// Mesh derivation on the device and the three stand-alone stencil kernels.
#include <cstdarg>
#include <cstdio>

#include "engine.cuh"
#include "weno.cuh"

namespace f2d {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return F2D_ERR_CUDA;
}

// ---------------------------------------------------------------------------
// Mesh.finalize on the device
// ---------------------------------------------------------------------------
// meshes.py:77-86 (mskx, msky, mskv) and noslip.py:4-36 (slip coefficient,
// always 0 or 1 so it is kept as int8).
__global__ void k_masks(const int8_t *__restrict__ msk, int8_t *__restrict__ mskx,
                        int8_t *__restrict__ msky, int8_t *__restrict__ mskv,
                        int8_t *__restrict__ slip, int n2, int n1, int nh, int noslip) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= n1) return;
    size_t k = (size_t)j * n1 + i;
    int m = msk[k];
    int mw = i > 0 ? msk[k - 1] : 0;
    int ms = j > 0 ? msk[k - n1] : 0;
    int msw = (i > 0 && j > 0) ? msk[k - n1 - 1] : 0;
    mskx[k] = (i > 0) ? (int8_t)(m * mw) : 0;
    msky[k] = (j > 0) ? (int8_t)(m * ms) : 0;
    mskv[k] = (i > 0 && j > 0) ? (int8_t)(ms * msw * m * mw) : 0;
    int coef = m + mw + ms + msw;
    int freeslip = coef == 4, nos = coef > 0;
    int s = freeslip;
    if (noslip & F2D_NOSLIP_ALL) s = nos;
    else {
        if ((noslip & F2D_NOSLIP_LEFT) && i == nh) s = nos;
        if ((noslip & F2D_NOSLIP_RIGHT) && i == n1 - nh) s = nos;
        if ((noslip & F2D_NOSLIP_BOTTOM) && j == nh) s = nos;
        if ((noslip & F2D_NOSLIP_TOP) && j == n2 - nh) s = nos;
    }
    slip[k] = (int8_t)s;
}

// meshes.py:146-186 in flat index space; `shift` carries the sign the caller
// passes there.  m is any array whose non-zero entries mean "usable".
__device__ __forceinline__ int order_at(const int8_t *__restrict__ m, long n, long i, long shift,
                                        int maxorder) {
    auto M = [&](long k) { return m[k] != 0 ? 1 : 0; };
    int s2, s4, s6;
    if (shift > 0) {
        s2 = (i - shift >= 0) ? M(i - shift) + M(i) : 0;
        s4 = ((i - 2 * shift >= 0) && (i + shift < n)) ? M(i - 2 * shift) + M(i + shift) + s2 : 0;
        s6 = ((i - 3 * shift >= 0) && (i + 2 * shift < n)) ? M(i - 3 * shift) + M(i + 2 * shift) + s4 : 0;
    } else {
        s2 = (i - shift < n) ? M(i - shift) + M(i) : 0;
        s4 = ((i - 2 * shift < n) && (i + shift >= 0)) ? M(i - 2 * shift) + M(i + shift) + s2 : 0;
        s6 = ((i - 3 * shift < n) && (i + 2 * shift >= 0)) ? M(i - 3 * shift) + M(i + 2 * shift) + s4 : 0;
    }
    int o = (s6 == 6) ? 6 : ((s4 == 4) ? 4 : ((s2 == 2) ? 2 : 0));
    return o < maxorder ? o : maxorder;
}

// meshes.py:88-104
__global__ void k_orders(const int8_t *__restrict__ msk, const int8_t *__restrict__ mskx,
                         const int8_t *__restrict__ msky, const int8_t *__restrict__ slip,
                         int8_t *ocx, int8_t *ocy, int8_t *ovx, int8_t *ovy, int8_t *okx,
                         int8_t *oky, long n, long n1, int maxorder) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ocx[i] = (int8_t)order_at(msk, n, i, 1, maxorder);
    ocy[i] = (int8_t)order_at(msk, n, i, n1, maxorder);
    ovx[i] = (int8_t)(order_at(slip, n, i, -1, maxorder) * msky[i]);
    ovy[i] = (int8_t)(order_at(slip, n, i, -n1, maxorder) * mskx[i]);
    okx[i] = (int8_t)order_at(mskx, n, i, -1, maxorder);
    oky[i] = (int8_t)order_at(msky, n, i, -n1, maxorder);
}

// one byte per point carrying every mask / stencil-order value a fused kernel needs (engine.cuh)
__global__ void k_pack_masks(int n2, int n1, int dpitch, const int8_t *__restrict__ msk, const int8_t *__restrict__ mskx,
                             const int8_t *__restrict__ msky, const int8_t *__restrict__ slip,
                             const int8_t *__restrict__ ovx, const int8_t *__restrict__ ovy,
                             const int8_t *__restrict__ okx, const int8_t *__restrict__ oky,
                             const int8_t *__restrict__ ocx, const int8_t *__restrict__ ocy,
                             uint8_t *__restrict__ smask, uint8_t *__restrict__ dmask, uint8_t *__restrict__ tmask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= dpitch) return;
    uint8_t d = 0;
    if (i < n1) {
        size_t k = (size_t)j * n1 + i;
        smask[k] = (uint8_t)((ovx[k] >> 1) | ((ovy[k] >> 1) << 2) | ((mskx[k] != 0) << 4) | ((msky[k] != 0) << 5));
        tmask[k] = (uint8_t)((ocx[k] >> 1) | ((ocy[k] >> 1) << 2) | ((msk[k] != 0) << 4));
        d = (uint8_t)((mskx[k] != 0) | ((msky[k] != 0) << 1) | ((msk[k] != 0) << 2) | ((slip[k] != 0) << 3) |
                      ((okx[k] >> 1) << 4) | ((oky[k] >> 1) << 6));
    }
    dmask[(size_t)j * dpitch + i] = d;
}

int build_mesh(f2d_ctx *c, const int8_t *h_msk) {
    const int nh = c->nh, n1 = c->n1, n2 = c->n2;
    int8_t *msk = c->m("msk");
    if (h_msk) {
        F2D_CUDA(cudaMemcpyAsync(msk, h_msk, c->n, cudaMemcpyHostToDevice, c->stream));
    } else {
        // meshes.py:70-75
        std::vector<int8_t> m(c->n, 0);
        int j0 = c->cfg.yperiodic ? 0 : nh, j1 = c->cfg.yperiodic ? n2 : n2 - nh;
        int i0 = c->cfg.xperiodic ? 0 : nh, i1 = c->cfg.xperiodic ? n1 : n1 - nh;
        for (int j = j0; j < j1; j++)
            for (int i = i0; i < i1; i++) m[(size_t)j * n1 + i] = 1;
        F2D_CUDA(cudaMemcpyAsync(msk, m.data(), c->n, cudaMemcpyHostToDevice, c->stream));
        F2D_CUDA(cudaStreamSynchronize(c->stream));
    }
    dim3 blk(128), grd((n1 + 127) / 128, n2);
    k_masks<<<grd, blk, 0, c->stream>>>(msk, c->m("mskx"), c->m("msky"), c->m("mskv"), c->m("slip"),
                                        n2, n1, nh, c->cfg.noslip);
    long n = (long)c->n;
    k_orders<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(
        msk, c->m("mskx"), c->m("msky"), c->m("slip"), c->m("oc.x"), c->m("oc.y"), c->m("ov.x"),
        c->m("ov.y"), c->m("ok.x"), c->m("ok.y"), n, n1, c->cfg.maxorder);
    c->launches += 2;
    F2D_CUDA(cudaGetLastError());
    c->dpitch = (n1 + 15) & ~15;
    if (!c->smask) F2D_CUDA(cudaMalloc(&c->smask, c->n));
    if (!c->dmask) F2D_CUDA(cudaMalloc(&c->dmask, (size_t)c->dpitch * n2));
    if (!c->tmask) F2D_CUDA(cudaMalloc(&c->tmask, c->n));
    k_pack_masks<<<dim3((c->dpitch + 127) / 128, n2), 128, 0, c->stream>>>(
        n2, n1, c->dpitch, msk, c->m("mskx"), c->m("msky"), c->m("slip"), c->m("ov.x"), c->m("ov.y"), c->m("ok.x"),
        c->m("ok.y"), c->m("oc.x"), c->m("oc.y"), c->smask, c->dmask, c->tmask);
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    c->mesh_ready = true;
    return F2D_OK;
}

// ---------------------------------------------------------------------------
// weno.py:346-405: flat-index kernels, one thread per element.
// Negative indices wrap like Python's (numba does the same); reads past the
// end are undefined in the reference and clamped here.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double at(const double *__restrict__ a, long k, long n) {
    if (k < 0) k += n;
    if (k < 0 || k >= n) k = 0;
    return a[k];
}

template <int M>
__global__ void k_compflux(double *__restrict__ flx, const double *__restrict__ U,
                           const double *__restrict__ q, const int8_t *__restrict__ o, long n,
                           long s) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ord = o[i];
    double r = 0;
    if (ord > 0) {
        double u = U[i];
        // compflux window: q[i-3s .. i+2s]
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
        double w2 = at(q, i - s, n), w3 = q[i];
        if (ord > 2) { w1 = at(q, i - 2 * s, n); w4 = at(q, i + s, n); }
        if (ord > 4) { w0 = at(q, i - 3 * s, n); w5 = at(q, i + 2 * s, n); }
        r = recon<M>(ord, u, w0, w1, w2, w3, w4, w5) * u;
    }
    flx[i] = r;
}

template <int M>
__global__ void k_vortexforce(double *__restrict__ du, const double *__restrict__ V,
                              const double *__restrict__ q, const int8_t *__restrict__ o, long n,
                              long s, long s2, double sign) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ord = o[i];
    double r = 0;
    if (ord > 0) {
        double Vm = 0.25 * (((V[i] + at(V, i + s, n)) + at(V, i - s2, n)) + at(V, i + s - s2, n));
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
        double w2 = q[i], w3 = at(q, i + s, n);
        if (ord > 2) { w1 = at(q, i - s, n); w4 = at(q, i + 2 * s, n); }
        if (ord > 4) { w0 = at(q, i - 2 * s, n); w5 = at(q, i + 3 * s, n); }
        r = (sign * recon<M>(ord, Vm, w0, w1, w2, w3, w4, w5)) * Vm;
    }
    du[i] = r;
}

template <int M>
__global__ void k_innerproduct(double *__restrict__ ke, const double *__restrict__ U,
                               const double *__restrict__ q, const int8_t *__restrict__ o, long n,
                               long s) {
    long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int ord = o[i];
    if (ord > 0) {
        double Um = 0.5 * (U[i] + at(U, i + s, n));
        double w0 = 0, w1 = 0, w4 = 0, w5 = 0;
        double w2 = q[i], w3 = at(q, i + s, n);
        if (ord > 2) { w1 = at(q, i - s, n); w4 = at(q, i + 2 * s, n); }
        if (ord > 4) { w0 = at(q, i - 2 * s, n); w5 = at(q, i + 3 * s, n); }
        ke[i] += recon<M>(ord, Um, w0, w1, w2, w3, w4, w5) * Um;
    }
}

#define DISPATCH_METHOD(method, KERNEL, ...)                                       \
    switch (method) {                                                              \
    case F2D_METHOD_WENO: KERNEL<WENO><<<grd, 256, 0, c->stream>>>(__VA_ARGS__); break;         \
    case F2D_METHOD_UPWIND: KERNEL<UPWIND><<<grd, 256, 0, c->stream>>>(__VA_ARGS__); break;     \
    case F2D_METHOD_CENTERED: KERNEL<CENTERED><<<grd, 256, 0, c->stream>>>(__VA_ARGS__); break; \
    case F2D_METHOD_CWENO: KERNEL<CWENO><<<grd, 256, 0, c->stream>>>(__VA_ARGS__); break;       \
    default: set_error("unknown method %d", method); return F2D_ERR_ARG;           \
    }

int op_compflux(f2d_ctx *c, double *flx, const double *U, const double *q, const int8_t *o,
                long n, long s, int method) {
    unsigned grd = (unsigned)((n + 255) / 256);
    DISPATCH_METHOD(method, k_compflux, flx, U, q, o, n, s);
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}
int op_vortexforce(f2d_ctx *c, double *du, const double *V, const double *q, const int8_t *o,
                   long n, long s, long s2, int sign, int method) {
    unsigned grd = (unsigned)((n + 255) / 256);
    DISPATCH_METHOD(method, k_vortexforce, du, V, q, o, n, s, s2, (double)sign);
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}
int op_innerproduct(f2d_ctx *c, double *ke, const double *U, const double *q, const int8_t *o,
                    long n, long s, int method) {
    unsigned grd = (unsigned)((n + 255) / 256);
    DISPATCH_METHOD(method, k_innerproduct, ke, U, q, o, n, s);
    c->launches++;
    F2D_CUDA(cudaGetLastError());
    return F2D_OK;
}

// meshes.py:135-143
__global__ void k_fill(double *__restrict__ a, int n2, int n1, int nh) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n2 * 2 * nh) return;
    int j = t / (2 * nh), k = t % (2 * nh);
    double *row = a + (size_t)j * n1;
    if (k < nh) row[k] = row[n1 - 2 * nh + k];
    else row[n1 - nh + (k - nh)] = row[nh + (k - nh)];
}

// the same in y for a truly periodic direction (param.ywrap; the reference's yperiodic never
// wraps, SURVEY note Y): halo rows are copies of the interior rows ny away.  Runs after the
// x copy and over every column, so the corners are periodic images too.
__global__ void k_fill_y(double *__restrict__ a, int n2, int n1, int nh) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int k = blockIdx.y;                      // 0 .. 2 nh - 1
    if (i >= n1) return;
    const int ny = n2 - 2 * nh;
    int j = k < nh ? k : n2 - nh + (k - nh);
    int src = k < nh ? j + ny : j - ny;
    a[(size_t)j * n1 + i] = a[(size_t)src * n1 + i];
}

int op_fill(f2d_ctx *c, double *a) {
    if (c->cfg.xperiodic) {
        int tot = c->n2 * 2 * c->nh;
        k_fill<<<(tot + 127) / 128, 128, 0, c->stream>>>(a, c->n2, c->n1, c->nh);
        c->launches++;
        F2D_CUDA(cudaGetLastError());
    }
    if (c->cfg.yperiodic == 2) {
        k_fill_y<<<dim3((c->n1 + 127) / 128, 2 * c->nh), 128, 0, c->stream>>>(a, c->n2, c->n1, c->nh);
        c->launches++;
        F2D_CUDA(cudaGetLastError());
    }
    return F2D_OK;
}

}  // namespace f2d
