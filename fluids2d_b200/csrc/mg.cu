// Matrix-free masked geometric multigrid for the reference's 5-point operators
// (elliptic.py:114-195), used as a CG preconditioner or as a stand-alone
// V-cycle iteration.  It replaces Poisson2D's SuperLU factorisation
// (elliptic.py:78, :83).
//
//   operator   L = -A (symmetric positive semi-definite):
//              (L x)_P = diag_P x_P - sum_nb c_nb x_nb,
//              c = dy/dx (W,E), dx/dy (S,N) across open faces,
//              diag = sum(open c) + maindiag            cell centres, Neumann
//              diag = 2(dy/dx + dx/dy) + maindiag       vertices, Dirichlet
//   fine level operates in place on the reference-layout (n2,n1) arrays, one
//              byte of mask bits per point (no coefficient arrays);
//   coarsening 2x2 aggregation of cells; coarse couplings are half the sum of
//              the fine couplings crossing the coarse face (rediscretisation
//              on masked grids), mass terms add up;
//   transfer   mask-renormalised bilinear prolongation P (weights 9,3,3,1 over
//              the fluid parents, so constants are preserved next to walls),
//              restriction R = P^T  -> symmetric V-cycle, valid CG preconditioner;
//   smoother   red-black Gauss-Seidel, nu1 sweeps (R,B) down, nu2 (B,R) up.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "engine.cuh"
#include "reduce.cuh"
#include "mg_tiles.cuh"

namespace f2d {

#define LAUNCH_CHECK(c)                 \
    do {                                \
        (c)->launches++;                \
        F2D_CUDA(cudaGetLastError());   \
    } while (0)

// device scalar slots
// device scalar slots (doubles)
//   k_cg_resid  -> S_RR, S_SUMR, S_FF      k_cg_update -> S_RR, S_SUMR
//   up leg / k_dot2 -> S_RZNEW, S_SUMZ     k_cg_dir_apply -> S_PQ, S_RZ0 + (it & 1)
enum { S_RR = 0, S_SUMR = 1, S_FF = 2, S_RZ0 = 3, S_RZ1 = 4, S_PQ = 5, S_RZNEW = 6, S_SUMZ = 7, S_TMP = 8,
       S_COMP = 32 /* .. +7: per-component sums (k_comp_sums) */, S_TRUE = 40 /* .. +2: true residual at exit */ };

// ------------------------------------------------------------------ helpers --
__device__ __forceinline__ bool fine_index(const FineView &F, int j, int i, long &idx) {
    int aj = F.oj + j, ai = F.oi + i;
    if (aj < 0 || aj >= F.n2 || ai < 0 || ai >= F.n1) return false;
    idx = (long)aj * F.n1 + ai;
    return true;
}

struct Stencil {
    double cw, ce, cs, cn, diag;
    long iw, ie, is, in;
};

__device__ __forceinline__ Stencil fine_stencil(const FineView &F, int j, int i, long idx, uint8_t c) {
    Stencil s;
    s.cw = (c & NB_W) ? F.cx : 0.0;
    s.ce = (c & NB_E) ? F.cx : 0.0;
    s.cs = (c & NB_S) ? F.cy : 0.0;
    s.cn = (c & NB_N) ? F.cy : 0.0;
    s.diag = F.dirichlet ? (2.0 * (F.cx + F.cy) + F.shift) : (((s.cw + s.ce) + s.cs) + s.cn + F.shift);
    s.iw = idx - 1;
    s.ie = idx + 1;
    if (F.periodic) {
        if (i == 0) s.iw = idx + (F.nx - 1);
        if (i == F.nx - 1) s.ie = idx - (F.nx - 1);
    }
    s.is = idx - F.n1;
    s.in = idx + F.n1;
    if (F.periodic_y) {
        if (j == 0) s.is = idx + (long)(F.ny - 1) * F.n1;
        if (j == F.ny - 1) s.in = idx - (long)(F.ny - 1) * F.n1;
    }
    return s;
}

// sum_nb c_nb x_nb with closed faces skipped (their neighbours may hold anything)
__device__ __forceinline__ double fine_offdiag(const Stencil &s, uint8_t c, const double *__restrict__ x) {
    double a = 0.0;
    if (c & NB_W) a += s.cw * x[s.iw];
    if (c & NB_E) a += s.ce * x[s.ie];
    if (c & NB_S) a += s.cs * x[s.is];
    if (c & NB_N) a += s.cn * x[s.in];
    return a;
}

// shared-memory arrays per level of the single-CTA tail kernel: x, b, r, cx, cy, dinv, w
constexpr int TAIL_ARRAYS = 7;

// normaliser of the prolongation weights.  Neumann: renormalise over the fluid
// parents (constants are interpolated exactly next to walls).  Dirichlet: a
// masked parent IS the wall value 0, so the plain bilinear weights stand.
__device__ __forceinline__ double w16_of(uint8_t c, int dirichlet) {
    if (dirichlet) return 16.0;
    return (double)(9 + ((c & NB_PJ) ? 3 : 0) + ((c & NB_PI) ? 3 : 0) + ((c & NB_PJI) ? 1 : 0));
}

__device__ __forceinline__ void coarse_idx(const CoarseView &V, int J, int I, long &c, long &w,
                                           long &e, long &s, long &n) {
    c = (long)(J + 1) * V.pitch + I + 1;
    w = c - 1;
    e = c + 1;
    if (V.periodic) {
        if (I == 0) w = c + (V.nx - 1);
        if (I == V.nx - 1) e = c - (V.nx - 1);
    }
    s = c - V.pitch;
    n = c + V.pitch;
    if (V.periodic_y) {
        if (J == 0) s = c + (long)(V.ny - 1) * V.pitch;
        if (J == V.ny - 1) n = c - (long)(V.ny - 1) * V.pitch;
    }
}

// parents of a fine cell (logical j,i): own aggregate (J0,I0) and the next
// nearest in each direction; returns false for a parent outside the grid
__device__ __forceinline__ void parents(int j, int i, int &J0, int &Jn, int &I0, int &In, int pj_off = 0) {
    J0 = (j >> 1) + pj_off;
    Jn = J0 + ((j & 1) ? 1 : -1);
    I0 = i >> 1;
    In = I0 + ((i & 1) ? 1 : -1);
}

// -------------------------------------------------------------- fine level --
template <bool ZERO_GUESS>
__global__ void __launch_bounds__(256)
k_smooth0(FineView F, double *__restrict__ x, const double *__restrict__ f, double fscale, int color) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.nx || j >= F.ny || ((i + j) & 1) != color) return;
    long idx;
    if (!fine_index(F, j, i, idx)) return;
    uint8_t c = F.nb[idx];
    if (!(c & NB_SELF)) return;
    Stencil s = fine_stencil(F, j, i, idx, c);
    if (s.diag <= 0.0) return;
    double a = ZERO_GUESS ? 0.0 : fine_offdiag(s, c, x);
    x[idx] = (fscale * f[idx] + a) / s.diag;
}

// r~ = (f - L x) / W16   (pre-scaled for the R = P^T gather)
__global__ void __launch_bounds__(256)
k_resid0(FineView F, const double *__restrict__ x, const double *__restrict__ f, double fscale,
         double *__restrict__ r) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.nx || j >= F.ny) return;
    long idx;
    if (!fine_index(F, j, i, idx)) return;
    uint8_t c = F.nb[idx];
    if (!(c & NB_SELF)) return;
    Stencil s = fine_stencil(F, j, i, idx, c);
    double res = fscale * f[idx] - (s.diag * x[idx] - fine_offdiag(s, c, x));
    r[idx] = res / w16_of(c, F.dirichlet);
}

// b_c = sum over the 4x4 fine cells around the aggregate of wy*wx*r~
__global__ void __launch_bounds__(256)
k_restrict0(FineView F, const double *__restrict__ r, CoarseView C, CT *__restrict__ bc) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= C.nx || J >= C.ny) return;
    long cc = (long)(J + 1) * C.pitch + I + 1;
    if (!(C.code[cc] & NB_SELF)) { bc[cc] = 0.0; return; }
    double acc = 0.0;
#pragma unroll
    for (int a = -1; a <= 2; a++) {
        int j = 2 * J + a;
        if (F.periodic_y) { if (j < 0) j += F.ny; else if (j >= F.ny) j -= F.ny; }
        if (j < 0 || j >= F.ny) continue;
        double wy = (a == 0 || a == 1) ? 3.0 : 1.0;
#pragma unroll
        for (int b = -1; b <= 2; b++) {
            int i = 2 * I + b;
            if (F.periodic) { if (i < 0) i += F.nx; else if (i >= F.nx) i -= F.nx; }
            else if (i < 0 || i >= F.nx) continue;
            double wx = (b == 0 || b == 1) ? 3.0 : 1.0;
            long idx;
            if (!fine_index(F, j, i, idx)) continue;
            if (F.nb[idx] & NB_SELF) acc += wy * wx * r[idx];
        }
    }
    bc[cc] = acc;
}

// x_f += (9 x00 + 3 xn0 + 3 x0n + xnn) / W16
__global__ void __launch_bounds__(256)
k_prolong0(FineView F, double *__restrict__ x, CoarseView C, const CT *__restrict__ xc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.nx || j >= F.ny) return;
    long idx;
    if (!fine_index(F, j, i, idx)) return;
    uint8_t c = F.nb[idx];
    if (!(c & NB_SELF)) return;
    int J0, Jn, I0, In;
    parents(j, i, J0, Jn, I0, In);
    if (C.periodic) { if (In < 0) In += C.nx; else if (In >= C.nx) In -= C.nx; }
    if (C.periodic_y) { if (Jn < 0) Jn += C.ny; else if (Jn >= C.ny) Jn -= C.ny; }
    long r0 = (long)(J0 + 1) * C.pitch, rn = (long)(Jn + 1) * C.pitch;   // halo rows absorb Jn=-1, ny
    double v = 9.0 * xc[r0 + I0 + 1];
    if (c & NB_PJ) v += 3.0 * xc[rn + I0 + 1];
    if (c & NB_PI) v += 3.0 * xc[r0 + In + 1];
    if (c & NB_PJI) v += xc[rn + In + 1];
    x[idx] += v / w16_of(c, F.dirichlet);
}

// 2-D grid-stride walk over the fine window (for kernels with reductions):
// blockIdx.y strides over rows, threads over columns; no integer divisions
#define FINE_LOOP(F)                                                                      \
    for (int j = blockIdx.y; j < (F).ny; j += gridDim.y)                                  \
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (F).nx; i += gridDim.x * blockDim.x)

// r = f - L x ; rr = r.r ; ff = f.f ; sum = 1.r      (CG start / convergence check)
__global__ void __launch_bounds__(256)
k_cg_resid(FineView F, const double *__restrict__ x, const double *__restrict__ f, double fscale,
           double *__restrict__ r, double *part, unsigned int *count, double *out) {
    double v[3] = {0.0, 0.0, 0.0};
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        uint8_t c = F.nb[idx];
        if (!(c & NB_SELF)) continue;
        Stencil s = fine_stencil(F, j, i, idx, c);
        double ff = fscale * f[idx];
        double res = ff - (s.diag * x[idx] - fine_offdiag(s, c, x));
        if (r) r[idx] = res;
        if (j < F.jo0 || j >= F.jo1) continue;
        v[0] += res * res;
        v[1] += res;
        v[2] += ff * ff;
    }
    grid_reduce<OpSum, 3>(v, part, count, out);
}

// r -= mean(r): the all-Neumann operator is singular (constants); a right-hand
// side that is not orthogonal to them (e.g. a divergence that is pure rounding
// noise) has no solution, so CG runs on its projection.  Every later residual
// stays orthogonal because the columns of L sum to zero.
__global__ void __launch_bounds__(256)
k_cg_project(FineView F, double *__restrict__ r, const double *__restrict__ scal, double inv_n) {
    double mean = scal[S_SUMR] * inv_n;
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        if (!(F.nb[idx] & NB_SELF)) continue;
        r[idx] -= mean;
    }
}

// Several connected fluid components (enclosed lakes, disconnected basins): the
// all-Neumann operator has one constant per component in its null space
// (elliptic.py:186-190: the diagonal is minus the sum of the existing
// off-diagonals, component by component).  v <- v - mean_c(v) is then done
// explicitly, in two passes: the per-component sums ...
struct CompMeans { double inv_n[Multigrid::MAXCOMP]; };

template <typename T>
__global__ void __launch_bounds__(256)
k_comp_sums(FineView F, const T *__restrict__ v, const uint8_t *__restrict__ comp, double *part,
            unsigned int *count, double *out) {
    double a[Multigrid::MAXCOMP];
#pragma unroll
    for (int k = 0; k < Multigrid::MAXCOMP; k++) a[k] = 0.0;
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        if (j < F.jo0 || j >= F.jo1) continue;
        const int cid = comp[idx];
        if (cid >= Multigrid::MAXCOMP) continue;
        const double x = (double)v[idx];
#pragma unroll
        for (int k = 0; k < Multigrid::MAXCOMP; k++) a[k] += (cid == k) ? x : 0.0;
    }
    grid_reduce<OpSum, Multigrid::MAXCOMP>(a, part, count, out);
}

// ... and the subtraction.  rr_slot >= 0: the squared norm kept in that slot loses
// sum_c (sum_c v)^2 / N_c, i.e. becomes the norm of the projected vector.
template <typename T>
__global__ void __launch_bounds__(256)
k_comp_sub(FineView F, T *__restrict__ v, const uint8_t *__restrict__ comp, double *__restrict__ scal,
           CompMeans M, int rr_slot) {
    __shared__ double mean[Multigrid::MAXCOMP];
    if (threadIdx.x < Multigrid::MAXCOMP && threadIdx.y == 0) mean[threadIdx.x] = scal[S_COMP + threadIdx.x] * M.inv_n[threadIdx.x];
    __syncthreads();
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        const int cid = comp[idx];
        if (cid >= Multigrid::MAXCOMP) continue;
        v[idx] = (T)((double)v[idx] - mean[cid]);
    }
    if (rr_slot >= 0 && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && threadIdx.y == 0) {
        double d = 0.0;
        for (int k = 0; k < Multigrid::MAXCOMP; k++) d += scal[S_COMP + k] * scal[S_COMP + k] * M.inv_n[k];
        scal[rr_slot] = fmax(scal[rr_slot] - d, 0.0);
    }
}

// q = L p ; pq = p.q
__global__ void __launch_bounds__(256)
k_cg_apply(FineView F, const double *__restrict__ p, double *__restrict__ q, double *part,
           unsigned int *count, double *out) {
    double v[1] = {0.0};
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        uint8_t c = F.nb[idx];
        if (!(c & NB_SELF)) continue;
        Stencil s = fine_stencil(F, j, i, idx, c);
        double pv = p[idx];
        double qv = s.diag * pv - fine_offdiag(s, c, p);
        q[idx] = qv;
        v[0] += pv * qv;
    }
    grid_reduce<OpSum, 1>(v, part, count, out);
}

// x += alpha p ; r -= alpha q ; (rr, sum r)          alpha = rz / pq
__global__ void __launch_bounds__(256)
k_cg_update(FineView F, double *__restrict__ x, double *__restrict__ r,
            const double *__restrict__ p, const double *__restrict__ q,
            const double *__restrict__ scal, int rz_slot, double *part, unsigned int *count,
            double *out) {
    double pq = scal[S_PQ];
    double alpha = pq != 0.0 ? scal[rz_slot] / pq : 0.0;
    double v[2] = {0.0, 0.0};
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        if (!(F.nb[idx] & NB_SELF)) continue;
        x[idx] += alpha * p[idx];
        double rv = r[idx] - alpha * q[idx];
        r[idx] = rv;
        if (j < F.jo0 || j >= F.jo1) continue;
        v[0] += rv * rv;
        v[1] += rv;
    }
    grid_reduce<OpSum, 2>(v, part, count, out);
}

// (sum (r - mean r) z, sum z) over the unknowns   (un-fused path; the fused up
// leg produces the same two numbers)
__global__ void __launch_bounds__(256)
k_dot2(FineView F, const double *__restrict__ r, const double *__restrict__ z,
       const double *__restrict__ scal, int sumr_slot, double inv_n, double *part,
       unsigned int *count, double *out) {
    double mr = sumr_slot >= 0 ? scal[sumr_slot] * inv_n : 0.0;
    double v[2] = {0.0, 0.0};
    FINE_LOOP(F) {
        long idx;
        if (!fine_index(F, j, i, idx)) continue;
        if (!(F.nb[idx] & NB_SELF)) continue;
        if (j < F.jo0 || j >= F.jo1) continue;
        v[0] += (r[idx] - mr) * z[idx];
        v[1] += z[idx];
    }
    grid_reduce<OpSum, 2>(v, part, count, out);
}

// pnew = (z - mean z) + beta pold ;  q = L pnew ;  pq = pnew.q
//   beta = rz_new / rz_old (0 on the first iteration); rz_new is stored in the
//   slot of this iteration's parity for k_cg_update and the next direction.
//   pold / pnew are distinct buffers because q needs pnew at the neighbours.
//   singular operators: z is projected on the complement of the constants, so
//   that rounding in the V-cycle cannot feed the null space.
//   A CTA forms pnew once per point of a (64+2) x (16+2) window in shared
//   memory and applies the 5-point operator from there.
#ifndef F2D_CGX
#define F2D_CGX 64
#define F2D_CGY 16
#endif
constexpr int CGX = F2D_CGX, CGY = F2D_CGY;   // tile of the CG vector kernels: 256 threads as CGX x CGTY, CGY / CGTY rows each
constexpr int CGTY = 256 / CGX;
static_assert(CGX % 32 == 0 && 256 % CGX == 0 && CGY % CGTY == 0, "CG tile shape");

// the two expressions both paths of k_cg_dir_apply evaluate, with explicit roundings
__device__ __forceinline__ double cg_pnew(double z, double mz, double beta, double pold) {
    return __fma_rn(beta, pold, __dsub_rn(z, mz));
}
__device__ __forceinline__ double cg_q(const FineView &F, double diag, double pc, double w, double e, double s, double n) {
    return __fma_rn(diag, pc, -__fma_rn(F.cx, __dadd_rn(w, e), __dmul_rn(F.cy, __dadd_rn(s, n))));
}
__device__ __forceinline__ double cg_diag(const FineView &F, double cw, double ce, double cs, double cn) {
    return F.dirichlet ? __fma_rn(2.0, __dadd_rn(F.cx, F.cy), F.shift)
                       : __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(cw, ce), cs), cn), F.shift);
}

// The search direction is STORED in fp32 (the preconditioned residual z already is):
//   pnew = fl32((z - mean z) + beta pold)
// and q = L pnew is never stored: k_cg_dir_apply only needs its dot product with pnew,
// the update kernel re-applies the 5-point operator to the fp32 window it reads anyway.
// L pnew itself is evaluated in fp64 in both (for smooth p it is a difference of nearly
// equal numbers: in fp32 the cancellation left p.q -- and with it alpha -- without a
// single correct digit on a 4096^2 grid, and CG diverged).  Per fine
// point the two CG kernels move 13 + 37 bytes instead of 29 + 49.  x and r stay fp64 and
// are updated with the SAME rounded direction (x += alpha p, r -= alpha L p with L p in
// fp64), so r remains the residual of x to fp64 rounding whatever alpha is; the rounding
// of p and alpha only perturbs the conjugacy of successive directions at the 1e-7 level,
// far below what the 4-5 iterations of a solve can see.
//
//   tile_open[tile] != 0 (k_cg_tile_flags): the whole window lies inside the array (its W / E
//   halo column may be the periodic image) and holds unknowns only -> no mask bytes, no bounds tests, NO shared
//   memory and no barrier: a thread owns 4 consecutive rows of one column, issues all its
//   loads first, and gets its W / E neighbours from the adjacent lanes (the two edge lanes
//   of a warp load theirs).  Both paths evaluate the same expressions with the same
//   thread-to-row map: same bits, whichever tiles are open.
// flat offset of the W neighbour of a tile's first column / the E neighbour of its last one
__device__ __forceinline__ void cg_side_offsets(const FineView &F, int i0, long &woff, long &eoff) {
    woff = (F.periodic && i0 == 0) ? (long)(F.nx - 1) : -1L;
    eoff = (F.periodic && i0 + CGX == F.nx) ? -(long)(F.nx - 1) : 1L;
}
struct CgF32 { float mz, beta; };
template <typename TZ>
__device__ __forceinline__ float cg_pnew32(TZ z, float pold, const CgF32 &K, double mz, double beta) {
    if constexpr (sizeof(TZ) == 4) return __fmaf_rn(K.beta, pold, __fsub_rn((float)z, K.mz));
    else return (float)cg_pnew((double)z, mz, beta, (double)pold);      // un-fused cross-check path: z is fp64
}

template <typename TZ>
__global__ void __launch_bounds__(256, 6)
k_cg_dir_apply(FineView F, const TZ *__restrict__ z, const float *__restrict__ pold,
               float *__restrict__ pnew, double *__restrict__ scal, int it,
               int singular, double inv_n, double *part, unsigned int *count,
               const uint8_t *__restrict__ tile_open) {
    __shared__ float sp[CGY + 2][CGX + 2];
    double rznew = scal[S_RZNEW];
    double mz = singular ? scal[S_SUMZ] * inv_n : 0.0;
    double beta = 0.0;
    if (it > 0) { double rzold = scal[S_RZ0 + ((it - 1) & 1)]; beta = rzold != 0.0 ? rznew / rzold : 0.0; }
    const CgF32 K{(float)mz, (float)beta};
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int ntx = (F.nx + CGX - 1) / CGX, nty = (F.ny + CGY - 1) / CGY, ntiles = ntx * nty;
    const int stride = gridDim.x * gridDim.y;
    const double diag_open = cg_diag(F, F.cx, F.cx, F.cy, F.cy);
    constexpr int R = CGY / CGTY;
    double v[1] = {0.0};
    int tile = blockIdx.y * gridDim.x + blockIdx.x;
    uint8_t flag = (tile_open != nullptr && tile < ntiles) ? tile_open[tile] : 0;
    for (; tile < ntiles; tile += stride) {
        const int i0 = (tile % ntx) * CGX, j0 = (tile / ntx) * CGY;
        const bool open = flag != 0;                                    // block-uniform
        if (tile_open != nullptr && tile + stride < ntiles) flag = tile_open[tile + stride];   // in flight during this tile
        if (open) {
            const int lane = threadIdx.x & 31;
            const long idx0 = (long)(F.oj + j0 + R * threadIdx.y - 1) * F.n1 + F.oi + i0 + threadIdx.x;   // row above my first
            TZ zv[R + 2], zw[R];
            float pv[R + 2], pw[R];
#pragma unroll
            for (int u = 0; u < R + 2; u++) { zv[u] = z[idx0 + (long)u * F.n1]; pv[u] = pold[idx0 + (long)u * F.n1]; }
            long woff, eoff;
            cg_side_offsets(F, i0, woff, eoff);
            const long side = lane == 0 ? (threadIdx.x == 0 ? woff : -1L) : (threadIdx.x == CGX - 1 ? eoff : 1L);
            if (lane == 0 || lane == 31) {
#pragma unroll
                for (int u = 0; u < R; u++) {
                    zw[u] = z[idx0 + (long)(u + 1) * F.n1 + side];
                    pw[u] = pold[idx0 + (long)(u + 1) * F.n1 + side];
                }
            }
            float pn[R + 2];
#pragma unroll
            for (int u = 0; u < R + 2; u++) pn[u] = cg_pnew32<TZ>(zv[u], pv[u], K, mz, beta);
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int j = j0 + R * threadIdx.y + u;
                const float pf = pn[u + 1];
                float w = __shfl_up_sync(0xffffffffu, pf, 1), e = __shfl_down_sync(0xffffffffu, pf, 1);
                if (lane == 0) w = cg_pnew32<TZ>(zw[u], pw[u], K, mz, beta);
                if (lane == 31) e = cg_pnew32<TZ>(zw[u], pw[u], K, mz, beta);
                pnew[idx0 + (long)(u + 1) * F.n1] = pf;
                const double pc = (double)pf;
                const double qv = cg_q(F, diag_open, pc, (double)w, (double)e, (double)pn[u], (double)pn[u + 2]);
                if (j >= F.jo0 && j < F.jo1) v[0] = __fma_rn(pc, qv, v[0]);
            }
            continue;
        }
        __syncthreads();        // the previous masked tile of this CTA is done with sp
        for (int t = tid; t < (CGY + 2) * (CGX + 2); t += 256) {
            int a = t / (CGX + 2), b = t - a * (CGX + 2);
            int j = j0 - 1 + a, i = i0 - 1 + b;
            if (F.periodic) { if (i < 0) i += F.nx; else if (i >= F.nx) i -= F.nx; }
            if (F.periodic_y) { if (j < 0) j += F.ny; else if (j >= F.ny) j -= F.ny; }
            float pv = 0.0f;
            long idx;
            if (j >= 0 && j < F.ny && i >= 0 && i < F.nx && fine_index(F, j, i, idx) && (F.nb[idx] & NB_SELF))
                pv = cg_pnew32<TZ>(z[idx], pold[idx], K, mz, beta);
            sp[a][b] = pv;
        }
        __syncthreads();
        const int i = i0 + threadIdx.x;
#pragma unroll
        for (int r = 0; r < R; r++) {
            const int a = 1 + R * threadIdx.y + r, b = 1 + threadIdx.x, j = j0 + R * threadIdx.y + r;   // the open path's thread-to-row map
            long idx;
            if (i >= F.nx || j >= F.ny || !fine_index(F, j, i, idx)) continue;
            uint8_t c = F.nb[idx];
            if (!(c & NB_SELF)) continue;
            double cw = (c & NB_W) ? F.cx : 0.0, ce = (c & NB_E) ? F.cx : 0.0;
            double cs = (c & NB_S) ? F.cy : 0.0, cn = (c & NB_N) ? F.cy : 0.0;
            const double diag = cg_diag(F, cw, ce, cs, cn);
            const float pf = sp[a][b];
            const double pc = (double)pf;
            pnew[idx] = pf;
            // closed faces lead to points that are not unknowns: their sp entry is 0
            const double qv = cg_q(F, diag, pc, (double)sp[a][b - 1], (double)sp[a][b + 1], (double)sp[a - 1][b], (double)sp[a + 1][b]);
            if (j >= F.jo0 && j < F.jo1) v[0] = __fma_rn(pc, qv, v[0]);
        }
    }
    grid_reduce<OpSum, 1>(v, part, count, scal + S_PQ);
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) scal[S_RZ0 + (it & 1)] = rznew;
}

// x += alpha p ; r -= alpha L p ; (rr, sum r)          alpha = rz / pq.   L p in fp64 from
// the fp32 direction; same tiles and thread-to-row map as k_cg_dir_apply.
#ifndef F2D_UPD_CTAS
#define F2D_UPD_CTAS 4
#endif
__global__ void __launch_bounds__(256, F2D_UPD_CTAS)      // 4 -> 64 registers: at 40 (6 CTAs) the loaded window was spilled as it arrived
k_cg_update_p(FineView F, double *__restrict__ x, double *__restrict__ r, const float *__restrict__ p,
              const double *__restrict__ scal, int rz_slot, double *part, unsigned int *count, double *out,
              const uint8_t *__restrict__ tile_open) {
    // half-height tiles (2 rows per thread): 0.123 ms against 0.141 with the 16-row tiles of k_cg_dir_apply
    // (fewer registers in flight per thread, more warps resident); the open flags are those of the 16-row tiles
    constexpr int UY = CGY / 2;
    static_assert(UY % CGTY == 0, "update tile");
    __shared__ float sp[UY + 2][CGX + 2];
    const double pq = scal[S_PQ];
    const double alpha = pq != 0.0 ? scal[rz_slot] / pq : 0.0;
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int ntx = (F.nx + CGX - 1) / CGX, nty = (F.ny + UY - 1) / UY, ntiles = ntx * nty;
    const int stride = gridDim.x * gridDim.y;
    const double diag_open = cg_diag(F, F.cx, F.cx, F.cy, F.cy);
    constexpr int R = UY / CGTY;
    double v[2] = {0.0, 0.0};
    int tile = blockIdx.y * gridDim.x + blockIdx.x;
    auto flag_of = [&](int t) { return tile_open[((t / ntx) * UY / CGY) * ntx + t % ntx]; };   // the 16-row tile this half tile lies in
    uint8_t flag = (tile_open != nullptr && tile < ntiles) ? flag_of(tile) : 0;
    for (; tile < ntiles; tile += stride) {
        const int i0 = (tile % ntx) * CGX, j0 = (tile / ntx) * UY;
        const bool open = flag != 0;                                    // block-uniform
        if (tile_open != nullptr && tile + stride < ntiles) flag = flag_of(tile + stride);   // in flight during this tile
        if (open) {
            const int lane = threadIdx.x & 31;
            const long idx0 = (long)(F.oj + j0 + R * threadIdx.y - 1) * F.n1 + F.oi + i0 + threadIdx.x;   // row above my first
            float pn[R + 2], pw[R];
            double xv[R], rv[R];
#pragma unroll
            for (int u = 0; u < R + 2; u++) pn[u] = p[idx0 + (long)u * F.n1];
#pragma unroll
            for (int u = 0; u < R; u++) { xv[u] = x[idx0 + (long)(u + 1) * F.n1]; rv[u] = r[idx0 + (long)(u + 1) * F.n1]; }
            long woff, eoff;
            cg_side_offsets(F, i0, woff, eoff);
            const long side = lane == 0 ? (threadIdx.x == 0 ? woff : -1L) : (threadIdx.x == CGX - 1 ? eoff : 1L);
            if (lane == 0 || lane == 31) {
#pragma unroll
                for (int u = 0; u < R; u++) pw[u] = p[idx0 + (long)(u + 1) * F.n1 + side];
            }
#pragma unroll
            for (int u = 0; u < R; u++) {
                const int j = j0 + R * threadIdx.y + u;
                const float pf = pn[u + 1];
                float w = __shfl_up_sync(0xffffffffu, pf, 1), e = __shfl_down_sync(0xffffffffu, pf, 1);
                if (lane == 0) w = pw[u];
                if (lane == 31) e = pw[u];
                const double pc = (double)pf;
                const double qv = cg_q(F, diag_open, pc, (double)w, (double)e, (double)pn[u], (double)pn[u + 2]);
                const long idx = idx0 + (long)(u + 1) * F.n1;
                x[idx] = __fma_rn(alpha, pc, xv[u]);
                const double rn = __fma_rn(-alpha, qv, rv[u]);
                r[idx] = rn;
                if (j >= F.jo0 && j < F.jo1) { v[0] = __fma_rn(rn, rn, v[0]); v[1] += rn; }
            }
            continue;
        }
        __syncthreads();        // the previous masked tile of this CTA is done with sp
        for (int t = tid; t < (UY + 2) * (CGX + 2); t += 256) {
            int a = t / (CGX + 2), b = t - a * (CGX + 2);
            int j = j0 - 1 + a, i = i0 - 1 + b;
            if (F.periodic) { if (i < 0) i += F.nx; else if (i >= F.nx) i -= F.nx; }
            if (F.periodic_y) { if (j < 0) j += F.ny; else if (j >= F.ny) j -= F.ny; }
            float pv = 0.0f;
            long idx;
            if (j >= 0 && j < F.ny && i >= 0 && i < F.nx && fine_index(F, j, i, idx) && (F.nb[idx] & NB_SELF)) pv = p[idx];
            sp[a][b] = pv;
        }
        __syncthreads();
        const int i = i0 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < R; q++) {
            const int a = 1 + R * threadIdx.y + q, b = 1 + threadIdx.x, j = j0 + R * threadIdx.y + q;   // the open path's thread-to-row map
            long idx;
            if (i >= F.nx || j >= F.ny || !fine_index(F, j, i, idx)) continue;
            uint8_t c = F.nb[idx];
            if (!(c & NB_SELF)) continue;
            double cw = (c & NB_W) ? F.cx : 0.0, ce = (c & NB_E) ? F.cx : 0.0;
            double cs = (c & NB_S) ? F.cy : 0.0, cn = (c & NB_N) ? F.cy : 0.0;
            const double diag = cg_diag(F, cw, ce, cs, cn);
            const double pc = (double)sp[a][b];
            const double qv = cg_q(F, diag, pc, (double)sp[a][b - 1], (double)sp[a][b + 1], (double)sp[a - 1][b], (double)sp[a + 1][b]);
            x[idx] = __fma_rn(alpha, pc, x[idx]);
            const double rn = __fma_rn(-alpha, qv, r[idx]);
            r[idx] = rn;
            if (j >= F.jo0 && j < F.jo1) { v[0] = __fma_rn(rn, rn, v[0]); v[1] += rn; }
        }
    }
    grid_reduce<OpSum, 2>(v, part, count, out);
}

// x = sum_k w_k g_k (the first guess, extrapolated from the history of this stage's
// solutions) and r = f - L x, rr, sum r, ff in ONE pass: the guess is formed once per
// point of the (64+2) x (16+2) window in shared memory instead of being written by one
// kernel and read back (with its four neighbours) by the next.  x must not alias a g_k.
struct GuessW { const double *g[6]; double w[6]; int n; };
__global__ void __launch_bounds__(256, 4)
k_cg_resid_guess(FineView F, GuessW G, double *__restrict__ x, const double *__restrict__ f, double fscale,
                 double *__restrict__ r, double *part, unsigned int *count, double *out,
                 const uint8_t *__restrict__ tile_open) {
    __shared__ double sp[CGY + 2][CGX + 2];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int ntx = (F.nx + CGX - 1) / CGX, nty = (F.ny + CGY - 1) / CGY;
    const double diag_open = cg_diag(F, F.cx, F.cx, F.cy, F.cy);
    double v[3] = {0.0, 0.0, 0.0};
    auto guess_at = [&](long idx) {
        double a = G.w[0] * G.g[0][idx];
#pragma unroll
        for (int m = 1; m < 6; m++)
            if (m < G.n) a += G.w[m] * G.g[m][idx];
        return a;
    };
    for (int tile = blockIdx.y * gridDim.x + blockIdx.x; tile < ntx * nty; tile += gridDim.x * gridDim.y) {
        const int i0 = (tile % ntx) * CGX, j0 = (tile / ntx) * CGY;
        const bool open = tile_open != nullptr && tile_open[tile];     // block-uniform
        __syncthreads();
        if (open) {
            const long base = (long)(F.oj + j0 - 1) * F.n1 + F.oi + i0 - 1;
            long woff, eoff;
            cg_side_offsets(F, i0, woff, eoff);
            for (int t = tid; t < (CGY + 2) * (CGX + 2); t += 256) {
                int a = t / (CGX + 2), b = t - a * (CGX + 2);
                long col = b == 0 ? 1 + woff : (b == CGX + 1 ? CGX + eoff : (long)b);     // halo columns may wrap
                (&sp[0][0])[t] = guess_at(base + (long)a * F.n1 + col);
            }
            __syncthreads();
            const long idx0 = base + (long)(1 + threadIdx.y) * F.n1 + 1 + threadIdx.x;
#pragma unroll
            for (int q = 0; q < CGY / CGTY; q++) {
                const int a = 1 + threadIdx.y + CGTY * q, b = 1 + threadIdx.x, j = j0 + threadIdx.y + CGTY * q;
                const long idx = idx0 + (long)(CGTY * q) * F.n1;
                const double xc = sp[a][b];
                const double ff = fscale * f[idx];
                const double res = ff - cg_q(F, diag_open, xc, sp[a][b - 1], sp[a][b + 1], sp[a - 1][b], sp[a + 1][b]);
                x[idx] = xc;
                r[idx] = res;
                if (j >= F.jo0 && j < F.jo1) { v[0] += res * res; v[1] += res; v[2] += ff * ff; }
            }
            continue;
        }
        for (int t = tid; t < (CGY + 2) * (CGX + 2); t += 256) {
            int a = t / (CGX + 2), b = t - a * (CGX + 2);
            int j = j0 - 1 + a, i = i0 - 1 + b;
            if (F.periodic) { if (i < 0) i += F.nx; else if (i >= F.nx) i -= F.nx; }
            if (F.periodic_y) { if (j < 0) j += F.ny; else if (j >= F.ny) j -= F.ny; }
            double xv = 0.0;
            long idx;
            if (j >= 0 && j < F.ny && i >= 0 && i < F.nx && fine_index(F, j, i, idx) && (F.nb[idx] & NB_SELF)) xv = guess_at(idx);
            sp[a][b] = xv;
        }
        __syncthreads();
        const int i = i0 + threadIdx.x;
#pragma unroll
        for (int q = 0; q < CGY / CGTY; q++) {
            const int a = 1 + threadIdx.y + CGTY * q, b = 1 + threadIdx.x, j = j0 + threadIdx.y + CGTY * q;
            long idx;
            if (i >= F.nx || j >= F.ny || !fine_index(F, j, i, idx)) continue;
            uint8_t c = F.nb[idx];
            if (!(c & NB_SELF)) continue;
            double cw = (c & NB_W) ? F.cx : 0.0, ce = (c & NB_E) ? F.cx : 0.0;
            double cs = (c & NB_S) ? F.cy : 0.0, cn = (c & NB_N) ? F.cy : 0.0;
            const double diag = cg_diag(F, cw, ce, cs, cn);
            const double xc = sp[a][b];
            const double ff = fscale * f[idx];
            // closed faces lead to points that are not unknowns: their sp entry is 0
            const double res = ff - cg_q(F, diag, xc, sp[a][b - 1], sp[a][b + 1], sp[a - 1][b], sp[a + 1][b]);
            x[idx] = xc;
            r[idx] = res;
            if (j >= F.jo0 && j < F.jo1) { v[0] += res * res; v[1] += res; v[2] += ff * ff; }
        }
    }
    grid_reduce<OpSum, 3>(v, part, count, out);
}

// which tiles of k_cg_dir_apply are open: one CTA per tile
__global__ void __launch_bounds__(256) k_cg_tile_flags(FineView F, uint8_t *__restrict__ flags) {
    const int ntx = (F.nx + CGX - 1) / CGX;
    const int tile = blockIdx.x;
    const int i0 = (tile % ntx) * CGX, j0 = (tile / ntx) * CGY;
    // x-periodic: the first / last tile of a row takes its W / E halo column from the other end
    // of the window (cg_side_offsets); everything else must lie inside window and array
    const bool wl = F.periodic && i0 == 0, wr = F.periodic && i0 + CGX == F.nx;
    int good = j0 - 1 >= 0 && j0 + CGY + 1 <= F.ny && (wl || i0 - 1 >= 0) && (wr || i0 + CGX + 1 <= F.nx) && i0 + CGX <= F.nx &&
               F.oj + j0 - 1 >= 0 && F.oj + j0 + CGY + 1 <= F.n2 && F.oi + i0 - (wl ? 0 : 1) >= 0 &&
               F.oi + i0 + CGX + (wr ? 0 : 1) <= F.n1;
    if (good) {
        const long row0 = (long)(F.oj + j0 - 1) * F.n1 + F.oi;
        for (int t = threadIdx.x; t < (CGY + 2) * (CGX + 2); t += blockDim.x) {
            int a = t / (CGX + 2), b = t - a * (CGX + 2);
            int i = i0 - 1 + b;
            if (i < 0) i += F.nx; else if (i >= F.nx) i -= F.nx;
            if ((F.nb[row0 + (long)a * F.n1 + i] & 31) != 31) good = 0;    // an unknown with four open faces
        }
    }
    good = __syncthreads_and(good);
    if (threadIdx.x == 0) flags[tile] = (uint8_t)good;
}

// y = A x = -L x on the unknowns, 0 elsewhere inside the window
__global__ void __launch_bounds__(256)
k_apply_A(FineView F, const double *__restrict__ x, double *__restrict__ y) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.nx || j >= F.ny) return;
    long idx;
    if (!fine_index(F, j, i, idx)) return;
    uint8_t c = F.nb[idx];
    if (!(c & NB_SELF)) return;
    Stencil s = fine_stencil(F, j, i, idx, c);
    y[idx] = fine_offdiag(s, c, x) - s.diag * x[idx];
}

// ------------------------------------------------------------ coarse levels --
template <bool ZERO_GUESS>
__global__ void __launch_bounds__(256)
k_smooth(CoarseView V, CT *__restrict__ x, const CT *__restrict__ b, int color) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= V.nx || J >= V.ny || ((I + J) & 1) != color) return;
    long c, w, e, s, n;
    coarse_idx(V, J, I, c, w, e, s, n);
    double di = V.dinv[c];
    if (di == 0.0) return;
    double a = 0.0;
    if (!ZERO_GUESS) a = V.cx[c] * x[w] + V.cx[e] * x[e] + V.cy[c] * x[s] + V.cy[n] * x[n];
    x[c] = (b[c] + a) * di;
}

__global__ void __launch_bounds__(256)
k_resid(CoarseView V, const CT *__restrict__ x, const CT *__restrict__ b,
        CT *__restrict__ r) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= V.nx || J >= V.ny) return;
    long c, w, e, s, n;
    coarse_idx(V, J, I, c, w, e, s, n);
    double di = V.dinv[c];
    if (di == 0.0) { r[c] = 0.0; return; }
    double a = V.cx[c] * x[w] + V.cx[e] * x[e] + V.cy[c] * x[s] + V.cy[n] * x[n];
    double res = b[c] - (x[c] / di - a);
    r[c] = res / w16_of(V.code[c], V.dirichlet);
}

__global__ void __launch_bounds__(256)
k_restrict(CoarseView Vf, const CT *__restrict__ r, CoarseView C, CT *__restrict__ bc) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= C.nx || J >= C.ny) return;
    long cc = (long)(J + 1) * C.pitch + I + 1;
    if (!(C.code[cc] & NB_SELF)) { bc[cc] = 0.0; return; }
    double acc = 0.0;
#pragma unroll
    for (int a = -1; a <= 2; a++) {
        int j = 2 * J + a;
        if (Vf.periodic_y) { if (j < 0) j += Vf.ny; else if (j >= Vf.ny) j -= Vf.ny; }
        if (j < 0 || j >= Vf.ny) continue;
        double wy = (a == 0 || a == 1) ? 3.0 : 1.0;
#pragma unroll
        for (int b = -1; b <= 2; b++) {
            int i = 2 * I + b;
            if (Vf.periodic) { if (i < 0) i += Vf.nx; else if (i >= Vf.nx) i -= Vf.nx; }
            else if (i < 0 || i >= Vf.nx) continue;
            double wx = (b == 0 || b == 1) ? 3.0 : 1.0;
            acc += wy * wx * r[(long)(j + 1) * Vf.pitch + i + 1];   // r~ is 0 off the unknowns
        }
    }
    bc[cc] = acc;
}

__global__ void __launch_bounds__(256)
k_prolong(CoarseView Vf, CT *__restrict__ x, CoarseView C, const CT *__restrict__ xc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Vf.nx || j >= Vf.ny) return;
    long idx = (long)(j + 1) * Vf.pitch + i + 1;
    uint8_t c = Vf.code[idx];
    if (!(c & NB_SELF)) return;
    int J0, Jn, I0, In;
    parents(j, i, J0, Jn, I0, In);
    if (C.periodic) { if (In < 0) In += C.nx; else if (In >= C.nx) In -= C.nx; }
    if (C.periodic_y) { if (Jn < 0) Jn += C.ny; else if (Jn >= C.ny) Jn -= C.ny; }
    long r0 = (long)(J0 + 1) * C.pitch, rn = (long)(Jn + 1) * C.pitch;
    double v = 9.0 * xc[r0 + I0 + 1];
    if (c & NB_PJ) v += 3.0 * xc[rn + I0 + 1];
    if (c & NB_PI) v += 3.0 * xc[r0 + In + 1];
    if (c & NB_PJI) v += xc[rn + In + 1];
    x[idx] += v / w16_of(c, Vf.dirichlet);
}

// coarsest level: nsw sweeps (R,B) then nsw sweeps (B,R) from a zero guess, one CTA
__global__ void __launch_bounds__(1024)
k_coarsest(CoarseView V, CT *__restrict__ x, const CT *__restrict__ b, int nsw) {
    int npts = V.ny * V.nx;
    for (int t = threadIdx.x; t < npts; t += blockDim.x) {
        int J = t / V.nx, I = t - J * V.nx;
        x[(long)(J + 1) * V.pitch + I + 1] = 0.0;
    }
    __syncthreads();
    for (int sweep = 0; sweep < 4 * nsw; sweep++) {
        int color = (sweep < 2 * nsw) ? (sweep & 1) : 1 - (sweep & 1);
        for (int t = threadIdx.x; t < npts; t += blockDim.x) {
            int J = t / V.nx, I = t - J * V.nx;
            if (((I + J) & 1) != color) continue;
            long c, w, e, s, n;
            coarse_idx(V, J, I, c, w, e, s, n);
            double di = V.dinv[c];
            if (di == 0.0) continue;
            double a = V.cx[c] * x[w] + V.cx[e] * x[e] + V.cy[c] * x[s] + V.cy[n] * x[n];
            x[c] = (b[c] + a) * di;
        }
        __syncthreads();
    }
}

// ----------------------------------------------------------------- set-up ---
// solver mask -> fine bits (elliptic.py:102-111, :144-165 neighbour rules)
__global__ void k_build_nb(FineView F, const int8_t *__restrict__ sm, uint8_t *__restrict__ nb) {
    int ai = blockIdx.x * blockDim.x + threadIdx.x;
    int aj = blockIdx.y * blockDim.y + threadIdx.y;
    if (ai >= F.n1 || aj >= F.n2) return;
    long idx = (long)aj * F.n1 + ai;
    uint8_t c = 0;
    int i = ai - F.oi;
    bool inwin = i >= 0 && i < F.nx;
    if (inwin && sm[idx]) {
        c = NB_SELF;
        long iw = idx - 1, ie = idx + 1;
        bool hw = ai > 0, he = ai < F.n1 - 1;
        if (F.periodic) {
            if (i == 0) iw = idx + (F.nx - 1);
            if (i == F.nx - 1) ie = idx - (F.nx - 1);
            hw = he = true;
        } else {
            hw = hw && i > 0;
            he = he && i < F.nx - 1;
        }
        if (hw && sm[iw]) c |= NB_W;
        if (he && sm[ie]) c |= NB_E;
        const int j = aj - F.oj;
        if (F.periodic_y && j == 0) { if (sm[idx + (long)(F.ny - 1) * F.n1]) c |= NB_S; }
        else if (aj > 0 && sm[idx - F.n1]) c |= NB_S;
        if (F.periodic_y && j == F.ny - 1) { if (sm[idx - (long)(F.ny - 1) * F.n1]) c |= NB_N; }
        else if (aj < F.n2 - 1 && sm[idx + F.n1]) c |= NB_N;
    }
    nb[idx] = c;
}

// level 0 -> level 1 coefficients
// Dirichlet wall couplings: the aggregate centre of level l sits (2^l+1)/2 fine
// spacings from the wall, so the coupling shrinks by (2^(l-1)+1)/(2^l+1) per
// level (2/3, 3/5, 5/9, ... -> 1/2) instead of the 1/2 of interior faces.
__global__ void k_coarsen0(FineView F, Level L1, int periodic, double wallfac, int pj_off) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= L1.nx || J >= L1.ny) return;
    long cc = (long)(J + 1) * L1.pitch + I + 1;
    double cx = 0, cy = 0, mass = 0, wall = 0;
    int fluid = 0;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            int j = 2 * (J - pj_off) + a, i = 2 * I + b;
            if (j < 0 || j >= F.ny || i >= F.nx) continue;
            long idx;
            if (!fine_index(F, j, i, idx)) continue;
            uint8_t c = F.nb[idx];
            if (!(c & NB_SELF)) continue;
            fluid = 1;
            if (b == 0 && (c & NB_W)) cx += F.cx;
            if (a == 0 && (c & NB_S)) cy += F.cy;
            mass += F.shift;
            if (F.dirichlet)
                wall += F.cx * (2 - ((c & NB_W) ? 1 : 0) - ((c & NB_E) ? 1 : 0)) +
                        F.cy * (2 - ((c & NB_S) ? 1 : 0) - ((c & NB_N) ? 1 : 0));
        }
    L1.cx[cc] = 0.5 * cx;
    L1.cy[cc] = 0.5 * cy;
    L1.mass[cc] = mass;
    L1.wall[cc] = wallfac * wall;
    L1.code[cc] = fluid ? NB_SELF : 0;
}

// level l -> l+1 coefficients (l >= 1)
__global__ void k_coarsen(Level Lf, Level Lc, int periodic, double wallfac, int pj_off) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= Lc.nx || J >= Lc.ny) return;
    long cc = (long)(J + 1) * Lc.pitch + I + 1;
    double cx = 0, cy = 0, mass = 0, wall = 0;
    int fluid = 0;
    for (int a = 0; a < 2; a++)
        for (int b = 0; b < 2; b++) {
            int j = 2 * (J - pj_off) + a, i = 2 * I + b;
            if (j < 0 || j >= Lf.ny || i >= Lf.nx) continue;
            long idx = (long)(j + 1) * Lf.pitch + i + 1;
            if (!(Lf.code[idx] & NB_SELF)) continue;
            fluid = 1;
            if (b == 0) cx += Lf.cx[idx];
            if (a == 0) cy += Lf.cy[idx];
            mass += Lf.mass[idx];
            wall += Lf.wall[idx];
        }
    Lc.cx[cc] = 0.5 * cx;
    Lc.cy[cc] = 0.5 * cy;
    Lc.mass[cc] = mass;
    Lc.wall[cc] = wallfac * wall;
    Lc.code[cc] = fluid ? NB_SELF : 0;
}

__global__ void k_build_dinv(Level L, int periodic) {      // periodic: bit 0 = x, bit 1 = y
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= L.nx || J >= L.ny) return;
    long c = (long)(J + 1) * L.pitch + I + 1;
    long e = c + 1, n = c + L.pitch;
    if ((periodic & 1) && I == L.nx - 1) e = c - (L.nx - 1);
    if ((periodic & 2) && J == L.ny - 1) n = c - (long)(L.ny - 1) * L.pitch;
    double d = 0.0;
    if (L.code[c] & NB_SELF) {
        double diag = L.cx[c] + L.cx[e] + L.cy[c] + L.cy[n] + L.mass[c] + L.wall[c];
        d = diag > 0.0 ? 1.0 / diag : 0.0;
    }
    L.dinv[c] = d;
    if (d == 0.0) L.code[c] = 0;
}

// NB_REG: the point is an unknown whose four couplings and inverse diagonal are
// the level's regular values (mg_tiles.cuh, open tiles).  Runs when every
// coefficient of the level, ghost rows included, is final.
__global__ void k_mark_regular(Level L, int periodic, CT cx0, CT cy0, CT dinv0) {
    int I = blockIdx.x * blockDim.x + threadIdx.x;
    int J = blockIdx.y * blockDim.y + threadIdx.y;
    if (I >= L.nx || J >= L.ny) return;
    long c = (long)(J + 1) * L.pitch + I + 1;
    long e = c + 1, n = c + L.pitch;
    if ((periodic & 1) && I == L.nx - 1) e = c - (L.nx - 1);
    else if (I == L.nx - 1) return;                      // closed domain: no regular east face
    if ((periodic & 2) && J == L.ny - 1) n = c - (long)(L.ny - 1) * L.pitch;
    uint8_t code = L.code[c];
    if (!(code & NB_SELF)) return;
    if (L.cx[c] == cx0 && L.cx[e] == cx0 && L.cy[c] == cy0 && L.cy[n] == cy0 && L.dinv[c] == dinv0)
        L.code[c] = code | NB_REG;
}

// regular coefficients of level l >= 1 (k_coarsen0 / k_coarsen / k_build_dinv away from
// walls and masks: couplings stay cx, cy; the mass term quadruples per level)
static void mark_regular(f2d_ctx *c, const FineView &F, std::vector<Level> &lev, int first, int last, int level0,
                         int xper);

// which of the three non-own parents of each fine cell are unknowns on the coarse level
__device__ __forceinline__ uint8_t parent_bits(int j, int i, const Level &Lc, int periodic, int pj_off) {
    int J0, Jn, I0, In;
    parents(j, i, J0, Jn, I0, In, pj_off);
    if (periodic & 2) { if (Jn < 0) Jn += Lc.ny; else if (Jn >= Lc.ny) Jn -= Lc.ny; }
    bool jin = Jn >= 0 && Jn < Lc.ny, iin = true;
    if (periodic & 1) { if (In < 0) In += Lc.nx; else if (In >= Lc.nx) In -= Lc.nx; }
    else iin = In >= 0 && In < Lc.nx;
    uint8_t c = 0;
    if (jin && (Lc.code[(long)(Jn + 1) * Lc.pitch + I0 + 1] & NB_SELF)) c |= NB_PJ;
    if (iin && (Lc.code[(long)(J0 + 1) * Lc.pitch + In + 1] & NB_SELF)) c |= NB_PI;
    if (jin && iin && (Lc.code[(long)(Jn + 1) * Lc.pitch + In + 1] & NB_SELF)) c |= NB_PJI;
    return c;
}

__global__ void k_parent_bits0(FineView F, uint8_t *__restrict__ nb, Level Lc, int periodic, int pj_off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= F.nx || j >= F.ny) return;
    long idx;
    if (!fine_index(F, j, i, idx)) return;
    uint8_t c = nb[idx];
    if (!(c & NB_SELF)) return;
    nb[idx] = (c & 31) | parent_bits(j, i, Lc, periodic, pj_off);
}

__global__ void k_parent_bits(Level Lf, Level Lc, int periodic, int pj_off) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= Lf.nx || j >= Lf.ny) return;
    long idx = (long)(j + 1) * Lf.pitch + i + 1;
    uint8_t c = Lf.code[idx];
    if (!(c & NB_SELF)) return;
    Lf.code[idx] = (c & 31) | parent_bits(j, i, Lc, periodic, pj_off);
}

__global__ void k_solver_mask(const int8_t *__restrict__ m, int8_t *__restrict__ sm, int n2, int n1,
                              int nh, int xper) {      // xper: bit 0 = x wraps, bit 1 = y wraps: halo columns / rows are images
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int j = blockIdx.y;
    if (i >= n1) return;
    long k = (long)j * n1 + i;
    int8_t v = m[k] != 0;
    if ((xper & 1) && (i < nh || i >= n1 - nh)) v = 0;
    if ((xper & 2) && (j < nh || j >= n2 - nh)) v = 0;
    sm[k] = v;
}

// ------------------------------------------------------------------- host ---
// Connected components (4-neighbours, x-periodic wrap inside the window) of the
// unknowns of the solver mask `h`, by row runs + union-find.  Returns the number of
// components; `comp` gets 0 .. MAXCOMP-1 for the MAXCOMP largest ones (by size, ties
// by first appearance) and 0xFF elsewhere, `sizes` their point counts in that order.
// `closed_only` (slab mode): count only components that touch neither the first
// nor the last row of the array (those may continue on a neighbouring rank).
static int label_components(const std::vector<int8_t> &h, int n2, int n1, const FineView &F,
                            std::vector<uint8_t> *comp, std::vector<int64_t> *sizes, bool closed_only = false) {
    struct Run { int j, i0, i1, lab; };
    std::vector<Run> runs;
    std::vector<int> parent;
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    auto unite = [&](int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); };
    size_t prev0 = 0, prev1 = 0;
    for (int j = 0; j < n2; j++) {
        const int8_t *row = h.data() + (size_t)j * n1;
        const size_t cur0 = runs.size();
        for (int i = 0; i < n1;) {
            if (!row[i]) { i++; continue; }
            int e = i;
            while (e < n1 && row[e]) e++;
            int lab = (int)parent.size();
            parent.push_back(lab);
            runs.push_back(Run{j, i, e, lab});
            i = e;
        }
        const size_t cur1 = runs.size();
        // overlap with the runs of the previous row (both lists are sorted by column)
        size_t a = prev0;
        for (size_t b = cur0; b < cur1; b++) {
            while (a < prev1 && runs[a].i1 <= runs[b].i0) a++;
            for (size_t k = a; k < prev1 && runs[k].i0 < runs[b].i1; k++) unite(runs[k].lab, runs[b].lab);
        }
        if (F.periodic && cur1 > cur0 && runs[cur0].i0 <= F.oi && runs[cur1 - 1].i1 >= F.oi + F.nx)
            unite(runs[cur0].lab, runs[cur1 - 1].lab);
        prev0 = cur0; prev1 = cur1;
    }
    if (F.periodic_y && n2 > 0) {     // rows oj and oj + ny - 1 are neighbours
        std::vector<const Run *> first, last;
        for (const Run &r : runs) {
            if (r.j == F.oj) first.push_back(&r);
            else if (r.j == F.oj + F.ny - 1) last.push_back(&r);
        }
        for (const Run *a : first)
            for (const Run *b : last)
                if (a->i0 < b->i1 && b->i0 < a->i1) unite(a->lab, b->lab);
    }
    std::vector<int64_t> cnt(parent.size(), 0);
    std::vector<char> open(parent.size(), 0);
    for (const Run &r : runs) {
        int root = find(r.lab);
        cnt[root] += r.i1 - r.i0;
        if (r.j == 0 || r.j == n2 - 1) open[root] = 1;
    }
    std::vector<int> roots;
    for (size_t k = 0; k < parent.size(); k++)
        if (parent[k] == (int)k && cnt[k] > 0 && !(closed_only && open[k])) roots.push_back((int)k);
    std::stable_sort(roots.begin(), roots.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
    if (comp) {
        std::vector<int> id(parent.size(), 0xFF);
        for (size_t k = 0; k < roots.size() && k < (size_t)Multigrid::MAXCOMP; k++) id[roots[k]] = (int)k;
        comp->assign((size_t)n2 * n1, 0xFF);
        for (const Run &r : runs) {
            uint8_t v = (uint8_t)id[find(r.lab)];
            std::fill(comp->begin() + (size_t)r.j * n1 + r.i0, comp->begin() + (size_t)r.j * n1 + r.i1, v);
        }
    }
    if (sizes) {
        sizes->clear();
        for (int r : roots) sizes->push_back(cnt[r]);
    }
    return (int)roots.size();
}

static dim3 blk() { return dim3(64, 4); }
// grid of the CG vector kernels: 256-thread blocks, x over columns, y strides rows
static dim3 cg_grid(const f2d_ctx *c, const FineView &F) {
    int gx = std::max(1, std::min((F.nx + 255) / 256, 16));
    int gy = std::max(1, std::min(F.ny, (c->nsm * 8) / gx));
    return dim3(gx, gy);
}
static dim3 grd(int nx, int ny) { return dim3((nx + 63) / 64, (ny + 3) / 4); }

static void mark_regular(f2d_ctx *c, const FineView &F, std::vector<Level> &lev, int first, int last, int level0,
                         int xper) {
    // lev[first..last] are multigrid levels level0 + (l - first)
    for (int l = first; l <= last; l++) {
        Level &L = lev[l];
        const int lvl = level0 + (l - first);
        L.cx0 = (CT)F.cx; L.cy0 = (CT)F.cy;
        double mass = F.shift;
        for (int k = 0; k < lvl; k++) mass = ((mass + mass) + mass) + mass;      // the order k_coarsen adds them in
        const CT faces = ((L.cx0 + L.cx0) + L.cy0) + L.cy0;      // k_build_dinv adds the four CT couplings first
        double diag = (double)faces + mass + 0.0;
        L.dinv0 = diag > 0.0 ? (CT)(1.0 / diag) : CT(0);
        k_mark_regular<<<grd(L.nx, L.ny), blk(), 0, c->stream>>>(L, xper, L.cx0, L.cy0, L.dinv0);
    }
}

static CoarseView view_of(const Level &L, int periodic, int dirichlet, int periodic_y = 0) {
    CoarseView V;
    V.ny = L.ny; V.nx = L.nx; V.pitch = L.pitch; V.periodic = periodic; V.periodic_y = periodic_y;
    V.cx = L.cx; V.cy = L.cy; V.dinv = L.dinv; V.code = L.code;
    V.dirichlet = dirichlet;
    V.pj_off = 0;
    return V;
}

void mg_free(f2d_ctx *c, int which) {
    Multigrid &M = c->mg[which];
    for (Multigrid::IterGraph &g : M.graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    M.graphs.clear();
    cudaFree(M.nb);
    cudaFree(M.cg_open); M.cg_open = nullptr;
    cudaFree(M.comp); M.comp = nullptr;
    for (double *p : {M.r, M.z, M.q}) cudaFree(p);
    for (float *p : {M.zf, M.zf2, M.p, M.p2}) cudaFree(p);
    for (Level &L : M.lev) {
        for (CT *p : {L.x, L.x2, L.b, L.r, L.cx, L.cy, L.dinv}) cudaFree(p);
        cudaFree(L.mass);
        cudaFree(L.wall);
        cudaFree(L.code);
    }
    for (Level &L : M.glev) {
        for (CT *p : {L.x, L.x2, L.b, L.r, L.cx, L.cy, L.dinv}) cudaFree(p);
        cudaFree(L.mass);
        cudaFree(L.wall);
        cudaFree(L.code);
    }
    M = Multigrid();
}

static int alloc_level(f2d_ctx *c, Level &L, int ny, int nx) {
    L.ny = ny; L.nx = nx;
    L.pitch = (L.nx + 2 + 1) & ~1;
    L.n = (size_t)(L.ny + 2) * L.pitch;
    for (CT **p : {&L.x, &L.x2, &L.b, &L.r, &L.cx, &L.cy, &L.dinv}) {
        F2D_CUDA(cudaMalloc(p, L.n * sizeof(CT)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, L.n * sizeof(CT), c->stream));
    }
    for (double **p : {&L.mass, &L.wall}) {
        F2D_CUDA(cudaMalloc(p, L.n * sizeof(double)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, L.n * sizeof(double), c->stream));
    }
    F2D_CUDA(cudaMalloc(&L.code, L.n));
    F2D_CUDA(cudaMemsetAsync(L.code, 0, L.n, c->stream));
    return F2D_OK;
}

static void free_setup_arrays(std::vector<Level> &lev) {
    for (size_t l = 1; l < lev.size(); l++) {
        cudaFree(lev[l].mass); lev[l].mass = nullptr;
        cudaFree(lev[l].wall); lev[l].wall = nullptr;
    }
}

// where the tail (single-CTA) part of the hierarchy starts, checking that the
// x, b, r arrays of all tail levels fit in shared memory
static int choose_tail(std::vector<Level> &lev, int first_allowed, int &tail) {
    const int n = (int)lev.size();
    tail = n - 1;
    for (int l = n - 1; l >= std::max(first_allowed, 1); l--)
        if ((long)lev[l].ny * lev[l].nx <= 4096 && n - l <= 16) tail = l;
    auto need = [&](int from) {
        size_t b = 0;
        for (int l = from; l < n; l++) b += (size_t)TAIL_ARRAYS * (lev[l].ny + 2) * (lev[l].nx + 2) * sizeof(CT);
        return b;
    };
    while (tail < n - 1 && need(tail) > 200 * 1024) tail++;
    if (need(tail) > 200 * 1024) {
        set_error("coarsest multigrid level too large (%d x %d)", lev.back().ny, lev.back().nx);
        return F2D_ERR_UNSUPPORTED;
    }
    return F2D_OK;
}

static int mg_build_slab(f2d_ctx *c, int which);

int mg_build(f2d_ctx *c, int which) {
    mg_free(c, which);
    if (c->dist.on) return mg_build_slab(c, which);
    Multigrid &M = c->mg[which];
    M.which = which;
    const int n1 = c->n1, n2 = c->n2, nh = c->nh;
    const int xper = c->cfg.xperiodic;
    const int yper = c->cfg.yperiodic == 2;      // param.ywrap: a true periodic direction (not the reference's mask-only quirk)
    const int per = xper | (yper << 1);          // what the set-up kernels take: bit 0 = x wraps, bit 1 = y wraps
    const bool vert = which != F2D_SOLVER_CENTERS;

    // solver mask and its bounding box
    int8_t *sm;
    F2D_CUDA(cudaMalloc(&sm, c->n));
    k_solver_mask<<<dim3((n1 + 127) / 128, n2), 128, 0, c->stream>>>(c->m(vert ? "mskv" : "msk"), sm,
                                                                      n2, n1, nh, per);
    LAUNCH_CHECK(c);
    std::vector<int8_t> h(c->n);
    F2D_CUDA(cudaMemcpyAsync(h.data(), sm, c->n, cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    int jmin = n2, jmax = -1, imin = n1, imax = -1;
    int64_t cnt = 0;
    for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++)
            if (h[(size_t)j * n1 + i]) {
                cnt++;
                jmin = std::min(jmin, j); jmax = std::max(jmax, j);
                imin = std::min(imin, i); imax = std::max(imax, i);
            }
    M.nunknown = cnt;
    M.n_global = (double)cnt;
    FineView &F = M.fine;
    F.n2 = n2; F.n1 = n1; F.periodic = xper; F.periodic_y = yper; F.dirichlet = vert;
    F.cx = c->dy / c->dx; F.cy = c->dx / c->dy;
    F.shift = (which == F2D_SOLVER_HELMHOLTZ) ? c->area * c->cfg.f0 * c->cfg.f0 / (c->cfg.g * c->cfg.H) : 0.0;
    if (cnt == 0) {
        cudaFree(sm);
        F.ny = F.nx = 0; F.oj = F.oi = 0;
        M.built = true;
        return F2D_OK;
    }
    // origin aligned with the interior corner (nh, nh) so that power-of-two
    // interiors aggregate cleanly; pushed out by a power of two if fluid
    // extends into the halo (the reference's yperiodic quirk, SURVEY note Y)
    auto origin = [&](int lo) {
        int o = nh;
        int span = 1;
        while (o > lo) { o = nh - span; span *= 2; }
        return o;
    };
    if (yper) { F.oj = nh; F.ny = c->cfg.ny; }
    else { F.oj = origin(jmin); F.ny = jmax + 1 - F.oj; }
    F.jo0 = 0; F.jo1 = F.ny; F.pj_off = 0;
    if (xper) { F.oi = nh; F.nx = c->cfg.nx; }
    else { F.oi = origin(imin); F.nx = imax + 1 - F.oi; }

    F2D_CUDA(cudaMalloc(&M.nb, c->n));
    F.nb = M.nb;
    k_build_nb<<<grd(n1, n2), blk(), 0, c->stream>>>(F, sm, M.nb);
    LAUNCH_CHECK(c);
    {   // open tiles of k_cg_dir_apply
        const int ntiles = ((F.nx + CGX - 1) / CGX) * ((F.ny + CGY - 1) / CGY);
        F2D_CUDA(cudaMalloc(&M.cg_open, std::max(ntiles, 1)));
        if (ntiles > 0) { k_cg_tile_flags<<<ntiles, 256, 0, c->stream>>>(F, M.cg_open); LAUNCH_CHECK(c); }
    }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(sm);
    // the operator has constants in its null space: all-Neumann centres; Dirichlet vertices only when the
    // domain is periodic both ways and has no wall at all (every vertex an unknown)
    M.singular = F.shift == 0.0 && (!vert || (xper && yper && cnt == (int64_t)F.ny * F.nx));
    if (!vert && F.shift == 0.0) {   // one null-space constant per connected component
        std::vector<uint8_t> comp;
        std::vector<int64_t> sizes;
        M.ncomp = label_components(h, n2, n1, F, &comp, &sizes);
        if (M.ncomp > 1) {
            F2D_CUDA(cudaMalloc(&M.comp, c->n));
            F2D_CUDA(cudaMemcpy(M.comp, comp.data(), c->n, cudaMemcpyHostToDevice));
            for (int k = 0; k < Multigrid::MAXCOMP; k++) M.inv_nc[k] = k < (int)sizes.size() ? 1.0 / (double)sizes[k] : 0.0;
        }
    }

    // level sizes
    std::vector<std::pair<int, int>> sizes;
    sizes.push_back({F.ny, F.nx});
    while (true) {
        int ny = sizes.back().first, nx = sizes.back().second;
        if ((long)ny * nx <= 16 || sizes.size() >= 24) break;
        if (xper && (nx & 1)) break;
        if (yper && (ny & 1)) break;
        if (ny == 1 && nx == 1) break;
        sizes.push_back({yper ? ny / 2 : (ny + 1) / 2, xper ? nx / 2 : (nx + 1) / 2});
    }
    if (sizes.size() < 2) {
        if ((xper && (F.nx & 1)) || (yper && (F.ny & 1))) {
            set_error("a periodic elliptic solve needs an even number of points in that direction (got %d x %d)", F.ny, F.nx);
            return F2D_ERR_UNSUPPORTED;
        }
        sizes.push_back({yper ? F.ny / 2 : (F.ny + 1) / 2, xper ? F.nx / 2 : (F.nx + 1) / 2});
    }
    M.lev.resize(sizes.size());
    for (size_t l = 1; l < sizes.size(); l++) {
        Level &L = M.lev[l];
        L.ny = sizes[l].first; L.nx = sizes[l].second;
        L.pitch = (L.nx + 2 + 1) & ~1;
        L.n = (size_t)(L.ny + 2) * L.pitch;
        for (CT **p : {&L.x, &L.x2, &L.b, &L.r, &L.cx, &L.cy, &L.dinv}) {
            F2D_CUDA(cudaMalloc(p, L.n * sizeof(CT)));
            F2D_CUDA(cudaMemsetAsync(*p, 0, L.n * sizeof(CT), c->stream));
        }
        for (double **p : {&L.mass, &L.wall}) {
            F2D_CUDA(cudaMalloc(p, L.n * sizeof(double)));
            F2D_CUDA(cudaMemsetAsync(*p, 0, L.n * sizeof(double), c->stream));
        }
        F2D_CUDA(cudaMalloc(&L.code, L.n));
        F2D_CUDA(cudaMemsetAsync(L.code, 0, L.n, c->stream));
    }
    M.lev[0].ny = F.ny; M.lev[0].nx = F.nx;
    // levels from M.tail on are small enough for the single-CTA tail kernel
    M.tail = (int)M.lev.size() - 1;
    // measured at 4096^2 (ms per step): tail from 64^2 down 11.20, from 32^2 10.98, from 16^2 10.96 --
    // a handful of tile CTAs relax a 64^2 level faster than one CTA does between block-wide barriers
    static const long tail_points = getenv("F2D_TAIL_POINTS") ? atol(getenv("F2D_TAIL_POINTS")) : 1024;
    for (int l = (int)M.lev.size() - 1; l >= 1; l--)
        if ((long)M.lev[l].ny * M.lev[l].nx <= tail_points && (int)M.lev.size() - l <= 16) M.tail = l;
    {   // x, b, r and the coefficient copies of every tail level must fit in shared memory
        auto need = [&](int from) {
            size_t b = 0;
            for (size_t l = from; l < M.lev.size(); l++) b += (size_t)TAIL_ARRAYS * (M.lev[l].ny + 2) * (M.lev[l].nx + 2) * sizeof(CT);
            return b;
        };
        while (M.tail < (int)M.lev.size() - 1 && need(M.tail) > 200 * 1024) M.tail++;
        if (need(M.tail) > 200 * 1024) { set_error("coarsest multigrid level too large (%d x %d)", M.lev.back().ny, M.lev.back().nx); return F2D_ERR_UNSUPPORTED; }
    }
    // coefficients, level by level
    k_coarsen0<<<grd(M.lev[1].nx, M.lev[1].ny), blk(), 0, c->stream>>>(F, M.lev[1], per, 2.0 / 3.0, 0);
    LAUNCH_CHECK(c);
    k_build_dinv<<<grd(M.lev[1].nx, M.lev[1].ny), blk(), 0, c->stream>>>(M.lev[1], per);
    LAUNCH_CHECK(c);
    for (size_t l = 1; l + 1 < M.lev.size(); l++) {
        Level &Lc = M.lev[l + 1];
        double pw = std::ldexp(1.0, (int)l);   // 2^l, producing level l+1
        k_coarsen<<<grd(Lc.nx, Lc.ny), blk(), 0, c->stream>>>(M.lev[l], Lc, per, (pw + 1.0) / (2.0 * pw + 1.0), 0);
        LAUNCH_CHECK(c);
        k_build_dinv<<<grd(Lc.nx, Lc.ny), blk(), 0, c->stream>>>(Lc, per);
        LAUNCH_CHECK(c);
    }
    // prolongation weights
    k_parent_bits0<<<grd(F.nx, F.ny), blk(), 0, c->stream>>>(F, M.nb, M.lev[1], per, 0);
    LAUNCH_CHECK(c);
    for (size_t l = 1; l + 1 < M.lev.size(); l++) {
        k_parent_bits<<<grd(M.lev[l].nx, M.lev[l].ny), blk(), 0, c->stream>>>(M.lev[l], M.lev[l + 1], per, 0);
        LAUNCH_CHECK(c);
    }
    if (M.tail > 1) { mark_regular(c, F, M.lev, 1, M.tail - 1, 1, per); LAUNCH_CHECK(c); }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t l = 1; l < M.lev.size(); l++) {   // set-up only arrays
        cudaFree(M.lev[l].mass); M.lev[l].mass = nullptr;
        cudaFree(M.lev[l].wall); M.lev[l].wall = nullptr;
    }
    for (double **p : {&M.r, &M.z, &M.q}) {
        F2D_CUDA(cudaMalloc(p, c->n * sizeof(double)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, c->n * sizeof(double), c->stream));
    }
    for (float **p : {&M.zf, &M.zf2, &M.p, &M.p2}) {
        F2D_CUDA(cudaMalloc(p, c->n * sizeof(float)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, c->n * sizeof(float), c->stream));
    }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    M.built = true;
    return F2D_OK;
}

// ---------------------------------------------------------------------------
// slab mode (f2d_dist_init): the hierarchy of a y-slab.  Levels 0..T-1 ("tile
// levels") are local: owned rows plus G ghost rows per interface, refreshed by
// dist_exchange; from level T on the grid is small, every rank keeps the whole
// (global) tail hierarchy and solves it redundantly after an all-gather.
// ---------------------------------------------------------------------------
static int exchange_level(f2d_ctx *c, Level &L, bool with_dinv) {
    void *ct[3] = {L.cx, L.cy, L.dinv};
    F2D_TRY(dist_exchange(c, with_dinv ? 3 : 2, ct, (size_t)L.pitch * sizeof(CT), L.ny, 1));
    if (L.mass) {
        void *db[2] = {L.mass, L.wall};
        F2D_TRY(dist_exchange(c, 2, db, (size_t)L.pitch * sizeof(double), L.ny, 1));
    }
    return dist_exchange1(c, L.code, (size_t)L.pitch, L.ny, 1);
}

static int mg_build_slab(f2d_ctx *c, int which) {
    Multigrid &M = c->mg[which];
    const Dist &D = c->dist;
    M.which = which;
    const int n1 = c->n1, n2 = c->n2, nh = c->nh, G = D.G;
    const int xper = c->cfg.xperiodic;
    const bool vert = which != F2D_SOLVER_CENTERS;
    const int gny = c->cfg.reserved[3];          // global interior rows
    if (c->cfg.solver_kind != 0) { set_error("slab mode supports the fused PCG solver only"); return F2D_ERR_UNSUPPORTED; }
    if (gny <= 0) { set_error("global ny missing from the configuration"); return F2D_ERR_ARG; }
    M.gs = D.south ? G : 0;
    M.gn = D.north ? G : 0;

    int8_t *sm;
    F2D_CUDA(cudaMalloc(&sm, c->n));
    k_solver_mask<<<dim3((n1 + 127) / 128, n2), 128, 0, c->stream>>>(c->m(vert ? "mskv" : "msk"), sm, n2, n1, nh, xper);
    LAUNCH_CHECK(c);
    std::vector<int8_t> h(c->n);
    F2D_CUDA(cudaMemcpyAsync(h.data(), sm, c->n, cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));

    FineView &F = M.fine;
    F.n2 = n2; F.n1 = n1; F.periodic = xper; F.periodic_y = 0; F.dirichlet = vert;
    F.cx = c->dy / c->dx; F.cy = c->dx / c->dy;
    F.shift = (which == F2D_SOLVER_HELMHOLTZ) ? c->area * c->cfg.f0 * c->cfg.f0 / (c->cfg.g * c->cfg.H) : 0.0;
    M.singular = !vert && F.shift == 0.0;
    F.oj = D.south ? 0 : nh;
    F.ny = (D.north ? n2 : n2 - nh) - F.oj;
    F.oi = nh; F.nx = c->cfg.nx;
    F.jo0 = M.gs; F.jo1 = F.ny - M.gn;
    F.pj_off = M.gs / 2;
    const int owned0 = F.jo1 - F.jo0;
    // unknowns outside the window (wall halos) are not supported in slab mode
    int64_t cnt = 0;
    for (int j = 0; j < n2; j++)
        for (int i = 0; i < n1; i++)
            if (h[(size_t)j * n1 + i]) {
                int lj = j - F.oj, li = i - F.oi;
                if (lj < 0 || lj >= F.ny || li < 0 || li >= F.nx) {
                    cudaFree(sm);
                    set_error("slab mode needs the fluid inside the interior rows/columns");
                    return F2D_ERR_UNSUPPORTED;
                }
                if (lj >= F.jo0 && lj < F.jo1) cnt++;
            }
    M.nunknown = cnt;
    double pocket = 0.0;
    if (!vert && D.world > 1 && cnt > 0) {
        // a fluid pocket closed inside this slab is a component of its own; the per-component
        // null-space projection needs global labels, which slab mode does not build
        if (label_components(h, n2, n1, F, nullptr, nullptr, true) > 0) pocket = 1.0;
    }
    {
        double v[2] = {(double)cnt, pocket};
        F2D_CUDA(cudaMemcpyAsync(c->d_scal + 24, v, 2 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        F2D_TRY(dist_allreduce(c, c->d_scal + 24, 2, false));
        F2D_CUDA(cudaMemcpyAsync(v, c->d_scal + 24, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        F2D_CUDA(cudaStreamSynchronize(c->stream));
        M.n_global = v[0];
        if (v[1] > 0.0) {      // every rank takes this branch
            cudaFree(sm);
            set_error("slab mode: the fluid has a pocket disconnected from the main basin "
                      "(the per-component Neumann null space is single-GPU only)");
            return F2D_ERR_UNSUPPORTED;
        }
    }
    F2D_CUDA(cudaMalloc(&M.nb, c->n));
    F.nb = M.nb;
    k_build_nb<<<grd(n1, n2), blk(), 0, c->stream>>>(F, sm, M.nb);
    LAUNCH_CHECK(c);
    {   // open tiles of k_cg_dir_apply
        const int ntiles = ((F.nx + CGX - 1) / CGX) * ((F.ny + CGY - 1) / CGY);
        F2D_CUDA(cudaMalloc(&M.cg_open, std::max(ntiles, 1)));
        if (ntiles > 0) { k_cg_tile_flags<<<ntiles, 256, 0, c->stream>>>(F, M.cg_open); LAUNCH_CHECK(c); }
    }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(sm);
    if (M.n_global == 0) { M.built = true; return F2D_OK; }

    // tile levels: T = first level whose GLOBAL grid fits the tail kernel
    int T = 1;
    {
        long gy = gny, gx = F.nx;
        int own = owned0;
        for (T = 1;; T++) {
            if ((gy & 1) || (own & 1) || (xper && (gx & 1))) {
                set_error("slab mode: rows per rank (%d) and nx must be divisible by 2^%d", owned0, T);
                return F2D_ERR_UNSUPPORTED;
            }
            gy /= 2; own /= 2; gx = xper ? gx / 2 : (gx + 1) / 2;
            if (gy * gx <= 4096) break;
            if (own < G) { set_error("slab mode: too few rows per rank (%d) for %d ranks", owned0, D.world); return F2D_ERR_UNSUPPORTED; }
        }
    }
    M.lev.resize(T + 1);
    M.lev[0].ny = F.ny; M.lev[0].nx = F.nx;
    {
        int own = owned0, nx = F.nx;
        for (int l = 1; l <= T; l++) {
            own /= 2; nx = xper ? nx / 2 : (nx + 1) / 2;
            F2D_TRY(alloc_level(c, M.lev[l], M.gs + own + M.gn, nx));
        }
    }
    M.tail = T;
    const int pj = F.pj_off;
    // local coefficients, level by level, ghosts from the neighbours
    k_coarsen0<<<grd(M.lev[1].nx, M.lev[1].ny), blk(), 0, c->stream>>>(F, M.lev[1], xper, 2.0 / 3.0, pj);
    LAUNCH_CHECK(c);
    for (int l = 1; l <= T; l++) {
        Level &L = M.lev[l];
        F2D_TRY(exchange_level(c, L, false));
        k_build_dinv<<<grd(L.nx, L.ny), blk(), 0, c->stream>>>(L, xper);
        LAUNCH_CHECK(c);
        F2D_TRY(exchange_level(c, L, true));
        if (l < T) {
            double pw = std::ldexp(1.0, l);
            k_coarsen<<<grd(M.lev[l + 1].nx, M.lev[l + 1].ny), blk(), 0, c->stream>>>(L, M.lev[l + 1], xper, (pw + 1.0) / (2.0 * pw + 1.0), pj);
            LAUNCH_CHECK(c);
        }
    }
    k_parent_bits0<<<grd(F.nx, F.ny), blk(), 0, c->stream>>>(F, M.nb, M.lev[1], xper, pj);
    LAUNCH_CHECK(c);
    for (int l = 1; l < T; l++) {
        k_parent_bits<<<grd(M.lev[l].nx, M.lev[l].ny), blk(), 0, c->stream>>>(M.lev[l], M.lev[l + 1], xper, pj);
        LAUNCH_CHECK(c);
    }
    if (T > 1) { mark_regular(c, F, M.lev, 1, T - 1, 1, xper); LAUNCH_CHECK(c); }
    // global tail hierarchy: level T gathered from the owners, coarser ones derived
    {
        const Level &LT = M.lev[T];
        const int ownT = LT.ny - M.gs - M.gn;
        M.tail_y0 = D.rank * ownT;
        std::vector<std::pair<int, int>> sizes;
        sizes.push_back({ownT * D.world, LT.nx});
        while (true) {
            int ny = sizes.back().first, nx = sizes.back().second;
            if ((long)ny * nx <= 16 || sizes.size() >= 16) break;
            if (xper && (nx & 1)) break;
            if (ny == 1 && nx == 1) break;
            sizes.push_back({(ny + 1) / 2, xper ? nx / 2 : (nx + 1) / 2});
        }
        M.glev.resize(sizes.size());
        for (size_t l = 0; l < sizes.size(); l++) F2D_TRY(alloc_level(c, M.glev[l], sizes[l].first, sizes[l].second));
        Level &G0 = M.glev[0];
        if (G0.pitch != LT.pitch) { set_error("internal: tail pitch mismatch"); return F2D_ERR_STATE; }
        const size_t r0 = (size_t)(1 + M.gs) * LT.pitch, rows = (size_t)ownT * LT.pitch;
        F2D_TRY(dist_allgather_rows(c, LT.cx + r0, G0.cx + G0.pitch, rows * sizeof(CT)));
        F2D_TRY(dist_allgather_rows(c, LT.cy + r0, G0.cy + G0.pitch, rows * sizeof(CT)));
        F2D_TRY(dist_allgather_rows(c, LT.mass + r0, G0.mass + G0.pitch, rows * sizeof(double)));
        F2D_TRY(dist_allgather_rows(c, LT.wall + r0, G0.wall + G0.pitch, rows * sizeof(double)));
        F2D_TRY(dist_allgather_rows(c, LT.code + r0, G0.code + G0.pitch, rows));
        k_build_dinv<<<grd(G0.nx, G0.ny), blk(), 0, c->stream>>>(G0, xper);
        LAUNCH_CHECK(c);
        for (size_t l = 0; l + 1 < M.glev.size(); l++) {
            Level &Lc = M.glev[l + 1];
            double pw = std::ldexp(1.0, (int)l + T);
            k_coarsen<<<grd(Lc.nx, Lc.ny), blk(), 0, c->stream>>>(M.glev[l], Lc, xper, (pw + 1.0) / (2.0 * pw + 1.0), 0);
            LAUNCH_CHECK(c);
            k_build_dinv<<<grd(Lc.nx, Lc.ny), blk(), 0, c->stream>>>(Lc, xper);
            LAUNCH_CHECK(c);
        }
        for (size_t l = 0; l + 1 < M.glev.size(); l++) {
            k_parent_bits<<<grd(M.glev[l].nx, M.glev[l].ny), blk(), 0, c->stream>>>(M.glev[l], M.glev[l + 1], xper, 0);
            LAUNCH_CHECK(c);
        }
        int dummy;
        std::vector<Level> chk(M.glev.size() + 1);
        for (size_t l = 0; l < M.glev.size(); l++) { chk[l + 1].ny = M.glev[l].ny; chk[l + 1].nx = M.glev[l].nx; }
        F2D_TRY(choose_tail(chk, 1, dummy));
        if (dummy != 1) { set_error("slab mode: the gathered tail grid does not fit one CTA"); return F2D_ERR_UNSUPPORTED; }
    }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    free_setup_arrays(M.lev);
    for (Level &L : M.glev) { cudaFree(L.mass); L.mass = nullptr; cudaFree(L.wall); L.wall = nullptr; }
    for (double **p : {&M.r, &M.z, &M.q}) {
        F2D_CUDA(cudaMalloc(p, c->n * sizeof(double)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, c->n * sizeof(double), c->stream));
    }
    for (float **p : {&M.zf, &M.zf2, &M.p, &M.p2}) {
        F2D_CUDA(cudaMalloc(p, c->n * sizeof(float)));
        F2D_CUDA(cudaMemsetAsync(*p, 0, c->n * sizeof(float), c->stream));
    }
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    M.built = true;
    return F2D_OK;
}

// ---------------------------------------------------------------------------
// single-CTA tail: the whole sub-V-cycle of the levels whose grids are small
// (<= 1024 points on one GPU, <= 4096 gathered points in slab mode) in one launch.  Everything lives in shared memory: x, b and
// the residual of every tail level, and a copy of the coefficients (couplings,
// inverse diagonal, prolongation normaliser) made once at the start with all
// loads in flight -- the kernel is a chain of ~80 short phases separated by
// barriers, and a global (L2) load in each of them was most of its run time.
// ---------------------------------------------------------------------------
struct TailLevel {
    int ny, nx, pitch;          // global arrays: (ny+2) x pitch
    int sp, off;                // shared arrays: pitch nx+2, offset (in CT) of x; b, r, cx, cy, dinv, w follow
    CT *x;
    const CT *b;
    const CT *cx, *cy, *dinv;
    const uint8_t *code;
};
struct TailArgs {
    int nlev, periodic, dirichlet, nu1, nu2, nsw;
    TailLevel lev[16];
};

#define TAIL_LOOP(L)                                            \
    for (int J = threadIdx.x >> 5; J < (L).ny; J += 32)         \
        for (int I = threadIdx.x & 31; I < (L).nx; I += 32)

struct TailSm { CT *x, *b, *r, *cx, *cy, *dinv, *w; };
__device__ __forceinline__ TailSm tail_sm(CT *sm, const TailLevel &L) {
    int n = (L.ny + 2) * L.sp;
    CT *p = sm + L.off;
    return TailSm{p, p + n, p + 2 * n, p + 3 * n, p + 4 * n, p + 5 * n, p + 6 * n};
}

// periodic: bit 0 = x wraps, bit 1 = y wraps
__device__ __forceinline__ CT tail_offdiag(const TailLevel &L, const TailSm &S, int periodic, int J, int I) {
    int s = (J + 1) * L.sp + I + 1, w = s - 1, e = s + 1, so = s - L.sp, no = s + L.sp;
    if (periodic & 1) {
        if (I == 0) w = s + (L.nx - 1);
        if (I == L.nx - 1) e = s - (L.nx - 1);
    }
    if (periodic & 2) {
        if (J == 0) so = s + (L.ny - 1) * L.sp;
        if (J == L.ny - 1) no = s - (L.ny - 1) * L.sp;
    }
    return S.cx[s] * S.x[w] + S.cx[e] * S.x[e] + S.cy[s] * S.x[so] + S.cy[no] * S.x[no];
}

// (tried: warp 0 alone walking the levels of <= 64 points with __syncwarp instead of 1024-thread barriers --
// 0.041 ms against 0.033 ms: the serial warp loses more than the ~45 barriers cost)
__device__ __forceinline__ void tail_relax(const TailLevel &L, CT *sm, int periodic, int color, bool zero) {
    TailSm S = tail_sm(sm, L);
    // only the points of this colour are visited: column I = 2k + ((J + color) & 1)
    for (int J = threadIdx.x >> 5; J < L.ny; J += 32)
        for (int I = 2 * (threadIdx.x & 31) + ((J + color) & 1); I < L.nx; I += 64) {
            int s = (J + 1) * L.sp + I + 1;
            CT di = S.dinv[s];
            CT a = zero ? CT(0) : tail_offdiag(L, S, periodic, J, I);
            S.x[s] = (S.b[s] + a) * di;        // di == 0 off the unknowns: stays 0
        }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_mg_tail(const __grid_constant__ TailArgs Ain) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // level descriptors are indexed dynamically: keep them in shared memory, not
    // in a local-memory copy of the parameter block
    __shared__ TailArgs A;
    for (int t = threadIdx.x; t < (int)(sizeof(TailArgs) / sizeof(int)); t += blockDim.x)
        reinterpret_cast<int *>(&A)[t] = reinterpret_cast<const int *>(&Ain)[t];
    __syncthreads();
    CT *sm = reinterpret_cast<CT *>(smem_raw);
    const int per = A.periodic;
    // coefficients of every tail level (halo rows / columns included: the east and
    // north faces of the last column / row live there), and the right-hand side
    // of the first tail level, which comes from the level above
    for (int l = 0; l < A.nlev; l++) {
        const TailLevel &L = A.lev[l];
        TailSm S = tail_sm(sm, L);
        const int n = (L.ny + 2) * L.sp;
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            int j = t / L.sp, i = t - j * L.sp;
            long g = (long)j * L.pitch + i;
            S.cx[t] = __ldg(L.cx + g);
            S.cy[t] = __ldg(L.cy + g);
            S.dinv[t] = __ldg(L.dinv + g);
            S.w[t] = (CT)w16_of(__ldg(L.code + g), A.dirichlet);
            if (l == 0) S.b[t] = __ldg(L.b + g);
        }
    }
    __syncthreads();
    for (int l = 0; l < A.nlev - 1; l++) {
        const TailLevel &L = A.lev[l], &C = A.lev[l + 1];
        TailSm S = tail_sm(sm, L), SC = tail_sm(sm, C);
        for (int t = threadIdx.x; t < (L.ny + 2) * L.sp; t += blockDim.x) S.x[t] = CT(0);
        __syncthreads();
        for (int s = 0; s < A.nu1; s++) {
            tail_relax(L, sm, per, 0, s == 0);
            tail_relax(L, sm, per, 1, false);
        }
        TAIL_LOOP(L) {   // residual / prolongation normaliser
            int sidx = (J + 1) * L.sp + I + 1;
            CT di = S.dinv[sidx], res = CT(0);
            if (di != CT(0)) {
                CT a = tail_offdiag(L, S, per, J, I);
                res = (S.b[sidx] - (S.x[sidx] / di - a)) / S.w[sidx];
            }
            S.r[sidx] = res;
        }
        __syncthreads();
        TAIL_LOOP(C) {   // restriction R = P^T
            CT acc = CT(0);
#pragma unroll
            for (int a = -1; a <= 2; a++) {
                int j = 2 * J + a;
                if (per & 2) { if (j < 0) j += L.ny; else if (j >= L.ny) j -= L.ny; }
                if (j < 0 || j >= L.ny) continue;
                CT wy = (a == 0 || a == 1) ? CT(3) : CT(1);
#pragma unroll
                for (int b = -1; b <= 2; b++) {
                    int i = 2 * I + b;
                    if (per & 1) { if (i < 0) i += L.nx; else if (i >= L.nx) i -= L.nx; }
                    if (i < 0 || i >= L.nx) continue;
                    CT wx = (b == 0 || b == 1) ? CT(3) : CT(1);
                    acc += wy * wx * S.r[(j + 1) * L.sp + i + 1];
                }
            }
            SC.b[(J + 1) * C.sp + I + 1] = acc;
        }
        __syncthreads();
    }
    {   // coarsest: nsw sweeps (R,B) then nsw sweeps (B,R) from zero
        const TailLevel &L = A.lev[A.nlev - 1];
        TailSm S = tail_sm(sm, L);
        for (int t = threadIdx.x; t < (L.ny + 2) * L.sp; t += blockDim.x) S.x[t] = CT(0);
        __syncthreads();
        for (int s = 0; s < 4 * A.nsw; s++)
            tail_relax(L, sm, per, (s < 2 * A.nsw) ? (s & 1) : 1 - (s & 1), s == 0);
    }
    for (int l = A.nlev - 2; l >= 0; l--) {
        const TailLevel &L = A.lev[l], &C = A.lev[l + 1];
        TailSm S = tail_sm(sm, L), SC = tail_sm(sm, C);
        TAIL_LOOP(L) {
            int sidx = (J + 1) * L.sp + I + 1;
            if (S.dinv[sidx] == CT(0)) continue;      // not an unknown (set-up clears the code where 1/diag is 0)
            int J0, Jn, I0, In;
            parents(J, I, J0, Jn, I0, In);
            if (per & 1) In = wrap_mod(In, C.nx);
            if (per & 2) Jn = wrap_mod(Jn, C.ny);
            int r0 = (J0 + 1) * C.sp, rn = (Jn + 1) * C.sp;   // halo rows/cols hold 0
            CT v = CT(9) * SC.x[r0 + I0 + 1] + CT(3) * (SC.x[rn + I0 + 1] + SC.x[r0 + In + 1]) + SC.x[rn + In + 1];
            S.x[sidx] += v / S.w[sidx];
        }
        __syncthreads();
        for (int s = 0; s < A.nu2; s++) {
            tail_relax(L, sm, per, 1, false);
            tail_relax(L, sm, per, 0, false);
        }
    }
    {   // the correction of the first tail level goes back to global memory
        const TailLevel &L = A.lev[0];
        TailSm S = tail_sm(sm, L);
        TAIL_LOOP(L) L.x[(long)(J + 1) * L.pitch + I + 1] = S.x[(J + 1) * L.sp + I + 1];
    }
}

// ---------------------------------------------------------------------------
// host side of the fused V-cycle
// ---------------------------------------------------------------------------
template <class K>
static int set_smem(K kernel, size_t bytes) {
    F2D_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return F2D_OK;
}

static CoarseArrays<CT> arrays_of(const Level &L, int periodic, int dirichlet, int pj_off, int periodic_y) {
    CoarseArrays<CT> A;
    A.ny = L.ny; A.nx = L.nx; A.pitch = L.pitch; A.periodic = periodic; A.periodic_y = periodic_y; A.dirichlet = dirichlet; A.pj_off = pj_off;
    A.cx = L.cx; A.cy = L.cy; A.dinv = L.dinv; A.code = L.code;
    A.cx0 = L.cx0; A.cy0 = L.cy0; A.dinv0 = L.dinv0;
    return A;
}

// where the finished correction of coarse level l lives: the tail kernel works
// in place, the tile kernels write their up leg to the second buffer
static const CT *level_result(const Multigrid &M, int l) { return l >= M.tail ? M.lev[l].x : M.lev[l].x2; }

// persistent grid of k_cg_dir_apply: exactly the CTAs that are resident at once
template <typename TZ>
static dim3 dir_apply_grid(const f2d_ctx *c) {
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_dir_apply<TZ>, 256, 0) != cudaSuccess || per_sm < 1)
            per_sm = 4;
    }
    return dim3(c->nsm * per_sm);
}

static dim3 update_grid(const f2d_ctx *c) {
    static int per_sm = 0;
    if (per_sm == 0) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cg_update_p, 256, 0) != cudaSuccess || per_sm < 1)
            per_sm = 4;
    }
    return dim3(c->nsm * per_sm);
}

// F2D_NO_OPEN=1: every tile takes the generic (masked) path -- A/B timing and the
// test that both paths give the same bits
static int allow_open_tiles() {
    static const int v = getenv("F2D_NO_OPEN") == nullptr;
    return v;
}

// fine-level legs.  FT = float: the CG preconditioner (z = M r) relaxes in fp32
// on the fp64 residual and stores z in fp32; FT = double: plain V-cycles on x itself.
template <typename FT, int NU, bool ZERO>
static int launch_down0(f2d_ctx *c, Multigrid &M, const FT *xin, FT *xout, const double *f, double fscale, int sumr_slot) {
    constexpr int WJ = 64, H = halo_down(NU, ZERO), TJ = WJ - 2 * H, TI = TW - 2 * H;
    const FineView &F = M.fine;
    FineLevel L{F};
    auto kern = k_mg_down<FT, FT, double, CT, true, ZERO, NU, WJ, FineLevel>;
    size_t smem = Window<FT, true, WJ>::bytes();
    static bool once = false;
    if (!once) { F2D_TRY(set_smem(kern, smem)); once = true; }
    dim3 g((F.nx + TI - 1) / TI, (F.ny + TJ - 1) / TJ);
    const Level &C = M.lev[1];
    kern<<<g, TILE_THREADS, smem, c->stream>>>(L, xin, xout, f, fscale, c->d_scal, sumr_slot, 1.0 / M.n_global,
                                               DownArgs{C.ny, C.nx, C.pitch, allow_open_tiles()}, C.b);
    LAUNCH_CHECK(c);
    return F2D_OK;
}

template <typename FT, int NU, bool DOT>
static int launch_up0(f2d_ctx *c, Multigrid &M, const FT *xin, FT *xout, const double *f, double fscale, int sumr_slot) {
    constexpr int WJ = 64, H = halo_up(NU), TJ = WJ - 2 * H, TI = TW - 2 * H;
    const FineView &F = M.fine;
    FineLevel L{F};
    auto kern = k_mg_up<FT, FT, double, CT, true, DOT, NU, WJ, FineLevel>;
    size_t smem = ((Window<FT, true, WJ>::bytes() + 15) & ~size_t(15)) + (WJ / 2 + 3) * (TW / 2 + 3) * sizeof(CT);
    static bool once = false;
    if (!once) { F2D_TRY(set_smem(kern, smem)); once = true; }
    dim3 g((F.nx + TI - 1) / TI, (F.ny + TJ - 1) / TJ);
    if (DOT && (size_t)g.x * g.y * 2 > c->part_capacity) { set_error("reduction scratch too small"); return F2D_ERR_STATE; }
    const Level &C = M.lev[1];
    kern<<<g, TILE_THREADS, smem, c->stream>>>(L, xin, xout, f, fscale, c->d_scal, sumr_slot, 1.0 / M.n_global,
                                               UpArgs{C.ny, C.nx, C.pitch, F.periodic, allow_open_tiles(), F.periodic_y},
                                               level_result(M, 1), c->d_part, c->d_count, c->d_scal + S_RZNEW);
    LAUNCH_CHECK(c);
    if (DOT) {
        k_fold_partials<OpSum, 2><<<1, 1024, 0, c->stream>>>(c->d_part, g.x * g.y, c->d_scal + S_RZNEW);
        LAUNCH_CHECK(c);
    }
    return F2D_OK;
}

template <int NU, int WJ>
static int launch_down(f2d_ctx *c, Multigrid &M, int l) {
    constexpr int H = halo_down(NU, true), TJ = WJ - 2 * H, TI = TW - 2 * H;
    Level &Lv = M.lev[l];
    CoarseLevel<CT> L{arrays_of(Lv, M.fine.periodic, M.fine.dirichlet, M.fine.pj_off, M.fine.periodic_y)};
    auto kern = k_mg_down<CT, CT, CT, CT, false, true, NU, WJ, CoarseLevel<CT>>;
    size_t smem = Window<CT, false, WJ>::bytes();
    static bool once = false;
    if (!once) { F2D_TRY(set_smem(kern, smem)); once = true; }
    dim3 g((Lv.nx + TI - 1) / TI, (Lv.ny + TJ - 1) / TJ);
    const Level &C = M.lev[l + 1];
    kern<<<g, TILE_THREADS, smem, c->stream>>>(L, Lv.x, Lv.x, Lv.b, 1.0, c->d_scal, -1, 0.0,
                                               DownArgs{C.ny, C.nx, C.pitch, allow_open_tiles()}, C.b);
    LAUNCH_CHECK(c);
    return F2D_OK;
}

template <int NU, int WJ>
static int launch_up(f2d_ctx *c, Multigrid &M, int l) {
    constexpr int H = halo_up(NU), TJ = WJ - 2 * H, TI = TW - 2 * H;
    Level &Lv = M.lev[l];
    CoarseLevel<CT> L{arrays_of(Lv, M.fine.periodic, M.fine.dirichlet, M.fine.pj_off, M.fine.periodic_y)};
    auto kern = k_mg_up<CT, CT, CT, CT, false, false, NU, WJ, CoarseLevel<CT>>;
    size_t smem = ((Window<CT, false, WJ>::bytes() + 15) & ~size_t(15)) + (WJ / 2 + 3) * (TW / 2 + 3) * sizeof(CT);
    static bool once = false;
    if (!once) { F2D_TRY(set_smem(kern, smem)); once = true; }
    dim3 g((Lv.nx + TI - 1) / TI, (Lv.ny + TJ - 1) / TJ);
    const Level &C = M.lev[l + 1];
    kern<<<g, TILE_THREADS, smem, c->stream>>>(L, Lv.x, Lv.x2, Lv.b, 1.0, c->d_scal, -1, 0.0,
                                               UpArgs{C.ny, C.nx, C.pitch, M.fine.periodic, allow_open_tiles(), M.fine.periodic_y}, level_result(M, l + 1),
                                               nullptr, nullptr, nullptr);
    LAUNCH_CHECK(c);
    return F2D_OK;
}

// rows from which a coarse level gets 64-row windows instead of 32-row ones
static int tall_window_rows() {
    static const int v = getenv("F2D_WJ_SWITCH") ? atoi(getenv("F2D_WJ_SWITCH")) : 1024;   // experiments
    return v;
}
template <int NU>
static int coarse_down(f2d_ctx *c, Multigrid &M, int l) {
    return M.lev[l].ny >= tall_window_rows() ? launch_down<NU, 64>(c, M, l) : launch_down<NU, 32>(c, M, l);
}
template <int NU>
static int coarse_up(f2d_ctx *c, Multigrid &M, int l) {
    return M.lev[l].ny >= tall_window_rows() ? launch_up<NU, 64>(c, M, l) : launch_up<NU, 32>(c, M, l);
}

#define NU_SWITCH(nu, CALL)                                   \
    switch (nu) {                                             \
    case 1: { constexpr int NU = 1; F2D_TRY(CALL); } break;   \
    case 2: { constexpr int NU = 2; F2D_TRY(CALL); } break;   \
    default: { constexpr int NU = 3; F2D_TRY(CALL); } break;  \
    }

static int launch_tail(f2d_ctx *c, Multigrid &M) {
    TailArgs A;
    const bool slab = c->dist.on;
    // slab mode: the tail works on the gathered (global) copies of its levels
    std::vector<Level> &TL = slab ? M.glev : M.lev;
    const int first = slab ? 0 : M.tail;
    const int nlev = (int)TL.size();
    A.nlev = nlev - first;
    A.periodic = M.fine.periodic | (M.fine.periodic_y << 1); A.dirichlet = M.fine.dirichlet;
    A.nu1 = c->cfg.nu1 > 0 ? c->cfg.nu1 : 2;
    A.nu2 = c->cfg.nu2 > 0 ? c->cfg.nu2 : 2;
    const Level &last = TL[nlev - 1];
    A.nsw = last.ny * last.nx <= 64 ? 8 : 24;
    if (const char *e = getenv("F2D_TAIL_NSW")) A.nsw = atoi(e);      // experiments
    int off = 0;
    for (int l = first; l < nlev; l++) {
        const Level &L = TL[l];
        int sp = L.nx + 2;
        A.lev[l - first] = TailLevel{L.ny, L.nx, L.pitch, sp, off, L.x, L.b, L.cx, L.cy, L.dinv, L.code};
        off += TAIL_ARRAYS * (L.ny + 2) * sp;
    }
    if (slab) {   // right-hand side: every rank's owned rows of level `tail`
        const Level &LT = M.lev[M.tail];
        const int ownT = LT.ny - M.gs - M.gn;
        F2D_TRY(dist_allgather_rows(c, LT.b + (size_t)(1 + M.gs) * LT.pitch, M.glev[0].b + M.glev[0].pitch,
                                    (size_t)ownT * LT.pitch * sizeof(CT)));
    }
    size_t smem = (size_t)off * sizeof(CT);
    static size_t configured = 0;
    if (smem > configured) { F2D_TRY(set_smem(k_mg_tail, smem)); configured = smem; }
    k_mg_tail<<<1, 1024, smem, c->stream>>>(A);
    LAUNCH_CHECK(c);
    if (slab) {   // my rows (owned + ghosts) of the global correction
        Level &LT = M.lev[M.tail];
        const Level &G0 = M.glev[0];
        F2D_CUDA(cudaMemcpyAsync(LT.x + LT.pitch, G0.x + (size_t)(1 + M.tail_y0 - M.gs) * G0.pitch,
                                 (size_t)LT.ny * LT.pitch * sizeof(CT), cudaMemcpyDeviceToDevice, c->stream));
    }
    return F2D_OK;
}

// One fused V-cycle on  L x = fscale*f - mean  (mean taken from scal[sumr_slot]/N
// when sumr_slot >= 0).  zero_guess: x is only written.  with_dot: the up leg
// leaves (sum f x, sum x) in S_RZNEW, S_SUMZ.
static int exchange_fine(f2d_ctx *c, const Multigrid &M, double *a) {
    return dist_exchange1(c, a, (size_t)M.fine.n1 * sizeof(double), M.fine.ny, M.fine.oj);
}
static int exchange_coarse(f2d_ctx *c, const Level &L, CT *a) {
    return dist_exchange1(c, a, (size_t)L.pitch * sizeof(CT), L.ny, 1);
}

// ghost rows of what the down leg of level l wrote: x_l (M.z on the fine level)
// and b_{l+1} (unless level l+1 is gathered for the tail)
static int exchange_leg(f2d_ctx *c, Multigrid &M, int l, void *fine_x, size_t fine_elem) {
    const Dist &D = c->dist;
    if (!D.on || (!D.south && !D.north)) return F2D_OK;
    struct Part { char *base; size_t row_bytes; long nrows, row0; } parts[2];
    int np = 0;
    if (l == 0) parts[np++] = Part{(char *)fine_x, (size_t)M.fine.n1 * fine_elem, M.fine.ny, M.fine.oj};
    else parts[np++] = Part{(char *)M.lev[l].x, (size_t)M.lev[l].pitch * sizeof(CT), M.lev[l].ny, 1};
    if (l + 1 < M.tail) parts[np++] = Part{(char *)M.lev[l + 1].b, (size_t)M.lev[l + 1].pitch * sizeof(CT), M.lev[l + 1].ny, 1};
    return dist_exchange_parts(c, np, &parts[0].base, &parts[0].row_bytes, &parts[0].nrows, &parts[0].row0, sizeof(Part));
}

// FT = float: preconditioner z = M f (zero guess; M.zf -> xout = M.zf2, fp32);
// FT = double: one V-cycle on x itself (x -> M.z -> x, fp64).
template <typename FT>
static int vcycle_fused(f2d_ctx *c, Multigrid &M, FT *work, const FT *xin, FT *xout, const double *f, double fscale,
                        bool zero_guess, int sumr_slot, bool with_dot) {
    // down leg: zero guess writes `work`; otherwise reads xin and writes `work`
    // (out of place).  up leg: reads `work`, writes xout.
    // Slab mode: the ghost rows of everything a leg wrote are refreshed from the
    // owners before the next leg reads them.
    const bool slab = c->dist.on;
    const int nu1 = std::min(c->cfg.nu1 > 0 ? c->cfg.nu1 : 2, 3), nu2 = std::min(c->cfg.nu2 > 0 ? c->cfg.nu2 : 2, 3);
    if (zero_guess) { NU_SWITCH(nu1, (launch_down0<FT, NU, true>(c, M, work, work, f, fscale, sumr_slot))); }
    else { NU_SWITCH(nu1, (launch_down0<FT, NU, false>(c, M, xin, work, f, fscale, sumr_slot))); }
    // (each leg's outputs travel in ONE NCCL group: x of this level and the
    // restricted right-hand side of the next)
    if (slab) F2D_TRY(exchange_leg(c, M, 0, work, sizeof(FT)));
    for (int l = 1; l < M.tail; l++) {
        NU_SWITCH(nu1, (coarse_down<NU>(c, M, l)));
        if (slab) F2D_TRY(exchange_leg(c, M, l, nullptr, 0));
    }
    F2D_TRY(launch_tail(c, M));
    for (int l = M.tail - 1; l >= 1; l--) {
        NU_SWITCH(nu2, (coarse_up<NU>(c, M, l)));
        if (slab) F2D_TRY(exchange_coarse(c, M.lev[l], M.lev[l].x2));
    }
    if (with_dot) { NU_SWITCH(nu2, (launch_up0<FT, NU, true>(c, M, work, xout, f, fscale, sumr_slot))); }
    else { NU_SWITCH(nu2, (launch_up0<FT, NU, false>(c, M, work, xout, f, fscale, sumr_slot))); }
    if (slab) {
        F2D_TRY(dist_exchange1(c, xout, (size_t)M.fine.n1 * sizeof(FT), M.fine.ny, M.fine.oj));
        if (with_dot) F2D_TRY(dist_allreduce(c, c->d_scal + S_RZNEW, 2, false));
    }
    return F2D_OK;
}

// The same V-cycle with one kernel per half-sweep (solver_kind & 2): kept as the
// cross-check of the tile kernels.
static int vcycle_unfused(f2d_ctx *c, Multigrid &M, double *x, const double *f, double fscale, bool zero_guess) {
    const FineView &F = M.fine;
    const int xper = F.periodic;
    const int nu1 = c->cfg.nu1 > 0 ? c->cfg.nu1 : 2, nu2 = c->cfg.nu2 > 0 ? c->cfg.nu2 : 2;
    const int nlev = (int)M.lev.size();
    cudaStream_t st = c->stream;
    dim3 g0 = grd(F.nx, F.ny);
    for (int s = 0; s < nu1; s++)
        for (int col = 0; col < 2; col++) {
            if (zero_guess && s == 0 && col == 0) {
                k_smooth0<true><<<g0, blk(), 0, st>>>(F, x, f, 0.0, 1);
                LAUNCH_CHECK(c);
                k_smooth0<true><<<g0, blk(), 0, st>>>(F, x, f, fscale, 0);
            } else
                k_smooth0<false><<<g0, blk(), 0, st>>>(F, x, f, fscale, col);
            LAUNCH_CHECK(c);
        }
    // M.q is free while a V-cycle runs (CG only uses it between apply and update)
    k_resid0<<<g0, blk(), 0, st>>>(F, x, f, fscale, M.q);
    LAUNCH_CHECK(c);
    {
        CoarseView C1 = view_of(M.lev[1], xper, F.dirichlet, F.periodic_y);
        k_restrict0<<<grd(C1.nx, C1.ny), blk(), 0, st>>>(F, M.q, C1, M.lev[1].b);
        LAUNCH_CHECK(c);
    }
    for (int l = 1; l < nlev - 1; l++) {
        Level &L = M.lev[l];
        CoarseView V = view_of(L, xper, F.dirichlet, F.periodic_y);
        dim3 g = grd(L.nx, L.ny);
        F2D_CUDA(cudaMemsetAsync(L.x, 0, L.n * sizeof(CT), st));
        for (int s = 0; s < nu1; s++)
            for (int col = 0; col < 2; col++) {
                if (s == 0 && col == 0) k_smooth<true><<<g, blk(), 0, st>>>(V, L.x, L.b, col);
                else k_smooth<false><<<g, blk(), 0, st>>>(V, L.x, L.b, col);
                LAUNCH_CHECK(c);
            }
        k_resid<<<g, blk(), 0, st>>>(V, L.x, L.b, L.r);
        LAUNCH_CHECK(c);
        CoarseView C = view_of(M.lev[l + 1], xper, F.dirichlet, F.periodic_y);
        k_restrict<<<grd(C.nx, C.ny), blk(), 0, st>>>(V, L.r, C, M.lev[l + 1].b);
        LAUNCH_CHECK(c);
    }
    {
        Level &L = M.lev[nlev - 1];
        int npts = L.ny * L.nx;
        k_coarsest<<<1, 1024, 0, st>>>(view_of(L, xper, F.dirichlet, F.periodic_y), L.x, L.b, npts <= 64 ? 8 : 24);
        LAUNCH_CHECK(c);
    }
    for (int l = nlev - 2; l >= 1; l--) {
        Level &L = M.lev[l];
        CoarseView V = view_of(L, xper, F.dirichlet, F.periodic_y);
        dim3 g = grd(L.nx, L.ny);
        k_prolong<<<g, blk(), 0, st>>>(V, L.x, view_of(M.lev[l + 1], xper, F.dirichlet, F.periodic_y), M.lev[l + 1].x);
        LAUNCH_CHECK(c);
        for (int s = 0; s < nu2; s++)
            for (int col = 1; col >= 0; col--) {
                k_smooth<false><<<g, blk(), 0, st>>>(V, L.x, L.b, col);
                LAUNCH_CHECK(c);
            }
    }
    k_prolong0<<<g0, blk(), 0, st>>>(F, x, view_of(M.lev[1], xper, F.dirichlet, F.periodic_y), M.lev[1].x);
    LAUNCH_CHECK(c);
    for (int s = 0; s < nu2; s++)
        for (int col = 1; col >= 0; col--) {
            k_smooth0<false><<<g0, blk(), 0, st>>>(F, x, f, fscale, col);
            LAUNCH_CHECK(c);
        }
    return F2D_OK;
}

static int read_scalars(f2d_ctx *c, int first, int count) {
    F2D_CUDA(cudaMemcpyAsync(c->h_scal + first, c->d_scal + first, count * sizeof(double),
                             cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}

static int zero_unknowns(f2d_ctx *c, const FineView &F, double *x) {
    for (int col = 0; col < 2; col++) {
        k_smooth0<true><<<grd(F.nx, F.ny), blk(), 0, c->stream>>>(F, x, x, 0.0, col);
        LAUNCH_CHECK(c);
    }
    return F2D_OK;
}

int mg_solve(f2d_ctx *c, int which, const double *b, double bscale, double *x, int *iters_out,
             double *relres_out, const GuessSpec *guess) {
    if (which < 0 || which > 2) { set_error("solver id %d", which); return F2D_ERR_ARG; }
    Multigrid &M = c->mg[which];
    if (!M.built) { set_error("solver %d not built (call f2d_set_mask; Helmholtz needs a qg/rsw model)", which); return F2D_ERR_STATE; }
    if (iters_out) *iters_out = 0;
    if (relres_out) *relres_out = 0.0;
    if (M.n_global == 0) return op_fill(c, x);
    const FineView &F = M.fine;
    cudaStream_t st = c->stream;
    const double rtol = c->cfg.solver_rtol > 0 ? c->cfg.solver_rtol : 1e-12;
    const int maxit = c->cfg.solver_maxit > 0 ? c->cfg.solver_maxit : 100;
    const double fscale = -bscale;   // L = -A
    const dim3 nblk = cg_grid(c, F);
    const uint8_t *cg_open = allow_open_tiles() ? M.cg_open : nullptr;
    const bool singular = M.singular;
    const bool plain = (c->cfg.solver_kind & 1) != 0, unfused = (c->cfg.solver_kind & 2) != 0;
    const double N = M.n_global, inv_n = 1.0 / N;
    double *S = c->d_scal;
    int it = 0, first_ok = 0;   // first_ok: an unchecked iteration that had already converged
    double relres = 0.0;
    bool conv = false;
    auto projected = [&](double rr, double sum) { return singular ? std::max(rr - sum * sum * inv_n, 0.0) : rr; };

    GuessW GW;
    GW.n = guess ? guess->n : 0;
    for (int k = 0; k < 6; k++) { GW.g[k] = guess ? guess->g[k] : nullptr; GW.w[k] = guess ? guess->w[k] : 0.0; }
    for (int k = 0; k < GW.n; k++)
        if (GW.g[k] == x) { set_error("internal: the first guess reads the array it is written to"); return F2D_ERR_STATE; }
    if (plain) {
        if (GW.n > 0) {      // x = first guess (the residual this leaves in M.r is not used)
            k_cg_resid_guess<<<update_grid(c), dim3(CGX, CGTY), 0, st>>>(F, GW, x, b, fscale, M.r, c->d_part, c->d_count, S + S_RR, cg_open);
            LAUNCH_CHECK(c);
        }
        // plain V-cycle iteration on x itself
        for (it = 0; it <= maxit; it++) {
            k_cg_resid<<<nblk, 256, 0, st>>>(F, x, b, fscale, nullptr, c->d_part, c->d_count, S + S_RR);
            LAUNCH_CHECK(c);
            F2D_TRY(read_scalars(c, S_RR, 3));
            double ff = c->h_scal[S_FF];
            if (ff == 0.0) { F2D_TRY(zero_unknowns(c, F, x)); conv = true; relres = 0.0; break; }
            relres = std::sqrt(projected(c->h_scal[S_RR], c->h_scal[S_SUMR]) / ff);
            if (!(relres > rtol)) { conv = true; break; }
            if (it == maxit) break;
            if (unfused) F2D_TRY(vcycle_unfused(c, M, x, b, fscale, false));
            else F2D_TRY(vcycle_fused<double>(c, M, M.z, x, x, b, fscale, false, -1, false));
        }
    } else {
        // preconditioned conjugate gradients, M^-1 = one V-cycle from a zero guess
        const bool multi = singular && M.ncomp > 1 && M.comp != nullptr;   // per-component null space
        const bool lazy = singular && !multi;                              // one component: mean removed inside the kernels
        CompMeans CM;
        for (int k = 0; k < Multigrid::MAXCOMP; k++) CM.inv_n[k] = M.inv_nc[k];
        // v <- v - mean_c(v) on every component (multi only); rr_slot: keep the squared norm in step
        auto project_r = [&](int rr_slot) -> int {
            k_comp_sums<double><<<nblk, 256, 0, st>>>(F, M.r, M.comp, c->d_part, c->d_count, S + S_COMP);
            LAUNCH_CHECK(c);
            k_comp_sub<double><<<nblk, 256, 0, st>>>(F, M.r, M.comp, S, CM, rr_slot);
            LAUNCH_CHECK(c);
            return F2D_OK;
        };
        auto projected_l = [&](double rr, double sum) { return lazy ? std::max(rr - sum * sum * inv_n, 0.0) : rr; };
        double ff = 0.0;
        // r = b - A x, its norms on the host; returns the relative residual
        bool first_resid = true;
        auto true_residual = [&](double *rel) -> int {
            if (first_resid && GW.n > 0)      // the first guess is formed on the way
                k_cg_resid_guess<<<update_grid(c), dim3(CGX, CGTY), 0, st>>>(F, GW, x, b, fscale, M.r, c->d_part, c->d_count, S + S_RR, cg_open);
            else
                k_cg_resid<<<nblk, 256, 0, st>>>(F, x, b, fscale, M.r, c->d_part, c->d_count, S + S_RR);
            first_resid = false;
            LAUNCH_CHECK(c);
            F2D_TRY(dist_allreduce(c, S + S_RR, 3, false));
            if (multi) F2D_TRY(project_r(S_RR));
            if (c->dist.on) F2D_TRY(exchange_fine(c, M, M.r));
            F2D_TRY(read_scalars(c, S_RR, 3));
            if (multi) F2D_TRY(read_scalars(c, S_COMP, Multigrid::MAXCOMP));
            ff = c->h_scal[S_FF];
            *rel = ff > 0.0 ? std::sqrt(projected_l(c->h_scal[S_RR], c->h_scal[S_SUMR]) / ff) : 0.0;
            return F2D_OK;
        };
        // the first guess may come from anywhere; an extrapolated one is formed from arrays whose
        // ghost rows are current and is itself current on them
        if (c->dist.on && GW.n == 0) F2D_TRY(exchange_fine(c, M, x));
        F2D_TRY(true_residual(&relres));
        if (ff == 0.0) {   // b == 0: the solution is 0 (up to the Neumann null space)
            F2D_TRY(zero_unknowns(c, F, x));
            conv = true;
        } else {
            if (!(relres > rtol)) conv = true;
            if (singular) {
                // how far the right-hand side is from the range of the operator: |sum_c b| / sqrt(N_c b.b)
                // per component (r = b - A x has the component sums of b: the columns of A sum to zero)
                double worst = 0.0;
                if (multi) {
                    for (int k = 0; k < Multigrid::MAXCOMP; k++)
                        worst = std::max(worst, std::fabs(c->h_scal[S_COMP + k]) * std::sqrt(M.inv_nc[k] / ff));
                } else worst = std::fabs(c->h_scal[S_SUMR]) * std::sqrt(inv_n / ff);
                M.rhs_incompat = std::max(M.rhs_incompat, worst);
            }
        }
        double best = relres;
        static const bool debug = getenv("F2D_DEBUG") != nullptr;
        if (debug) fprintf(stderr, "[f2d] solve %d: initial relres %.3e (%d component%s)\n", which, relres, M.ncomp, M.ncomp > 1 ? "s" : "");
        float *pold = M.p, *pnew = M.p2;
        const int slot = lazy ? S_SUMR : -1;   // lazy projection r - mean(r)
        // One iteration = V-cycle + direction/apply + update (+ exchanges and
        // all-reduces): a fixed sequence of ~16 launches.  It is captured once
        // per parity class (first / odd / even iteration: the rz slot and the
        // p ping-pong alternate) into a CUDA graph and replayed, which takes
        // the launch and NCCL enqueue cost off the host.
        auto iteration = [&](int iter, float *po, float *pn) -> int {
            if (unfused) {
                if (lazy) { k_cg_project<<<nblk, 256, 0, st>>>(F, M.r, S, inv_n); LAUNCH_CHECK(c); }
                F2D_TRY(vcycle_unfused(c, M, M.z, M.r, 1.0, true));
                if (multi) {
                    k_comp_sums<double><<<nblk, 256, 0, st>>>(F, M.z, M.comp, c->d_part, c->d_count, S + S_COMP);
                    LAUNCH_CHECK(c);
                    k_comp_sub<double><<<nblk, 256, 0, st>>>(F, M.z, M.comp, S, CM, -1);
                    LAUNCH_CHECK(c);
                }
                k_dot2<<<nblk, 256, 0, st>>>(F, M.r, M.z, S, -1, inv_n, c->d_part, c->d_count, S + S_RZNEW);
                LAUNCH_CHECK(c);
                k_cg_dir_apply<double><<<dir_apply_grid<double>(c), dim3(CGX, CGTY), 0, st>>>(F, M.z, po, pn, S, iter, lazy ? 1 : 0, inv_n,
                                                             c->d_part, c->d_count, cg_open);
            } else {
                F2D_TRY((vcycle_fused<float>(c, M, M.zf, nullptr, M.zf2, M.r, 1.0, true, slot, true)));
                if (multi) {   // z <- z - mean_c(z); r.z is unchanged because r is projected
                    k_comp_sums<float><<<nblk, 256, 0, st>>>(F, M.zf2, M.comp, c->d_part, c->d_count, S + S_COMP);
                    LAUNCH_CHECK(c);
                    k_comp_sub<float><<<nblk, 256, 0, st>>>(F, M.zf2, M.comp, S, CM, -1);
                    LAUNCH_CHECK(c);
                }
                k_cg_dir_apply<float><<<dir_apply_grid<float>(c), dim3(CGX, CGTY), 0, st>>>(F, M.zf2, po, pn, S, iter, lazy ? 1 : 0, inv_n,
                                                            c->d_part, c->d_count, cg_open);
            }
            LAUNCH_CHECK(c);
            F2D_TRY(dist_allreduce(c, S + S_PQ, 1, false));
            k_cg_update_p<<<update_grid(c), dim3(CGX, CGTY), 0, st>>>(F, x, M.r, pn, S, S_RZ0 + (iter & 1), c->d_part, c->d_count, S + S_RR, cg_open);
            LAUNCH_CHECK(c);
            F2D_TRY(dist_allreduce(c, S + S_RR, 2, false));
            if (multi) F2D_TRY(project_r(S_RR));
            if (c->dist.on) F2D_TRY(exchange_fine(c, M, M.r));
            return F2D_OK;
        };
        static const bool no_graph = getenv("F2D_NO_GRAPH") != nullptr;
        const bool use_graph = !no_graph && !unfused && M.warm;
        // The previous solve of this system needed M.expect iterations (the
        // first guess makes consecutive solves alike): run that many back to
        // back, recording their residual norms asynchronously, and only then
        // take the host round trip that decides convergence.
        const int expected = M.expect;
        const int nocheck = (use_graph && !debug && M.expect > 1) ? std::min(M.expect - 1, 255) : 0;
        // CG decides convergence on the residual its recurrence carries.  A solve that
        // needed clearly more iterations than the previous one of the same system is
        // re-checked against the true residual b - A x, and restarted from it if the
        // recurrence had drifted (once: on ill-conditioned grids b - A x itself bottoms out a
        // few times above the recurrence, and a second restart could not do better).
        for (int attempt = 0; attempt < 2; attempt++) {
            int k = 0;                                   // iteration within this (re)start
            for (; !conv && it < maxit; it++, k++) {
                const int cls = k == 0 ? 0 : ((k & 1) ? 1 : 2);
                if (use_graph) {
                    Multigrid::IterGraph *G = nullptr;
                    for (Multigrid::IterGraph &g : M.graphs) if (g.x == x && g.cls == cls) { G = &g; break; }
                    if (!G) {
                        if (M.graphs.size() >= 96) {            // a caller cycling through many arrays: start over
                            for (Multigrid::IterGraph &g : M.graphs) cudaGraphExecDestroy(g.exec);
                            M.graphs.clear();
                        }
                        cudaGraph_t graph = nullptr;
                        const int64_t l0 = c->launches, e0 = c->exchanges;
                        F2D_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                        int rc = iteration(k, pold, pnew);
                        cudaError_t ce = cudaStreamEndCapture(st, &graph);
                        if (rc != F2D_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                        F2D_CUDA(ce);
                        cudaGraphExec_t exec = nullptr;
                        F2D_CUDA(cudaGraphInstantiate(&exec, graph, 0));
                        cudaGraphDestroy(graph);
                        M.graphs.push_back(Multigrid::IterGraph{x, cls, exec, c->launches - l0, c->exchanges - e0});
                        G = &M.graphs.back();
                        c->launches = l0; c->exchanges = e0;    // counted when the graph runs
                    }
                    F2D_CUDA(cudaGraphLaunch(G->exec, st));
                    c->launches += G->launches;
                    c->exchanges += G->exchanges;
                } else {
                    F2D_TRY(iteration(k, pold, pnew));
                }
                std::swap(pold, pnew);
                if (attempt == 0 && it < nocheck && it + 1 < maxit) {
                    F2D_CUDA(cudaMemcpyAsync(c->h_hist + 2 * it, S + S_RR, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
                    continue;
                }
                F2D_TRY(read_scalars(c, S_RR, 2));
                if (attempt == 0 && nocheck > 0 && it == nocheck)        // what the unchecked iterations did
                    for (int j = 0; j < nocheck && !first_ok; j++) {
                        double rj = std::sqrt(projected_l(c->h_hist[2 * j], c->h_hist[2 * j + 1]) / ff);
                        if (!(rj > rtol)) first_ok = j + 1;
                        best = std::min(best, rj);
                    }
                relres = std::sqrt(projected_l(c->h_scal[S_RR], c->h_scal[S_SUMR]) / ff);
                if (debug) fprintf(stderr, "[f2d]   it %d relres %.3e\n", it + 1, relres);
                if (!(relres > rtol)) { conv = true; it++; break; }
                best = std::min(best, relres);
                if (!(relres < 1e6 * best)) { it++; break; }   // diverging: give up, report
            }
            if (!conv || ff == 0.0 || attempt == 1) break;
            const bool unusual = it > (expected > 0 ? expected + 2 : 12);
            if (!unusual) break;
            double tr = 0.0;
            F2D_TRY(true_residual(&tr));
            if (debug) fprintf(stderr, "[f2d]   exit check after %d iterations: recurrence %.3e, true %.3e\n", it, relres, tr);
            if (!(tr > 4.0 * rtol)) break;      // rounding in b - A x itself sits a little above the recurrence
            relres = tr;
            conv = false;                        // drifted: restart CG from the true residual
            pold = M.p; pnew = M.p2;
        }
    }
    M.warm = true;      // every kernel attribute is set by now: later solves may capture graphs
    if (!plain) M.expect = conv ? (first_ok ? first_ok : it) : 0;
    c->nsolves++;
    c->niters += it;
    c->max_relres = std::max(c->max_relres, relres);
    if (iters_out) *iters_out = it;
    if (relres_out) *relres_out = relres;
    if (c->dist.on) {
        F2D_TRY(exchange_fine(c, M, x));
        F2D_TRY(p2p_check(c));
    }
    F2D_TRY(op_fill(c, x));
    if (!conv || !std::isfinite(relres)) {
        set_error("elliptic solve %d: relative residual %.3e after %d iterations (rtol %.1e)", which,
                  relres, it, rtol);
        return F2D_ERR_NOTCONV;
    }
    return F2D_OK;
}

int mg_apply(f2d_ctx *c, int which, const double *x, double *y) {
    if (which < 0 || which > 2) { set_error("solver id %d", which); return F2D_ERR_ARG; }
    Multigrid &M = c->mg[which];
    if (!M.built) { set_error("solver %d not built", which); return F2D_ERR_STATE; }
    F2D_CUDA(cudaMemsetAsync(y, 0, c->n * sizeof(double), c->stream));
    if (M.nunknown == 0) return F2D_OK;
    k_apply_A<<<grd(M.fine.nx, M.fine.ny), blk(), 0, c->stream>>>(M.fine, x, y);
    LAUNCH_CHECK(c);
    return F2D_OK;
}

// bench.py hook (see bench_step_kernel): fine-level multigrid / CG kernels of
// the cell-centre solver, timed alone.  Operates on the CG work vectors only.
int bench_mg_kernel(f2d_ctx *c, const char *name, int reps, float *ms, double *bytes) {
    // the solver the model's time step calls: pressure (cell centres) for the projecting
    // models, the vertex Helmholtz operator for qgrsw
    const int which = c->cfg.model == F2D_MODEL_QGRSW ? F2D_SOLVER_HELMHOLTZ : F2D_SOLVER_CENTERS;
    Multigrid &M = c->mg[which];
    if (!M.built || M.nunknown == 0) { set_error("solver not built"); return F2D_ERR_STATE; }
    const FineView &F = M.fine;
    std::string k(name);
    double npts = (double)F.ny * F.nx;
    dim3 g0 = grd(F.nx, F.ny);
    const dim3 nblk = cg_grid(c, F);
    for (int pass = 0; pass < 2; pass++) {
        int n = pass == 0 ? 2 : reps;
        if (pass == 1) F2D_CUDA(cudaEventRecord(c->ev0, c->stream));
        for (int r = 0; r < n; r++) {
            if (k == "mg.down0") {
                // fused down leg, level 0: R r (fp64) bits, W z (fp32) + b1 (fp32, 1/4 of the points)
                F2D_TRY((launch_down0<float, 2, true>(c, M, M.zf, M.zf, M.r, 1.0, -1)));
                *bytes = npts * (8 + 1 + 4 + 1);
            } else if (k == "mg.up0") {
                // fused up leg, level 0: R z (fp32) r (fp64) bits x1 (fp32, 1/4), W z2 (fp32)
                F2D_TRY((launch_up0<float, 2, true>(c, M, M.zf, M.zf2, M.r, 1.0, -1)));
                *bytes = npts * (4 + 8 + 1 + 1 + 4);
            } else if (k == "mg.down1") {
                // open tiles (regular coefficients) do not read cx, cy, dinv: R b code, W x + b2 (1/4)
                F2D_TRY((coarse_down<2>(c, M, 1)));
                *bytes = (double)M.lev[1].ny * M.lev[1].nx * (allow_open_tiles() ? (4 + 1 + 4 + 1) : (4 * 4 + 1 + 4 + 1));
            } else if (k == "mg.up1") {
                // R x b code + x2 of level 2 (1/4), W x2
                F2D_TRY((coarse_up<2>(c, M, 1)));
                *bytes = (double)M.lev[1].ny * M.lev[1].nx * (allow_open_tiles() ? (2 * 4 + 1 + 1 + 4) : (5 * 4 + 1 + 1 + 4));
            } else if (k == "mg.tail") {
                F2D_TRY(launch_tail(c, M));
                *bytes = 0;
            } else if (k == "mg.smooth_halfsweep") {
                // one colour: R x(other colour) f(own) W x(own), 1 mask byte per updated point
                k_smooth0<false><<<g0, blk(), 0, c->stream>>>(F, M.z, M.r, 1.0, r & 1);
                *bytes = npts * (1.5 * 8 + 0.5);
                LAUNCH_CHECK(c);
            } else if (k == "cg.dir_apply") {
                // R z2 (fp32) p (fp32) bits, W p2 (fp32)
                k_cg_dir_apply<float><<<dir_apply_grid<float>(c), dim3(CGX, CGTY), 0, c->stream>>>(F, M.zf2, M.p, M.p2, c->d_scal + 16, 0, 0, 0.0, c->d_part, c->d_count,
                                                                            allow_open_tiles() ? M.cg_open : nullptr);
                *bytes = npts * (3 * 4 + 1);
                LAUNCH_CHECK(c);
            } else if (k == "cg.update") {
                // R x r (fp64) p (fp32) bits, W x r
                k_cg_update_p<<<update_grid(c), dim3(CGX, CGTY), 0, c->stream>>>(F, M.z, M.r, M.p, c->d_scal + 16, S_RZ0, c->d_part, c->d_count, c->d_scal + 16 + S_TMP,
                                                                            allow_open_tiles() ? M.cg_open : nullptr);
                *bytes = npts * (4 * 8 + 4 + 1);
                LAUNCH_CHECK(c);
            } else { set_error("unknown kernel '%s'", name); return F2D_ERR_ARG; }
        }
    }
    F2D_CUDA(cudaEventRecord(c->ev1, c->stream));
    F2D_CUDA(cudaEventSynchronize(c->ev1));
    F2D_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    *ms /= reps;
    return F2D_OK;
}

}  // namespace f2d
