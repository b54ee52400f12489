// Shared-memory tile kernels of the multigrid V-cycle.
//
// One V-cycle leg on one level is ONE kernel:
//   k_mg_down : nu1 red-black Gauss-Seidel sweeps + residual + restriction
//               (R = P^T) to the next coarser level
//   k_mg_up   : prolongation of the coarse correction + nu2 sweeps (+ the CG
//               dot products r.z and 1.z on the fine level)
// instead of one kernel per half-sweep.  A CTA owns a window of WJ x 64 points
// in shared memory; each half-sweep is exact on a region one ring smaller than
// the previous one (the domain of dependence of red-black relaxation grows by
// one point per half-sweep), so the interior TJ x TI = (WJ-2H) x (64-2H)
// points come out bit-identical to grid-wide sweeps, whatever the tiling.  HBM
// traffic per leg drops from ~(4 nu + 3) array passes to one read of the
// inputs and one write of the outputs.
//
// Shared-memory layout: the two colours of the checkerboard are stored apart,
// [colour][row][k] with 32 entries per row, so a warp that relaxes one colour
// of one row touches 32 consecutive words of each operand: no bank conflicts.
//   row parity rp = colour of column 0 of that row
//   colour c point k sits at column b = 2k + (c ^ rp)
//   its W / E neighbours are (1-c, k + o - 1) and (1-c, k + o), o = c ^ rp
//   its S / N neighbours are (1-c, row -+ 1, k)
//
// A leg that reads x with a halo must not write it in place (a neighbouring CTA
// may already have stored its interior): xin / xout are distinct buffers; only
// the zero-guess down leg, which never reads x, may alias them.
//
// FINE  (level 0): fp64, operator from one byte of mask bits per point
// COARSE (l >= 1): CT (fp32) arrays of face couplings and inverse diagonals.
//
// OPEN tiles.  The tile kernels are instruction-issue bound, and half of their
// instructions are bounds tests, periodic wraps and mask-bit look-ups that only
// matter next to a wall.  A CTA whose whole window lies inside the array and
// whose mask bytes are all 0xFF (every point an unknown, every face open, every
// parent fluid -- the vast majority of tiles of any domain) takes a path
// without them: constant diagonal, constant prolongation weight, unconditional
// stores.  The decision is block-uniform (__syncthreads_and); both paths round
// identically (explicit _rn intrinsics below), so the result does not depend
// on which tiles are open, on the tiling, or on the slab decomposition.
#pragma once
#include "engine.cuh"
#include "reduce.cuh"

namespace f2d {

constexpr int TW = 64;    // window width in points
constexpr int TK = 32;    // points of one colour per window row
constexpr int TILE_THREADS = 512;
constexpr int TILE_WARPS = TILE_THREADS / 32;
// resident CTAs per SM the fp32 fine-level legs are compiled for (3 -> 40 registers)
#ifndef F2D_FINE_CTAS
#define F2D_FINE_CTAS 3
#endif

__host__ __device__ constexpr int halo_down(int nu, bool zero) { return zero ? 2 * nu + 1 : 2 * nu + 2; }
__host__ __device__ constexpr int halo_up(int nu) { return 2 * nu; }

template <typename T>
struct CoarseArrays {       // level l >= 1, halo-padded (ny+2) x pitch
    int ny, nx, pitch, periodic, periodic_y, dirichlet, pj_off;
    const T *cx, *cy, *dinv;
    const uint8_t *code;
    T cx0, cy0, dinv0;      // coefficients of a regular point (NB_REG): all faces open, no wall, no mask nearby
};

// sum of the four neighbour contributions / Gauss-Seidel update / residual with
// explicit roundings: the generic and the open-tile path give the same bits
__device__ __forceinline__ float nb_sum(float cx, float cy, float xw, float xe, float xs, float xn) {
    return __fmaf_rn(cx, __fadd_rn(xw, xe), __fmul_rn(cy, __fadd_rn(xs, xn)));
}
__device__ __forceinline__ double nb_sum(double cx, double cy, double xw, double xe, double xs, double xn) {
    return __fma_rn(cx, __dadd_rn(xw, xe), __dmul_rn(cy, __dadd_rn(xs, xn)));
}
// four couplings (coarse levels): w, e, s, n in this order in both paths
__device__ __forceinline__ float nb_sum4(float cw, float xw, float ce, float xe, float cs, float xs, float cn, float xn) {
    return __fmaf_rn(cw, xw, __fmaf_rn(ce, xe, __fmaf_rn(cs, xs, __fmul_rn(cn, xn))));
}
__device__ __forceinline__ double nb_sum4(double cw, double xw, double ce, double xe, double cs, double xs, double cn,
                                          double xn) {
    return __fma_rn(cw, xw, __fma_rn(ce, xe, __fma_rn(cs, xs, __dmul_rn(cn, xn))));
}
__device__ __forceinline__ float gs_new(float f, float off, float dinv) { return __fmul_rn(__fadd_rn(f, off), dinv); }
__device__ __forceinline__ double gs_new(double f, double off, double dinv) { return __dmul_rn(__dadd_rn(f, off), dinv); }
// (f - (diag x - off)) * w
__device__ __forceinline__ float res_val(float f, float diag, float x, float off, float w) {
    return __fmul_rn(__fsub_rn(f, __fsub_rn(__fmul_rn(diag, x), off)), w);
}
__device__ __forceinline__ double res_val(double f, double diag, double x, double off, double w) {
    return __dmul_rn(__dsub_rn(f, __dsub_rn(__dmul_rn(diag, x), off)), w);
}
__device__ __forceinline__ float prol_val(float x00, float xn0, float x0n, float xnn) {
    return __fadd_rn(__fmaf_rn(3.0f, __fadd_rn(xn0, x0n), __fmul_rn(9.0f, x00)), xnn);
}
__device__ __forceinline__ double prol_val(double x00, double xn0, double x0n, double xnn) {
    return __dadd_rn(__fma_rn(3.0, __dadd_rn(xn0, x0n), __dmul_rn(9.0, x00)), xnn);
}
__device__ __forceinline__ float mul_add(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double mul_add(double a, double b, double c) { return __fma_rn(a, b, c); }

// global -> shared without a register in between (LDGSTS): the copy is in flight
// while the thread goes on, which keeps the window fetch inside the register budget
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async sizes");
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int wrap_mod(int i, int n) {
    i %= n;
    return i < 0 ? i + n : i;
}
// a window is at most TW + a few columns wide: one wrap is enough unless the
// grid itself is narrower than the window
__device__ __forceinline__ int wrap_col(int i, int n) {
    if (n >= 2 * TW) {
        if (i < 0) i += n;
        else if (i >= n) i -= n;
        return i;
    }
    return wrap_mod(i, n);
}

// ---- level descriptors: how a window point maps to global memory ------------
// row(j): offset of logical (j, 0) or -1;  col(i): offset to add, or -1
struct FineLevel {
    FineView F;
    __device__ __forceinline__ int ny() const { return F.ny; }
    __device__ __forceinline__ int nx() const { return F.nx; }
    __device__ __forceinline__ int dirichlet() const { return F.dirichlet; }
    __device__ __forceinline__ int pj_off() const { return F.pj_off; }
    __device__ __forceinline__ bool owned(int j) const { return j >= F.jo0 && j < F.jo1; }
    __device__ __forceinline__ long row(int j) const {
        if (F.periodic_y) j = wrap_col(j, F.ny);
        int aj = F.oj + j;
        if (j < 0 || j >= F.ny || aj < 0 || aj >= F.n2) return -1;
        return (long)aj * F.n1 + F.oi;
    }
    __device__ __forceinline__ int col(int i) const {
        if (F.periodic) i = wrap_col(i, F.nx);
        else if (i < 0 || i >= F.nx) return -1;
        int ai = F.oi + i;
        return (ai < 0 || ai >= F.n1) ? -1 : i;
    }
    // the window [j0, j0+nj) x [i0, i0+ni) needs no bounds test and no wrap
    __device__ __forceinline__ bool inside(int j0, int i0, int nj, int ni) const {
        return j0 >= 0 && j0 + nj <= F.ny && F.oj + j0 >= 0 && F.oj + j0 + nj <= F.n2 &&
               i0 >= 0 && i0 + ni <= F.nx && F.oi + i0 >= 0 && F.oi + i0 + ni <= F.n1;
    }
    __device__ __forceinline__ long base(int j, int i) const { return (long)(F.oj + j) * F.n1 + F.oi + i; }
    __device__ __forceinline__ int stride() const { return F.n1; }
    __device__ __forceinline__ const uint8_t *bits_ptr() const { return F.nb; }
    static constexpr unsigned FULL = 0xFFu;     // unknown, four open faces, three fluid parents
};

template <typename T>
struct CoarseLevel {
    CoarseArrays<T> A;
    __device__ __forceinline__ int ny() const { return A.ny; }
    __device__ __forceinline__ int nx() const { return A.nx; }
    __device__ __forceinline__ int dirichlet() const { return A.dirichlet; }
    __device__ __forceinline__ int pj_off() const { return A.pj_off; }
    __device__ __forceinline__ bool owned(int j) const { return true; }
    __device__ __forceinline__ long row(int j) const {
        if (A.periodic_y) j = wrap_col(j, A.ny);
        if (j < 0 || j >= A.ny) return -1;
        return (long)(j + 1) * A.pitch + 1;
    }
    __device__ __forceinline__ int col(int i) const {
        if (A.periodic) return wrap_col(i, A.nx);
        return (i < 0 || i >= A.nx) ? -1 : i;
    }
    __device__ __forceinline__ bool inside(int j0, int i0, int nj, int ni) const {
        return j0 >= 0 && j0 + nj <= A.ny && i0 >= 0 && i0 + ni <= A.nx;
    }
    __device__ __forceinline__ long base(int j, int i) const { return (long)(j + 1) * A.pitch + 1 + i; }
    __device__ __forceinline__ int stride() const { return A.pitch; }
    __device__ __forceinline__ const uint8_t *bits_ptr() const { return A.code; }
    static constexpr unsigned FULL = NB_SELF | NB_REG | NB_PJ | NB_PI | NB_PJI;
};

// ---- shared-memory window -----------------------------------------------------
template <typename T, bool FINE, int WJ>
struct Window {
    T *X, *Fv;              // [2][WJ][TK]
    T *CX, *CY, *DI;        // COARSE only
    uint8_t *B;             // mask / parent bits
    T *tab_dinv, *tab_diag; // FINE only: 32 entries indexed by bits & 31
    T *tab_invw;            // 8 entries indexed by bits >> 5: 1 / prolongation normaliser
    T cxf, cyf;             // FINE couplings

    static constexpr int ROWS_PER_WARP = (WJ + TILE_WARPS - 1) / TILE_WARPS;

    __device__ __forceinline__ static int at(int c, int a, int k) { return (c * WJ + a) * TK + k; }

    __host__ __device__ static constexpr size_t bytes() {
        size_t n = (size_t)2 * WJ * TK;
        return (FINE ? 2 : 5) * n * sizeof(T) + n + 72 * sizeof(T);
    }

    __device__ void carve(unsigned char *smem) {
        size_t n = (size_t)2 * WJ * TK;
        T *p = reinterpret_cast<T *>(smem);
        X = p; p += n;
        Fv = p; p += n;
        CX = CY = DI = nullptr;
        if (!FINE) { CX = p; p += n; CY = p; p += n; DI = p; p += n; }
        tab_dinv = p; p += 32; tab_diag = p; p += 32; tab_invw = p; p += 8;
        B = reinterpret_cast<uint8_t *>(p);
    }

    // tables: FINE diagonal by open-face bits; prolongation normaliser by parent bits
    __device__ __forceinline__ void fill_tables(const FineView *F, int dirichlet) {
        int c = threadIdx.x;
        if (FINE && c < 32) {
            double d = 0.0;
            if (c & NB_SELF) {
                if (F->dirichlet) d = 2.0 * (F->cx + F->cy) + F->shift;
                else d = F->cx * (((c & NB_W) ? 1 : 0) + ((c & NB_E) ? 1 : 0)) +
                         F->cy * (((c & NB_S) ? 1 : 0) + ((c & NB_N) ? 1 : 0)) + F->shift;
            }
            tab_diag[c] = (T)d;
            tab_dinv[c] = d > 0.0 ? (T)(1.0 / d) : T(0);
        }
        if (c >= 32 && c < 40) {
            int b = c - 32;     // bit0 = NB_PJ, bit1 = NB_PI, bit2 = NB_PJI
            double w = dirichlet ? 16.0 : 9.0 + 3.0 * (b & 1) + 3.0 * ((b >> 1) & 1) + ((b >> 2) & 1);
            tab_invw[b] = (T)(1.0 / w);
        }
    }

    // one red-black half-sweep of colour `col` on the ring-`m` interior.
    // A warp owns rows a0, a0+16, ... which all have the parity of a0, so the
    // column offset o, the bounds test and every neighbour offset are hoisted;
    // the unrolled body is loads at constant offsets from one base pointer.
    template <bool NO_NEIGHBOURS>
    __device__ __forceinline__ void relax(int col, int m, int par0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const int a0 = m + warp;
        const int o = col ^ ((par0 + a0) & 1), b = 2 * k + o;
        if (b < m || b >= TW - m) return;
        const int p0 = at(col, a0, k), q0 = at(1 - col, a0, k);
        const int nrow = (WJ - m - a0 + TILE_WARPS - 1) / TILE_WARPS;
        constexpr int S = TILE_WARPS * TK;
        T *xp = X + p0;
        const T *xq = X + q0 + o, *fp = Fv + p0;
#pragma unroll
        for (int r = 0; r < ROWS_PER_WARP; r++) {
            if (r >= nrow) break;
            T off = T(0);
            if (!NO_NEIGHBOURS) {
                const T *q = xq + r * S;
                T xw = q[-1], xe = q[0], xs = q[-o - TK], xn = q[-o + TK];
                if constexpr (FINE) off = nb_sum(cxf, cyf, xw, xe, xs, xn);
                else off = nb_sum4(CX[p0 + r * S], xw, CX[q0 + o + r * S], xe, CY[p0 + r * S], xs, CY[q0 + TK + r * S], xn);
            }
            if constexpr (FINE) xp[r * S] = gs_new(fp[r * S], off, tab_dinv[B[p0 + r * S] & 31]);
            else xp[r * S] = gs_new(fp[r * S], off, DI[p0 + r * S]);
        }
    }

    // residual, pre-multiplied by 1/normaliser of the prolongation, into Fv (ring m)
    __device__ __forceinline__ void residual(int m, int par0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const int a0 = m + warp;
        const int rp = (par0 + a0) & 1;
        const int nrow = (WJ - m - a0 + TILE_WARPS - 1) / TILE_WARPS;
        constexpr int S = TILE_WARPS * TK;
#pragma unroll
        for (int col = 0; col < 2; col++) {
            const int o = col ^ rp, b = 2 * k + o;
            if (b < m || b >= TW - m) continue;
            const int p0 = at(col, a0, k), q0 = at(1 - col, a0, k) + o;
#pragma unroll
            for (int r = 0; r < ROWS_PER_WARP; r++) {
                if (r >= nrow) break;
                const int p = p0 + r * S;
                const T *q = X + q0 + r * S;
                T xw = q[-1], xe = q[0], xs = q[-o - TK], xn = q[-o + TK];
                uint8_t bits = B[p];
                T off, diag;
                if constexpr (FINE) {
                    off = nb_sum(cxf, cyf, xw, xe, xs, xn);
                    diag = tab_diag[bits & 31];
                } else {
                    off = nb_sum4(CX[p], xw, CX[q0 + r * S], xe, CY[p], xs, CY[q0 - o + TK + r * S], xn);
                    T di = DI[p];
                    diag = di != T(0) ? T(1) / di : T(0);
                }
                Fv[p] = (bits & NB_SELF) ? res_val(Fv[p], diag, X[p], off, tab_invw[bits >> 5]) : T(0);
            }
        }
    }

    // global -> shared: every load of a thread is issued before its first use
    template <bool LOAD_X, class Lev, typename TX, typename TF>
    __device__ __forceinline__ void load(const Lev &L, const TX *__restrict__ xin, const TF *__restrict__ fin,
                                         double fscale, double fshift, int wj0, int wi0, int par0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const int c0 = L.col(wi0 + 2 * k), c1 = L.col(wi0 + 2 * k + 1);
        constexpr int R = WJ / TILE_WARPS;
        static_assert(WJ % TILE_WARPS == 0, "window rows must be a multiple of the warp count");
        TF fv[R][2];
        TX xv[R][2];
        T cxv[R][2], cyv[R][2], div[R][2];
        uint8_t bits[R][2];
#pragma unroll
        for (int r = 0; r < R; r++) {
            long rb = L.row(wj0 + warp + r * TILE_WARPS);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int ch = h ? c1 : c0;
                bool ok = rb >= 0 && ch >= 0;
                long g = ok ? rb + ch : 0;
                fv[r][h] = TF(0); xv[r][h] = TX(0);
                cxv[r][h] = cyv[r][h] = div[r][h] = T(0);
                bits[r][h] = 0;
                if (ok) {
                    if constexpr (FINE) {
                        bits[r][h] = L.F.nb[g];
                        fv[r][h] = fin[g];
                        if (LOAD_X) xv[r][h] = xin[g];
                    } else {
                        bits[r][h] = L.A.code[g];
                        fv[r][h] = fin[g];
                        if (LOAD_X) xv[r][h] = xin[g];
                        cxv[r][h] = L.A.cx[g]; cyv[r][h] = L.A.cy[g]; div[r][h] = L.A.dinv[g];
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            int a = warp + r * TILE_WARPS, rp = (par0 + a) & 1;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int p = at(rp ^ h, a, k);
                bool self = bits[r][h] & NB_SELF;
                B[p] = bits[r][h];
                if constexpr (FINE) {
                    // masked entries of the field arrays may hold anything: select, do not multiply
                    // (the residual is scaled and shifted in fp64 before it is narrowed)
                    Fv[p] = self ? (T)__fma_rn(fscale, (double)fv[r][h], -fshift) : T(0);
                    X[p] = self ? (T)xv[r][h] : T(0);
                } else {
                    Fv[p] = (T)fv[r][h];
                    X[p] = (T)xv[r][h];
                    CX[p] = cxv[r][h]; CY[p] = cyv[r][h]; DI[p] = div[r][h];
                }
            }
        }
    }

    // ---- open tiles: fixed ownership, f and x in registers --------------------
    // (COARSE: cxf / cyf hold the level's regular couplings)
    // Thread (warp w, lane k) owns rows w, w+16, ... x columns 2k, 2k+1 for the
    // whole leg.  Its rows share one parity rp, so its colour-c points sit in
    // column 2k + (c ^ rp); registers are indexed [row][colour].  Of the four
    // neighbours of a point one is the thread's own point of the other colour
    // (a register), one belongs to the next lane and two to the next warps: three
    // shared-memory loads instead of five, none for f, no mask byte, no table.
    static constexpr int RO = WJ / TILE_WARPS;

    // the window without bounds tests or wraps (Lev::inside): global -> registers.
    // Only issues the loads; the caller does its set-up work before it looks at them.
    template <bool LOAD_X, class Lev, typename TX, typename TF>
    __device__ __forceinline__ void fetch_inside(const Lev &L, const TX *__restrict__ xin, const TF *__restrict__ fin,
                                                 int wj0, int wi0, TF (&fv)[RO][2], TX (&xv)[RO][2],
                                                 uint8_t (&bits)[RO][2]) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const uint8_t *__restrict__ nb = L.bits_ptr();
        const long g0 = L.base(wj0 + warp, wi0 + 2 * k);
        const long st = (long)TILE_WARPS * L.stride();
#pragma unroll
        for (int r = 0; r < RO; r++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long g = g0 + r * st + h;
                bits[r][h] = nb[g];
                fv[r][h] = fin[g];
                xv[r][h] = TX(0);
                if (LOAD_X) xv[r][h] = xin[g];
            }
        }
    }
    // x of the thread's points: global -> X directly (cp.async), no registers
    template <class Lev, typename TX>
    __device__ __forceinline__ void fetch_x_async(const Lev &L, const TX *__restrict__ xin, int wj0, int wi0, int par0) {
        static_assert(sizeof(TX) == sizeof(T), "x is stored in the type it is relaxed in");
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const TX *src = xin + L.base(wj0 + warp, wi0 + 2 * k);
        const long st = (long)TILE_WARPS * L.stride();
        const int rp = (par0 + warp) & 1;
#pragma unroll
        for (int r = 0; r < RO; r++) {
            cp_async<sizeof(T)>(X + at(rp, warp + r * TILE_WARPS, k), src + r * st);
            cp_async<sizeof(T)>(X + at(rp ^ 1, warp + r * TILE_WARPS, k), src + r * st + 1);
        }
    }
    // every mask byte this thread fetched has all the bits of `full`
    __device__ __forceinline__ static bool all_open(const uint8_t (&bits)[RO][2], unsigned full) {
        unsigned all = 0xFFu;
#pragma unroll
        for (int r = 0; r < RO; r++) all &= bits[r][0] & bits[r][1];
        return (all & full) == full;
    }

    // registers -> shared memory, exactly what load() leaves (an inside tile that is not open)
    // X_ASYNC: x already sits in X (fetch_x_async): only clear the masked entries
    template <bool X_ASYNC, typename TX, typename TF>
    __device__ __forceinline__ void stash(const TF (&fv)[RO][2], const TX (&xv)[RO][2], const uint8_t (&bits)[RO][2],
                                          double fscale, double fshift, int par0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RO; r++) {
            int a = warp + r * TILE_WARPS, rp = (par0 + a) & 1;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int p = at(rp ^ h, a, k);
                bool self = bits[r][h] & NB_SELF;
                B[p] = bits[r][h];
                Fv[p] = self ? (T)__fma_rn(fscale, (double)fv[r][h], -fshift) : T(0);
                if (!X_ASYNC) X[p] = self ? (T)xv[r][h] : T(0);
                else if (!self) X[p] = T(0);
            }
        }
    }

    // [row][column] -> [row][colour] registers (+ scale/shift/narrow of f)
    // X_ASYNC: x comes from X, where fetch_x_async put it (already by colour)
    template <bool X_ASYNC, typename TX, typename TF>
    __device__ __forceinline__ void own(const TF (&fv)[RO][2], const TX (&xv)[RO][2], double fscale, double fshift,
                                        int rp, T (&fr)[RO][2], T (&xr)[RO][2]) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RO; r++) {
            T f0 = (T)__fma_rn(fscale, (double)fv[r][0], -fshift), f1 = (T)__fma_rn(fscale, (double)fv[r][1], -fshift);
            fr[r][0] = rp ? f1 : f0; fr[r][1] = rp ? f0 : f1;
            if (X_ASYNC) {
                xr[r][0] = X[at(0, warp + r * TILE_WARPS, k)];
                xr[r][1] = X[at(1, warp + r * TILE_WARPS, k)];
            } else {
                T x0 = (T)xv[r][0], x1 = (T)xv[r][1];
                xr[r][0] = rp ? x1 : x0; xr[r][1] = rp ? x0 : x1;
            }
        }
    }

    // xr -> X (both colours, whole window)
    __device__ __forceinline__ void publish(const T (&xr)[RO][2]) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RO; r++) {
            X[at(0, warp + r * TILE_WARPS, k)] = xr[r][0];
            X[at(1, warp + r * TILE_WARPS, k)] = xr[r][1];
        }
    }

    // half-sweep of colour c (a constant once the caller's loop is unrolled) on ring m
    template <bool NO_NEIGHBOURS>
    __device__ __forceinline__ void relax_open(int c, int m, int rp, const T (&fr)[RO][2], T (&xr)[RO][2], T dinv0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        const int o = c ^ rp, b = 2 * k + o;
        if (b < m || b >= TW - m) return;
        constexpr int S = TILE_WARPS * TK;
        T *xp = X + at(c, warp, k);
        const T *xq = X + at(1 - c, warp, k);
        const int side = 2 * o - 1;     // the W (o = 0) or E (o = 1) neighbour is another lane's point
#pragma unroll
        for (int r = 0; r < RO; r++) {
            const int a = warp + r * TILE_WARPS;
            if (a < m || a >= WJ - m) continue;
            T off = T(0);
            if (!NO_NEIGHBOURS) {
                const T *q = xq + r * S;
                T xl = q[side], xs = q[-TK], xn = q[TK], mine = xr[r][1 - c];
                if constexpr (FINE) off = nb_sum(cxf, cyf, mine, xl, xs, xn);     // xw + xe commutes: no need to know which is which
                else off = nb_sum4(cxf, o ? mine : xl, cxf, o ? xl : mine, cyf, xs, cyf, xn);
            }
            T v = gs_new(fr[r][c], off, dinv0);
            xr[r][c] = v;
            xp[r * S] = v;
        }
    }

    // residual (x 1/16) of both colours on ring m into Fv
    __device__ __forceinline__ void residual_open(int m, int rp, const T (&fr)[RO][2], const T (&xr)[RO][2], T diag0,
                                                  T invw0) {
        const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
        constexpr int S = TILE_WARPS * TK;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int o = c ^ rp, b = 2 * k + o;
            if (b < m || b >= TW - m) continue;
            const T *xq = X + at(1 - c, warp, k);
            T *fp = Fv + at(c, warp, k);
            const int side = 2 * o - 1;
#pragma unroll
            for (int r = 0; r < RO; r++) {
                const int a = warp + r * TILE_WARPS;
                if (a < m || a >= WJ - m) continue;
                const T *q = xq + r * S;
                T xl = q[side], xs = q[-TK], xn = q[TK], mine = xr[r][1 - c];
                T off;
                if constexpr (FINE) off = nb_sum(cxf, cyf, mine, xl, xs, xn);
                else off = nb_sum4(cxf, o ? mine : xl, cxf, o ? xl : mine, cyf, xs, cyf, xn);
                fp[r * S] = res_val(fr[r][c], diag0, xr[r][c], off, invw0);
            }
        }
    }
};

// ---------------------------------------------------------------------------
// DOWN leg.   x <- NU sweeps (R,B) on  L x = f ;  bc <- P^T (f - L x)
//   FINE : f = fscale * fin - fshift (fshift = mean of fin when the operator is
//          singular and CG asks for the projected residual), x in the (n2,n1)
//          array; ZERO = first guess is zero (x is only written).
//   TC   : element type of the next coarser level
// ---------------------------------------------------------------------------
//   TX / TF: storage types of x and f in global memory (the fine level of the
//          CG preconditioner relaxes in fp32 on an fp64 residual)
struct DownArgs {
    int nyc, nxc, pitchc;   // next coarser level
    int allow_open;         // 0: every tile takes the generic path (F2D_NO_OPEN, A/B tests)
};

// restriction R = P^T of the residual in W.Fv: weights (1,3,3,1) x (1,3,3,1) over
// the 4 x 4 fine cells around the aggregate.  In the colour-separated layout the
// four columns of a row are (cA,kA) (cB,kB) (cA,kC) (cB,kD) with cB = 1 - cA, and
// cA flips from one row to the next: four loads at fixed offsets per row.  A warp
// takes one coarse row at a time (lane = coarse column): consecutive lanes read
// consecutive words, no bank conflicts and no division.
template <typename T, typename TC, bool FINE, int WJ, int H, class Lev>
__device__ __forceinline__ void restrict_tile(const Window<T, FINE, WJ> &W, const Lev &L, const DownArgs &A,
                                              TC *__restrict__ bc, int tj0, int ti0, int par0) {
    constexpr int TJ = WJ - 2 * H, TI = TW - 2 * H;
    static_assert(TI / 2 <= 32, "one lane per coarse column");
    const int warp = threadIdx.x >> 5, ci = threadIdx.x & 31;
    if (ci >= TI / 2) return;
    const int I = ti0 / 2 + ci;
    if (I >= A.nxc) return;
    const int b0 = 2 * ci + H;                             // = 2I - wi0
    const int kA = (b0 - 1) >> 1, kB = b0 >> 1, kC = (b0 + 1) >> 1, kD = (b0 + 2) >> 1;
    constexpr int PL = WJ * TK;                            // one colour plane
    for (int cj = warp; cj < TJ / 2; cj += TILE_WARPS) {
        int Jc = tj0 / 2 + cj + L.pj_off();                // aggregate of fine rows 2J, 2J+1: its row in the coarse array
        if (Jc >= A.nyc) break;
        const int a0 = 2 * cj + H;                         // = 2J - wj0
        int cA = (par0 + a0 - 1 + b0 - 1) & 1;             // colour of (a0-1, b0-1)
        const T *row = W.Fv + (a0 - 1) * TK;
        T acc = T(0);
#pragma unroll
        for (int da = 0; da < 4; da++) {
            const T *pa = row + cA * PL, *pb = row + (1 - cA) * PL;
            T v = (pa[kA] + pb[kD]) + T(3) * (pb[kB] + pa[kC]);
            acc += (da == 1 || da == 2) ? T(3) * v : v;
            row += TK;
            cA ^= 1;
        }
        bc[(long)(Jc + 1) * A.pitchc + I + 1] = (TC)acc;
    }
}

// generic path: everything after the window is loaded
template <typename T, typename TX, typename TC, bool FINE, bool ZERO, int NU, int WJ, class Lev>
__device__ __forceinline__ void down_body(Window<T, FINE, WJ> &W, const Lev &L, TX *__restrict__ xout,
                                          const DownArgs &A, TC *__restrict__ bc, int tj0, int ti0) {
    constexpr int H = halo_down(NU, ZERO);
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    // ---- NU sweeps, red then black
#pragma unroll
    for (int hs = 0; hs < 2 * NU; hs++) {
        if (ZERO && hs == 0) W.template relax<true>(0, 0, par0);
        else W.template relax<false>(hs & 1, ZERO ? hs : hs + 1, par0);
        __syncthreads();
    }
    // ---- residual on the tile +- 1, then write x and the restricted residual
    W.residual(H - 1, par0);
    {
        const int c0 = L.col(wi0 + 2 * k), c1 = L.col(wi0 + 2 * k + 1);
#pragma unroll
        for (int r = 0; r < Window<T, FINE, WJ>::ROWS_PER_WARP; r++) {
            int a = H + warp + r * TILE_WARPS;
            if (a >= WJ - H) break;
            int rp = (par0 + a) & 1, j = wj0 + a;
            if (j >= L.ny()) continue;             // (y-periodic images are another tile's)
            long rb = L.row(j);
            if (rb < 0) continue;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int b = 2 * k + h, p = W.at(rp ^ h, a, k), ch = h ? c1 : c0;
                if (b < H || b >= TW - H || ch < 0) continue;
                if (!(W.B[p] & NB_SELF)) continue;
                if (wi0 + b >= L.nx()) continue;   // periodic images are another tile's
                xout[rb + ch] = (TX)W.X[p];
            }
        }
    }
    __syncthreads();
    restrict_tile<T, TC, FINE, WJ, H>(W, L, A, bc, tj0, ti0, par0);
}

// open tile: registers fr / xr hold the thread's 2 x RO points
template <typename T>
struct OpenConst { T dinv0, diag0, invw0; };

// regular-point constants: FINE from the tables, COARSE from the level (the generic
// path's expressions, so that both give the same bits)
template <typename T, bool FINE, int WJ, class Lev>
__device__ __forceinline__ OpenConst<T> open_const(Window<T, FINE, WJ> &W, const Lev &L) {
    OpenConst<T> K;
    if constexpr (FINE) {
        K.dinv0 = W.tab_dinv[31]; K.diag0 = W.tab_diag[31];
    } else {
        W.cxf = L.A.cx0; W.cyf = L.A.cy0;
        K.dinv0 = L.A.dinv0;
        K.diag0 = K.dinv0 != T(0) ? T(1) / K.dinv0 : T(0);
    }
    K.invw0 = W.tab_invw[7];
    return K;
}

template <typename T, typename TX, typename TC, bool FINE, bool ZERO, int NU, int WJ, class Lev>
__device__ __forceinline__ void down_open(Window<T, FINE, WJ> &W, const Lev &L, TX *__restrict__ xout,
                                          const DownArgs &A, TC *__restrict__ bc, int tj0, int ti0,
                                          const T (&fr)[WJ / TILE_WARPS][2], T (&xr)[WJ / TILE_WARPS][2]) {
    constexpr int H = halo_down(NU, ZERO), RO = WJ / TILE_WARPS;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int rp = (par0 + warp) & 1;
    const OpenConst<T> K = open_const<T, FINE, WJ>(W, L);
    const T dinv0 = K.dinv0, diag0 = K.diag0, invw0 = K.invw0;
    if (!ZERO) { W.publish(xr); __syncthreads(); }
#pragma unroll
    for (int hs = 0; hs < 2 * NU; hs++) {
        if (ZERO && hs == 0) W.template relax_open<true>(0, 0, rp, fr, xr, dinv0);
        else W.template relax_open<false>(hs & 1, ZERO ? hs : hs + 1, rp, fr, xr, dinv0);
        __syncthreads();
    }
    W.residual_open(H - 1, rp, fr, xr, diag0, invw0);
    {
        const long g0 = L.base(wj0 + warp, wi0 + 2 * k);
        const long st = (long)TILE_WARPS * L.stride();
#pragma unroll
        for (int r = 0; r < RO; r++) {
            const int a = warp + r * TILE_WARPS;
            if (a < H || a >= WJ - H) continue;
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const int h = c ^ rp, b = 2 * k + h;
                if (b < H || b >= TW - H) continue;
                xout[g0 + r * st + h] = (TX)xr[r][c];
            }
        }
    }
    __syncthreads();
    restrict_tile<T, TC, FINE, WJ, H>(W, L, A, bc, tj0, ti0, par0);
}

template <typename T, typename TX, typename TF, typename TC, bool FINE, bool ZERO, int NU, int WJ, class Lev>
__global__ void __launch_bounds__(TILE_THREADS, (FINE && sizeof(T) == 4) ? F2D_FINE_CTAS : 2)
k_mg_down(Lev L, const TX *__restrict__ xin, TX *__restrict__ xout, const TF *__restrict__ fin, double fscale,
          const double *__restrict__ scal, int sumr_slot, double inv_n, DownArgs A, TC *__restrict__ bc) {
    constexpr int H = halo_down(NU, ZERO);
    constexpr int TJ = WJ - 2 * H, TI = TW - 2 * H;
    extern __shared__ __align__(16) unsigned char smem[];
    Window<T, FINE, WJ> W;
    W.carve(smem);
    const int tj0 = blockIdx.y * TJ, ti0 = blockIdx.x * TI;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    double fshift = 0.0;
    {
        // every global load is issued before the set-up work: the tables (an fp64
        // division) and the mean are computed while the window is in flight
        constexpr int RO = WJ / TILE_WARPS;
        const bool inside = A.allow_open && L.inside(wj0, wi0, WJ, TW);     // block-uniform
        TF fv[RO][2];
        TX xv[RO][2];
        uint8_t bits[RO][2];
        if (FINE && sumr_slot >= 0) fshift = scal[sumr_slot];
        if (inside) W.template fetch_inside<!ZERO>(L, xin, fin, wj0, wi0, fv, xv, bits);
        if constexpr (FINE) {
            W.cxf = (T)L.F.cx; W.cyf = (T)L.F.cy;
            W.fill_tables(&L.F, L.dirichlet());
        } else W.fill_tables(nullptr, L.dirichlet());
        fshift *= inv_n;
        if (inside) {
            if (__syncthreads_and(W.all_open(bits, Lev::FULL))) {
                T fr[RO][2], xr[RO][2];
                W.template own<false>(fv, xv, fscale, fshift, (par0 + (threadIdx.x >> 5)) & 1, fr, xr);
                down_open<T, TX, TC, FINE, ZERO, NU, WJ>(W, L, xout, A, bc, tj0, ti0, fr, xr);
                return;
            }
            // inside but not open: FINE has everything it needs in registers, COARSE
            // fetches the coefficient arrays as well
            if constexpr (FINE) W.template stash<false>(fv, xv, bits, fscale, fshift, par0);
            else W.template load<!ZERO>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
        } else {
            W.template load<!ZERO>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
        }
    }
    __syncthreads();
    down_body<T, TX, TC, FINE, ZERO, NU, WJ>(W, L, xout, A, bc, tj0, ti0);
}

// ---------------------------------------------------------------------------
// UP leg.   x <- x + P xc ;  NU sweeps (B,R) ;  [DOT: out = (sum f x, sum x)]
// ---------------------------------------------------------------------------
struct UpArgs {
    int nyc, nxc, pitchc, periodic_c;
    int allow_open;
    int periodic_yc;
};

template <typename T, typename TX, typename TF, typename TC, bool FINE, bool DOT, int NU, int WJ, class Lev>
__device__ __forceinline__ void up_body(Window<T, FINE, WJ> &W, const Lev &L, const TC *XC, TX *__restrict__ xout,
                                        const TF *__restrict__ fin, double fscale, double fshift, int tj0, int ti0,
                                        double (&acc)[2]) {
    constexpr int H = halo_up(NU);
    constexpr int CI = TW / 2 + 3;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int cj0 = (wj0 >> 1) - 1, ci0 = (wi0 >> 1) - 1;
    // ---- prolongation on the whole window
#pragma unroll
    for (int r = 0; r < WJ / TILE_WARPS; r++) {
        int a = warp + r * TILE_WARPS;
        int rp = (par0 + a) & 1, j = wj0 + a;
        int J0 = (j >> 1) - cj0, Jn = J0 + ((j & 1) ? 1 : -1);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int b = 2 * k + h, p = W.at(rp ^ h, a, k);
            uint8_t bits = W.B[p];
            if (!(bits & NB_SELF)) continue;
            int i = wi0 + b;
            int I0 = (i >> 1) - ci0, In = I0 + ((i & 1) ? 1 : -1);
            T v = prol_val((T)XC[J0 * CI + I0], (T)XC[Jn * CI + I0], (T)XC[J0 * CI + In], (T)XC[Jn * CI + In]);
            W.X[p] = mul_add(v, W.tab_invw[bits >> 5], W.X[p]);
        }
    }
    __syncthreads();
    // ---- NU sweeps, black then red
#pragma unroll
    for (int hs = 0; hs < 2 * NU; hs++) {
        W.template relax<false>(1 - (hs & 1), hs + 1, par0);
        __syncthreads();
    }
    // ---- write the interior (+ dots)
    const int c0 = L.col(wi0 + 2 * k), c1 = L.col(wi0 + 2 * k + 1);
#pragma unroll
    for (int r = 0; r < Window<T, FINE, WJ>::ROWS_PER_WARP; r++) {
        int a = H + warp + r * TILE_WARPS;
        if (a >= WJ - H) break;
        int rp = (par0 + a) & 1, j = wj0 + a;
        if (j >= L.ny()) continue;                 // (y-periodic images are another tile's)
        long rb = L.row(j);
        if (rb < 0) continue;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            int b = 2 * k + h, p = W.at(rp ^ h, a, k), ch = h ? c1 : c0;
            if (b < H || b >= TW - H || ch < 0) continue;
            if (!(W.B[p] & NB_SELF)) continue;
            if (wi0 + b >= L.nx()) continue;   // periodic images are another tile's
            T xv = W.X[p];
            xout[rb + ch] = (TX)xv;
            if (DOT && L.owned(j)) {
                // the dot uses the fp64 residual, not its narrowed copy in shared memory
                double fv = __fma_rn(fscale, (double)fin[rb + ch], -fshift);
                acc[0] = __fma_rn(fv, (double)xv, acc[0]); acc[1] += (double)xv;
            }
        }
    }
}

// open tile: registers fr / xr hold the thread's 2 x RO points
template <typename T, typename TX, typename TF, typename TC, bool FINE, bool DOT, int NU, int WJ, class Lev>
__device__ __forceinline__ void up_open(Window<T, FINE, WJ> &W, const Lev &L, const TC *XC, TX *__restrict__ xout,
                                        const TF *__restrict__ fin, double fscale, double fshift, int tj0,
                                        int ti0, const T (&fr)[WJ / TILE_WARPS][2], T (&xr)[WJ / TILE_WARPS][2],
                                        double (&acc)[2]) {
    constexpr int H = halo_up(NU), RO = WJ / TILE_WARPS;
    constexpr int CI = TW / 2 + 3;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    const int rp = (par0 + warp) & 1;
    const int cj0 = (wj0 >> 1) - 1, ci0 = (wi0 >> 1) - 1;
    const OpenConst<T> K = open_const<T, FINE, WJ>(W, L);
    const T dinv0 = K.dinv0, invw0 = K.invw0;
    // ---- prolongation of the thread's own points
    {
        // columns 2k (h = 0) and 2k+1 (h = 1): own parent I0, the other parent one column to the side
        const int i0 = wi0 + 2 * k;
        const int I0a = (i0 >> 1) - ci0, Ina = I0a + ((i0 & 1) ? 1 : -1);
        const int I0b = ((i0 + 1) >> 1) - ci0, Inb = I0b + (((i0 + 1) & 1) ? 1 : -1);
#pragma unroll
        for (int r = 0; r < RO; r++) {
            const int j = wj0 + warp + r * TILE_WARPS;
            const int J0 = (j >> 1) - cj0, Jn = J0 + ((j & 1) ? 1 : -1);
            const TC *r0 = XC + J0 * CI, *rn = XC + Jn * CI;
            T va = prol_val((T)r0[I0a], (T)rn[I0a], (T)r0[Ina], (T)rn[Ina]);
            T vb = prol_val((T)r0[I0b], (T)rn[I0b], (T)r0[Inb], (T)rn[Inb]);
            // column h holds colour h ^ rp
            T v0 = rp ? vb : va, v1 = rp ? va : vb;
            xr[r][0] = mul_add(v0, invw0, xr[r][0]);
            xr[r][1] = mul_add(v1, invw0, xr[r][1]);
        }
    }
    W.publish(xr);
    __syncthreads();
    // ---- NU sweeps, black then red
#pragma unroll
    for (int hs = 0; hs < 2 * NU; hs++) {
        W.template relax_open<false>(1 - (hs & 1), hs + 1, rp, fr, xr, dinv0);
        __syncthreads();
    }
    // ---- write the interior (+ dots): from shared memory, with the generic
    // path's thread-to-row map, so that a CTA's partial sums (and with them the
    // CG scalars) do not depend on which path it took.  The fp64 residual of the
    // dots is re-read in one batch (all loads in flight) before anything is stored.
    constexpr int RW = Window<T, FINE, WJ>::ROWS_PER_WARP;
    const long g0 = L.base(wj0 + H + warp, wi0 + 2 * k);
    const long st = (long)TILE_WARPS * L.stride();
    TF fq[RW][2];
    if (DOT) {
#pragma unroll
        for (int r = 0; r < RW; r++) {
            const bool rowok = H + warp + r * TILE_WARPS < WJ - H;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int b = 2 * k + h;
                fq[r][h] = (rowok && b >= H && b < TW - H) ? fin[g0 + r * st + h] : TF(0);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RW; r++) {
        const int a = H + warp + r * TILE_WARPS;
        if (a >= WJ - H) break;
        const int rpa = (par0 + a) & 1;
        const bool mine = L.owned(wj0 + a);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int b = 2 * k + h;
            if (b < H || b >= TW - H) continue;
            const T xv = W.X[W.at(rpa ^ h, a, k)];
            xout[g0 + r * st + h] = (TX)xv;
            if (DOT && mine) {
                double f = __fma_rn(fscale, (double)fq[r][h], -fshift);
                acc[0] = __fma_rn(f, (double)xv, acc[0]); acc[1] += (double)xv;
            }
        }
    }
}

template <typename T, typename TX, typename TF, typename TC, bool FINE, bool DOT, int NU, int WJ, class Lev>
__global__ void __launch_bounds__(TILE_THREADS, (FINE && sizeof(T) == 4) ? F2D_FINE_CTAS : 2)
k_mg_up(Lev L, const TX *__restrict__ xin, TX *__restrict__ xout, const TF *__restrict__ fin, double fscale,
        const double *__restrict__ scal, int sumr_slot, double inv_n, UpArgs A, const TC *__restrict__ xc,
        double *part, unsigned int *count, double *out) {
    constexpr int H = halo_up(NU);
    constexpr int TJ = WJ - 2 * H, TI = TW - 2 * H;
    constexpr int CJ = WJ / 2 + 3, CI = TW / 2 + 3;
    extern __shared__ __align__(16) unsigned char smem[];
    Window<T, FINE, WJ> W;
    W.carve(smem);
    TC *XC = reinterpret_cast<TC *>(smem + ((Window<T, FINE, WJ>::bytes() + 15) & ~size_t(15)));
    const int tj0 = blockIdx.y * TJ, ti0 = blockIdx.x * TI;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    double fshift = 0.0;
    if (FINE && sumr_slot >= 0) fshift = scal[sumr_slot];
    // ---- coarse window: global -> shared asynchronously (no registers held)
    const int cj0 = (wj0 >> 1) - 1, ci0 = (wi0 >> 1) - 1;
    const int Jc0 = cj0 + L.pj_off();
    if (Jc0 >= 0 && Jc0 + CJ <= A.nyc && ci0 >= 0 && ci0 + CI <= A.nxc) {   // block-uniform: no bounds, no wrap
        const TC *src = xc + (long)(Jc0 + 1) * A.pitchc + ci0 + 1;
        for (int t = threadIdx.x; t < CJ * CI; t += TILE_THREADS) {
            int a = t / CI, b = t - a * CI;
            cp_async<sizeof(TC)>(XC + t, src + a * A.pitchc + b);
        }
    } else {
        for (int t = threadIdx.x; t < CJ * CI; t += TILE_THREADS) {
            int a = t / CI, b = t - a * CI;
            int J = Jc0 + a, I = ci0 + b;
            if (A.periodic_yc) J = wrap_col(J, A.nyc);
            bool ok = J >= 0 && J < A.nyc;
            if (ok) {
                if (A.periodic_c) I = wrap_col(I, A.nxc);
                ok = I >= 0 && I < A.nxc;
            }
            if (ok) cp_async<sizeof(TC)>(XC + t, xc + (long)(J + 1) * A.pitchc + I + 1);
            else XC[t] = TC(0);
        }
    }
    double acc[2] = {0.0, 0.0};
    bool done = false;
    {
        constexpr int RO = WJ / TILE_WARPS;
        const bool inside = A.allow_open && L.inside(wj0, wi0, WJ, TW);     // block-uniform
        TF fv[RO][2];
        TX xv[RO][2];
        uint8_t bits[RO][2];
        if (inside) {
            W.fetch_x_async(L, xin, wj0, wi0, par0);
            W.template fetch_inside<false>(L, xin, fin, wj0, wi0, fv, xv, bits);
        }
        // set-up work while the window is in flight
        if constexpr (FINE) {
            W.cxf = (T)L.F.cx; W.cyf = (T)L.F.cy;
            W.fill_tables(&L.F, L.dirichlet());
        } else W.fill_tables(nullptr, L.dirichlet());
        fshift *= inv_n;
        if (inside) {
            const bool mine = W.all_open(bits, Lev::FULL);
            cp_async_wait_all();
            if (__syncthreads_and(mine)) {
                T fr[RO][2], xr[RO][2];
                W.template own<true>(fv, xv, fscale, fshift, (par0 + (threadIdx.x >> 5)) & 1, fr, xr);
                up_open<T, TX, TF, TC, FINE, DOT, NU, WJ>(W, L, XC, xout, fin, fscale, fshift, tj0, ti0, fr, xr, acc);
                done = true;
            } else {
                if constexpr (FINE) W.template stash<true>(fv, xv, bits, fscale, fshift, par0);
                else W.template load<true>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
            }
        } else {
            W.template load<true>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
            cp_async_wait_all();
        }
    }
    if (!done) {
        __syncthreads();
        up_body<T, TX, TF, TC, FINE, DOT, NU, WJ>(W, L, XC, xout, fin, fscale, fshift, tj0, ti0, acc);
    }
    if (DOT) block_partials<OpSum, 2>(acc, part);     // folded by k_fold_partials (launch_up0)
}

}  // namespace f2d
