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
#pragma once
#include "engine.cuh"
#include "reduce.cuh"

namespace f2d {

constexpr int TW = 64;    // window width in points
constexpr int TK = 32;    // points of one colour per window row
constexpr int TILE_THREADS = 512;
constexpr int TILE_WARPS = TILE_THREADS / 32;

__host__ __device__ constexpr int halo_down(int nu, bool zero) { return zero ? 2 * nu + 1 : 2 * nu + 2; }
__host__ __device__ constexpr int halo_up(int nu) { return 2 * nu; }

template <typename T>
struct CoarseArrays {       // level l >= 1, halo-padded (ny+2) x pitch
    int ny, nx, pitch, periodic, dirichlet, pj_off;
    const T *cx, *cy, *dinv;
    const uint8_t *code;
};

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
        if (j < 0 || j >= A.ny) return -1;
        return (long)(j + 1) * A.pitch + 1;
    }
    __device__ __forceinline__ int col(int i) const {
        if (A.periodic) return wrap_col(i, A.nx);
        return (i < 0 || i >= A.nx) ? -1 : i;
    }
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
            T acc = fp[r * S];
            if (!NO_NEIGHBOURS) {
                const T *q = xq + r * S;
                T xw = q[-1], xe = q[0], xs = q[-o - TK], xn = q[-o + TK];
                if constexpr (FINE) acc += cxf * (xw + xe) + cyf * (xs + xn);
                else acc += CX[p0 + r * S] * xw + CX[q0 + o + r * S] * xe + CY[p0 + r * S] * xs + CY[q0 + TK + r * S] * xn;
            }
            if constexpr (FINE) xp[r * S] = acc * tab_dinv[B[p0 + r * S] & 31];
            else xp[r * S] = acc * DI[p0 + r * S];
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
                    off = cxf * (xw + xe) + cyf * (xs + xn);
                    diag = tab_diag[bits & 31];
                } else {
                    off = CX[p] * xw + CX[q0 + r * S] * xe + CY[p] * xs + CY[q0 - o + TK + r * S] * xn;
                    T di = DI[p];
                    diag = di != T(0) ? T(1) / di : T(0);
                }
                T res = (bits & NB_SELF) ? (Fv[p] - (diag * X[p] - off)) * tab_invw[bits >> 5] : T(0);
                Fv[p] = res;    // each thread only overwrites what it alone reads
            }
        }
    }

    __device__ __forceinline__ T &Fat(int par0, int a, int b) {
        int c = (par0 + a + b) & 1;
        return Fv[at(c, a, b >> 1)];
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
                    Fv[p] = self ? (T)(fscale * (double)fv[r][h] - fshift) : T(0);
                    X[p] = self ? (T)xv[r][h] : T(0);
                } else {
                    Fv[p] = (T)fv[r][h];
                    X[p] = (T)xv[r][h];
                    CX[p] = cxv[r][h]; CY[p] = cyv[r][h]; DI[p] = div[r][h];
                }
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
template <typename T, typename TX, typename TF, typename TC, bool FINE, bool ZERO, int NU, int WJ, class Lev>
__global__ void __launch_bounds__(TILE_THREADS, (FINE && sizeof(T) == 4) ? 3 : 2)
k_mg_down(Lev L, const TX *__restrict__ xin, TX *__restrict__ xout, const TF *__restrict__ fin, double fscale,
          const double *__restrict__ scal, int sumr_slot, double inv_n,
          int nyc, int nxc, int pitchc, TC *__restrict__ bc) {
    constexpr int H = halo_down(NU, ZERO);
    constexpr int TJ = WJ - 2 * H, TI = TW - 2 * H;
    extern __shared__ __align__(16) unsigned char smem[];
    Window<T, FINE, WJ> W;
    W.carve(smem);
    const int tj0 = blockIdx.y * TJ, ti0 = blockIdx.x * TI;
    const int wj0 = tj0 - H, wi0 = ti0 - H;
    const int par0 = (wj0 + wi0) & 1;
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    double fshift = 0.0;
    if constexpr (FINE) {
        W.cxf = (T)L.F.cx; W.cyf = (T)L.F.cy;
        W.fill_tables(&L.F, L.dirichlet());
        if (sumr_slot >= 0) fshift = scal[sumr_slot] * inv_n;
    } else W.fill_tables(nullptr, L.dirichlet());
    W.template load<!ZERO>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
    __syncthreads();
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
    // restriction R = P^T: weights (1,3,3,1) x (1,3,3,1) over the 4 x 4 fine cells
    // around the aggregate.  In the colour-separated layout the four columns of
    // a row are (cA,kA) (cB,kB) (cA,kC) (cB,kD) with cB = 1 - cA, and cA flips
    // from one row to the next: four loads at fixed offsets per row.
    for (int t = threadIdx.x; t < (TJ / 2) * (TI / 2); t += TILE_THREADS) {
        int cj = t / (TI / 2), ci = t - cj * (TI / 2);
        int J = tj0 / 2 + cj, I = ti0 / 2 + ci;            // aggregate of fine rows 2J, 2J+1
        int Jc = J + L.pj_off();                           // its row in the coarse array
        if (Jc >= nyc || I >= nxc) continue;
        const int a0 = 2 * cj + H, b0 = 2 * ci + H;        // = 2J - wj0, 2I - wi0
        const int kA = (b0 - 1) >> 1, kB = b0 >> 1, kC = (b0 + 1) >> 1, kD = (b0 + 2) >> 1;
        int cA = (par0 + a0 - 1 + b0 - 1) & 1;             // colour of (a0-1, b0-1)
        constexpr int PL = WJ * TK;                        // one colour plane
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
        bc[(long)(Jc + 1) * pitchc + I + 1] = (TC)acc;
    }
}

// ---------------------------------------------------------------------------
// UP leg.   x <- x + P xc ;  NU sweeps (B,R) ;  [DOT: out = (sum f x, sum x)]
// ---------------------------------------------------------------------------
template <typename T, typename TX, typename TF, typename TC, bool FINE, bool DOT, int NU, int WJ, class Lev>
__global__ void __launch_bounds__(TILE_THREADS, (FINE && sizeof(T) == 4) ? 3 : 2)
k_mg_up(Lev L, const TX *__restrict__ xin, TX *__restrict__ xout, const TF *__restrict__ fin, double fscale,
        const double *__restrict__ scal, int sumr_slot, double inv_n,
        int nyc, int nxc, int pitchc, int periodic_c, const TC *__restrict__ xc,
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
    const int warp = threadIdx.x >> 5, k = threadIdx.x & 31;
    double fshift = 0.0;
    if constexpr (FINE) {
        W.cxf = (T)L.F.cx; W.cyf = (T)L.F.cy;
        W.fill_tables(&L.F, L.dirichlet());
        if (sumr_slot >= 0) fshift = scal[sumr_slot] * inv_n;
    } else W.fill_tables(nullptr, L.dirichlet());
    // ---- coarse window
    const int cj0 = (wj0 >> 1) - 1, ci0 = (wi0 >> 1) - 1;
    for (int t = threadIdx.x; t < CJ * CI; t += TILE_THREADS) {
        int a = t / CI, b = t - a * CI;
        int J = cj0 + a + L.pj_off(), I = ci0 + b;
        TC v = TC(0);
        if (J >= 0 && J < nyc) {
            if (periodic_c) I = wrap_col(I, nxc);
            if (I >= 0 && I < nxc) v = xc[(long)(J + 1) * pitchc + I + 1];
        }
        XC[t] = v;
    }
    W.template load<true>(L, xin, fin, fscale, fshift, wj0, wi0, par0);
    __syncthreads();
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
            T v = T(9) * (T)XC[J0 * CI + I0] + T(3) * ((T)XC[Jn * CI + I0] + (T)XC[J0 * CI + In]) + (T)XC[Jn * CI + In];
            W.X[p] += v * W.tab_invw[bits >> 5];
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
    double acc[2] = {0.0, 0.0};
    {
        const int c0 = L.col(wi0 + 2 * k), c1 = L.col(wi0 + 2 * k + 1);
#pragma unroll
        for (int r = 0; r < Window<T, FINE, WJ>::ROWS_PER_WARP; r++) {
            int a = H + warp + r * TILE_WARPS;
            if (a >= WJ - H) break;
            int rp = (par0 + a) & 1, j = wj0 + a;
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
                    double fv = fscale * (double)fin[rb + ch] - fshift;
                    acc[0] += fv * (double)xv; acc[1] += (double)xv;
                }
            }
        }
    }
    if (DOT) grid_reduce<OpSum, 2>(acc, part, count, out);
}

}  // namespace f2d
