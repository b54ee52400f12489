// Internal structures shared by the translation units of libf2d.so.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/f2d.h"

namespace f2d {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define F2D_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t _e = (call);                                               \
        if (_e != cudaSuccess) return f2d::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define F2D_TRY(call)                  \
    do {                               \
        int _s = (call);               \
        if (_s != F2D_OK) return _s;   \
    } while (0)

// mask bits of the fine multigrid level (one byte per grid point)
enum : uint8_t {
    NB_SELF = 1,   // the point is an unknown
    NB_W = 2, NB_E = 4, NB_S = 8, NB_N = 16,   // coupling to that neighbour is open
    NB_PJ = 32,    // coarse parent (Jn, I0) is fluid      (prolongation weights)
    NB_PI = 64,    //               (J0, In)
    NB_PJI = 128,  //               (Jn, In)
    // coarse levels (their codes carry no face bits): the point and its four faces have
    // the level's regular coefficients -- no wall, no mask nearby (open tiles, mg_tiles.cuh)
    NB_REG = 2
};

// element type of the coarse multigrid levels: the coarse-grid correction only
// has to be accurate to a fraction of the error it removes, so levels l >= 1
// are stored and relaxed in fp32 (half the traffic and shared memory); the fine
// level, the residuals and every CG scalar stay fp64.
typedef float CT;

// Fine level: a window of the reference-layout (n2, n1) arrays.
struct FineView {
    int ny, nx;        // logical size of the window
    int oj, oi;        // array coordinates of logical (0,0); may be negative
    int n2, n1;        // array shape
    int periodic;      // x wraps inside the window
    int periodic_y;    // y wraps inside the window (param.ywrap: a true periodic direction, SURVEY note Y)
    int dirichlet;     // vertices: fixed diagonal; centres: Neumann
    double cx, cy;     // dy/dx, dx/dy   (elliptic.py:138-139)
    double shift;      // maindiag       (elliptic.py:186-190)
    const uint8_t *nb;
    int jo0, jo1;      // rows this rank owns (reductions); [0, ny) on one GPU
    int pj_off;        // parent row of fine row j is (j >> 1) + pj_off (ghost rows south)
};

// Coarse level l >= 1: halo-padded arrays (ny+2) x pitch, element (J,I) at
// (J+1)*pitch + I+1.
struct CoarseView {
    int ny, nx, pitch;
    int periodic, periodic_y;
    int dirichlet;
    int pj_off;
    const CT *cx;         // coupling across the west face of (J,I)
    const CT *cy;         // coupling across the south face
    const CT *dinv;       // 1/diagonal, 0 where not an unknown
    const uint8_t *code;  // NB_SELF | NB_P* bits
};

struct Level {
    int ny = 0, nx = 0, pitch = 0;
    size_t n = 0;                       // (ny+2)*pitch
    CT *x = nullptr, *x2 = nullptr, *b = nullptr, *r = nullptr;
    CT *cx = nullptr, *cy = nullptr, *dinv = nullptr;
    double *mass = nullptr, *wall = nullptr;   // set-up only
    uint8_t *code = nullptr;
    CT cx0 = 0, cy0 = 0, dinv0 = 0;            // coefficients of a regular point (NB_REG)
};

// y-slab decomposition: one context per GPU/process.  Every array of a rank
// carries G ghost rows at each interface with a neighbour (at every multigrid
// level); a kernel's results are exact on the owned rows as long as its domain
// of dependence is <= G rows, and the ghosts are refreshed from the owners
// after each kernel (dist.cu).
struct Dist {
    bool on = false;
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;
    int G = 8;                       // ghost rows at an interface
    bool south = false, north = false;
    // peer-to-peer exchange (dist.cu): neighbours' arrays mapped through CUDA IPC
    bool p2p = false;
    unsigned long long *flags = nullptr;
    struct Geo { long long rows_total, pad, row_bytes; };
    std::vector<void *> peer_south, peer_north;      // index 0 = the flag block
    std::vector<Geo> rec_south, rec_north;
    std::map<const void *, int> index;               // my array -> slot
};

struct Multigrid {
    bool built = false;
    bool singular = false;            // the operator has a null space of constants (all-Neumann, or Dirichlet
                                      // vertices of a doubly periodic domain without any wall)
    int which = 0;
    FineView fine{};
    uint8_t *nb = nullptr;            // fine bits, (n2,n1)
    uint8_t *cg_open = nullptr;       // per tile of k_cg_dir_apply: all unknowns, no wrap (mg.cu)
    std::vector<Level> lev;           // lev[0] unused except sizes; lev[l>=1] coarse
    std::vector<Level> glev;          // slab mode: global (replicated) copies of the tail levels
    int gs = 0, gn = 0;               // ghost rows south / north of the owned rows, every level
    int jo0 = 0, jo1 = 0;             // owned logical rows of the fine level [jo0, jo1)
    int tail_y0 = 0;                  // slab mode: first owned row of level `tail` in the global tail grid
    double n_global = 0;              // unknowns over all ranks
    // CUDA graphs of one PCG iteration, keyed by (solution array, first / odd / even iteration):
    // inside f2d_step the solution array rotates through the first-guess history slots
    struct IterGraph { const double *x; int cls; cudaGraphExec_t exec; int64_t launches, exchanges; };
    std::vector<IterGraph> graphs;
    bool warm = false;
    int expect = 0;        // iterations the previous solve needed: that many run without a host check
    double *r = nullptr, *z = nullptr, *q = nullptr;   // CG residual; work vectors of the un-fused cross-check path (n2,n1)
    float *p = nullptr, *p2 = nullptr;                 // CG search direction, ping-pong, fp32 (mg.cu: k_cg_dir_apply)
    float *zf = nullptr, *zf2 = nullptr;   // preconditioned residual z = M r: fp32 (mg_tiles.cuh)
    int tail = 1;                     // first level handled by the single-CTA tail kernel
    int64_t nunknown = 0;
    // Neumann null space: one constant per CONNECTED fluid component (elliptic.py:186-190 gives
    // every component its own singular block).  One component: the mean is removed lazily
    // inside the kernels.  Several: `comp` labels the unknowns of the MAXCOMP largest
    // components (0xFF elsewhere) and r, z are projected explicitly per component.
    static constexpr int MAXCOMP = 8;
    int ncomp = 1;                    // connected components of the unknowns
    uint8_t *comp = nullptr;          // (n2,n1), only when ncomp > 1
    double inv_nc[MAXCOMP] = {0, 0, 0, 0, 0, 0, 0, 0};
    double rhs_incompat = 0;          // max_c |sum_c b| / sqrt(N_c b.b) over the solves so far
};

}  // namespace f2d

namespace f2d {
// Last solutions of one RK stage's elliptic solve.  The solver works IN the slot that
// becomes the newest entry (the oldest one, which the extrapolation no longer reads), so a
// solution is never copied into the history: depth = extrapolation order + 1.
struct GuessHistory {
    static constexpr int MAXD = 7;
    double *g[MAXD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double t[MAXD] = {0, 0, 0, 0, 0, 0, 0};     // model time and step length each was computed at
    double dt[MAXD] = {1, 1, 1, 1, 1, 1, 1};
    int valid = 0;
};
// first guess x0 = sum_k w[k] g[k] (k < n), formed inside the initial-residual kernel
struct GuessSpec {
    const double *g[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    double w[6] = {0, 0, 0, 0, 0, 0};
    int n = 0;
};
}  // namespace f2d

struct f2d_ctx {
    f2d_config cfg{};
    int n1 = 0, n2 = 0, nh = 0;
    size_t n = 0;
    double dx = 0, dy = 0, area = 0, idx2 = 0, idy2 = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int64_t launches = 0, exchanges = 0;
    int nsm = 148;

    // mesh (int8, (n2,n1))
    std::map<std::string, int8_t *> mesh;
    bool mesh_ready = false;
    // state + scratch (float64, (n2,n1))
    std::map<std::string, double *> fields;
    // a field whose storage currently is a first-guess history slot (step.cu: guess_begin):
    // its own allocation, to be handed back / freed
    std::map<std::string, double *> field_home;
    bool U_stale = false;                    // U = sharp(u) is formed on demand (step.cu: ensure_U)
    // mask bits packed for the fused stencil kernels (ops.cu: k_pack_masks), rebuilt by f2d_set_mask:
    //   smask (n2, n1):        ov.x/2 | ov.y/2 << 2 | mskx << 4 | msky << 5           (stage kernel: 1 B instead of 4)
    //   dmask (n2, dpitch):    mskx | msky << 1 | msk << 2 | slip << 3 | ok.x/2 << 4 | ok.y/2 << 6
    //                          (diagnostic kernel: 1 B instead of 6; dpitch = n1 rounded up to 16 so that TMA can fetch boxes)
    //   tmask (n2, n1):        oc.x/2 | oc.y/2 << 2 | msk << 4                        (scalar transport kernel: 1 B instead of 3)
    uint8_t *smask = nullptr, *dmask = nullptr, *tmask = nullptr;
    int dpitch = 0;
    // TMA tensor maps (128-byte CUtensorMap blobs) of the (n2,n1) arrays, keyed by (base pointer, box)
    struct TmaBlob { alignas(64) unsigned char b[128]; bool ok; };
    std::map<std::pair<const void *, long>, TmaBlob> tma_cache;
    std::vector<std::string> prognostic;     // leaf names, e.g. "u.x","u.y"
    int nstages = 3;
    double *hb = nullptr;                    // topography (zeros by default)
    double *tmp[4] = {nullptr, nullptr, nullptr, nullptr};   // work arrays (n2,n1)

    f2d::Multigrid mg[3];
    f2d::Dist dist;
    f2d::GuessHistory guess[3];
    // history output (io.py:12-32): float32 staging + a copy stream of its own
    struct IoStage { float *d = nullptr; cudaEvent_t filled = nullptr, drained = nullptr; };
    std::map<std::string, IoStage> io_stage;
    cudaStream_t io_stream = nullptr;
    // model.add_forcing with a device pattern: ds.<leaf> += amplitude * pattern
    struct Forcing { double *pattern = nullptr; double amplitude = 0.0; };
    std::map<std::string, Forcing> forcing;
    double sim_t = 0.0, sim_dt = 1.0;   // clock of f2d_step: start of the current step, its length
    bool tracer = false;            // param.tracer: extra advected scalar "tracer" (equations.py:217-226)
    int guess_order = 4;            // 0 off, 1 previous step, 2 linear, 3 quadratic, 4 cubic ... 6
    int stage_hint = -1;
    // reductions
    double *d_scal = nullptr;       // device scalars
    double *d_part = nullptr;       // per-block partial sums
    size_t part_capacity = 0;       // doubles
    unsigned int *d_count = nullptr;
    double *h_scal = nullptr;       // pinned mirror
    double *h_hist = nullptr;       // pinned: (rr, sum r) of the PCG iterations that ran unchecked
    // solver statistics
    int64_t nsolves = 0, niters = 0;
    double max_relres = 0;

    int8_t *m(const char *k) { return mesh.at(k); }
    double *f(const std::string &k) { return fields.at(k); }
    bool has(const std::string &k) const { return fields.count(k) != 0; }
};

namespace f2d {
// ops.cu
int build_mesh(f2d_ctx *c, const int8_t *h_msk);
int op_fill(f2d_ctx *c, double *a);
// step.cu
int model_rhs(f2d_ctx *c, int k);
int model_addto(f2d_ctx *c, int ncoef, const double *coefs);
int model_diag(f2d_ctx *c);
int model_step(f2d_ctx *c, double dt, int nsteps);
int model_step_lfra(f2d_ctx *c, double dt, int first, double gamma);
int max_abs_U(f2d_ctx *c, double *out);
int bulk_sums(f2d_ctx *c, int row0, double *out);
int set_forcing(f2d_ctx *c, const std::string &leaf, const double *h_pattern, double amplitude);
int download_f32(f2d_ctx *c, const double *src, const std::string &key, float *h_dst);
// mg.cu
int mg_build(f2d_ctx *c, int which);
void mg_free(f2d_ctx *c, int which);
int mg_solve(f2d_ctx *c, int which, const double *b, double bscale, double *x, int *iters,
             double *relres, const GuessSpec *guess = nullptr);
int ensure_U(f2d_ctx *c);
int mg_apply(f2d_ctx *c, int which, const double *x, double *y);
// dist.cu
int dist_exchange(f2d_ctx *c, int narr, void *const *base, size_t row_bytes, long nrows, long row0);
int dist_exchange1(f2d_ctx *c, void *base, size_t row_bytes, long nrows, long row0);
int dist_exchange_parts(f2d_ctx *c, int n, char *const *base, const size_t *row_bytes, const long *nrows,
                        const long *row0, size_t stride);
int dist_allreduce(f2d_ctx *c, double *d_vals, int n, bool max_op);
int dist_allgather_rows(f2d_ctx *c, const void *src_rows, void *dst, size_t bytes_per_rank);
int dist_init(f2d_ctx *c, int rank, int world, const char *unique_id);
int dist_unique_id(char *out);
void dist_free(f2d_ctx *c);
int p2p_setup(f2d_ctx *c, const std::vector<void *> &arrays, const std::vector<long long> &rows_total,
              const std::vector<long long> &pad, const std::vector<long long> &row_bytes);
void p2p_teardown(f2d_ctx *c);
int p2p_check(f2d_ctx *c);
int bench_mg_kernel(f2d_ctx *c, const char *name, int reps, float *ms, double *bytes);
int bench_step_kernel(f2d_ctx *c, const char *name, int reps, float *ms, double *bytes);
}  // namespace f2d
