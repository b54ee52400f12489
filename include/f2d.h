/*
 * f2d.h -- C ABI of libf2d.so, the B200 (sm_100a) implementation of the
 * Fluids2d time-step hot path.
 *
 * The reference (pvthinker/Fluids2d) has no FFI: its seams are Python
 * callables on in-place numpy arrays.  Each entry point
 * below names the reference interface it stands in for
 * (paths relative to the reference tree).  A maintainer binds them with
 * ctypes -- see INTEGRATION.md and fluids2d_b200/_cabi.py.
 *
 * Conventions
 *   - every function returns 0 on success or a negative f2d_status; the text
 *     of the last error on the calling thread is f2d_last_error()
 *   - no C++ types, no torch types, no exceptions cross this boundary
 *   - arrays are C-contiguous, shape (n2, n1) = (ny+2*nh, nx+2*nh), float64
 *     unless stated; masks / stencil orders are int8   (meshes.py:122-132)
 *   - pointers named d_* are DEVICE pointers, h_* are HOST pointers
 *   - one host thread per context; work is enqueued on the context's stream
 */
#ifndef F2D_H
#define F2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct f2d_ctx f2d_ctx;

typedef enum {
    F2D_OK = 0,
    F2D_ERR_CUDA = -1,      /* a CUDA runtime call failed                     */
    F2D_ERR_ARG = -2,       /* bad argument (unknown name, bad size, NULL)    */
    F2D_ERR_STATE = -3,     /* call order (e.g. step before set_mask)         */
    F2D_ERR_NOTCONV = -4,   /* elliptic solve hit maxit before reaching rtol  */
    F2D_ERR_UNSUPPORTED = -5
} f2d_status;

/* param.py:1-7 string enums, in the order the reference lists them */
enum { F2D_MODEL_EULER = 0, F2D_MODEL_BOUSSINESQ = 1, F2D_MODEL_RSW = 2, F2D_MODEL_QGRSW = 3,
       F2D_MODEL_EULERPSI = 4, F2D_MODEL_QG = 5, F2D_MODEL_ADVECTION = 6, F2D_MODEL_VECTORADV = 7 };
enum { F2D_METHOD_WENO = 0, F2D_METHOD_UPWIND = 1, F2D_METHOD_CENTERED = 2, F2D_METHOD_CWENO = 3,
       F2D_METHOD_CLASSIC = 4 /* innerproduct only, operators.py:86-89 */ };
enum { F2D_INT_RK3 = 0, F2D_INT_EF = 1, F2D_INT_ENRK3 = 2, F2D_INT_LFRA = 3 };
/* noslip.py:15-34: bit flags; F2D_NOSLIP_ALL == param.noslip is True */
enum { F2D_NOSLIP_NONE = 0, F2D_NOSLIP_LEFT = 1, F2D_NOSLIP_RIGHT = 2, F2D_NOSLIP_BOTTOM = 4,
       F2D_NOSLIP_TOP = 8, F2D_NOSLIP_ALL = 16 };
/* elliptic.py:71-75: "c" centres/Neumann, "v" vertices/Dirichlet, plus the
 * vertex Helmholtz operator of meshes.py:42-47 */
enum { F2D_SOLVER_CENTERS = 0, F2D_SOLVER_VERTICES = 1, F2D_SOLVER_HELMHOLTZ = 2 };

/* Mirrors the attributes of param.py:13-59 that the hot path reads. */
typedef struct {
    int32_t model;          /* F2D_MODEL_*                                   */
    int32_t nx, ny, nh;     /* param.nx, param.ny, param.halowidth           */
    double Lx, Ly;
    int32_t xperiodic, yperiodic;   /* yperiodic: 1 = the reference's behaviour (the mask of the halo rows is 1,
                                     * nothing wraps: meshes.py:74, :135-143, elliptic.py:142); 2 = a true
                                     * periodic direction (halo rows are images, the Laplacian wraps): NEW */
    int32_t noslip;         /* F2D_NOSLIP_* flags                            */
    double f0, g, H;
    int32_t integrator;     /* F2D_INT_*                                     */
    int32_t compflux, vortexforce, innerproduct;  /* F2D_METHOD_*            */
    int32_t maxorder;       /* 2, 4 or 6                                     */
    int32_t device;         /* CUDA device ordinal                           */
    /* elliptic solver controls (new: the reference uses a direct solve)     */
    double solver_rtol;     /* ||b - A x|| <= rtol * ||b||; 0 -> 1e-12       */
    int32_t solver_maxit;   /* 0 -> 100                                      */
    int32_t solver_kind;    /* 0 MG-preconditioned CG, 1 plain V-cycles      */
    int32_t nu1, nu2;       /* red-black sweeps before / after; 0 -> 2       */
    int32_t reserved[8];    /* [1], [2], [3]: slab decomposition, see f2d_dist_init;
                             * [4]: non-zero when param.beta != 0 (qg: unsupported);
                             * [5]: non-zero adds the passive scalar of param.tracer
                             *      (states.py:23-34, equations.py:217-226) as the device
                             *      field "tracer" (tendencies "ds<k>.tracer");
                             * [0]: 1 + order of the first-guess extrapolation across
                             * time steps for the solves inside f2d_step (0 = default
                             * = cubic; 1 = off, 2 = previous step, 3 = linear, 4 = quadratic,
                             * 5 = cubic, up to 7) */
} f2d_config;

int f2d_version(void);
const char *f2d_last_error(void);
int f2d_device_count(int *count);

/* ---- context: stands in for Model.__init__ (model.py:14-24): Mesh + State +
 *      integrator scratch, all resident in HBM ------------------------------ */
int f2d_create(const f2d_config *cfg, f2d_ctx **out);
int f2d_destroy(f2d_ctx *ctx);
/* Borrow a caller's cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream);
 * NULL restores the context's own stream. */
int f2d_set_stream(f2d_ctx *ctx, void *cuda_stream);
int f2d_sync(f2d_ctx *ctx);

/* Mesh.finalize (meshes.py:34-52): takes the cell mask `msk` (host, int8,
 * (n2,n1); NULL = the default mask of meshes.py:70-75), derives mskx/msky/mskv
 * (:77-86), slipcoef (noslip.py:4-36), the six stencil-order arrays (:88-104,
 * set_order :146-186) on the device and rebuilds the multigrid hierarchies
 * that replace Poisson2D's splu factors (elliptic.py:71-78). */
int f2d_set_mask(f2d_ctx *ctx, const int8_t *h_msk);
/* name in {msk,mskx,msky,mskv,slip,oc.x,oc.y,ov.x,ov.y,ok.x,ok.y}; int8 out */
int f2d_get_mesh_array(f2d_ctx *ctx, const char *name, int8_t *h_out);
/* mesh.hb (rsw_with_topo.py:96-99); NULL resets to 0 */
int f2d_set_topography(f2d_ctx *ctx, const double *h_hb);

/* ---- State (states.py:37-79).  Field names: "u.x","u.y","U.x","U.y","omega",
 *      "ke","p","div","flx.x","flx.y","b","h","pv","psi","vomega","work","q","v.x","v.y";
 *      scratch tendencies
 *      "ds0.u.x", "ds1.h", ... (integrators.py:66-67) ----------------------- */
int f2d_upload(f2d_ctx *ctx, const char *field, const double *h_src);
int f2d_download(f2d_ctx *ctx, const char *field, double *h_dst);
/* device address of a field NOW: the fused stage kernels write u* (h*, b*) and the solves their
 * solution into other arrays and swap the pointers, so the address is valid until the next
 * f2d_step / f2d_diag only */
int f2d_field_ptr(f2d_ctx *ctx, const char *field, double **d_ptr);

/* Model.add_forcing (model.py:121-123; addforcingterm, equations.py:229-238) for
 * forcings of the form  ds.<leaf> += amplitude * pattern  (forced_convection.py:
 * 9-24): the pattern (n2*n1 doubles, host) is kept on the device and added after
 * the model's tendency in every stage of f2d_step / f2d_rhs, so a forced run
 * needs no host round trip.  h_pattern == NULL only changes the amplitude
 * (time-dependent coefficient, set once per step); amplitude 0 disables it.
 * leaf must be a prognostic field ("u.x", "b", "h", "tracer", ...). */
int f2d_set_forcing(f2d_ctx *ctx, const char *leaf, const double *h_pattern, double amplitude);

/* ---- observation points (model.py:48-51) --------------------------------- */
/* io.write (io.py:12-32; NetCDF variables are float32, io.py:65): convert the
 * field to float32 on the device and copy it to h_dst (pinned) on a separate
 * copy stream, so the step loop is not stalled; f2d_io_sync waits for all
 * outstanding history copies before the host reads h_dst. */
int f2d_download_f32(f2d_ctx *ctx, const char *field, float *h_dst);
int f2d_io_sync(f2d_ctx *ctx);
/* diagnostics.Bulk.__call__ (diagnostics.py:39-62): the whole-array sums behind
 * its ke / ens / vort / angular averages, one pass on the device (all-reduced
 * over slabs).  out6 = [sum ke, sum omega^2, sum omega, sum U.y*xv, sum U.x*yu,
 * sum msk];  row0 = first global row of this context's array (0 on one GPU). */
int f2d_bulk_sums(f2d_ctx *ctx, int row0, double *out6);

/* ---- time stepping ------------------------------------------------------ */
/* RKIntegrator.step (integrators.py:76-79) with rk3/ef/enrk3 (:82-124):
 * nsteps fused steps at fixed dt, state stays on the device. */
int f2d_step(f2d_ctx *ctx, double dt, int nsteps);
/* LFRAintegrator.step (integrators.py:20-53): one leap-frog step with the
 * Robert-Asselin filter; scratch sets ds0 = sb, ds1 = sa, ds2 = ds as in the
 * reference's `scratch` list; first = (time.ite == 0). */
int f2d_step_lfra(f2d_ctx *ctx, double dt, int first, double gamma);
/* The same step in the reference's granularity, so that host callbacks
 * (model.add_forcing, equations.py:229-238) can run between the pieces:
 * ds_k = rhs(state) (equations.py:11-15 etc.) */
int f2d_rhs(f2d_ctx *ctx, int k);
/* addto(s, c0, ds0, ..., c_{n-1}, ds_{n-1}) on the prognostic fields
 * (integrators.py:154-196) */
int f2d_addto(f2d_ctx *ctx, int ncoef, const double *coefs);
/* diag(state) (equations.py:17-22 etc.) */
int f2d_diag(f2d_ctx *ctx);
/* Model.set_dt (model.py:71-87): max|U.x| + max|U.y| over the whole arrays */
int f2d_max_abs_U(f2d_ctx *ctx, double *h_out);

/* ---- elliptic: Poisson2D.solve(b, x) (elliptic.py:80-87).  Solves
 *      A x = bscale * b on the fluid points of `which`, starting from the
 *      contents of x, leaves masked entries untouched, then mesh.fill(x).
 *      d_b / d_x are device arrays (n2,n1).  iters/relres may be NULL. ------ */
int f2d_solve(f2d_ctx *ctx, int which, const double *d_b, double bscale, double *d_x,
              int *iters, double *relres);
/* y = A x for the same operators (pins the matrix of elliptic.py:114-195) */
int f2d_apply_laplacian(f2d_ctx *ctx, int which, const double *d_x, double *d_y);
/* statistics of the solves issued by f2d_step / f2d_diag since the last call */
int f2d_solver_stats(f2d_ctx *ctx, int64_t *nsolves, int64_t *niters, double *max_relres);
/* structure of solver `which`: connected components of its unknowns (the
 * all-Neumann operator of elliptic.py:186-190 has one null-space constant per
 * component; each gets its own projection), multigrid levels, and how far the
 * right-hand sides seen since the last call were from the operator's range:
 * max over solves and components of |sum_c b| / sqrt(N_c b.b) (0 = compatible; a
 * right-hand side that is pure rounding noise, e.g. the divergence of an already
 * projected velocity, reads O(1) and is harmless: it is projected). */
int f2d_solver_info(f2d_ctx *ctx, int which, int *ncomponents, int *nlevels, double *rhs_incompat);

/* ---- the three kernels of weno.py:412-436, one launch each, flat-index
 *      semantics of the reference (s, s2 are flat strides), device arrays of
 *      n elements ---------------------------------------------------------- */
int f2d_compflux(f2d_ctx *ctx, double *d_flx, const double *d_U, const double *d_q,
                 const int8_t *d_o, int64_t n, int64_t s, int method);
int f2d_vortexforce(f2d_ctx *ctx, double *d_du, const double *d_V, const double *d_q,
                    const int8_t *d_o, int64_t n, int64_t s, int64_t s2, int sign, int method);
int f2d_innerproduct(f2d_ctx *ctx, double *d_ke, const double *d_U, const double *d_q,
                     const int8_t *d_o, int64_t n, int64_t s, int method);
/* Mesh.fill (meshes.py:135-143) */
int f2d_fill(f2d_ctx *ctx, double *d_a);

/* ---- plumbing: device / pinned-host memory for callers without torch ------ */
int f2d_malloc(f2d_ctx *ctx, size_t bytes, void **d_ptr);
int f2d_free(f2d_ctx *ctx, void *d_ptr);
int f2d_memcpy_h2d(f2d_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
int f2d_memcpy_d2h(f2d_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
int f2d_host_alloc(size_t bytes, void **h_ptr);   /* page-locked */
int f2d_host_free(void *h_ptr);
/* CUDA-event stopwatch on the context's stream (bench.py) */
int f2d_timer_start(f2d_ctx *ctx);
int f2d_timer_stop(f2d_ctx *ctx, float *ms);
/* bench.py: time ONE kernel alone (`reps` launches between two events on the
 * context's stream; average ms per launch) and report its algorithmic bytes
 * per launch.  Names: advection, flux_div (the one-kernel scalar transport of
 * rsw / boussinesq where TMA applies), rk_update, divergence, project_diag,
 * diag, qg_pv, qg_back (as the model has them), mg.down0, mg.up0, mg.down1,
 * mg.up1, mg.tail, cg.dir_apply, cg.update.
 * Clobbers scratch arrays and diagnostics: call it last. */
int f2d_bench_kernel(f2d_ctx *ctx, const char *name, int reps, float *ms, double *alg_bytes);
/* ---- y-slab decomposition over several GPUs, one process and one context per
 *      GPU (new: the reference is single-process).  The context is created
 *      for the LOCAL slab: cfg.ny = rows of the local arrays - 2*nh, where the
 *      local arrays hold the owned rows plus cfg.reserved[1] (south) and
 *      cfg.reserved[2] (north) ghost rows -- 8 at an interface with a
 *      neighbour, 0 at a physical wall (whose nh halo rows are part of the
 *      arrays as usual); cfg.reserved[3] = global ny.  f2d_dist_init joins the
 *      NCCL communicator (id from f2d_dist_unique_id on rank 0, passed around
 *      by the host program) and must precede f2d_set_mask.  Afterwards
 *      f2d_step / f2d_solve / f2d_max_abs_U exchange ghost rows and reduce
 *      scalars themselves; state arrays are uploaded / downloaded per slab.
 *      Models: euler, boussinesq, rsw, qgrsw (F2D_ERR_UNSUPPORTED otherwise). */
int f2d_dist_unique_id(char *id128);
int f2d_dist_init(f2d_ctx *ctx, int rank, int world, const char *id128);
/* refresh the ghost rows of one field from the owners (after an upload) */
int f2d_dist_exchange(f2d_ctx *ctx, const char *field);
int f2d_exchange_count(f2d_ctx *ctx, int64_t *count);
/* number of kernels this context has launched so far */
int f2d_launch_count(f2d_ctx *ctx, int64_t *count);

#ifdef __cplusplus
}
#endif
#endif /* F2D_H */
