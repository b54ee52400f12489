// extern "C" surface of libf2d.so (see include/f2d.h).
#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "engine.cuh"

namespace f2d {
const char *last_error();
int op_compflux(f2d_ctx *, double *, const double *, const double *, const int8_t *, long, long, int);
int op_vortexforce(f2d_ctx *, double *, const double *, const double *, const int8_t *, long, long,
                   long, int, int);
int op_innerproduct(f2d_ctx *, double *, const double *, const double *, const int8_t *, long, long, int);
}  // namespace f2d

using namespace f2d;

#define NEED(cond, ...)                \
    do {                               \
        if (!(cond)) {                 \
            set_error(__VA_ARGS__);    \
            return F2D_ERR_ARG;        \
        }                              \
    } while (0)

// Every entry point that takes a context makes the context's device current
// (two Models with different param.device may share a process,
// tools.run_twin_experiments) and keeps C++ exceptions (std::map::at on a field
// the model does not have, bad_alloc) on this side of the C boundary.
#define CTX_ENTER(c)                                        \
    NEED(c, "null ctx");                                    \
    F2D_CUDA(cudaSetDevice((c)->cfg.device))

template <class Fn>
static int guarded(Fn &&fn) {
    try {
        return fn();
    } catch (const std::out_of_range &) {
        set_error("a field or mesh array this call needs does not exist for this model");
        return F2D_ERR_ARG;
    } catch (const std::exception &e) {
        set_error("internal error: %s", e.what());
        return F2D_ERR_STATE;
    }
}

static const char *MESH_NAMES[] = {"msk", "mskx", "msky", "mskv", "slip", "oc.x", "oc.y",
                                   "ov.x", "ov.y", "ok.x", "ok.y"};

extern "C" {

int f2d_version(void) { return 100; }
const char *f2d_last_error(void) { return last_error(); }

int f2d_device_count(int *count) {
    NEED(count, "null count");
    F2D_CUDA(cudaGetDeviceCount(count));
    return F2D_OK;
}

static int alloc_field(f2d_ctx *c, const std::string &name) {
    double *p;
    F2D_CUDA(cudaMalloc(&p, c->n * sizeof(double)));
    F2D_CUDA(cudaMemsetAsync(p, 0, c->n * sizeof(double), c->stream));
    c->fields[name] = p;
    return F2D_OK;
}

int f2d_create(const f2d_config *cfg, f2d_ctx **out) {
    NEED(cfg && out, "null argument");
    NEED(cfg->nx > 0 && cfg->ny > 0, "nx, ny must be positive");
    NEED(cfg->nh == 3, "halowidth must be 3 (the widest stencil reaches 3 cells, weno.py:353-355)");
    NEED(cfg->nx >= 2 * cfg->nh && cfg->ny >= 2 * cfg->nh, "nx, ny must be >= 2*halowidth");
    NEED(cfg->model >= 0 && cfg->model <= F2D_MODEL_VECTORADV, "unknown model %d", cfg->model);
    NEED(cfg->integrator >= 0 && cfg->integrator <= F2D_INT_LFRA, "unknown integrator %d", cfg->integrator);
    NEED(cfg->maxorder == 2 || cfg->maxorder == 4 || cfg->maxorder == 6, "maxorder must be 2, 4 or 6");
    NEED(cfg->vortexforce >= 0 && cfg->vortexforce <= 3, "unknown vortexforce method");
    NEED(cfg->compflux >= 0 && cfg->compflux <= 3, "unknown compflux method");
    NEED(cfg->innerproduct >= 0 && cfg->innerproduct <= 4, "unknown innerproduct method");
    NEED(cfg->yperiodic != 2 || cfg->model == F2D_MODEL_EULER || cfg->model == F2D_MODEL_BOUSSINESQ,
         "a truly periodic y direction (ywrap) is available for the euler and boussinesq models");
    F2D_CUDA(cudaSetDevice(cfg->device));
    f2d_ctx *c = new f2d_ctx();
    c->cfg = *cfg;
    c->guess_order = cfg->reserved[0] > 0 ? cfg->reserved[0] - 1 : 4;   // guess_order + 1 (0 = default)
    c->nh = cfg->nh;
    c->n1 = cfg->nx + 2 * cfg->nh;
    c->n2 = cfg->ny + 2 * cfg->nh;
    c->n = (size_t)c->n1 * c->n2;
    c->dx = cfg->Lx / cfg->nx;       // meshes.py:24-26
    c->dy = cfg->Ly / (cfg->reserved[3] > 0 ? cfg->reserved[3] : cfg->ny);   // slab: global ny
    c->area = c->dx * c->dy;
    c->idx2 = 1 / (c->dx * c->dx);   // operators.py:61-62
    c->idy2 = 1 / (c->dy * c->dy);
    *out = c;
    cudaDeviceProp prop;
    F2D_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    c->nsm = prop.multiProcessorCount;
    F2D_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    F2D_CUDA(cudaEventCreate(&c->ev0));
    F2D_CUDA(cudaEventCreate(&c->ev1));
    for (const char *m : MESH_NAMES) {
        int8_t *p;
        F2D_CUDA(cudaMalloc(&p, c->n));
        F2D_CUDA(cudaMemsetAsync(p, 0, c->n, c->stream));
        c->mesh[m] = p;
    }
    // states.py:7-17
    std::vector<std::string> names;
    switch (cfg->model) {
    case F2D_MODEL_EULER:
        names = {"u.x", "u.y", "U.x", "U.y", "omega", "ke", "p", "div"};
        c->prognostic = {"u.x", "u.y"};
        break;
    case F2D_MODEL_BOUSSINESQ:
        names = {"b", "u.x", "u.y", "U.x", "U.y", "omega", "ke", "p", "div", "flx.x", "flx.y"};
        c->prognostic = {"b", "u.x", "u.y"};
        break;
    case F2D_MODEL_RSW:
        names = {"u.x", "u.y", "h", "U.x", "U.y", "omega", "ke", "p", "flx.x", "flx.y", "pv"};
        c->prognostic = {"u.x", "u.y", "h"};
        break;
    case F2D_MODEL_QGRSW:
        names = {"u.x", "u.y", "h", "U.x", "U.y", "omega", "ke", "p", "flx.x", "flx.y", "pv", "psi"};
        c->prognostic = {"u.x", "u.y", "h"};
        break;
    case F2D_MODEL_EULERPSI:
        names = {"omega", "U.x", "U.y", "psi", "vomega", "flx.x", "flx.y"};
        c->prognostic = {"omega"};
        break;
    case F2D_MODEL_QG:
        names = {"pv", "U.x", "U.y", "h", "flx.x", "flx.y", "work", "psi"};
        c->prognostic = {"pv"};
        break;
    case F2D_MODEL_ADVECTION:
        names = {"q", "U.x", "U.y", "flx.x", "flx.y"};
        c->prognostic = {"q"};
        break;
    case F2D_MODEL_VECTORADV:
        names = {"v.x", "v.y", "U.x", "U.y", "omega", "q"};
        c->prognostic = {"v.x", "v.y"};
        break;
    }
    c->tracer = cfg->reserved[5] != 0;
    if (c->tracer) {                    // states.py:23-34: one more prognostic scalar
        names.push_back("tracer");
        c->prognostic.push_back("tracer");
        if (std::find(names.begin(), names.end(), "flx.x") == names.end()) {
            names.push_back("flx.x");
            names.push_back("flx.y");
        }
    }
    c->nstages = cfg->integrator == F2D_INT_EF ? 1 : 3;
    for (auto &nm : names) F2D_TRY(alloc_field(c, nm));
    for (int k = 0; k < c->nstages; k++)
        for (auto &leaf : c->prognostic) F2D_TRY(alloc_field(c, "ds" + std::to_string(k) + "." + leaf));
    F2D_CUDA(cudaMalloc(&c->hb, c->n * sizeof(double)));
    F2D_CUDA(cudaMemsetAsync(c->hb, 0, c->n * sizeof(double), c->stream));
    if (cfg->model == F2D_MODEL_EULER || cfg->model == F2D_MODEL_BOUSSINESQ || cfg->model == F2D_MODEL_RSW)
        for (int t = 0; t < ((cfg->model == F2D_MODEL_EULER) ? 2 : 3); t++) {       // u* (and the updated scalar) of the fused stage kernels
            F2D_CUDA(cudaMalloc(&c->tmp[t], c->n * sizeof(double)));
            F2D_CUDA(cudaMemsetAsync(c->tmp[t], 0, c->n * sizeof(double), c->stream));
        }
    F2D_CUDA(cudaMalloc(&c->d_scal, 64 * sizeof(double)));
    F2D_CUDA(cudaMemsetAsync(c->d_scal, 0, 64 * sizeof(double), c->stream));
    c->part_capacity = std::max<size_t>(4 * 8192, 4 * (c->n / (40 * 40) + 1024));
    F2D_CUDA(cudaMalloc(&c->d_part, c->part_capacity * sizeof(double)));
    F2D_CUDA(cudaMalloc(&c->d_count, sizeof(unsigned int)));
    F2D_CUDA(cudaMemsetAsync(c->d_count, 0, sizeof(unsigned int), c->stream));
    F2D_CUDA(cudaMallocHost(&c->h_scal, 64 * sizeof(double)));
    F2D_CUDA(cudaMallocHost(&c->h_hist, 512 * sizeof(double)));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}

int f2d_destroy(f2d_ctx *c) {
    if (!c) return F2D_OK;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    for (int w = 0; w < 3; w++) mg_free(c, w);
    dist_free(c);
    for (auto &G : c->guess) for (double *g : G.g) cudaFree(g);
    for (auto &kv : c->mesh) cudaFree(kv.second);
    for (auto &kv : c->field_home) c->fields[kv.first] = kv.second;     // history slots are freed above
    for (auto &kv : c->fields) cudaFree(kv.second);
    for (auto &kv : c->forcing) cudaFree(kv.second.pattern);
    if (c->io_stream) cudaStreamSynchronize(c->io_stream);
    for (auto &kv : c->io_stage) {
        cudaFree(kv.second.d);
        cudaEventDestroy(kv.second.filled);
        cudaEventDestroy(kv.second.drained);
    }
    if (c->io_stream) cudaStreamDestroy(c->io_stream);
    cudaFree(c->hb);
    cudaFree(c->smask);
    cudaFree(c->tmask);
    cudaFree(c->dmask);
    for (double *t : c->tmp) cudaFree(t);
    cudaFree(c->d_scal);
    cudaFree(c->d_part);
    cudaFree(c->d_count);
    cudaFreeHost(c->h_scal);
    cudaFreeHost(c->h_hist);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    cudaStreamDestroy(c->own_stream);
    delete c;
    return F2D_OK;
}

int f2d_set_stream(f2d_ctx *c, void *s) {
    CTX_ENTER(c);
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return F2D_OK;
}

int f2d_sync(f2d_ctx *c) {
    CTX_ENTER(c);
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}

static int set_mask_impl(f2d_ctx *c, const int8_t *h_msk) {
    if (c->dist.on) {   // neighbours map my arrays: unmap everywhere before anything is freed
        p2p_teardown(c);
        F2D_TRY(dist_allreduce(c, c->d_scal + 26, 1, false));
        F2D_CUDA(cudaStreamSynchronize(c->stream));
    }
    F2D_TRY(build_mesh(c, h_msk));
    for (auto &G : c->guess) G.valid = 0;     // a new mask invalidates the solve history
    for (auto &kv : c->field_home) {          // fields living in a history slot go home (contents do not matter:
        c->fields[kv.first] = kv.second;       // a new mask restarts from whatever first guess the caller uploads)
    }
    c->field_home.clear();
    // meshes.py:39-47
    F2D_TRY(mg_build(c, F2D_SOLVER_CENTERS));
    F2D_TRY(mg_build(c, F2D_SOLVER_VERTICES));
    if (c->cfg.model == F2D_MODEL_RSW || c->cfg.model == F2D_MODEL_QGRSW || c->cfg.model == F2D_MODEL_QG)
        F2D_TRY(mg_build(c, F2D_SOLVER_HELMHOLTZ));
    if (c->dist.on) {
        // every array the time loop exchanges, in the same order on every rank
        std::vector<void *> arr;
        std::vector<long long> rows, pad, rb;
        auto add = [&](void *p, long long r, long long pd, long long b) { if (p) { arr.push_back(p); rows.push_back(r); pad.push_back(pd); rb.push_back(b); } };
        const long long fb = (long long)c->n1 * sizeof(double);
        for (auto &kv : c->fields) add(kv.second, c->n2, 0, fb);
        add(c->tmp[0], c->n2, 0, fb);
        add(c->tmp[1], c->n2, 0, fb);
        add(c->tmp[2], c->n2, 0, fb);
        for (int w = 0; w < 3; w++) {
            Multigrid &M = c->mg[w];
            if (!M.built) continue;
            for (double *p : {M.r, M.z, M.q}) add(p, c->n2, 0, fb);
            for (float *p : {M.zf, M.zf2, M.p, M.p2}) add(p, c->n2, 0, (long long)c->n1 * sizeof(float));
            for (size_t l = 1; l < M.lev.size(); l++) {
                Level &L = M.lev[l];
                for (CT *p : {L.x, L.x2, L.b}) add(p, L.ny + 2, 1, (long long)L.pitch * sizeof(CT));
            }
        }
        F2D_TRY(p2p_setup(c, arr, rows, pad, rb));
    }
    return F2D_OK;
}

int f2d_set_mask(f2d_ctx *c, const int8_t *h_msk) {
    CTX_ENTER(c);
    return guarded([&] { return set_mask_impl(c, h_msk); });
}

int f2d_get_mesh_array(f2d_ctx *c, const char *name, int8_t *h_out) {
    NEED(c && name && h_out, "null argument");
    CTX_ENTER(c);
    auto it = c->mesh.find(name);
    NEED(it != c->mesh.end(), "unknown mesh array '%s'", name);
    F2D_CUDA(cudaMemcpyAsync(h_out, it->second, c->n, cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}

int f2d_set_topography(f2d_ctx *c, const double *h_hb) {
    CTX_ENTER(c);
    if (h_hb) F2D_CUDA(cudaMemcpyAsync(c->hb, h_hb, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    else F2D_CUDA(cudaMemsetAsync(c->hb, 0, c->n * sizeof(double), c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}

static int find_field(f2d_ctx *c, const char *field, double **p) {
    NEED(c && field, "null argument");
    F2D_CUDA(cudaSetDevice(c->cfg.device));
    auto it = c->fields.find(field);
    NEED(it != c->fields.end(), "unknown field '%s' for this model", field);
    if (c->U_stale && (!strcmp(field, "U.x") || !strcmp(field, "U.y"))) F2D_TRY(ensure_U(c));   // formed on demand
    *p = it->second;
    return F2D_OK;
}

int f2d_upload(f2d_ctx *c, const char *field, const double *h_src) {
    double *p;
    F2D_TRY(find_field(c, field, &p));
    NEED(h_src, "null source");
    F2D_CUDA(cudaMemcpyAsync(p, h_src, c->n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    return F2D_OK;
}

int f2d_download(f2d_ctx *c, const char *field, double *h_dst) {
    double *p;
    F2D_TRY(find_field(c, field, &p));
    NEED(h_dst, "null destination");
    F2D_CUDA(cudaMemcpyAsync(h_dst, p, c->n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    return F2D_OK;
}

int f2d_download_f32(f2d_ctx *c, const char *field, float *h_dst) {
    double *p;
    F2D_TRY(find_field(c, field, &p));
    NEED(h_dst, "null destination");
    return guarded([&] { return download_f32(c, p, field, h_dst); });
}

int f2d_set_forcing(f2d_ctx *c, const char *leaf, const double *h_pattern, double amplitude) {
    NEED(c && leaf, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return set_forcing(c, leaf, h_pattern, amplitude); });
}

int f2d_io_sync(f2d_ctx *c) {
    CTX_ENTER(c);
    if (c->io_stream) F2D_CUDA(cudaStreamSynchronize(c->io_stream));
    return F2D_OK;
}

int f2d_bulk_sums(f2d_ctx *c, int row0, double *out6) {
    NEED(c && out6, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return bulk_sums(c, row0, out6); });
}

int f2d_field_ptr(f2d_ctx *c, const char *field, double **d_ptr) {
    NEED(d_ptr, "null out pointer");
    return find_field(c, field, d_ptr);
}

int f2d_step(f2d_ctx *c, double dt, int nsteps) {
    CTX_ENTER(c);
    return guarded([&] { return model_step(c, dt, nsteps); });
}
int f2d_step_lfra(f2d_ctx *c, double dt, int first, double gamma) {
    CTX_ENTER(c);
    return guarded([&] { return model_step_lfra(c, dt, first, gamma); });
}
int f2d_rhs(f2d_ctx *c, int k) {
    CTX_ENTER(c);
    return guarded([&] { return model_rhs(c, k); });
}
int f2d_addto(f2d_ctx *c, int ncoef, const double *coefs) {
    NEED(c && coefs, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return model_addto(c, ncoef, coefs); });
}
int f2d_diag(f2d_ctx *c) {
    CTX_ENTER(c);
    return guarded([&] { return model_diag(c); });
}
int f2d_max_abs_U(f2d_ctx *c, double *h_out) {
    NEED(c && h_out, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return max_abs_U(c, h_out); });
}

int f2d_solve(f2d_ctx *c, int which, const double *d_b, double bscale, double *d_x, int *iters,
              double *relres) {
    NEED(c && d_b && d_x, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return mg_solve(c, which, d_b, bscale, d_x, iters, relres); });
}
int f2d_apply_laplacian(f2d_ctx *c, int which, const double *d_x, double *d_y) {
    NEED(c && d_x && d_y, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return mg_apply(c, which, d_x, d_y); });
}
int f2d_solver_stats(f2d_ctx *c, int64_t *nsolves, int64_t *niters, double *max_relres) {
    NEED(c, "null ctx");
    if (nsolves) *nsolves = c->nsolves;
    if (niters) *niters = c->niters;
    if (max_relres) *max_relres = c->max_relres;
    c->nsolves = c->niters = 0;
    c->max_relres = 0;
    return F2D_OK;
}

int f2d_solver_info(f2d_ctx *c, int which, int *ncomponents, int *nlevels, double *rhs_incompat) {
    NEED(c, "null ctx");
    NEED(which >= 0 && which <= 2, "solver id %d", which);
    Multigrid &M = c->mg[which];
    NEED(M.built, "solver %d not built", which);
    if (ncomponents) *ncomponents = M.ncomp;
    if (nlevels) *nlevels = (int)M.lev.size() + (c->dist.on ? (int)M.glev.size() - 1 : 0);
    if (rhs_incompat) { *rhs_incompat = M.rhs_incompat; M.rhs_incompat = 0.0; }    // since the last call
    return F2D_OK;
}

int f2d_compflux(f2d_ctx *c, double *flx, const double *U, const double *q, const int8_t *o,
                 int64_t n, int64_t s, int method) {
    NEED(c && flx && U && q && o, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return op_compflux(c, flx, U, q, o, (long)n, (long)s, method); });
}
int f2d_vortexforce(f2d_ctx *c, double *du, const double *V, const double *q, const int8_t *o,
                    int64_t n, int64_t s, int64_t s2, int sign, int method) {
    NEED(c && du && V && q && o, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return op_vortexforce(c, du, V, q, o, (long)n, (long)s, (long)s2, sign, method); });
}
int f2d_innerproduct(f2d_ctx *c, double *ke, const double *U, const double *q, const int8_t *o,
                     int64_t n, int64_t s, int method) {
    NEED(c && ke && U && q && o, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return op_innerproduct(c, ke, U, q, o, (long)n, (long)s, method); });
}
int f2d_fill(f2d_ctx *c, double *d_a) {
    NEED(c && d_a, "null argument");
    CTX_ENTER(c);
    return guarded([&] { return op_fill(c, d_a); });
}

int f2d_malloc(f2d_ctx *c, size_t bytes, void **d_ptr) {
    NEED(c && d_ptr, "null argument");
    F2D_CUDA(cudaSetDevice(c->cfg.device));
    F2D_CUDA(cudaMalloc(d_ptr, bytes));
    return F2D_OK;
}
int f2d_free(f2d_ctx *c, void *d_ptr) {
    CTX_ENTER(c);
    F2D_CUDA(cudaFree(d_ptr));
    return F2D_OK;
}
int f2d_memcpy_h2d(f2d_ctx *c, void *d_dst, const void *h_src, size_t bytes) {
    NEED(c && d_dst && h_src, "null argument");
    CTX_ENTER(c);
    F2D_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, c->stream));
    return F2D_OK;
}
int f2d_memcpy_d2h(f2d_ctx *c, void *h_dst, const void *d_src, size_t bytes) {
    NEED(c && h_dst && d_src, "null argument");
    CTX_ENTER(c);
    F2D_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    return F2D_OK;
}
int f2d_host_alloc(size_t bytes, void **h_ptr) {
    NEED(h_ptr, "null argument");
    F2D_CUDA(cudaMallocHost(h_ptr, bytes));
    return F2D_OK;
}
int f2d_host_free(void *h_ptr) {
    F2D_CUDA(cudaFreeHost(h_ptr));
    return F2D_OK;
}
int f2d_timer_start(f2d_ctx *c) {
    CTX_ENTER(c);
    F2D_CUDA(cudaEventRecord(c->ev0, c->stream));
    return F2D_OK;
}
int f2d_timer_stop(f2d_ctx *c, float *ms) {
    NEED(c && ms, "null argument");
    CTX_ENTER(c);
    F2D_CUDA(cudaEventRecord(c->ev1, c->stream));
    F2D_CUDA(cudaEventSynchronize(c->ev1));
    F2D_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return F2D_OK;
}
int f2d_bench_kernel(f2d_ctx *c, const char *name, int reps, float *ms, double *alg_bytes) {
    NEED(c && name && ms && alg_bytes && reps > 0, "bad argument");
    CTX_ENTER(c);
    return guarded([&] {
        if (!strncmp(name, "mg.", 3) || !strncmp(name, "cg.", 3)) return bench_mg_kernel(c, name, reps, ms, alg_bytes);
        return bench_step_kernel(c, name, reps, ms, alg_bytes);
    });
}
int f2d_dist_unique_id(char *id128) {
    NEED(id128, "null argument");
    return dist_unique_id(id128);
}
int f2d_dist_init(f2d_ctx *c, int rank, int world, const char *id128) {
    NEED(c && id128, "null argument");
    NEED(!c->mesh_ready, "f2d_dist_init must precede f2d_set_mask");
    CTX_ENTER(c);
    return guarded([&] { return dist_init(c, rank, world, id128); });
}
int f2d_dist_exchange(f2d_ctx *c, const char *field) {
    double *p;
    F2D_TRY(find_field(c, field, &p));
    return dist_exchange1(c, p, (size_t)c->n1 * sizeof(double), c->n2, 0);
}
int f2d_exchange_count(f2d_ctx *c, int64_t *count) {
    NEED(c && count, "null argument");
    *count = c->exchanges;
    return F2D_OK;
}
int f2d_launch_count(f2d_ctx *c, int64_t *count) {
    NEED(c && count, "null argument");
    *count = c->launches;
    return F2D_OK;
}

}  // extern "C"
