// y-slab decomposition plumbing: ghost-row exchange between neighbouring ranks
// and the scalar reductions of CG / CFL, over NCCL (NVLink / NVSwitch on an
// 8 x B200 box).  Rows are contiguous in every array of the path (row-major,
// yshift = n1), so a ghost exchange is one contiguous message per direction and
// array: no packing kernels.
//
// NCCL is bound at run time (dlopen), not at link time: a process that also
// imports torch must end up with ONE libnccl.so.2, and torch's bundled copy is
// newer than the system one.  Whoever loads first wins the soname, so the
// library is only opened when a slab context is actually created (the Python
// layer pre-loads torch's copy when it can find it).
#include <dlfcn.h>

#include "engine.cuh"

namespace f2d {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return F2D_OK;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("cannot load libnccl.so.2: %s", dlerror()); return F2D_ERR_UNSUPPORTED; }
#define SYM(field, name)                                                             \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));         \
    if (!g_nccl.field) { set_error("libnccl lacks %s", name); return F2D_ERR_UNSUPPORTED; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    g_nccl.handle = h;
    return F2D_OK;
}
#define ncclGetUniqueId g_nccl.GetUniqueId
#define ncclCommInitRank g_nccl.CommInitRank
#define ncclCommDestroy g_nccl.CommDestroy
#define ncclGroupStart g_nccl.GroupStart
#define ncclGroupEnd g_nccl.GroupEnd
#define ncclSend g_nccl.Send
#define ncclRecv g_nccl.Recv
#define ncclAllReduce g_nccl.AllReduce
#define ncclAllGather g_nccl.AllGather
#define ncclGetErrorString g_nccl.GetErrorString

#define F2D_NCCL(call)                                                            \
    do {                                                                          \
        ncclResult_t _r = (call);                                                 \
        if (_r != ncclSuccess) {                                                  \
            set_error("NCCL error %d (%s) at %s:%d", (int)_r, ncclGetErrorString(_r), __FILE__, __LINE__); \
            return F2D_ERR_CUDA;                                                  \
        }                                                                         \
    } while (0)

// Logical rows [0, nrows) of each array start at array row `row0`; the first /
// last G logical rows are ghosts of the south / north neighbour's boundary rows.
int dist_exchange(f2d_ctx *c, int narr, void *const *base, size_t row_bytes, long nrows, long row0) {
    const Dist &D = c->dist;
    if (!D.on || (!D.south && !D.north)) return F2D_OK;
    const size_t bytes = (size_t)D.G * row_bytes;
    F2D_NCCL(ncclGroupStart());
    for (int a = 0; a < narr; a++) {
        char *p = static_cast<char *>(base[a]) + (size_t)row0 * row_bytes;
        if (D.north) {
            F2D_NCCL(ncclSend(p + (size_t)(nrows - 2 * D.G) * row_bytes, bytes, ncclChar, D.rank + 1, D.comm, c->stream));
            F2D_NCCL(ncclRecv(p + (size_t)(nrows - D.G) * row_bytes, bytes, ncclChar, D.rank + 1, D.comm, c->stream));
        }
        if (D.south) {
            F2D_NCCL(ncclSend(p + (size_t)D.G * row_bytes, bytes, ncclChar, D.rank - 1, D.comm, c->stream));
            F2D_NCCL(ncclRecv(p, bytes, ncclChar, D.rank - 1, D.comm, c->stream));
        }
    }
    F2D_NCCL(ncclGroupEnd());
    c->exchanges++;
    return F2D_OK;
}

// several arrays of different geometry in one NCCL group; the arguments point
// into an array of structs with the given stride (bytes)
int dist_exchange_parts(f2d_ctx *c, int n, char *const *base, const size_t *row_bytes, const long *nrows,
                        const long *row0, size_t stride) {
    const Dist &D = c->dist;
    if (!D.on || (!D.south && !D.north)) return F2D_OK;
    auto at = [&](const void *p, int k) { return reinterpret_cast<const char *>(p) + (size_t)k * stride; };
    F2D_NCCL(ncclGroupStart());
    for (int k = 0; k < n; k++) {
        char *b = *reinterpret_cast<char *const *>(at(base, k));
        size_t rb = *reinterpret_cast<const size_t *>(at(row_bytes, k));
        long nr = *reinterpret_cast<const long *>(at(nrows, k)), r0 = *reinterpret_cast<const long *>(at(row0, k));
        char *p = b + (size_t)r0 * rb;
        const size_t bytes = (size_t)D.G * rb;
        if (D.north) {
            F2D_NCCL(ncclSend(p + (size_t)(nr - 2 * D.G) * rb, bytes, ncclChar, D.rank + 1, D.comm, c->stream));
            F2D_NCCL(ncclRecv(p + (size_t)(nr - D.G) * rb, bytes, ncclChar, D.rank + 1, D.comm, c->stream));
        }
        if (D.south) {
            F2D_NCCL(ncclSend(p + (size_t)D.G * rb, bytes, ncclChar, D.rank - 1, D.comm, c->stream));
            F2D_NCCL(ncclRecv(p, bytes, ncclChar, D.rank - 1, D.comm, c->stream));
        }
    }
    F2D_NCCL(ncclGroupEnd());
    c->exchanges++;
    return F2D_OK;
}

int dist_exchange1(f2d_ctx *c, void *base, size_t row_bytes, long nrows, long row0) {
    void *b[1] = {base};
    return dist_exchange(c, 1, b, row_bytes, nrows, row0);
}

int dist_allreduce(f2d_ctx *c, double *d_vals, int n, bool max_op) {
    const Dist &D = c->dist;
    if (!D.on || D.world == 1) return F2D_OK;
    F2D_NCCL(ncclAllReduce(d_vals, d_vals, n, ncclDouble, max_op ? ncclMax : ncclSum, D.comm, c->stream));
    return F2D_OK;
}

// every rank contributes bytes_per_rank from src_rows; dst receives world * bytes_per_rank
int dist_allgather_rows(f2d_ctx *c, const void *src_rows, void *dst, size_t bytes_per_rank) {
    const Dist &D = c->dist;
    if (!D.on || D.world == 1) {
        F2D_CUDA(cudaMemcpyAsync(dst, src_rows, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream));
        return F2D_OK;
    }
    F2D_NCCL(ncclAllGather(src_rows, dst, bytes_per_rank, ncclChar, D.comm, c->stream));
    return F2D_OK;
}

int dist_init(f2d_ctx *c, int rank, int world, const char *unique_id) {
    Dist &D = c->dist;
    if (D.on) { set_error("f2d_dist_init called twice"); return F2D_ERR_STATE; }
    if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank %d / world %d", rank, world); return F2D_ERR_ARG; }
    const int gs = c->cfg.reserved[1], gn = c->cfg.reserved[2];
    if ((rank > 0) != (gs > 0) || (rank < world - 1) != (gn > 0)) {
        set_error("rank %d of %d does not match the ghost rows of the context (south %d, north %d)", rank, world, gs, gn);
        return F2D_ERR_ARG;
    }
    if ((gs && gs != D.G) || (gn && gn != D.G)) { set_error("ghost width must be %d", D.G); return F2D_ERR_ARG; }
    if (c->cfg.model != F2D_MODEL_EULER && c->cfg.model != F2D_MODEL_BOUSSINESQ) {
        set_error("slab decomposition is implemented for the euler and boussinesq models");
        return F2D_ERR_UNSUPPORTED;
    }
    if (c->cfg.yperiodic) { set_error("yperiodic is not supported with slabs"); return F2D_ERR_UNSUPPORTED; }
    F2D_TRY(nccl_load());
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "unique id size");
    memcpy(&id, unique_id, sizeof(id));
    F2D_CUDA(cudaSetDevice(c->cfg.device));
    F2D_NCCL(ncclCommInitRank(&D.comm, world, id, rank));
    D.rank = rank; D.world = world;
    D.south = rank > 0; D.north = rank < world - 1;
    D.on = true;
    return F2D_OK;
}

int dist_unique_id(char *out) {
    F2D_TRY(nccl_load());
    ncclUniqueId id;
    F2D_NCCL(ncclGetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return F2D_OK;
}

void dist_free(f2d_ctx *c) {
    if (c->dist.on && c->dist.comm) ncclCommDestroy(c->dist.comm);
    c->dist = Dist();
}

}  // namespace f2d
