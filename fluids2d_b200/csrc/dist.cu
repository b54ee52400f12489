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

#include <algorithm>
#include <cstdlib>

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

// ---------------------------------------------------------------------------
// Peer-to-peer ghost exchange.  NCCL send/recv costs ~20 us per exchange on
// this box (its own kernel + protocol), and a PCG iteration needs ~14 of them.
// Every array that is exchanged in the time loop is a cudaMalloc allocation,
// so each rank maps its two neighbours' copies (CUDA IPC over NVLink) once at
// set-up and an exchange becomes ONE small kernel that
//   1. tells both neighbours "my ghost rows may be overwritten" (READY),
//   2. waits for their READY, then stores its boundary rows straight into
//      their ghost rows (coalesced 8-byte stores through the peer mapping),
//   3. fences, signals DONE and waits for the neighbours' DONE.
// Flags are epoch counters in device memory (so the kernel is CUDA-graph safe);
// every wait is bounded and raises an error flag instead of hanging.
// ---------------------------------------------------------------------------
enum { FL_READY_S = 0, FL_READY_N = 16, FL_DONE_S = 32, FL_DONE_N = 48, FL_EPOCH = 64, FL_ARRIVE = 80,
       FL_ERROR = 96, FL_WORDS = 112 };

struct P2PPart {
    const char *src_n, *src_s;     // my boundary rows next to the north / south interface
    char *dst_n, *dst_s;           // the neighbours' ghost rows (peer mappings)
    size_t bytes;                  // G * row_bytes
};
struct P2PArgs {
    int np, has_south, has_north;
    P2PPart part[4];
    unsigned long long *mine, *south, *north;   // flag blocks
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ bool wait_flag(const unsigned long long *p, unsigned long long want, unsigned long long *err) {
    long long t0 = clock64();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000LL) { atomicExch(err, 1ULL); return false; }   // ~2 s: give up
    }
    return true;
}

__global__ void __launch_bounds__(256) k_p2p_exchange(P2PArgs A) {
    __shared__ unsigned long long epoch;
    if (threadIdx.x == 0) {
        epoch = ld_acquire_sys(A.mine + FL_EPOCH) + 1;
        if (blockIdx.x == 0) {          // my ghost rows are free: every earlier kernel of this stream is done
            if (A.has_north) st_release_sys(A.north + FL_READY_S, epoch);
            if (A.has_south) st_release_sys(A.south + FL_READY_N, epoch);
        }
        if (A.has_north) wait_flag(A.mine + FL_READY_N, epoch, A.mine + FL_ERROR);
        if (A.has_south) wait_flag(A.mine + FL_READY_S, epoch, A.mine + FL_ERROR);
    }
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < A.np; k++) {
        const P2PPart &P = A.part[k];
        const size_t n8 = P.bytes >> 3;
        if (A.has_north) {
            const unsigned long long *s = reinterpret_cast<const unsigned long long *>(P.src_n);
            unsigned long long *d = reinterpret_cast<unsigned long long *>(P.dst_n);
            for (size_t i = tid; i < n8; i += nt) d[i] = s[i];
        }
        if (A.has_south) {
            const unsigned long long *s = reinterpret_cast<const unsigned long long *>(P.src_s);
            unsigned long long *d = reinterpret_cast<unsigned long long *>(P.dst_s);
            for (size_t i = tid; i < n8; i += nt) d[i] = s[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = atomicAdd(A.mine + FL_ARRIVE, 1ULL);
        if (t == gridDim.x - 1) {       // last CTA: everybody's stores are fenced
            if (A.has_north) st_release_sys(A.north + FL_DONE_S, epoch);
            if (A.has_south) st_release_sys(A.south + FL_DONE_N, epoch);
            if (A.has_north) wait_flag(A.mine + FL_DONE_N, epoch, A.mine + FL_ERROR);
            if (A.has_south) wait_flag(A.mine + FL_DONE_S, epoch, A.mine + FL_ERROR);
            A.mine[FL_ARRIVE] = 0;
            st_release_sys(A.mine + FL_EPOCH, epoch);
        }
    }
}

struct PeerRecord {
    cudaIpcMemHandle_t handle;
    long long rows_total;      // rows of the array (coarse arrays: including their two halo rows)
    long long pad;             // 1 for halo-padded coarse arrays, 0 for (n2, n1) arrays
    long long row_bytes;
};

void p2p_teardown(f2d_ctx *c) {
    Dist &D = c->dist;
    for (void *p : D.peer_south) if (p) cudaIpcCloseMemHandle(p);
    for (void *p : D.peer_north) if (p) cudaIpcCloseMemHandle(p);
    D.peer_south.clear(); D.peer_north.clear(); D.index.clear();
    D.rec_south.clear(); D.rec_north.clear();
    if (D.flags) { cudaFree(D.flags); D.flags = nullptr; }
    D.p2p = false;
}

// Map the neighbours' copies of every array listed in `arrays` (same order on
// every rank).  rows_total / pad / row_bytes describe each array's geometry.
int p2p_setup(f2d_ctx *c, const std::vector<void *> &arrays, const std::vector<long long> &rows_total,
              const std::vector<long long> &pad, const std::vector<long long> &row_bytes) {
    Dist &D = c->dist;
    p2p_teardown(c);
    // Opt-in (F2D_P2P=1).  Measured on 2 and 4 B200s: 19.1 / 19.6 ms per step
    // against 19.3 / 18.9 ms with NCCL send/recv inside the CUDA graph -- both
    // are two NVLink flag round trips plus one launch per exchange, so the peer
    // path buys nothing here and NCCL (validated up to 8 GPUs) stays the default.
    static const bool on = getenv("F2D_P2P") != nullptr;
    if (!D.on || D.world == 1 || !on) return F2D_OK;
    const int K = (int)arrays.size() + 1;      // + the flag block
    F2D_CUDA(cudaMalloc(&D.flags, FL_WORDS * sizeof(unsigned long long)));
    F2D_CUDA(cudaMemsetAsync(D.flags, 0, FL_WORDS * sizeof(unsigned long long), c->stream));
    std::vector<PeerRecord> mine(K);
    memset(mine.data(), 0, K * sizeof(PeerRecord));
    for (int k = 0; k < K; k++) {
        void *p = k == 0 ? (void *)D.flags : arrays[k - 1];
        cudaError_t e = cudaIpcGetMemHandle(&mine[k].handle, p);
        if (e != cudaSuccess) { cudaGetLastError(); return F2D_OK; }   // no IPC here: stay on NCCL
        if (k > 0) { mine[k].rows_total = rows_total[k - 1]; mine[k].pad = pad[k - 1]; mine[k].row_bytes = row_bytes[k - 1]; }
    }
    PeerRecord *d_mine, *d_all;
    const size_t bytes = (size_t)K * sizeof(PeerRecord);
    F2D_CUDA(cudaMalloc(&d_mine, bytes));
    F2D_CUDA(cudaMalloc(&d_all, bytes * D.world));
    F2D_CUDA(cudaMemcpyAsync(d_mine, mine.data(), bytes, cudaMemcpyHostToDevice, c->stream));
    F2D_NCCL(ncclAllGather(d_mine, d_all, bytes, ncclChar, D.comm, c->stream));
    std::vector<PeerRecord> all((size_t)K * D.world);
    F2D_CUDA(cudaMemcpyAsync(all.data(), d_all, bytes * D.world, cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_mine); cudaFree(d_all);
    D.peer_south.assign(K, nullptr); D.peer_north.assign(K, nullptr);
    D.rec_south.assign(K, {0, 0, 0}); D.rec_north.assign(K, {0, 0, 0});
    bool ok = true;
    for (int k = 0; k < K && ok; k++) {
        if (D.south) {
            const PeerRecord &r = all[(size_t)(D.rank - 1) * K + k];
            if (cudaIpcOpenMemHandle(&D.peer_south[k], r.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
            D.rec_south[k] = {r.rows_total, r.pad, r.row_bytes};
        }
        if (D.north && ok) {
            const PeerRecord &r = all[(size_t)(D.rank + 1) * K + k];
            if (cudaIpcOpenMemHandle(&D.peer_north[k], r.handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = false;
            D.rec_north[k] = {r.rows_total, r.pad, r.row_bytes};
        }
    }
    // every rank must take the same path
    double flag = ok ? 0.0 : 1.0;
    F2D_CUDA(cudaMemcpyAsync(c->d_scal + 25, &flag, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    F2D_TRY(dist_allreduce(c, c->d_scal + 25, 1, true));
    F2D_CUDA(cudaMemcpyAsync(&flag, c->d_scal + 25, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    if (flag != 0.0) { cudaGetLastError(); p2p_teardown(c); return F2D_OK; }
    for (int k = 1; k < K; k++) D.index[arrays[k - 1]] = k;
    D.p2p = true;
    return F2D_OK;
}

// all parts mapped on the neighbours?  then one kernel does the exchange
static bool p2p_try(f2d_ctx *c, int n, char *const *base, const size_t *row_bytes, const long *nrows,
                    const long *row0, size_t stride, int *status) {
    Dist &D = c->dist;
    if (!D.p2p || n > 4) return false;
    auto at = [&](const void *p, int k) { return reinterpret_cast<const char *>(p) + (size_t)k * stride; };
    P2PArgs A;
    A.np = n; A.has_south = D.south; A.has_north = D.north;
    A.mine = D.flags;
    A.south = reinterpret_cast<unsigned long long *>(D.peer_south[0]);
    A.north = reinterpret_cast<unsigned long long *>(D.peer_north[0]);
    for (int k = 0; k < n; k++) {
        char *b = *reinterpret_cast<char *const *>(at(base, k));
        auto it = D.index.find(b);
        if (it == D.index.end()) return false;
        const int idx = it->second;
        const size_t rb = *reinterpret_cast<const size_t *>(at(row_bytes, k));
        const long nr = *reinterpret_cast<const long *>(at(nrows, k)), r0 = *reinterpret_cast<const long *>(at(row0, k));
        P2PPart &P = A.part[k];
        P.bytes = (size_t)D.G * rb;
        if (P.bytes & 7) return false;
        P.src_n = b + (size_t)(r0 + nr - 2 * D.G) * rb;
        P.src_s = b + (size_t)(r0 + D.G) * rb;
        P.dst_n = P.dst_s = nullptr;
        if (D.north) {      // the north neighbour's south ghost rows start right after its pad rows
            if ((size_t)D.rec_north[idx].row_bytes != rb) return false;
            P.dst_n = static_cast<char *>(D.peer_north[idx]) + (size_t)D.rec_north[idx].pad * rb;
        }
        if (D.south) {      // the south neighbour's north ghost rows are its last G rows before the pad
            if ((size_t)D.rec_south[idx].row_bytes != rb) return false;
            P.dst_s = static_cast<char *>(D.peer_south[idx]) +
                      (size_t)(D.rec_south[idx].rows_total - D.rec_south[idx].pad - D.G) * rb;
        }
    }
    size_t maxb = 0;
    for (int k = 0; k < n; k++) maxb = std::max(maxb, A.part[k].bytes);
    int nblk = (int)std::min<size_t>(16, std::max<size_t>(1, maxb / (256 * 8 * 4)));
    k_p2p_exchange<<<nblk, 256, 0, c->stream>>>(A);
    c->launches++;
    c->exchanges++;
    cudaError_t e = cudaGetLastError();
    *status = e == cudaSuccess ? F2D_OK : cuda_fail(e, "k_p2p_exchange", __FILE__, __LINE__);
    return true;
}

int p2p_check(f2d_ctx *c) {
    Dist &D = c->dist;
    if (!D.p2p) return F2D_OK;
    unsigned long long err = 0;
    F2D_CUDA(cudaMemcpyAsync(&err, D.flags + FL_ERROR, sizeof(err), cudaMemcpyDeviceToHost, c->stream));
    F2D_CUDA(cudaStreamSynchronize(c->stream));
    if (err) { set_error("peer-to-peer ghost exchange timed out waiting for a neighbour"); return F2D_ERR_CUDA; }
    return F2D_OK;
}

// Logical rows [0, nrows) of each array start at array row `row0`; the first /
// last G logical rows are ghosts of the south / north neighbour's boundary rows.
int dist_exchange(f2d_ctx *c, int narr, void *const *base, size_t row_bytes, long nrows, long row0) {
    const Dist &D = c->dist;
    if (!D.on || (!D.south && !D.north)) return F2D_OK;
    if (D.p2p && narr <= 4) {
        struct Part { char *base; size_t row_bytes; long nrows, row0; } parts[4];
        for (int a = 0; a < narr; a++) parts[a] = Part{static_cast<char *>(base[a]), row_bytes, nrows, row0};
        int st = F2D_OK;
        if (p2p_try(c, narr, &parts[0].base, &parts[0].row_bytes, &parts[0].nrows, &parts[0].row0, sizeof(Part), &st)) return st;
    }
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
    {
        int st = F2D_OK;
        if (p2p_try(c, n, base, row_bytes, nrows, row0, stride, &st)) return st;
    }
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
    if (c->cfg.model != F2D_MODEL_EULER && c->cfg.model != F2D_MODEL_BOUSSINESQ && c->cfg.model != F2D_MODEL_RSW &&
        c->cfg.model != F2D_MODEL_QGRSW) {
        set_error("slab decomposition is implemented for the euler, boussinesq, rsw and qgrsw models");
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
    p2p_teardown(c);
    if (c->dist.on && c->dist.comm) ncclCommDestroy(c->dist.comm);
    c->dist = Dist();
}

}  // namespace f2d
