// Deterministic grid-wide reductions: warp shuffles -> shared memory -> one
// partial per block -> the last block to finish folds the partials in index
// order.  No floating-point atomics, so results are bit-reproducible run to run.
#pragma once
#include <cuda_runtime.h>

namespace f2d {

struct OpSum {
    __device__ static double id() { return 0.0; }
    __device__ static double ap(double a, double b) { return a + b; }
};
struct OpMax {
    __device__ static double id() { return 0.0; }
    __device__ static double ap(double a, double b) { return fmax(a, b); }
};

__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

template <class Op>
__device__ __forceinline__ double warp_reduce(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = Op::ap(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Reduce NV values per thread across the block; result valid in thread 0.
template <class Op, int NV>
__device__ __forceinline__ void block_reduce(double (&v)[NV]) {
    __shared__ double sm[NV][32];
    int tid = threadIdx.y * blockDim.x + threadIdx.x;
    int lane = tid & 31, wid = tid >> 5;
    int nw = (blockDim.x * blockDim.y + 31) >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
        v[k] = warp_reduce<Op>(v[k]);
        if (lane == 0) sm[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double t = lane < nw ? sm[k][lane] : Op::id();
            v[k] = warp_reduce<Op>(t);
        }
    }
    __syncthreads();
}

// Two-kernel variant for kernels whose CTAs are short-lived (one tile each): the
// CTA only stores its partial and leaves -- no fence, no atomic round trip, no
// second barrier while it holds its slot -- and k_fold_partials, launched right
// behind it, folds the partials in a fixed order.
template <class Op, int NV>
__device__ __forceinline__ void block_partials(double (&v)[NV], double *part) {
    block_reduce<Op, NV>(v);
    if (threadIdx.y * blockDim.x + threadIdx.x == 0) {
        unsigned nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
#pragma unroll
        for (int k = 0; k < NV; k++) part[(size_t)k * nblocks + bid] = v[k];
    }
}

template <class Op, int NV>
__global__ void __launch_bounds__(1024) k_fold_partials(const double *__restrict__ part, unsigned nblocks, double *out) {
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        acc[k] = Op::id();
        for (unsigned b = threadIdx.x; b < nblocks; b += blockDim.x)
            acc[k] = Op::ap(acc[k], part[(size_t)k * nblocks + b]);
    }
    block_reduce<Op, NV>(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) out[k] = acc[k];
    }
}

// Grid-wide: every block calls this with its per-thread values.  `part` holds
// NV * nblocks doubles, `count` one zero-initialised counter (reset on exit),
// `out[k]` receives the result.  nblocks = gridDim.x*gridDim.y.
template <class Op, int NV>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double *part, unsigned int *count,
                                            double *out) {
    block_reduce<Op, NV>(v);
    int tid = threadIdx.y * blockDim.x + threadIdx.x;
    unsigned nblocks = gridDim.x * gridDim.y;
    unsigned bid = blockIdx.y * gridDim.x + blockIdx.x;
    __shared__ bool last;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) part[(size_t)k * nblocks + bid] = v[k];
        // release the partials / acquire everybody else's: fence.acq_rel (MEMBAR.ALL.GPU)
        // is enough for this message-passing pattern and much cheaper than the
        // sequentially consistent fence __threadfence() compiles to (MEMBAR.SC.GPU),
        // which every thread of the CTA would sit behind at the barrier below
        fence_acq_rel_gpu();
        unsigned t = atomicAdd(count, 1u);
        last = (t == nblocks - 1);
        if (t == nblocks - 1) fence_acq_rel_gpu();
    }
    __syncthreads();
    if (!last) return;
    int nt = blockDim.x * blockDim.y;
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
        acc[k] = Op::id();
        // fixed order: thread t folds partials t, t+nt, ... then the block tree
        for (unsigned b = tid; b < nblocks; b += nt)
            acc[k] = Op::ap(acc[k], __ldcg(&part[(size_t)k * nblocks + b]));
    }
    block_reduce<Op, NV>(acc);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; k++) out[k] = acc[k];
        *count = 0;
    }
}

}  // namespace f2d
