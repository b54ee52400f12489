// Tensor Memory Accelerator plumbing (sm_90+/sm_100a): 2-D tiled tensor maps over
// the reference-layout (n2, n1) arrays, and the few PTX wrappers a kernel needs to
// pull a box of such an array into shared memory with ONE instruction issued by one
// thread:  cp.async.bulk.tensor.2d (UTMALDG in SASS) completing on an mbarrier.
//
// Why it fits this path: every stencil kernel reads a (TY + 2h) x (TX + 2h) window of
// several fields per output tile.  With TMA the window arrives without a single
// per-thread address computation or bounds test -- coordinates outside the array are
// filled with zeros by the hardware, which is exactly what the reference's guarded
// index windows (operators.py:49-77) need at the array edge.
//
// cuTensorMapEncodeTiled is a driver entry point; libf2d.so links the runtime
// statically and does not link libcuda, so the symbol is resolved through
// cudaGetDriverEntryPoint at first use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace f2d {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// nullptr when the driver has no tensor-map support (callers fall back to plain loads)
inline EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// Tensor map of a row-major (rows, cols) array of `elem_bytes`-wide elements (8: fp64,
// 4: fp32, 1: bytes), row pitch `pitch_elems`, fetched in boxes of box_rows x box_cols.
// Requirements of the hardware: base 16-byte aligned, pitch in bytes a multiple of 16,
// box_cols * elem_bytes a multiple of 16, box dims <= 256.  Returns false if they do not
// hold (or there is no encoder).  At load time the INNER coordinate of a box must also be a
// multiple of 16 bytes (an even column for fp64): an odd one raises "illegal instruction"
// on sm_100 (measured with scripts/tma_probe.cu; negative and out-of-range coordinates are
// fine and zero-filled).  Row coordinates are unconstrained.
inline bool tma_make_2d(CUtensorMap *map, const void *base, int elem_bytes, long rows, long cols, long pitch_elems,
                        int box_rows, int box_cols) {
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((pitch_elems * elem_bytes) & 15) || ((box_cols * elem_bytes) & 15) ||
        box_rows > 256 || box_cols > 256)
        return false;
    CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)(pitch_elems * elem_bytes)};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// make the initialised barrier visible to the async (TMA) proxy
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// box whose first element is (row c1, column c0) of the mapped array -> shared memory
// (densely packed, box_cols elements per row); completes `bytes` on the barrier
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
#endif

}  // namespace f2d
