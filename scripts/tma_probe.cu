// stand-alone probe of the TMA box load used by k_stage_tma (build: nvcc -gencode arch=compute_100a,code=sm_100a)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../fluids2d_b200/csrc/tma.cuh"
using namespace f2d;

struct Maps { CUtensorMap a, b; };

template <int BW, int BH>
__global__ void probe(const __grid_constant__ Maps M, int c0, int c1, double *out, int which) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    double *s = reinterpret_cast<double *>(smem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, BW * BH * 8);
        tma_load_2d(s, which ? &M.b : &M.a, c0, c1, &bar);
    }
    mbar_wait(&bar, 0);
    for (int t = threadIdx.x; t < BW * BH; t += blockDim.x) out[t] = s[t];
}

int main(int argc, char **argv) {
    const int n2 = 46, n1 = 46;
    std::vector<double> h(n2 * n1);
    for (int k = 0; k < n2 * n1; k++) h[k] = k;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&o, 70 * 22 * 8);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    Maps M;
    bool ok1 = tma_make_2d(&M.a, d, 8, n2, n1, n1, 22, 70);
    bool ok2 = tma_make_2d(&M.b, d, 8, n2, n1, n1, 18, 66);
    printf("encode: %d %d\n", ok1, ok2);
    int which = argc > 1 ? atoi(argv[1]) : 0, c0 = argc > 2 ? atoi(argv[2]) : 0, c1 = argc > 3 ? atoi(argv[3]) : 0;
    cudaMemset(o, 0, 70 * 22 * 8);
    if (which == 0) probe<70, 22><<<1, 128, 70 * 22 * 8>>>(M, c0, c1, o, 0);
    else probe<66, 18><<<1, 128, 66 * 18 * 8>>>(M, c0, c1, o, 1);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<double> r(70 * 22);
    cudaMemcpy(r.data(), o, r.size() * 8, cudaMemcpyDeviceToHost);
    int bw = which ? 66 : 70;
    printf("which %d c0 %d c1 %d: %s  s[5*bw+7]=%g (expect %g)\n", which, c0, c1, cudaGetErrorString(e),
           r[5 * bw + 7], (double)((c1 + 5) * n1 + c0 + 7));
    return e != cudaSuccess;
}
