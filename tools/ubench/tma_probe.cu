// TMA OOB probe: ./tma_probe W3 rows pitch boxB boxR x y
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../comfyui-distributed_b200/csrc/usdu_tma.cuh"
using namespace usdu;
__global__ void k(const __grid_constant__ CUtensorMap m, int x, int y, int bytes, unsigned* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) tma::mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { tma::mbar_expect_tx(&bar, bytes); tma::load_2d(sm, &m, x, y, &bar); }
    tma::mbar_wait(&bar, 0);
    unsigned s = 0;
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) s += sm[i];
    atomicAdd(out, s);
}
int main(int argc, char** argv) {
    int W3 = atoi(argv[1]), rows = atoi(argv[2]), pitch = atoi(argv[3]), bB = atoi(argv[4]), bR = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
    unsigned char* d; cudaMalloc(&d, (size_t)pitch * rows); cudaMemset(d, 1, (size_t)pitch * rows);
    unsigned* out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
    CUtensorMap m; if (!tma::encode_u8_2d(&m, d, W3, rows, pitch, bB, bR)) { printf("encode failed\n"); return 1; }
    k<<<1, 128, bB * bR>>>(m, x, y, bB * bR, out);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned h = 0; cudaMemcpy(&h, out, 4, cudaMemcpyDeviceToHost);
    printf("W3=%d rows=%d box=%dx%d x=%d y=%d -> %s sum=%u\n", W3, rows, bB, bR, x, y, cudaGetErrorString(e), h);
    return 0;
}
