// microbench_fp64.cu — latency and throughput of dependent DADD chains (the Kahan steps of the tie pass) on one SM's worth
// of warps: CHAINS independent chains per thread, ITER x 8 dependent adds per chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench_fp64.bin scripts/microbench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;

template < int CHAINS >
__global__ void __launch_bounds__(1024, 1) kernel(double* out, double seed, long long* clocks) {
    double a[CHAINS];
    #pragma unroll
    for(int c = 0; c < CHAINS; ++c) { a[c] = seed * (threadIdx.x + 1 + c); }
    const double b = seed * 1e-3;
    const long long t0 = clock64();
    for(int i = 0; i < ITER; ++i) {
        #pragma unroll
        for(int u = 0; u < 8; ++u) {
            #pragma unroll
            for(int c = 0; c < CHAINS; ++c) { a[c] = __dadd_rn(a[c], b); }
        }
    }
    const long long t1 = clock64();
    double s = 0;
    #pragma unroll
    for(int c = 0; c < CHAINS; ++c) { s += a[c]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if(threadIdx.x == 0 && blockIdx.x == 0) { *clocks = t1 - t0; }
}

template < int CHAINS >
void run() {
    double* out; long long* clocks; long long host = 0;
    cudaMalloc(&out, 148 * 1024 * 8);
    cudaMalloc(&clocks, 8);
    for(int warps = 1; warps <= 32; warps *= 2) {
        kernel< CHAINS ><<< 148, warps * 32 >>>(out, 1.2345, clocks);
        kernel< CHAINS ><<< 148, warps * 32 >>>(out, 1.2345, clocks);
        cudaMemcpy(&host, clocks, 8, cudaMemcpyDeviceToHost);
        const double adds = double(ITER) * 8 * CHAINS;
        printf("DADD chains/thread %d  warps/SM %2d  clocks %9lld  clocks per dependent step %.2f  warp-DADD/clk/SM %.3f\n", CHAINS, warps, host, host / (double(ITER) * 8), adds * warps / host);
    }
    cudaFree(out); cudaFree(clocks);
}

int main() {
    run< 1 >(); run< 2 >(); run< 4 >();
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
