// microbench.cu — per-SM throughput of the integer instructions the pruned whitelist scan leans on
// (POPC, LOP3, SHF, IMAD, VIMNMX, uniform LDS), in warp instructions per clock per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench.bin scripts/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
constexpr int UNROLL = 8;

template < int OP >
__global__ void __launch_bounds__(1024, 1) kernel(uint32_t* out, uint32_t seed, long long* clocks) {
    __shared__ uint4 table[256];
    if(threadIdx.x < 256) { table[threadIdx.x] = make_uint4(threadIdx.x * seed, seed, threadIdx.x, 7u); }
    __syncthreads();
    uint32_t a[UNROLL];
    #pragma unroll
    for(int u = 0; u < UNROLL; ++u) { a[u] = seed * (threadIdx.x + 1) + u; }
    uint32_t b = seed ^ 0x5555u, c = seed | 3u;
    const long long t0 = clock64();
    for(int i = 0; i < ITER; ++i) {
        #pragma unroll
        for(int u = 0; u < UNROLL; ++u) {
            if(OP == 0) { a[u] = __popc(a[u]) + b; }                                  // POPC + IADD
            if(OP == 1) { asm volatile("lop3.b32 %0, %0, %1, %2, 0xBE;" : "+r"(a[u]) : "r"(b), "r"(c)); }
            if(OP == 2) { a[u] = __funnelshift_r(a[u], b, 7); }                        // SHF
            if(OP == 3) { a[u] = a[u] * c + b; }                                      // IMAD
            if(OP == 4) { a[u] = min(min(a[u], b), c + u); }                          // VIMNMX3?
            if(OP == 5) { const uint4 v = table[(i + u) & 255]; a[u] ^= v.x + v.y + v.z + v.w; }   // uniform LDS.128 + 3 IADD3/LOP
            if(OP == 6) { a[u] = __popc(a[u] & b); }                                  // LOP + POPC chain
            if(OP == 7) { a[u] = (a[u] + b) ^ c; }                                    // IADD + LOP baseline for OP 0
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    #pragma unroll
    for(int u = 0; u < UNROLL; ++u) { s ^= a[u]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if(threadIdx.x == 0 && blockIdx.x == 0) { *clocks = t1 - t0; }
}

template < int OP >
void run(const char* name, int ops_per_iteration) {
    uint32_t* out; long long* clocks; long long host = 0;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&clocks, 8);
    for(int warps = 4; warps <= 32; warps *= 2) {
        kernel< OP ><<< 148, warps * 32 >>>(out, 12345u, clocks);
        kernel< OP ><<< 148, warps * 32 >>>(out, 12345u, clocks);
        cudaMemcpy(&host, clocks, 8, cudaMemcpyDeviceToHost);
        const double warp_instructions = double(ITER) * UNROLL * ops_per_iteration * warps;
        printf("%-28s warps/SM %2d  clocks %9lld  warp-instr/clk/SM %.3f (counting %d instr per step)\n", name, warps, host, warp_instructions / host, ops_per_iteration);
    }
    cudaFree(out); cudaFree(clocks);
}

int main() {
    run< 0 >("POPC+IADD", 2);
    run< 7 >("IADD+LOP (baseline)", 2);
    run< 6 >("LOP+POPC", 2);
    run< 1 >("LOP3", 1);
    run< 2 >("SHF", 1);
    run< 3 >("IMAD", 1);
    run< 4 >("MIN3", 1);
    run< 5 >("uniform LDS.128 + 4 int", 5);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
