// fp64 pipe microbenchmark for the B200: DFMA throughput as a function of resident warps and ILP, alone and
// interleaved with independent integer multiply-adds (IMAD).  Answers: (1) the fp64 peak (DFMA per clock per SM);
// (2) whether an fp64 instruction costs the warp scheduler one issue slot or two, i.e. whether a 1:1 mix of fp64 and
// other instructions can run at 1 instruction per clock per scheduler or only at 2/3.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/fp64_bench tools/fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() {
    long long c;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)::"memory");
    return c;
}
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
    return t;
}

template <int ILP, int IPF>
__global__ void __launch_bounds__(256) k(double* out, unsigned* iout, long long* cyc, int iters, double a, double b, unsigned m) {
    const long long c0 = clk();
    const unsigned long long g0 = gtime();
    double acc[ILP];
    unsigned xi[ILP][IPF > 0 ? IPF : 1];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
        for (int q = 0; q < IPF; ++q) xi[i][q] = threadIdx.x + 17 * i + q;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                acc[i] = fma(acc[i], a, b);
#pragma unroll
                for (int q = 0; q < IPF; ++q) xi[i][q] = xi[i][q] * m + 0x7F4A7C15u;
            }
        }
    }
    const long long c1 = clk();
    const unsigned long long g1 = gtime();
    if (threadIdx.x == 0) {
        cyc[2 * blockIdx.x] = c1 - c0;
        cyc[2 * blockIdx.x + 1] = (long long)(g1 - g0);
    }
    double s = 0;
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        s += acc[i];
#pragma unroll
        for (int q = 0; q < IPF; ++q) x ^= xi[i][q];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = x;
}

template <int ILP, int IPF>
void run(int ctas_per_sm, int sms, double* out, unsigned* iout, long long* cyc) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ILP, IPF><<<sms * ctas_per_sm, 256>>>(out, iout, cyc, 10, 0.999, 1e-3, 0x9E3779B9u);
    cudaEventRecord(e0);
    k<ILP, IPF><<<sms * ctas_per_sm, 256>>>(out, iout, cyc, iters, 0.999, 1e-3, 0x9E3779B9u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c[2];
    cudaMemcpy(c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double ghz = (double)c[0] / (double)c[1];  // SM clock seen by CTA 0 (clock64 ticks per globaltimer ns)
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k<ILP, IPF>, 256, 0);
    // per SMSP: 2 * ctas_per_sm warps, each issuing iters * 8 * ILP DFMA (+ IPF IMAD each), in ms * ghz * 1e6 cycles
    const double warp_fma_per_smsp = 2.0 * ctas_per_sm * iters * 8.0 * ILP;
    const double ipc_fp64 = warp_fma_per_smsp / (ms * ghz * 1e6);
    printf("warps/SMSP=%2d (resident CTAs/SM possible %d) ILP=%d IMAD-per-DFMA=%d : %.3f ms  SM clock %.3f GHz  fp64 warp-inst/clk/SMSP %.3f (DFMA/clk/SM %.1f)  all warp-inst/clk/SMSP %.3f   %.2f TDFMA/s\n",
           ctas_per_sm * 2, resident, ILP, IPF, ms, ghz, ipc_fp64, ipc_fp64 * 128, ipc_fp64 * (1 + IPF),
           warp_fma_per_smsp * 32 * 4 * sms / ms * 1e-9);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    unsigned* iout;
    long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaMalloc(&iout, sizeof(unsigned) * sms * 8 * 256);
    cudaMalloc(&cyc, sizeof(long long) * sms * 8 * 2);
    for (int c : {2, 4, 8}) {
        run<1, 0>(c, sms, out, iout, cyc);
        run<2, 0>(c, sms, out, iout, cyc);
        run<4, 0>(c, sms, out, iout, cyc);
    }
    for (int c : {2, 4, 8}) {
        run<4, 1>(c, sms, out, iout, cyc);
        run<4, 2>(c, sms, out, iout, cyc);
        run<4, 3>(c, sms, out, iout, cyc);
    }
    return 0;
}
