// Write-only bandwidth microbenchmark with the generator's store pattern:
// P planes of 32-byte rows + one 8-byte weight per event.  nvcc -arch=sm_100a -O3 tools/store_bench.cu -o store_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__device__ __forceinline__ void st_row(double* p, double a, double b, double c, double d) {
    if (MODE == 0) asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    if (MODE == 1) asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    if (MODE == 2) asm volatile("st.global.wt.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    if (MODE == 3) asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
    if (MODE == 4) { asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(a), "d"(b) : "memory"); asm volatile("st.global.cs.v2.f64 [%0], {%1,%2};" ::"l"(p + 2), "d"(c), "d"(d) : "memory"); }
}

// BLOCKED=0: tile = blockIdx + k*grid (what the generator does); 1: each CTA owns a contiguous run of tiles
template <int P, int MODE, int BLOCKED>
__global__ void __launch_bounds__(256) store_kernel(double* w, double* parts, long long n, long long stride) {
    const long long tiles = (n + 255) / 256;
    long long t0 = blockIdx.x, t1 = tiles, dt = gridDim.x;
    if (BLOCKED) { long long per = (tiles + gridDim.x - 1) / gridDim.x; t0 = per * blockIdx.x; t1 = min(tiles, t0 + per); dt = 1; }
    for (long long tile = t0; tile < t1; tile += dt) {
        const long long i = tile * 256 + threadIdx.x;
        if (i < n) {
            const double v = (double)i;
            w[i] = v;
#pragma unroll
            for (int s = 0; s < P; ++s) st_row<MODE>(parts + s * stride + 4 * i, v, v + 1, v + 2, v + 3);
        }
    }
}

template <int P, int MODE, int BLOCKED>
float run(double* w, double* parts, long long n, int grid, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) store_kernel<P, MODE, BLOCKED><<<grid, 256>>>(w, parts, n, 4 * n);
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) store_kernel<P, MODE, BLOCKED><<<grid, 256>>>(w, parts, n, 4 * n);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

// Shared-memory staged TMA bulk stores: each thread writes its rows into a tile image in smem, one
// elected thread issues one cp.async.bulk per plane (+1 for the weights); two tile buffers alternate.
template <int P>
__global__ void __launch_bounds__(256) tma_store_kernel(double* w, double* parts, long long n, long long stride) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int TILE_BYTES = 256 * 8 + P * 256 * 32;
    const long long tiles = (n + 255) / 256;
    int buf = 0;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x, buf ^= 1) {
        unsigned char* base = smem + buf * TILE_BYTES;
        // the bulk copies that read this buffer two iterations ago must have finished reading it
        if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncthreads();
        const long long i = tile * 256 + threadIdx.x;
        const double v = (double)i;
        double* sw = reinterpret_cast<double*>(base);
        sw[threadIdx.x] = v;
#pragma unroll
        for (int s = 0; s < P; ++s) {
            double2* row = reinterpret_cast<double2*>(base + 256 * 8 + s * 256 * 32 + threadIdx.x * 32);
            row[0] = make_double2(v, v + 1);
            row[1] = make_double2(v + 2, v + 3);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const long long first = tile * 256;
            const int cnt = (int)min(256LL, n - first);
            unsigned int sa = (unsigned int)__cvta_generic_to_shared(base);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(w + first), "r"(sa), "r"(cnt * 8) : "memory");
#pragma unroll
            for (int s = 0; s < P; ++s)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(parts + s * stride + 4 * first),
                             "r"(sa + 256 * 8 + s * 256 * 32), "r"(cnt * 32) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int P>
float run_tma(double* w, double* parts, long long n, int grid, int reps) {
    constexpr int SMEM = 2 * (256 * 8 + P * 256 * 32);
    CK(cudaFuncSetAttribute(tma_store_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) tma_store_kernel<P><<<grid, 256, SMEM>>>(w, parts, n, 4 * n);
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) tma_store_kernel<P><<<grid, 256, SMEM>>>(w, parts, n, 4 * n);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

// non-persistent grid: CTA b handles the K consecutive tiles [b*K, (b+1)*K)
template <int P>
__global__ void __launch_bounds__(256) chunk_kernel(double* w, double* parts, long long n, long long stride, int K) {
    const long long tiles = (n + 255) / 256;
    const long long t0 = (long long)blockIdx.x * K, t1 = min(tiles, t0 + K);
    for (long long tile = t0; tile < t1; ++tile) {
        const long long i = tile * 256 + threadIdx.x;
        if (i < n) {
            const double v = (double)i;
            w[i] = v;
#pragma unroll
            for (int s = 0; s < P; ++s) st_row<0>(parts + s * stride + 4 * i, v, v + 1, v + 2, v + 3);
        }
    }
}
// persistent grid with a dynamic tile counter (tiles are claimed in address order)
template <int P>
__global__ void __launch_bounds__(256) dyn_kernel(double* w, double* parts, long long n, long long stride, unsigned long long* counter) {
    __shared__ long long s_tile[2];
    const long long tiles = (n + 255) / 256;
    if (threadIdx.x == 0) s_tile[0] = (long long)atomicAdd(counter, 1ULL);
    __syncthreads();
    int ph = 0;
    while (true) {
        const long long tile = s_tile[ph];
        if (tile >= tiles) break;
        if (threadIdx.x == 0) s_tile[ph ^ 1] = (long long)atomicAdd(counter, 1ULL);  // claim the next one early
        const long long i = tile * 256 + threadIdx.x;
        if (i < n) {
            const double v = (double)i;
            w[i] = v;
#pragma unroll
            for (int s = 0; s < P; ++s) st_row<0>(parts + s * stride + 4 * i, v, v + 1, v + 2, v + 3);
        }
        __syncthreads();
        ph ^= 1;
    }
}

int main(int argc, char** argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 100000000LL;
    constexpr int P = 3;
    double *w, *parts;
    CK(cudaMalloc(&w, n * 8)); CK(cudaMalloc(&parts, n * 32 * P));
    const double gb = (8.0 + 32.0 * P) * n / 1e9;
    int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    // cudaMemset reference
    {
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaMemsetAsync(parts, 0, n * 32 * P)); CK(cudaEventRecord(a));
        for (int i = 0; i < 5; ++i) CK(cudaMemsetAsync(parts, 0, n * 32 * P));
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        printf("cudaMemset %.2f GB: %.3f ms  %.0f GB/s\n", n * 32.0 * P / 1e9, ms / 5, n * 32.0 * P / 1e6 / (ms / 5));
    }
    const int grids[] = {sms * 2, sms * 4, sms * 8, sms * 16, (int)((n + 255) / 256)};
    for (int g : grids) {
        printf("grid=%d (%.1f CTAs/SM)\n", g, (double)g / sms);
#define RUN(MODE, BLOCKED, NAME) { float ms = run<P, MODE, BLOCKED>(w, parts, n, g, 5); printf("  %-34s %.3f ms  %.0f GB/s\n", NAME, ms, gb / ms * 1e3); }
        RUN(0, 0, "cs.v4 strided-tiles");
        RUN(0, 1, "cs.v4 blocked-tiles");
        RUN(1, 0, "default.v4 strided-tiles");
        RUN(2, 0, "wt.v4 strided-tiles");
        RUN(3, 0, "L1::no_allocate.v4 strided-tiles");
        RUN(4, 0, "cs.2xv2 strided-tiles");
        if (g <= sms * 8) { float ms = run_tma<P>(w, parts, n, g, 5); printf("  %-34s %.3f ms  %.0f GB/s\n", "smem + cp.async.bulk (TMA) store", ms, gb / ms * 1e3); }
    }
    for (int K : {1, 2, 4, 8, 16, 32, 64, 128}) {
        const long long tiles = (n + 255) / 256;
        const int g = (int)((tiles + K - 1) / K);
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        chunk_kernel<P><<<g, 256>>>(w, parts, n, 4 * n, K);
        CK(cudaEventRecord(a));
        for (int i = 0; i < 5; ++i) chunk_kernel<P><<<g, 256>>>(w, parts, n, 4 * n, K);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
        printf("chunk K=%-3d grid=%-7d %.3f ms  %.0f GB/s\n", K, g, ms, gb / ms * 1e3);
    }
    {
        unsigned long long* ctr; CK(cudaMalloc(&ctr, 8));
        for (int per : {2, 4, 8}) {
            cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            CK(cudaMemsetAsync(ctr, 0, 8)); dyn_kernel<P><<<sms * per, 256>>>(w, parts, n, 4 * n, ctr);
            CK(cudaEventRecord(a));
            for (int i = 0; i < 5; ++i) { CK(cudaMemsetAsync(ctr, 0, 8)); dyn_kernel<P><<<sms * per, 256>>>(w, parts, n, 4 * n, ctr); }
            CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
            float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= 5;
            printf("dynamic counter, %d CTAs/SM: %.3f ms  %.0f GB/s\n", per, ms, gb / ms * 1e3);
        }
    }
    return 0;
}
