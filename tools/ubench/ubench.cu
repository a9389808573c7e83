// DEVELOPMENT TOOL: latency micro-measurements on the target GPU that the solve-kernel design
// depends on (dependent fp64 chains, shuffles, rsqrt, barriers, L2-hit pointer chasing).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, const int* chain, int n_chain, long long* cyc) {
    __shared__ double sh[4096];
    double x = out[threadIdx.x], y = 1.000001;
    long long t0, t1;
    int it = 4096;
    // 1. dependent DFMA chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = fma(x, y, 1e-9);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0);
    // 2. shfl + dfma chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = fma(__shfl_sync(0xffffffffu, x, i & 31), y, x);
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0);
    // 3. rsqrt(double) chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = rsqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0);
    // 4. 1/sqrt chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = 1.0 / sqrt(x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0);
    // 5. __syncthreads
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0);
    // 6. dependent global pointer chase (L2 hits after the first pass; chain is large enough to miss L1)
    int p = threadIdx.x;
    for (int i = 0; i < 2048; ++i) p = chain[p];
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) p = chain[p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0);
    // 7. dependent shared-memory chase
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = (double)((i * 7 + 3) & 4095);
    __syncthreads(); t0 = clock64();
    int q = threadIdx.x;
    for (int i = 0; i < it; ++i) q = (int)sh[q];
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0);
    // 8. double division chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = 1.0 / (x + 2.0);
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = (t1 - t0);
    // 9. subwarp reduction (5 shuffles + adds)
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) { for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o); x *= 1e-3; }
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = (t1 - t0);
    // 10. DADD chain
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) x = x + y;
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = (t1 - t0);
    // 11. small-chain pointer chase (fits L1)
    p = threadIdx.x & 1023;
    for (int i = 0; i < 2048; ++i) p = chain[n_chain + p];
    __syncthreads(); t0 = clock64();
    for (int i = 0; i < it; ++i) p = chain[n_chain + p];
    t1 = clock64(); if (threadIdx.x == 0) cyc[10] = (t1 - t0);
    out[threadIdx.x] = x + p + q;
}
int main() {
    const int n = 1 << 22;  // 16 MB of int: misses L1, stays in L2
    int* h = new int[n + 1024];
    // single random cycle
    unsigned s = 12345u; int* perm = new int[n];
    for (int i = 0; i < n; ++i) perm[i] = i;
    for (int i = n - 1; i > 0; --i) { s = s * 1664525u + 1013904223u; int j = s % (i + 1); int t = perm[i]; perm[i] = perm[j]; perm[j] = t; }
    for (int i = 0; i < n; ++i) h[perm[i]] = perm[(i + 1) % n];
    for (int i = 0; i < 1024; ++i) h[n + i] = (i * 7 + 3) & 1023;
    int* d; double* o; long long* c;
    cudaMalloc(&d, (n + 1024) * 4); cudaMalloc(&o, 1024 * 8); cudaMalloc(&c, 16 * 8);
    cudaMemcpy(d, h, (n + 1024) * 4, cudaMemcpyHostToDevice); cudaMemset(o, 0, 1024 * 8);
    const char* nm[] = {"dfma chain", "shfl+dfma chain", "rsqrt(double) chain", "1/sqrt chain", "__syncthreads", "L2-hit pointer chase",
                        "smem chase (+cvt)", "double division chain", "warp reduce (5 shfl+add) + mul", "dadd chain", "L1-hit pointer chase"};
    for (int threads : {32, 256, 512}) {
        k<<<1, threads>>>(o, d, n, c); cudaDeviceSynchronize();
        k<<<1, threads>>>(o, d, n, c); cudaDeviceSynchronize();
        long long hc[16]; cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
        printf("threads %d (%s)\n", threads, cudaGetErrorString(cudaGetLastError()));
        for (int i = 0; i < 11; ++i) printf("  %-34s %8.1f cycles/op\n", nm[i], hc[i] / 4096.0);
    }
    return 0;
}
