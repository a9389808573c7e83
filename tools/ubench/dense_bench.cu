// DEVELOPMENT TOOL: cycles of the dense-tail routines of csrc/chol.cuh on one CTA, checked against a host Cholesky.
#include <cstdio>
#include <cmath>
#include <vector>
#include "../../sqpsolver.jl_b200/csrc/chol.cuh"

__global__ void k(const double* A, int Tn, double* out, double* rhs, long long* cyc, int reps) {
    extern __shared__ double sm[];
    const int Tp = (Tn + 3) & ~3;
    double* D = sm;
    double* dinv = sm + Tp * (Tp + 1) / 2;
    double* y = dinv + Tn;
    long long tf = 0, ts = 0;
    for (int r = 0; r < reps; ++r) {
        for (int i = threadIdx.x; i < Tp * (Tp + 1) / 2; i += blockDim.x) D[i] = i < Tn * (Tn + 1) / 2 ? A[i] : 0.0;
        __syncthreads();
        for (int i = Tn + threadIdx.x; i < Tp; i += blockDim.x) D[i * (i + 1) / 2 + i] = 1.0;
        for (int i = threadIdx.x; i < Tn; i += blockDim.x) y[i] = rhs[i];
        __syncthreads();
        long long t0 = clock64();
        dense_factor(D, dinv, Tn);
        long long t1 = clock64();
        if (threadIdx.x < 32) dense_solve_warp(D, dinv, y, Tn);
        __syncthreads();
        long long t2 = clock64();
        tf += t1 - t0; ts += t2 - t1;
    }
    for (int i = threadIdx.x; i < Tn; i += blockDim.x) out[i] = y[i];
    if (threadIdx.x == 0) { cyc[0] = tf / reps; cyc[1] = ts / reps; }
}

int main() {
    for (int Tn : {48, 90, 92, 96, 128})
        for (int threads : {256, 512}) {
            std::vector<double> M(Tn * Tn, 0.0), P(Tn * (Tn + 1) / 2), b(Tn), x(Tn);
            unsigned s = 1u;
            auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (double)(s >> 8) / (1 << 24) - 0.5; };
            std::vector<double> G(Tn * Tn);
            for (auto& g : G) g = rnd();
            for (int i = 0; i < Tn; ++i)
                for (int j = 0; j <= i; ++j) {
                    double a = (i == j) ? Tn * 0.3 : 0.0;
                    for (int k = 0; k < Tn; ++k) a += G[i * Tn + k] * G[j * Tn + k];
                    M[i * Tn + j] = M[j * Tn + i] = a;
                    P[i * (i + 1) / 2 + j] = a;
                }
            for (auto& v : b) v = rnd();
            double *dA, *dout, *drhs; long long* dc;
            cudaMalloc(&dA, P.size() * 8); cudaMalloc(&dout, Tn * 8); cudaMalloc(&drhs, Tn * 8); cudaMalloc(&dc, 16);
            cudaMemcpy(dA, P.data(), P.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(drhs, b.data(), Tn * 8, cudaMemcpyHostToDevice);
            size_t sh = ((size_t)(Tn + 4) * (Tn + 5) / 2 + 2 * Tn + 8) * 8;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
            k<<<1, threads, sh>>>(dA, Tn, dout, drhs, dc, 20);
            cudaError_t e = cudaDeviceSynchronize();
            long long hc[2]; cudaMemcpy(hc, dc, 16, cudaMemcpyDeviceToHost); cudaMemcpy(x.data(), dout, Tn * 8, cudaMemcpyDeviceToHost);
            double res = 0.0;
            for (int i = 0; i < Tn; ++i) { double a = -b[i]; for (int j = 0; j < Tn; ++j) a += M[i * Tn + j] * x[j]; res = fmax(res, fabs(a)); }
            printf("T=%3d threads=%3d  factor %7lld cycles (%5.0f/col)  solve %6lld cycles (%4.0f/step)  |Ax-b|=%.2e  %s\n", Tn, threads, hc[0],
                   (double)hc[0] / Tn, hc[1], (double)hc[1] / (2 * Tn), res, cudaGetErrorString(e));
            cudaFree(dA); cudaFree(dout); cudaFree(drhs); cudaFree(dc);
        }
    return 0;
}
