// Developer micro-benchmark (B200): latency of dependent DFMA / FFMA / MUFU / F2F chains for one warp,
// and per-SM throughput with many warps.  nvcc -arch=sm_100a -O3 -o lat lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ double stepd(double a, double b, double c) { return fma(a, b, c); }
__global__ void lat_dfma(double* out, int n, long long* cyc) {
    double a = out[threadIdx.x], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); a = fma(a, b, c); }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_ffma(float* out, int n, long long* cyc) {
    float a = out[threadIdx.x], b = 1.0000001f, c = 1e-9f;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = fmaf(a, b, c); a = fmaf(a, b, c); a = fmaf(a, b, c); a = fmaf(a, b, c); }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_f2f(double* out, int n, long long* cyc) {
    double a = out[threadIdx.x];
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { float f = (float)a; a = (double)f + 1.0; f = (float)a; a = (double)f + 1.0; }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_rsq(float* out, int n, long long* cyc) {
    float a = out[threadIdx.x] + 2.0f;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) { a = rsqrtf(a) + 1.0f; a = rsqrtf(a) + 1.0f; a = rsqrtf(a) + 1.0f; a = rsqrtf(a) + 1.0f; }
    long long t1 = clock64();
    out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <typename T, int ILP>
__global__ void thr(T* out, int n) {
    T a[ILP]; for (int k = 0; k < ILP; ++k) a[k] = out[threadIdx.x] + (T)k;
    T b = out[1] + (T)1.0000001, c = out[2] + (T)1e-9;
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < ILP; ++k) a[k] = a[k] * b + c;
    }
    T s = 0; for (int k = 0; k < ILP; ++k) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double* d; float* f; long long* cyc; cudaMalloc(&d, 1 << 24); cudaMalloc(&f, 1 << 24); cudaMalloc(&cyc, 64);
    cudaMemset(d, 0, 1 << 24); cudaMemset(f, 0, 1 << 24);
    long long h; int n = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        lat_dfma<<<1, 32>>>(d, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("DFMA dependent latency   %.2f cycles\n", h / (4.0 * n));
        lat_ffma<<<1, 32>>>(f, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("FFMA dependent latency   %.2f cycles\n", h / (4.0 * n));
        lat_f2f<<<1, 32>>>(d, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("F2F64->32 + F2F32->64 + DADD latency   %.2f cycles per triple\n", h / (2.0 * n));
        lat_rsq<<<1, 32>>>(f, n, cyc); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); if (rep) printf("MUFU.RSQ + FADD latency %.2f cycles\n", h / (4.0 * n));
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int dev = 0, sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    printf("SMs %d clock %d kHz\n", sms, clk);
    for (int warps = 1; warps <= 16; warps *= 2) {
        float ms; int it = 8192;
        for (int which = 0; which < 4; ++which) {
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) thr<double, 8><<<sms * 4, warps * 32 / 4 < 32 ? 32 : warps * 32 / 4>>>(d, it);
                if (which == 1) thr<float, 8><<<sms * 4, warps * 32 / 4 < 32 ? 32 : warps * 32 / 4>>>(f, it);
                if (which == 2) thr<double, 2><<<sms * 4, warps * 32 / 4 < 32 ? 32 : warps * 32 / 4>>>(d, it);
                if (which == 3) thr<float, 16><<<sms * 4, warps * 32 / 4 < 32 ? 32 : warps * 32 / 4>>>(f, it);
                cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
            }
            int ilp = which == 0 ? 8 : which == 1 ? 8 : which == 2 ? 2 : 16;
            int thr_per_block = warps * 32 / 4 < 32 ? 32 : warps * 32 / 4;
            double fl = 2.0 * 4 * ilp * it * (double)sms * 4 * thr_per_block;
            printf("warps/SM %2d %s ILP %2d: %.2f TFLOP/s  (%.1f FMA/clk/SM at nominal clock)\n", (thr_per_block / 32) * 4, (which % 2 == 0) ? "FP64" : "FP32", ilp,
                   fl / ms / 1e9, fl / 2 / (ms * 1e-3) / sms / (clk * 1e3));
        }
    }
    return 0;
}
