#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float f0 = threadIdx.x, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3;
    long long i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    int cnt = 0;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b); }
        if (MODE == 1) { x0 = x0 + a; x1 = x1 + a; x2 = x2 + a; x3 = x3 + a; }
        if (MODE == 2) { cnt += (x0 > a) + (x1 > b) + (x2 > a) + (x3 > b); x0 = __longlong_as_double(__double_as_longlong(x0) + i); x1 = __longlong_as_double(__double_as_longlong(x1) + i);  x2 = __longlong_as_double(__double_as_longlong(x2) + i); x3 = __longlong_as_double(__double_as_longlong(x3) + i);}
        if (MODE == 3) { f0 = fmaf(f0, (float)a, (float)b); f1 = fmaf(f1, (float)a, (float)b); f2 = fmaf(f2, (float)a, (float)b); f3 = fmaf(f3, (float)a, (float)b); }
        if (MODE == 4) { cnt += (i0 > (long long)a) + (i1 > (long long)b) + (i2 > (long long)a) + (i3 > (long long)b); i0 += i; i1 += i; i2 += i; i3 += i; }
        if (MODE == 5) { x0 = fmin(x0, a + i); x1 = fmin(x1, b + i); x2 = fmin(x2, a - i); x3 = fmin(x3, b - i); }
        if (MODE == 6) { x0 = (double)f0 - a; x1 = (double)f1 - a; x2 = (double)f2 - a; x3 = (double)f3 - a; f0 += 1.f; f1 += 1.f; f2 += 1.f; f3 += 1.f; cnt += (x0 < x1) + (x2 < x3); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + f0 + f1 + f2 + f3 + i0 + i1 + i2 + i3 + cnt;
}
template <int MODE> void run(const char* name, double* out, int ops_per_iter) {
    int iters = 20000; cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * 4, 256>>>(out, 100, 1.0000001, 0.5);
    cudaEventRecord(a); k<MODE><<<148 * 4, 256>>>(out, iters, 1.0000001, 0.5); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double per_sm_clk = (double)148 * 4 * 256 * iters * ops_per_iter / (ms * 1e-3) / 148 / 1.9e9;
    printf("%-28s %.3f ms  -> %.1f thread-ops/clk/SM\n", name, ms, per_sm_clk);
}
int main() { double* out; cudaMalloc(&out, 148 * 4 * 256 * 8);
    run<0>("DFMA", out, 4); run<1>("DADD", out, 4); run<2>("DSETP (+int add)", out, 4); run<3>("FFMA", out, 4); run<4>("ISETP64 (+add)", out, 4); run<5>("DMNMX fmin (+dadd)", out, 4); run<6>("F2F+DADD x4, DSETP x2", out, 4);
    return 0; }
