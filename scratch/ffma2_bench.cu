// microbenchmark: FFMA vs FFMA2 vs FADD vs FADD2 issue throughput on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, float2 s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, s.x, s.y); a[i].y = fmaf(a[i].y, s.x, s.y); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], s, s); }
            if (MODE == 2) { a[i].x = a[i].x + s.x; a[i].y = a[i].y + s.y; }
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], s); }
        }
    }
    float2 r = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { r.x += a[i].x; r.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name) {
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    k<MODE><<<148 * 8, 256>>>(out, 100, make_float2(0.999f, 0.001f));
    cudaEventRecord(e0);
    k<MODE><<<148 * 8, 256>>>(out, iters, make_float2(0.999f, 0.001f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double lane_ops = 148.0 * 8 * 256 * (double)iters * 16;   // scalar-equivalent ops
    printf("%s: %.3f ms  %.2f T lane-ops/s\n", name, ms, lane_ops / ms / 1e9);
    cudaFree(out);
}
int main() { run<0>("FFMA  "); run<1>("FFMA2 "); run<2>("FADD  "); run<3>("FADD2 "); return 0; }
