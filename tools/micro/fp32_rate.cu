// Development microbenchmark: issue rate of rounded FP32 mul/add chains (scalar vs packed f32x2) per SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int MODE>
__global__ void k(float* out, int iters, float y0)
{
    float acc[8], nrm[8], w[8];
    for (int i = 0; i < 8; i++) { acc[i] = 0; nrm[i] = 0; w[i] = threadIdx.x * 0.001f + i; }
    float y = y0;
    if (MODE == 0) {           // scalar: 8 x (mul, add, add) per step
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
#pragma unroll
                for (int i = 0; i < 8; i++) { acc[i] = __fadd_rn(acc[i], __fmul_rn(w[(i + s) & 7], y)); nrm[i] = __fadd_rn(nrm[i], w[(i + s) & 7]); }
                y = __fadd_rn(y, 1e-7f);
            }
        }
    } else if (MODE == 1) {    // scalar fused (for reference): ffma + add
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
#pragma unroll
                for (int i = 0; i < 8; i++) { acc[i] = __fmaf_rn(w[(i + s) & 7], y, acc[i]); nrm[i] = __fadd_rn(nrm[i], w[(i + s) & 7]); }
                y = __fadd_rn(y, 1e-7f);
            }
        }
    } else {                   // packed: 4 x (mul2, add2, add2) per step
        f2 A[4], N[4], W[8];
        for (int i = 0; i < 4; i++) { A[i] = pk2(0, 0); N[i] = pk2(0, 0); }
        for (int i = 0; i < 8; i++) W[i] = pk2(w[i], w[(i + 1) & 7]);
        f2 Y = pk2(y, y);
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int s = 0; s < 8; s++) {
#pragma unroll
                for (int i = 0; i < 4; i++) { A[i] = add2(A[i], mul2(W[(2 * i + s) & 7], Y)); N[i] = add2(N[i], W[(2 * i + s) & 7]); }
                Y = add2(Y, pk2(1e-7f, 1e-7f));
            }
        }
        for (int i = 0; i < 4; i++) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(A[i])); acc[2 * i] = a; acc[2 * i + 1] = b;
                                      asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(N[i])); nrm[2 * i] = a; nrm[2 * i + 1] = b; }
    }
    float r = 0;
    for (int i = 0; i < 8; i++) r += acc[i] + nrm[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm)
{
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    const int iters = 20000;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148 * ctas_per_sm, threads>>>(out, 100, 1.0f);
    cudaEventRecord(a);
    k<MODE><<<148 * ctas_per_sm, threads>>>(out, iters, 1.0f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // mac = (mul, add, add) per candidate-step: 8 steps x 8 candidates per iteration per thread
    const double macs = (double)iters * 64 * threads * ctas_per_sm * 148;
    printf("%-14s threads/CTA %4d CTAs/SM %d : %.3f ms  %.1f G mac/s/SM-clk-equiv: %.2f mac/clk/SM (1.965 GHz)\n", name, threads, ctas_per_sm, ms,
           macs / ms / 1e6, macs / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}

int main()
{
    for (int th : {128, 256, 512, 1024}) { run<0>("scalar mul+add", th, 1); run<1>("scalar ffma", th, 1); run<2>("packed f32x2", th, 1); }
    run<0>("scalar mul+add", 256, 2); run<2>("packed f32x2", 256, 2);
    return 0;
}
