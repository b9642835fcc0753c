// Development microbenchmark: inner loop of the pipelined resampler (8 phases x 2 periods per thread, 39 window positions)
// with the warp-uniform taps read from shared memory (two 128-bit loads) or from the constant bank (kernel parameters).
#include <cstdio>
#include <cuda_runtime.h>
struct Taps { float h[20 * 39 * 8]; };
template <int MODE>
__global__ void __launch_bounds__(640, 1) k(float2* out, const float* hq_g, const __grid_constant__ Taps tp, int iters)
{
    extern __shared__ __align__(16) float smem[];
    float* s_hq = smem;                       // 6240 floats
    float2* s_in = reinterpret_cast<float2*>(smem + 6240);   // 64 periods * 147 + 64 frames
    const int tid = threadIdx.x, q = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 6240; i += 640) s_hq[i] = hq_g[i];
    for (int i = tid; i < 64 * 147 + 64; i += 640) s_in[i] = make_float2(i * 1e-4f, -i * 1e-4f);
    __syncthreads();
    float2 a2[2][8];
    for (int p = 0; p < 2; p++) for (int g = 0; g < 8; g++) a2[p][g] = make_float2(0.f, 0.f);
    const float2* x2 = s_in + lane * 147 + q;
    for (int it = 0; it < iters; it++) {
#pragma unroll 3
        for (int m = 0; m < 39; m++) {
            float4 h0, h1;
            if (MODE == 0) {
                h0 = *reinterpret_cast<const float4*>(s_hq + q * 312 + m * 8);
                h1 = *reinterpret_cast<const float4*>(s_hq + q * 312 + m * 8 + 4);
            } else {
                const float* c = tp.h + q * 312 + m * 8;
                h0 = make_float4(c[0], c[1], c[2], c[3]); h1 = make_float4(c[4], c[5], c[6], c[7]);
            }
#pragma unroll
            for (int p = 0; p < 2; p++) {
                const float2 x = x2[m + p * 32 * 147];
                a2[p][0] = __ffma2_rn(x, make_float2(h0.x, h0.x), a2[p][0]);
                a2[p][1] = __ffma2_rn(x, make_float2(h0.y, h0.y), a2[p][1]);
                a2[p][2] = __ffma2_rn(x, make_float2(h0.z, h0.z), a2[p][2]);
                a2[p][3] = __ffma2_rn(x, make_float2(h0.w, h0.w), a2[p][3]);
                a2[p][4] = __ffma2_rn(x, make_float2(h1.x, h1.x), a2[p][4]);
                a2[p][5] = __ffma2_rn(x, make_float2(h1.y, h1.y), a2[p][5]);
                a2[p][6] = __ffma2_rn(x, make_float2(h1.z, h1.z), a2[p][6]);
                a2[p][7] = __ffma2_rn(x, make_float2(h1.w, h1.w), a2[p][7]);
            }
        }
    }
    float2 r = make_float2(0.f, 0.f);
    for (int p = 0; p < 2; p++) for (int g = 0; g < 8; g++) { r.x += a2[p][g].x; r.y += a2[p][g].y; }
    out[blockIdx.x * 640 + tid] = r;
}
template <int MODE> void run(const char* name, float2* out, float* hq, const Taps& tp)
{
    const size_t smem = 6240 * 4 + (64 * 147 + 64) * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, 640, smem>>>(out, hq, tp, 10);
    cudaEventRecord(a);
    const int iters = 2000;
    k<MODE><<<148, 640, smem>>>(out, hq, tp, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double ffma2 = (double)iters * 39 * 16 * 640 * 148;
    printf("%-28s %.3f ms  FFMA2 pipe utilisation %.1f %% (2 cycles each, 4 x 32 lanes per SM)  err %s\n", name, ms,
           100.0 * (ffma2 / 32 * 2) / (ms * 1e-3 * 148 * 1.965e9 * 4), cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    float2* out; cudaMalloc(&out, 148 * 640 * 8);
    float* hq; cudaMalloc(&hq, 6240 * 4);
    static Taps tp; for (int i = 0; i < 6240; i++) tp.h[i] = 1e-3f * (i % 97);
    cudaMemcpy(hq, tp.h, 6240 * 4, cudaMemcpyHostToDevice);
    run<0>("taps from shared memory", out, hq, tp);
    run<1>("taps from the constant bank", out, hq, tp);
    return 0;
}
