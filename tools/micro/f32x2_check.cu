// Development: are the packed f32x2 operations of sm_100 bit identical per component to the scalar ones -- signed zeros,
// denormals, infinities included -- and what do they cost?   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -O3
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__global__ void check(const float2* a, const float2* b, int n, unsigned long long* bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float2 x = a[i], y = b[i];
    const float2 m2 = __fmul2_rn(x, y), s2 = __fadd2_rn(x, y);
    const float2 mz = __ffma2_rn(x, y, make_float2(-0.f, -0.f));
    const float m0 = __fmul_rn(x.x, y.x), m1 = __fmul_rn(x.y, y.y), s0 = __fadd_rn(x.x, y.x), s1 = __fadd_rn(x.y, y.y);
    const auto ne = [](float p, float q) { return __float_as_uint(p) != __float_as_uint(q) && !(p != p && q != q); };
    if (ne(m2.x, m0) || ne(m2.y, m1)) atomicAdd(&bad[0], 1ull);
    if (ne(s2.x, s0) || ne(s2.y, s1)) atomicAdd(&bad[1], 1ull);
    if (ne(mz.x, m0) || ne(mz.y, m1)) atomicAdd(&bad[2], 1ull);
    if ((ne(m2.x, m0) && m0 == 0.f && m2.x == 0.f) || (ne(m2.y, m1) && m1 == 0.f && m2.y == 0.f)) atomicAdd(&bad[3], 1ull);   // differs only in the sign of zero
}

template <int MODE>
__global__ void rate(float2* out, float h, int iters)
{
    float2 acc[8];
    for (int k = 0; k < 8; k++) acc[k] = make_float2(threadIdx.x * 1e-3f + k, k * 0.5f);
    float2 x = make_float2(1.0001f, 0.9999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (MODE == 0) { acc[k].x = __fadd_rn(acc[k].x, __fmul_rn(x.x, h)); acc[k].y = __fadd_rn(acc[k].y, __fmul_rn(x.y, h)); }
            else if (MODE == 1) acc[k] = __fadd2_rn(acc[k], __fmul2_rn(x, make_float2(h, h)));
            else acc[k] = __fadd2_rn(acc[k], __ffma2_rn(x, make_float2(h, h), make_float2(-0.f, -0.f)));
        }
        x.x = __fadd_rn(x.x, 1e-7f);
    }
    float2 s = make_float2(0.f, 0.f);
    for (int k = 0; k < 8; k++) { s.x += acc[k].x; s.y += acc[k].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    const int n = 1 << 22;
    std::vector<float2> a(n), b(n);
    uint32_t st = 12345u;
    auto rnd = [&]() { st = st * 1664525u + 1013904223u; return st; };
    const float special[] = {0.f, -0.f, 1.f, -1.f, 1e-45f, -1e-45f, 3e-39f, 1e38f, -1e38f, INFINITY, -INFINITY, 0.5f, -0.25f};
    for (int i = 0; i < n; i++) {
        float v[4];
        for (int k = 0; k < 4; k++) {
            const uint32_t r = rnd();
            if ((r & 7) == 0) v[k] = special[(r >> 8) % 13];
            else { uint32_t bits = rnd(); if ((bits & 0x7f800000u) == 0x7f800000u) bits &= 0xbfffffffu; memcpy(&v[k], &bits, 4); }
        }
        a[i] = make_float2(v[0], v[1]); b[i] = make_float2(v[2], v[3]);
    }
    float2 *da, *db; unsigned long long* dbad;
    cudaMalloc(&da, n * sizeof(float2)); cudaMalloc(&db, n * sizeof(float2)); cudaMalloc(&dbad, 4 * sizeof(unsigned long long));
    cudaMemcpy(da, a.data(), n * sizeof(float2), cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * sizeof(float2), cudaMemcpyHostToDevice);
    cudaMemset(dbad, 0, 4 * sizeof(unsigned long long));
    check<<<n / 256, 256>>>(da, db, n, dbad);
    unsigned long long bad[4];
    cudaMemcpy(bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost);
    printf("pairs %d: __fmul2_rn != __fmul_rn: %llu (of which only the sign of a zero: %llu); __fadd2_rn != __fadd_rn: %llu; __ffma2_rn(x, y, -0) != __fmul_rn: %llu\n",
           n, bad[0], bad[3], bad[1], bad[2]);
    float2* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 3; mode++) {
        float ms = 0;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) rate<0><<<148 * 8, 256>>>(out, 0.999f, iters); else if (mode == 1) rate<1><<<148 * 8, 256>>>(out, 0.999f, iters); else rate<2><<<148 * 8, 256>>>(out, 0.999f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double ops = 148.0 * 8 * 256 * (double)iters * 8 * 2 * 2;     // rounded mul + add per component
        printf("mode %d (%s): %.3f ms, %.2f T rounded ops/s\n", mode, mode == 0 ? "scalar fmul + fadd" : mode == 1 ? "fmul2 + fadd2" : "ffma2(x,y,-0) + fadd2", ms, ops / ms / 1e9);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
