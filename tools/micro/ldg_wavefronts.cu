// micro-benchmark: cost of a divergent 64-byte record fetch as 4 x LDG.128 vs 2 x LDG.256 (one record per lane,
// records scattered over a 32 MB table), with `active` of 32 lanes participating.  Prints ns per warp-fetch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ void ld256(const void *p, float4 &a, float4 &b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
}
template <int WIDE> __global__ void k(const float4 *__restrict__ tab, const int *__restrict__ idx, float *out, int iters, int active, int n) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= active) return;
    int i = idx[t % n];
    float acc = 0;
    for (int it = 0; it < iters; it++) {
        float4 a, b, c, d;
        const float4 *p = tab + 4 * size_t(i);
        if (WIDE) { ld256(p, a, b); ld256(p + 2, c, d); }
        else { a = __ldg(p); b = __ldg(p + 1); c = __ldg(p + 2); d = __ldg(p + 3); }
        acc += a.x + b.y + c.z + d.w;
        i = (__float_as_int(a.w) & 0x7fffffff) % n;  // dependent chase, like a BVH child reference
    }
    out[t] = acc;
}
int main() {
    const int n = 1 << 19;  // 512k records x 64 B = 32 MB (L2 resident)
    std::vector<float4> h(4 * size_t(n));
    srand(1);
    for (int i = 0; i < n; i++) for (int q = 0; q < 4; q++) { int r = rand() % n; float f; memcpy(&f, &r, 4); h[4 * size_t(i) + q] = make_float4(1, 2, 3, f); }
    std::vector<int> hi(n); for (int i = 0; i < n; i++) hi[i] = rand() % n;
    float4 *d; int *di; float *o;
    cudaMalloc(&d, h.size() * 16); cudaMalloc(&di, n * 4); cudaMalloc(&o, 148 * 12 * 128 * 4);
    cudaMemcpy(d, h.data(), h.size() * 16, cudaMemcpyHostToDevice); cudaMemcpy(di, hi.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int active : {32, 16, 8, 4})
        for (int wide = 0; wide < 2; wide++) {
            const int iters = 2000, grid = 148 * 12;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (wide) k<1><<<grid, 128>>>(d, di, o, iters, active, n); else k<0><<<grid, 128>>>(d, di, o, iters, active, n);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fetches = double(grid) * 4 * iters;  // warp-level record fetches
            printf("active=%2d %s: %.3f ms, %.2f ns per warp-fetch per SM-slot, %.1f G lane-records/s\n", active, wide ? "2xLDG.256" : "4xLDG.128", ms,
                   ms * 1e6 / fetches * 148, fetches * active / ms / 1e6);
        }
    return 0;
}
