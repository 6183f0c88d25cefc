// gather_probe.cu -- what does a random 16-byte (or 8-byte) table read cost in DRAM sectors on B200, per load flavour?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/gather_probe tools/gather_probe.cu
// Run:   build/gather_probe            (times per variant)
//        ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --csv build/gather_probe   (DRAM bytes per variant)
// The lookup kernel of the prefilter reads 13.8 M random 16-byte entries of a 1 GB table per batch; this probe reproduces
// only that access with four independent loads in flight per thread.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int V> __device__ __forceinline__ ulonglong2 ld16(const ulonglong2 *p) {
    ulonglong2 r;
    if (V == 0) asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 1) asm volatile("ld.global.cs.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 2) asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.L2::64B.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 4) asm volatile("ld.global.nc.L2::128B.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 5) asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 6) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol));
    }
    if (V == 7) asm volatile("ld.global.lu.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 8) asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
    if (V == 9) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(r.x), "=l"(r.y) : "l"(p), "l"(pol));
    }
    return r;
}

// n random 16-byte reads of table[entries]; 4 in flight per thread
template <int V>
__global__ void __launch_bounds__(256) gather16(const ulonglong2 *__restrict__ t, uint32_t entries, uint32_t n, uint32_t seed,
                                                unsigned long long *out) {
    unsigned long long acc = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        ulonglong2 e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t j = i + u * stride;
            const uint32_t k = (uint32_t)(((uint64_t)mix(j * 2654435761u + seed) * entries) >> 32);
            e[u] = (j < n) ? ld16<V>(t + k) : make_ulonglong2(0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += e[u].x ^ e[u].y;
    }
    if (acc == 0x1234567ull) out[0] = acc;
}

// the same with 8-byte entries (table half the size)
template <int V>
__global__ void __launch_bounds__(256) gather8(const uint64_t *__restrict__ t, uint32_t entries, uint32_t n, uint32_t seed,
                                               unsigned long long *out) {
    unsigned long long acc = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += 4 * stride) {
        uint64_t e[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t j = i + u * stride;
            const uint32_t k = (uint32_t)(((uint64_t)mix(j * 2654435761u + seed) * entries) >> 32);
            e[u] = 0;
            if (j < n) {
                if (V == 0) asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(e[u]) : "l"(t + k));
                if (V == 1) asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(e[u]) : "l"(t + k));
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += e[u];
    }
    if (acc == 0x1234567ull) out[0] = acc;
}

__global__ void fill(uint64_t *p, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = i * 0x9E3779B97F4A7C15ull;
}

template <class F> static float timeit(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; r++) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main(int argc, char **argv) {
    const uint32_t entries = 64000000u;          // 20^6 k-mers
    const uint32_t n = 13800000u;                // occupied look-ups of one 38 k-frame batch
    uint64_t *t;
    unsigned long long *out;
    CHECK(cudaMalloc(&t, (size_t)entries * 16));
    CHECK(cudaMalloc(&out, 8));
    fill<<<1184, 256>>>(t, (size_t)entries * 2);
    CHECK(cudaDeviceSynchronize());
    const int grid = 148 * 6;
    const char *names[] = {"nc", "cs", "cg", "nc.L2::64B", "nc.L2::128B", "nc.L1::no_allocate", "nc.L1::no_allocate + policy evict_first", "lu", "cv",
                           "nc + policy evict_last"};
    uint32_t seed = 1;
#define RUN(V) { float ms = timeit([&] { gather16<V><<<grid, 256>>>((const ulonglong2 *)t, entries, n, seed++, out); }); \
                 printf("gather16 %-40s %.4f ms  %.1f M reads/ms-> %.0f GB/s at 32 B, %.0f GB/s at 64 B, %.0f GB/s at 128 B per read\n", names[V], ms, n / ms / 1e3, \
                        n * 32.0 / ms / 1e6, n * 64.0 / ms / 1e6, n * 128.0 / ms / 1e6); }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9)
    { float ms = timeit([&] { gather8<0><<<grid, 256>>>(t, entries, n, seed++, out); }); printf("gather8 nc (512 MB table) %.4f ms\n", ms); }
    { float ms = timeit([&] { gather8<1><<<grid, 256>>>(t, entries, n, seed++, out); }); printf("gather8 cs (512 MB table) %.4f ms\n", ms); }
    // compact tables: only the occupied third of the k-mers has an entry
    { float ms = timeit([&] { gather16<0><<<grid, 256>>>((const ulonglong2 *)t, 21700000u, n, seed++, out); }); printf("gather16 nc, 347 MB compact table %.4f ms\n", ms); }
    { float ms = timeit([&] { gather8<0><<<grid, 256>>>(t, 21700000u, n, seed++, out); }); printf("gather8 nc, 174 MB compact table %.4f ms\n", ms); }
    CHECK(cudaDeviceSynchronize());
    return 0;
}
