// Micro-benchmark: per-lane gather of 128-byte records from an L2-resident array with a BVH-like skew (half of the
// accesses go to a 64 KB hot set, like the top levels of a tree), as 8 x LDG.128 vs 4 x 256-bit loads of several flavours.
// Questions: (1) does the L1TEX data pipe charge per lane and instruction (so 256-bit loads relieve it)?
//            (2) which 256-bit load flavours allocate in L1?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
struct alignas(32) U8 { uint32_t v[8]; };
#define LD256(NAME, PTX)                                                                                                   \
    __device__ __forceinline__ U8 NAME(const void *p) {                                                                    \
        U8 r;                                                                                                              \
        asm volatile(PTX " {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                                               \
                     : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p)); \
        return r;                                                                                                          \
    }
LD256(ld_nc, "ld.global.nc.v8.u32")
LD256(ld_plain, "ld.global.v8.u32")
LD256(ld_ca, "ld.global.ca.v8.u32")
LD256(ld_nc_el, "ld.global.nc.L1::evict_last.v8.u32")
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
template <int MODE>
__global__ void __launch_bounds__(128) k(const uint4 *nodes, uint32_t n, uint32_t *out, int iters, uint32_t hot) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x, acc = 0;
    for (int it = 0; it < iters; it++) {
        s = hash(s);
        const uint32_t i = (s & 0x10000u) ? (s >> 17) % hot : s % n;
        const uint4 *p = nodes + (size_t)i * 8;
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) { uint4 v = __ldg(p + j); acc += v.x ^ v.y ^ v.z ^ v.w; }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
                U8 v = MODE == 1 ? ld_nc(p + 2 * j) : MODE == 2 ? ld_plain(p + 2 * j) : MODE == 3 ? ld_ca(p + 2 * j) : ld_nc_el(p + 2 * j);
                acc += v.v[0] ^ v.v[1] ^ v.v[2] ^ v.v[3] ^ v.v[4] ^ v.v[5] ^ v.v[6] ^ v.v[7];
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main() {
    const uint32_t n = 173873; const int iters = 256;
    uint4 *nodes; uint32_t *out;
    cudaMalloc(&nodes, (size_t)n * 128); cudaMemset(nodes, 1, (size_t)n * 128);
    const int grid = 148 * 6, block = 128;
    cudaMalloc(&out, grid * block * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[5] = {"8 x LDG.128 (__ldg)", "4 x ld.global.nc.v8", "4 x ld.global.v8", "4 x ld.global.ca.v8", "4 x ld.global.nc.L1::evict_last.v8"};
    for (uint32_t hot : {512u, 173873u}) {
        for (int mode = 0; mode < 5; mode++) {
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                switch (mode) {
                    case 0: k<0><<<grid, block>>>(nodes, n, out, iters, hot); break;
                    case 1: k<1><<<grid, block>>>(nodes, n, out, iters, hot); break;
                    case 2: k<2><<<grid, block>>>(nodes, n, out, iters, hot); break;
                    case 3: k<3><<<grid, block>>>(nodes, n, out, iters, hot); break;
                    default: k<4><<<grid, block>>>(nodes, n, out, iters, hot); break;
                }
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            double recs = (double)grid * block * iters;
            printf("hot set %6u records | %-36s: %.3f ms, %6.2f G records/s, %7.1f GB/s logical\n", hot, names[mode], best, recs / best / 1e6, recs * 128 / best / 1e6);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
