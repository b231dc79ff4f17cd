// Bandwidth probe: how fast can B200 stream a [N, K/2]-byte matrix with (a) fully coalesced reads and
// (b) the GEMV kernel's access pattern (per warp-load: 8 rows x 64 B), without any compute?
// build+run on the GPU box: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/bw_probe.cu -o /tmp/bw_probe && /tmp/bw_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__global__ void coalesced(const uint4* __restrict__ w, size_t n16, uint32_t* out) {
    uint32_t acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < n16) ? ldg_stream_u4(w + i + u * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0x12345678) out[0] = acc;
}

// GEMV pattern: CTA of 4 warps, item = 16-row tile, warp kpart handles super-chunks kpart, kpart+4, ...; lane (g,t): rows g, g+8; 16 B at t*16
template <int DEPTH>
__global__ void gemv_pattern(const uint8_t* __restrict__ w, uint32_t n, uint32_t row_bytes, uint32_t* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const uint32_t tiles = n / 16, chunks = row_bytes / 64;
    uint32_t acc = 0;
    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint8_t* ra = w + (size_t)(tile * 16 + g) * row_bytes + t * 16;
        const uint8_t* rb = w + (size_t)(tile * 16 + g + 8) * row_bytes + t * 16;
        for (uint32_t c0 = warp * 4; c0 < chunks; c0 += 16 * DEPTH) {
            uint4 va[DEPTH][4], vb[DEPTH][4];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t c = c0 + d * 16 + j;
                    if (c < chunks) { va[d][j] = ldg_stream_u4(ra + (size_t)c * 64); vb[d][j] = ldg_stream_u4(rb + (size_t)c * 64); }
                    else { va[d][j] = make_uint4(0,0,0,0); vb[d][j] = make_uint4(0,0,0,0); }
                }
#pragma unroll
            for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc ^= va[d][j].x ^ va[d][j].y ^ va[d][j].z ^ va[d][j].w ^ vb[d][j].x ^ vb[d][j].y ^ vb[d][j].z ^ vb[d][j].w;
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}

int main() {
    const uint32_t shapes[][2] = {{28672, 2048}, {128256, 2048}, {6144, 2048}, {4096, 7168}};
    uint32_t* out; cudaMalloc(&out, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (auto& sh : shapes) {
        uint32_t n = sh[0], rb = sh[1];
        size_t bytes = (size_t)n * rb;
        int copies = (int)(400e6 / bytes) + 2;
        uint8_t* w; cudaMalloc(&w, bytes * copies); cudaMemset(w, 1, bytes * copies);
        auto run = [&](const char* name, auto launch) {
            for (int i = 0; i < 3; ++i) launch(i % copies);
            cudaEventRecord(a);
            const int iters = 20;
            for (int i = 0; i < iters; ++i) launch(i % copies);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("%-28s n=%6u row_bytes=%5u  %8.2f us  %8.1f GB/s\n", name, n, rb, ms * 1000 / iters, bytes / (ms * 1e-3 / iters) / 1e9);
        };
        run("coalesced 148x8x256", [&](int c) { coalesced<<<148 * 8, 256>>>((const uint4*)(w + bytes * c), bytes / 16, out); });
        run("gemv pattern depth1 444", [&](int c) { gemv_pattern<1><<<444, 128>>>(w + bytes * c, n, rb, out); });
        run("gemv pattern depth2 444", [&](int c) { gemv_pattern<2><<<444, 128>>>(w + bytes * c, n, rb, out); });
        run("gemv pattern depth2 592", [&](int c) { gemv_pattern<2><<<592, 128>>>(w + bytes * c, n, rb, out); });
        run("gemv pattern depth2 1184", [&](int c) { gemv_pattern<2><<<1184, 128>>>(w + bytes * c, n, rb, out); });
        cudaFree(w);
    }
    return 0;
}
