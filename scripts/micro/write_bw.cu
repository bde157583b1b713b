// Micro-benchmark: achievable HBM write bandwidth on B200 for the store patterns the stash writers use.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_bw write_bw.cu && ./write_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE> __global__ void wr(uint4* dst, size_t n16, int iters_per_thread) {
    // each warp writes 512 contiguous bytes per instruction (like st_chunk_g); grid-stride over 512 B blocks
    size_t warp_global = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    size_t nwarps = (size_t)gridDim.x * (blockDim.x >> 5);
    int lane = threadIdx.x & 31;
    uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
    for (size_t blk = warp_global; blk * 32 < n16; blk += nwarps) {
        uint4* p = dst + blk * 32 + lane;
        if (MODE == 0) *p = v;
        if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        if (MODE == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
}
// bulk S2G: each CTA repeatedly bulk-stores a 64 KB smem tile
__global__ void wr_bulk(uint8_t* dst, size_t bytes, int chunk) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < chunk / 4; i += blockDim.x) ((uint32_t*)sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        for (size_t ofs = (size_t)blockIdx.x * chunk; ofs + chunk <= bytes; ofs += (size_t)gridDim.x * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + ofs), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 10; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 10;
}
int main() {
    size_t bytes = (size_t)2 << 30; uint8_t* d; cudaMalloc(&d, bytes);
    size_t n16 = bytes / 16;
    for (int ctas : {148, 296, 592, 1184}) for (int thr : {128, 256, 1024}) {
        float t0 = timeit([&] { wr<0><<<ctas, thr>>>((uint4*)d, n16, 0); });
        float t1 = timeit([&] { wr<1><<<ctas, thr>>>((uint4*)d, n16, 0); });
        float t2 = timeit([&] { wr<2><<<ctas, thr>>>((uint4*)d, n16, 0); });
        printf("st.v4 ctas=%4d thr=%4d: default %.0f GB/s  .cs %.0f GB/s  no_alloc %.0f GB/s\n", ctas, thr, bytes / t0 / 1e6, bytes / t1 / 1e6, bytes / t2 / 1e6);
    }
    // same stores from a kernel that also holds a big shared-memory carve-out (what is left for L1 shrinks)
    for (int smem_kb : {0, 64, 128, 200, 226}) {
        cudaFuncSetAttribute(wr<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
        float t0 = timeit([&] { wr<0><<<148, 128, smem_kb * 1024>>>((uint4*)d, n16, 0); });
        float t1 = timeit([&] { wr<0><<<148, 256, smem_kb * 1024>>>((uint4*)d, n16, 0); });
        printf("st.v4 with %3d KB dynamic smem: 128 thr %.0f GB/s, 256 thr %.0f GB/s\n", smem_kb, bytes / t0 / 1e6, bytes / t1 / 1e6);
    }
    cudaFuncSetAttribute(wr_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int chunk : {16384, 65536}) {
        float t = timeit([&] { wr_bulk<<<148, 128, 65536>>>(d, bytes, chunk); });
        printf("bulk s2g chunk=%d: %.0f GB/s\n", chunk, bytes / t / 1e6);
    }
    float tm = timeit([&] { cudaMemsetAsync(d, 1, bytes); });
    printf("cudaMemset: %.0f GB/s\n", bytes / tm / 1e6);
    uint8_t* e; cudaMalloc(&e, bytes);
    float tc = timeit([&] { cudaMemcpyAsync(e, d, bytes, cudaMemcpyDeviceToDevice); });
    printf("cudaMemcpy D2D: %.0f GB/s (read+write)\n", 2 * bytes / tc / 1e6);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
