// Micro-benchmark: HBM read bandwidth of cp.async.bulk global->shared rings (one producer/consumer thread per CTA),
// as a function of piece size and ring depth -- the access pattern of the wgrad kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_load_bw bulk_load_bw.cu && ./bulk_load_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// one thread: keeps `slots` loads of `piece` bytes in flight; pieces of a CTA are `stride` apart (tile images far apart)
__global__ void ring(const uint8_t* src, size_t total, int piece, int slots, size_t stride_pieces) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t bars[32];
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const size_t n_pieces = total / piece;
        size_t issued = 0, done = 0;
        size_t idx = blockIdx.x;                    // piece index, grid-strided
        uint32_t parity[32] = {0};
        // prologue
        for (int s = 0; s < slots && idx < n_pieces; ++s, idx += gridDim.x, ++issued) {
            asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bars[s])), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * piece)), "l"(src + idx * (size_t)piece), "r"(piece), "r"(smem_u32(&bars[s])) : "memory");
        }
        int s = 0;
        while (done < issued) {
            while (!try_wait(smem_u32(&bars[s]), parity[s])) {}
            parity[s] ^= 1; ++done;
            if (idx < n_pieces) {
                asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bars[s])), "r"(piece) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + (size_t)s * piece)), "l"(src + idx * (size_t)piece), "r"(piece), "r"(smem_u32(&bars[s])) : "memory");
                idx += gridDim.x; ++issued;
            }
            if (++s == slots) s = 0;
        }
    }
}
int main() {
    size_t bytes = (size_t)2 << 30; uint8_t* d; cudaMalloc(&d, bytes); cudaMemset(d, 1, bytes);
    cudaFuncSetAttribute(ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int cfg[][2] = {{65536, 3}, {32768, 2}, {32768, 3}, {32768, 4}, {32768, 6}, {16384, 4}, {16384, 6}, {16384, 12}, {8192, 8}, {8192, 12}, {8192, 24}, {4096, 24}};
    for (auto& c : cfg) {
        for (int i = 0; i < 2; ++i) ring<<<148, 32, c[0] * c[1]>>>(d, bytes, c[0], c[1], 0);
        cudaEventRecord(a);
        for (int i = 0; i < 5; ++i) ring<<<148, 32, c[0] * c[1]>>>(d, bytes, c[0], c[1], 0);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
        printf("piece %6d B x %2d slots (%3d KB in flight/SM): %.0f GB/s\n", c[0], c[1], c[0] * c[1] / 1024, bytes / ms / 1e6);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
