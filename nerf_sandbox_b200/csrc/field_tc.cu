// K1 (bf16 tensor-core mode) -- positional encoder + NeRF MLP as ONE persistent tcgen05 kernel.
//
// One CTA per SM walks pairs of 128-point tiles.  Per pair, the whole chain
//     gamma(x) -> mlp.0..7 (skip concat at layer 4) -> feature -> color_fc -> heads
// runs without touching HBM for activations:
//   * A operand (activations, bf16) lives in shared memory in the tcgen05 canonical K-major layout
//     (8x16B core matrices, no swizzle); the epilogue of layer l writes the A operand of layer l+1 in place;
//   * B operand (weights, bf16) is pre-packed by nsb_pack_weights in exactly that shared-memory image, so a
//     K=32 slab is one contiguous block streamed by the TMA engine (cp.async.bulk + mbarrier tx-count)
//     through a 4-stage ring, shared by the two tiles of the pair;
//   * accumulators (fp32) live in TMEM: 2 tiles x 256 columns = all 512 columns;
//   * roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread), warps 4-7 / 8-11 = epilogue
//     groups of tile A / tile B (thread == accumulator row: tcgen05.ld 32x32b, bias + ReLU, bf16 pack,
//     conflict-free st.shared, fence.proxy.async, mbarrier arrive).  While one tile's epilogue runs, the
//     tensor core works on the other tile.
//   * sigma_out (256->1) and color_out (128->3) are CUDA-core dot products inside the epilogues of layer 7
//     and color_fc (fp32), so raw [r,g,b,sigma] leaves the chip as one float4 per point.
// Training adds a bf16 stash of every layer input (bulk-stored tile images) for the backward kernels.
#include <cuda_bf16.h>
#include <cstdio>
#include "nsb_common.cuh"

namespace nsb {
namespace tc {

constexpr int TILE_M = 128;
constexpr int kStageBytes = 16384;              // K=32 x N=256 bf16
constexpr int kStages = 4;
constexpr int kActBytes = TILE_M * 256 * 2;     // 65536
constexpr int kGxBytes = TILE_M * 64 * 2;       // 16384 (gamma(x), later gamma(d))
constexpr int kSmemAct = 0;
constexpr int kSmemGx = 2 * kActBytes;                      // 131072
constexpr int kSmemRing = kSmemGx + 2 * kGxBytes;           // 163840
constexpr int kSmemBar = kSmemRing + kStages * kStageBytes; // 229376
constexpr int kSmemBytes = kSmemBar + 256;
constexpr int kThreads = 384;
constexpr int kNumMmaLayers = 10;               // mlp.0..7, feature, color_fc
constexpr uint32_t kTmemCols = 512;

// bf16 image offsets (bytes) of the 10 MMA layers: [K/8][N][8] bf16 each
__constant__ uint32_t c_layer_ofs[kNumMmaLayers] = {0,      32768,  163840, 294912, 425984,
                                                    589824, 720896, 851968, 983040, 1114112};
constexpr uint32_t kWeightImageBytes = 1114112 + 288 * 128 * 2;   // 1,187,840
// fp32 tail: biases of the 10 MMA layers (9x256 + 128), w_sigma[256], b_sigma, Wo[3][128], bo[3]
constexpr uint32_t kBiasOfs = kWeightImageBytes;                  // floats from here
constexpr int kBiasFloats = 9 * 256 + 128;
constexpr int kWsigOfs = kBiasFloats;                             // float index within the tail
constexpr int kBsigOfs = kWsigOfs + 256;
constexpr int kWoOfs = kBsigOfs + 4;                              // keep 16B alignment
constexpr int kBoOfs = kWoOfs + 384;
constexpr int kTailFloats = kBoOfs + 4;

__host__ __device__ inline int layer_nslabs(int l) { return l == 0 ? 2 : (l == 4 ? 10 : (l == 9 ? 9 : 8)); }
__host__ __device__ inline int layer_N(int l) { return l == 9 ? 128 : 256; }
__host__ __device__ inline int layer_bias_ofs(int l) { return l * 256; }
constexpr int kSlabsPerPair = 2 + 3 * 8 + 10 + 3 * 8 + 8 + 9;   // 77

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a protocol bug traps (kills this context) instead of hanging the GPU
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_ns();
    while (!mbar_try_wait(bar, parity)) {
        if (global_ns() - t0 > 2000000000ull) {   // 2 s
            printf("nsb tc: mbarrier timeout bar=%u parity=%u block=%d thread=%d\n", bar, parity, blockIdx.x, threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm_100):
// core matrix = 8 rows x 16 bytes; SBO = byte stride between 8-row groups; LBO = byte stride between the
// two 8-element K chunks of one K=16 MMA.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128
__device__ __forceinline__ uint32_t make_idesc(int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}

// ---- kernel parameters --------------------------------------------------------------------------------
struct FwdParams {
    const float* rays_o; const float* rays_d; const float* z; const float* ray_norm; const float* viewdirs;   // FROM_ENC=false
    const float* enc_pos; const float* enc_dir;                                                                  // FROM_ENC=true
    const uint8_t* packed;        // bf16 image + fp32 tail
    float* raw;                   // [Q,4]
    uint8_t* stash;               // training: per tile, images of the layer inputs (see StashLayout); else null
    float* dbg; int dbg_layer;    // debug: post-activation fp32 of one layer, [Q,256]
    int64_t Q; int N;             // points, samples per ray
    int64_t num_tiles;
};

// Stash of one tile (bytes), every block a shared-memory image ([K/8][128 rows][8] bf16):
//   gx (K=64) 16384 | h1..h8 (inputs of mlp.1..7 and of feature/sigma; K=256) 8 x 65536 | feat (K=256) 65536 |
//   gd (K=32) 8192 | c (color_fc output, K=128) 32768
constexpr size_t kStashGx = 0, kStashH = 16384, kStashFeat = kStashH + 8 * 65536, kStashGd = kStashFeat + 65536,
                 kStashC = kStashGd + 8192, kStashTile = kStashC + 32768;   // 647,168 B / tile = 5056 B / point

// write 8 consecutive K elements (one 16-byte core-matrix row) of row r at K-chunk `k8`
__device__ __forceinline__ void st_chunk(uint32_t base, int k8, int r, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (uint32_t)k8 * 2048u + (uint32_t)r * 16u), "r"(a), "r"(b),
                 "r"(c), "r"(d)
                 : "memory");
}

// gamma(x) of this thread's point -> gx buffer (K=64: [x(3) | sin(2^k x_d) k-major (30) | cos (30) | 0])
__device__ __forceinline__ void encode_pos(uint32_t gx, int r, float px, float py, float pz) {
    float e[64];
    e[0] = px; e[1] = py; e[2] = pz; e[63] = 0.f;
    float s[3], c[3];
    sincosf(px, &s[0], &c[0]); sincosf(py, &s[1], &c[1]); sincosf(pz, &s[2], &c[2]);
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            e[3 + 3 * k + d] = s[d]; e[33 + 3 * k + d] = c[d];
            const float s2 = 2.0f * s[d] * c[d], c2 = fmaf(-2.0f * s[d], s[d], 1.0f);   // angle doubling
            s[d] = s2; c[d] = c2;
        }
    }
#pragma unroll
    for (int k8 = 0; k8 < 8; ++k8)
        st_chunk(gx, k8, r, pack_bf16(e[8 * k8], e[8 * k8 + 1]), pack_bf16(e[8 * k8 + 2], e[8 * k8 + 3]),
                 pack_bf16(e[8 * k8 + 4], e[8 * k8 + 5]), pack_bf16(e[8 * k8 + 6], e[8 * k8 + 7]));
}
// gamma(d) -> first 4 K-chunks of the gx buffer (K=32: [v(3) | sin (12) | cos (12) | 0 x5])
__device__ __forceinline__ void encode_dir(uint32_t gx, int r, float vx, float vy, float vz) {
    float e[32];
#pragma unroll
    for (int i = 27; i < 32; ++i) e[i] = 0.f;
    e[0] = vx; e[1] = vy; e[2] = vz;
    float s[3], c[3];
    sincosf(vx, &s[0], &c[0]); sincosf(vy, &s[1], &c[1]); sincosf(vz, &s[2], &c[2]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            e[3 + 3 * k + d] = s[d]; e[15 + 3 * k + d] = c[d];
            const float s2 = 2.0f * s[d] * c[d], c2 = fmaf(-2.0f * s[d], s[d], 1.0f);
            s[d] = s2; c[d] = c2;
        }
    }
#pragma unroll
    for (int k8 = 0; k8 < 4; ++k8)
        st_chunk(gx, k8, r, pack_bf16(e[8 * k8], e[8 * k8 + 1]), pack_bf16(e[8 * k8 + 2], e[8 * k8 + 3]),
                 pack_bf16(e[8 * k8 + 4], e[8 * k8 + 5]), pack_bf16(e[8 * k8 + 6], e[8 * k8 + 7]));
}
// materialised encodings (NeRF.forward boundary): copy a row of `n` floats, zero-padded to 8*chunks
__device__ __forceinline__ void copy_enc_row(uint32_t gx, int r, const float* __restrict__ src, int n, int chunks) {
    for (int k8 = 0; k8 < chunks; ++k8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (8 * k8 + j < n && src) ? src[8 * k8 + j] : 0.f;
        st_chunk(gx, k8, r, pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
}

template <bool FROM_ENC>
__global__ void __launch_bounds__(kThreads, 1) field_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers: full[4] empty[4] in_ready[2] acc_full[2]; then the TMEM base address
    const uint32_t bar_full = sbase + kSmemBar, bar_empty = bar_full + 8 * kStages, bar_in = bar_empty + 8 * kStages,
                   bar_acc = bar_in + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + kSmemBar + 8 * (2 * kStages + 4));

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(bar_in + 8 * t, 128); mbar_init(bar_acc + 8 * t, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {   // TMEM allocation (whole SM: 2 tiles x 256 fp32 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t num_pairs = (p.num_tiles + 1) / 2;
    const float* tail = reinterpret_cast<const float*>(p.packed + kBiasOfs);

    if (warp == 0) {
        // ===================================== TMA producer =====================================
        if (lane == 0) {
            uint32_t stage = 0, round = 0;
            for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
                for (int l = 0; l < kNumMmaLayers; ++l) {
                    const uint32_t bytes = 32u * (uint32_t)layer_N(l) * 2u;
                    const uint8_t* src = p.packed + c_layer_ofs[l];
                    const int ns = layer_nslabs(l);
                    for (int s = 0; s < ns; ++s) {
                        mbar_wait(bar_empty + 8 * stage, (round & 1) ^ 1);
                        mbar_expect_tx(bar_full + 8 * stage, bytes);
                        bulk_g2s(sbase + kSmemRing + stage * kStageBytes, src + (size_t)s * bytes, bytes, bar_full + 8 * stage);
                        if (++stage == kStages) { stage = 0; ++round; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================== MMA issuer ========================================
        if (lane == 0) {
            uint32_t stage = 0, round = 0, use = 0;   // use = how many layers both tiles went through (barrier parity)
            for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
                for (int l = 0; l < kNumMmaLayers; ++l, ++use) {
                    const int N = layer_N(l);
                    const uint32_t idesc = make_idesc(N);
                    const uint32_t lbo_b = (uint32_t)N * 16u;
                    const int ns = layer_nslabs(l);
                    for (int s = 0; s < ns; ++s) {
                        // A-operand source of this K=32 slab
                        uint32_t a_off;   // byte offset inside the tile's act / gx buffer
                        bool from_gx;
                        if (l == 0) { from_gx = true; a_off = (uint32_t)s * 4u * 2048u; }
                        else if (l == 4 && s >= 8) { from_gx = true; a_off = (uint32_t)(s - 8) * 4u * 2048u; }
                        else if (l == 9 && s == 8) { from_gx = true; a_off = 0; }
                        else { from_gx = false; a_off = (uint32_t)s * 4u * 2048u; }
                        const uint32_t b_addr = sbase + kSmemRing + stage * kStageBytes;
#pragma unroll
                        for (int t = 0; t < 2; ++t) {
                            if (s == 0) { mbar_wait(bar_in + 8 * t, use & 1); }
                            if (t == 0) { mbar_wait(bar_full + 8 * stage, round & 1); }
                            tc_fence_after();
                            const uint32_t a_addr = sbase + (from_gx ? (kSmemGx + t * kGxBytes) : (kSmemAct + t * kActBytes)) + a_off;
                            const uint32_t d_tmem = tmem_base + (uint32_t)t * 256u;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint64_t adesc = make_desc(a_addr + ks * 2 * 2048, 2048, 128);
                                const uint64_t bdesc = make_desc(b_addr + ks * 2 * lbo_b, lbo_b, 128);
                                tc_mma(d_tmem, adesc, bdesc, idesc, (s > 0 || ks > 0) ? 1u : 0u);
                            }
                            if (s == ns - 1) tc_commit(bar_acc + 8 * t);      // accumulator of tile t complete
                        }
                        tc_commit(bar_empty + 8 * stage);                     // ring slot free once these MMAs retire
                        if (++stage == kStages) { stage = 0; ++round; }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================================== epilogue groups ===================================
        const int t = (warp - 4) >> 2;                 // tile of the pair: 0 = A, 1 = B
        const int r = (int)threadIdx.x - 128 - t * 128; // accumulator row == TMEM lane == point within the tile
        const uint32_t act = sbase + kSmemAct + t * kActBytes, gx = sbase + kSmemGx + t * kGxBytes;
        const uint32_t tmem_row = tmem_base + (uint32_t)t * 256u + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t use = 0;
        for (int64_t pair = blockIdx.x; pair < num_pairs; pair += gridDim.x) {
            const int64_t tile = pair * 2 + t;
            const int64_t q = tile * TILE_M + r;
            const bool valid = tile < p.num_tiles && q < p.Q;
            const int64_t qc = valid ? q : 0;
            // ---- layer-0 input: gamma(x) ----
            float vdir[3] = {0.f, 0.f, 1.f};
            if (FROM_ENC) {
                copy_enc_row(gx, r, valid ? p.enc_pos + qc * kPosDim : nullptr, kPosDim, 8);
            } else {
                const int64_t b = qc / p.N;
                const float zz = valid ? p.z[qc] : 0.f;
                const float zm = p.ray_norm ? zz * p.ray_norm[b] : zz;
                const float px = fmaf(p.rays_d[b * 3 + 0], zm, p.rays_o[b * 3 + 0]);
                const float py = fmaf(p.rays_d[b * 3 + 1], zm, p.rays_o[b * 3 + 1]);
                const float pz = fmaf(p.rays_d[b * 3 + 2], zm, p.rays_o[b * 3 + 2]);
                encode_pos(gx, r, px, py, pz);
                const float* vs = p.viewdirs ? p.viewdirs : p.rays_d;
                const float vx = vs[b * 3 + 0], vy = vs[b * 3 + 1], vz = vs[b * 3 + 2];
                const float inv = 1.0f / fmaxf(sqrtf(vx * vx + vy * vy + vz * vz), 1e-12f);
                vdir[0] = vx * inv; vdir[1] = vy * inv; vdir[2] = vz * inv;
            }
            fence_async_smem();
            mbar_arrive(bar_in + 8 * t);
            float sig = 0.f, rgb[3] = {0.f, 0.f, 0.f};
            for (int l = 0; l < kNumMmaLayers; ++l, ++use) {
                mbar_wait(bar_acc + 8 * t, use & 1);
                tc_fence_after();
                const int N = layer_N(l);
                const float* bias = tail + layer_bias_ofs(l);
                const bool relu = l != 8;
                for (int c0 = 0; c0 < N; c0 += 32) {
                    uint32_t v[32];
                    tc_ld32(tmem_row + (uint32_t)c0, v);
                    tc_wait_ld();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
                        f[j] = __uint_as_float(v[j]) + bb.x; f[j + 1] = __uint_as_float(v[j + 1]) + bb.y;
                        f[j + 2] = __uint_as_float(v[j + 2]) + bb.z; f[j + 3] = __uint_as_float(v[j + 3]) + bb.w;
                    }
                    if (relu) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
                    }
                    if (l == 7) {          // sigma_out on the fp32 activations (mlps.py:265)
                        const float* ws = tail + kWsigOfs + c0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 w4 = __ldg(reinterpret_cast<const float4*>(ws + j));
                            sig = fmaf(f[j], w4.x, sig); sig = fmaf(f[j + 1], w4.y, sig);
                            sig = fmaf(f[j + 2], w4.z, sig); sig = fmaf(f[j + 3], w4.w, sig);
                        }
                    }
                    if (l == 9) {          // color_out on the fp32 color_fc activations (mlps.py:273)
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const float* wo = tail + kWoOfs + ch * 128 + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 w4 = __ldg(reinterpret_cast<const float4*>(wo + j));
                                rgb[ch] = fmaf(f[j], w4.x, rgb[ch]); rgb[ch] = fmaf(f[j + 1], w4.y, rgb[ch]);
                                rgb[ch] = fmaf(f[j + 2], w4.z, rgb[ch]); rgb[ch] = fmaf(f[j + 3], w4.w, rgb[ch]);
                            }
                        }
                    }
                    if (p.dbg && l == p.dbg_layer && valid) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) p.dbg[q * 256 + c0 + j] = f[j];
                    }
                    if (l != 9) {          // next layer's A operand, in place
#pragma unroll
                        for (int j8 = 0; j8 < 4; ++j8)
                            st_chunk(act, (c0 >> 3) + j8, r, pack_bf16(f[8 * j8], f[8 * j8 + 1]), pack_bf16(f[8 * j8 + 2], f[8 * j8 + 3]),
                                     pack_bf16(f[8 * j8 + 4], f[8 * j8 + 5]), pack_bf16(f[8 * j8 + 6], f[8 * j8 + 7]));
                    }
                }
                if (l == 8) {              // gamma(d) for color_fc replaces gamma(x) (layer 4 has retired)
                    if (FROM_ENC) copy_enc_row(gx, r, valid ? p.enc_dir + qc * kDirDim : nullptr, kDirDim, 4);
                    else encode_dir(gx, r, vdir[0], vdir[1], vdir[2]);
                }
                if (l != 9) {
                    tc_fence_before();
                    fence_async_smem();
                    mbar_arrive(bar_in + 8 * t);
                }
            }
            if (valid) {
                reinterpret_cast<float4*>(p.raw)[q] = make_float4(rgb[0] + tail[kBoOfs], rgb[1] + tail[kBoOfs + 1],
                                                                   rgb[2] + tail[kBoOfs + 2], sig + tail[kBsigOfs]);
            }
            tc_fence_before();   // the next pair's first MMA overwrites this accumulator: order our tcgen05.ld before it
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- weight packing: flat fp32 params -> bf16 shared-memory images + fp32 tail ------------------------
__global__ void pack_tc_kernel(const float* __restrict__ params, uint8_t* __restrict__ out) {
    // MMA layer m (0..9) <- parameter layer index: 0..7 trunk, 8 feature, 10 color_fc
    for (int m = blockIdx.y; m < kNumMmaLayers; m += gridDim.y) {
        const int pl = m < 9 ? m : 10;
        const LayerDesc d = layer_desc(pl);
        const int N = d.N, Kp = d.Kpad;
        __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(out + c_layer_ofs[m]);
        for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < N * Kp; idx += gridDim.x * blockDim.x) {
            const int k = idx / N, n = idx % N;                 // image order: [k/8][n][k%8]
            const float w = k < d.K ? params[d.w_off + (int64_t)n * d.K + k] : 0.f;
            img[((size_t)(k >> 3) * N + n) * 8 + (k & 7)] = __float2bfloat16_rn(w);
        }
        float* tail = reinterpret_cast<float*>(out + kBiasOfs);
        for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x)
            tail[layer_bias_ofs(m) + n] = params[d.b_off + n];
    }
    if (blockIdx.y == 0) {
        float* tail = reinterpret_cast<float*>(out + kBiasOfs);
        const LayerDesc ds = layer_desc(9), dc = layer_desc(11);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 256; i += gridDim.x * blockDim.x) tail[kWsigOfs + i] = params[ds.w_off + i];
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 384; i += gridDim.x * blockDim.x) tail[kWoOfs + i] = params[dc.w_off + i];
        if (blockIdx.x == 0 && threadIdx.x < 4) {
            tail[kBsigOfs + threadIdx.x] = threadIdx.x == 0 ? params[ds.b_off] : 0.f;
            tail[kBoOfs + threadIdx.x] = threadIdx.x < 3 ? params[dc.b_off + threadIdx.x] : 0.f;
        }
    }
}

}  // namespace tc

size_t tc_packed_bytes() { return align_up(tc::kWeightImageBytes + tc::kTailFloats * sizeof(float), 256); }

size_t tc_workspace_bytes(int64_t Q, int stash) {
    const int64_t tiles = cdiv(Q, tc::TILE_M);
    return 256 + (stash ? (size_t)tiles * tc::kStashTile : 0);
}

int tc_pack(const float* params, void* packed_bf16, cudaStream_t st) {
    tc::pack_tc_kernel<<<dim3(32, tc::kNumMmaLayers), 256, 0, st>>>(params, reinterpret_cast<uint8_t*>(packed_bf16));
    NSB_LAUNCH_CHECK("pack_tc_kernel");
    return NSB_OK;
}

static int check_arch() {
    static int ok = -1;
    if (ok < 0) {
        int dev = 0, major = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
        ok = major == 10 ? 1 : 0;
    }
    return ok ? NSB_OK : NSB_E_ARCH;
}

template <bool FROM_ENC>
static int launch_fwd(tc::FwdParams& p, cudaStream_t st) {
    NSB_TRY(check_arch());
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(tc::field_fwd_kernel<FROM_ENC>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes) != cudaSuccess)
            return check_launch("cudaFuncSetAttribute(field_fwd_kernel)");
        attr_set = true;
    }
    p.num_tiles = cdiv(p.Q, tc::TILE_M);
    const int64_t pairs = (p.num_tiles + 1) / 2;
    const int grid = (int)(pairs < num_sms() ? pairs : num_sms());
    tc::field_fwd_kernel<FROM_ENC><<<grid, tc::kThreads, tc::kSmemBytes, st>>>(p);
    NSB_LAUNCH_CHECK("field_fwd_kernel");
    return NSB_OK;
}

int tc_field_fwd_rays(const float* rays_o, const float* rays_d, const float* z, const float* ray_norm, const float* viewdirs,
                      const void* packed, float* raw, void* ws, int64_t B, int N, int stash, cudaStream_t st) {
    tc::FwdParams p{};
    p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.ray_norm = ray_norm; p.viewdirs = viewdirs;
    p.packed = reinterpret_cast<const uint8_t*>(packed); p.raw = raw;
    p.stash = stash ? reinterpret_cast<uint8_t*>(ws) + 256 : nullptr;
    p.Q = B * (int64_t)N; p.N = N;
    return launch_fwd<false>(p, st);
}

int tc_field_fwd_enc(const float* enc_pos, const float* enc_dir, const void* packed, float* raw, void* ws, int64_t Q,
                     int stash, cudaStream_t st) {
    tc::FwdParams p{};
    p.enc_pos = enc_pos; p.enc_dir = enc_dir;
    p.packed = reinterpret_cast<const uint8_t*>(packed); p.raw = raw;
    p.stash = stash ? reinterpret_cast<uint8_t*>(ws) + 256 : nullptr;
    p.Q = Q; p.N = 1;
    return launch_fwd<true>(p, st);
}

int tc_field_bwd(const float*, const void*, float*, void*, int64_t, cudaStream_t) { return NSB_E_BADARG; }

// debug hook used by tests: run the forward and dump the fp32 post-activation of one layer
int tc_debug_layer(const float* rays_o, const float* rays_d, const float* z, const float* ray_norm, const float* viewdirs,
                   const void* packed, float* raw, float* dbg, int layer, int64_t B, int N, cudaStream_t st) {
    tc::FwdParams p{};
    p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.ray_norm = ray_norm; p.viewdirs = viewdirs;
    p.packed = reinterpret_cast<const uint8_t*>(packed); p.raw = raw; p.dbg = dbg; p.dbg_layer = layer;
    p.Q = B * (int64_t)N; p.N = N;
    return launch_fwd<false>(p, st);
}

}  // namespace nsb
