// K1 (fp32 mode, training) -- the layer GEMMs of the fp32 parity mode on the tensor cores.
//
// The fp32 mode keeps every activation and gradient in HBM as fp32 (field_fp32.cu: concat-free buffers, one GEMM per layer
// and role).  This file runs those GEMMs on tcgen05 without giving up fp32 accuracy: each fp32 operand element is split on
// the fly into NT bf16 terms  x = x0 + x1 (+ x2)  (x0 = bf16(x), x1 = bf16(x - x0), ...; bf16 keeps fp32's exponent, so
// activations, weights and the tiny back-propagated gradients all split without scaling), and the product is the sum of the
// term products with i + j < NT (3 MMAs for NT = 2, 6 for NT = 3), accumulated in fp32 in TMEM: x0 * y0 in one accumulator,
// the cross products in a second one (see the MMA loop for why).  NT = 3 carries 24 significand bits per operand -- the same
// information as the FFMA path.
//
// One CTA = one 128 x 128 output tile; 256 threads.  Per K = 32 chunk every thread loads its share of the fp32 A and B
// tiles from global memory (prefetched one chunk ahead in registers), splits it and writes the bf16 term images into a
// shared-memory stage in the canonical K-major core-matrix layout -- operands stored with the OTHER index contiguous
// (dgrad's weights, both wgrad operands) are transposed by the same pass, so every MMA is K-major (an MN-major MMA costs
// 3-4x more).  One thread issues the MMAs of the chunk and commits them to the stage's barrier; the
// stage is rewritten only after that commit has arrived.  Two CTAs share an SM (96 KB of stages, 256 TMEM columns each), so
// one CTA's conversion overlaps the other's MMAs.  Epilogues as in field_fp32.cu: bias + ReLU; addend + mask; split-K atomics.
#include <cuda_bf16.h>
#include <cstdlib>
#include "nsb_common.cuh"
#include "tc_ptx.cuh"

namespace nsb {
namespace tc {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 32;
constexpr int SG_THREADS = 256;
constexpr int SG_TERM_BYTES = SG_BM * SG_BK * 2;          // one bf16 term image of a 128 x 32 operand chunk: [k/8][row][8]
constexpr uint32_t SG_TMEM_COLS = 256;      // two fp32 accumulators of 128 columns: main (x0 * y0) and cross (every other term product)

template <int NT> struct SgCfg {
    static constexpr int kStageBytes = 2 * NT * SG_TERM_BYTES;            // A terms then B terms
    static constexpr int kStages = NT == 3 ? 2 : 3;                       // 96 KB either way -> two CTAs per SM
    static constexpr int kSmemBytes = kStages * kStageBytes + 128;        // + barriers and the TMEM slot
};

__device__ __forceinline__ uint32_t pack_bf16_pair(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// x[0..8) -> NT bf16 term vectors (16 bytes each); the residual of each step is exact in fp32
template <int NT>
__device__ __forceinline__ void split8(float (&x)[8], uint4 (&out)[NT]) {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        uint32_t p[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p[j] = pack_bf16_pair(x[2 * j], x[2 * j + 1]);
        out[t] = make_uint4(p[0], p[1], p[2], p[3]);
        if (t + 1 < NT) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                x[2 * j] -= __uint_as_float(p[j] << 16);
                x[2 * j + 1] -= __uint_as_float(p[j] & 0xFFFF0000u);
            }
        }
    }
}

// Operand chunk loader: rows [row0, row0 + 128) x contraction [k0, k0 + 32) of P, 16 floats per thread in two tasks.
//   TR = false: P[(row0 + r) * ld + k], k contiguous: task = (row r, group of 8 k) -- two float4 loads;
//   TR = true : P[k * ld + row0 + r], r contiguous:   task = (row r, group of 8 k) -- eight scalar loads, lanes over r.
// Either way a task yields the 8 consecutive k of one row = one 16-byte core-matrix row per term.
template <bool TR>
__device__ __forceinline__ void sg_task(int t, int& r, int& kg) {
    if (!TR) { r = (t & 7) | ((t >> 5) << 3); kg = (t >> 3) & 3; }      // 8 lanes = 8 rows of one k group: conflict-free stores
    else { r = t & 127; kg = t >> 7; }
}
template <bool TR>
__device__ __forceinline__ void sg_load(const float* __restrict__ P, int64_t ld, int64_t row0, int64_t rows, int64_t k0,
                                        int64_t kend, int tid, float (&reg)[2][8]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int r, kg;
        sg_task<TR>(tid + i * SG_THREADS, r, kg);
        const int64_t row = row0 + r, k = k0 + kg * 8;
        if (!TR) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (row < rows && k < kend) {
                const float4* src = reinterpret_cast<const float4*>(P + row * ld + k);
                a = __ldg(src); b = __ldg(src + 1);
            }
            reg[i][0] = a.x; reg[i][1] = a.y; reg[i][2] = a.z; reg[i][3] = a.w;
            reg[i][4] = b.x; reg[i][5] = b.y; reg[i][6] = b.z; reg[i][7] = b.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) reg[i][j] = (row < rows && k + j < kend) ? __ldg(P + (k + j) * ld + row) : 0.f;
        }
    }
}
template <bool TR, int NT>
__device__ __forceinline__ void sg_store(uint8_t* img, int tid, float (&reg)[2][8]) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int r, kg;
        sg_task<TR>(tid + i * SG_THREADS, r, kg);
        uint4 terms[NT];
        split8<NT>(reg[i], terms);
#pragma unroll
        for (int t = 0; t < NT; ++t) *reinterpret_cast<uint4*>(img + t * SG_TERM_BYTES + kg * 2048 + r * 16) = terms[t];
    }
}

template <bool AT, bool BT, int EPI, int NT>
__global__ void __launch_bounds__(SG_THREADS) split_gemm_kernel(const GemmArgs g, int tiles_n, int tiles_mn) {
    using Cfg = SgCfg<NT>;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_empty = sbase + Cfg::kStages * Cfg::kStageBytes;            // one per stage
    const uint32_t bar_done = bar_empty + 8 * Cfg::kStages;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + Cfg::kStages * Cfg::kStageBytes + 64);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    int64_t tile = blockIdx.x;
    int64_t kbeg = 0, kend = g.Kdim;
    if (EPI == EPI_WGRAD) {
        const int64_t split = tile / tiles_mn;
        tile -= split * tiles_mn;
        kbeg = split * g.k_per_split;
        kend = kbeg + g.k_per_split < g.Kdim ? kbeg + g.k_per_split : g.Kdim;
    }
    const int64_t m0 = (tile / tiles_n) * SG_BM, n0 = (tile % tiles_n) * SG_BN;

    if (tid == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) mbar_init(bar_empty + 8 * s, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)),
                     "r"(SG_TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t nchunks = kend > kbeg ? (kend - kbeg + SG_BK - 1) / SG_BK : 0;
    float ra[2][8], rb[2][8];
    if (nchunks > 0) {
        sg_load<AT>(g.A, g.lda, m0, g.Mdim, kbeg, kend, tid, ra);
        sg_load<BT>(g.B, g.ldb, n0, g.Ndim, kbeg, kend, tid, rb);
    }
    // kind::f16, D = f32, A = B = bf16, both K-major, N = 128, M = 128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(SG_BN >> 3) << 17) | ((uint32_t)(SG_BM >> 4) << 24);
    const uint64_t dhi = desc_hi(2048, 128);          // LBO: between the two 8-wide k groups of a K = 16 step; SBO: between 8-row groups
    for (int64_t c = 0; c < nchunks; ++c) {
        const int s = (int)(c % Cfg::kStages);
        const int64_t use = c / Cfg::kStages;
        if (use > 0) mbar_wait(bar_empty + 8 * s, (uint32_t)((use - 1) & 1));      // the MMAs that read this stage have completed
        uint8_t* stage = smem + s * Cfg::kStageBytes;
        sg_store<AT, NT>(stage, tid, ra);
        sg_store<BT, NT>(stage + NT * SG_TERM_BYTES, tid, rb);
        if (c + 1 < nchunks) {
            sg_load<AT>(g.A, g.lda, m0, g.Mdim, kbeg + (c + 1) * SG_BK, kend, tid, ra);
            sg_load<BT>(g.B, g.ldb, n0, g.Ndim, kbeg + (c + 1) * SG_BK, kend, tid, rb);
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a0 = sbase + s * Cfg::kStageBytes, b0 = a0 + NT * SG_TERM_BYTES;
#pragma unroll
            for (int ks = 0; ks < SG_BK / 16; ++ks) {
                // Two accumulators.  tcgen05 adds each MMA into the fp32 accumulator with TRUNCATION (measured: same-sign sums
                // come out low by ~0.25 ulp per MMA, scripts/dbg_split_gemm.py), so the error grows with the number of MMAs that
                // touch an accumulator of full magnitude.  Only the x0 * y0 product (K / 16 MMAs) goes to the main accumulator;
                // the 2 (NT = 2) or 5 (NT = 3) cross products are 2^-8 .. 2^-16 of it and collect in their own accumulator,
                // where their truncation is that much smaller; the epilogue adds the two in registers (round to nearest).
#pragma unroll
                for (int sum = NT - 1; sum >= 1; --sum) {          // smallest products first
#pragma unroll
                    for (int ta = 0; ta <= sum; ++ta) {
                        const int tb = sum - ta;
                        const bool first = c == 0 && ks == 0 && sum == NT - 1 && ta == 0;
                        tc_mma(tmem_base + 128, desc_at(dhi, a0 + ta * SG_TERM_BYTES + ks * 4096), desc_at(dhi, b0 + tb * SG_TERM_BYTES + ks * 4096),
                               idesc, first ? 0u : 1u);
                    }
                }
                tc_mma(tmem_base, desc_at(dhi, a0 + ks * 4096), desc_at(dhi, b0 + ks * 4096), idesc, (c == 0 && ks == 0) ? 0u : 1u);
            }
            tc_commit(bar_empty + 8 * s);
            if (c + 1 == nchunks) tc_commit(bar_done);
        }
    }
    if (nchunks > 0) {
        mbar_wait(bar_done, 0);
        tc_fence_after();
        // ---- epilogue: thread = accumulator row (TMEM lane), warps 0-3 columns 0-63, warps 4-7 columns 64-127
        const int rowl = 32 * (warp & 3) + lane;
        const int64_t m = m0 + rowl;
        const int cb = 64 * (warp >> 2);
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t v[16], vx[16];
            tc_ld16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cb + 16 * cc), v);
            tc_ld16(tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(128 + cb + 16 * cc), vx);
            tc_wait_ld();
            pin16(v); pin16(vx);
            const int64_t n = n0 + cb + 16 * cc;
            if (m < g.Mdim && n < g.Ndim) {
                float f[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + __uint_as_float(vx[j]);
                if (EPI == EPI_FWD) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n) + q);
                        float4 o = make_float4(f[4 * q] + bb.x, f[4 * q + 1] + bb.y, f[4 * q + 2] + bb.z, f[4 * q + 3] + bb.w);
                        if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                        reinterpret_cast<float4*>(g.C + m * g.ldc + n)[q] = o;
                    }
                } else if (EPI == EPI_DGRAD) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 o = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                        if (g.addend) {
                            const float4 ad = reinterpret_cast<const float4*>(g.addend + m * g.ldadd + n)[q];
                            o.x += ad.x; o.y += ad.y; o.z += ad.z; o.w += ad.w;
                        }
                        if (g.mask) {
                            const float4 mk = __ldg(reinterpret_cast<const float4*>(g.mask + m * g.ldm + n) + q);
                            o.x = mk.x > 0.f ? o.x : 0.f; o.y = mk.y > 0.f ? o.y : 0.f;
                            o.z = mk.z > 0.f ? o.z : 0.f; o.w = mk.w > 0.f ? o.w : 0.f;
                        }
                        reinterpret_cast<float4*>(g.C + m * g.ldc + n)[q] = o;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (n + j < g.n_valid) atomicAdd(g.C + m * g.ldc + n + j, f[j]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(SG_TMEM_COLS) : "memory");
    }
}

template <bool AT, bool BT, int EPI, int NT>
static int launch_split_gemm(const GemmArgs& g, cudaStream_t st) {
    using Cfg = SgCfg<NT>;
    auto kern = split_gemm_kernel<AT, BT, EPI, NT>;
    static bool configured = false;          // per instantiation; the attribute is per function, not per device context state we track
    if (!configured) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes) != cudaSuccess) {
            cudaGetLastError();
            return NSB_E_CUDA;
        }
        configured = true;
    }
    const int64_t tiles_m = cdiv(g.Mdim, SG_BM), tiles_n = cdiv(g.Ndim, SG_BN);
    int64_t splits = 1;
    GemmArgs a = g;
    if (EPI == EPI_WGRAD) {
        // split the contraction (points) so that every SM holds two CTAs; k_per_split a multiple of the chunk
        int64_t want = (int64_t)num_sms() * 2 / (tiles_m * tiles_n);
        if (want < 1) want = 1;
        int64_t kps = cdiv(cdiv(g.Kdim, want), SG_BK) * SG_BK;
        if (kps < 8 * SG_BK) kps = 8 * SG_BK;
        splits = cdiv(g.Kdim, kps);
        a.k_per_split = kps;
    }
    const int64_t grid = tiles_m * tiles_n * splits;
    if (grid <= 0) return NSB_OK;
    if (grid > 0x7fffffffLL) return NSB_E_BADARG;
    kern<<<(unsigned)grid, SG_THREADS, Cfg::kSmemBytes, st>>>(a, (int)tiles_n, (int)(tiles_m * tiles_n));
    NSB_LAUNCH_CHECK("split_gemm_kernel");
    return NSB_OK;
}

}  // namespace tc

// number of bf16 terms per operand: NSB_SPLIT_TERMS=2 -> 16 significand bits (3 MMAs), default 3 -> 24 bits (6 MMAs)
static int split_terms() {
    static const int nt = [] { const char* e = getenv("NSB_SPLIT_TERMS"); return (e && e[0] == '2') ? 2 : 3; }();
    return nt;
}

int split_gemm(const GemmArgs& g, int role, cudaStream_t st) {
    // alignment contract of the loaders / epilogues (all layer buffers of field_fp32.cu satisfy it)
    if ((g.lda & 3) || (g.ldb & 3) || (role != EPI_WGRAD && ((g.Kdim & 7) || (g.ldc & 3) || (g.Ndim & 15)))) return NSB_E_BADARG;
    const bool t3 = split_terms() == 3;
    switch (role) {
        case EPI_FWD: return t3 ? tc::launch_split_gemm<false, false, EPI_FWD, 3>(g, st) : tc::launch_split_gemm<false, false, EPI_FWD, 2>(g, st);
        case EPI_DGRAD: return t3 ? tc::launch_split_gemm<false, true, EPI_DGRAD, 3>(g, st) : tc::launch_split_gemm<false, true, EPI_DGRAD, 2>(g, st);
        case EPI_WGRAD: return t3 ? tc::launch_split_gemm<true, true, EPI_WGRAD, 3>(g, st) : tc::launch_split_gemm<true, true, EPI_WGRAD, 2>(g, st);
    }
    return NSB_E_BADARG;
}

}  // namespace nsb

// Test hook (not in include/nsb.h): one GEMM of the given role on caller buffers.
extern "C" int nsb_debug_split_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                                    int64_t N, int64_t K, int role, const float* bias, int relu, const float* mask, int64_t ldm,
                                    const float* addend, int64_t ldadd, int n_valid, void* stream) {
    nsb::GemmArgs g{};
    g.A = A; g.lda = lda; g.B = B; g.ldb = ldb; g.C = C; g.ldc = ldc; g.Mdim = M; g.Ndim = N; g.Kdim = K;
    g.bias = bias; g.relu = relu; g.mask = mask; g.ldm = ldm; g.addend = addend; g.ldadd = ldadd; g.n_valid = n_valid;
    return nsb::split_gemm(g, role, nsb::as_stream(stream));
}
