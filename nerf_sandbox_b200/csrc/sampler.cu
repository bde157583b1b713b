// K2 -- ray samplers: stratified coarse z, sample_pdf (inverse CDF), fused resample+merge.
// One warp per ray; per-ray CDF/edges staged in shared memory; coalesced row loads/stores.
// Compiled with -fmad=false: every mul/add rounds separately like the reference's ATen ops, which is
// what makes the stratified samples and the searchsorted indices bit-exact.
#include "nsb_common.cuh"

namespace nsb {

constexpr int kWarpsPerBlock = 4;

// ---------------------------------------------------------------------------------------------------
// stratified z -- train/trainer.py:901-908
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float coarse_z_at(int i, int Nc, float near_, float far_) {
    const float t = linspace01(i, Nc);
    return near_ * (1.0f - t) + far_ * t;                      // trainer.py:902
}

__global__ void stratified_kernel(float* __restrict__ z, const float* __restrict__ U, int64_t B, int Nc,
                                  float near_, float far_, int jitter, uint64_t seed, uint64_t offset, const uint64_t* step_dev) {
    if (step_dev) offset += 8 * *step_dev;
    const int64_t total = B * (int64_t)Nc;
    // four consecutive samples per thread: one Philox4x32 block feeds all four uniforms
    for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4; base < total;
         base += (int64_t)gridDim.x * blockDim.x * 4) {
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (jitter && !U) rnd = philox4(seed, offset, (uint64_t)(base >> 2));
        const uint32_t rw[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int64_t idx = base + e;
            if (idx >= total) break;
            const int i = (int)(idx % Nc);
            const float zi = coarse_z_at(i, Nc, near_, far_);
            if (!jitter) { z[idx] = zi; continue; }
            const float zl = i > 0 ? coarse_z_at(i - 1, Nc, near_, far_) : zi;
            const float zr = i < Nc - 1 ? coarse_z_at(i + 1, Nc, near_, far_) : zi;
            const float lower = i > 0 ? 0.5f * (zi + zl) : zi;     // :904-905  (mids = 0.5*(z[1:]+z[:-1]))
            const float upper = i < Nc - 1 ? 0.5f * (zr + zi) : zi; // :906
            const float u = U ? U[idx] : u01(rw[e]);
            z[idx] = lower + (upper - lower) * u;                  // :907 (the sort at :908 is the identity)
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// per-warp PDF machinery shared by sample_pdf and resample_merge
// ---------------------------------------------------------------------------------------------------
// cdf[0..M] from w[0..M-1] held as a callable; inclusive warp scan with carry across 32-wide chunks.
template <typename WFn>
__device__ __forceinline__ void build_cdf(float* cdf, int M, int lane, WFn wfn) {
    float part = 0.f;
    for (int j = lane; j < M; j += 32) part += wfn(j);
    const float total = warp_sum(part);                        // sampling_utils.py:39
    float carry = 0.f;
    if (lane == 0) cdf[0] = 0.f;
    for (int base = 0; base < M; base += 32) {
        const int j = base + lane;
        float v = j < M ? wfn(j) / total : 0.f;                // pdf
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const float t = __shfl_up_sync(0xffffffffu, v, d);
            if (lane >= d) v += t;
        }
        v += carry;                                            // :40 cumsum
        if (j < M) cdf[j + 1] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// searchsorted(cdf, u, right=True): number of entries <= u  (sampling_utils.py:51)
__device__ __forceinline__ int upper_bound(const float* cdf, int n, float u) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ float invert(const float* cdf, const float* edges, int M, float u, int* ind_out) {
    const int ind = upper_bound(cdf, M + 1, u);
    const int below = min(max(ind - 1, 0), M);                 // :52
    const int above = min(max(ind, 1), M);                     // :53
    const float c_lo = cdf[below], c_hi = cdf[above];
    float denom = c_hi - c_lo;
    if (denom < 1e-5f) denom = 1.0f;                           // :62
    const float t = (u - c_lo) / denom;
    const float e_lo = edges[below], e_hi = edges[above];
    *ind_out = ind;
    return e_lo + t * (e_hi - e_lo);                           // :64
}

// edges from midpoints, sampling_utils.py:24-33
template <typename MFn>
__device__ __forceinline__ void build_edges_from_mids(float* edges, int M, int lane, MFn mid) {
    if (M == 1) {
        if (lane == 0) { const float m = mid(0); edges[0] = m - 0.5f * 1e-3f; edges[1] = m + 0.5f * 1e-3f; }
        return;
    }
    for (int j = lane; j <= M; j += 32) {
        float e;
        if (j == 0) e = mid(0) - 0.5f * (mid(1) - mid(0));
        else if (j == M) e = mid(M - 1) + 0.5f * (mid(M - 1) - mid(M - 2));
        else e = 0.5f * (mid(j) + mid(j - 1));
        edges[j] = e;
    }
}

__global__ void sample_pdf_kernel(const float* __restrict__ bins, int bins_cols, const float* __restrict__ weights,
                                  int M, const float* __restrict__ u_in, const float* __restrict__ cdf_in,
                                  float* __restrict__ out, int64_t* __restrict__ inds_out, int64_t B, int n,
                                  int deterministic, uint64_t seed, uint64_t offset) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* edges = smem + (size_t)warp * 2 * (M + 1);
    float* cdf = edges + (M + 1);
    for (int64_t b = blockIdx.x * (int64_t)kWarpsPerBlock + warp; b < B; b += (int64_t)gridDim.x * kWarpsPerBlock) {
        const float* brow = bins + b * bins_cols;
        const float* wrow = weights + b * M;
        if (bins_cols == M + 1) { for (int j = lane; j <= M; j += 32) edges[j] = brow[j]; }
        else build_edges_from_mids(edges, M, lane, [&](int j) { return brow[j]; });
        if (cdf_in) { for (int j = lane; j <= M; j += 32) cdf[j] = cdf_in[b * (M + 1) + j]; }
        else build_cdf(cdf, M, lane, [&](int j) { return fmaxf(wrow[j] + 1e-5f, 0.0f); });   // :38
        __syncwarp();
        for (int s = lane; s < n; s += 32) {
            float u;
            if (deterministic) u = linspace01(s, n);                                            // :44-46
            else u = u_in ? u_in[b * n + s] : philox_uniform(seed, offset, (uint64_t)(b * n + s));
            int ind;
            out[b * n + s] = invert(cdf, edges, M, u, &ind);
            if (inds_out) inds_out[b * n + s] = (int64_t)ind;
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------
// resample + merge -- trainer.py:926-934, :981 (eval: render_utils.py:388-395)
// ---------------------------------------------------------------------------------------------------
// bitonic sort of `len` (power of two) floats held in shared memory by one warp
__device__ __forceinline__ void warp_bitonic_sort(float* a, int len, int lane) {
    for (int k = 2; k <= len; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < len; i += 32) {
                const int p = i ^ j;
                if (p > i) {
                    const float x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncwarp();
        }
    }
}

// first index with a[idx] >= v  /  first index with a[idx] > v, on a sorted shared-memory array
__device__ __forceinline__ int lower_bound_f(const float* a, int n, float v) {
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// The merge of sort(cat(zc, zf)) (trainer.py:981) is done by rank: zc is sorted (stratified), zf is sorted because the
// inverse CDF is monotone in u -- directly for the deterministic linspace u, after a bitonic sort of the Nf samples for
// random u.  out[i + #{zf < zc[i]}] = zc[i];  out[j + #{zc <= zf[j]}] = zf[j]  (coarse first on ties; values-only, so the
// result is bit-identical to any sort).
__global__ void resample_merge_kernel(const float* __restrict__ zc, const float* __restrict__ w_c,
                                      const float* __restrict__ u_in, float* __restrict__ z_all,
                                      float* __restrict__ z_fine, int64_t B, int Nc, int Nf, int sort_len,
                                      int deterministic, uint64_t seed, uint64_t offset, const uint64_t* step_dev) {
    if (step_dev) offset += 8 * *step_dev;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int M = Nc - 1;
    const int Nt = Nc + Nf;
    const int per_warp = 3 * Nc + sort_len + Nt;                // zc | edges | cdf | fine (padded) | merged row
    float* zrow = smem + (size_t)warp * per_warp;
    float* edges = zrow + Nc;
    float* cdf = edges + Nc;
    float* fine = cdf + Nc;
    float* merged = fine + sort_len;
    for (int64_t b = blockIdx.x * (int64_t)kWarpsPerBlock + warp; b < B; b += (int64_t)gridDim.x * kWarpsPerBlock) {
        const float* wrow = w_c + b * Nc;
        for (int j = lane; j < Nc; j += 32) zrow[j] = zc[b * Nc + j];
        __syncwarp();
        auto mid = [&](int j) { return 0.5f * (zrow[j + 1] + zrow[j]); };                      // :926
        build_edges_from_mids(edges, M, lane, mid);
        // weights_bins = 0.5*(w[1:]+w[:-1]) + 1e-5 (:927-928); sample_pdf adds another 1e-5 and clamps (:38)
        build_cdf(cdf, M, lane, [&](int j) { return fmaxf((0.5f * (wrow[j + 1] + wrow[j]) + 1e-5f) + 1e-5f, 0.0f); });
        __syncwarp();
        const bool draw_sorted = !deterministic && !u_in;
        if (draw_sorted) {
            // In-kernel draws: the merge below only needs the SORTED fine samples, and sorted iid uniforms are exactly the
            // normalised partial sums of Nf+1 iid exponentials (order statistics of the uniform distribution), so they are
            // generated in order: u_(k) = (E_1 + .. + E_k) / (E_1 + .. + E_{Nf+1}).  One prefix sum instead of a 28-stage
            // bitonic sort, and one Philox block per four draws.  Same distribution as sort(rand()); the explicit-u path
            // below keeps the draw-then-sort form.
            const int cnt = (Nf + 1 + 31) >> 5;                       // draws per lane, a contiguous block
            const int s0 = lane * cnt;
            // Philox blocks per lane: at least ceil(cnt / 4), so the lanes' counter ranges never overlap (8 up to Nf = 1023)
            const uint64_t ctr_stride = (uint64_t)max(8, (cnt + 3) >> 2);
            float local = 0.f;
            for (int q = 0; q < cnt; q += 4) {
                const uint4 r = philox4(seed, offset, ((uint64_t)b * 32 + lane) * ctr_stride + (uint64_t)(q >> 2));
                const uint32_t rw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int sidx = s0 + q + e;
                    if (q + e < cnt && sidx <= Nf) {
                        const float ex = -__logf(((float)(rw[e] >> 8) + 1.0f) * (1.0f / 16777216.0f));     // Exp(1), argument in (0,1]
                        local += ex;
                        if (sidx < Nf) fine[sidx] = ex;                      // (the last one only enters the total)
                    }
                }
            }
            float incl = local;                                       // inclusive scan of the lane totals
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float tsum = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += tsum;
            }
            const float total = __shfl_sync(0xffffffffu, incl, 31);
            // this lane's partial sums run from the previous lane's inclusive total to its own (clamped: monotone across
            // lane boundaries whatever the rounding of the in-lane additions)
            float run = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) run = 0.f;
            __syncwarp();
            for (int q = 0; q < cnt; ++q) {
                const int sidx = s0 + q;
                if (sidx < Nf) {
                    run = fminf(run + fine[sidx], incl);
                    int ind;
                    const float zf = invert(cdf, edges, M, fminf(run / total, 1.0f), &ind);
                    fine[sidx] = zf;
                    if (z_fine) z_fine[b * Nf + sidx] = zf;
                }
            }
        } else {
            for (int s = lane; s < Nf; s += 32) {
                const float u = deterministic ? linspace01(s, Nf) : u_in[b * Nf + s];
                int ind;
                const float zf = invert(cdf, edges, M, u, &ind);
                fine[s] = zf;
                if (z_fine) z_fine[b * Nf + s] = zf;
            }
        }
        if (!deterministic && !draw_sorted) {
            for (int j = Nf + lane; j < sort_len; j += 32) fine[j] = __int_as_float(0x7f800000);   // +inf pad
            __syncwarp();
            warp_bitonic_sort(fine, sort_len, lane);
        } else {
            __syncwarp();
        }
        for (int i = lane; i < Nc; i += 32) { const float v = zrow[i]; merged[i + lower_bound_f(fine, Nf, v)] = v; }
        for (int j = lane; j < Nf; j += 32) { const float v = fine[j]; merged[j + upper_bound(zrow, Nc, v)] = v; }
        __syncwarp();
        for (int j = lane; j < Nt; j += 32) z_all[b * Nt + j] = merged[j];
        __syncwarp();
    }
}

static int grid_for_rays(int64_t B) {
    const int64_t want = cdiv(B, kWarpsPerBlock);
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace nsb

using namespace nsb;

extern "C" int nsb_stratified_z(float* z, const float* U, int64_t B, int Nc, float near_, float far_, int jitter,
                                uint64_t seed, uint64_t offset, void* stream) {
    if (B == 0) return NSB_OK;
    if (!z || B < 0 || Nc < 1) return NSB_E_BADARG;
    const int64_t total = B * Nc;
    const int64_t want = cdiv(total, 256 * 4);
    const int grid = (int)(want < (int64_t)num_sms() * 8 ? want : (int64_t)num_sms() * 8);
    stratified_kernel<<<grid, 256, 0, as_stream(stream)>>>(z, U, B, Nc, near_, far_, jitter, seed, offset, g_step_dev);
    NSB_LAUNCH_CHECK("stratified_kernel");
    return NSB_OK;
}

extern "C" int nsb_sample_pdf(const float* bins, int bins_cols, const float* weights, int M, const float* u,
                              const float* cdf_in, float* out, int64_t* inds_out, int64_t B, int n, int deterministic,
                              uint64_t seed, uint64_t offset, void* stream) {
    if (B == 0) return NSB_OK;
    if (!bins || !weights || !out || M < 1 || n < 1 || B < 0) return NSB_E_BADARG;
    if (bins_cols != M && bins_cols != M + 1) return NSB_E_BADARG;          // sampling_utils.py:34-35
    const size_t smem = (size_t)kWarpsPerBlock * 2 * (M + 1) * sizeof(float);
    if (smem > 200 * 1024) return NSB_E_BADARG;
    if (smem > 48 * 1024) cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    sample_pdf_kernel<<<grid_for_rays(B), kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
        bins, bins_cols, weights, M, u, cdf_in, out, inds_out, B, n, deterministic, seed, offset);
    NSB_LAUNCH_CHECK("sample_pdf_kernel");
    return NSB_OK;
}

extern "C" int nsb_resample_merge(const float* zc, const float* w_c, const float* u, float* z_all, float* z_fine,
                                  int64_t B, int Nc, int Nf, int deterministic, uint64_t seed, uint64_t offset,
                                  void* stream) {
    if (B == 0) return NSB_OK;
    if (!zc || !w_c || !z_all || Nc < 2 || Nf < 1 || B < 0) return NSB_E_BADARG;
    int sort_len = 32;
    while (sort_len < Nf) sort_len <<= 1;
    const size_t smem = (size_t)kWarpsPerBlock * (3 * Nc + sort_len + Nc + Nf) * sizeof(float);
    if (smem > 200 * 1024) return NSB_E_BADARG;
    if (smem > 48 * 1024) cudaFuncSetAttribute(resample_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    resample_merge_kernel<<<grid_for_rays(B), kWarpsPerBlock * 32, smem, as_stream(stream)>>>(
        zc, w_c, u, z_all, z_fine, B, Nc, Nf, sort_len, deterministic, seed, offset, g_step_dev);
    NSB_LAUNCH_CHECK("resample_merge_kernel");
    return NSB_OK;
}
