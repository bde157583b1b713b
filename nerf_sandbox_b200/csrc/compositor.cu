// K3 -- alpha compositor, forward and backward, one warp per ray.
// Restates volume_render_rays (utils/render_utils.py:108-167) and, in the RAW variants, the head
// activations of nerf_forward_pass (:236-246) so that raw MLP logits are consumed directly.
// The exclusive cumprod transmittance is a multiplicative warp scan carried across 32-sample chunks;
// rgb/depth/acc reductions ride the same pass.  The backward recomputes T (never reads it from HBM),
// stages alpha/T per warp in shared memory and walks the ray in reverse with a suffix-sum scan.
#include "nsb_common.cuh"

namespace nsb {

constexpr int kCompWarps = 4;

struct Sample { float r, g, b, sigma; };

// RAW=false: rgb[B,N,3], sigma[B,N].  RAW=true: raw[B*N,4] logits (+ optional noise).
template <bool RAW>
__device__ __forceinline__ Sample load_sample(const float* __restrict__ rgb_or_raw, const float* __restrict__ sigma,
                                              const float* __restrict__ noise, float noise_std, bool add_noise,
                                              uint64_t seed, uint64_t offset, int64_t q, float* pre_out,
                                              const float* pre_in = nullptr, bool softplus = false) {
    Sample s;
    if (RAW) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(rgb_or_raw) + q);
        s.r = 1.0f / (1.0f + expf(-v.x));                     // :236 sigmoid
        s.g = 1.0f / (1.0f + expf(-v.y));
        s.b = 1.0f / (1.0f + expf(-v.z));
        float pre = v.w;
        if (pre_in) pre = *pre_in;                            // reverse pass: noisy pre-activation kept from pass 1
        else if (add_noise) pre += (noise ? noise[q] : hash_normal(seed, offset, (uint64_t)q)) * noise_std;   // :239-241
        if (pre_out) *pre_out = pre;
        // :243-246 relu, or F.softplus (beta = 1, threshold = 20: identity above it)
        s.sigma = softplus ? (pre > 20.0f ? pre : log1pf(expf(pre))) : fmaxf(pre, 0.0f);
    } else {
        s.r = rgb_or_raw[q * 3 + 0]; s.g = rgb_or_raw[q * 3 + 1]; s.b = rgb_or_raw[q * 3 + 2];
        s.sigma = sigma[q];
    }
    return s;
}

__device__ __forceinline__ float delta_at(const float* __restrict__ zrow, int i, int N, bool inf_last, float rn,
                                          bool has_rn) {
    float d = (i == N - 1) ? (inf_last ? 1e10f : 0.0f) : (zrow[i + 1] - zrow[i]);   // :131-136
    if (has_rn) d *= rn;                                                               // :139-141
    return d;
}

// inclusive multiplicative scan across the warp
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v *= t;
    }
    return v;
}

__device__ __forceinline__ float finalize_color(float c) {   // :165 nan_to_num(nan=0,posinf=1,neginf=0).clamp(0,1)
    if (isnan(c)) c = 0.0f;
    return fminf(fmaxf(c, 0.0f), 1.0f);
}

template <bool RAW>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_fwd_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ sigma,
                     const float* __restrict__ noise, float noise_std, const float* __restrict__ z,
                     const float* __restrict__ ray_norm, float* __restrict__ comp, float* __restrict__ weights,
                     float* __restrict__ acc_out, float* __restrict__ depth_out, int64_t B, int N, uint32_t flags,
                     float eps, uint64_t seed, uint64_t offset, const uint64_t* step_dev) {
    if (step_dev) offset += 8 * *step_dev;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool white = flags & NSB_WHITE_BKGD, inf_last = flags & NSB_INFINITE_LAST_BIN;
    const bool add_noise = RAW && (flags & NSB_TRAINING) && noise_std > 0.0f;
    const bool softplus = RAW && (flags & NSB_SIGMA_SOFTPLUS);
    for (int64_t b = blockIdx.x * (int64_t)kCompWarps + warp; b < B; b += (int64_t)gridDim.x * kCompWarps) {
        const float* zrow = z + b * N;
        const bool has_rn = ray_norm != nullptr;
        const float rn = has_rn ? ray_norm[b] : 1.0f;
        float carry = 1.0f, sw = 0.f, swz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const bool valid = i < N;
            float f = 1.0f, alpha = 0.f, zi = 0.f;
            Sample s = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                zi = zrow[i];
                s = load_sample<RAW>(rgb_or_raw, sigma, noise, noise_std, add_noise, seed, offset, b * N + i, nullptr, nullptr, softplus);
                const float sdt = fminf(fmaxf(s.sigma * delta_at(zrow, i, N, inf_last, rn, has_rn), 0.0f), 60.0f);  // :144
                alpha = 1.0f - expf(-sdt);                     // :145
                f = (1.0f - alpha) + eps;                      // :149
            }
            const float incl = warp_scan_mul(f, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;                      // :148-150 exclusive cumprod
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            float w = T * alpha;                               // :153
            if (!isfinite(w)) w = 0.0f;                        // :154
            if (valid) {
                if (weights) weights[b * N + i] = w;
                sw += w; swz += w * zi; sr += w * s.r; sg += w * s.g; sb += w * s.b;
            }
        }
        sw = warp_sum(sw); swz = warp_sum(swz); sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb);
        if (lane == 0) {
            const float acc = fminf(fmaxf(sw, 0.0f), 1.0f);    // :156
            const float bg = white ? (1.0f - acc) : 0.0f;      // :161-162
            comp[b * 3 + 0] = finalize_color(sr + bg);
            comp[b * 3 + 1] = finalize_color(sg + bg);
            comp[b * 3 + 2] = finalize_color(sb + bg);
            if (acc_out) acc_out[b] = acc;
            if (depth_out) depth_out[b] = swz / (acc + eps);   // :157
        }
    }
}

// Backward.  Shared memory per warp: alpha[N], T[N].
template <bool RAW>
__global__ void __launch_bounds__(kCompWarps * 32)
composite_bwd_kernel(const float* __restrict__ rgb_or_raw, const float* __restrict__ sigma,
                     const float* __restrict__ noise, float noise_std, const float* __restrict__ z,
                     const float* __restrict__ ray_norm, const float* __restrict__ g_comp,
                     const float* __restrict__ g_weights, const float* __restrict__ g_acc,
                     const float* __restrict__ g_depth, float* __restrict__ d_rgb_or_raw, float* __restrict__ d_sigma,
                     int64_t B, int N, uint32_t flags, float eps, uint64_t seed, uint64_t offset, const uint64_t* step_dev) {
    if (step_dev) offset += 8 * *step_dev;
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* s_alpha = smem + (size_t)warp * 3 * N;
    float* s_T = s_alpha + N;
    float* s_pre = s_T + N;
    const bool white = flags & NSB_WHITE_BKGD, inf_last = flags & NSB_INFINITE_LAST_BIN;
    const bool add_noise = RAW && (flags & NSB_TRAINING) && noise_std > 0.0f;
    const bool softplus = RAW && (flags & NSB_SIGMA_SOFTPLUS);
    for (int64_t b = blockIdx.x * (int64_t)kCompWarps + warp; b < B; b += (int64_t)gridDim.x * kCompWarps) {
        const float* zrow = z + b * N;
        const bool has_rn = ray_norm != nullptr;
        const float rn = has_rn ? ray_norm[b] : 1.0f;
        // ---- pass 1: forward recompute, totals
        float carry = 1.0f, sw = 0.f, swz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
        for (int base = 0; base < N; base += 32) {
            const int i = base + lane;
            const bool valid = i < N;
            float f = 1.0f, alpha = 0.f, zi = 0.f;
            Sample s = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                zi = zrow[i];
                float pre1 = 0.f;
                s = load_sample<RAW>(rgb_or_raw, sigma, noise, noise_std, add_noise, seed, offset, b * N + i, &pre1, nullptr, softplus);
                if (RAW) s_pre[i] = pre1;
                const float sdt = fminf(fmaxf(s.sigma * delta_at(zrow, i, N, inf_last, rn, has_rn), 0.0f), 60.0f);
                alpha = 1.0f - expf(-sdt);
                f = (1.0f - alpha) + eps;
            }
            const float incl = warp_scan_mul(f, lane);
            float excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 1.0f;
            const float T = carry * excl;
            carry *= __shfl_sync(0xffffffffu, incl, 31);
            float w = T * alpha;
            if (!isfinite(w)) w = 0.0f;
            if (valid) {
                s_alpha[i] = alpha; s_T[i] = T;
                sw += w; swz += w * zi; sr += w * s.r; sg += w * s.g; sb += w * s.b;
            }
        }
        sw = warp_sum(sw); swz = warp_sum(swz); sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb);
        __syncwarp();
        const float acc = fminf(fmaxf(sw, 0.0f), 1.0f);
        const float bg = white ? (1.0f - acc) : 0.0f;
        // clamp/nan_to_num masks on the composite (:165): grad passes on the closed interval, finite only
        float gc[3];
        {
            const float craw[3] = {sr + bg, sg + bg, sb + bg};
#pragma unroll
            for (int c = 0; c < 3; ++c)
                gc[c] = (isfinite(craw[c]) && craw[c] >= 0.0f && craw[c] <= 1.0f) ? g_comp[b * 3 + c] : 0.0f;
        }
        const float inv = 1.0f / (acc + eps);
        const float gd = g_depth ? g_depth[b] : 0.0f;
        float g_accv = g_acc ? g_acc[b] : 0.0f;
        if (white) g_accv -= gc[0] + gc[1] + gc[2];
        g_accv -= gd * swz * inv * inv;                        // depth = swz/(acc+eps)
        const float g_s = (sw >= 0.0f && sw <= 1.0f) ? g_accv : 0.0f;   // clamp(0,1) mask at :156
        // ---- pass 2: reverse walk with suffix sum of G_i * w_i
        float suffix_carry = 0.0f;
        const int nchunks = (N + 31) / 32;
        for (int ch = nchunks - 1; ch >= 0; --ch) {
            const int i = ch * 32 + lane;
            const bool valid = i < N;
            float Gw = 0.f, G = 0.f, T = 0.f, alpha = 0.f, delta = 0.f, pre = 0.f, w = 0.f;
            Sample s = {0.f, 0.f, 0.f, 0.f};
            if (valid) {
                alpha = s_alpha[i]; T = s_T[i];
                s = load_sample<RAW>(rgb_or_raw, sigma, noise, noise_std, add_noise, seed, offset, b * N + i, &pre, RAW ? &s_pre[i] : nullptr, softplus);
                delta = delta_at(zrow, i, N, inf_last, rn, has_rn);
                const float wraw = T * alpha;
                const bool fin = isfinite(wraw);
                w = fin ? wraw : 0.0f;
                G = (g_weights ? g_weights[b * N + i] : 0.0f) + g_s + s.r * gc[0] + s.g * gc[1] + s.b * gc[2] +
                    gd * zrow[i] * inv;
                if (!fin) G = 0.0f;                            // nan_to_num backward
                Gw = G * w;
            }
            // reverse inclusive scan (suffix within the chunk)
            float sfx = Gw;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const float t = __shfl_down_sync(0xffffffffu, sfx, d);
                if (lane + d < 32) sfx += t;
            }
            const float chunk_total = __shfl_sync(0xffffffffu, sfx, 0);
            const float suffix_excl = (sfx - Gw) + suffix_carry;   // sum_{j>i} G_j w_j
            suffix_carry += chunk_total;
            if (valid) {
                const float f = (1.0f - alpha) + eps;
                const float d_alpha = G * T - suffix_excl / f;     // cumprod_backward, division form
                const float d_sdt = d_alpha * (1.0f - alpha);      // d/dsdt (1-exp(-sdt)) = exp(-sdt)
                const float sd = s.sigma * delta;
                float ds = (sd >= 0.0f && sd <= 60.0f) ? d_sdt * delta : 0.0f;   // clamp masks :144
                const int64_t q = b * N + i;
                if (RAW) {
                    float4 o;
                    o.x = w * gc[0] * s.r * (1.0f - s.r);
                    o.y = w * gc[1] * s.g * (1.0f - s.g);
                    o.z = w * gc[2] * s.b * (1.0f - s.b);
                    // relu / softplus backward (softplus: sigmoid(pre), 1 above the threshold)
                    o.w = softplus ? (pre > 20.0f ? ds : ds / (1.0f + expf(-pre))) : (pre > 0.0f ? ds : 0.0f);
                    reinterpret_cast<float4*>(d_rgb_or_raw)[q] = o;
                } else {
                    d_rgb_or_raw[q * 3 + 0] = w * gc[0];
                    d_rgb_or_raw[q * 3 + 1] = w * gc[1];
                    d_rgb_or_raw[q * 3 + 2] = w * gc[2];
                    d_sigma[q] = ds;
                }
            }
        }
        __syncwarp();
    }
}

static int comp_grid(int64_t B) {
    const int64_t want = cdiv(B, kCompWarps);
    const int64_t cap = (int64_t)num_sms() * 16;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

template <bool RAW>
static int launch_bwd(const float* a, const float* sigma, const float* noise, float noise_std, const float* z,
                      const float* rn, const float* g_comp, const float* g_w, const float* g_a, const float* g_d,
                      float* d0, float* d1, int64_t B, int N, uint32_t flags, float eps, uint64_t seed, uint64_t off,
                      void* stream) {
    const size_t smem = (size_t)kCompWarps * 3 * N * sizeof(float);
    if (smem > 200 * 1024) return NSB_E_BADARG;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(composite_bwd_kernel<RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    composite_bwd_kernel<RAW><<<comp_grid(B), kCompWarps * 32, smem, as_stream(stream)>>>(
        a, sigma, noise, noise_std, z, rn, g_comp, g_w, g_a, g_d, d0, d1, B, N, flags, eps, seed, off, g_step_dev);
    NSB_LAUNCH_CHECK("composite_bwd_kernel");
    return NSB_OK;
}

// ---- loss -- train/trainer.py:999-1006 ------------------------------------------------------------
__device__ __forceinline__ float guard01(float x, bool* pass) {   // nan_to_num(nan=0,posinf=1,neginf=0).clamp(0,1)
    *pass = isfinite(x) && x >= 0.0f && x <= 1.0f;
    if (isnan(x)) x = 0.0f;
    return fminf(fmaxf(x, 0.0f), 1.0f);
}

__global__ void mse_zero_kernel(float* scalars) { if (threadIdx.x < 4) scalars[threadIdx.x] = 0.0f; }

__global__ void mse_loss_kernel(const float* __restrict__ comp_c, const float* __restrict__ comp_f,
                                const float* __restrict__ target, float* __restrict__ g_c, float* __restrict__ g_f,
                                float* __restrict__ scalars, int64_t n, float grad_scale) {
    float sc = 0.f, sf = 0.f;
    const float inv_n = 1.0f / (float)n;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        bool pt, pc, pf;
        const float t = guard01(target[i], &pt);
        if (comp_c) {
            const float c = guard01(comp_c[i], &pc);
            const float d = c - t; sc += d * d;
            if (g_c) g_c[i] = pc ? 2.0f * d * inv_n * grad_scale : 0.0f;
        }
        const float f = guard01(comp_f[i], &pf);
        const float d = f - t; sf += d * d;
        if (g_f) g_f[i] = pf ? 2.0f * d * inv_n * grad_scale : 0.0f;
    }
    sc = warp_sum(sc); sf = warp_sum(sf);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&scalars[2], sc * inv_n); atomicAdd(&scalars[3], sf * inv_n); }
}

// small batches (the training step: 3 x 1024 values): one block does zero + reduce + finish, deterministically
__global__ void __launch_bounds__(256) mse_single_block_kernel(const float* __restrict__ comp_c, const float* __restrict__ comp_f,
                                                               const float* __restrict__ target, float* __restrict__ g_c,
                                                               float* __restrict__ g_f, float* __restrict__ scalars, int64_t n,
                                                               float grad_scale) {
    __shared__ float s_c[8], s_f[8];
    float sc = 0.f, sf = 0.f;
    const float inv_n = 1.0f / (float)n;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        bool pt, pc, pf;
        const float t = guard01(target[i], &pt);
        if (comp_c) {
            const float c = guard01(comp_c[i], &pc);
            const float d = c - t; sc += d * d;
            if (g_c) g_c[i] = pc ? 2.0f * d * inv_n * grad_scale : 0.0f;
        }
        const float f = guard01(comp_f[i], &pf);
        const float d = f - t; sf += d * d;
        if (g_f) g_f[i] = pf ? 2.0f * d * inv_n * grad_scale : 0.0f;
    }
    sc = warp_sum(sc); sf = warp_sum(sf);
    if ((threadIdx.x & 31) == 0) { s_c[threadIdx.x >> 5] = sc; s_f[threadIdx.x >> 5] = sf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mc = 0.f, mf = 0.f;
        for (int w = 0; w < 8; ++w) { mc += s_c[w]; mf += s_f[w]; }
        mc *= inv_n; mf *= inv_n;
        scalars[2] = mc; scalars[3] = mf;
        scalars[0] = mc + mf;                                              // :1005
        scalars[1] = -10.0f * log10f(fmaxf(mf, 1e-10f));                   // trainer.py:77-78
    }
}

__global__ void mse_finish_kernel(float* scalars) {
    if (threadIdx.x == 0) {
        scalars[0] = scalars[2] + scalars[3];                              // :1005
        scalars[1] = -10.0f * log10f(fmaxf(scalars[3], 1e-10f));           // trainer.py:77-78
    }
}

}  // namespace nsb

using namespace nsb;

extern "C" int nsb_composite_fwd(const float* rgb, const float* sigma, const float* z, const float* ray_norm,
                                 float* comp, float* weights, float* acc, float* depth, int64_t B, int N,
                                 uint32_t flags, float eps, void* stream) {
    if (B == 0) return NSB_OK;
    if (!rgb || !sigma || !z || !comp || N < 1 || B < 0) return NSB_E_BADARG;
    composite_fwd_kernel<false><<<comp_grid(B), kCompWarps * 32, 0, as_stream(stream)>>>(
        rgb, sigma, nullptr, 0.f, z, ray_norm, comp, weights, acc, depth, B, N, flags, eps, 0, 0, nullptr);
    NSB_LAUNCH_CHECK("composite_fwd_kernel");
    return NSB_OK;
}

extern "C" int nsb_composite_bwd(const float* rgb, const float* sigma, const float* z, const float* ray_norm,
                                 const float* g_comp, const float* g_weights, const float* g_acc, const float* g_depth,
                                 float* d_rgb, float* d_sigma, int64_t B, int N, uint32_t flags, float eps,
                                 void* stream) {
    if (B == 0) return NSB_OK;
    if (!rgb || !sigma || !z || !g_comp || !d_rgb || !d_sigma || N < 1 || B < 0) return NSB_E_BADARG;
    return launch_bwd<false>(rgb, sigma, nullptr, 0.f, z, ray_norm, g_comp, g_weights, g_acc, g_depth, d_rgb, d_sigma,
                             B, N, flags, eps, 0, 0, stream);
}

extern "C" int nsb_composite_raw_fwd(const float* raw, const float* noise, float noise_std, const float* z,
                                     const float* ray_norm, float* comp, float* weights, float* acc, float* depth,
                                     int64_t B, int N, uint32_t flags, uint64_t seed, uint64_t offset, void* stream) {
    if (B == 0) return NSB_OK;
    if (!raw || !z || !comp || N < 1 || B < 0) return NSB_E_BADARG;
    composite_fwd_kernel<true><<<comp_grid(B), kCompWarps * 32, 0, as_stream(stream)>>>(
        raw, nullptr, noise, noise_std, z, ray_norm, comp, weights, acc, depth, B, N, flags, 1e-10f, seed, offset, g_step_dev);
    NSB_LAUNCH_CHECK("composite_raw_fwd_kernel");
    return NSB_OK;
}

extern "C" int nsb_composite_raw_bwd(const float* raw, const float* noise, float noise_std, const float* z,
                                     const float* ray_norm, const float* g_comp, float* d_raw, int64_t B, int N,
                                     uint32_t flags, uint64_t seed, uint64_t offset, void* stream) {
    if (B == 0) return NSB_OK;
    if (!raw || !z || !g_comp || !d_raw || N < 1 || B < 0) return NSB_E_BADARG;
    return launch_bwd<true>(raw, nullptr, noise, noise_std, z, ray_norm, g_comp, nullptr, nullptr, nullptr, d_raw,
                            nullptr, B, N, flags, 1e-10f, seed, offset, stream);
}

extern "C" int nsb_mse_loss(const float* comp_c, const float* comp_f, const float* target, float* g_c, float* g_f,
                            float* scalars, int64_t B, float grad_scale, void* stream) {
    if (!comp_f || !target || !scalars || B < 1) return NSB_E_BADARG;
    const int64_t n = B * 3;
    if (n <= 16384) {
        mse_single_block_kernel<<<1, 256, 0, as_stream(stream)>>>(comp_c, comp_f, target, g_c, g_f, scalars, n, grad_scale);
        NSB_LAUNCH_CHECK("mse_single_block_kernel");
        return NSB_OK;
    }
    mse_zero_kernel<<<1, 32, 0, as_stream(stream)>>>(scalars);
    NSB_LAUNCH_CHECK("mse_zero_kernel");
    const int grid = (int)(cdiv(n, 256) < 296 ? cdiv(n, 256) : 296);
    mse_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>(comp_c, comp_f, target, g_c, g_f, scalars, n, grad_scale);
    NSB_LAUNCH_CHECK("mse_loss_kernel");
    mse_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(scalars);
    NSB_LAUNCH_CHECK("mse_finish_kernel");
    return NSB_OK;
}
