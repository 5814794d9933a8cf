// K-DPZ: DPZivkovicAGMMBGS (USTC_BGS type 11): the reference's own Zivkovic adaptive-GMM
// implementation, package_bgs/dp/ZivkovicAGMM.cpp:98-372 (SubtractPixel) driven per frame by
// package_bgs/dp/DPZivkovicAGMMBGS.cpp:32-84.  Same kernel skeleton as K-MOG2 (one pixel's mode list in registers,
// constant indices only), but its own arithmetic: the learning factor uses the OLD weight (k = alpha / weight), the
// weights are renormalised by a division, a new mode renormalises again, the list is sorted by swaps, the variance
// is clamped to [4, 5 * 36], and the plugin's output is the HIGH-threshold mask (2 x threshold); img_bgmodel is never
// written.  State: the MOG2 tile layout (kernels.h), plane q = mode*5 + {0 weight, 1 sigma, 2..4 mean}, K <= 5 modes.
// One thread per pixel: a warp reads each plane of its half tile as one 128-byte row.  fp32 unfused (-fmad=false),
// IEEE division; the double-precision running sum that counts the background modes is kept.
#include "common.cuh"
#include "kernels.h"

namespace bgsb {

struct DpzMode { float w, s, m0, m1, m2; };

__device__ __forceinline__ void dpz_swap(DpzMode &a, DpzMode &b) { DpzMode t = a; a = b; b = t; }

template <int K>
__global__ void __launch_bounds__(256)
dpz_kernel(DpzLaunch L)
{
    pdl_entry();
    const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (unsigned)L.npx) return;
    const size_t s = blockIdx.y;
    float *const q = L.state + s * MOG2_PLANES * L.pstride + mog2_tile_off(p);
    uint8_t *const nmp = L.nmodes + s * L.pstride + p;
    const uint8_t *in = L.frame + s * L.frame_stride + (size_t)p * 3;
    const float p0 = (float)in[0], p1 = (float)in[1], p2 = (float)in[2];
    const float low_thr = L.low_thr, high_thr = 2 * low_thr, alpha = L.alpha;          // DPZivkovicAGMMBGS.cpp:59-61
    const float bg_threshold = 0.75f, variance = 36.0f, complexity_prior = 0.05f;       // ZivkovicAGMM.cpp:64-67
    const float one_min_alpha = 1 - alpha;                                                // :108
    const float prune = -alpha * complexity_prior;                                        // :110

    int n = L.fresh ? 0 : (int)*nmp;
    DpzMode md[K];
#pragma unroll
    for (int m = 0; m < K; m++) {
        if (m < n) {
            md[m].w = q[(m * 5) * MOG2_TILE]; md[m].s = q[(m * 5 + 1) * MOG2_TILE]; md[m].m0 = q[(m * 5 + 2) * MOG2_TILE];
            md[m].m1 = q[(m * 5 + 3) * MOG2_TILE]; md[m].m2 = q[(m * 5 + 4) * MOG2_TILE];
        } else {
            md[m].w = 0.f; md[m].s = 0.f; md[m].m0 = 0.f; md[m].m1 = 0.f; md[m].m2 = 0.f;
        }
    }
    // number of modes that make up the background (:116-128)
    int bg_gauss = 0;
    {
        double sum = 0.0;
        bool stop = false;
#pragma unroll
        for (int m = 0; m < K; m++) {
            if (m < n && !stop) {
                if (sum < (double)bg_threshold) { bg_gauss++; sum += (double)md[m].w; }
                else stop = true;
            }
        }
    }
    bool fits = false, bg_high = false;
    float total = 0.0f;
#pragma unroll
    for (int m = 0; m < K; m++) {
        if (m < n) {                                            // n shrinks when a mode is pruned (:131, :235, :249)
            float weight = md[m].w;
            bool matched = false;
            if (!fits) {
                const float var = md[m].s;
                const float d0 = md[m].m0 - p0, d1 = md[m].m1 - p1, d2 = md[m].m2 - p2;
                const float dist = (d0 * d0 + d1 * d1 + d2 * d2);
                if (dist < high_thr * var && m < bg_gauss) bg_high = true;               // :153-154
                if (dist < low_thr * var) {                                              // :157
                    fits = true; matched = true;
                    const float k = alpha / weight;                                      // :168, the old weight
                    weight = one_min_alpha * weight + prune;
                    weight += alpha;
                    md[m].w = weight;
                    md[m].m0 = md[m].m0 - k * d0; md[m].m1 = md[m].m1 - k * d1; md[m].m2 = md[m].m2 - k * d2;
                    const float sigmanew = var + k * (dist - var);                       // :183
                    md[m].s = sigmanew < 4 ? 4 : sigmanew > 5 * variance ? 5 * variance : sigmanew;   // :186
                    bool moving = true;                                                  // :212-227
#pragma unroll
                    for (int l = K - 1; l > 0; l--) {
                        if (l <= m && moving) {
                            if (md[l].w > md[l - 1].w) dpz_swap(md[l], md[l - 1]);
                            else moving = false;
                        }
                    }
                }
            }
            if (!matched) {                                                              // :231-238, :245-252
                weight = one_min_alpha * weight + prune;
                if (weight < -prune) { weight = 0.0f; n--; }
                md[m].w = weight;
            }
            total += weight;
        }
    }
#pragma unroll
    for (int m = 0; m < K; m++)
        if (m < n) md[m].w = md[m].w / total;                                            // :257-260
    if (!fits) {                                                                         // :263-345
        if (n != K) n++;
        const float wnew = (n == 1) ? 1.f : alpha;
#pragma unroll
        for (int m = 0; m < K; m++)
            if (m == n - 1) { md[m].w = wnew; }
        float s2 = 0.0f;
#pragma unroll
        for (int m = 0; m < K; m++)
            if (m < n) s2 += md[m].w;
        const float inv = 1.0f / s2;
#pragma unroll
        for (int m = 0; m < K; m++)
            if (m < n) md[m].w *= inv;
#pragma unroll
        for (int m = 0; m < K; m++)
            if (m == n - 1) { md[m].m0 = p0; md[m].m1 = p1; md[m].m2 = p2; md[m].s = variance; }
        bool moving = true;
#pragma unroll
        for (int l = K - 1; l > 0; l--) {
            if (l <= n - 1 && moving) {
                if (md[l].w > md[l - 1].w) dpz_swap(md[l], md[l - 1]);
                else moving = false;
            }
        }
    }
#pragma unroll
    for (int m = 0; m < K; m++) {
        if (m < n) {
            q[(m * 5) * MOG2_TILE] = md[m].w; q[(m * 5 + 1) * MOG2_TILE] = md[m].s; q[(m * 5 + 2) * MOG2_TILE] = md[m].m0;
            q[(m * 5 + 3) * MOG2_TILE] = md[m].m1; q[(m * 5 + 4) * MOG2_TILE] = md[m].m2;
        }
    }
    *nmp = (uint8_t)n;
    L.fg[s * L.fg_stride + p] = bg_high ? 0 : 255;                                       // :360-367, Bgs.h:41-42
}

int launch_dpz(const DpzLaunch &L, int nstreams, cudaStream_t stream)
{
    const dim3 grid((unsigned)((L.npx + 255) / 256), (unsigned)nstreams);
    switch (L.K) {
    case 1: launch_pdl(dpz_kernel<1>, grid, dim3(256), 0, stream, L); break;
    case 2: launch_pdl(dpz_kernel<2>, grid, dim3(256), 0, stream, L); break;
    case 3: launch_pdl(dpz_kernel<3>, grid, dim3(256), 0, stream, L); break;
    case 4: launch_pdl(dpz_kernel<4>, grid, dim3(256), 0, stream, L); break;
    case 5: launch_pdl(dpz_kernel<5>, grid, dim3(256), 0, stream, L); break;
    default: set_error("DPZivkovicAGMM: gaussians must be 1..5"); return BGSB_ERR_ARG;
    }
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
}

}  // namespace bgsb
