// Generic per-pixel MOG2 update (one pixel, one frame, every case of the algorithm), shared by the
// straight kernel (mog2.cu) and by the generic phase of the production kernels (mog2_t1.cu).
// Restates cv::BackgroundSubtractorMOG2's MOG2Invoker loop body + detectShadowGMM +
// getBackgroundImage (bgfg_gaussmix2.cpp; reference call sites package_bgs/MixtureOfGaussianV2BGS.cpp:56,59),
// spec SURVEY.md Appendix A.4.  All register indices are compile-time (fully unrolled).
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace bgsb {

constexpr int K = MOG2_K;

struct Mode { float w, v, b, g, r; };

// One pixel, one frame.  `md` is the per-pixel mode list (registers), n its live length.
// Returns the raw MOG2 mask value {0, shadow, 255} and the background colour of the updated model.
template <bool SHADOWS>
__device__ __forceinline__ unsigned mog2_pixel(Mode (&md)[K], int &n, float x0, float x1, float x2,
                                               float aT, float a1, float prune, const Mog2Launch &L,
                                               unsigned &bgB, unsigned &bgG, unsigned &bgR, bool want_bg)
{
    bool bgflag = false, fits = false;
    float tw = 0.f;
    const float nprune = -prune;

#pragma unroll
    for (int m = 0; m < K; m++) {
        if (m < n) {                                   // LIVE bound: pruning below shortens the walk
            float wt = a1 * md[m].w + prune;
            int pos = m;
            if (!fits) {
                float var = md[m].v;
                float d0 = md[m].b - x0, d1 = md[m].g - x1, d2 = md[m].r - x2;
                float dist2 = d0 * d0 + d1 * d1 + d2 * d2;
                if (tw < L.TB && dist2 < L.Tb * var) bgflag = true;
                if (dist2 < L.Tg * var) {
                    fits = true;
                    wt += aT;
                    float k = aT / wt;
                    float nb = md[m].b - k * d0, ng = md[m].g - k * d1, nr = md[m].r - k * d2;
                    float vn = var + k * (dist2 - var);
                    vn = fmaxf(vn, L.varMin);
                    vn = fminf(vn, L.varMax);
                    // keep the list sorted by weight: bubble the matched mode up past every
                    // predecessor whose (already updated) weight is not larger
#pragma unroll
                    for (int i = m; i > 0; i--) {
                        if (pos == i && !(wt < md[i - 1].w)) { md[i] = md[i - 1]; pos = i - 1; }
                    }
#pragma unroll
                    for (int i = 0; i <= m; i++)
                        if (pos == i) { md[i].v = vn; md[i].b = nb; md[i].g = ng; md[i].r = nr; }
                }
            }
            if (wt < nprune) { wt = 0.f; n--; }
#pragma unroll
            for (int i = 0; i <= m; i++)
                if (pos == i) md[i].w = wt;
            tw += wt;
        }
    }

    // renormalise
    float inv = 0.f;
    if (fabsf(tw) > 1.1920929e-07f) inv = 1.f / tw;
#pragma unroll
    for (int m = 0; m < K; m++)
        if (m < n) md[m].w *= inv;

    // no mode explains the pixel: insert a new one (replace the weakest if the list is full)
    if (!fits && aT > 0.f) {
        if (n < K) n++;
        int pos = n - 1;
        float wn;
        if (n == 1) wn = 1.f;
        else {
            wn = aT;
#pragma unroll
            for (int i = 0; i < K - 1; i++)
                if (i < n - 1) md[i].w *= a1;
        }
#pragma unroll
        for (int i = K - 1; i > 0; i--) {
            if (pos == i && !(aT < md[i - 1].w)) { md[i] = md[i - 1]; pos = i - 1; }
        }
#pragma unroll
        for (int i = 0; i < K; i++)
            if (pos == i) { md[i].w = wn; md[i].v = L.varInit; md[i].b = x0; md[i].g = x1; md[i].r = x2; }
    }

    // classification
    unsigned raw = 0;
    if (!bgflag) {
        raw = 255;
        if (SHADOWS) {
            // detectShadowGMM: walk the background modes, test brightness ratio + chroma distortion
            float tW = 0.f;
            bool done = false;
#pragma unroll
            for (int m = 0; m < K; m++) {
                if (!done && m < n) {
                    float num = x0 * md[m].b + x1 * md[m].g + x2 * md[m].r;
                    float den = md[m].b * md[m].b + md[m].g * md[m].g + md[m].r * md[m].r;
                    if (den == 0.f) { done = true; }
                    else {
                        if (num <= den && num >= L.tau * den) {
                            float a = num / den;
                            float e0 = a * md[m].b - x0, e1 = a * md[m].g - x1, e2 = a * md[m].r - x2;
                            float e = e0 * e0 + e1 * e1 + e2 * e2;
                            if (e < L.Tb * md[m].v * a * a) { raw = (unsigned)L.shadow_value; done = true; }
                        }
                        if (!done) { tW += md[m].w; if (tW > L.TB) done = true; }
                    }
                }
            }
        }
    }

    // getBackgroundImage on the updated model
    if (want_bg) {
        float aB = 0.f, aG = 0.f, aR = 0.f, t2 = 0.f;
        bool stop = false;
#pragma unroll
        for (int m = 0; m < K; m++) {
            if (!stop && m < n) {
                float w = md[m].w;
                aB += w * md[m].b; aG += w * md[m].g; aR += w * md[m].r;
                t2 += w;
                if (t2 > L.TB) stop = true;
            }
        }
        float iv = 0.f;
        if (fabsf(t2) > 1.1920929e-07f) iv = 1.f / t2;
        bgB = sat_u8_rint(aB * iv); bgG = sat_u8_rint(aG * iv); bgR = sat_u8_rint(aR * iv);
    }
    return raw;
}


}  // namespace bgsb
