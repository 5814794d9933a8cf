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

    // Walk the live modes: decay every weight, find the first mode that explains the pixel (its update is
    // deferred to one block below -- the walk itself stays short), prune.
    float wcmp = 0.f, wfin = 0.f;                      // matched mode: weight before / after the prune test
    float fb = 0.f, fg_ = 0.f, fr = 0.f, fvar = 0.f, fd0 = 0.f, fd1 = 0.f, fd2 = 0.f, fdist2 = 0.f;
    int f = -1;
#pragma unroll
    for (int m = 0; m < K; m++) {
        if (m < n) {                                   // LIVE bound: pruning below shortens the walk
            float wt = a1 * md[m].w + prune;
            bool here = false;
            if (!fits) {
                float var = md[m].v;
                float d0 = md[m].b - x0, d1 = md[m].g - x1, d2 = md[m].r - x2;
                float dist2 = d0 * d0 + d1 * d1 + d2 * d2;
                if (tw < L.TB && dist2 < L.Tb * var) bgflag = true;
                if (dist2 < L.Tg * var) {
                    fits = true; here = true; f = m;
                    wt += aT;
                    wcmp = wt;                         // the sort compares the weight before the prune test
                    fb = md[m].b; fg_ = md[m].g; fr = md[m].r; fvar = var;
                    fd0 = d0; fd1 = d1; fd2 = d2; fdist2 = dist2;
                }
            }
            if (wt < nprune) { wt = 0.f; n--; }
            if (here) wfin = wt;
            md[m].w = wt;
            tw += wt;
        }
    }
    if (fits) {
        Mode upd;
        const float k = aT / wcmp;
        upd.w = wfin;
        upd.b = fb - k * fd0; upd.g = fg_ - k * fd1; upd.r = fr - k * fd2;
        float vn = fvar + k * (fdist2 - fvar);
        vn = fmaxf(vn, L.varMin);
        upd.v = fminf(vn, L.varMax);
        // keep the list sorted by weight: the matched mode moves up past every predecessor whose (already
        // decayed) weight is not larger.  Written as a shift of the predecessors with the mode carried in
        // registers -- no position variable indexes the list, so every index stays a constant and the list
        // stays in registers (a tracked position turns into a dynamically indexed array in local memory).
        // Slots above f are untouched by the sort, so running it after the walk changes nothing.
        bool placed = false;
#pragma unroll
        for (int i = K - 1; i > 0; i--) {
            if (i <= f && !placed) {
                if (!(wcmp < md[i - 1].w)) md[i] = md[i - 1];
                else { md[i] = upd; placed = true; }
            }
        }
        if (!placed) md[0] = upd;
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
        Mode ins;
        ins.v = L.varInit; ins.b = x0; ins.g = x1; ins.r = x2;
        if (n == 1) ins.w = 1.f;
        else {
            ins.w = aT;
#pragma unroll
            for (int i = 0; i < K - 1; i++)
                if (i < n - 1) md[i].w *= a1;
        }
        // the new mode starts in slot n-1 and moves up like a matched one (constant indices only, see above)
        bool placed = false;
#pragma unroll
        for (int i = K - 1; i > 0; i--) {
            if (i <= n - 1 && !placed) {
                if (!(aT < md[i - 1].w)) md[i] = md[i - 1];
                else { md[i] = ins; placed = true; }
            }
        }
        if (!placed) md[0] = ins;
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
