// K-DPS: the three simple per-pixel models of the reference's DP package (sibling plugins, SURVEY 8f N3 widened):
//
//   DPAdaptiveMedianBGS  USTC_BGS type 9   package_bgs/dp/AdaptiveMedianBGS.cpp:53-140   (wrapper DPAdaptiveMedianBGS.cpp:28-82)
//   DPMeanBGS            USTC_BGS type 12  package_bgs/dp/MeanBGS.cpp:32-131            (wrapper DPMeanBGS.cpp:28-84)
//   DPWrenGABGS          USTC_BGS type 13  package_bgs/dp/WrenGA.cpp:47-173             (wrapper DPWrenGABGS.cpp:28-84)
//
// Every wrapper does, per frame: Subtract (mask from the CURRENT model; the plugin's output is the HIGH-threshold mask),
// clear the low mask, Update on every pixel (the cleared mask makes the learning-frames test vacuous).  img_bgmodel is
// never written.  The reference makes two full passes with per-pixel accessor calls; here one thread owns 4 pixels
// (12 frame bytes = three words), reads the model once, writes mask and model once:
//   AdaptiveMedian  3 in + 3 model + 1 mask (+ 3 model on the frames whose number is 1 modulo samplingRate)  =  7 / 10 B/px
//   Mean            3 in + 12 mean read + 12 written + 1 mask                                                = 28 B/px
//   WrenGA          3 in + 16 (mu, var) read + 16 written + 1 mask                                           = 36 B/px
// fp32 arithmetic in the reference's order, unfused (the library is compiled with -fmad=false; the operations are
// written with the _rn intrinsics anyway).  Parity: bit-exact against the CPU restatement, which is pinned to a build of
// the reference's own sources (tests/golden/golden_dp.json, tests/test_oracle_pin.py).
#include "common.cuh"
#include "kernels.h"

namespace bgsb {

// the 12 bytes of 4 pixels as three words (B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3); `n` valid pixels, zero fill
__device__ __forceinline__ void dps_load12(const uint8_t *p, bool vec, int n, unsigned (&w)[3])
{
    if (vec) {
        const unsigned *q = reinterpret_cast<const unsigned *>(p);
        w[0] = ld_stream_u32(q); w[1] = ld_stream_u32(q + 1); w[2] = ld_stream_u32(q + 2);
    } else {
        w[0] = w[1] = w[2] = 0u;
        for (int i = 0; i < 3 * n; i++) w[i >> 2] |= (unsigned)p[i] << (8 * (i & 3));
    }
}
__device__ __forceinline__ void dps_store12(uint8_t *p, bool vec, int n, const unsigned (&w)[3])
{
    if (vec) {
        unsigned *q = reinterpret_cast<unsigned *>(p);
        st_stream_u32(q, w[0]); st_stream_u32(q + 1, w[1]); st_stream_u32(q + 2, w[2]);
    } else {
        for (int i = 0; i < 3 * n; i++) p[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    }
}
__device__ __forceinline__ void dps_store_mask(uint8_t *p, bool vec, int n, unsigned m)
{
    if (vec) st_stream_u32(p, m);
    else for (int i = 0; i < n; i++) p[i] = (uint8_t)(m >> (8 * i));
}
__device__ __forceinline__ unsigned dps_byte(const unsigned (&w)[3], int i) { return (w[i >> 2] >> (8 * (i & 3))) & 0xffu; }

__global__ void __launch_bounds__(256)
dps_median_kernel(DpsLaunch L)
{
    pdl_entry();
    const long long px0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (px0 >= L.npx) return;
    const int n = (int)min(4LL, (long long)L.npx - px0);
    const size_t s = blockIdx.y;
    const uint8_t *fp = L.frame + s * L.frame_stride + px0 * 3;
    uint8_t *mp = L.median + s * L.median_stride + px0 * 3;
    uint8_t *gp = L.fg + s * L.fg_stride + px0;
    const bool v12 = n == 4 && ((reinterpret_cast<uintptr_t>(fp) | reinterpret_cast<uintptr_t>(mp)) & 3) == 0;
    const bool v4 = n == 4 && (reinterpret_cast<uintptr_t>(gp) & 3) == 0;
    unsigned in[3], med[3];
    dps_load12(fp, v12, n, in);
    if (L.fresh) { med[0] = in[0]; med[1] = in[1]; med[2] = in[2]; }           // InitModel :53-63
    else dps_load12(mp, v12, n, med);
    // SubtractPixel :92-111: foreground unless all three |pixel - median| <= high
    const unsigned hi = L.high_u * 0x01010101u;
    unsigned gt[3];
#pragma unroll
    for (int k = 0; k < 3; k++) gt[k] = __vcmpgtu4(__vabsdiffu4(in[k], med[k]), hi) & 0x01010101u;
    const unsigned p0 = gt[0] & 0x00ffffffu, p1 = __byte_perm(gt[0], gt[1], 0x0543u) & 0x00ffffffu;
    const unsigned p2 = __byte_perm(gt[1], gt[2], 0x0432u) & 0x00ffffffu, p3 = gt[2] >> 8;
    const unsigned m = (p0 ? 0xffu : 0u) | (p1 ? 0xff00u : 0u) | (p2 ? 0xff0000u : 0u) | (p3 ? 0xff000000u : 0u);
    dps_store_mask(gp, v4, n, m);
    // Update :65-90 (every pixel: the wrapper clears the update mask): one step towards the pixel, per channel
    if (L.update) {
#pragma unroll
        for (int k = 0; k < 3; k++)
            med[k] = med[k] + (__vcmpgtu4(in[k], med[k]) & 0x01010101u) - (__vcmpltu4(in[k], med[k]) & 0x01010101u);   // no byte carries
    }
    if (L.update || L.fresh) dps_store12(mp, v12, n, med);
}

// one pixel of MeanBGS / WrenGA: x = its three bytes as floats, m = its model (Mean: 3 means; WrenGA: 3 means + variance)
template <int KIND>
__device__ __forceinline__ unsigned dps_float_pixel(const DpsLaunch &L, const float (&x)[3], float (&m)[4])
{
    if (L.fresh) {                                       // InitModel (MeanBGS.cpp:40-52, WrenGA.cpp:67-85)
        m[0] = x[0]; m[1] = x[1]; m[2] = x[2];
        if (KIND == DPS_WREN) m[3] = 36.0f;
    }
    unsigned fg;
    if (KIND == DPS_MEAN) {
        float dist = 0.f;                                // SubtractPixel :77-99
#pragma unroll
        for (int c = 0; c < 3; c++) { const float d = __fsub_rn(x[c], m[c]); dist = __fadd_rn(dist, __fmul_rn(d, d)); }
        fg = dist > L.high_f ? 255u : 0u;
#pragma unroll
        for (int c = 0; c < 3; c++)                      // Update :54-75
            m[c] = __fadd_rn(__fmul_rn(L.alpha, m[c]), __fmul_rn(L.one_minus_alpha, x[c]));
    } else {
        float d[3], dist = 0.f;                          // SubtractPixel :121-147
#pragma unroll
        for (int c = 0; c < 3; c++) { d[c] = __fsub_rn(m[c], x[c]); dist = __fadd_rn(dist, __fmul_rn(d[c], d[c])); }
        fg = dist > __fmul_rn(L.high_f, m[3]) ? 255u : 0u;
#pragma unroll
        for (int c = 0; c < 3; c++) m[c] = __fsub_rn(m[c], __fmul_rn(L.alpha, d[c]));           // Update :97-105
        const float sig = __fadd_rn(m[3], __fmul_rn(L.alpha, __fsub_rn(dist, m[3])));
        m[3] = sig < 4.f ? 4.f : (sig > 180.f ? 180.f : sig);                                   // 5 * m_variance
    }
    return fg;
}

template <int KIND>
__global__ void __launch_bounds__(256)
dps_float_kernel(DpsLaunch L)
{
    pdl_entry();
    constexpr int NP = KIND == DPS_MEAN ? 3 : 4;
    const long long px0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (px0 >= L.npx) return;
    const int n = (int)min(4LL, (long long)L.npx - px0);
    const size_t s = blockIdx.y;
    const uint8_t *fp = L.frame + s * L.frame_stride + px0 * 3;
    uint8_t *gp = L.fg + s * L.fg_stride + px0;
    float *st = L.state + s * NP * L.pstride + px0;      // plane q: st + q * pstride
    const bool v12 = n == 4 && (reinterpret_cast<uintptr_t>(fp) & 3) == 0;
    const bool v4 = n == 4 && (reinterpret_cast<uintptr_t>(gp) & 3) == 0;
    const bool v16 = n == 4 && (reinterpret_cast<uintptr_t>(st) & 15) == 0 && (L.pstride & 3) == 0;
    unsigned in[3];
    dps_load12(fp, v12, n, in);
    float pl[NP][4];
#pragma unroll
    for (int q = 0; q < NP; q++) {
        if (L.fresh) { pl[q][0] = pl[q][1] = pl[q][2] = pl[q][3] = 0.f; }
        else if (v16) { const float4 v = ld_stream_f4(st + q * L.pstride); pl[q][0] = v.x; pl[q][1] = v.y; pl[q][2] = v.z; pl[q][3] = v.w; }
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) pl[q][j] = j < n ? st[q * L.pstride + j] : 0.f;
        }
    }
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float x[3], mod[4];
#pragma unroll
        for (int c = 0; c < 3; c++) x[c] = (float)dps_byte(in, 3 * j + c);
#pragma unroll
        for (int q = 0; q < NP; q++) mod[q] = pl[q][j];
        m |= dps_float_pixel<KIND>(L, x, mod) << (8 * j);
#pragma unroll
        for (int q = 0; q < NP; q++) pl[q][j] = mod[q];
    }
    dps_store_mask(gp, v4, n, m);
#pragma unroll
    for (int q = 0; q < NP; q++) {
        if (v16) st_stream_f4(st + q * L.pstride, make_float4(pl[q][0], pl[q][1], pl[q][2], pl[q][3]));
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < n) st[q * L.pstride + j] = pl[q][j];
        }
    }
}

int launch_dp_simple(const DpsLaunch &L, int nstreams, cudaStream_t stream)
{
    const long long groups = ((long long)L.npx + 3) / 4;
    const dim3 grid((unsigned)((groups + 255) / 256), (unsigned)nstreams);
    if (L.kind == DPS_MEDIAN) launch_pdl(dps_median_kernel, grid, dim3(256), 0, stream, L);
    else if (L.kind == DPS_MEAN) launch_pdl(dps_float_kernel<DPS_MEAN>, grid, dim3(256), 0, stream, L);
    else if (L.kind == DPS_WREN) launch_pdl(dps_float_kernel<DPS_WREN>, grid, dim3(256), 0, stream, L);
    else { set_error("launch_dp_simple: bad kind %d", L.kind); return BGSB_ERR_ARG; }
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
}

}  // namespace bgsb
