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

// ---------------------------------------------------------------------------------------------------------------
// K-PRATI: DPPratiMediodBGS, USTC_BGS type 14 (package_bgs/dp/PratiMediodBGS.cpp:52-275, wrapper DPPratiMediodBGS.cpp:28-88).
// Pure integer.  The reference keeps, per pixel, std::vectors of up to H sampled pixels and of each sample's sum of L-inf
// distances to the others; the medoid (smallest sum, first wins) is the background.  The wrapper clears the update mask,
// so every pixel is sampled on the same frames: the per-pixel vectors are whole frames (samples[s]) and whole planes
// (dist[s], 16 bit: a sum never exceeds (H + 1) * 255), buffer length n and write position pos are host counters.
//   subtract (every frame >= H):  L-inf distance to the medoid against low / high threshold, then the hysteresis of
//       Combine(): high -> foreground; low -> foreground iff one of the 8 neighbours is high; image border -> background.
//       A thread owns 4 pixels; the neighbour test (rare: low-but-not-high pixels) recomputes the neighbours' distance.
//   update (frames with frame_num % samplingRate == 0):  one pass over the n samples per pixel -- remove the replaced
//       sample's distances (ring full), add the new pixel's, track the medoid -- then the new pixel takes the slot.
//       The quirks are the reference's and are kept: the replaced sample still competes with its old sum, and the new
//       sample's sum includes its distance to the sample it replaces.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned prati_linf3(unsigned px)           // max of the three low bytes
{
    const unsigned a = px & 0xffu, b = (px >> 8) & 0xffu, c = (px >> 16) & 0xffu;
    return max(a, max(b, c));
}
// pixel j (0..3) of three words as B | G << 8 | R << 16 (+ garbage in the top byte)
__device__ __forceinline__ unsigned prati_px(const unsigned (&w)[3], int j)
{
    return j == 0 ? w[0] : (j == 1 ? __byte_perm(w[0], w[1], 0x0543u) : (j == 2 ? __byte_perm(w[1], w[2], 0x0432u) : (w[2] >> 8)));
}

__global__ void __launch_bounds__(256)
prati_subtract_kernel(PratiLaunch L)
{
    pdl_entry();
    const int npx = L.w * L.h;
    const long long px0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (px0 >= npx) return;
    const int n = (int)min(4LL, (long long)npx - px0);
    const size_t s = blockIdx.y;
    const uint8_t *frame = L.frame + s * L.frame_stride;
    const uint8_t *median = L.state + s * L.stream_bytes + (size_t)L.H * L.plane3;
    uint8_t *gp = L.fg + s * L.fg_stride + px0;
    const uint8_t *fp = frame + px0 * 3, *mp = median + px0 * 3;
    const bool v12 = n == 4 && ((reinterpret_cast<uintptr_t>(fp) | reinterpret_cast<uintptr_t>(mp)) & 3) == 0;
    const bool v4 = n == 4 && (reinterpret_cast<uintptr_t>(gp) & 3) == 0;
    unsigned in[3], med[3], d[3];
    dps_load12(fp, v12, n, in);
    dps_load12(mp, v12, n, med);
#pragma unroll
    for (int k = 0; k < 3; k++) d[k] = __vabsdiffu4(in[k], med[k]);
    unsigned m = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (j >= n) break;
        const unsigned dist = prati_linf3(prati_px(d, j));                  // CalculateMasks :204-234
        unsigned out = 0;
        const int p = (int)px0 + j, r = p / L.w, c = p - r * L.w;
        if (r > 0 && c > 0 && r < L.h - 1 && c < L.w - 1) {                 // Combine :167-202
            if (dist > L.high) out = 255u;
            else if (dist > L.low) {
                for (int dr = -1; dr <= 1 && !out; dr++)
                    for (int dc = -1; dc <= 1; dc++) {
                        if (!dr && !dc) continue;
                        const size_t q = (size_t)(p + dr * L.w + dc) * 3;
                        const unsigned nd = max(max(__sad((int)frame[q], (int)median[q], 0u), __sad((int)frame[q + 1], (int)median[q + 1], 0u)),
                                                __sad((int)frame[q + 2], (int)median[q + 2], 0u));
                        if (nd > L.high) { out = 255u; break; }
                    }
            }
        }
        m |= out << (8 * j);
    }
    dps_store_mask(gp, v4, n, m);
}

__global__ void __launch_bounds__(256)
prati_update_kernel(PratiLaunch L)
{
    pdl_entry();
    const int npx = L.w * L.h;
    const long long px0 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
    if (px0 >= npx) return;
    const int n = (int)min(4LL, (long long)npx - px0);
    const size_t s = blockIdx.y;
    uint8_t *st = L.state + s * L.stream_bytes;
    uint8_t *samples = st, *median = st + (size_t)L.H * L.plane3;
    unsigned short *dist = reinterpret_cast<unsigned short *>(st + (size_t)(L.H + 1) * L.plane3);
    const uint8_t *fp = L.frame + s * L.frame_stride + px0 * 3;
    const bool vin = n == 4 && (reinterpret_cast<uintptr_t>(fp) & 3) == 0;
    const bool vec = n == 4;                                                // the state planes are padded and 16-byte aligned
    const bool full = L.n == L.H;
    unsigned in[3], old[3] = {0u, 0u, 0u}, med[3];
    dps_load12(fp, vin, n, in);
    if (full) dps_load12(samples + (size_t)L.pos * L.plane3 + px0 * 3, vec, n, old);
    dps_load12(median + px0 * 3, vec, n, med);                              // kept where no sample wins (cannot happen: n >= 0 -> the new pixel does)
    unsigned best[4], Lsum[4] = {0u, 0u, 0u, 0u}, mpx[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { best[j] = 0x7fffffffu; mpx[j] = prati_px(med, j) & 0x00ffffffu; }
    for (int k = 0; k < L.n; k++) {
        unsigned sp[3], dn[3], dol[3];
        dps_load12(samples + (size_t)k * L.plane3 + px0 * 3, vec, n, sp);
        unsigned short *dp = dist + (size_t)k * L.plane1 + px0;
        unsigned ds[4];
        if (vec) { const uint2 v = *reinterpret_cast<const uint2 *>(dp); ds[0] = v.x & 0xffffu; ds[1] = v.x >> 16; ds[2] = v.y & 0xffffu; ds[3] = v.y >> 16; }
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) ds[j] = j < n ? dp[j] : 0u;
        }
#pragma unroll
        for (int q = 0; q < 3; q++) { dn[q] = __vabsdiffu4(sp[q], in[q]); dol[q] = __vabsdiffu4(sp[q], old[q]); }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const unsigned dnew = prati_linf3(prati_px(dn, j));
            if (full) ds[j] -= prati_linf3(prati_px(dol, j));               // Update :84-95
            ds[j] += dnew;                                                  // UpdateMediod :143-155
            if (ds[j] < best[j]) { best[j] = ds[j]; mpx[j] = prati_px(sp, j) & 0x00ffffffu; }
            Lsum[j] += dnew;
        }
        if (vec) *reinterpret_cast<uint2 *>(dp) = make_uint2(ds[0] | (ds[1] << 16), ds[2] | (ds[3] << 16));
        else {
#pragma unroll
            for (int j = 0; j < 4; j++) if (j < n) dp[j] = (unsigned short)ds[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (Lsum[j] < best[j]) mpx[j] = prati_px(in, j) & 0x00ffffffu;      // the new point is the medoid :160-164
    // medoid image, the new sample and its sum into the slot (:97-100 / :119-121)
    unsigned mo[3];
    mo[0] = mpx[0] | (mpx[1] << 24); mo[1] = (mpx[1] >> 8) | (mpx[2] << 16); mo[2] = (mpx[2] >> 16) | (mpx[3] << 8);
    dps_store12(median + px0 * 3, vec, n, mo);
    const int slot = full ? L.pos : L.n;
    dps_store12(samples + (size_t)slot * L.plane3 + px0 * 3, vec, n, in);
    unsigned short *dq = dist + (size_t)slot * L.plane1 + px0;
    if (vec) *reinterpret_cast<uint2 *>(dq) = make_uint2(Lsum[0] | (Lsum[1] << 16), Lsum[2] | (Lsum[3] << 16));
    else {
#pragma unroll
        for (int j = 0; j < 4; j++) if (j < n) dq[j] = (unsigned short)Lsum[j];
    }
}

int launch_prati_subtract(const PratiLaunch &L, int nstreams, cudaStream_t stream)
{
    const long long groups = ((long long)L.w * L.h + 3) / 4;
    launch_pdl(prati_subtract_kernel, dim3((unsigned)((groups + 255) / 256), (unsigned)nstreams), dim3(256), 0, stream, L);
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
}

int launch_prati_update(const PratiLaunch &L, int nstreams, cudaStream_t stream)
{
    const long long groups = ((long long)L.w * L.h + 3) / 4;
    launch_pdl(prati_update_kernel, dim3((unsigned)((groups + 255) / 256), (unsigned)nstreams), dim3(256), 0, stream, L);
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
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
