// K-GEN: deterministic synthetic video of SURVEY.md 8(d), generated straight into HBM so that
// roofline runs do not cross PCIe.  Integer-only: tracking_b200/synth.py (numpy) and the C oracle
// (oracle/c/bgs_oracle.c: orc_synth_frame) produce byte-identical frames.
//   frame(s,t)[y,x,c] = clamp_u8(B[y,x,c] + N(s,t,y,x,c)), overwritten by 12 moving rectangles.
#include "common.cuh"
#include "kernels.h"

namespace bgsb {

__device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256)
synth_kernel(uint8_t *frames, int T, int w, int h, int t0, uint32_t seed0)
{
    const long long npx = (long long)w * h;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    const int t = blockIdx.y, s = blockIdx.z;
    const int x = (int)(p % w), y = (int)(p / w);
    const int tt = t0 + t;
    const uint32_t seed = seed0 + (uint32_t)s;
    const int scale = (w >= 3840) ? 2 : 1;
    unsigned v[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int B = 32 + (((x * 5 + y * 3 + 64 * c) >> 3) & 127) + 48 * (((x >> 6) ^ (y >> 6)) & 1);
        uint32_t hsh = ((uint32_t)x * 73856093U) ^ ((uint32_t)y * 19349663U) ^
                       ((uint32_t)tt * 83492791U) ^ ((uint32_t)c * 2654435761U) ^ seed;
        int N = (int)(mix32(hsh) % 13U) - 6;
        v[c] = (unsigned)min(max(B + N, 0), 255);
    }
#pragma unroll 1
    for (int r = 0; r < 12; r++) {     // later rectangles overwrite earlier ones
        int rw = 100 * scale, rh = 80 * scale;
        int x0 = (100 + 150 * r + 17 * tt) % (w - 120 * scale);
        int y0 = (60 + 83 * r + 5 * tt) % (h - 90 * scale);
        if (x >= x0 && x < x0 + rw && y >= y0 && y < y0 + rh) {
            v[0] = (unsigned)((40 * r) & 255); v[1] = (unsigned)(255 - 20 * r); v[2] = 128u;
        }
    }
    uint8_t *o = frames + (((size_t)s * T + t) * npx + p) * 3;
    o[0] = (uint8_t)v[0]; o[1] = (uint8_t)v[1]; o[2] = (uint8_t)v[2];
}

// Mode-churn video (the dense case of SURVEY 8(d): every pixel keeps all K = 5 mixture modes live): each pixel shows one
// of five well-separated colours, redrawn every second frame by a hash of (x, y, t / 2, stream), plus noise in [-2, 2].
//   colour i = (20 + 50 i, 230 - 45 i, (90 + 110 i) mod 256)
__global__ void __launch_bounds__(256)
synth_churn_kernel(uint8_t *frames, int T, int w, int h, int t0, uint32_t seed0)
{
    const long long npx = (long long)w * h;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npx) return;
    const int t = blockIdx.y, s = blockIdx.z;
    const int x = (int)(p % w), y = (int)(p / w);
    const int tt = t0 + t;
    const uint32_t seed = seed0 + (uint32_t)s;
    const uint32_t hi = ((uint32_t)x * 73856093U) ^ ((uint32_t)y * 19349663U) ^ ((uint32_t)(tt >> 1) * 83492791U) ^ (seed * 2246822519U);
    const int i = (int)(mix32(hi) % 5U);
    const int col[3] = {20 + 50 * i, 230 - 45 * i, (90 + 110 * i) & 255};
    uint8_t *o = frames + (((size_t)s * T + t) * npx + p) * 3;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const uint32_t hsh = ((uint32_t)x * 73856093U) ^ ((uint32_t)y * 19349663U) ^
                             ((uint32_t)tt * 83492791U) ^ ((uint32_t)c * 2654435761U) ^ seed;
        const int N = (int)(mix32(hsh) % 5U) - 2;
        o[c] = (uint8_t)min(max(col[c] + N, 0), 255);
    }
}

int launch_synth_churn(uint8_t *d_frames, int nstreams, int T, int w, int h, int t0, uint32_t seed0, cudaStream_t stream)
{
    if (w <= 0 || h <= 0) { set_error("synth: empty frame"); return BGSB_ERR_ARG; }
    const long long npx = (long long)w * h;
    dim3 grid((unsigned)((npx + 255) / 256), (unsigned)T, (unsigned)nstreams);
    synth_churn_kernel<<<grid, 256, 0, stream>>>(d_frames, T, w, h, t0, seed0);
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
}

int launch_synth(uint8_t *d_frames, int nstreams, int T, int w, int h, int t0, uint32_t seed0,
                 cudaStream_t stream)
{
    if (w <= 240 || h <= 180) { set_error("synth: frame must be larger than 240x180"); return BGSB_ERR_ARG; }
    const long long npx = (long long)w * h;
    dim3 grid((unsigned)((npx + 255) / 256), (unsigned)T, (unsigned)nstreams);
    synth_kernel<<<grid, 256, 0, stream>>>(d_frames, T, w, h, t0, seed0);
    BGSB_LAUNCH_CHECK();
    return BGSB_OK;
}

}  // namespace bgsb
