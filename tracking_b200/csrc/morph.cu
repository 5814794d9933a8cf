// K-MORPH: cv::erode / cv::dilate with the default 3x3 rectangular element on {0,255} masks,
// as bit-parallel AND/OR on bit-packed rows, shared-memory tiled, whole op chains fused per tile.
//
// Replaces cv::erode/cv::dilate(mask, cv::Mat()) and cvErode/cvDilate(mask, mask, NULL, n) call
// sites (package_bgs/jmo/CMultiLayerBGS.cpp:1614-1615, package_bgs/tb/PixelUtils.cpp:73-74,
// package_bgs/av/VuMeter.cpp:68) and the open/close of the stock FG detectors reached through
// ustc_src/trackingMain.cpp:616-618.  Semantics: anchor at the centre, pixels
// outside the image are ignored (erode pads +inf, dilate pads -inf), `iterations = n` == n passes.
//
// One CTA owns a tile of TH rows x TW 32-px words.  It packs the input bytes of the tile plus a
// halo of R = sum(iterations) rows (and one halo word = 32 px on each side) into shared memory,
// runs every pass of the chain there (ping-pong buffers, one __syncthreads per pass; each pass
// is 3 shifted ANDs/ORs per word and per row), and unpacks the interior back to bytes.  HBM
// traffic is one byte read and one byte written per pixel for the whole chain.
#include <string.h>
#include <algorithm>
#include "ccl_internal.h"
#include "kernels.h"

namespace bgsb {

constexpr int MORPH_TW = 62;        // interior words per tile (+2 halo words = 64)
constexpr int MORPH_RMAX = 8;       // max total iterations fused in one launch
constexpr int MORPH_SW = MORPH_TW + 2;
// interior rows per tile: a template parameter of the kernel, 16 in production

struct MorphChain {
    int n;
    signed char op[16];
    signed char iters[16];
};

// bits of word k of an image row that lie inside the image
__device__ __forceinline__ unsigned inside_mask(int y, int k, int w, int h, int wpr)
{
    if (y < 0 || y >= h || k < 0 || k >= wpr) return 0u;
    if (k < wpr - 1 || (w & 31) == 0) return 0xffffffffu;
    return (1u << (w & 31)) - 1u;
}

// 32 mask bytes (as eight words) -> one word of "byte is non-zero" bits.  Per word, bit 7 of every byte says
// non-zero; the flags of two words (the first shifted down by 4: bits 3, 11, 19, 27 and 7, 15, 23, 31) go through ONE
// multiply by 2^0 + 2^7 + 2^14 + 2^21, which lands them on bits 24..31 in pixel order -- the other partial products
// fall on distinct bits below 24 or above 31, so nothing carries.
__device__ __forceinline__ unsigned nz_flags(unsigned v) { return (((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v) & 0x80808080u; }
__device__ __forceinline__ unsigned pack32(const unsigned (&ws)[8])
{
    unsigned word = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const unsigned g = (nz_flags(ws[2 * i]) >> 4) | nz_flags(ws[2 * i + 1]);
        word |= ((g * 0x00204081u) >> 24) << (8 * i);
    }
    return word;
}

// nibble -> four {0,255} bytes: one multiply puts bit j of the nibble on bit 7 of byte j, PRMT in its
// sign-replicating mode fills the bytes
__device__ __forceinline__ unsigned unpack4(unsigned nib)
{
    unsigned d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(nib * 0x10204080u), "r"(0u), "r"(0xBA98u));   // selector bit 3: replicate the msb
    return d;
}

// Thread (tx, ty) = (tid & 63, tid >> 6) owns shared-memory column tx and rows ty, ty+4, ...: the column's
// in-image masks are computed once, and no index is divided.
// IN_BITS: the input is already bit-packed ([nimages][h][wpr] words, e.g. the MOG2 kernel's packed mask) -- the tile
// load is one word per thread and row instead of 32 bytes.  Outputs (either or both): {0,255} bytes, and / or
// bit-packed words; with `parent` the words' runs also become the labeller's union-find nodes (ccl_internal.h), which
// saves the labeller's own first launch.
template <bool IN_BITS, int MORPH_TH>
__global__ void __launch_bounds__(256)
morph_kernel(const MorphIO io, int w, int h, int wpr, MorphChain chain, int R, int ntx, int nty, int ntiles)
{
    pdl_entry();
    constexpr int MORPH_ROWS = MORPH_TH + 2 * MORPH_RMAX;
    __shared__ unsigned buf[2][MORPH_ROWS][MORPH_SW];
    // a CTA takes tiles tile = blockIdx.x, blockIdx.x + gridDim.x, ... of all images (normally one tile per CTA; the
    // pipeline caps the grid so that the kernel runs beside the plugin kernel without evicting its CTAs)
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tbx = tile % ntx, tby = (tile / ntx) % nty;
    const int img = tile / (ntx * nty);
    const size_t npx = (size_t)w * h, nwords = (size_t)wpr * h;
    const uint8_t *in = IN_BITS ? nullptr : io.in_bytes + img * npx;
    const unsigned *inb = IN_BITS ? io.in_bits + img * nwords : nullptr;
    uint8_t *out = io.out_bytes ? io.out_bytes + img * npx : nullptr;
    unsigned *outb = io.out_bits ? io.out_bits + img * nwords : nullptr;
    int *parent = (io.parent && outb) ? io.parent + img * npx : nullptr;
    const int k0 = tbx * MORPH_TW - 1;                 // word index of smem column 0 (halo)
    const int y0 = tby * MORPH_TH - R;                 // image row of smem row 0 (halo)
    const int rows = MORPH_TH + 2 * R;
    const int rblock = (rows + 3) >> 2;                 // consecutive rows per thread in the passes (4 row groups)
    const int tw = min(MORPH_TW, wpr - tbx * MORPH_TW) + 2;   // smem columns in use
    const int c = threadIdx.x & 63, ty = threadIdx.x >> 6;
    const int k = k0 + c;
    const bool col_used = c < tw;
    // in-image bits of this column and of its neighbours (row validity is added per row)
    auto colmask = [&](int kk) -> unsigned {
        if (kk < 0 || kk >= wpr) return 0u;
        if (kk < wpr - 1 || (w & 31) == 0) return 0xffffffffu;
        return (1u << (w & 31)) - 1u;
    };
    const unsigned cm = colmask(k), cl = colmask(k - 1), cr = colmask(k + 1);
    const bool vec = !IN_BITS && cm == 0xffffffffu && ((reinterpret_cast<uintptr_t>(in) | (size_t)w) & 15) == 0;   // 32 whole, aligned bytes

    // ---- load: bytes -> bits (bit i of word k = pixel 32k+i is set), or the packed words as they are ----
    if (col_used) {
        for (int r = ty; r < rows; r += 4) {
            const int y = y0 + r;
            unsigned word = 0;
            if (y >= 0 && y < h && cm) {
                if (IN_BITS) word = inb[(size_t)y * wpr + k] & cm;
                else {
                    const uint8_t *p = in + (size_t)y * w + (size_t)k * 32;
                    if (vec) {
                        const uint4 *q = reinterpret_cast<const uint4 *>(p);
                        const uint4 a = __ldg(q), b = __ldg(q + 1);
                        const unsigned ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                        word = pack32(ws);
                    } else {
                        const int nvalid = min(32, w - k * 32);
                        for (int i = 0; i < nvalid; i++) word |= (p[i] ? 1u : 0u) << i;
                    }
                }
            }
            buf[0][r][c] = word;
        }
    }
    __syncthreads();

    // ---- the chain, entirely in shared memory ----
    int cur = 0;
    for (int o = 0; o < chain.n; o++) {
        const bool dil = chain.op[o] == BGSB_MORPH_DILATE;
        for (int it = 0; it < chain.iters[o]; it++) {
            // The 3x3 rectangle is separable: h(row) = the row combined with its two horizontal neighbours, and the
            // result is h(r-1) op h(r) op h(r+1).  A thread owns a block of consecutive rows of its column and slides
            // the three h values through registers: three shared-memory loads per output word instead of nine.
            const int r0 = ty * rblock, r1 = min(rows, r0 + rblock);
            if (col_used && r0 < r1) {
                auto hrow = [&](int rr) -> unsigned {
                    unsigned L = 0, M = 0, Rt = 0;
                    if (rr >= 0 && rr < rows) {
                        M = buf[cur][rr][c];
                        if (c > 0) L = buf[cur][rr][c - 1];
                        if (c + 1 < tw) Rt = buf[cur][rr][c + 1];
                    }
                    if (!dil) {   // outside-image pixels are neutral (all ones) for erosion
                        const bool rowin = (y0 + rr >= 0 && y0 + rr < h);
                        M |= rowin ? ~cm : 0xffffffffu;
                        L |= rowin ? ~cl : 0xffffffffu;
                        Rt |= rowin ? ~cr : 0xffffffffu;
                    }
                    const unsigned left = (M << 1) | (L >> 31);      // neighbour x-1
                    const unsigned right = (M >> 1) | (Rt << 31);    // neighbour x+1
                    return dil ? (M | left | right) : (M & left & right);
                };
                unsigned hp = hrow(r0 - 1), hc = hrow(r0);
                for (int r = r0; r < r1; r++) {
                    const unsigned hn = hrow(r + 1);
                    const unsigned acc = dil ? (hp | hc | hn) : (hp & hc & hn);
                    const int y = y0 + r;
                    buf[cur ^ 1][r][c] = (y >= 0 && y < h) ? (acc & cm) : 0u;
                    hp = hc; hc = hn;
                }
            }
            __syncthreads();
            cur ^= 1;
        }
    }

    // ---- store the interior: packed words (+ labeller nodes) and / or {0,255} bytes ----
    if (c >= 1 && c < tw - 1) {
        for (int r = ty; r < MORPH_TH; r += 4) {
            const int y = tby * MORPH_TH + r;
            if (y >= h) break;
            const unsigned word = buf[cur][r + R][c];
            if (outb) {
                outb[(size_t)y * wpr + k] = word;
                if (parent) ccl_init_word(parent, ccl_border(word, y, k, w, h, wpr, io.zero_border), y * w + k * 32);
            }
            if (out) {
                uint8_t *p = out + (size_t)y * w + (size_t)k * 32;
                if (cm == 0xffffffffu && ((reinterpret_cast<uintptr_t>(out) | (size_t)w) & 15) == 0) {
                    uint4 *q = reinterpret_cast<uint4 *>(p);
                    q[0] = make_uint4(unpack4(word & 0xfu), unpack4((word >> 4) & 0xfu), unpack4((word >> 8) & 0xfu),
                                      unpack4((word >> 12) & 0xfu));
                    q[1] = make_uint4(unpack4((word >> 16) & 0xfu), unpack4((word >> 20) & 0xfu), unpack4((word >> 24) & 0xfu),
                                      unpack4(word >> 28));
                } else {
                    const int nvalid = min(32, w - k * 32);
                    for (int i = 0; i < nvalid; i++) p[i] = ((word >> i) & 1u) ? 255 : 0;
                }
            }
        }
    }
    __syncthreads();                                   // the tile buffers are reused by the CTA's next tile
    }
}

static int parse_chain(const int *ops, int nops, int (*items)[2], int &ni, int &total)
{
    ni = 0; total = 0;
    for (int i = 0; i < nops; i++) {
        int op = ops[2 * i], it = ops[2 * i + 1];
        if (op != BGSB_MORPH_ERODE && op != BGSB_MORPH_DILATE) { set_error("morph: bad op %d", op); return BGSB_ERR_ARG; }
        if (it < 0 || it > 64) { set_error("morph: iterations out of range"); return BGSB_ERR_ARG; }
        while (it > 0) {
            int take = it < MORPH_RMAX ? it : MORPH_RMAX;
            if (ni >= 64) { set_error("morph: chain too long"); return BGSB_ERR_ARG; }
            items[ni][0] = op; items[ni][1] = take; ni++;
            it -= take; total += take;
        }
    }
    return BGSB_OK;
}

// The chain on any mix of byte / bit-packed inputs and outputs.  Passes of at most MORPH_RMAX total iterations share a
// launch; between launches the mask stays bit-packed (1/8 of the byte traffic).  An empty chain still converts
// (bytes != 0 -> set bits -> {0,255} bytes).
int launch_morph_chain_io(const MorphIO &io_, int w, int h, int nimages, const int *ops, int nops, cudaStream_t stream)
{
    if ((io_.in_bytes == nullptr) == (io_.in_bits == nullptr)) { set_error("morph: exactly one input form"); return BGSB_ERR_ARG; }
    if (!io_.out_bytes && !io_.out_bits) { set_error("morph: no output"); return BGSB_ERR_ARG; }
    int items[64][2], ni = 0, total = 0;
    int rc = parse_chain(ops, nops, items, ni, total);
    if (rc) return rc;
    // group items into launches
    int nlaunch = 0, lstart[65], lend[65];
    for (int i = 0; i < ni;) {
        int r = 0, j = i;
        while (j < ni && r + items[j][1] <= MORPH_RMAX && j - i < 16) { r += items[j][1]; j++; }
        lstart[nlaunch] = i; lend[nlaunch] = j; nlaunch++;
        i = j;
    }
    if (nlaunch == 0) { lstart[0] = lend[0] = 0; nlaunch = 1; }          // pure conversion
    // a CTA reads a halo that another CTA may already have overwritten: in-place runs go through a packed temporary
    const bool inplace = (io_.in_bytes && io_.in_bytes == io_.out_bytes) || (io_.in_bits && io_.in_bits == io_.out_bits);
    if (inplace && nlaunch == 1 && lend[0] > lstart[0]) { lstart[1] = lend[1] = ni; nlaunch = 2; }
    const int wpr = (w + 31) / 32;
    const size_t tbytes = (size_t)wpr * h * nimages * sizeof(unsigned);
    unsigned *tmp[2] = {nullptr, nullptr};
    const int ntmp = nlaunch > 2 ? 2 : (nlaunch > 1 ? 1 : 0);
    for (int i = 0; i < ntmp; i++) BGSB_CUDA(cudaMallocAsync(&tmp[i], tbytes, stream));
    const int TH = 16;      // (64-row tiles for batches were measured and dropped: OPEN on 8 masks 12.7 -> 26.3 us)
    const int ntx = (wpr + MORPH_TW - 1) / MORPH_TW, nty = (h + TH - 1) / TH;
    const long long ntiles_ll = (long long)ntx * nty * nimages;
    if (ntiles_ll > 0x7fffffffLL) { set_error("morph: too many tiles"); return BGSB_ERR_ARG; }
    const int ntiles = (int)ntiles_ll;
    const dim3 grid((unsigned)(io_.max_ctas > 0 ? std::min(ntiles, io_.max_ctas) : ntiles));
    const unsigned *src_bits = nullptr;
    for (int l = 0; l < nlaunch; l++) {
        const bool first = (l == 0), last = (l == nlaunch - 1);
        MorphIO io;
        memset(&io, 0, sizeof(io));
        if (first) { io.in_bytes = io_.in_bytes; io.in_bits = io_.in_bits; }
        else io.in_bits = src_bits;
        if (last) { io.out_bytes = io_.out_bytes; io.out_bits = io_.out_bits; io.parent = io_.parent; io.zero_border = io_.zero_border; }
        else io.out_bits = (src_bits == tmp[0]) ? tmp[1] : tmp[0];
        MorphChain ch;
        ch.n = 0;
        int R = 0;
        for (int i = lstart[l]; i < lend[l]; i++) {
            ch.op[ch.n] = (signed char)items[i][0]; ch.iters[ch.n] = (signed char)items[i][1]; ch.n++;
            R += items[i][1];
        }
        if (io.in_bits) launch_pdl(morph_kernel<true, 16>, dim3(grid), dim3(256), 0, stream, io, w, h, wpr, ch, R, ntx, nty, ntiles);
        else launch_pdl(morph_kernel<false, 16>, dim3(grid), dim3(256), 0, stream, io, w, h, wpr, ch, R, ntx, nty, ntiles);
        BGSB_LAUNCH_CHECK();
        src_bits = io.out_bits;
    }
    for (int i = 0; i < ntmp; i++) BGSB_CUDA(cudaFreeAsync(tmp[i], stream));
    return BGSB_OK;
}

// byte masks in and out (bgsb_morph_dev).  An empty chain is a no-op COPY (iterations = 0): values pass
// through unchanged.
int launch_morph_chain(const uint8_t *d_in, uint8_t *d_out, int w, int h, int nimages, const int *ops, int nops,
                       cudaStream_t stream)
{
    int items[64][2], ni = 0, total = 0;
    int rc = parse_chain(ops, nops, items, ni, total);
    if (rc) return rc;
    if (total == 0) {
        if (d_in != d_out) BGSB_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)w * h * nimages, cudaMemcpyDeviceToDevice, stream));
        return BGSB_OK;
    }
    MorphIO io;
    memset(&io, 0, sizeof(io));
    io.in_bytes = d_in; io.out_bytes = d_out;
    return launch_morph_chain_io(io, w, h, nimages, ops, nops, stream);
}

}  // namespace bgsb

using namespace bgsb;

extern "C" {

int bgsb_morph_dev(const uint8_t *d_mask, int w, int h, int nimages, const int *ops, int nops, uint8_t *d_out,
                   void *stream)
{
    BGSB_REQUIRE(d_mask && d_out, "null");
    BGSB_REQUIRE(w > 0 && h > 0 && nimages > 0, "empty image");
    BGSB_REQUIRE(nops >= 0 && (nops == 0 || ops), "ops");
    return launch_morph_chain(d_mask, d_out, w, h, nimages, ops, nops, (cudaStream_t)stream);
}

// Host-buffer convenience call (cvErode(mask, mask, NULL, n) pattern).  The two device staging buffers are cached
// per host thread and device (grow-only) so that a per-frame caller does not pay cudaMalloc/cudaFree every call.
namespace {
struct MorphScratch {
    uint8_t *a = nullptr, *b = nullptr;
    size_t cap = 0;
    int device = -1;
    // small masks (the reference's 320x240 class) go through pinned staging on a private stream: two short
    // asynchronous copies and one synchronisation instead of two blocking pageable copies on the legacy stream
    static constexpr size_t PIN_MAX = 512 * 1024;
    uint8_t *pin = nullptr;
    cudaStream_t stream = nullptr;
    ~MorphScratch() { if (a) cudaFree(a); if (b) cudaFree(b); if (pin) cudaFreeHost(pin); if (stream) cudaStreamDestroy(stream); }
};
thread_local MorphScratch g_morph_scratch;
}  // namespace

int bgsb_morph(const uint8_t *mask, int w, int h, size_t stride, const int *ops, int nops, uint8_t *out,
               size_t out_stride)
{
    BGSB_REQUIRE(mask && out, "null");
    BGSB_REQUIRE(w > 0 && h > 0, "empty image");
    BGSB_REQUIRE(stride >= (size_t)w && out_stride >= (size_t)w, "stride smaller than a row");
    const size_t bytes = (size_t)w * h;
    int dev = 0;
    BGSB_CUDA(cudaGetDevice(&dev));
    MorphScratch &S = g_morph_scratch;
    if (S.device != dev || S.cap < bytes) {
        if (S.a) cudaFree(S.a);
        if (S.b) cudaFree(S.b);
        if (S.device != dev && S.stream) { cudaStreamDestroy(S.stream); S.stream = nullptr; }   // a stream belongs to its device
        S.a = S.b = nullptr; S.cap = 0; S.device = dev;
        BGSB_CUDA(cudaMalloc(&S.a, bytes));
        BGSB_CUDA(cudaMalloc(&S.b, bytes));
        S.cap = bytes;
    }
    if (!S.stream) BGSB_CUDA(cudaStreamCreateWithFlags(&S.stream, cudaStreamNonBlocking));
    if (bytes <= MorphScratch::PIN_MAX) {
        if (!S.pin) BGSB_CUDA(cudaMallocHost(&S.pin, 2 * MorphScratch::PIN_MAX));
        for (int y = 0; y < h; y++) memcpy(S.pin + (size_t)y * w, mask + (size_t)y * stride, (size_t)w);
        BGSB_CUDA(cudaMemcpyAsync(S.a, S.pin, bytes, cudaMemcpyHostToDevice, S.stream));
        int rc = launch_morph_chain(S.a, S.b, w, h, 1, ops, nops, S.stream);
        if (rc) return rc;
        uint8_t *back = S.pin + MorphScratch::PIN_MAX;
        BGSB_CUDA(cudaMemcpyAsync(back, S.b, bytes, cudaMemcpyDeviceToHost, S.stream));
        BGSB_CUDA(cudaStreamSynchronize(S.stream));
        for (int y = 0; y < h; y++) memcpy(out + (size_t)y * out_stride, back + (size_t)y * w, (size_t)w);
        return BGSB_OK;
    }
    BGSB_CUDA(cudaMemcpy2DAsync(S.a, w, mask, stride, w, h, cudaMemcpyHostToDevice, S.stream));
    int rc = launch_morph_chain(S.a, S.b, w, h, 1, ops, nops, S.stream);
    if (rc) return rc;
    BGSB_CUDA(cudaMemcpy2DAsync(out, out_stride, S.b, w, w, h, cudaMemcpyDeviceToHost, S.stream));
    BGSB_CUDA(cudaStreamSynchronize(S.stream));
    return BGSB_OK;
}

}  // extern "C"
