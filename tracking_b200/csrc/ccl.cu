// K-CCL: 8-connected component labelling of the foreground mask on the GPU, with canonical
// (raster-first) labels, per-component bounding box / area, and the RETR_EXTERNAL nesting test.
//
// Replaces steps 1-2 (and the integer sums of step 4) of CvBlobDetectorCC::DetectNewBlob
// (OpenCV legacy `cvCreateBlobDetectorCC`, referenced at ustc_src/trackingMain.cpp:56,626 and
// readme.md:4-10; spec = SURVEY.md Appendix A.6):
//     cvThreshold(pIB, pIB, 128, 255, CV_THRESH_BINARY); cvFindContours(pIB, ..., CV_RETR_EXTERNAL);
//     CvRect r = ((CvContour*)cnt)->rect;   cvMoments(cvGetSubRect(pFGMask,&mat,r), &m, 0);
//
// Algorithm (all kernels one thread per 32-pixel word of the bit-packed mask; every launch covers a
// whole batch of images, blockIdx.y = image):
//   pack     bytes > 128 -> bits (optionally clearing the 1-px frame, OpenCV 2.4 behaviour)
//   init     every horizontal run (within a word) is a union-find node named by its first pixel
//   merge    unions found with bit tricks on (this row, row above): a run pair is linked exactly
//            once (8-connectivity: vertical + the two diagonals that are not already implied);
//            lock-free union by atomicMin with path halving, so a component's root is its
//            minimum = raster-first pixel
//   flatten  parent[node] = root; mark roots; count roots per 256-word block
//   rank     scan of the block counts + in-block scan -> canonical label = 1 + rank of the root
//   label    write labels, accumulate bbox/area per component with atomics
//   nest     RETR_EXTERNAL drops components enclosed in a hole of another one.  That requires the
//            enclosed component's bounding box to lie STRICTLY inside the other's, so one small
//            kernel checks the boxes; only images where such a pair exists run the background pass:
//   bg pass  the same init/merge/flatten on the 4-connected background; regions touching the frame
//            are "outer"; the background pixel left of a component's first pixel lies in the region
//            surrounding it, and the component is external iff that region is outer.
//            (Typical masks have no such pair and skip it: the background is >98 % of the pixels.)
#include <limits.h>
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace bgsb {

__device__ __forceinline__ unsigned in_mask(int k, int w, int wpr)
{
    if (k < 0 || k >= wpr) return 0u;
    if (k < wpr - 1 || (w & 31) == 0) return 0xffffffffu;
    return (1u << (w & 31)) - 1u;
}

// first bit of the run of ones in v that contains bit b
__device__ __forceinline__ int run_start(unsigned v, int b)
{
    unsigned below = ~v & ((b == 0) ? 0u : (0xffffffffu >> (32 - b)));
    return below ? 32 - __clz(below) : 0;
}

__device__ __forceinline__ int find_root(const int *parent, int p)
{
    int q = parent[p];
    while (q != p) { p = q; q = parent[p]; }
    return p;
}

// find with path halving: every visited node is re-pointed at its grandparent.  Re-pointing uses
// atomicMin so that parents only ever decrease (an ancestor always has a smaller index), which keeps
// the lock-free unions below valid while other threads compress the same paths.
__device__ __forceinline__ int find_compress(int *parent, int p)
{
    int q = parent[p];
    while (q != p) {
        int g = parent[q];
        if (g != q) atomicMin(&parent[p], g);
        p = q; q = g;
    }
    return p;
}

__device__ __forceinline__ void unite(int *parent, int a, int b)
{
    while (true) {
        a = find_compress(parent, a);
        b = find_compress(parent, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }      // a > b: hang a under b
        int old = atomicMin(&parent[a], b);
        if (old == a) return;
        a = old;                                      // somebody re-parented a meanwhile: retry
    }
}

// pack + foreground init in one pass: a word's run starts become their own union-find roots right away (the
// init step only needs the word itself), and the per-block root counters / background-pass flag are cleared.
__global__ void __launch_bounds__(256)
ccl_pack_kernel(const uint8_t *__restrict__ mask, unsigned *__restrict__ bits, int *__restrict__ parent,
                int *__restrict__ blockcount, int *__restrict__ need_bg, int nblocks, int w, int h, int wpr, int zero_border,
                size_t img_px, size_t img_words)
{
    pdl_entry();
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (threadIdx.x == 0) {
        blockcount[blockIdx.y * (size_t)(nblocks + 1) + blockIdx.x] = 0;
        if (blockIdx.x == 0) need_bg[blockIdx.y] = 0;
    }
    if (wi >= h * wpr) return;
    mask += blockIdx.y * img_px; bits += blockIdx.y * img_words; parent += blockIdx.y * img_px;
    int y = wi / wpr, k = wi - y * wpr;
    const uint8_t *p = mask + (size_t)y * w + (size_t)k * 32;
    int nvalid = min(32, w - k * 32);
    unsigned word = 0;
    if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const uint4 *q = reinterpret_cast<const uint4 *>(p);
        uint4 a = __ldg(q), b = __ldg(q + 1);
        unsigned ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned v = ws[i];
            unsigned nz = (((v & 0xffu) > 128u) ? 1u : 0u) | ((((v >> 8) & 0xffu) > 128u) ? 2u : 0u) |
                          ((((v >> 16) & 0xffu) > 128u) ? 4u : 0u) | (((v >> 24) > 128u) ? 8u : 0u);
            word |= nz << (4 * i);
        }
    } else {
        for (int i = 0; i < nvalid; i++) word |= (p[i] > 128 ? 1u : 0u) << i;
    }
    if (zero_border) {
        if (y == 0 || y == h - 1) word = 0;
        if (k == 0) word &= ~1u;
        if (k == wpr - 1) word &= ~(1u << ((w - 1) & 31));
    }
    bits[wi] = word;
    const int base = y * w + k * 32;
    unsigned s = word & ~(word << 1);                  // run starts inside the word
    while (s) {
        const int b = __ffs(s) - 1; s &= s - 1;
        parent[base + b] = base + b;
    }
}

// BG = false: foreground runs;  BG = true: background runs (only for images flagged by the nest check)
template <bool BG>
__global__ void __launch_bounds__(BG ? 1024 : 256)
ccl_init_kernel(const unsigned *__restrict__ bits, int *__restrict__ parent, uint8_t *__restrict__ outer,
                int *__restrict__ blockcount, const int *__restrict__ need_bg, int w, int h, int wpr, size_t img_px,
                size_t img_words, int nblocks)
{
    pdl_entry();
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (BG) { if (!need_bg[blockIdx.y]) return; }
    else if (threadIdx.x == 0) {
        blockcount[blockIdx.y * (size_t)(nblocks + 1) + blockIdx.x] = 0;
        if (blockIdx.x == 0) (const_cast<int *>(need_bg))[blockIdx.y] = 0;
    }
    if (wi >= h * wpr) return;
    bits += blockIdx.y * img_words; parent += blockIdx.y * img_px; outer += blockIdx.y * img_px;
    int y = wi / wpr, k = wi - y * wpr;
    unsigned v = bits[wi];
    if (BG) v = ~v & in_mask(k, w, wpr);
    int base = y * w + k * 32;
    unsigned s = v & ~(v << 1);
    while (s) {
        int b = __ffs(s) - 1; s &= s - 1;
        parent[base + b] = base + b;
        if (BG) outer[base + b] = 0;
    }
}

template <bool DIAG>
__device__ __forceinline__ void merge_class(int *parent, unsigned v, unsigned vl, unsigned vr, unsigned u,
                                            unsigned ul, unsigned ur, int base, int base_up, bool has_up, int k)
{
    // horizontal continuation across the word boundary
    if ((v & 1u) && (vl >> 31)) unite(parent, base, base - 32 + run_start(vl, 31));
    if (!has_up) return;
    unsigned up_m1 = (u << 1) | (ul >> 31), cur_m1 = (v << 1) | (vl >> 31);
    unsigned A = v & u & ~(cur_m1 & up_m1);
    while (A) {
        int x = __ffs(A) - 1; A &= A - 1;
        unite(parent, base + run_start(v, x), base_up + run_start(u, x));
    }
    if (DIAG) {
        unsigned up_p1 = (u >> 1) | (ur << 31), cur_p1 = (v >> 1) | (vr << 31);
        unsigned B = v & ~u & up_m1 & ~cur_m1;
        while (B) {
            int x = __ffs(B) - 1; B &= B - 1;
            int other = (x == 0) ? base_up - 32 + run_start(ul, 31) : base_up + run_start(u, x - 1);
            unite(parent, base + run_start(v, x), other);
        }
        unsigned C = v & ~u & up_p1 & ~cur_p1;
        while (C) {
            int x = __ffs(C) - 1; C &= C - 1;
            int other = (x == 31) ? base_up + 32 : base_up + run_start(u, x + 1);
            unite(parent, base + run_start(v, x), other);
        }
    }
    (void)k;
}

template <bool BG>
__global__ void __launch_bounds__(BG ? 1024 : 256)
ccl_merge_kernel(const unsigned *__restrict__ bits, int *parent, const int *__restrict__ need_bg, int w, int h, int wpr,
                 size_t img_px, size_t img_words)
{
    pdl_entry();
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (BG && !need_bg[blockIdx.y]) return;
    if (wi >= h * wpr) return;
    bits += blockIdx.y * img_words; parent += blockIdx.y * img_px;
    int y = wi / wpr, k = wi - y * wpr;
    unsigned v = bits[wi];
    if (!BG && v == 0) return;                     // no foreground in this word: nothing to link
    unsigned vl = k > 0 ? bits[wi - 1] : 0u, vr = k + 1 < wpr ? bits[wi + 1] : 0u;
    unsigned u = 0, ul = 0, ur = 0;
    bool has_up = y > 0;
    if (has_up) {
        u = bits[wi - wpr];
        ul = k > 0 ? bits[wi - wpr - 1] : 0u;
        ur = k + 1 < wpr ? bits[wi - wpr + 1] : 0u;
    }
    int base = y * w + k * 32, base_up = base - w;
    if (!BG) {
        merge_class<true>(parent, v, vl, vr, u, ul, ur, base, base_up, has_up, k);          // 8-connected
    } else {
        // background = complement inside the image, 4-connected
        unsigned mk = in_mask(k, w, wpr), ml = in_mask(k - 1, w, wpr), mr = in_mask(k + 1, w, wpr);
        merge_class<false>(parent, ~v & mk, ~vl & ml, ~vr & mr, has_up ? (~u & mk) : 0u, has_up ? (~ul & ml) : 0u,
                           has_up ? (~ur & mr) : 0u, base, base_up, has_up, k);
    }
}

template <bool BG>
__global__ void __launch_bounds__(BG ? 1024 : 256)
ccl_flatten_kernel(const unsigned *__restrict__ bits, int *parent, uint8_t *outer, unsigned *__restrict__ rootbits,
                   int *__restrict__ blockcount, const int *__restrict__ need_bg, int w, int h, int wpr, size_t img_px,
                   size_t img_words, int nblocks)
{
    pdl_entry();
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (BG && !need_bg[blockIdx.y]) return;
    if (wi >= h * wpr) return;
    bits += blockIdx.y * img_words; parent += blockIdx.y * img_px; outer += blockIdx.y * img_px;
    rootbits += blockIdx.y * img_words; blockcount += blockIdx.y * (size_t)(nblocks + 1);
    int y = wi / wpr, k = wi - y * wpr;
    unsigned v = bits[wi];
    int base = y * w + k * 32;
    if (!BG) {
        unsigned roots = 0;
        unsigned s = v & ~(v << 1);
        while (s) {
            int b = __ffs(s) - 1; s &= s - 1;
            int r = find_root(parent, base + b);
            parent[base + b] = r;
            if (r == base + b) roots |= 1u << b;
        }
        rootbits[wi] = roots;
        if (roots) atomicAdd(&blockcount[blockIdx.x], __popc(roots));     // roots per 256-word block, for the rank scan
    } else {
        // background runs: flatten, and flag regions that touch the image frame as "outer"
        unsigned vb = ~v & in_mask(k, w, wpr);
        const bool edge_row = (y == 0 || y == h - 1);
        unsigned s = vb & ~(vb << 1);
        while (s) {
            int b = __ffs(s) - 1; s &= s - 1;
            int r = find_root(parent, base + b);
            parent[base + b] = r;
            unsigned rest = ~(vb >> b);
            int len = rest ? __ffs(rest) - 1 : 32 - b;
            bool touches = edge_row || (k == 0 && b == 0) || (k * 32 + b + len - 1 == w - 1);
            if (touches) outer[r] = 1;
        }
    }
}

struct CompRaw { int label, first_index, xmin, ymin, xmax, ymax, area, external; };

// one CTA per image: exclusive scan of the per-block root counts (nblocks <= a few thousand)
__global__ void __launch_bounds__(1024)
ccl_blockscan_kernel(int *blockcount, int nblocks, int *ncomp)
{
    pdl_entry();
    __shared__ int sums[1024];
    int *bc = blockcount + blockIdx.x * (size_t)(nblocks + 1);
    const int tid = threadIdx.x;
    const int chunk = (nblocks + 1023) / 1024;
    const int lo = min(nblocks, tid * chunk), hi = min(nblocks, lo + chunk);
    int acc = 0;
    for (int i = lo; i < hi; i++) acc += bc[i];
    sums[tid] = acc;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // Hillis-Steele inclusive scan
        int v = tid >= off ? sums[tid - off] : 0;
        __syncthreads();
        sums[tid] += v;
        __syncthreads();
    }
    int run = sums[tid] - acc;
    for (int i = lo; i < hi; i++) { int c = bc[i]; bc[i] = run; run += c; }
    if (tid == 1023) { bc[nblocks] = sums[1023]; ncomp[blockIdx.x] = sums[1023]; }
}

// per 256-word block: exclusive rank of every word's roots, component table rows initialised
__global__ void __launch_bounds__(256)
ccl_rank_kernel(const unsigned *__restrict__ rootbits, int *__restrict__ wordrank, const int *__restrict__ blockcount,
                CompRaw *comp, int cap, int w, int h, int wpr, size_t img_words, int nblocks)
{
    pdl_entry();
    __shared__ int wsum[8];
    rootbits += blockIdx.y * img_words; wordrank += blockIdx.y * img_words;
    blockcount += blockIdx.y * (size_t)(nblocks + 1); comp += blockIdx.y * (size_t)cap;
    const int wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int nwords = h * wpr;
    unsigned r = wi < nwords ? rootbits[wi] : 0u;
    int c = __popc(r);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += v;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    int base = blockcount[blockIdx.x];
    for (int i = 0; i < wid; i++) base += wsum[i];
    int run = base + incl - c;
    if (wi >= nwords) return;
    wordrank[wi] = run;
    const int y = wi / wpr, k = wi - y * wpr;
    while (r) {
        int b = __ffs(r) - 1; r &= r - 1;
        if (run < cap) {
            CompRaw cr;
            cr.label = run + 1; cr.first_index = y * w + k * 32 + b;
            cr.xmin = INT_MAX; cr.ymin = INT_MAX; cr.xmax = -1; cr.ymax = -1; cr.area = 0; cr.external = 0;
            comp[run] = cr;
        }
        run++;
    }
}

template <bool LABELS>       // LABELS = false: component table only (no 32-register label row per thread)
__global__ void __launch_bounds__(256)
ccl_label_kernel(const unsigned *__restrict__ bits, const int *__restrict__ parent, const uint8_t *__restrict__ outer,
                 const unsigned *__restrict__ rootbits, const int *__restrict__ wordrank, CompRaw *comp, int cap,
                 int *__restrict__ labels, int w, int h, int wpr, size_t img_px, size_t img_words)
{
    pdl_entry();
    int wi = blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= h * wpr) return;
    bits += blockIdx.y * img_words; parent += blockIdx.y * img_px; outer += blockIdx.y * img_px;
    rootbits += blockIdx.y * img_words; wordrank += blockIdx.y * img_words; comp += blockIdx.y * (size_t)cap;
    if (LABELS) labels += blockIdx.y * img_px;
    int y = wi / wpr, k = wi - y * wpr;
    unsigned v = bits[wi];
    int base = y * w + k * 32;
    int nvalid = min(32, w - k * 32);
    int lab[LABELS ? 32 : 1];
    if (LABELS) {
#pragma unroll
        for (int i = 0; i < 32; i++) lab[i] = 0;
    }
    unsigned s = v & ~(v << 1);
    while (s) {
        int b = __ffs(s) - 1; s &= s - 1;
        int p = base + b;
        int r = parent[p];
        int ry = r / w, rx = r - ry * w;
        int rwi = ry * wpr + (rx >> 5);
        int rank = wordrank[rwi] + __popc(rootbits[rwi] & ((rx & 31) ? (0xffffffffu >> (32 - (rx & 31))) : 0u));
        unsigned rest = ~(v >> b);
        int len = rest ? __ffs(rest) - 1 : 32 - b;
        if (LABELS) {
#pragma unroll
            for (int i = 0; i < 32; i++)
                if (i >= b && i < b + len) lab[i] = rank + 1;
        }
        if (rank < cap) {
            CompRaw *c = comp + rank;
            int x0 = k * 32 + b;
            atomicMin(&c->xmin, x0); atomicMax(&c->xmax, x0 + len - 1);
            atomicMin(&c->ymin, y); atomicMax(&c->ymax, y);
            atomicAdd(&c->area, len);
            if (r == p) c->external = 1;           // provisional; nested components are found by the bg pass
        }
    }
    if (LABELS) {
        int *o = labels + base;
        if (nvalid == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                reinterpret_cast<int4 *>(o)[i] = make_int4(lab[4 * i], lab[4 * i + 1], lab[4 * i + 2], lab[4 * i + 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; i++)
                if (i < nvalid) o[i] = lab[i];
        }
    }
}

// Does any component's bounding box lie strictly inside another's?  (Necessary for a component to sit in a
// hole of another one.)  grid = (16 tiles of 256 components, images); the j-loop runs over shared-memory tiles.
// More than 4096 components: not worth checking, run the bg pass.  need_bg[] was zeroed by ccl_pack_kernel.
__global__ void __launch_bounds__(256)
ccl_nest_kernel(const CompRaw *__restrict__ comp, const int *__restrict__ ncomp, int cap, int *__restrict__ need_bg)
{
    pdl_entry();
    __shared__ int4 box[256];
    const int img = blockIdx.y;
    const int n = ncomp[img];
    if (n > 4096 || n > cap) { if (blockIdx.x == 0 && threadIdx.x == 0) need_bg[img] = 1; return; }
    if ((int)(blockIdx.x * 256) >= n) return;
    const CompRaw *c = comp + (size_t)img * cap;
    const int i = blockIdx.x * 256 + threadIdx.x;
    int x0 = 0, y0 = 0, x1 = 0, y1 = 0;
    if (i < n) { x0 = c[i].xmin; y0 = c[i].ymin; x1 = c[i].xmax; y1 = c[i].ymax; }
    bool inside = false;
    for (int t0 = 0; t0 < n; t0 += 256) {
        const int j = t0 + threadIdx.x;
        box[threadIdx.x] = j < n ? make_int4(c[j].xmin, c[j].ymin, c[j].xmax, c[j].ymax) : make_int4(1 << 30, 1 << 30, -1, -1);
        __syncthreads();
        if (i < n) {
#pragma unroll 8
            for (int q = 0; q < 256; q++) {
                const int4 b = box[q];
                inside |= (b.x < x0) & (b.y < y0) & (b.z > x1) & (b.w > y1);
            }
        }
        __syncthreads();
    }
    if (inside) need_bg[img] = 1;
}

// After the bg pass: a component is external iff the background region left of its first pixel is outer.
__global__ void __launch_bounds__(256)
ccl_resolve_external_kernel(const unsigned *__restrict__ bits, const int *__restrict__ parent,
                            const uint8_t *__restrict__ outer, CompRaw *comp, const int *__restrict__ ncomp, int cap,
                            const int *__restrict__ need_bg, int w, int wpr, size_t img_px, size_t img_words)
{
    pdl_entry();
    const int img = blockIdx.y;
    if (!need_bg[img]) return;
    const int n = min(ncomp[img], cap);
    bits += img * img_words; parent += img * img_px; outer += img * img_px;
    // the component count is only known on the device: a small grid strides over the table
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        CompRaw *c = comp + (size_t)img * cap + i;
        const int r = c->first_index;
        const int ry = r / w, rx = r - ry * w;
        int ext = 1;
        if (rx > 0) {
            const int lk = (rx - 1) >> 5, lb = (rx - 1) & 31;
            const unsigned lv = ~bits[ry * wpr + lk] & in_mask(lk, w, wpr);
            const int node = ry * w + lk * 32 + run_start(lv, lb);
            ext = outer[parent[node]];
        }
        c->external = ext;
    }
}

// cvMoments(ROI, binary=0): pixel-value weighted raw moments, ROI-relative coordinates.
__global__ void __launch_bounds__(256)
rect_moments_kernel(const uint8_t *__restrict__ mask, int w, int h, const int *__restrict__ rects,
                    unsigned long long *out)
{
    pdl_entry();
    const int ri = blockIdx.x;
    const int rx = rects[4 * ri], ry = rects[4 * ri + 1], rw = rects[4 * ri + 2], rh = rects[4 * ri + 3];
    unsigned long long m[6] = {0, 0, 0, 0, 0, 0};
    // each y-slice of blocks takes every gridDim.y-th row
    for (int yy = blockIdx.y; yy < rh; yy += gridDim.y) {
        int y = ry + yy;
        if (y < 0 || y >= h) continue;
        const uint8_t *row = mask + (size_t)y * w;
        unsigned long long r0 = 0, r1 = 0, r2 = 0;       // sum v, sum v*x, sum v*x*x over this row
        for (int xx = threadIdx.x; xx < rw; xx += blockDim.x) {
            int x = rx + xx;
            if (x < 0 || x >= w) continue;
            unsigned long long v = row[x];
            r0 += v; r1 += v * xx; r2 += v * (unsigned long long)xx * xx;
        }
        m[0] += r0; m[1] += r1; m[2] += r0 * yy; m[3] += r2; m[4] += r0 * (unsigned long long)yy * yy; m[5] += r1 * yy;
    }
    __shared__ unsigned long long red[6][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        unsigned long long v = m[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[i][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned long long v = 0;
        for (int i = 0; i < 8; i++) v += red[threadIdx.x][i];
        if (v) atomicAdd(out + 6 * ri + threadIdx.x, v);
    }
}

}  // namespace bgsb

using namespace bgsb;

struct bgsb_ccl {
    int device = 0, max_w = 0, max_h = 0, max_images = 1;
    int w = 0, h = 0, nimages = 0;
    size_t img_px = 0, img_words = 0;      // per-image strides of the work buffers
    int max_blocks = 0;
    unsigned *d_bits = nullptr, *d_rootbits = nullptr;
    int *d_parent = nullptr, *d_wordrank = nullptr, *d_ncomp = nullptr, *d_blockcount = nullptr, *d_need_bg = nullptr;
    int force_bg = 0;                       // 1: always run the background pass (A/B and tests)
    uint8_t *d_outer = nullptr, *d_mask_own = nullptr;
    int32_t *d_labels_own = nullptr;
    CompRaw *d_comp = nullptr;
    int cap = 0;
    const uint8_t *last_mask = nullptr;
    cudaStream_t last_stream = nullptr;
    cudaStream_t own_stream = nullptr;
    bool labelled = false;
    // host round trips of the per-frame path: pinned staging (count + the first PIN_COMPS table rows, or moments) and
    // device buffers for the rectangle queries that live as long as the context (no allocator calls per frame)
    static constexpr int PIN_COMPS = 2048;
    uint8_t *h_pin = nullptr;               // 64 + PIN_COMPS * sizeof(CompRaw) bytes
    int *d_rects = nullptr;
    unsigned long long *d_mom = nullptr;
    int rect_cap = 0;
};

extern "C" {

int bgsb_ccl_create_batch(bgsb_ccl **out, int device, int max_w, int max_h, int max_images)
{
    BGSB_REQUIRE(out, "null out");
    BGSB_REQUIRE(max_w > 0 && max_h > 0 && (long long)max_w * max_h < (1LL << 30), "bad size");
    BGSB_REQUIRE(max_images >= 1 && max_images <= 4096, "max_images out of range");
    BGSB_CUDA(cudaSetDevice(device));
    bgsb_ccl *c = new bgsb_ccl();
    c->device = device; c->max_w = max_w; c->max_h = max_h; c->max_images = max_images;
    c->img_px = (size_t)max_w * max_h;
    c->img_words = (size_t)((max_w + 31) / 32) * max_h;
    c->max_blocks = (int)((c->img_words + 255) / 256);
    // the most 8-connected components an image can hold: isolated pixels on every second row and column
    c->cap = ((max_w + 1) / 2) * ((max_h + 1) / 2) + 64;
    const size_t N = (size_t)max_images;
    cudaError_t e = cudaSuccess;
    auto A = [&](void **p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes); };
    A((void **)&c->d_bits, N * c->img_words * 4);
    A((void **)&c->d_rootbits, N * c->img_words * 4);
    A((void **)&c->d_wordrank, N * c->img_words * 4);
    A((void **)&c->d_parent, N * c->img_px * 4);
    A((void **)&c->d_outer, N * c->img_px);
    A((void **)&c->d_mask_own, c->img_px);
    A((void **)&c->d_comp, N * (size_t)c->cap * sizeof(CompRaw));
    A((void **)&c->d_ncomp, N * sizeof(int));
    A((void **)&c->d_blockcount, N * (size_t)(c->max_blocks + 1) * sizeof(int));
    A((void **)&c->d_need_bg, N * sizeof(int));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        set_error("bgsb_ccl_create: %s", cudaGetErrorString(e));
        bgsb_ccl_destroy(c);
        return BGSB_ERR_CUDA;
    }
    *out = c;
    return BGSB_OK;
}

int bgsb_ccl_create(bgsb_ccl **out, int device, int max_w, int max_h)
{
    return bgsb_ccl_create_batch(out, device, max_w, max_h, 1);
}

void bgsb_ccl_destroy(bgsb_ccl *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_bits); cudaFree(c->d_rootbits); cudaFree(c->d_wordrank); cudaFree(c->d_parent);
    cudaFree(c->d_outer); cudaFree(c->d_mask_own); cudaFree(c->d_comp); cudaFree(c->d_ncomp);
    cudaFree(c->d_blockcount); cudaFree(c->d_need_bg); cudaFree(c->d_labels_own);
    cudaFree(c->d_rects); cudaFree(c->d_mom);
    if (c->h_pin) cudaFreeHost(c->h_pin);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int bgsb_ccl_label_batch_dev(bgsb_ccl *c, const uint8_t *d_masks, int w, int h, int nimages, int zero_border,
                             int32_t *d_labels, void *stream_)
{
    BGSB_REQUIRE(c && d_masks, "null");
    BGSB_REQUIRE(w > 0 && h > 0 && w <= c->max_w && h <= c->max_h, "image larger than the labeller was created for");
    BGSB_REQUIRE(nimages >= 1 && nimages <= c->max_images, "more images than the labeller was created for");
    BGSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    const int wpr = (w + 31) / 32, nwords = wpr * h;
    const int threads = 256, nblocks = (nwords + threads - 1) / threads;
    const size_t ipx = (size_t)w * h, iw = (size_t)nwords;       // dense per-image strides for this geometry
    dim3 grid(nblocks, nimages);
    launch_pdl(ccl_pack_kernel, dim3(grid), dim3(threads), 0, stream, d_masks, c->d_bits, c->d_parent, c->d_blockcount, c->d_need_bg,
               nblocks, w, h, wpr, zero_border, ipx, iw);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_merge_kernel<false>, dim3(grid), dim3(threads), 0, stream, c->d_bits, c->d_parent, c->d_need_bg, w, h, wpr, ipx, iw);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_flatten_kernel<false>, dim3(grid), dim3(threads), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_rootbits,
                                                           c->d_blockcount, c->d_need_bg, w, h, wpr, ipx, iw, nblocks);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_blockscan_kernel, dim3(nimages), dim3(1024), 0, stream, c->d_blockcount, nblocks, c->d_ncomp);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_rank_kernel, dim3(grid), dim3(threads), 0, stream, c->d_rootbits, c->d_wordrank, c->d_blockcount, c->d_comp, c->cap, w, h, wpr,
                                                  iw, nblocks);
    BGSB_LAUNCH_CHECK();
    if (d_labels)
        launch_pdl(ccl_label_kernel<true>, dim3(grid), dim3(threads), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_rootbits,
                   c->d_wordrank, c->d_comp, c->cap, d_labels, w, h, wpr, ipx, iw);
    else
        launch_pdl(ccl_label_kernel<false>, dim3(grid), dim3(threads), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_rootbits,
                   c->d_wordrank, c->d_comp, c->cap, d_labels, w, h, wpr, ipx, iw);
    BGSB_LAUNCH_CHECK();
    // RETR_EXTERNAL: background pass only for images where a bounding box lies strictly inside another
    if (c->force_bg) {
        std::vector<int> ones(nimages, 1);
        BGSB_CUDA(cudaMemcpyAsync(c->d_need_bg, ones.data(), nimages * sizeof(int), cudaMemcpyHostToDevice, stream));
        BGSB_CUDA(cudaStreamSynchronize(stream));
    } else {
        launch_pdl(ccl_nest_kernel, dim3(dim3(16, nimages)), dim3(256), 0, stream, c->d_comp, c->d_ncomp, c->cap, c->d_need_bg);
        BGSB_LAUNCH_CHECK();
    }
    // The background pass exits at once for images that do not need it.  In a batch fat CTAs keep that exit cheap
    // (4x fewer CTAs to retire); a single image keeps 256-thread CTAs so that a pass that IS taken fills the SMs.
    const int bthreads = nimages >= 8 ? 1024 : 256;
    const dim3 bgrid((nwords + bthreads - 1) / bthreads, nimages);
    launch_pdl(ccl_init_kernel<true>, bgrid, dim3(bthreads), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_blockcount, c->d_need_bg,
                                                       w, h, wpr, ipx, iw, nblocks);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_merge_kernel<true>, bgrid, dim3(bthreads), 0, stream, c->d_bits, c->d_parent, c->d_need_bg, w, h, wpr, ipx, iw);
    BGSB_LAUNCH_CHECK();
    launch_pdl(ccl_flatten_kernel<true>, bgrid, dim3(bthreads), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_rootbits,
                                                          c->d_blockcount, c->d_need_bg, w, h, wpr, ipx, iw, nblocks);
    BGSB_LAUNCH_CHECK();
    {
        // at most cap components per image; the kernel exits early past the real count
        const int maxc = std::min(c->cap, ((w + 1) / 2) * ((h + 1) / 2));
        dim3 rgrid(std::min((maxc + 255) / 256, 16), nimages);
        launch_pdl(ccl_resolve_external_kernel, dim3(rgrid), dim3(256), 0, stream, c->d_bits, c->d_parent, c->d_outer, c->d_comp, c->d_ncomp,
                                                               c->cap, c->d_need_bg, w, wpr, ipx, iw);
        BGSB_LAUNCH_CHECK();
    }
    c->w = w; c->h = h; c->nimages = nimages; c->last_mask = d_masks; c->last_stream = stream; c->labelled = true;
    return BGSB_OK;
}

int bgsb_ccl_set_param(bgsb_ccl *c, const char *key, double v)
{
    BGSB_REQUIRE(c && key, "null");
    if (std::string(key) == "forceBackgroundPass") c->force_bg = (v != 0);
    else { set_error("bgsb_ccl_set_param: unknown key '%s'", key); return BGSB_ERR_ARG; }
    return BGSB_OK;
}

int bgsb_ccl_label_dev(bgsb_ccl *c, const uint8_t *d_mask, int w, int h, int zero_border, int32_t *d_labels,
                       void *stream_)
{
    return bgsb_ccl_label_batch_dev(c, d_mask, w, h, 1, zero_border, d_labels, stream_);
}

int bgsb_ccl_components_of(bgsb_ccl *c, int image, bgsb_component *out, int capacity, int *n)
{
    BGSB_REQUIRE(c && n, "null");
    if (!c->labelled) { set_error("bgsb_ccl_components: nothing labelled yet"); return BGSB_ERR_STATE; }
    BGSB_REQUIRE(image >= 0 && image < c->nimages, "image index");
    BGSB_CUDA(cudaSetDevice(c->device));
    // one round trip: the count and the first PIN_COMPS rows come back together through pinned memory
    if (!c->h_pin) BGSB_CUDA(cudaMallocHost(&c->h_pin, 64 + (size_t)bgsb_ccl::PIN_COMPS * sizeof(CompRaw)));
    static_assert(sizeof(CompRaw) == sizeof(bgsb_component), "layout");
    const int want = (out && capacity > 0) ? std::min(std::min(capacity, c->cap), (int)bgsb_ccl::PIN_COMPS) : 0;
    BGSB_CUDA(cudaMemcpyAsync(c->h_pin, c->d_ncomp + image, sizeof(int), cudaMemcpyDeviceToHost, c->last_stream));
    if (want)
        BGSB_CUDA(cudaMemcpyAsync(c->h_pin + 64, c->d_comp + (size_t)image * c->cap, (size_t)want * sizeof(CompRaw),
                                  cudaMemcpyDeviceToHost, c->last_stream));
    BGSB_CUDA(cudaStreamSynchronize(c->last_stream));
    const int cnt = *reinterpret_cast<const int *>(c->h_pin);
    *n = cnt;
    if (cnt > c->cap) { set_error("component table overflow (%d > %d)", cnt, c->cap); return BGSB_ERR_CAPACITY; }
    if (!out || capacity <= 0) return BGSB_OK;
    if (cnt > capacity) { set_error("caller table too small (%d > %d)", cnt, capacity); return BGSB_ERR_CAPACITY; }
    if (cnt == 0) return BGSB_OK;
    memcpy(out, c->h_pin + 64, (size_t)std::min(cnt, want) * sizeof(CompRaw));
    if (cnt > want)
        BGSB_CUDA(cudaMemcpy(out + want, c->d_comp + (size_t)image * c->cap + want, (size_t)(cnt - want) * sizeof(CompRaw),
                             cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; i++) {       // (xmin,ymin,xmax,ymax) -> (x,y,w,h)
        out[i].w = out[i].w - out[i].x + 1;
        out[i].h = out[i].h - out[i].y + 1;
    }
    return BGSB_OK;
}

int bgsb_ccl_components(bgsb_ccl *c, bgsb_component *out, int capacity, int *n)
{
    return bgsb_ccl_components_of(c, 0, out, capacity, n);
}

int bgsb_ccl_rect_moments(bgsb_ccl *c, const int32_t *rects, int nrects, uint64_t *out)
{
    BGSB_REQUIRE(c && out && (rects || nrects == 0), "null");
    if (!c->labelled) { set_error("bgsb_ccl_rect_moments: nothing labelled yet"); return BGSB_ERR_STATE; }
    if (nrects == 0) return BGSB_OK;
    BGSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->last_stream;
    if (nrects > c->rect_cap) {                              // grows rarely; no allocator calls on the per-frame path
        cudaFree(c->d_rects); cudaFree(c->d_mom); c->d_rects = nullptr; c->d_mom = nullptr; c->rect_cap = 0;
        const int cap = std::max(1024, nrects * 2);
        BGSB_CUDA(cudaMalloc(&c->d_rects, (size_t)cap * 16));
        BGSB_CUDA(cudaMalloc(&c->d_mom, (size_t)cap * 48));
        c->rect_cap = cap;
    }
    if (!c->h_pin) BGSB_CUDA(cudaMallocHost(&c->h_pin, 64 + (size_t)bgsb_ccl::PIN_COMPS * sizeof(CompRaw)));
    const size_t pin_bytes = (size_t)bgsb_ccl::PIN_COMPS * sizeof(CompRaw);
    const bool staged = (size_t)nrects * 48 <= pin_bytes;     // moments come back through pinned memory when they fit
    BGSB_CUDA(cudaMemcpyAsync(c->d_rects, rects, (size_t)nrects * 16, cudaMemcpyHostToDevice, st));
    BGSB_CUDA(cudaMemsetAsync(c->d_mom, 0, (size_t)nrects * 48, st));
    dim3 grid(nrects, 16);
    launch_pdl(rect_moments_kernel, dim3(grid), dim3(256), 0, st, c->last_mask, c->w, c->h, c->d_rects, c->d_mom);
    BGSB_LAUNCH_CHECK();
    BGSB_CUDA(cudaMemcpyAsync(staged ? (void *)(c->h_pin + 64) : (void *)out, c->d_mom, (size_t)nrects * 48,
                              cudaMemcpyDeviceToHost, st));
    BGSB_CUDA(cudaStreamSynchronize(st));
    if (staged) memcpy(out, c->h_pin + 64, (size_t)nrects * 48);
    return BGSB_OK;
}

int bgsb_ccl_label(bgsb_ccl *c, const uint8_t *mask, int w, int h, size_t stride, int zero_border, int32_t *labels,
                   bgsb_component *out, int capacity, int *n)
{
    BGSB_REQUIRE(c && mask && n, "null");
    BGSB_REQUIRE(stride >= (size_t)w, "stride smaller than a row");
    BGSB_REQUIRE(w > 0 && h > 0 && w <= c->max_w && h <= c->max_h, "image larger than the labeller was created for");
    BGSB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->own_stream;
    BGSB_CUDA(cudaMemcpy2DAsync(c->d_mask_own, w, mask, stride, w, h, cudaMemcpyHostToDevice, st));
    if (labels && !c->d_labels_own) BGSB_CUDA(cudaMalloc(&c->d_labels_own, c->img_px * 4));
    int rc = bgsb_ccl_label_dev(c, c->d_mask_own, w, h, zero_border, labels ? c->d_labels_own : nullptr, st);
    if (rc) return rc;
    if (labels) BGSB_CUDA(cudaMemcpyAsync(labels, c->d_labels_own, (size_t)w * h * 4, cudaMemcpyDeviceToHost, st));
    return bgsb_ccl_components(c, out, capacity, n);
}

}  // extern "C"
