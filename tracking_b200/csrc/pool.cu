// Stream pool: N camera streams of one geometry over the GPUs of this process (BASELINE configs 4 / 5).
//
// The reference runs ONE stream per process: its main loop (ustc_src/trackingMain.cpp:161-166) reads a frame and
// pushes it through USTC_BGS::Process, the mask clean-up and the blob detector before it reads the next one;
// FrameProcessor::process (FrameProcessor.cpp:176-195) does the same for the plugin family.  The pool is that loop
// for many cameras: stream s lives on devices[s % ndevices]; every GPU has
//   * one host worker thread that owns the GPU's context work (nothing is enqueued for a GPU from any other thread, so
//     the launch cost of G GPUs is paid in parallel, not G times in a row),
//   * one stream-group pipeline (bgsb_pipeline: plugin -> erode/dilate -> labelling, one launch sequence for all the
//     GPU's streams),
//   * rings of `ring` slots of page-locked host memory for the frames coming in and the masks / component tables
//     going out, and matching device rings, on three CUDA streams (upload / kernels / download) chained by events, so
//     that the upload of slot k+1 overlaps the kernels of slot k and the download of slot k-1.
// Streams never exchange data: no collective, nothing crosses NVLink.
#include <string.h>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "kernels.h"

using namespace bgsb;

namespace {

constexpr int POOL_RING_MAX = 8;

struct Job { int slot; int want_masks; uint64_t seq; };

struct GpuGroup {
    int device = 0;
    std::vector<int> streams;                 // global stream ids, in group order
    bgsb_pipeline *pipe = nullptr;
    uint8_t *h_in[POOL_RING_MAX] = {}, *d_in[POOL_RING_MAX] = {};
    uint8_t *h_mask[POOL_RING_MAX] = {}, *d_mask[POOL_RING_MAX] = {};
    int32_t *h_tab[POOL_RING_MAX] = {}, *d_tab[POOL_RING_MAX] = {};
    cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_up[POOL_RING_MAX] = {}, ev_k[POOL_RING_MAX] = {}, ev_done[POOL_RING_MAX] = {};
    bool used[POOL_RING_MAX] = {};
    // worker
    std::thread th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<Job> jobs;
    bool stop = false;
    uint64_t enq_seq[POOL_RING_MAX] = {};      // sequence number of the last job the worker has fully enqueued per slot
    int rc[POOL_RING_MAX] = {};
    int valid[POOL_RING_MAX] = {};
    int has_mask[POOL_RING_MAX] = {};
    std::string err[POOL_RING_MAX];
};

}  // namespace

struct bgsb_pool {
    int algo = 0, nstreams = 0, w = 0, h = 0, ring = 2, table_rows = 256;
    std::vector<GpuGroup *> groups;
    std::vector<int> group_of, index_in_group;      // per global stream
    uint64_t seq = 0;
    uint64_t want_seq[POOL_RING_MAX] = {};           // sequence number of the last submit per slot
    size_t frame_bytes = 0, mask_bytes = 0, tab_ints = 0;
};

static void worker_main(bgsb_pool *P, GpuGroup *G)
{
    cudaSetDevice(G->device);
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> lk(G->mu);
            G->cv.wait(lk, [&] { return G->stop || !G->jobs.empty(); });
            if (G->jobs.empty()) return;           // stop requested and nothing left to enqueue
            j = G->jobs.front();
            G->jobs.pop_front();
        }
        const int k = j.slot;
        const size_t S = G->streams.size();
        int rc = BGSB_OK, valid = 0;
        std::string err;
        auto fail = [&](cudaError_t e, const char *what) {
            if (e != cudaSuccess && rc == BGSB_OK) { rc = BGSB_ERR_CUDA; err = std::string(what) + ": " + cudaGetErrorString(e); (void)cudaGetLastError(); }
            return e != cudaSuccess;
        };
        do {
            // the device input slot was last read by the kernels of its previous use; its outputs by that use's downloads
            if (G->used[k]) {
                if (fail(cudaStreamWaitEvent(G->s_h2d, G->ev_k[k], 0), "wait(kernels of the slot's previous use)")) break;
                if (fail(cudaStreamWaitEvent(G->s_k, G->ev_done[k], 0), "wait(downloads of the slot's previous use)")) break;
            }
            if (fail(cudaMemcpyAsync(G->d_in[k], G->h_in[k], S * P->frame_bytes, cudaMemcpyHostToDevice, G->s_h2d), "upload")) break;
            if (fail(cudaEventRecord(G->ev_up[k], G->s_h2d), "record")) break;
            if (fail(cudaStreamWaitEvent(G->s_k, G->ev_up[k], 0), "wait(upload)")) break;
            int bgv = 0;
            rc = bgsb_pipeline_process_dev(G->pipe, G->d_in[k], P->w, P->h, j.want_masks ? G->d_mask[k] : nullptr, nullptr, nullptr,
                                           &valid, &bgv, G->s_k);
            if (rc) { err = bgsb_last_error(); break; }
            if (fail(cudaEventRecord(G->ev_k[k], G->s_k), "record")) break;
            if (fail(cudaStreamWaitEvent(G->s_d2h, G->ev_k[k], 0), "wait(kernels)")) break;
            if (valid) {
                // clean-up and labelling run on the pipeline's own stream beside the NEXT frame set's plugin kernel; only
                // the download stream waits for them (masks and tables), the kernel stream goes straight on
                rc = bgsb_pipeline_join_dev(G->pipe, G->s_d2h);
                if (!rc) rc = bgsb_pipeline_tables_dev(G->pipe, G->d_tab[k], P->table_rows, G->s_d2h);
                if (rc) { err = bgsb_last_error(); break; }
            }
            if (valid) {
                if (j.want_masks && fail(cudaMemcpyAsync(G->h_mask[k], G->d_mask[k], S * P->mask_bytes, cudaMemcpyDeviceToHost, G->s_d2h), "download masks")) break;
                if (fail(cudaMemcpyAsync(G->h_tab[k], G->d_tab[k], S * P->tab_ints * sizeof(int32_t), cudaMemcpyDeviceToHost, G->s_d2h), "download tables")) break;
            }
            if (fail(cudaEventRecord(G->ev_done[k], G->s_d2h), "record")) break;
            G->used[k] = true;
        } while (0);
        {
            std::lock_guard<std::mutex> lk(G->mu);
            G->rc[k] = rc; G->err[k] = err; G->valid[k] = valid; G->has_mask[k] = valid && j.want_masks;
            G->enq_seq[k] = j.seq;
        }
        G->cv_done.notify_all();
    }
}

static void pool_free(bgsb_pool *P)
{
    for (GpuGroup *G : P->groups) {
        if (G->th.joinable()) {
            { std::lock_guard<std::mutex> lk(G->mu); G->stop = true; }
            G->cv.notify_all();
            G->th.join();
        }
        cudaSetDevice(G->device);
        cudaDeviceSynchronize();
        if (G->pipe) bgsb_pipeline_destroy(G->pipe);
        for (int k = 0; k < POOL_RING_MAX; k++) {
            if (G->h_in[k]) cudaFreeHost(G->h_in[k]);
            if (G->h_mask[k]) cudaFreeHost(G->h_mask[k]);
            if (G->h_tab[k]) cudaFreeHost(G->h_tab[k]);
            cudaFree(G->d_in[k]); cudaFree(G->d_mask[k]); cudaFree(G->d_tab[k]);
            if (G->ev_up[k]) cudaEventDestroy(G->ev_up[k]);
            if (G->ev_k[k]) cudaEventDestroy(G->ev_k[k]);
            if (G->ev_done[k]) cudaEventDestroy(G->ev_done[k]);
        }
        if (G->s_h2d) cudaStreamDestroy(G->s_h2d);
        if (G->s_k) cudaStreamDestroy(G->s_k);
        if (G->s_d2h) cudaStreamDestroy(G->s_d2h);
        delete G;
    }
    P->groups.clear();
}

extern "C" {

int bgsb_pool_create(bgsb_pool **out, int algo, int nstreams, const int *devices, int ndevices, int w, int h, int ring)
{
    BGSB_REQUIRE(out && devices, "null");
    BGSB_REQUIRE(nstreams >= 1 && ndevices >= 1 && ndevices <= 64, "nstreams >= 1, 1 <= ndevices <= 64");
    BGSB_REQUIRE(w > 0 && h > 0 && (long long)w * h < (1LL << 27), "bad geometry");
    BGSB_REQUIRE(ring >= 1 && ring <= POOL_RING_MAX, "ring in [1,8]");
    bgsb_pool *P = new bgsb_pool();
    P->algo = algo; P->nstreams = nstreams; P->w = w; P->h = h; P->ring = ring;
    P->frame_bytes = (size_t)w * h * 3; P->mask_bytes = (size_t)w * h;
    P->tab_ints = (size_t)(P->table_rows + 1) * 8;
    P->group_of.assign(nstreams, 0); P->index_in_group.assign(nstreams, 0);
    const int ng = std::min(ndevices, nstreams);
    int rc = BGSB_OK;
    for (int g = 0; g < ng && rc == BGSB_OK; g++) {
        GpuGroup *G = new GpuGroup();
        P->groups.push_back(G);
        G->device = devices[g];
        for (int s = g; s < nstreams; s += ng) {           // stream s -> GPU s mod G
            P->group_of[s] = g; P->index_in_group[s] = (int)G->streams.size();
            G->streams.push_back(s);
        }
        const size_t S = G->streams.size();
        cudaError_t e = cudaSetDevice(G->device);
        if (e == cudaSuccess) rc = bgsb_pipeline_create(&G->pipe, algo, G->device, (int)S);
        if (rc) break;
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G->s_h2d, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G->s_k, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&G->s_d2h, cudaStreamNonBlocking);
        for (int k = 0; k < ring && e == cudaSuccess; k++) {
            e = cudaHostAlloc((void **)&G->h_in[k], S * P->frame_bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaHostAlloc((void **)&G->h_mask[k], S * P->mask_bytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaHostAlloc((void **)&G->h_tab[k], S * P->tab_ints * sizeof(int32_t), cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaMalloc(&G->d_in[k], S * P->frame_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&G->d_mask[k], S * P->mask_bytes);
            if (e == cudaSuccess) e = cudaMalloc(&G->d_tab[k], S * P->tab_ints * sizeof(int32_t));
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G->ev_up[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G->ev_k[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&G->ev_done[k], cudaEventDisableTiming);
        }
        if (e != cudaSuccess) {
            set_error("bgsb_pool_create (device %d, %zu streams): %s", G->device, S, cudaGetErrorString(e));
            (void)cudaGetLastError();
            rc = BGSB_ERR_CUDA;
        }
    }
    if (rc) { pool_free(P); delete P; return rc; }
    for (GpuGroup *G : P->groups) G->th = std::thread(worker_main, P, G);
    *out = P;
    return BGSB_OK;
}

void bgsb_pool_destroy(bgsb_pool *P)
{
    if (!P) return;
    pool_free(P);
    delete P;
}

int bgsb_pool_set_param(bgsb_pool *P, const char *key, double v)
{
    BGSB_REQUIRE(P && key, "null");
    for (GpuGroup *G : P->groups) {
        int rc = bgsb_pipeline_set_param(G->pipe, key, v);
        if (rc) return rc;
    }
    return BGSB_OK;
}

int bgsb_pool_set_morph(bgsb_pool *P, const int *ops, int nops)
{
    BGSB_REQUIRE(P, "null");
    for (GpuGroup *G : P->groups) {
        int rc = bgsb_pipeline_set_morph(G->pipe, ops, nops);
        if (rc) return rc;
    }
    return BGSB_OK;
}

int bgsb_pool_device_of(bgsb_pool *P, int stream, int *device)
{
    BGSB_REQUIRE(P && device && stream >= 0 && stream < P->nstreams, "stream index");
    *device = P->groups[P->group_of[stream]]->device;
    return BGSB_OK;
}

uint8_t *bgsb_pool_frame_buffer(bgsb_pool *P, int stream, int slot)
{
    if (!P || stream < 0 || stream >= P->nstreams || slot < 0 || slot >= P->ring) return nullptr;
    GpuGroup *G = P->groups[P->group_of[stream]];
    return G->h_in[slot] + (size_t)P->index_in_group[stream] * P->frame_bytes;
}

int bgsb_pool_submit(bgsb_pool *P, int slot, int want_masks)
{
    BGSB_REQUIRE(P && slot >= 0 && slot < P->ring, "slot index");
    const uint64_t seq = ++P->seq;
    P->want_seq[slot] = seq;
    for (GpuGroup *G : P->groups) {
        { std::lock_guard<std::mutex> lk(G->mu); G->jobs.push_back(Job{slot, want_masks, seq}); }
        G->cv.notify_one();
    }
    return BGSB_OK;
}

int bgsb_pool_wait(bgsb_pool *P, int slot, int *valid)
{
    BGSB_REQUIRE(P && slot >= 0 && slot < P->ring, "slot index");
    if (valid) *valid = 0;
    if (P->want_seq[slot] == 0) { set_error("bgsb_pool_wait: nothing was submitted for slot %d", slot); return BGSB_ERR_STATE; }
    int all_valid = 1;
    for (GpuGroup *G : P->groups) {
        {
            std::unique_lock<std::mutex> lk(G->mu);
            G->cv_done.wait(lk, [&] { return G->enq_seq[slot] >= P->want_seq[slot]; });
            if (G->rc[slot]) { set_error("bgsb_pool (device %d): %s", G->device, G->err[slot].c_str()); return G->rc[slot]; }
            all_valid &= G->valid[slot];
        }
        BGSB_CUDA(cudaEventSynchronize(G->ev_done[slot]));
    }
    if (valid) *valid = all_valid;
    return BGSB_OK;
}

const uint8_t *bgsb_pool_mask(bgsb_pool *P, int stream, int slot)
{
    if (!P || stream < 0 || stream >= P->nstreams || slot < 0 || slot >= P->ring) return nullptr;
    GpuGroup *G = P->groups[P->group_of[stream]];
    if (!G->has_mask[slot]) return nullptr;
    return G->h_mask[slot] + (size_t)P->index_in_group[stream] * P->mask_bytes;
}

int bgsb_pool_components(bgsb_pool *P, int stream, int slot, bgsb_component *out, int capacity, int *n)
{
    BGSB_REQUIRE(P && n && stream >= 0 && stream < P->nstreams && slot >= 0 && slot < P->ring, "bad args");
    GpuGroup *G = P->groups[P->group_of[stream]];
    if (!G->valid[slot]) { *n = 0; set_error("bgsb_pool_components: the slot holds no mask (plugin warm-up frame)"); return BGSB_ERR_STATE; }
    const int32_t *t = G->h_tab[slot] + (size_t)P->index_in_group[stream] * P->tab_ints;
    const int cnt = t[0];
    *n = cnt;
    if (!out || capacity <= 0) return BGSB_OK;
    if (cnt > capacity) { set_error("caller table too small (%d > %d)", cnt, capacity); return BGSB_ERR_CAPACITY; }
    if (cnt > P->table_rows) {
        set_error("stream %d has %d components, the pool keeps the first %d per frame (noisy mask?)", stream, cnt, P->table_rows);
        memcpy(out, t + 8, (size_t)P->table_rows * sizeof(bgsb_component));
        return BGSB_ERR_CAPACITY;
    }
    memcpy(out, t + 8, (size_t)cnt * sizeof(bgsb_component));
    return BGSB_OK;
}

int bgsb_pool_info(bgsb_pool *P, int *ngroups, int *ring, int *table_rows)
{
    BGSB_REQUIRE(P, "null");
    if (ngroups) *ngroups = (int)P->groups.size();
    if (ring) *ring = P->ring;
    if (table_rows) *table_rows = P->table_rows;
    return BGSB_OK;
}

}  // extern "C"
