// Host side of libmrgingham_b200.so: the C ABI of include/mrgingham_b200.h over the kernels.
// No CPU implementation of anything lives here: if CUDA is unusable every entry point fails loudly.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <vector>
#include <algorithm>
#include <atomic>
#include <thread>
#include <condition_variable>

#include "../../include/mrgingham_b200.h"
#include "kernels.cuh"
#include "find_grid.hh"
#include <chrono>

using namespace mrgb200;

#define API extern "C" __attribute__((visibility("default")))

#define MSG(fmt, ...) fprintf(stderr, "%s:%d in %s(): " fmt " Sorry.\n", __FILE__, __LINE__, __func__, ##__VA_ARGS__)

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            MSG("CUDA failure '%s' in " #expr ".", cudaGetErrorString(_e));              \
            return -1;                                                                   \
        }                                                                                \
    } while (0)

static int next_pow2(long long v) { int p = 1; while (p < v) p <<= 1; return p; }
static int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int cv_round_div(int n, int d)  // cvRound(n/d): round half to even, d a power of two
{
    int q = n / d, r = n % d;
    if (2*r > d || (2*r == d && (q & 1))) q++;
    return q;
}

namespace
{
struct DeviceBuffer
{
    void*  p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want)
    {
        if (want <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr; bytes = 0;
        CUDA_TRY(cudaMalloc(&p, want));
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
};
struct PinnedBuffer
{
    void*  p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want)
    {
        if (want <= bytes) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; bytes = 0;
        CUDA_TRY(cudaMallocHost(&p, want));
        bytes = want;
        return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; bytes = 0; }
};

struct KernelTimer
{
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> spans;
    std::vector<cudaEvent_t> pool;
    float ms = 0; int launches = 0;
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e; cudaEventCreate(&e); return e;
    }
    void reset() { for (auto& s : spans) { pool.push_back(s.first); pool.push_back(s.second); } spans.clear(); ms = 0; launches = 0; }
    void resolve()
    {
        for (auto& s : spans) { float t = 0; if (cudaEventElapsedTime(&t, s.first, s.second) == cudaSuccess) ms += t; pool.push_back(s.first); pool.push_back(s.second); }
        spans.clear();
    }
    void release() { reset(); for (auto e : pool) cudaEventDestroy(e); pool.clear(); }
};

// Makes `dev` the calling thread's current device for the lifetime of the guard and puts the caller's own
// device back afterwards: no entry point of this library changes the caller's current device (the reference,
// CPU code, has no such side effect).
struct DeviceGuard
{
    int prev = -1; bool ok = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = prev == dev || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;      // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(det)                                                                \
    DeviceGuard _dg((det)->device);                                                      \
    if (!_dg.ok) { MSG("Could not select CUDA device %d.", (det)->device); return -1; }
}

struct mrg_b200_detector
{
    mrg_b200_detector_config cfg;
    int          device = 0;
    cudaStream_t own_stream = nullptr;
    std::mutex   mtx;

    // Per-chunk scratch, double-buffered: chunk c uses slot c&1, so the clustering kernel of chunk
    // c (aux stream) and the host->device copy of chunk c+1 (copy stream) overlap the ChESS kernel
    // of chunk c+1 (main stream).
    struct Slot
    {
        DeviceBuffer stage, blurred, equalized, pre_scratch, level_img, cand, counts, table, dfs, records, xy, outcounts;
        cudaEvent_t staged = nullptr, k1done = nullptr, k2done = nullptr;
        bool used = false;
    } slot[2];
    cudaStream_t aux_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t  ev_fork = nullptr;
    // scratch for single frames whose candidate list overflowed the default capacity
    DeviceBuffer big_cand, big_table, big_dfs, big_records;
    // refinement
    DeviceBuffer pts, lvls;
    // results of the batch in flight
    PinnedBuffer h_xy, h_counts, h_candcounts;
    // detectors of the one-image pool: pageable caller memory goes through this pinned buffer (a plain memcpy, then
    // one DMA on the detector's own stream), so concurrent callers do not queue on the driver's pageable-copy path
    bool         stage_via_pinned = false;
    PinnedBuffer h_stage;
    cudaEvent_t  h_stage_free = nullptr;      // recorded after the last copy out of h_stage

    bool profiling = false;
    KernelTimer timers[3];
    int k2_smem_cands = kClusterSmemCands;   // adapted to the candidate counts of the previous batch (collect_locked)
    int k2_rows = 0, k2_cols = 0, k2_level = -1;   // ... of this geometry
    static constexpr int kBlobDepth = 4;       // chunks of the blob path kept in flight
    BlobWorkspace* blobs[kBlobDepth] = {};
    Slot blob_slot[kBlobDepth];                // their staged frames (host input)
    cudaStream_t blob_stream[kBlobDepth] = {};
    float blob_ms = 0;
    // board finder: the frames of the chunk being worked on, kept on the device across the level loop
    std::mutex   boards_mtx;
    DeviceBuffer boards_frames, boards_gather;
    mrg_b200_detector* boards_helper = nullptr;   // second detector of the board finder (find_boards): chunks alternate between the two
    DeviceBuffer dbg_xyd;                      // mrg_b200_debug_dump_corners(): the corners as doubles
    DeviceBuffer mixed_stage, mixed_srcs;     // mrg_b200_find_corners_mixed_batch(): one size group, contiguous; its sources

    struct Pending
    {
        bool active = false;
        const uint8_t* images; int on_device, nframes, rows, cols, level, mp; size_t pitch, fstride;
        cudaStream_t stream;
    } pending;
};

namespace
{
struct Launch
{
    mrg_b200_detector* det; int which; cudaStream_t stream; cudaEvent_t e0 = nullptr;
    Launch(mrg_b200_detector* d, int w, cudaStream_t s) : det(d), which(w), stream(s)
    {
        det->timers[which].launches++;
        if (det->profiling) { e0 = det->timers[which].get(); cudaEventRecord(e0, stream); }
    }
    ~Launch()
    {
        if (det->profiling) { cudaEvent_t e1 = det->timers[which].get(); cudaEventRecord(e1, stream); det->timers[which].spans.push_back({e0, e1}); }
    }
};

// batch launch sizes: see enqueue_locked()
constexpr int kTaperMinFrames = 128;

int chess_sparse(mrg_b200_detector* det, const FrameSet& fs, cand_t* cand, uint32_t* counts, int cap, cudaStream_t stream)
{
    Launch l(det, 0, stream);
    if (det->cfg.kernel_variant == 1) { CUDA_TRY(launch_chess_sparse_simple(fs, cand, counts, cap, stream)); return 0; }
    if (det->cfg.kernel_variant != 2)
    {
        bool launched = false;
        CUDA_TRY(launch_chess_sparse_cascade(fs, cand, counts, cap, stream, &launched));
        if (launched) return 0;
    }
    CUDA_TRY(launch_chess_sparse_tiled(fs, cand, counts, cap, stream));
    return 0;
}

// geometry of a pyramid level of a rows x cols frame
struct LevelGeom { int w, h, pitch; size_t frame_bytes; };
LevelGeom level_geom(int rows, int cols, int level)
{
    LevelGeom g;
    g.w = cv_round_div(cols, 1 << level);
    g.h = cv_round_div(rows, 1 << level);
    g.pitch = round_up(std::max(g.w, 1), 16);
    g.frame_bytes = (size_t)g.pitch * std::max(g.h, 1);
    return g;
}

// Puts `n` frames on the device (if they are not there already; the copy goes on `cstream` and
// `stream` is made to wait for it) and, for level > 0, builds the level image on `stream`.
// On return `out` describes the image the detector kernels must read.
int stage_frames(mrg_b200_detector* det, mrg_b200_detector::Slot& S, const uint8_t* images, int on_device, int n, int rows, int cols,
                 size_t pitch, size_t fstride, int level, cudaStream_t stream, cudaStream_t cstream, FrameSet* out,
                 bool preprocess = false)
{
    FrameSet src;
    if (on_device)
    {
        src.base = images; src.frame_stride = fstride; src.pitch = (int)pitch;
    }
    else
    {
        const int spitch = round_up(cols, 16);
        const size_t sframe = (size_t)spitch * rows;
        if (S.stage.ensure(sframe * n)) return -1;
        if (det->stage_via_pinned)
        {
            if (det->h_stage.ensure(sframe * n)) return -1;
            if (!det->h_stage_free) CUDA_TRY(cudaEventCreateWithFlags(&det->h_stage_free, cudaEventDisableTiming));
            else                    CUDA_TRY(cudaEventSynchronize(det->h_stage_free));
            uint8_t* hp = (uint8_t*)det->h_stage.p;
            for (int i = 0; i < n; i++)
                if (pitch == (size_t)spitch) memcpy(hp + i * sframe, images + i * fstride, sframe);
                else for (int y = 0; y < rows; y++) memcpy(hp + i * sframe + (size_t)y * spitch, images + i * fstride + (size_t)y * pitch, cols);
            CUDA_TRY(cudaMemcpyAsync(S.stage.p, hp, sframe * n, cudaMemcpyHostToDevice, cstream));
            CUDA_TRY(cudaEventRecord(det->h_stage_free, cstream));
        }
        else if (fstride == pitch * (size_t)rows && pitch == (size_t)spitch)
            // dense rows on both sides: one linear copy
            CUDA_TRY(cudaMemcpyAsync(S.stage.p, images, sframe * n, cudaMemcpyHostToDevice, cstream));
        else if (fstride == pitch * (size_t)rows)
            CUDA_TRY(cudaMemcpy2DAsync(S.stage.p, spitch, images, pitch, cols, (size_t)rows * n, cudaMemcpyHostToDevice, cstream));
        else
            for (int i = 0; i < n; i++)
                CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)S.stage.p + i * sframe, spitch, images + i * fstride, pitch, cols, rows,
                                           cudaMemcpyHostToDevice, cstream));
        if (cstream != stream)
        {
            CUDA_TRY(cudaEventRecord(S.staged, cstream));
            CUDA_TRY(cudaStreamWaitEvent(stream, S.staged, 0));
        }
        src.base = (const uint8_t*)S.stage.p; src.frame_stride = sframe; src.pitch = spitch;
    }
    src.w = cols; src.h = rows; src.nframes = n;
    if (preprocess && det->cfg.clahe)
    {
        // the reference CLI's --clahe: normalize to [0,255], then CLAHE(clipLimit 8) (mrgingham-from-image.cc:71-80)
        const int epitch = round_up(cols, 16);
        const size_t eframe = (size_t)epitch * rows;
        if (S.equalized.ensure(eframe * n) || S.pre_scratch.ensure(clahe_scratch_bytes(n))) return -1;
        CUDA_TRY(launch_normalize_clahe(src, true, 8.0, (uint8_t*)S.equalized.p, epitch, eframe, S.pre_scratch.p, stream));
        src.base = (const uint8_t*)S.equalized.p; src.frame_stride = eframe; src.pitch = epitch;
    }
    if (preprocess && det->cfg.blur_radius > 0)
    {
        // the reference CLI's default preprocessing, on the device: the detector reads the blurred copy
        const int bpitch = round_up(cols, 16);
        const size_t bframe = (size_t)bpitch * rows;
        if (S.blurred.ensure(bframe * n)) return -1;
        CUDA_TRY(launch_box_blur(src, det->cfg.blur_radius, (uint8_t*)S.blurred.p, bpitch, bframe, stream));
        src.base = (const uint8_t*)S.blurred.p; src.frame_stride = bframe; src.pitch = bpitch;
    }
    if (level == 0) { *out = src; return 0; }

    const LevelGeom g = level_geom(rows, cols, level);
    if (S.level_img.ensure(g.frame_bytes * n)) return -1;
    {
        Launch l(det, 2, stream);
        CUDA_TRY(launch_pyramid(src, level, (uint8_t*)S.level_img.p, g.pitch, g.frame_bytes, g.w, g.h, stream));
    }
    out->base = (const uint8_t*)S.level_img.p; out->frame_stride = g.frame_bytes; out->pitch = g.pitch;
    out->w = g.w; out->h = g.h; out->nframes = n;
    return 0;
}

// mp = per-frame output capacity of the call being served. It is an argument everywhere below and never written
// back into det->cfg: the public batch calls pass the configured max_points (what the caller sized xy_out by),
// the board finder and the one-image entry points grow a private value when a frame has more corners.
int ensure_chunk_scratch(mrg_b200_detector* det, mrg_b200_detector::Slot& S, int n, int mp_)
{
    const size_t cap = det->cfg.candidate_capacity, mp = (size_t)mp_;
    if (S.cand.ensure(sizeof(cand_t) * cap * n)) return -1;
    if (S.counts.ensure(sizeof(uint32_t) * n)) return -1;
    if (S.table.ensure(sizeof(uint32_t) * 2 * cap * n)) return -1;
    if (S.dfs.ensure(sizeof(uint32_t) * cap * n)) return -1;
    if (S.records.ensure(cluster_record_bytes() * 2 * mp * n)) return -1;
    if (S.xy.ensure(sizeof(int32_t) * 2 * mp * n)) return -1;
    if (S.outcounts.ensure(sizeof(int32_t) * n)) return -1;
    return 0;
}

// One frame whose candidate list (or component list) overflowed the chunk scratch: run it again
// on the GPU with worst-case capacities.
int rerun_big(mrg_b200_detector* det, const uint8_t* image, int on_device, int rows, int cols, size_t pitch,
              int level, int mp, cudaStream_t stream, int32_t* xy_out, int32_t* count_out, int32_t* candcount_out)
{
    mrg_b200_detector::Slot& S = det->slot[0];
    FrameSet fs;
    if (stage_frames(det, S, image, on_device, 1, rows, cols, pitch, pitch * rows, level, stream, stream, &fs, true)) return -1;
    const int cap = next_pow2((long long)fs.w * fs.h);
    const int reccap = cap / 2 + 1;
    if (det->big_cand.ensure(sizeof(cand_t) * cap)) return -1;
    if (det->big_table.ensure(sizeof(uint32_t) * 2 * cap)) return -1;
    if (det->big_dfs.ensure(sizeof(uint32_t) * cap)) return -1;
    if (det->big_records.ensure(cluster_record_bytes() * reccap)) return -1;
    if (ensure_chunk_scratch(det, S, 1, mp)) return -1;
    CUDA_TRY(cudaMemsetAsync(S.counts.p, 0, sizeof(uint32_t), stream));
    if (chess_sparse(det, fs, (cand_t*)det->big_cand.p, (uint32_t*)S.counts.p, cap, stream)) return -1;
    ClusterParams p; p.smem_cands = det->k2_smem_cands; p.level = level; p.cand_capacity = cap; p.max_points = mp; p.record_capacity = reccap; p.records = det->big_records.p;
    {
        Launch l(det, 1, stream);
        CUDA_TRY(launch_cluster_find(fs, p, (cand_t*)det->big_cand.p, (uint32_t*)S.counts.p, (uint32_t*)det->big_table.p,
                                     (uint32_t*)det->big_dfs.p, (int32_t*)S.xy.p, nullptr, (int32_t*)S.outcounts.p, stream));
    }
    CUDA_TRY(cudaMemcpyAsync(xy_out, S.xy.p, sizeof(int32_t) * 2 * mp, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(count_out, S.outcounts.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaMemcpyAsync(candcount_out, S.counts.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (*count_out < 0) { MSG("Frame still overflows at worst-case capacity; this is a bug."); return -1; }
    return 0;
}

int enqueue_locked(mrg_b200_detector* det, const uint8_t* images, int on_device, int nframes, int rows, int cols,
                   size_t pitch, size_t fstride, int level, int mp, cudaStream_t stream)
{
    if (det->pending.active) { MSG("A batch is already in flight on this detector: collect it first."); return -1; }
    if (nframes < 0 || rows <= 0 || cols <= 0 || rows > 32767 || cols > 32767 || pitch < (size_t)cols)
    { MSG("Bad batch geometry (nframes=%d rows=%d cols=%d pitch=%zu).", nframes, rows, cols, pitch); return -1; }
    if (rows > det->cfg.max_rows || cols > det->cfg.max_cols)
    { MSG("Frame %dx%d exceeds the detector's configured maximum %dx%d.", cols, rows, det->cfg.max_cols, det->cfg.max_rows); return -1; }
    if (level < 0 || level > 10)
    { MSG("Got an unreasonable image_pyramid_level = %d.", level); return -1; }
    DEVICE_GUARD(det);
    for (auto& t : det->timers) t.reset();

    // the clustering kernel's shared memory is sized from the previous batch's candidate counts; frames of another
    // geometry say nothing about this batch's: start from the largest size again (a list that outgrows the choice
    // takes the slow global-scratch path)
    if (rows != det->k2_rows || cols != det->k2_cols || level != det->k2_level)
    { det->k2_smem_cands = kClusterSmemCands; det->k2_rows = rows; det->k2_cols = cols; det->k2_level = level; }
    const int cap = det->cfg.candidate_capacity;
    if (det->h_xy.ensure(sizeof(int32_t) * 2 * mp * std::max(nframes, 1))) return -1;
    if (det->h_counts.ensure(sizeof(int32_t) * std::max(nframes, 1))) return -1;
    if (det->h_candcounts.ensure(sizeof(int32_t) * std::max(nframes, 1))) return -1;

    const int chunk = det->cfg.max_frames;
    // MRG_B200_K2_STREAM=main runs the clustering kernel on the ChESS kernel's stream (no overlap of
    // K2(c) with K1(c+1)); default: its own high-priority stream
    static const bool k2_on_main = [] { const char* e = getenv("MRG_B200_K2_STREAM"); return e && !strcmp(e, "main"); }();
    cudaStream_t aux = k2_on_main ? stream : det->aux_stream, cpy = det->copy_stream;
    // Launch sizes. For HOST frames the kernels of the last chunk run after its copy with nothing left to overlap,
    // so the last max_frames of a batch go up in halves down to 128 frames (512 frames: 256, 128, 128): the
    // exposed tail is one short copy-less launch pair instead of a full one (+1.8 % end to end on 4K frames).
    // Device-resident batches keep full-size launches: the ChESS kernel loses more on short launches than the
    // shorter tail wins (measured 1.97 against 2.01 Tpix/s). MRG_B200_TAPER=0/1 forces it off/on for both.
    static const int taper_env = [] { const char* e = getenv("MRG_B200_TAPER"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool taper = taper_env < 0 ? !on_device : taper_env != 0;
    std::vector<int> sizes;
    for (int left = nframes; left > 0; )
    {
        int n = std::min(chunk, left);
        if (taper && left <= chunk && left >= 2 * kTaperMinFrames)
            n = std::min(left, std::max(kTaperMinFrames, ((left / 2 + 63) / 64) * 64));
        sizes.push_back(n);
        left -= n;
    }
    // size both slots before anything is in flight (growing a buffer frees it)
    for (int b = 0; b < 2 && b < (int)sizes.size(); b++)
    {
        mrg_b200_detector::Slot& S = det->slot[b];
        const int n = std::min(chunk, nframes);
        if (ensure_chunk_scratch(det, S, n, mp)) return -1;
        if (!on_device && S.stage.ensure((size_t)round_up(cols, 16) * rows * n)) return -1;
        if (level > 0 && S.level_img.ensure(level_geom(rows, cols, level).frame_bytes * n)) return -1;
        S.used = false;
    }
    CUDA_TRY(cudaEventRecord(det->ev_fork, stream));
    CUDA_TRY(cudaStreamWaitEvent(aux, det->ev_fork, 0));
    CUDA_TRY(cudaStreamWaitEvent(cpy, det->ev_fork, 0));
    int f0 = 0;
    for (int c = 0; c < (int)sizes.size(); f0 += sizes[c], c++)
    {
        const int n = sizes[c];
        mrg_b200_detector::Slot& S = det->slot[c & 1];
        if (S.used)
        {
            // the slot's previous chunk must have left the clustering kernel
            CUDA_TRY(cudaStreamWaitEvent(stream, S.k2done, 0));
            CUDA_TRY(cudaStreamWaitEvent(cpy, S.k2done, 0));
        }
        FrameSet fs;
        if (stage_frames(det, S, images + (size_t)f0 * fstride, on_device, n, rows, cols, pitch, fstride, level, stream, cpy, &fs, true)) return -1;
        CUDA_TRY(cudaMemsetAsync(S.counts.p, 0, sizeof(uint32_t) * n, stream));
        if (chess_sparse(det, fs, (cand_t*)S.cand.p, (uint32_t*)S.counts.p, cap, stream)) return -1;
        CUDA_TRY(cudaEventRecord(S.k1done, stream));
        CUDA_TRY(cudaStreamWaitEvent(aux, S.k1done, 0));
        ClusterParams p; p.smem_cands = det->k2_smem_cands; p.level = level; p.cand_capacity = cap; p.max_points = mp; p.record_capacity = 2 * mp; p.records = S.records.p;
        {
            Launch l(det, 1, aux);
            CUDA_TRY(launch_cluster_find(fs, p, (cand_t*)S.cand.p, (uint32_t*)S.counts.p, (uint32_t*)S.table.p,
                                         (uint32_t*)S.dfs.p, (int32_t*)S.xy.p, nullptr, (int32_t*)S.outcounts.p, aux));
        }
        CUDA_TRY(cudaMemcpyAsync((int32_t*)det->h_xy.p + (size_t)f0 * 2 * mp, S.xy.p, sizeof(int32_t) * 2 * mp * n, cudaMemcpyDeviceToHost, aux));
        CUDA_TRY(cudaMemcpyAsync((int32_t*)det->h_counts.p + f0, S.outcounts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, aux));
        CUDA_TRY(cudaMemcpyAsync((int32_t*)det->h_candcounts.p + f0, S.counts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, aux));
        CUDA_TRY(cudaEventRecord(S.k2done, aux));
        S.used = true;
    }
    // No join here: the clustering kernels and result copies of the last chunks run on the aux stream and
    // collect() waits for THEM (their events), not for `stream`. Whatever the caller enqueues on `stream` next --
    // another detector's batch over the same frames, as bench.py's overlapped passes do -- starts as soon as this
    // batch's ChESS kernels are done, beside this batch's clustering tail. The frames must stay valid and
    // unchanged until collect() returns.
    det->pending.active = true;
    det->pending.images = images; det->pending.on_device = on_device; det->pending.nframes = nframes;
    det->pending.rows = rows; det->pending.cols = cols; det->pending.level = level;
    det->pending.pitch = pitch; det->pending.fstride = fstride; det->pending.stream = stream; det->pending.mp = mp;
    return 0;
}

int collect_locked(mrg_b200_detector* det, int32_t* xy_out, int32_t* counts_out)
{
    if (!det->pending.active) { MSG("No batch in flight."); return -1; }
    auto& pd = det->pending;
    pd.active = false;
    DEVICE_GUARD(det);
    // everything the batch enqueued ends in the k2done event of one of the two slots
    for (int b = 0; b < 2; b++)
        if (det->slot[b].used) CUDA_TRY(cudaEventSynchronize(det->slot[b].k2done));
    for (auto& t : det->timers) t.resolve();
    const int mp = pd.mp;
    int32_t* hxy = (int32_t*)det->h_xy.p; int32_t* hc = (int32_t*)det->h_counts.p; int32_t* hcc = (int32_t*)det->h_candcounts.p;
    for (int f = 0; f < pd.nframes; f++)
        if (hc[f] < 0)
            if (rerun_big(det, pd.images + (size_t)f * pd.fstride, pd.on_device, pd.rows, pd.cols, pd.pitch, pd.level, mp, pd.stream,
                          hxy + (size_t)f * 2 * mp, hc + f, hcc + f)) return -1;
    if (xy_out)     memcpy(xy_out, hxy, sizeof(int32_t) * 2 * mp * pd.nframes);
    if (counts_out) memcpy(counts_out, hc, sizeof(int32_t) * pd.nframes);
    // Size the clustering kernel's shared memory for the next batch from this one's longest candidate
    // list (with headroom): 22 KB for lists up to 1024 instead of 90 KB lets its CTAs run beside the
    // ChESS kernel's. Lists that outgrow the choice still work (global scratch), only slower.
    {
        static const bool fixed = getenv("MRG_B200_K2_SMEM_FIXED") != nullptr;
        int longest = 0;
        for (int f = 0; f < pd.nframes; f++) longest = std::max(longest, hcc[f]);
        const int want = longest + longest / 4;
        if (pd.nframes > 0 && !fixed) det->k2_smem_cands = want <= 1024 ? 1024 : want <= 2048 ? 2048 : kClusterSmemCands;
    }
    return 0;
}
}

// =================================================================================================
// C. batched API
// =================================================================================================
API int mrg_b200_detector_create(mrg_b200_detector** out, const mrg_b200_detector_config* config)
{
    if (!out || !config) return -1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    {
        MSG("No usable CUDA device: this library has no CPU implementation.");
        return -1;
    }
    mrg_b200_detector* det = new mrg_b200_detector;
    det->cfg = *config;
    if (det->cfg.device < 0) { if (cudaGetDevice(&det->cfg.device) != cudaSuccess) { delete det; return -1; } }
    det->device = det->cfg.device;
    if (det->cfg.max_frames <= 0) det->cfg.max_frames = 64;
    if (det->cfg.max_rows <= 0) det->cfg.max_rows = 32767;
    if (det->cfg.max_cols <= 0) det->cfg.max_cols = 32767;
    if (det->cfg.candidate_capacity <= 0) det->cfg.candidate_capacity = 32768;
    det->cfg.candidate_capacity = next_pow2(det->cfg.candidate_capacity);
    if (det->cfg.max_points <= 0) det->cfg.max_points = 1024;
    if (det->cfg.blur_radius < 0 || det->cfg.blur_radius > 4) { MSG("blur_radius must be in [0,4]; got %d.", det->cfg.blur_radius); delete det; return -1; }
    DeviceGuard dg(det->device);
    bool ok = dg.ok &&
              cudaStreamCreateWithFlags(&det->own_stream, cudaStreamNonBlocking) == cudaSuccess &&
              // clustering runs at high priority so its few CTAs slot in as ChESS CTAs retire
              // (MRG_B200_K2_PRIORITY=0: A/B knob, default priority -- they then wait for the ChESS grid's tail)
              cudaStreamCreateWithPriority(&det->aux_stream, cudaStreamNonBlocking,
                                           (getenv("MRG_B200_K2_PRIORITY") && atoi(getenv("MRG_B200_K2_PRIORITY")) == 0) ? 0 : -1) == cudaSuccess &&
              cudaStreamCreateWithFlags(&det->copy_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&det->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int b = 0; b < 2 && ok; b++)
        ok = cudaEventCreateWithFlags(&det->slot[b].staged, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&det->slot[b].k1done, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&det->slot[b].k2done, cudaEventDisableTiming) == cudaSuccess;
    if (!ok)
    {
        MSG("Could not initialise CUDA device %d.", det->device);
        delete det;
        return -1;
    }
    *out = det;
    return 0;
}

API void mrg_b200_detector_destroy(mrg_b200_detector* det)
{
    if (!det) return;
    if (det->boards_helper) { mrg_b200_detector_destroy(det->boards_helper); det->boards_helper = nullptr; }
    DeviceGuard dg(det->device);
    for (int k = 0; k < mrg_b200_detector::kBlobDepth; k++)
    {
        if (det->blobs[k]) { blob_workspace_destroy(det->blobs[k]); det->blobs[k] = nullptr; }
        if (det->blob_stream[k]) { cudaStreamDestroy(det->blob_stream[k]); det->blob_stream[k] = nullptr; }
        det->blob_slot[k].stage.release();
    }
    cudaDeviceSynchronize();
    for (auto& S : det->slot)
    {
        for (DeviceBuffer* b : { &S.stage, &S.blurred, &S.equalized, &S.pre_scratch, &S.level_img, &S.cand, &S.counts, &S.table, &S.dfs, &S.records, &S.xy, &S.outcounts }) b->release();
        for (cudaEvent_t e : { S.staged, S.k1done, S.k2done }) if (e) cudaEventDestroy(e);
    }
    for (DeviceBuffer* b : { &det->big_cand, &det->big_table, &det->big_dfs, &det->big_records, &det->pts, &det->lvls, &det->boards_frames,
                             &det->boards_gather, &det->mixed_stage, &det->mixed_srcs, &det->dbg_xyd }) b->release();
    if (det->ev_fork) cudaEventDestroy(det->ev_fork);
    if (det->aux_stream) cudaStreamDestroy(det->aux_stream);
    if (det->copy_stream) cudaStreamDestroy(det->copy_stream);
    for (PinnedBuffer* b : { &det->h_xy, &det->h_counts, &det->h_candcounts, &det->h_stage }) b->release();
    if (det->h_stage_free) cudaEventDestroy(det->h_stage_free);
    for (auto& t : det->timers) t.release();
    if (det->own_stream) cudaStreamDestroy(det->own_stream);
    delete det;
}

// mrg_b200_find_corners_batch() with an explicit per-frame output capacity (xy_out is [nframes][mp][2])
static int corners_batch_mp(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                            int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                            int image_pyramid_level, int mp, int32_t* xy_out, int32_t* counts_out, void* stream)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (enqueue_locked(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, image_pyramid_level, mp,
                       stream ? (cudaStream_t)stream : det->own_stream)) return -1;
    return collect_locked(det, xy_out, counts_out);
}

API int mrg_b200_find_corners_batch_enqueue(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                            int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                            int image_pyramid_level, void* stream)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    return enqueue_locked(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, image_pyramid_level,
                          det->cfg.max_points, stream ? (cudaStream_t)stream : det->own_stream);
}

API int mrg_b200_find_corners_batch_collect(mrg_b200_detector* det, int32_t* xy_out, int32_t* counts_out)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    return collect_locked(det, xy_out, counts_out);
}

API int mrg_b200_find_corners_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                    int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                    int image_pyramid_level, int32_t* xy_out, int32_t* counts_out, void* stream)
{
    if (!det) return -1;
    return corners_batch_mp(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, image_pyramid_level,
                            det->cfg.max_points, xy_out, counts_out, stream);
}

// One call over images of ANY sizes (the reference CLI takes a glob of images and hands them to its workers one by
// one, whatever their sizes: mrgingham-from-image.cc:50-54, 374-379). The images are grouped by (rows, cols); each
// group is laid out contiguously on the device (host images: one copy each, straight into the group's staging
// buffer; device images that already lie at a fixed stride are read in place) and goes through the batch path.
API int mrg_b200_find_corners_mixed_batch(mrg_b200_detector* det, const mrg_b200_image_desc* images, int nimages,
                                          int images_on_device, int image_pyramid_level,
                                          int32_t* xy_out, int32_t* counts_out, void* stream_)
{
    if (!det || nimages < 0 || (nimages > 0 && (!images || !counts_out))) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    const int mp = det->cfg.max_points;
    for (int i = 0; i < nimages; i++)
    {
        const mrg_b200_image_desc& d = images[i];
        if (!d.data || d.rows <= 0 || d.cols <= 0 || d.row_pitch < (size_t)d.cols)
        { MSG("Bad image %d (data=%p rows=%d cols=%d pitch=%zu).", i, (const void*)d.data, d.rows, d.cols, d.row_pitch); return -1; }
        counts_out[i] = 0;
    }
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    std::vector<int> order(nimages);
    for (int i = 0; i < nimages; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b)
    {
        if (images[a].rows != images[b].rows) return images[a].rows < images[b].rows;
        if (images[a].cols != images[b].cols) return images[a].cols < images[b].cols;
        return images[a].data < images[b].data;         // frames cut from one allocation end up at a fixed stride
    });
    std::vector<int32_t> txy, tcounts;
    std::vector<GatherSrc> srcs;
    const int super = std::max(1, 2 * det->cfg.max_frames);
    for (int g0 = 0; g0 < nimages; )
    {
        const int rows = images[order[g0]].rows, cols = images[order[g0]].cols;
        int g1 = g0;
        while (g1 < nimages && images[order[g1]].rows == rows && images[order[g1]].cols == cols) g1++;
        for (int c0 = g0; c0 < g1; c0 += super)
        {
            const int n = std::min(super, g1 - c0);
            const mrg_b200_image_desc& first = images[order[c0]];
            // device images at a fixed stride, in this order, with rows TMA can address: read in place
            bool in_place = images_on_device != 0;
            size_t fstride = n > 1 ? (size_t)(images[order[c0 + 1]].data - first.data) : first.row_pitch * rows;
            for (int i = 0; i < n && in_place; i++)
            {
                const mrg_b200_image_desc& d = images[order[c0 + i]];
                in_place = d.row_pitch == first.row_pitch && d.data == first.data + (size_t)i * fstride;
            }
            if (in_place && n > 1 && (images[order[c0 + 1]].data < first.data || fstride < first.row_pitch * rows)) in_place = false;
            const uint8_t* base; size_t pitch;
            if (in_place) { base = first.data; pitch = first.row_pitch; }
            else
            {
                pitch = (size_t)round_up(cols, 16);
                fstride = pitch * rows;
                if (det->mixed_stage.ensure(fstride * n)) return -1;
                if (images_on_device)
                {
                    // one gather kernel for the whole group (a copy call per image would cost more than the detector)
                    srcs.resize(n);
                    for (int i = 0; i < n; i++) { srcs[i].data = images[order[c0 + i]].data; srcs[i].pitch = images[order[c0 + i]].row_pitch; }
                    if (det->mixed_srcs.ensure(sizeof(GatherSrc) * n)) return -1;
                    CUDA_TRY(cudaMemcpyAsync(det->mixed_srcs.p, srcs.data(), sizeof(GatherSrc) * n, cudaMemcpyHostToDevice, stream));
                    CUDA_TRY(cudaStreamSynchronize(stream));       // (srcs is reused by the next group)
                    CUDA_TRY(launch_gather_frames((const GatherSrc*)det->mixed_srcs.p, n, rows, cols, (uint8_t*)det->mixed_stage.p, pitch, fstride, stream));
                }
                else
                    for (int i = 0; i < n; i++)
                    {
                        const mrg_b200_image_desc& d = images[order[c0 + i]];
                        CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)det->mixed_stage.p + (size_t)i * fstride, pitch, d.data, d.row_pitch, cols, rows,
                                                   cudaMemcpyHostToDevice, stream));
                    }
                base = (const uint8_t*)det->mixed_stage.p;
            }
            if (enqueue_locked(det, base, 1, n, rows, cols, pitch, fstride, image_pyramid_level, mp, stream)) return -1;
            txy.resize((size_t)n * 2 * mp); tcounts.resize(n);
            if (collect_locked(det, txy.data(), tcounts.data())) return -1;
            for (int i = 0; i < n; i++)
            {
                const int dst = order[c0 + i];
                counts_out[dst] = tcounts[i];
                if (xy_out) memcpy(xy_out + (size_t)dst * 2 * mp, txy.data() + (size_t)i * 2 * mp, sizeof(int32_t) * 2 * mp);
            }
        }
        g0 = g1;
    }
    return 0;
}

// The sparse form of the ChESS response: for every frame the list {(x, y, r) : r > 15} inside [7,w-7) x [7,h-7) that
// the production ChESS kernel (K1) emits and the clustering kernel consumes -- K1 alone, nothing after it.
API int mrg_b200_chess_candidates_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                        int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                        int image_pyramid_level, uint64_t* cand_out, int cand_cap, int32_t* counts_out,
                                        void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    if (nframes < 0 || rows <= 0 || cols <= 0 || rows > 32767 || cols > 32767 || row_pitch < (size_t)cols || (cand_out && cand_cap <= 0))
    { MSG("Bad batch geometry."); return -1; }
    if (image_pyramid_level < 0 || image_pyramid_level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", image_pyramid_level); return -1; }
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    for (auto& t : det->timers) t.reset();
    const int cap = det->cfg.candidate_capacity;
    const int chunk = std::max(1, det->cfg.max_frames);
    std::vector<uint32_t> hc;
    for (int f0 = 0; f0 < nframes; f0 += chunk)
    {
        const int n = std::min(chunk, nframes - f0);
        mrg_b200_detector::Slot& S = det->slot[0];
        if (S.cand.ensure(sizeof(cand_t) * (size_t)cap * n) || S.counts.ensure(sizeof(uint32_t) * n)) return -1;
        FrameSet fs;
        if (stage_frames(det, S, images + (size_t)f0 * frame_stride, images_on_device, n, rows, cols, row_pitch, frame_stride,
                         image_pyramid_level, stream, stream, &fs, true)) return -1;
        CUDA_TRY(cudaMemsetAsync(S.counts.p, 0, sizeof(uint32_t) * n, stream));
        if (chess_sparse(det, fs, (cand_t*)S.cand.p, (uint32_t*)S.counts.p, cap, stream)) return -1;
        hc.resize(n);
        CUDA_TRY(cudaMemcpyAsync(hc.data(), S.counts.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
        for (int i = 0; i < n; i++)
        {
            if (counts_out) counts_out[f0 + i] = (int32_t)hc[i];           // may exceed the capacities: the list is then truncated
            const size_t keep = std::min<size_t>(std::min<size_t>(hc[i], (size_t)cap), (size_t)std::max(cand_cap, 0));
            if (cand_out && keep)
                CUDA_TRY(cudaMemcpy(cand_out + (size_t)(f0 + i) * cand_cap, (cand_t*)S.cand.p + (size_t)i * cap, sizeof(cand_t) * keep, cudaMemcpyDeviceToHost));
        }
    }
    for (auto& t : det->timers) t.resolve();
    return 0;
}

API int mrg_b200_chess_response_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                      int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                      int16_t* response, int response_on_device, void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    if (nframes <= 0 || rows <= 0 || cols <= 0 || row_pitch < (size_t)cols) { MSG("Bad batch geometry."); return -1; }
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    DEVICE_GUARD(det);
    const size_t fe = (size_t)rows * cols;
    const int chunk = response_on_device ? nframes : std::max(1, det->cfg.max_frames);
    for (int f0 = 0; f0 < nframes; f0 += chunk)
    {
        const int n = std::min(chunk, nframes - f0);
        mrg_b200_detector::Slot& S = det->slot[0];
        FrameSet fs;
        if (stage_frames(det, S, images + (size_t)f0 * frame_stride, images_on_device, n, rows, cols, row_pitch, frame_stride, 0, stream, stream, &fs)) return -1;
        int16_t* dresp;
        if (response_on_device) dresp = response + (size_t)f0 * fe;
        else
        {
            // scratch: reuse the candidate buffer allocation
            if (S.cand.ensure(sizeof(int16_t) * fe * n)) return -1;
            dresp = (int16_t*)S.cand.p;
        }
        if (det->cfg.kernel_variant == 1) CUDA_TRY(launch_chess_dense(fs, dresp, fe, stream));
        else                              CUDA_TRY(launch_chess_dense_tiled(fs, dresp, fe, stream));
        if (!response_on_device && cols > 2*kMargin && rows > 2*kMargin)
        {
            // only the interior travels back: the caller's border elements stay untouched (ChESS.c:62-63)
            for (int i = 0; i < n; i++)
            {
                const size_t off = (size_t)i * fe + (size_t)kMargin * cols + kMargin;
                CUDA_TRY(cudaMemcpy2DAsync(response + (size_t)(f0 + i) * fe + (size_t)kMargin * cols + kMargin, sizeof(int16_t) * cols,
                                           dresp + off, sizeof(int16_t) * cols,
                                           sizeof(int16_t) * (cols - 2*kMargin), rows - 2*kMargin, cudaMemcpyDeviceToHost, stream));
            }
        }
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return 0;
}

// clahe (optional) then blur (optional) of a batch into a dense output; the order the reference CLI applies them in
static int preprocess_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                            int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                            int clahe, int blur_radius, uint8_t* out, int out_on_device, void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (nframes < 0 || rows <= 0 || cols <= 0 || row_pitch < (size_t)cols) { MSG("Bad batch geometry."); return -1; }
    if (blur_radius < 0 || blur_radius > 4) { MSG("blur_radius must be in [0,4]; got %d.", blur_radius); return -1; }
    if (!clahe && blur_radius == 0) { MSG("Nothing to do: neither clahe nor a blur was asked for."); return -1; }
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    const size_t fe = (size_t)rows * cols;
    const int chunk = std::min(out_on_device && images_on_device ? std::max(nframes, 1) : std::max(1, det->cfg.max_frames), 65535);
    for (int f0 = 0; f0 < nframes; f0 += chunk)
    {
        const int n = std::min(chunk, nframes - f0);
        mrg_b200_detector::Slot& S = det->slot[0];
        FrameSet fs;
        if (stage_frames(det, S, images + (size_t)f0 * frame_stride, images_on_device, n, rows, cols, row_pitch, frame_stride, 0, stream, stream, &fs)) return -1;
        uint8_t* d = out + (size_t)f0 * fe;
        if (!out_on_device)
        {
            if (S.level_img.ensure(fe * n)) return -1;
            d = (uint8_t*)S.level_img.p;
        }
        if (clahe)
        {
            uint8_t* e = d; int epitch = cols; size_t eframe = fe;
            if (blur_radius > 0)
            {
                epitch = round_up(cols, 16); eframe = (size_t)epitch * rows;
                if (S.equalized.ensure(eframe * n)) return -1;
                e = (uint8_t*)S.equalized.p;
            }
            if (S.pre_scratch.ensure(clahe_scratch_bytes(n))) return -1;
            CUDA_TRY(launch_normalize_clahe(fs, true, 8.0, e, epitch, eframe, S.pre_scratch.p, stream));
            fs.base = e; fs.pitch = epitch; fs.frame_stride = eframe;
        }
        if (blur_radius > 0) CUDA_TRY(launch_box_blur(fs, blur_radius, d, cols, fe, stream));
        if (!out_on_device) CUDA_TRY(cudaMemcpyAsync(out + (size_t)f0 * fe, d, fe * n, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return 0;
}

// The reference CLI's handling of 16-bit images (mrgingham-from-image.cc:83-93), then its blur: 16 bits in, the 8-bit
// image the detector would be given out.
API int mrg_b200_preprocess16_batch(mrg_b200_detector* det, const uint16_t* images, int images_on_device,
                                    int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                    int clahe, int blur_radius, uint8_t* out, int out_on_device, void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (nframes < 0 || rows <= 0 || cols <= 0 || row_pitch < 2 * (size_t)cols || (row_pitch & 1) || (frame_stride & 1) || ((uintptr_t)images & 1))
    { MSG("Bad batch geometry (16-bit frames: pitch and stride in bytes, even, pitch >= 2*cols)."); return -1; }
    if (blur_radius < 0 || blur_radius > 4) { MSG("blur_radius must be in [0,4]; got %d.", blur_radius); return -1; }
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    const size_t fe = (size_t)rows * cols;
    // the 65536-bin histograms and tables of the 16-bit CLAHE take 24 MB per frame: modest chunks
    const int chunk = std::max(1, std::min(clahe ? 32 : 65535, out_on_device && images_on_device ? std::max(nframes, 1) : std::max(1, det->cfg.max_frames)));
    for (int f0 = 0; f0 < nframes; f0 += chunk)
    {
        const int n = std::min(chunk, nframes - f0);
        mrg_b200_detector::Slot& S = det->slot[0];
        const uint16_t* src = (const uint16_t*)((const uint8_t*)images + (size_t)f0 * frame_stride);
        size_t src_stride = frame_stride / 2; int src_pitch = (int)(row_pitch / 2);
        if (!images_on_device)
        {
            if (S.stage.ensure(fe * 2 * n)) return -1;
            for (int i = 0; i < n; i++)
                CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)S.stage.p + (size_t)i * fe * 2, (size_t)cols * 2, (const uint8_t*)src + (size_t)i * frame_stride, row_pitch,
                                           (size_t)cols * 2, rows, cudaMemcpyHostToDevice, stream));
            src = (const uint16_t*)S.stage.p; src_stride = fe; src_pitch = cols;
        }
        uint8_t* d = out + (size_t)f0 * fe;
        if (!out_on_device)
        {
            if (S.level_img.ensure(fe * n)) return -1;
            d = (uint8_t*)S.level_img.p;
        }
        uint8_t* e = d; int epitch = cols; size_t eframe = fe;
        if (blur_radius > 0)
        {
            epitch = round_up(cols, 16); eframe = (size_t)epitch * rows;
            if (S.equalized.ensure(eframe * n)) return -1;
            e = (uint8_t*)S.equalized.p;
        }
        if (clahe && S.pre_scratch.ensure(preproc16_scratch_bytes(n))) return -1;
        CUDA_TRY(launch_preprocess16(src, src_stride, src_pitch, cols, rows, n, clahe != 0, e, epitch, eframe, S.pre_scratch.p, stream));
        if (blur_radius > 0)
        {
            FrameSet fs; fs.base = e; fs.pitch = epitch; fs.frame_stride = eframe; fs.w = cols; fs.h = rows; fs.nframes = n;
            CUDA_TRY(launch_box_blur(fs, blur_radius, d, cols, fe, stream));
        }
        if (!out_on_device) CUDA_TRY(cudaMemcpyAsync(out + (size_t)f0 * fe, d, fe * n, cudaMemcpyDeviceToHost, stream));
        CUDA_TRY(cudaStreamSynchronize(stream));
    }
    return 0;
}

API int mrg_b200_box_blur_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                int blur_radius, uint8_t* out, int out_on_device, void* stream_)
{
    if (blur_radius < 1 || blur_radius > 4) { MSG("blur_radius must be in [1,4]; got %d.", blur_radius); return -1; }
    return preprocess_batch(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, 0, blur_radius, out, out_on_device, stream_);
}

API int mrg_b200_preprocess_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                  int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                  int clahe, int blur_radius, uint8_t* out, int out_on_device, void* stream_)
{
    return preprocess_batch(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, clahe != 0, blur_radius, out, out_on_device, stream_);
}

API int mrg_b200_pyramid_level(mrg_b200_detector* det, const uint8_t* image, int rows, int cols, size_t row_pitch,
                               int level, uint8_t* out, int* orows, int* ocols)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (level < 0 || level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", level); return -1; }
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    DEVICE_GUARD(det);
    cudaStream_t stream = det->own_stream;
    FrameSet fs;
    if (stage_frames(det, det->slot[0], image, 0, 1, rows, cols, row_pitch, row_pitch * rows, level, stream, stream, &fs)) return -1;
    *orows = fs.h; *ocols = fs.w;
    if (out && fs.w > 0 && fs.h > 0)
        CUDA_TRY(cudaMemcpy2DAsync(out, fs.w, fs.base, fs.pitch, fs.w, fs.h, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    return 0;
}

API int mrg_b200_last_kernel_ms(mrg_b200_detector* det, int which, float* ms, int* launches)
{
    if (!det || which < 0 || which > 3) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (which == 3) { if (ms) *ms = det->blob_ms; if (launches) *launches = 0; return 0; }
    if (ms) *ms = det->timers[which].ms;
    if (launches) *launches = det->timers[which].launches;
    return 0;
}
API void mrg_b200_set_profiling(mrg_b200_detector* det, int enabled) { if (det) { std::lock_guard<std::mutex> g(det->mtx); det->profiling = enabled != 0; } }

API int mrg_b200_last_candidate_counts(mrg_b200_detector* det, int32_t* counts_out, int nframes)
{
    if (!det || !det->h_candcounts.p) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    memcpy(counts_out, det->h_candcounts.p, sizeof(int32_t) * nframes);
    return 0;
}

API const char* mrg_b200_version(void) { return "mrgingham_b200 0.1 (sm_100a)"; }

API int mrg_b200_device_count(void)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess) return 0;
    return ndev;
}

static int blobs_batch_mp(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                          int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                          int mp, int32_t* xy_out, int32_t* counts_out, void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    if (nframes < 0 || rows <= 0 || cols <= 0 || rows > 32767 || cols > 32767 || row_pitch < (size_t)cols)
    { MSG("Bad batch geometry (nframes=%d rows=%d cols=%d pitch=%zu).", nframes, rows, cols, row_pitch); return -1; }
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    constexpr int D = mrg_b200_detector::kBlobDepth;
    for (int k = 0; k < D; k++)
    {
        if (!det->blobs[k]) det->blobs[k] = blob_workspace_create();
        if (!det->blob_stream[k]) CUDA_TRY(cudaStreamCreateWithFlags(&det->blob_stream[k], cudaStreamNonBlocking));
    }
    det->blob_ms = 0;
    // Every border is followed by its own lane, but a frame-sized outline still takes one lane a few milliseconds,
    // during which most of the GPU has nothing left to do for that chunk: D chunks are kept in flight, each on its
    // own stream and workspace, so that the tails of some chunks' border walks run beside other chunks' kernels,
    // and the host-side grouping of a finished chunk beside all of them.
    int chunk = std::max(1, std::min(det->cfg.max_frames, 64));
    if (const char* e = getenv("MRG_B200_BLOB_CHUNK")) { const int v = atoi(e); if (v >= 1 && v <= 256) chunk = std::min(v, det->cfg.max_frames); }
    CUDA_TRY(cudaEventRecord(det->ev_fork, stream));
    for (int k = 0; k < D; k++) CUDA_TRY(cudaStreamWaitEvent(det->blob_stream[k], det->ev_fork, 0));   // the frames may come from work queued on `stream`
    struct InFlight { FrameSet fs; int f0 = 0, n = 0; bool live = false; } fl[D];
    float ms = 0;
    auto finish = [&](int k) -> int
    {
        InFlight& c = fl[k];
        if (!c.live) return 0;
        c.live = false;
        const int rc = blob_finish(det->blobs[k], xy_out + (size_t)c.f0 * 2 * mp, counts_out + c.f0, mp, det->profiling ? &ms : nullptr);
        if (rc < 0) return -1;
        if (rc == 1)
        {
            // some frame of the chunk needs more scratch than the default: frame by frame, on the GPU
            for (int i = 0; i < c.n; i++)
            {
                FrameSet one = c.fs; one.base = c.fs.base + (size_t)i * c.fs.frame_stride; one.nframes = 1;
                float ms1 = 0;
                if (blob_find_frames(det->blobs[k], one, xy_out + (size_t)(c.f0 + i) * 2 * mp, counts_out + c.f0 + i, mp, det->blob_stream[k],
                                     det->profiling ? &ms1 : nullptr)) return -1;
                ms += ms1;
                blob_workspace_reset_capacity(det->blobs[k]);
            }
        }
        return 0;
    };
    int rc = 0, ci = 0;
    for (int f0 = 0; f0 < nframes && !rc; f0 += chunk, ci++)
    {
        const int k = ci % D, n = std::min(chunk, nframes - f0);
        if (finish(k)) { rc = -1; break; }
        if (stage_frames(det, det->blob_slot[k], images + (size_t)f0 * frame_stride, images_on_device, n, rows, cols, row_pitch, frame_stride,
                         0, det->blob_stream[k], det->blob_stream[k], &fl[k].fs)) { rc = -1; break; }
        if (blob_enqueue(det->blobs[k], fl[k].fs, det->blob_stream[k])) { rc = -1; break; }
        fl[k].f0 = f0; fl[k].n = n; fl[k].live = true;
    }
    // (after a failure the chunks still in flight are waited for, their results dropped)
    for (int j = 0; j < D; j++) { const int k = (ci + j) % D; if (finish(k)) rc = -1; }
    det->blob_ms = ms;
    return rc;
}

API int mrg_b200_find_blobs_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                  int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                  int32_t* xy_out, int32_t* counts_out, void* stream_)
{
    if (!det) return -1;
    return blobs_batch_mp(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, det->cfg.max_points,
                          xy_out, counts_out, stream_);
}

// =================================================================================================
// A/B. single-image entry points on a process-wide default detector
// =================================================================================================
// A call borrows a private detector for its duration from a per-device pool (created on first use, at most
// kPoolMax per device; further callers wait for one to come back). N host threads -- the reference CLI's -j N,
// mrgingham-from-image.cc:374-379 -- therefore run on N streams with N sets of scratch, concurrently, and the
// detector a call gets lives on the calling thread's CURRENT device (a thread working on cuda:1 stays there).
namespace
{
constexpr int kMaxDevices = 64, kPoolMax = 16;
constexpr int kDefaultPoints = 4096;     // first guess at a frame's corner count; grown per call, never stored
struct DefaultPool
{
    std::mutex m;
    std::condition_variable cv;
    std::vector<mrg_b200_detector*> idle[kMaxDevices];
    int created[kMaxDevices] = {};
};
DefaultPool& default_pool()
{
    static DefaultPool* p = new DefaultPool;   // never destroyed: the CUDA runtime may be gone by static-destruction time
    return *p;
}
struct DefaultLease
{
    mrg_b200_detector* det = nullptr;
    int dev = -1;
    DefaultLease()
    {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
        {
            MSG("No usable CUDA device: this library has no CPU implementation.");
            dev = -1;
            return;
        }
        DefaultPool& P = default_pool();
        {
            std::unique_lock<std::mutex> lk(P.m);
            for (;;)
            {
                if (!P.idle[dev].empty()) { det = P.idle[dev].back(); P.idle[dev].pop_back(); return; }
                if (P.created[dev] < kPoolMax) { P.created[dev]++; break; }
                P.cv.wait(lk);
            }
        }
        mrg_b200_detector_config cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.device = dev; cfg.max_frames = 1; cfg.candidate_capacity = 65536; cfg.max_points = kDefaultPoints;
        if (mrg_b200_detector_create(&det, &cfg))
        {
            det = nullptr;
            std::lock_guard<std::mutex> lk(P.m);
            P.created[dev]--;
            P.cv.notify_one();
            return;
        }
        // one-image calls hand over pageable host memory: stage it through the detector's own pinned buffer
        static const bool pin = [] { const char* e = getenv("MRG_B200_PIN_STAGE"); return !e || atoi(e) != 0; }();
        det->stage_via_pinned = pin;
    }
    ~DefaultLease()
    {
        if (!det) return;
        DefaultPool& P = default_pool();
        std::lock_guard<std::mutex> lk(P.m);
        P.idle[dev].push_back(det);
        P.cv.notify_one();
    }
    DefaultLease(const DefaultLease&) = delete;
    DefaultLease& operator=(const DefaultLease&) = delete;
};
}

// Return conventions of the one-image calls: 0 = nothing found (including the reference's own error paths, which
// report "no points"), < 0 = the GPU path FAILED (CUDA error, no device). The two are kept apart all the way up:
// the drop-in bridge symbols turn a failure into the reference's error return, never into "found nothing".
API int mrg_b200_find_chessboard_corners(const uint8_t* image, int Nrows, int Ncols, int stride, int level,
                                         int* xy_out, int cap)
{
    // the reference's error paths give "no points" (find_chessboard_corners.cc:433-441, :461-466)
    if (level < 0 || level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", level); return 0; }
    if (level == 0 && Nrows > 1 && stride != Ncols) { MSG("I can only handle continuous arrays (stride == width) currently."); return 0; }
    DefaultLease L;
    if (!L.det) return -1;
    int mp = kDefaultPoints;
    std::vector<int32_t> xy;
    int32_t n = 0;
    for (;;)
    {
        xy.resize((size_t)2 * mp);
        if (corners_batch_mp(L.det, image, 0, 1, Nrows, Ncols, (size_t)stride, (size_t)stride * Nrows, level, mp, xy.data(), &n, nullptr)) return -1;
        if (n <= mp) break;
        mp = next_pow2(n);         // more corners than the first guess: look again with room for all of them
    }
    for (int i = 0; i < n && i < cap; i++) { xy_out[2*i] = xy[2*i]; xy_out[2*i + 1] = xy[2*i + 1]; }
    return n;
}

API int mrg_b200_find_blobs(const uint8_t* image, int Nrows, int Ncols, int stride, int* xy_out, int cap)
{
    if (Nrows <= 0 || Ncols <= 0 || stride < Ncols) { MSG("Bad image geometry."); return 0; }
    DefaultLease L;
    if (!L.det) return -1;
    int mp = kDefaultPoints;
    std::vector<int32_t> xy;
    int32_t n = 0;
    for (;;)
    {
        xy.resize((size_t)2 * mp);
        if (blobs_batch_mp(L.det, image, 0, 1, Nrows, Ncols, (size_t)stride, (size_t)stride * Nrows, mp, xy.data(), &n, nullptr)) return -1;
        if (n <= mp) break;
        mp = next_pow2(n);
    }
    for (int i = 0; i < n && i < cap; i++) { xy_out[2*i] = xy[2*i]; xy_out[2*i + 1] = xy[2*i + 1]; }
    return n;
}

// The corners of one image as un-quantised doubles (what the reference's debug dump prints): the clustering kernel
// writes them beside the integers when asked to. Returns the count, -1 on failure, -2 if the frame needs the
// large-list path (the dump is then skipped).
static int corners_one_dbl(mrg_b200_detector* det, const uint8_t* image, int rows, int cols, int stride, int level, std::vector<double>* xyd)
{
    std::lock_guard<std::mutex> g(det->mtx);
    if (det->pending.active) return -1;
    DEVICE_GUARD(det);
    cudaStream_t stream = det->own_stream;
    mrg_b200_detector::Slot& S = det->slot[0];
    const int mp = 1 << 14;
    FrameSet fs;
    if (stage_frames(det, S, image, 0, 1, rows, cols, (size_t)stride, (size_t)stride * rows, level, stream, stream, &fs, true)) return -1;
    if (ensure_chunk_scratch(det, S, 1, mp)) return -1;
    if (det->dbg_xyd.ensure(sizeof(double) * 2 * mp)) return -1;
    const int cap = det->cfg.candidate_capacity;
    CUDA_TRY(cudaMemsetAsync(S.counts.p, 0, sizeof(uint32_t), stream));
    if (chess_sparse(det, fs, (cand_t*)S.cand.p, (uint32_t*)S.counts.p, cap, stream)) return -1;
    ClusterParams p; p.smem_cands = kClusterSmemCands; p.level = level; p.cand_capacity = cap; p.max_points = mp; p.record_capacity = 2 * mp; p.records = S.records.p;
    CUDA_TRY(launch_cluster_find(fs, p, (cand_t*)S.cand.p, (uint32_t*)S.counts.p, (uint32_t*)S.table.p, (uint32_t*)S.dfs.p,
                                 (int32_t*)S.xy.p, (double*)det->dbg_xyd.p, (int32_t*)S.outcounts.p, stream));
    int32_t n = 0;
    CUDA_TRY(cudaMemcpyAsync(&n, S.outcounts.p, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    if (n < 0 || n > mp) return -2;
    xyd->resize((size_t)2 * n);
    if (n) CUDA_TRY(cudaMemcpy(xyd->data(), det->dbg_xyd.p, sizeof(double) * 2 * n, cudaMemcpyDeviceToHost));
    return n;
}

// The reference's --debug artefacts of one corner-detector call (debug_dump.cu lists the files).
// refined_xy == NULL: the find branch (the corners are computed here once more, as doubles); otherwise the refinement
// branch: the points whose level is image_pyramid_level after the call are the ones that were refined.
API int mrg_b200_debug_dump_corners(const uint8_t* image, int Nrows, int Ncols, int stride, int image_pyramid_level,
                                    const char* debug_image_filename,
                                    const double* refined_xy, const signed char* refined_levels, int Npoints)
{
    const int level = image_pyramid_level;
    if (level < 0 || level > 10 || Nrows <= 0 || Ncols <= 0 || stride < Ncols) return -1;
    DefaultLease L;
    if (!L.det) return -1;
    const bool refinement = refined_xy != nullptr;
    char filename[256];
    // the level image
    int lh = 0, lw = 0;
    std::vector<uint8_t> limg((size_t)Nrows * Ncols);
    if (mrg_b200_pyramid_level(L.det, image, Nrows, Ncols, (size_t)stride, level, limg.data(), &lh, &lw)) return -1;
    snprintf(filename, sizeof(filename), "/tmp/mrgingham-scaled-processed-level%d.png", level);
    if (write_png_gray8(filename, limg.data(), lw, lh, (size_t)lw)) fprintf(stderr, "Wrote scaled,processed image to %s\n", filename);
    // the response, as the reference holds it: zeros where ChESS writes nothing
    std::vector<int16_t> resp((size_t)lw * lh, 0);
    if (lw > 2 * kMargin && lh > 2 * kMargin &&
        mrg_b200_chess_response_batch(L.det, limg.data(), 0, 1, lh, lw, (size_t)lw, (size_t)lw * lh, resp.data(), 0, nullptr)) return -1;
    std::vector<uint8_t> out8(resp.size());
    normalize_response_u8(resp.data(), resp.size(), out8.data());
    snprintf(filename, sizeof(filename), "/tmp/mrgingham-chess-response%s-level%d.png", refinement ? "-refinement" : "", level);
    if (write_png_gray8(filename, out8.data(), lw, lh, (size_t)lw)) fprintf(stderr, "Wrote a normalized ChESS response to %s\n", filename);
    for (auto& v : resp) if (v < 0) v = 0;
    normalize_response_u8(resp.data(), resp.size(), out8.data());
    snprintf(filename, sizeof(filename), "/tmp/mrgingham-chess-response%s-level%d-positive.png", refinement ? "-refinement" : "", level);
    if (write_png_gray8(filename, out8.data(), lw, lh, (size_t)lw)) fprintf(stderr, "Wrote positive-only, normalized ChESS response to %s\n", filename);
    // the corners
    std::vector<double> xyd;
    if (!refinement)
    {
        const int n = corners_one_dbl(L.det, image, Nrows, Ncols, stride, level, &xyd);
        if (n == -2) { MSG("Too many candidates for the debug corner dump; skipping it."); return 0; }
        if (n < 0) return -1;
        snprintf(filename, sizeof(filename), "/tmp/mrgingham-1-corners.vnl");
    }
    else
    {
        for (int i = 0; i < Npoints; i++)
            if (refined_levels[i] == level) { xyd.push_back(refined_xy[2*i]); xyd.push_back(refined_xy[2*i + 1]); }
        snprintf(filename, sizeof(filename), "/tmp/mrgingham-1-corners-refinement-level%d.vnl", level);
    }
    fprintf(stderr, "Writing self-plotting corner dump to %s\n", filename);
    if (!write_corner_vnl(filename, debug_image_filename, xyd.data(), (int)(xyd.size() / 2))) return -1;
    return 0;
}

API bool find_chessboard_corners_from_image_array_C(int Nrows, int Ncols, int stride, char* imagebuffer,
                                                    int image_pyramid_level, bool doblobs, bool debug,
                                                    bool (*add_points)(int* xy, int N, double scale, void* cookie), void* cookie)
{
    // mrgingham_pywrap_cplusplus_bridge.cc:50-56: blobs only at level 0, via find_blobs_from_image_array()
    if (doblobs && image_pyramid_level != 0) return false;
    int cap = 4096;
    std::vector<int> xy((size_t)2 * cap);
    auto look = [&]() -> int
    {
        return doblobs ? mrg_b200_find_blobs((const uint8_t*)imagebuffer, Nrows, Ncols, stride, xy.data(), cap)
                       : mrg_b200_find_chessboard_corners((const uint8_t*)imagebuffer, Nrows, Ncols, stride, image_pyramid_level, xy.data(), cap);
    };
    int n = look();
    if (n > cap) { cap = n; xy.resize((size_t)2 * cap); n = look(); }
    if (n < 0)
    {
        // The GPU path failed. "false with nothing added" means "found nothing" to the caller (mrgingham_pywrap.c:
        // 203-211 returns an empty array), so hand over an empty result first: false WITH a result is the bridge's
        // error return (mrgingham_pywrap.c:212-219 raises RuntimeError).
        MSG("The GPU corner detector failed; reporting an error, not 'no points'.");
        (*add_points)(xy.data(), 0, 1.0 / kFindGridScale, cookie);
        return false;
    }
    if (debug && !doblobs && n >= 0)
        mrg_b200_debug_dump_corners((const uint8_t*)imagebuffer, Nrows, Ncols, stride, image_pyramid_level, nullptr, nullptr, nullptr, 0);
    if (n == 0) return false;
    return (*add_points)(xy.data(), n, 1.0 / kFindGridScale, cookie);
}

// A failure here cannot be returned (void, as ChESS.h:31-34 declares it): it is reported on stderr and `response` is
// left untouched -- the host process is never taken down.
API void mrgingham_ChESS_response_5(int16_t* response, const uint8_t* image, int w, int h, int stride)
{
    if (w <= 2*kMargin || h <= 2*kMargin) return; // nothing is written for such sizes (ChESS.c:62-63)
    DefaultLease L;
    if (!L.det) { MSG("No CUDA device: response NOT computed, the output buffer is untouched."); return; }
    if (mrg_b200_chess_response_batch(L.det, image, 0, 1, h, w, (size_t)stride, (size_t)stride * h, response, 0, nullptr))
        MSG("ChESS response failed on the GPU: response NOT computed, the output buffer may be partly written.");
}

// Refinement of `n` frames (n <= max_frames, or n == 1 with big == true for a frame whose candidate
// list overflowed the chunk scratch). xy/levels/nref are HOST arrays [n][npoints][2], [n][npoints], [n].
static int refine_frames(mrg_b200_detector* det, cudaStream_t stream, const uint8_t* images, int on_device, int n,
                         int rows, int cols, size_t pitch, size_t fstride, int level,
                         double* xy, signed char* levels, int npoints, int32_t* nref, bool big)
{
    mrg_b200_detector::Slot& S = det->slot[0];
    FrameSet fs;
    if (stage_frames(det, S, images, on_device, n, rows, cols, pitch, fstride, level, stream, stream, &fs, true)) return -1;
    if (det->pts.ensure(sizeof(double) * 2 * npoints * n)) return -1;
    if (det->lvls.ensure((size_t)npoints * n)) return -1;
    if (ensure_chunk_scratch(det, S, n, det->cfg.max_points)) return -1;
    if (det->big_records.ensure(cluster_record_bytes() * (size_t)npoints * n)) return -1;
    int cap = det->cfg.candidate_capacity;
    cand_t* cand = (cand_t*)S.cand.p; uint32_t* table = (uint32_t*)S.table.p; uint32_t* dfs = (uint32_t*)S.dfs.p;
    if (big)
    {
        cap = next_pow2((long long)fs.w * fs.h);
        if (det->big_cand.ensure(sizeof(cand_t) * cap)) return -1;
        if (det->big_table.ensure(sizeof(uint32_t) * 2 * cap)) return -1;
        if (det->big_dfs.ensure(sizeof(uint32_t) * cap)) return -1;
        cand = (cand_t*)det->big_cand.p; table = (uint32_t*)det->big_table.p; dfs = (uint32_t*)det->big_dfs.p;
    }
    CUDA_TRY(cudaMemcpyAsync(det->pts.p, xy, sizeof(double) * 2 * npoints * n, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemcpyAsync(det->lvls.p, levels, (size_t)npoints * n, cudaMemcpyHostToDevice, stream));
    CUDA_TRY(cudaMemsetAsync(S.counts.p, 0, sizeof(uint32_t) * n, stream));
    if (chess_sparse(det, fs, cand, (uint32_t*)S.counts.p, cap, stream)) return -1;
    ClusterParams p; p.smem_cands = det->k2_smem_cands; p.level = level; p.cand_capacity = cap; p.max_points = det->cfg.max_points;
    p.record_capacity = npoints; p.records = det->big_records.p;
    {
        Launch l(det, 1, stream);
        CUDA_TRY(launch_cluster_refine(fs, p, cand, (uint32_t*)S.counts.p, table, dfs, (double*)det->pts.p, (signed char*)det->lvls.p,
                                       npoints, (int32_t*)S.outcounts.p, stream));
    }
    std::vector<int32_t> got(n);
    CUDA_TRY(cudaMemcpyAsync(got.data(), S.outcounts.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaStreamSynchronize(stream));
    bool any_overflow = false;
    for (int i = 0; i < n; i++) any_overflow |= got[i] < 0;
    if (!any_overflow)
    {
        CUDA_TRY(cudaMemcpy(xy, det->pts.p, sizeof(double) * 2 * npoints * n, cudaMemcpyDeviceToHost));
        CUDA_TRY(cudaMemcpy(levels, det->lvls.p, (size_t)npoints * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) nref[i] = got[i];
        return 0;
    }
    if (big) { MSG("Frame still overflows at worst-case capacity; this is a bug."); return -1; }
    // some candidate list overflowed: frames that did not are final, the others are re-run one by
    // one on the GPU with worst-case capacity (the kernel leaves an overflowed frame's points untouched)
    std::vector<double> hxy((size_t)2 * npoints * n); std::vector<signed char> hl((size_t)npoints * n);
    CUDA_TRY(cudaMemcpy(hxy.data(), det->pts.p, sizeof(double) * 2 * npoints * n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(hl.data(), det->lvls.p, (size_t)npoints * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
    {
        double* xi = xy + (size_t)2 * npoints * i; signed char* li = levels + (size_t)npoints * i;
        if (got[i] >= 0)
        {
            memcpy(xi, hxy.data() + (size_t)2 * npoints * i, sizeof(double) * 2 * npoints);
            memcpy(li, hl.data() + (size_t)npoints * i, npoints);
            nref[i] = got[i];
        }
        else if (refine_frames(det, stream, images + (size_t)i * fstride, on_device, 1, rows, cols, pitch, pitch * rows, level,
                               xi, li, npoints, nref + i, true)) return -1;
    }
    return 0;
}

API int mrg_b200_refine_corners_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                      int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                      int level, double* xy_inout, signed char* levels, int npoints,
                                      int32_t* nrefined_out, void* stream_)
{
    if (!det) return -1;
    std::lock_guard<std::mutex> g(det->mtx);
    if (det->pending.active) { MSG("A batch is in flight on this detector: collect it first."); return -1; }
    if (level < 0 || level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", level); return -1; }
    if (nframes < 0 || rows <= 0 || cols <= 0 || rows > 32767 || cols > 32767 || row_pitch < (size_t)cols || npoints < 0)
    { MSG("Bad batch geometry."); return -1; }
    for (int i = 0; i < nframes; i++) nrefined_out[i] = 0;
    if (npoints == 0 || nframes == 0) return 0;
    DEVICE_GUARD(det);
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    const int chunk = det->cfg.max_frames;
    for (int f0 = 0; f0 < nframes; f0 += chunk)
    {
        const int n = std::min(chunk, nframes - f0);
        if (refine_frames(det, stream, images + (size_t)f0 * frame_stride, images_on_device, n, rows, cols, row_pitch, frame_stride, level,
                          xy_inout + (size_t)2 * npoints * f0, levels + (size_t)npoints * f0, npoints, nrefined_out + f0, false)) return -1;
    }
    return 0;
}

API int mrg_b200_refine_chessboard_corners(const uint8_t* image, int Nrows, int Ncols, int stride, int level,
                                           double* xy_inout, signed char* levels, int Npoints)
{
    if (level < 0 || level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", level); return 0; }
    if (level == 0 && Nrows > 1 && stride != Ncols) { MSG("I can only handle continuous arrays (stride == width) currently."); return 0; }
    if (Npoints <= 0) return 0;
    DefaultLease L;
    if (!L.det) return -1;
    mrg_b200_detector* det = L.det;
    int32_t nref = 0;
    if (mrg_b200_refine_corners_batch(det, image, 0, 1, Nrows, Ncols, (size_t)stride, (size_t)stride * Nrows, level,
                                      xy_inout, levels, Npoints, &nref, nullptr)) return -1;
    return nref;
}

// ===========================================================================================
// Boards: corners (or blobs) -> grid -> refinement (mrgingham.cc:10-140)
// ===========================================================================================
namespace
{
// runs fn(i) for i in [0,n) on a few host threads
template <class F> void parallel_for(int n, F fn)
{
    const int nt = std::min<int>(std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), 16u), n / 4);
    if (nt <= 1) { for (int i = 0; i < n; i++) fn(i); return; }
    std::atomic<int> next(0);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&]() { for (int i; (i = next.fetch_add(1)) < n; ) fn(i); });
    for (auto& t : th) t.join();
}

// The frames idx[] of a device-resident chunk as one contiguous batch: in place when they are consecutive,
// otherwise copied (device to device, a few microseconds per frame) into the detector's gather buffer.
int gather_frames(mrg_b200_detector* det, cudaStream_t stream, const uint8_t* d_images, int rows, int cols, size_t pitch, size_t fstride,
                  const std::vector<int>& idx, const uint8_t** base, size_t* gpitch, size_t* gfstride)
{
    bool consecutive = true;
    for (size_t k = 1; k < idx.size(); k++) consecutive = consecutive && idx[k] == idx[k - 1] + 1;
    if (consecutive) { *base = d_images + (size_t)idx[0] * fstride; *gpitch = pitch; *gfstride = fstride; return 0; }
    DEVICE_GUARD(det);       // (also called from the board finder's helper threads)
    const size_t p = (size_t)round_up(cols, 16), fs = p * rows;
    if (det->boards_gather.ensure(fs * idx.size())) return -1;
    for (size_t k = 0; k < idx.size(); )
    {
        size_t e = k + 1;
        while (e < idx.size() && idx[e] == idx[e - 1] + 1 && fstride == pitch * rows) e++;
        CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)det->boards_gather.p + k * fs, p, d_images + (size_t)idx[k] * fstride, pitch, cols,
                                   (size_t)rows * (e - k), cudaMemcpyDeviceToDevice, stream));
        k = e;
    }
    *base = (const uint8_t*)det->boards_gather.p; *gpitch = p; *gfstride = fs;
    return 0;
}

// The reference's debug / debug_sequence of a one-image board call (mrgingham.cc:43-52): set by the entry point around
// its call, read by the grid search of find_boards_chunk(), which runs in the calling thread for a single frame.
thread_local const GridDebug* t_grid_debug = nullptr;
struct GridDebugScope
{
    GridDebug d; bool on;
    GridDebugScope(bool debug, int seq_x, int seq_y) : on(debug || (seq_x >= 0 && seq_y >= 0))
    {
        d.dump = debug; d.sequence = seq_x >= 0 && seq_y >= 0; d.seq_x = seq_x; d.seq_y = seq_y;
        if (on) t_grid_debug = &d;
    }
    ~GridDebugScope() { if (on) t_grid_debug = nullptr; }
};

// MRG_B200_BOARDS_TRACE=1: wall-clock time of every phase of a chunk on stderr (where a board call's time goes)
struct PhaseTrace
{
    bool on; std::chrono::steady_clock::time_point t;
    PhaseTrace() : on(getenv("MRG_B200_BOARDS_TRACE") != nullptr), t(std::chrono::steady_clock::now()) {}
    void mark(const char* what, int level, int frames)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "boards: %-14s level %d  %4d frames  %8.3f ms\n", what, level, frames, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

// One chunk of frames already on the device. level < 0: try levels 3,2,1,0 and keep the first that gives a grid
// (mrgingham.cc:127-138). strict_level0: a level-0 pass needs pitch == cols, as the reference's corner finder does.
int find_boards_chunk(mrg_b200_detector* det, const uint8_t* d_images, int n, int rows, int cols, size_t pitch, size_t fstride,
                      int gridn, int level, bool doblobs, bool refine, bool strict_level0,
                      double* xy_out, signed char* levels_out, int32_t* found_out, void* stream_)
{
    const int npts = gridn * gridn;
    cudaStream_t stream = stream_ ? (cudaStream_t)stream_ : det->own_stream;
    for (int i = 0; i < n; i++) found_out[i] = -1;
    std::vector<int32_t> xy, counts(n);
    std::vector<int> todo;
    const uint8_t* base; size_t gp, gfs;
    int mp = std::max(det->cfg.max_points, npts);      // output capacity of the corner passes below; grows here, never in det->cfg
    const int first_level = level < 0 ? 3 : level, last_level = level < 0 ? 0 : level;
    PhaseTrace trace;
    for (int L = first_level; L >= last_level; L--)
    {
        todo.clear();
        for (int i = 0; i < n; i++) if (found_out[i] < 0) todo.push_back(i);
        if (todo.empty()) break;
        if (L == 0 && !doblobs && strict_level0 && rows > 1 && pitch != (size_t)cols)
        { MSG("I can only handle continuous arrays (stride == width) currently."); break; }
        // the frames still without a grid, as one batch (results in the order of todo[])
        const int cnt = (int)todo.size();
        if (gather_frames(det, stream, d_images, rows, cols, pitch, fstride, todo, &base, &gp, &gfs)) return -1;
        for (;;)
        {
            xy.resize((size_t)2 * mp * cnt);
            const int rc = doblobs ? blobs_batch_mp(det, base, 1, cnt, rows, cols, gp, gfs, mp, xy.data(), counts.data(), stream)
                                   : corners_batch_mp(det, base, 1, cnt, rows, cols, gp, gfs, L, mp, xy.data(), counts.data(), stream);
            if (rc) return -1;
            int most = 0;
            for (int k = 0; k < cnt; k++) most = std::max(most, (int)counts[k]);
            if (most <= mp) break;
            mp = next_pow2(most);     // more points than room was made for: look again (a capacity private to this call)
        }
        trace.mark("corners", L, cnt);
        parallel_for(cnt, [&](int k)
        {
            const int i = todo[k];
            if (find_grid_from_points(xy.data() + (size_t)2 * mp * k, counts[k], gridn, xy_out + (size_t)2 * npts * i, cnt == 1 ? t_grid_debug : nullptr))
                found_out[i] = L;
        });
        trace.mark("grid search", L, cnt);
    }
    if (!refine || doblobs) return 0;

    // mrgingham.cc:81-99: every point starts at the level the grid was found at; refine level by level while
    // any point of the frame moves
    std::vector<signed char> own_levels;
    if (!levels_out) { own_levels.resize((size_t)npts * n); levels_out = own_levels.data(); }
    int top = 0;
    for (int i = 0; i < n; i++)
        if (found_out[i] >= 0) { top = std::max(top, found_out[i]); for (int k = 0; k < npts; k++) levels_out[(size_t)npts * i + k] = (signed char)found_out[i]; }
    std::vector<char> stopped(n, 0);
    std::vector<int32_t> nref(n);
    std::vector<double> txy;
    std::vector<signed char> tlv;
    for (int L = top - 1; L >= 0; L--)
    {
        todo.clear();
        for (int i = 0; i < n; i++) if (found_out[i] > L && !stopped[i]) todo.push_back(i);
        if (todo.empty()) break;
        const int cnt = (int)todo.size();
        if (gather_frames(det, stream, d_images, rows, cols, pitch, fstride, todo, &base, &gp, &gfs)) return -1;
        txy.resize((size_t)2 * npts * cnt); tlv.resize((size_t)npts * cnt);
        for (int k = 0; k < cnt; k++)
        {
            memcpy(&txy[(size_t)2 * npts * k], xy_out + (size_t)2 * npts * todo[k], sizeof(double) * 2 * npts);
            memcpy(&tlv[(size_t)npts * k], levels_out + (size_t)npts * todo[k], npts);
        }
        if (mrg_b200_refine_corners_batch(det, base, 1, cnt, rows, cols, gp, gfs, L, txy.data(), tlv.data(), npts, nref.data(), stream)) return -1;
        for (int k = 0; k < cnt; k++)
        {
            memcpy(xy_out + (size_t)2 * npts * todo[k], &txy[(size_t)2 * npts * k], sizeof(double) * 2 * npts);
            memcpy(levels_out + (size_t)npts * todo[k], &tlv[(size_t)npts * k], npts);
            if (nref[k] <= 0) stopped[todo[k]] = 1;
        }
        trace.mark("refine", L, cnt);
    }
    return 0;
}

// The same, for chunks of many frames, as a pipeline over two detectors (their streams and scratch): `det` runs the corner
// passes, `helper` the refinement passes, this thread's pool the grid searches.
//   * A pass covers ALL frames of the chunk, in place, whenever the frames it is needed for are a fair share of them:
//     gathering a scattered subset costs more than the pyramid + ChESS + clustering of the frames nobody asked about
//     (a level-L image is 4^-L of a frame; the gather copies whole frames). Results of frames that already have a grid
//     are ignored; their points carry level 0 in a refinement pass, which then leaves them alone.
//   * So the corner pass of level L-1 does not depend on the grid search of level L: it is enqueued before that search
//     starts and runs beside it (level 0, a full-resolution pass, only when some frame really needs it).
//   * Refinement at level L needs the searches of the levels above L and the refinement at L+1: it runs on the helper,
//     on its own host thread, beside the search of level L.
// Results are those of find_boards_chunk() (same passes over the same pixels, frame by frame).
int find_boards_chunk_pipelined(mrg_b200_detector* det, mrg_b200_detector* helper, const uint8_t* d_images, int n, int rows, int cols,
                                size_t pitch, size_t fstride, int gridn, int level, bool strict_level0,
                                double* xy_out, signed char* levels_out, int32_t* found_out)
{
    const int npts = gridn * gridn;
    cudaStream_t stream = det->own_stream;
    for (int i = 0; i < n; i++) found_out[i] = -1;
    int mp = std::max(det->cfg.max_points, npts);
    const int first_level = level < 0 ? 3 : level, last_level = level < 0 ? 0 : level;
    PhaseTrace trace;
    std::vector<signed char> own_levels;
    if (!levels_out) { own_levels.resize((size_t)npts * n); levels_out = own_levels.data(); }
    std::vector<char> stopped(n, 0);

    // ---- corner passes on det ----
    // Results: the points of frame k of the pass at pts[2 * off[k] ...], counts[k] of them -- copied out of the detector's
    // pinned result buffer compactly (the next pass, which starts before these are looked at, writes there again; a slot
    // of max_points per frame would be megabytes per pass to copy)
    std::vector<int32_t> pts, counts;
    std::vector<size_t> off;
    std::vector<int> pass_frames;                  // frames of the pass in flight, in result order (empty: all n, in place)
    auto enqueue = [&](int L, const std::vector<int>& only) -> int
    {
        const uint8_t* base = d_images; size_t gp = pitch, gfs = fstride; int cnt = n;
        pass_frames.clear();
        if (!only.empty() && (int)only.size() * 2 < n)
        {
            if (gather_frames(det, stream, d_images, rows, cols, pitch, fstride, only, &base, &gp, &gfs)) return -1;
            pass_frames = only; cnt = (int)only.size();
        }
        std::lock_guard<std::mutex> g(det->mtx);
        return enqueue_locked(det, base, 1, cnt, rows, cols, gp, gfs, L, mp, stream);
    };
    auto collect = [&](int L) -> int
    {
        for (;;)
        {
            const int cnt = pass_frames.empty() ? n : (int)pass_frames.size();
            counts.resize(cnt);
            {
                std::lock_guard<std::mutex> g(det->mtx);
                if (collect_locked(det, nullptr, counts.data())) return -1;
            }
            int most = 0;
            for (int k = 0; k < cnt; k++) most = std::max(most, (int)counts[k]);
            if (most <= mp)
            {
                off.assign(cnt + 1, 0);
                for (int k = 0; k < cnt; k++) off[k + 1] = off[k] + (size_t)std::max(counts[k], 0);
                pts.resize(2 * off[cnt]);
                const int32_t* hxy = (const int32_t*)det->h_xy.p;
                for (int k = 0; k < cnt; k++)
                    if (counts[k] > 0) memcpy(&pts[2 * off[k]], hxy + (size_t)2 * mp * k, sizeof(int32_t) * 2 * counts[k]);
                return 0;
            }
            mp = next_pow2(most);                  // more points than room was made for: the same pass again
            const std::vector<int> again = pass_frames;
            if (enqueue(L, again)) return -1;
        }
    };

    // ---- refinement passes on helper (mrgingham.cc:81-99: level by level while any point of the frame moves) ----
    auto refine_pass = [&](int L, std::vector<int> who) -> int
    {
        const int cnt = (int)who.size();
        const bool all = cnt * 4 >= n || L >= 2;
        const uint8_t* base = d_images; size_t gp = pitch, gfs = fstride;
        if (!all && gather_frames(helper, helper->own_stream, d_images, rows, cols, pitch, fstride, who, &base, &gp, &gfs)) return -1;
        const int m = all ? n : cnt;
        std::vector<double> txy((size_t)2 * npts * m, 0.0);
        std::vector<signed char> tlv((size_t)npts * m, 0);           // level 0: "already refined", never touched
        std::vector<int32_t> nref(m);
        for (int k = 0; k < cnt; k++)
        {
            const size_t slot = all ? who[k] : k;
            memcpy(&txy[2 * npts * slot], xy_out + (size_t)2 * npts * who[k], sizeof(double) * 2 * npts);
            memcpy(&tlv[npts * slot], levels_out + (size_t)npts * who[k], npts);
        }
        if (mrg_b200_refine_corners_batch(helper, base, 1, m, rows, cols, gp, gfs, L, txy.data(), tlv.data(), npts, nref.data(), nullptr)) return -1;
        for (int k = 0; k < cnt; k++)
        {
            const size_t slot = all ? who[k] : k;
            memcpy(xy_out + (size_t)2 * npts * who[k], &txy[2 * npts * slot], sizeof(double) * 2 * npts);
            memcpy(levels_out + (size_t)npts * who[k], &tlv[npts * slot], npts);
            if (nref[slot] <= 0) stopped[who[k]] = 1;
        }
        return 0;
    };
    std::thread refine_thread;
    int refine_rc = 0, rl = first_level - 1;       // rl: the next refinement level
    auto join_refine = [&]() { if (refine_thread.joinable()) refine_thread.join(); return refine_rc; };
    auto refine_who = [&](int L) { std::vector<int> who; for (int i = 0; i < n; i++) if (found_out[i] > L && !stopped[i]) who.push_back(i); return who; };

    int rc = 0;
    std::vector<int> todo;
    bool in_flight = false;
    if (!(first_level == 0 && strict_level0 && rows > 1 && pitch != (size_t)cols))
    { if (enqueue(first_level, todo)) return -1; in_flight = true; }
    else MSG("I can only handle continuous arrays (stride == width) currently.");
    for (int L = first_level; L >= last_level && in_flight; L--)
    {
        if (collect(L)) { rc = -1; break; }
        in_flight = false;
        trace.mark("corners", L, pass_frames.empty() ? n : (int)pass_frames.size());
        // what the search of this level looks at; the next level's pass starts before it
        std::vector<int> frames = pass_frames;                       // frames of xy/counts, in order
        if (frames.empty()) { frames.resize(n); for (int i = 0; i < n; i++) frames[i] = i; }
        const std::vector<int32_t> cur_pts = std::move(pts), cur_counts = std::move(counts);
        const std::vector<size_t> cur_off = std::move(off);
        todo.clear();
        for (int i = 0; i < n; i++) if (found_out[i] < 0) todo.push_back(i);
        const bool speculate = L - 1 >= last_level && L - 1 >= 1 && !todo.empty();
        if (speculate) { if (enqueue(L - 1, std::vector<int>())) { rc = -1; break; } in_flight = true; }
        if (rl == L)
        {
            if (join_refine()) { rc = -1; break; }
            std::vector<int> who = refine_who(rl);
            if (!who.empty()) { const int lv = rl; refine_thread = std::thread([&, lv, who]() { refine_rc = refine_pass(lv, who); }); }
            rl--;
        }
        std::vector<int> slot_of(n, -1);
        for (size_t k = 0; k < frames.size(); k++) slot_of[frames[k]] = (int)k;
        const int ntodo = (int)todo.size();
        parallel_for(ntodo, [&](int t)
        {
            const int i = todo[t], k = slot_of[i];
            if (k < 0) return;
            if (find_grid_from_points(cur_pts.data() + 2 * cur_off[k], cur_counts[k], gridn, xy_out + (size_t)2 * npts * i))
            {
                for (int q = 0; q < npts; q++) levels_out[(size_t)npts * i + q] = (signed char)L;
                found_out[i] = L;
            }
        });
        trace.mark("grid search", L, ntodo);
        todo.clear();
        for (int i = 0; i < n; i++) if (found_out[i] < 0) todo.push_back(i);
        if (todo.empty() || L == last_level) break;
        if (!in_flight)
        {
            // level 0: full-resolution frames, only those that still have no grid
            if (L - 1 == 0 && strict_level0 && rows > 1 && pitch != (size_t)cols)
            { MSG("I can only handle continuous arrays (stride == width) currently."); break; }
            if (enqueue(L - 1, todo)) { rc = -1; break; }
            in_flight = true;
        }
    }
    if (in_flight)
    {
        // a pass that ran ahead of a search that then found every grid: its results are not needed
        std::lock_guard<std::mutex> g(det->mtx);
        if (collect_locked(det, nullptr, nullptr)) rc = -1;
    }
    if (join_refine()) rc = -1;
    trace.mark("refine (tail)", rl + 1, 0);
    for (; rc == 0 && rl >= 0; rl--)
    {
        std::vector<int> who = refine_who(rl);
        if (who.empty()) continue;
        if (refine_pass(rl, who)) rc = -1;
        trace.mark("refine", rl, (int)who.size());
    }
    return rc;
}

int find_boards(mrg_b200_detector* det, const uint8_t* images, int on_device, int nframes, int rows, int cols,
                size_t pitch, size_t fstride, int gridn, int level, bool doblobs, bool refine, bool strict_level0,
                double* xy_out, signed char* levels_out, int32_t* found_out, void* stream_)
{
    if (!det) return -1;
    if (nframes < 0 || rows <= 0 || cols <= 0 || rows > 32767 || cols > 32767 || pitch < (size_t)cols) { MSG("Bad batch geometry."); return -1; }
    if (gridn < 2 || gridn > 181) { MSG("gridn must be in [2,181]; got %d.", gridn); return -1; }
    if (level > 10) { MSG("Got an unreasonable image_pyramid_level = %d.", level); for (int i = 0; i < nframes; i++) found_out[i] = -1; return 0; }
    if (doblobs && level != 0) { MSG("The blob detector works at image_pyramid_level 0 only."); return -1; }
    std::lock_guard<std::mutex> g(det->boards_mtx);
    const int npts = gridn * gridn;
    // Between the GPU passes of a chunk (corners at one level, refinement at one level) the host looks for grids,
    // and the GPU has nothing to do. So when the batch is longer than one chunk (max_frames), two detectors -- this
    // one and a helper it owns, with its own streams and scratch -- take the chunks in turns on two host threads:
    // one's grid search runs beside the other's kernels (1024 4K frames, 14x14: 15.6 k boards/s against 11.9 k for
    // one chunk at a time; cutting a single chunk into smaller ones for this loses more than it wins).
    const int chunk = std::max(1, det->cfg.max_frames);
    const int nchunks = (nframes + chunk - 1) / chunk;
    // chunks of many frames with refinement: the pipeline over both detectors inside every chunk (chunks one after another)
    const bool pipelined = !doblobs && refine && std::min(chunk, nframes) >= 16 && !getenv("MRG_B200_BOARDS_SERIAL");
    if ((nchunks >= 2 || pipelined) && !det->boards_helper)
    {
        mrg_b200_detector_config hc = det->cfg;
        hc.device = det->device;
        if (mrg_b200_detector_create(&det->boards_helper, &hc)) det->boards_helper = nullptr;       // (then: one detector, as before)
    }
    const bool piped = pipelined && det->boards_helper;
    if (stream_ && (nchunks >= 2 || piped) && det->boards_helper)
    {
        DEVICE_GUARD(det);
        CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream_));       // device frames may come from work queued there; the helper runs elsewhere
    }
    auto run_chunks = [&](mrg_b200_detector* dd, int first, int step, void* dstream) -> int
    {
        for (int c = first; c < nchunks; c += step)
        {
            const int f0 = c * chunk, n = std::min(chunk, nframes - f0);
            const uint8_t* d = images + (size_t)f0 * fstride;
            size_t dpitch = pitch, dfstride = fstride;
            if (!on_device)
            {
                // the frames go to the device once and stay there for every level and refinement pass
                DEVICE_GUARD(dd);
                cudaStream_t stream = dstream ? (cudaStream_t)dstream : dd->own_stream;
                dpitch = pitch == (size_t)cols ? (size_t)cols : (size_t)round_up(cols, 16);
                dfstride = dpitch * rows;
                if (dd->boards_frames.ensure(dfstride * n)) return -1;
                if (dd->stage_via_pinned)
                {
                    if (dd->h_stage.ensure(dfstride * n)) return -1;
                    if (!dd->h_stage_free) CUDA_TRY(cudaEventCreateWithFlags(&dd->h_stage_free, cudaEventDisableTiming));
                    else                   CUDA_TRY(cudaEventSynchronize(dd->h_stage_free));
                    uint8_t* hp = (uint8_t*)dd->h_stage.p;
                    for (int i = 0; i < n; i++)
                        for (int y = 0; y < rows; y++) memcpy(hp + i * dfstride + (size_t)y * dpitch, d + i * fstride + (size_t)y * pitch, cols);
                    CUDA_TRY(cudaMemcpyAsync(dd->boards_frames.p, hp, dfstride * n, cudaMemcpyHostToDevice, stream));
                    CUDA_TRY(cudaEventRecord(dd->h_stage_free, stream));
                }
                else if (dpitch == pitch && fstride == dfstride)
                    CUDA_TRY(cudaMemcpyAsync(dd->boards_frames.p, d, dfstride * n, cudaMemcpyHostToDevice, stream));
                else
                    for (int i = 0; i < n; i++)
                        CUDA_TRY(cudaMemcpy2DAsync((uint8_t*)dd->boards_frames.p + i * dfstride, dpitch, d + i * fstride, pitch, cols, rows,
                                                   cudaMemcpyHostToDevice, stream));
                CUDA_TRY(cudaStreamSynchronize(stream));
                d = (const uint8_t*)dd->boards_frames.p;
            }
            if (piped && n >= 16)
            {
                if (find_boards_chunk_pipelined(dd, det->boards_helper, d, n, rows, cols, dpitch, dfstride, gridn, level, strict_level0,
                                                xy_out + (size_t)2 * npts * f0, levels_out ? levels_out + (size_t)npts * f0 : nullptr, found_out + f0)) return -1;
            }
            else if (find_boards_chunk(dd, d, n, rows, cols, dpitch, dfstride, gridn, level, doblobs, refine, strict_level0,
                                  xy_out + (size_t)2 * npts * f0, levels_out ? levels_out + (size_t)npts * f0 : nullptr, found_out + f0, dstream)) return -1;
        }
        return 0;
    };
    if (piped) return run_chunks(det, 0, 1, nullptr);
    if (nchunks >= 2 && det->boards_helper)
    {
        int rc_b = 0;
        std::thread tb([&] { rc_b = run_chunks(det->boards_helper, 1, 2, nullptr); });
        const int rc_a = run_chunks(det, 0, 2, nullptr);
        tb.join();
        return rc_a || rc_b ? -1 : 0;
    }
    return run_chunks(det, 0, 1, stream_);
}
}   // namespace

API int mrg_b200_find_grid_from_points(const int* xy, int npoints, int gridn, double* xy_out)
{
    if (gridn < 2 || npoints < 0 || (npoints > 0 && !xy) || !xy_out) return 0;
    return find_grid_from_points(xy, npoints, gridn, xy_out) ? 1 : 0;
}

API int mrg_b200_find_grid_from_points_debug(const int* xy, int npoints, int gridn, double* xy_out,
                                             int debug, int debug_sequence_x, int debug_sequence_y)
{
    if (gridn < 2 || npoints < 0 || (npoints > 0 && !xy) || !xy_out) return 0;
    GridDebug d; d.dump = debug != 0; d.sequence = debug_sequence_x >= 0 && debug_sequence_y >= 0;
    d.seq_x = debug_sequence_x; d.seq_y = debug_sequence_y;
    return find_grid_from_points(xy, npoints, gridn, xy_out, &d) ? 1 : 0;
}

API int mrg_b200_voronoi_neighbours(const int* xy, int npoints, int* ring_off, int* ring, int ring_cap)
{
    if (npoints <= 0 || !xy || !ring_off || (ring_cap > 0 && !ring)) return -1;
    return voronoi_neighbours(xy, npoints, ring_off, ring, ring_cap);
}

API int mrg_b200_find_boards_batch(mrg_b200_detector* det, const uint8_t* images, int images_on_device,
                                   int nframes, int rows, int cols, size_t row_pitch, size_t frame_stride,
                                   int gridn, int image_pyramid_level, int doblobs, int refine,
                                   double* xy_out, signed char* levels_out, int32_t* found_level_out, void* stream)
{
    return find_boards(det, images, images_on_device, nframes, rows, cols, row_pitch, frame_stride, gridn, image_pyramid_level,
                       doblobs != 0, refine != 0, false, xy_out, levels_out, found_level_out, stream);
}

// Return values of the two one-image board calls: the reference's (level found / 1), "no grid" (-1 / 0), and,
// kept apart from "no grid", -2 = the GPU path failed.
API int mrg_b200_find_chessboard_from_image_array(const uint8_t* image, int Nrows, int Ncols, int stride,
                                                  int gridn, int image_pyramid_level, int refine,
                                                  double* xy_out, signed char* levels_out)
{
    if (Nrows <= 0 || Ncols <= 0 || stride < Ncols) { MSG("Bad image geometry."); return -1; }
    DefaultLease L;
    if (!L.det) return -2;
    int32_t found = -1;
    if (find_boards(L.det, image, 0, 1, Nrows, Ncols, (size_t)stride, (size_t)stride * Nrows, gridn, image_pyramid_level,
                    false, refine != 0, true, xy_out, levels_out, &found, nullptr)) return -2;
    return found;
}

API int mrg_b200_find_circle_grid_from_image_array(const uint8_t* image, int Nrows, int Ncols, int stride,
                                                   int gridn, double* xy_out)
{
    if (Nrows <= 0 || Ncols <= 0 || stride < Ncols) { MSG("Bad image geometry."); return 0; }
    DefaultLease L;
    if (!L.det) return -2;
    int32_t found = -1;
    if (find_boards(L.det, image, 0, 1, Nrows, Ncols, (size_t)stride, (size_t)stride * Nrows, gridn, 0,
                    true, false, true, xy_out, nullptr, &found, nullptr)) return -2;
    return found >= 0 ? 1 : 0;
}

// mrgingham_pywrap_cplusplus_bridge.cc:72-138, what the reference's Python find_board() binds
API bool find_chessboard_from_image_array_C(int Nrows, int Ncols, int stride, char* imagebuffer, const int gridn,
                                            int image_pyramid_level, bool doblobs, bool debug,
                                            int debug_sequence_x, int debug_sequence_y,
                                            bool (*add_points)(double* xy, int N, void* cookie), void* cookie)
{
    if (gridn < 2) return false;
    // debug: the grid finder's /tmp dumps and messages; debug_sequence_x/y >= 0: its trace of the walks from the point
    // nearest to that pixel (bridge.cc:97-104). The corner-level artefacts of `debug` come from mrg_b200_debug_dump_corners().
    GridDebugScope grid_debug(debug, debug_sequence_x, debug_sequence_y);
    if (debug && !doblobs && image_pyramid_level >= 0)
        mrg_b200_debug_dump_corners((const uint8_t*)imagebuffer, Nrows, Ncols, stride, image_pyramid_level, nullptr, nullptr, nullptr, 0);
    std::vector<double> xy((size_t)2 * gridn * gridn);
    int rc;
    if (doblobs)
    {
        if (image_pyramid_level != 0) return false;
        rc = mrg_b200_find_circle_grid_from_image_array((const uint8_t*)imagebuffer, Nrows, Ncols, stride, gridn, xy.data());
        if (rc == 0) return false;
    }
    else
    {
        rc = mrg_b200_find_chessboard_from_image_array((const uint8_t*)imagebuffer, Nrows, Ncols, stride, gridn, image_pyramid_level, 1,
                                                       xy.data(), nullptr);
        if (rc == -1) return false;
    }
    if (rc < 0)
    {
        // The reference's Python wrapper reads any false as "no chessboard" (mrgingham_pywrap.c:300-310), so a GPU
        // failure cannot be told apart through this symbol: it is made loud here instead of passing for "no board".
        MSG("The GPU board finder FAILED (CUDA error or no device); this is NOT 'no chessboard found'.");
        return false;
    }
    return (*add_points)(xy.data(), gridn * gridn, cookie);
}
