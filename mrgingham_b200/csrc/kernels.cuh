// Internal declarations shared by the kernel translation units and the host API (api.cu).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace mrgb200
{
// Detector constants of the reference (find_chessboard_corners.cc:18,22,27,29,38,39,564).
// Bit-exactness depends on them, so they are compile-time constants here as they are there.
constexpr int kPeakMin        = 120;  // RESPONSE_MIN_PEAK_THRESHOLD
constexpr int kRespMin        = 15;   // RESPONSE_MIN_THRESHOLD
constexpr int kComponentMinN  = 2;    // CONNECTED_COMPONENT_MIN_SIZE
constexpr int kVarWindowR     = 10;   // CONSTANCY_WINDOW_R
constexpr int kVarMin         = 400;  // STDEV_THRESHOLD^2
constexpr int kMargin         = 7;    // ChESS leaves a 7-pixel border unwritten
constexpr double kFindGridScale = 1000.0; // FIND_GRID_SCALE, mrgingham-internal.h:3

// A candidate = one pixel whose clamped ChESS response exceeds kRespMin, inside
// [7,w-7) x [7,h-7):   bits 47..32 = y, 31..16 = x, 15..0 = response.
// Numeric order of the word == raster order of the pixel.
typedef unsigned long long cand_t;
__host__ __device__ inline cand_t   cand_pack(int x, int y, int r) { return ((cand_t)y << 32) | ((cand_t)x << 16) | (cand_t)(unsigned)r; }
__host__ __device__ inline int      cand_x(cand_t c)   { return (int)((c >> 16) & 0xFFFF); }
__host__ __device__ inline int      cand_y(cand_t c)   { return (int)(c >> 32); }
__host__ __device__ inline int      cand_r(cand_t c)   { return (int)(c & 0xFFFF); }
__host__ __device__ inline uint32_t cand_key(cand_t c) { return (uint32_t)(c >> 16); }

// One level image of one batch, as the kernels see it
struct FrameSet
{
    const uint8_t* base;      // first frame
    size_t frame_stride;      // bytes between frames
    int    pitch;             // bytes between rows
    int    w, h;              // pixels
    int    nframes;
};

// Output of the clustering kernel per accepted component, before ordering/finalisation is
// done in the same kernel. Kept for debugging entry points.
struct ClusterParams
{
    int      level;           // pyramid level the FrameSet is at (coordinates scale by 2^level)
    int      cand_capacity;   // per-frame capacity of the candidate lists (power of two)
    int      max_points;      // per-frame capacity of the outputs
    int      record_capacity; // per-frame capacity of `records`
    void*    records;         // [nframes][record_capacity] scratch, cluster_record_bytes() each
    int      smem_cands;      // candidates per frame the clustering CTA keeps in shared memory: 1024, 2048 or
                              // 4096 (anything else = 4096), 22 bytes each. Longer lists use the global scratch.
};
size_t cluster_record_bytes();

// ---- launchers (each enqueues on `stream`, returns cudaGetLastError()) ----

// K0: level-L image from the full-resolution frames (model N1 of cv::resize INTER_LINEAR)
cudaError_t launch_pyramid(const FrameSet& src, int level, uint8_t* dst, int dst_pitch,
                           size_t dst_frame_stride, int ow, int oh, cudaStream_t stream);

// Kpre: cv::blur(Size(1+2R,1+2R)) of the reference CLI's default preprocessing (mrgingham-from-image.cc:106-111)
cudaError_t launch_box_blur(const FrameSet& src, int radius, uint8_t* dst, int dst_pitch, size_t dst_frame_stride,
                            cudaStream_t stream);

// Kpre2: cv::normalize(0,255,NORM_MINMAX) (if `normalize`) then cv::CLAHE(clip_limit, 8x8 tiles), the reference
// CLI's --clahe (mrgingham-from-image.cc:43-44, 71-80). scratch: clahe_scratch_bytes(nframes) bytes.
size_t clahe_scratch_bytes(int nframes);
cudaError_t launch_normalize_clahe(const FrameSet& fs, bool normalize, double clip_limit, uint8_t* dst, int dst_pitch,
                                   size_t dst_frame_stride, void* scratch, cudaStream_t stream);

// device-to-device gather of n equally-sized images lying anywhere into one contiguous, pitched batch
struct GatherSrc { const uint8_t* data; size_t pitch; };
cudaError_t launch_gather_frames(const GatherSrc* srcs, int n, int rows, int cols, uint8_t* dst, size_t dst_pitch,
                                 size_t dst_frame_stride, cudaStream_t stream);

// Kpre3: 16-bit frames -> 8-bit frames as the reference CLI does it (mrgingham-from-image.cc:83-93):
// [normalize(0,65535) + CLAHE(8) on 16 bits if clahe], then convertTo(CV_8U, 255/65535). preproc16.cu
size_t preproc16_scratch_bytes(int nframes);
cudaError_t launch_preprocess16(const uint16_t* src, size_t src_frame_stride_elems, int src_pitch_elems, int w, int h, int nframes,
                                bool clahe, uint8_t* dst, int dst_pitch, size_t dst_frame_stride, void* scratch, cudaStream_t stream);

// K1 (simple variant): ChESS response, one thread per pixel, emits candidates
cudaError_t launch_chess_sparse_simple(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                       int cand_capacity, cudaStream_t stream);
// K1 (tiled variant): TMA/shared-memory staged, packed-lane arithmetic
cudaError_t launch_chess_sparse_tiled(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                      int cand_capacity, cudaStream_t stream);
// K1 (cascade variant, production): TMA-staged, byte-lane filter cascade. *launched = false (and
// nothing enqueued) when the frames do not meet TMA's alignment rules.
cudaError_t launch_chess_sparse_cascade(const FrameSet& fs, cand_t* cand, uint32_t* counts,
                                        int cand_capacity, cudaStream_t stream, bool* launched);
// dense int16 response (interior only), for the ChESS_response_5 API
cudaError_t launch_chess_dense(const FrameSet& fs, int16_t* response, size_t response_frame_stride_elems,
                               cudaStream_t stream);

// the same through the tiled kernel's TMA ring and packed 16-bit lanes (production path of the dense API;
// launch_chess_dense above is the one-thread-per-pixel cross-check)
cudaError_t launch_chess_dense_tiled(const FrameSet& fs, int16_t* response, size_t response_frame_stride_elems,
                                     cudaStream_t stream);

// K2: per-frame clustering of the candidate lists, exact emulation of
// process_connected_components()'s find branch. One CTA per frame.
//   cand/counts      from K1 (lists are sorted in place)
//   scratch_table    [nframes][2*cand_capacity] uint32, only touched for frames whose list does
//                    not fit shared memory
//   scratch_dfs      [nframes][cand_capacity]   uint32, idem
//   xy_int           [nframes][max_points][2] int32  (scaled by 1000)
//   xy_dbl           [nframes][max_points][2] double (un-quantised), may be NULL
//   out_counts       [nframes] int32: corners found; -1 = candidate list overflowed
cudaError_t launch_cluster_find(const FrameSet& fs, const ClusterParams& p,
                                cand_t* cand, const uint32_t* counts,
                                uint32_t* scratch_table, uint32_t* scratch_dfs,
                                int32_t* xy_int, double* xy_dbl, int32_t* out_counts,
                                cudaStream_t stream);

// K2r: refinement branch. points/levels are [nframes][npoints] (in/out), out_refined [nframes]
// (-1 = candidate list overflowed)
cudaError_t launch_cluster_refine(const FrameSet& fs, const ClusterParams& p,
                                  cand_t* cand, const uint32_t* counts,
                                  uint32_t* scratch_table, uint32_t* scratch_dfs,
                                  double* points_xy, signed char* levels, int npoints,
                                  int32_t* out_refined, cudaStream_t stream);

// Blob path (find_blobs.cc:14-46 = cv::SimpleBlobDetector): blobs.cu. The workspace owns the device
// scratch (bit planes, mark planes, border points, records) and grows it on demand.
struct BlobWorkspace;
BlobWorkspace* blob_workspace_create();
void           blob_workspace_destroy(BlobWorkspace* ws);
// xy_out: HOST int32 [nframes][max_points][2] scaled by 1000, counts_out: HOST int32 [nframes].
// Synchronous on `stream`. ms_out (optional): device time of the three kernels (CUDA events).
// Returns 0; -1 on a CUDA failure; 1 if the scratch of a multi-frame chunk overflowed (nothing was
// produced: call again frame by frame, then blob_workspace_reset_capacity()).
void blob_workspace_reset_capacity(BlobWorkspace* ws);
// The two halves of blob_find_frames for callers that keep two chunks in flight (one workspace each): enqueue the
// kernels of a chunk on a stream; later wait for it and group its records on the host (same return values; ms is
// added to *ms_out).
int blob_enqueue(BlobWorkspace* ws, const FrameSet& fs, cudaStream_t stream);
int blob_finish(BlobWorkspace* ws, int32_t* xy_out, int32_t* counts_out, int max_points, float* ms_out);
int blob_find_frames(BlobWorkspace* ws, const FrameSet& fs, int32_t* xy_out, int32_t* counts_out, int max_points,
                     cudaStream_t stream, float* ms_out);

// debug_dump.cu: the reference's --debug artefacts (host code)
bool write_png_gray8(const char* path, const uint8_t* data, int w, int h, size_t pitch);
void normalize_response_u8(const int16_t* resp, size_t n, uint8_t* out);
bool write_corner_vnl(const char* path, const char* debug_image_filename, const double* xy, int n);

// largest shared-memory candidate capacity of the clustering kernel (ClusterParams::smem_cands)
constexpr int kClusterSmemCands = 4096;
}
