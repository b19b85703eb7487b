/*
  mrgingham_b200 -- C ABI of the B200-native (sm_100a) chessboard-corner detector.

  This library replaces ONE hot path of dkogan/mrgingham: the per-image corner detector
  (ChESS response -> pyramid level -> connected-component clustering -> sub-pixel centroid).
  All compute runs in hand-written CUDA kernels; there is no CPU fallback: every entry point
  fails (returns false / <0 and prints a diagnostic) when no CUDA device is usable.

  Section A re-exports, with identical names, signatures and error behaviour, the C symbols the
  reference library already exposes for this path, so a build of the reference can link this
  library in place of ChESS.c + find_chessboard_corners.cc.
  Section B is the C mirror of the reference's C++ API for the path (the C++ spellings live in
  include/mrgingham_b200/find_chessboard_corners.hh as inline adapters over these).
  Section C is additive: batched, device-resident entry points that the reference has no
  equivalent of (precedent for batch semantics: the Python ChESS_response_5 broadcasts over
  leading dimensions, mrgingham_pywrap.c:84-103).

  Conventions (same as the reference): 8-bit single-channel images, rows
  contiguous, `stride` in bytes; the caller owns every buffer; no exceptions cross this ABI;
  diagnostics go to stderr as "file:line in func(): ... Sorry.". All functions are thread-safe, and none
  changes the calling thread's current CUDA device. The one-image calls (sections A and B) run on the calling
  thread's current device, each on a detector borrowed from a per-device pool, so concurrent callers overlap.
*/
#ifndef MRGINGHAM_B200_H
#define MRGINGHAM_B200_H

#include <stdint.h>
#include <stddef.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===========================================================================================
   A. Drop-in symbols (names and signatures of the reference)
   =========================================================================================== */

/* Replaces ChESS.c:55-106 (declared in ChESS.h:31-34).
   response: dense int16 [h][w] HOST buffer; image: uint8 HOST buffer with row pitch `stride`.
   Writes 7 <= x < w-7, 7 <= y < h-7 only; every other element of `response` is left untouched,
   exactly as the reference does. The signature has no way to report a failure: if the GPU path fails (no
   device, CUDA error) a diagnostic goes to stderr and `response` is not (or only partly) written; the host
   process is never aborted. */
void mrgingham_ChESS_response_5(int16_t*       response,
                                const uint8_t* image,
                                int w, int h, int stride);

/* Replaces mrgingham_pywrap_cplusplus_bridge.cc:28-70 (declared in
   mrgingham_pywrap_cplusplus_bridge.h:10-23): the C bridge the reference's Python module binds.
   Returns false when nothing was found or on error (doblobs with a level other than 0 is one:
   ...bridge.cc:50-56); otherwise calls add_points(xy, N, 1/1000., cookie) once and returns its
   result. A failure of the GPU path is NOT "nothing found": add_points(xy, 0, ...) is called and false
   returned, which the reference's Python wrapper turns into RuntimeError (mrgingham_pywrap.c:203-219). doblobs selects the blob detector (find_blobs.cc:14-46) instead of the corner detector. */
bool find_chessboard_corners_from_image_array_C(int Nrows, int Ncols,
                                                int stride,
                                                char* imagebuffer, /* const */
                                                int image_pyramid_level,
                                                bool doblobs,
                                                bool debug,
                                                bool (*add_points)(int* xy, int N, double scale, void* cookie),
                                                void* cookie);

/* Replaces mrgingham_pywrap_cplusplus_bridge.cc:72-138 (declared in ...bridge.h:25-42): what the reference's
   Python find_board() binds. Corners (or blobs) -> grid of gridn x gridn -> refinement; image_pyramid_level < 0
   tries levels 3,2,1,0 and keeps the first that yields a grid. Returns false when no grid was found, on error,
   or for doblobs with a level other than 0; otherwise calls add_points(xy, gridn*gridn, cookie) once and returns
   its result. debug: the grid finder's /tmp dumps and messages (mrg_b200_find_grid_from_points_debug()) and, for a
   given level, the corner-level artefacts of mrg_b200_debug_dump_corners(); debug_sequence_x/y >= 0: the grid
   finder's stderr trace of the walks from the corner nearest to that pixel (bridge.cc:97-104). */
bool find_chessboard_from_image_array_C(int Nrows, int Ncols,
                                        int stride,
                                        char* imagebuffer, /* const */
                                        const int gridn,
                                        int image_pyramid_level,
                                        bool doblobs,
                                        bool debug,
                                        int debug_sequence_x,
                                        int debug_sequence_y,
                                        bool (*add_points)(double* xy, int N, void* cookie),
                                        void* cookie);

/* ===========================================================================================
   B. C mirror of the reference's C++ API for the path
   =========================================================================================== */

/* mrgingham::find_chessboard_corners_from_image_array(), find_chessboard_corners.cc:568-587.
   Returns the number of corners found (0 on any of the reference's error paths: level outside
   [0,10], level 0 with stride != Ncols) or <0 on a CUDA failure; writes min(N, cap) points into
   xy_out as (x,y) pairs scaled by 1000 (mrgingham-internal.h:3), in the reference's order. */
int mrg_b200_find_chessboard_corners(const uint8_t* image, int Nrows, int Ncols, int stride,
                                     int image_pyramid_level,
                                     int* xy_out, int cap);

/* The reference's --debug artefacts of ONE corner-detector call (find_chessboard_corners.cc:294-315, 346-348,
   391-392, 400-407, 452-459, 513-541): /tmp/mrgingham-scaled-processed-level%d.png, /tmp/mrgingham-chess-response
   [-refinement]-level%d[-positive].png (8-bit PNGs with the pixel values cv::normalize + cv::imwrite give) and the
   self-plotting corner list /tmp/mrgingham-1-corners.vnl (find branch: refined_xy == NULL; the corners as un-quantised
   doubles, "%f %f") or /tmp/mrgingham-1-corners-refinement-level%d.vnl (refinement branch: pass the arrays as they are
   AFTER the refinement call; the points whose level equals image_pyramid_level are listed). The C++ adapters and the
   bridge symbol call this when their `debug` argument is set. Returns 0 or <0. */
int mrg_b200_debug_dump_corners(const uint8_t* image, int Nrows, int Ncols, int stride,
                                int image_pyramid_level, const char* debug_image_filename,
                                const double* refined_xy, const signed char* refined_levels, int Npoints);

/* mrgingham::refine_chessboard_corners_from_image_array(), find_chessboard_corners.cc:591-619.
   xy_inout: Npoints (x,y) doubles in full-resolution pixels, level[]: per-point pyramid level;
   points with level[i] == image_pyramid_level+1 are refined in place and their level lowered.
   Returns the number of points refined, <0 on a CUDA failure. */
int mrg_b200_refine_chessboard_corners(const uint8_t* image, int Nrows, int Ncols, int stride,
                                       int image_pyramid_level,
                                       double* xy_inout, signed char* level, int Npoints);

/* mrgingham::find_blobs_from_image_array(), find_blobs.cc:14-46: cv::SimpleBlobDetector with
   minArea=20, maxArea=80000, minDistBetweenBlobs=5, blobColor=0, OpenCV defaults otherwise
   (behaviour pinned to OpenCV 4.13.0, which is what oracle/blob_oracle.c is checked against).
   Returns the number of blobs (<0 on a CUDA failure); writes min(N, cap) points into xy_out as
   (x,y) pairs scaled by 1000, in the reference's keypoint order. */
int mrg_b200_find_blobs(const uint8_t* image, int Nrows, int Ncols, int stride,
                        int* xy_out, int cap);

/* mrgingham::find_grid_from_points(), find_grid.cc:1216-1445 (host code; no GPU involved).
   xy: npoints (x,y) pairs scaled by 1000; xy_out: gridn*gridn (x,y) doubles in pixels, row by row from the
   board's top edge. Returns 1 if exactly one gridn x gridn grid was found, else 0 (xy_out untouched).
   The reference's neighbour graph comes from Boost.Polygon's Voronoi diagram; this library builds the same
   graph itself (exact Delaunay triangulation) - see DESIGN.md for what that does and does not pin. */
int mrg_b200_find_grid_from_points(const int* xy, int npoints, int gridn, double* xy_out);

/* The same with the reference's diagnostics (find_grid.cc:1216-1445 with debug / debug_sequence). debug != 0: the
   self-plotting /tmp/mrgingham-2-voronoi.vnl, /tmp/mrgingham-3-candidates[-detailed].vnl,
   /tmp/mrgingham-4-outer-edges[-detailed].vnl, /tmp/mrgingham-5-outer-edge-cycles,
   /tmp/mrgingham-6-identified-outer-edge-cycle (:387-480, :609-778) and the stderr messages that say where the search
   stopped. debug_sequence_x/y >= 0 (pixels): the stderr trace of every connection considered in the walks that start
   at the point nearest to that pixel (:216-310, :515-566). */
int mrg_b200_find_grid_from_points_debug(const int* xy, int npoints, int gridn, double* xy_out,
                                         int debug, int debug_sequence_x, int debug_sequence_y);

/* The neighbour graph the grid finder walks (the content of the reference's --debug Voronoi dump,
   find_grid.cc:391-430): ring[ring_off[i] .. ring_off[i+1]) = the points whose Voronoi cells share an edge with
   point i's cell, counter-clockwise in (x,y) from the +x direction. ring_off has npoints+1 entries. Returns
   the total ring length (only the first ring_cap entries are written) or <0. */
int mrg_b200_voronoi_neighbours(const int* xy, int npoints, int* ring_off, int* ring, int ring_cap);

/* mrgingham::find_chessboard_from_image_array(), mrgingham.cc:36-140. image_pyramid_level < 0: levels 3,2,1,0
   in turn until one yields a grid. refine != 0 (the reference: refinement_level != NULL): the grid's points
   are then refined level by level down to 0 (mrgingham.cc:81-99); levels_out (may be NULL) receives the level
   each point ended at. xy_out: gridn*gridn (x,y) doubles in full-resolution pixels.
   Returns the level the grid was found at, -1 if no grid was found (or on the reference's own error paths), -2 if
   the GPU path failed (CUDA error, no device) -- a failure is never reported as "not found". */
int mrg_b200_find_chessboard_from_image_array(const uint8_t* image, int Nrows, int Ncols, int stride,
                                              int gridn, int image_pyramid_level, int refine,
                                              double* xy_out, signed char* levels_out);

/* mrgingham::find_circle_grid_from_image_array(), mrgingham.cc:10-21: blobs -> grid. Returns 1 (found), 0 (not
   found) or -2 (the GPU path failed). */
int mrg_b200_find_circle_grid_from_image_array(const uint8_t* image, int Nrows, int Ncols, int stride,
                                               int gridn, double* xy_out);

/* ===========================================================================================
   C. Batched / device-resident entry points (additive)
   =========================================================================================== */

typedef struct mrg_b200_detector mrg_b200_detector;

typedef struct
{
    int device;               /* CUDA device ordinal; <0 = the calling thread's current device  */
    int max_frames;           /* frames processed per internal chunk (scratch is sized for it)  */
    int max_rows, max_cols;   /* largest frame this detector will be given                      */
    int candidate_capacity;   /* per-frame capacity of the sparse {response>15} list; rounded up
                                 to a power of two; 0 = default. Frames that overflow it are
                                 re-run on the GPU with a capacity of rows*cols.                 */
    int max_points;           /* per-frame output capacity (corners); 0 = default 1024           */
    int kernel_variant;       /* 0 = default (cascade); 1 = simple one-thread-per-pixel; 2 = tiled  */
    int blur_radius;          /* > 0: cv::blur(Size(1+2R,1+2R)) every frame on the GPU before the corner
                                 detector and the refinement, as the reference CLI does by default with
                                 R = 1 (mrgingham-from-image.cc:106-111); 0 = frames are used as given.
                                 Does not apply to the blob detector or the dense response.           */
    int clahe;                /* nonzero: cv::normalize(0,255,NORM_MINMAX) then cv::CLAHE(clipLimit 8, 8x8 tiles)
                                 on the GPU before the blur, the reference CLI's --clahe
                                 (mrgingham-from-image.cc:43-44, 71-80). Same scope as blur_radius.    */
} mrg_b200_detector_config;

/* returns 0 and a detector, or <0 */
int  mrg_b200_detector_create (mrg_b200_detector** det, const mrg_b200_detector_config* config);
void mrg_b200_detector_destroy(mrg_b200_detector* det);

/* Corner detection over a batch of equally-sized frames.
     images           first frame; frame i starts at images + i*frame_stride; rows are row_pitch apart
     images_on_device nonzero: device pointer (no copy); zero: host pointer (copied inside)
     xy_out           HOST int32 [nframes][max_points][2], scaled by 1000
     counts_out       HOST int32 [nframes]: corners found per frame (may exceed max_points; only
                      the first max_points are stored)
     stream           a cudaStream_t (as void*) to run on, or NULL for the detector's own stream
   Synchronous: results are in host memory on return. Returns 0 or <0. */
int mrg_b200_find_corners_batch(mrg_b200_detector* det,
                                const uint8_t* images, int images_on_device,
                                int nframes, int rows, int cols,
                                size_t row_pitch, size_t frame_stride,
                                int image_pyramid_level,
                                int32_t* xy_out, int32_t* counts_out,
                                void* stream);

/* The same over images of ANY sizes in one call: what the reference CLI does with a glob of images of whatever
   sizes, one per worker iteration (mrgingham-from-image.cc:50-54, 374-379). Images are grouped by size inside;
   each group goes through the batch path above (device images that lie at a fixed stride are read in place).
     images           HOST array of nimages descriptors; `data` is host or device memory per images_on_device
     xy_out           HOST int32 [nimages][max_points][2] (x, y scaled by 1000), in the order of `images`
     counts_out       HOST int32 [nimages]
   Every image must fit the detector's max_rows x max_cols. Synchronous. Returns 0 or <0. */
typedef struct mrg_b200_image_desc
{
    const uint8_t* data;
    int            rows, cols;
    size_t         row_pitch;    /* bytes between rows, >= cols */
} mrg_b200_image_desc;
int mrg_b200_find_corners_mixed_batch(mrg_b200_detector* det,
                                      const mrg_b200_image_desc* images, int nimages, int images_on_device,
                                      int image_pyramid_level,
                                      int32_t* xy_out, int32_t* counts_out, void* stream);

/* Same work, split so a caller can time the device part with its own CUDA events:
   enqueue() only enqueues (copies, kernels, result copies into the detector's pinned buffers) on
   `stream`; collect() synchronises that stream, handles overflowed frames and fills the outputs. */
int mrg_b200_find_corners_batch_enqueue(mrg_b200_detector* det,
                                        const uint8_t* images, int images_on_device,
                                        int nframes, int rows, int cols,
                                        size_t row_pitch, size_t frame_stride,
                                        int image_pyramid_level, void* stream);
int mrg_b200_find_corners_batch_collect(mrg_b200_detector* det,
                                        int32_t* xy_out, int32_t* counts_out);

/* Batched form of mrg_b200_refine_chessboard_corners(): frame i refines its own npoints points.
     xy_inout     HOST double [nframes][npoints][2], full-resolution pixels (in/out)
     levels       HOST int8   [nframes][npoints] (in/out)
     nrefined_out HOST int32  [nframes]
   Synchronous. Returns 0 or <0. */
int mrg_b200_refine_corners_batch(mrg_b200_detector* det,
                                  const uint8_t* images, int images_on_device,
                                  int nframes, int rows, int cols,
                                  size_t row_pitch, size_t frame_stride,
                                  int image_pyramid_level,
                                  double* xy_inout, signed char* levels, int npoints,
                                  int32_t* nrefined_out, void* stream);

/* Blob detection over a batch of equally-sized frames (the batched form of mrg_b200_find_blobs();
   arguments as for mrg_b200_find_corners_batch()). Synchronous. Returns 0 or <0. */
int mrg_b200_find_blobs_batch(mrg_b200_detector* det,
                              const uint8_t* images, int images_on_device,
                              int nframes, int rows, int cols,
                              size_t row_pitch, size_t frame_stride,
                              int32_t* xy_out, int32_t* counts_out,
                              void* stream);

/* Whole boards over a batch of equally-sized frames: the batched form of
   mrg_b200_find_chessboard_from_image_array() / mrg_b200_find_circle_grid_from_image_array(). Host frames are
   copied to the device once per chunk and stay there for every level and refinement pass; the corner /
   refinement kernels run over all frames that still need them, the grid finder on host threads.
     xy_out           HOST double [nframes][gridn*gridn][2], valid where found_level_out[i] >= 0
     levels_out       HOST int8 [nframes][gridn*gridn] or NULL (filled when refine != 0)
     found_level_out  HOST int32 [nframes]: level the grid was found at, or -1
   doblobs needs image_pyramid_level == 0 and ignores refine. Synchronous. Returns 0 or <0. */
int mrg_b200_find_boards_batch(mrg_b200_detector* det,
                               const uint8_t* images, int images_on_device,
                               int nframes, int rows, int cols,
                               size_t row_pitch, size_t frame_stride,
                               int gridn, int image_pyramid_level, int doblobs, int refine,
                               double* xy_out, signed char* levels_out, int32_t* found_level_out,
                               void* stream);

/* Dense ChESS response over a batch (the batched form of section A's function).
   response: int16 [nframes][rows][cols]; elements outside the 7-pixel interior are not written. */
int mrg_b200_chess_response_batch(mrg_b200_detector* det,
                                  const uint8_t* images, int images_on_device,
                                  int nframes, int rows, int cols,
                                  size_t row_pitch, size_t frame_stride,
                                  int16_t* response, int response_on_device,
                                  void* stream);

/* The sparse form of the response: per frame, every pixel of [7,w-7) x [7,h-7) whose ChESS response (at the given
   pyramid level, after the detector's configured preprocessing) exceeds 15 -- what the reference's clamp + "r > 15"
   membership test keeps (find_chessboard_corners.cc:527-529, :159-171) and all that the production ChESS kernel
   writes. Only that kernel runs. cand_out: HOST uint64 [nframes][cand_cap] or NULL, entry = y << 32 | x << 16 | r, in no
   particular order; counts_out: HOST int32 [nframes], the number found (entries beyond min(cand_cap, the detector's
   candidate_capacity) are not stored). Synchronous. Returns 0 or <0. */
int mrg_b200_chess_candidates_batch(mrg_b200_detector* det,
                                    const uint8_t* images, int images_on_device,
                                    int nframes, int rows, int cols,
                                    size_t row_pitch, size_t frame_stride,
                                    int image_pyramid_level,
                                    uint64_t* cand_out, int cand_cap, int32_t* counts_out,
                                    void* stream);

/* The box blur the reference CLI applies before the detector by default (mrgingham-from-image.cc:
   106-111, --blur R with R = 1): cv::blur(image, image, Size(1+2R, 1+2R)), BORDER_REFLECT_101.
   out: uint8 [nframes][rows][cols] dense, HOST or DEVICE (out_on_device); it must not overlap the
   input. blur_radius in [1,4]. Synchronous. Returns 0 or <0. */
int mrg_b200_box_blur_batch(mrg_b200_detector* det,
                            const uint8_t* images, int images_on_device,
                            int nframes, int rows, int cols,
                            size_t row_pitch, size_t frame_stride,
                            int blur_radius,
                            uint8_t* out, int out_on_device,
                            void* stream);

/* The whole preprocessing chain of the reference CLI (mrgingham-from-image.cc:71-111) as a stand-alone call:
   if clahe: cv::normalize(image, 0, 255, NORM_MINMAX) and cv::createCLAHE(8)->apply(); then, if
   blur_radius > 0, the blur above. Results are those of OpenCV 4.13.0. Arguments as for
   mrg_b200_box_blur_batch(); at least one of the two steps must be asked for. */
int mrg_b200_preprocess_batch(mrg_b200_detector* det,
                              const uint8_t* images, int images_on_device,
                              int nframes, int rows, int cols,
                              size_t row_pitch, size_t frame_stride,
                              int clahe, int blur_radius,
                              uint8_t* out, int out_on_device,
                              void* stream);

/* The same for 16-bit frames, as the reference CLI treats them (mrgingham-from-image.cc:83-93): with clahe,
   cv::normalize(0, 65535, NORM_MINMAX) and CLAHE(clipLimit 8) on the 16-bit data; then convertTo(CV_8U, 255./65535.);
   then the blur (blur_radius 0 = none). images: uint16 [nframes][rows][row_pitch / 2]; row_pitch and frame_stride
   in BYTES. out: uint8 [nframes][rows][cols], dense: the image the detector is then given. Returns 0 or <0. */
int mrg_b200_preprocess16_batch(mrg_b200_detector* det,
                                const uint16_t* images, int images_on_device,
                                int nframes, int rows, int cols,
                                size_t row_pitch, size_t frame_stride,
                                int clahe, int blur_radius,
                                uint8_t* out, int out_on_device, void* stream);

/* Pyramid level image (what the reference gets from cv::resize, find_chessboard_corners.cc:449-450).
   out: HOST uint8 [orows][ocols] dense; returns 0 and the size, or <0. */
int mrg_b200_pyramid_level(mrg_b200_detector* det,
                           const uint8_t* image, int rows, int cols, size_t row_pitch,
                           int image_pyramid_level,
                           uint8_t* out, int* orows, int* ocols);

/* Accumulated device time (CUDA events on the launching stream) of the kernels launched by the
   most recent batch call, and how many kernels that call launched. which: 0 = ChESS+candidate
   kernel, 1 = clustering kernel, 2 = pyramid kernel, 3 = the blob kernels (time only). */
int mrg_b200_last_kernel_ms(mrg_b200_detector* det, int which, float* ms, int* launches);
void mrg_b200_set_profiling(mrg_b200_detector* det, int enabled);

/* per-frame size of the sparse candidate list seen by the most recent batch call (HOST int32
   [nframes]); for tests and capacity planning */
int mrg_b200_last_candidate_counts(mrg_b200_detector* det, int32_t* counts_out, int nframes);

const char* mrg_b200_version(void);

/* number of CUDA devices this process can use; 0 means every other entry point will fail */
int mrg_b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
