// TEST INFRASTRUCTURE ONLY -- never shipped, never linked into the product library.
//
// extern "C" handles onto the UNMODIFIED reference hot path, so Python tests / the bench's
// cpu_baseline leg can call it through ctypes. Compiled by oracle/Makefile together with
// /root/reference/ChESS.c and /root/reference/find_chessboard_corners.cc (sources are compiled
// where they lie; nothing is copied) against the cv::Mat shim in oracle/shim/.
//
// Reference entry points used (file:line in /root/reference):
//   mrgingham_ChESS_response_5                         ChESS.c:55
//   mrgingham::find_chessboard_corners_from_image_array   find_chessboard_corners.cc:568
//   mrgingham::refine_chessboard_corners_from_image_array find_chessboard_corners.cc:591
#include <vector>
#include <stdint.h>
#include <string.h>

#include "find_chessboard_corners.hh"

extern "C" {
#include "ChESS.h"
}

#define API extern "C" __attribute__((visibility("default")))

API void ref_ChESS_response_5(int16_t* response, const uint8_t* image, int w, int h, int stride)
{
    mrgingham_ChESS_response_5(response, image, w, h, stride);
}

// returns the number of points the reference found (its vector's size); writes min(N,cap) of them
API int ref_find_chessboard_corners(const uint8_t* image, int rows, int cols, int stride,
                                    int level, int* xy_out, int cap)
{
    cv::Mat m(rows, cols, CV_8UC1, (void*)image, (size_t)stride);
    std::vector<mrgingham::PointInt> pts;
    mrgingham::find_chessboard_corners_from_image_array(&pts, m, level);
    const int N = (int)pts.size();
    for(int i = 0; i < N && i < cap; i++) { xy_out[2*i] = pts[i].x; xy_out[2*i+1] = pts[i].y; }
    return N;
}

// in/out points (full-resolution pixel coordinates, doubles) and levels; returns Nrefined
API int ref_refine_chessboard_corners(const uint8_t* image, int rows, int cols, int stride,
                                      int level, double* xy_inout, signed char* levels, int npoints)
{
    cv::Mat m(rows, cols, CV_8UC1, (void*)image, (size_t)stride);
    std::vector<mrgingham::PointDouble> pts(npoints);
    for(int i = 0; i < npoints; i++) { pts[i].x = xy_inout[2*i]; pts[i].y = xy_inout[2*i+1]; }
    const int N = mrgingham::refine_chessboard_corners_from_image_array(&pts, levels, m, level);
    for(int i = 0; i < npoints; i++) { xy_inout[2*i] = pts[i].x; xy_inout[2*i+1] = pts[i].y; }
    return N;
}

// The shim's cv::resize model, exposed so tests can pin it against the real cv2.resize
API int ref_shim_resize(const uint8_t* image, int rows, int cols, int stride, int level,
                        uint8_t* out, int* orows, int* ocols)
{
    cv::Mat m(rows, cols, CV_8UC1, (void*)image, (size_t)stride), d;
    cv::resize(m, d, cv::Size(), 1.0 / (double)(1 << level), 1.0 / (double)(1 << level), cv::INTER_LINEAR);
    *orows = d.rows; *ocols = d.cols;
    if(out) for(int y = 0; y < d.rows; y++) memcpy(out + (size_t)y*d.cols, d.data + (size_t)y*d.step, d.cols);
    return 0;
}
