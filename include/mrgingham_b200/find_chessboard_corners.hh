// C++ spellings of the reference's API for the hot path, as inline adapters over the C ABI in
// <mrgingham_b200.h>. Mirrors find_chessboard_corners.hh:12-72 of the reference: same namespace,
// names, argument order, defaults, return conventions and "append to the vector" semantics
// (find_chessboard_corners.cc:350 push_back()s; the vector is never cleared).
//
// The reference takes cv::Mat. Where OpenCV's C++ headers exist the cv::Mat overloads below are
// compiled; everywhere (this image has no OpenCV C++ headers) the same functions are available on
// mrgingham::ImageView, a 4-field view with the only members the reference reads from the Mat
// (rows, cols, step, data; type() and isContinuous() are expressed by construction).
#pragma once

#include <vector>
#include <stdexcept>
#include <stdio.h>
#include "../mrgingham_b200.h"

#if defined(__has_include)
#  if __has_include(<opencv2/core/core.hpp>)
#    include <opencv2/core/core.hpp>
#    define MRGINGHAM_B200_HAVE_OPENCV 1
#    if __has_include(<opencv2/highgui/highgui.hpp>)
#      include <opencv2/highgui/highgui.hpp>
#      define MRGINGHAM_B200_HAVE_OPENCV_IMREAD 1
#    endif
#  endif
#endif

namespace mrgingham
{
#ifndef MRGINGHAM_B200_POINT_TYPES
#define MRGINGHAM_B200_POINT_TYPES
    // point.hh:5-15 of the reference (layout relied on by the bridge: 2 ints / 2 doubles)
    struct PointInt    { int x, y;    PointInt(int _x = 0, int _y = 0) : x(_x), y(_y) {} };
    struct PointDouble { double x, y; PointDouble(double _x = 0, double _y = 0) : x(_x), y(_y) {} };
#endif

    struct ImageView
    {
        int rows, cols;
        size_t step;                 // bytes between rows
        const unsigned char* data;   // 8-bit, single channel
    };

    // find_chessboard_corners.cc:568-587. Appends to *points_scaled_out (scaled by 1000);
    // returns true iff at least one point is in the vector afterwards (the reference returns
    // points_scaled_out->size() > 0).
    inline bool find_chessboard_corners_from_image_array(std::vector<PointInt>* points_scaled_out,
                                                         const ImageView& image_input,
                                                         int image_pyramid_level,
                                                         bool debug = false,
                                                         const char* debug_image_filename = NULL)
    {
        int cap = 4096;
        std::vector<int> xy((size_t)2 * cap);
        int n = mrg_b200_find_chessboard_corners(image_input.data, image_input.rows, image_input.cols, (int)image_input.step,
                                                 image_pyramid_level, xy.data(), cap);
        if (n > cap)
        {
            cap = n; xy.resize((size_t)2 * cap);
            n = mrg_b200_find_chessboard_corners(image_input.data, image_input.rows, image_input.cols, (int)image_input.step,
                                                 image_pyramid_level, xy.data(), cap);
        }
        // n < 0: the GPU path failed (CUDA error, no device). There is no CPU fallback and "no points" would be a lie:
        // the failure is thrown (the reference's bool cannot carry it).
        if (n < 0) throw std::runtime_error("mrgingham_b200: the GPU corner detector failed (see stderr)");
        for (int i = 0; i < n; i++) points_scaled_out->push_back(PointInt(xy[2*i], xy[2*i + 1]));
        if (debug)      // the reference's /tmp dumps (find_chessboard_corners.cc:294-315, 452-459, 513-541)
            mrg_b200_debug_dump_corners(image_input.data, image_input.rows, image_input.cols, (int)image_input.step,
                                        image_pyramid_level, debug_image_filename, NULL, NULL, 0);
        return points_scaled_out->size() > 0;
    }

    // find_chessboard_corners.cc:591-619. Returns how many points were refined.
    inline int refine_chessboard_corners_from_image_array(std::vector<PointDouble>* points,
                                                          signed char* level,
                                                          const ImageView& image_input,
                                                          int image_pyramid_level,
                                                          bool debug = false,
                                                          const char* debug_image_filename = NULL)
    {
        static_assert(sizeof(PointDouble) == 2 * sizeof(double), "PointDouble must be 2 doubles");
        if (points->empty()) return 0;
        const int n = mrg_b200_refine_chessboard_corners(image_input.data, image_input.rows, image_input.cols, (int)image_input.step,
                                                         image_pyramid_level, &(*points)[0].x, level, (int)points->size());
        // callers stop refining on "<= 0 points refined" (mrgingham.cc:96): a GPU failure must not pass for that
        if (n < 0) throw std::runtime_error("mrgingham_b200: the GPU corner refinement failed (see stderr)");
        if (debug)      // the reference's /tmp dumps of a refinement pass (find_chessboard_corners.cc:300-305, 391-392, 513-541)
            mrg_b200_debug_dump_corners(image_input.data, image_input.rows, image_input.cols, (int)image_input.step,
                                        image_pyramid_level, debug_image_filename, &(*points)[0].x, level, (int)points->size());
        return n;
    }

    // Image files (find_chessboard_corners.cc:622-648, find_blobs.cc:48-64 go through cv::imread with
    // IMREAD_IGNORE_ORIENTATION | IMREAD_GRAYSCALE). Where OpenCV's imread is not available the file
    // variants below read binary 8-bit PGM ("P5") only; anything else fails like an unreadable file does
    // in the reference: the diagnostic "Couldn't open image" and false.
    inline bool read_pgm_p5(const char* filename, std::vector<unsigned char>* pixels, int* rows, int* cols)
    {
        FILE* fp = fopen(filename, "rb");
        if (!fp) return false;
        auto token = [&](int* v) -> bool
        {
            int c = fgetc(fp);
            for (;;)
            {
                while (c == ' ' || c == '\t' || c == '\n' || c == '\r') c = fgetc(fp);
                if (c != '#') break;
                while (c != '\n' && c != EOF) c = fgetc(fp);
            }
            if (c < '0' || c > '9') return false;
            long n = 0;
            while (c >= '0' && c <= '9') { n = n * 10 + (c - '0'); if (n > 100000) return false; c = fgetc(fp); }
            *v = (int)n;           // the single whitespace after the last header token has just been consumed
            return true;
        };
        int w = 0, h = 0, maxval = 0;
        bool ok = fgetc(fp) == 'P' && fgetc(fp) == '5' && token(&w) && token(&h) && token(&maxval) &&
                  w > 0 && h > 0 && maxval > 0 && maxval <= 255;
        if (ok)
        {
            pixels->resize((size_t)w * h);
            ok = fread(pixels->data(), 1, pixels->size(), fp) == pixels->size();
        }
        fclose(fp);
        *rows = h; *cols = w;
        return ok;
    }

#ifndef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    inline bool find_chessboard_corners_from_image_file(std::vector<PointInt>* points, const char* filename,
                                                        int image_pyramid_level, bool debug = false)
    {
        std::vector<unsigned char> px; int rows, cols;
        if (!read_pgm_p5(filename, &px, &rows, &cols))
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        const ImageView v = { rows, cols, (size_t)cols, px.data() };
        return find_chessboard_corners_from_image_array(points, v, image_pyramid_level, debug, filename);
    }
#endif

#ifdef MRGINGHAM_B200_HAVE_OPENCV
    inline bool mat_to_view(ImageView* v, const cv::Mat& m, const char* who)
    {
        if (m.type() != CV_8U)
        {
            // same diagnostic as find_chessboard_corners.cc:468-473
            fprintf(stderr, "%s:%d in %s(): I can only handle CV_8U arrays currently. Sorry.\n", __FILE__, __LINE__, who);
            return false;
        }
        v->rows = m.rows; v->cols = m.cols; v->step = m.step; v->data = m.data;
        return true;
    }
    inline bool find_chessboard_corners_from_image_array(std::vector<PointInt>* points_scaled_out, const cv::Mat& image_input,
                                                         int image_pyramid_level, bool debug = false,
                                                         const char* debug_image_filename = NULL)
    {
        ImageView v;
        if (!mat_to_view(&v, image_input, __func__)) return points_scaled_out->size() > 0;
        return find_chessboard_corners_from_image_array(points_scaled_out, v, image_pyramid_level, debug, debug_image_filename);
    }
#ifdef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    inline bool find_chessboard_corners_from_image_file(std::vector<PointInt>* points, const char* filename,
                                                        int image_pyramid_level, bool debug = false)
    {
        cv::Mat image = cv::imread(filename, cv::IMREAD_IGNORE_ORIENTATION | cv::IMREAD_GRAYSCALE);
        if (image.data == NULL)
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        return find_chessboard_corners_from_image_array(points, image, image_pyramid_level, debug, filename);
    }
#endif
    inline int refine_chessboard_corners_from_image_array(std::vector<PointDouble>* points, signed char* level,
                                                          const cv::Mat& image_input, int image_pyramid_level,
                                                          bool debug = false, const char* debug_image_filename = NULL)
    {
        ImageView v;
        if (!mat_to_view(&v, image_input, __func__)) return 0;
        return refine_chessboard_corners_from_image_array(points, level, v, image_pyramid_level, debug, debug_image_filename);
    }
#endif
}
