// C++ spellings of the reference's top-level API (mrgingham.hh:11-98) as inline adapters over the C ABI in
// <mrgingham_b200.h>: same namespace, names, argument order, defaults and return conventions as
// mrgingham.cc:10-171 and find_grid.cc:1216-1225. Points are APPENDED to points_out only when a grid was
// found (the reference's find_grid_from_points() push_back()s its rows at the very end, find_grid.cc:1435-1439).
// *refinement_level is realloc()ed here and must be free()d by the caller, as in the reference
// (mrgingham.cc:36-37,82). debug / debug_sequence only produce diagnostics and /tmp dumps in the reference and
// are accepted and ignored. cv::Mat overloads exist where OpenCV's headers do; mrgingham::ImageView always.
#pragma once

#include <stdlib.h>
#include "find_blobs.hh"

namespace mrgingham
{
    struct debug_sequence_t
    {
        bool     dodebug;
        PointInt pt;
        debug_sequence_t() : dodebug(false), pt() {}
    };

    // find_grid.cc:1216-1445
    inline bool find_grid_from_points(std::vector<PointDouble>& points_out, const std::vector<PointInt>& points,
                                      const int gridn, bool debug = false,
                                      const debug_sequence_t& debug_sequence = debug_sequence_t())
    {
        static_assert(sizeof(PointInt) == 2 * sizeof(int), "PointInt must be 2 ints");
        if (points.empty() || gridn < 2) return false;
        std::vector<double> xy((size_t)2 * gridn * gridn);
        if (mrg_b200_find_grid_from_points_debug(&points[0].x, (int)points.size(), gridn, xy.data(), debug ? 1 : 0,
                                                 debug_sequence.dodebug ? debug_sequence.pt.x : -1,
                                                 debug_sequence.dodebug ? debug_sequence.pt.y : -1) != 1) return false;
        for (int i = 0; i < gridn * gridn; i++) points_out.push_back(PointDouble(xy[2*i], xy[2*i + 1]));
        return true;
    }

    // mrgingham.cc:10-21
    inline bool find_circle_grid_from_image_array(std::vector<PointDouble>& points_out, const ImageView& image,
                                                  const int gridn, bool debug = false,
                                                  debug_sequence_t debug_sequence = debug_sequence_t())
    {
        (void)debug; (void)debug_sequence;
        if (gridn < 2) return false;
        std::vector<double> xy((size_t)2 * gridn * gridn);
        const int rc = mrg_b200_find_circle_grid_from_image_array(image.data, image.rows, image.cols, (int)image.step, gridn, xy.data());
        if (rc < 0) throw std::runtime_error("mrgingham_b200: the GPU circle-grid finder failed (see stderr)");
        if (rc != 1) return false;
        for (int i = 0; i < gridn * gridn; i++) points_out.push_back(PointDouble(xy[2*i], xy[2*i + 1]));
        return true;
    }

    // mrgingham.cc:106-140. Returns the pyramid level the grid was found at, or <0.
    inline int find_chessboard_from_image_array(std::vector<PointDouble>& points_out, signed char** refinement_level,
                                                const int gridn, const ImageView& image, int image_pyramid_level = -1,
                                                bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t(),
                                                const char* debug_image_filename = NULL)
    {
        (void)debug; (void)debug_sequence; (void)debug_image_filename;
        if (gridn < 2) return -1;
        const int N = gridn * gridn;
        std::vector<double> xy((size_t)2 * N);
        std::vector<signed char> lv(N);
        const int level = mrg_b200_find_chessboard_from_image_array(image.data, image.rows, image.cols, (int)image.step, gridn,
                                                                    image_pyramid_level, refinement_level != NULL, xy.data(), lv.data());
        if (level < -1) throw std::runtime_error("mrgingham_b200: the GPU board finder failed (see stderr)");   // -2: failure, not "no board"
        if (level < 0) return -1;
        const size_t first = points_out.size();
        for (int i = 0; i < N; i++) points_out.push_back(PointDouble(xy[2*i], xy[2*i + 1]));
        if (refinement_level != NULL && level > 0)
        {
            // the reference sizes the buffer for the whole vector (mrgingham.cc:80-84)
            const size_t total = points_out.size();
            *refinement_level = (signed char*)realloc((void*)*refinement_level, total * sizeof(**refinement_level));
            if (*refinement_level == NULL) return -1;
            for (size_t i = 0; i < first; i++) (*refinement_level)[i] = (signed char)level;
            for (int i = 0; i < N; i++) (*refinement_level)[first + i] = lv[i];
        }
        return level;
    }

#ifndef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    // mrgingham.cc:23-34, :142-171; without cv::imread these read binary 8-bit PGM only
    inline bool find_circle_grid_from_image_file(std::vector<PointDouble>& points_out, const char* filename, const int gridn,
                                                 bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t())
    {
        std::vector<unsigned char> px; int rows, cols;
        if (!read_pgm_p5(filename, &px, &rows, &cols))
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        const ImageView v = { rows, cols, (size_t)cols, px.data() };
        return find_circle_grid_from_image_array(points_out, v, gridn, debug, debug_sequence);
    }
    inline int find_chessboard_from_image_file(std::vector<PointDouble>& points_out, signed char** refinement_level,
                                               const int gridn, const char* filename, int image_pyramid_level = -1,
                                               bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t())
    {
        std::vector<unsigned char> px; int rows, cols;
        if (!read_pgm_p5(filename, &px, &rows, &cols))
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return -1;
        }
        const ImageView v = { rows, cols, (size_t)cols, px.data() };
        return find_chessboard_from_image_array(points_out, refinement_level, gridn, v, image_pyramid_level, debug, debug_sequence, filename);
    }
#endif

#ifdef MRGINGHAM_B200_HAVE_OPENCV
    inline bool find_circle_grid_from_image_array(std::vector<PointDouble>& points_out, const cv::Mat& image, const int gridn,
                                                  bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t())
    {
        ImageView v;
        if (!mat_to_view(&v, image, __func__)) return false;
        return find_circle_grid_from_image_array(points_out, v, gridn, debug, debug_sequence);
    }
    inline int find_chessboard_from_image_array(std::vector<PointDouble>& points_out, signed char** refinement_level,
                                                const int gridn, const cv::Mat& image, int image_pyramid_level = -1,
                                                bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t(),
                                                const char* debug_image_filename = NULL)
    {
        ImageView v;
        if (!mat_to_view(&v, image, __func__)) return -1;
        return find_chessboard_from_image_array(points_out, refinement_level, gridn, v, image_pyramid_level, debug, debug_sequence,
                                                debug_image_filename);
    }
#ifdef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    inline bool find_circle_grid_from_image_file(std::vector<PointDouble>& points_out, const char* filename, const int gridn,
                                                 bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t())
    {
        cv::Mat image = cv::imread(filename, cv::IMREAD_IGNORE_ORIENTATION | cv::IMREAD_GRAYSCALE);
        if (image.data == NULL)
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        return find_circle_grid_from_image_array(points_out, image, gridn, debug, debug_sequence);
    }
    inline int find_chessboard_from_image_file(std::vector<PointDouble>& points_out, signed char** refinement_level,
                                               const int gridn, const char* filename, int image_pyramid_level = -1,
                                               bool debug = false, debug_sequence_t debug_sequence = debug_sequence_t())
    {
        cv::Mat image = cv::imread(filename, cv::IMREAD_IGNORE_ORIENTATION | cv::IMREAD_GRAYSCALE);
        if (image.data == NULL)
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return -1;
        }
        return find_chessboard_from_image_array(points_out, refinement_level, gridn, image, image_pyramid_level, debug, debug_sequence,
                                                filename);
    }
#endif
#endif
}
