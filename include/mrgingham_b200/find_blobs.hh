// C++ spelling of the reference's blob API (find_blobs.hh:15-20) as an inline adapter over
// mrg_b200_find_blobs() in <mrgingham_b200.h>: same namespace, name, argument order, default and
// behaviour as find_blobs.cc:14-46 -- points are APPENDED scaled by 1000; with dodump the
// float32 keypoint coordinates are printed instead ("%f %f\n", find_blobs.cc:34-36; the values
// printed here are the scaled integers divided by 1000, i.e. rounded to 1e-3 px); returns true.
// The cv::Mat overload is compiled where OpenCV's C++ headers exist; mrgingham::ImageView works
// everywhere (see find_chessboard_corners.hh). The file variant (find_blobs.cc:48-64) uses cv::imread
// where OpenCV's highgui/imgcodecs headers exist and reads binary 8-bit PGM otherwise.
#pragma once

#include "find_chessboard_corners.hh"

namespace mrgingham
{
    inline bool find_blobs_from_image_array(std::vector<PointInt>* points, const ImageView& image, bool dodump = false)
    {
        int cap = 4096;
        std::vector<int> xy((size_t)2 * cap);
        int n = mrg_b200_find_blobs(image.data, image.rows, image.cols, (int)image.step, xy.data(), cap);
        if (n > cap)
        {
            cap = n; xy.resize((size_t)2 * cap);
            n = mrg_b200_find_blobs(image.data, image.rows, image.cols, (int)image.step, xy.data(), cap);
        }
        if (n < 0) throw std::runtime_error("mrgingham_b200: the GPU blob detector failed (see stderr)");
        for (int i = 0; i < n; i++)
        {
            if (dodump) printf("%f %f\n", xy[2*i] / 1000.0, xy[2*i + 1] / 1000.0);
            else        points->push_back(PointInt(xy[2*i], xy[2*i + 1]));
        }
        return true;
    }

#ifdef MRGINGHAM_B200_HAVE_OPENCV
    inline bool find_blobs_from_image_array(std::vector<PointInt>* points, const cv::Mat& image, bool dodump = false)
    {
        ImageView v;
        if (!mat_to_view(&v, image, __func__)) return true;
        return find_blobs_from_image_array(points, v, dodump);
    }
#endif
#ifndef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    // without cv::imread: binary 8-bit PGM only (see read_pgm_p5 in find_chessboard_corners.hh)
    inline bool find_blobs_from_image_file(std::vector<PointInt>* points, const char* filename, bool dodump = false)
    {
        std::vector<unsigned char> px; int rows, cols;
        if (!read_pgm_p5(filename, &px, &rows, &cols))
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        const ImageView v = { rows, cols, (size_t)cols, px.data() };
        return find_blobs_from_image_array(points, v, dodump);
    }
#endif
#ifdef MRGINGHAM_B200_HAVE_OPENCV_IMREAD
    inline bool find_blobs_from_image_file(std::vector<PointInt>* points, const char* filename, bool dodump = false)
    {
        cv::Mat image = cv::imread(filename, cv::IMREAD_IGNORE_ORIENTATION | cv::IMREAD_GRAYSCALE);
        if (image.data == NULL)
        {
            fprintf(stderr, "%s:%d in %s(): Couldn't open image '%s'. Sorry.\n", __FILE__, __LINE__, __func__, filename);
            return false;
        }
        return find_blobs_from_image_array(points, image, dodump);
    }
#endif
}
