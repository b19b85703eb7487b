// C++ spelling of the reference's blob API (find_blobs.hh:15-20) as an inline adapter over
// mrg_b200_find_blobs() in <mrgingham_b200.h>: same namespace, name, argument order, default and
// behaviour as find_blobs.cc:14-46 -- points are APPENDED scaled by 1000; with dodump the
// float32 keypoint coordinates are printed instead ("%f %f\n", find_blobs.cc:34-36; the values
// printed here are the scaled integers divided by 1000, i.e. rounded to 1e-3 px); returns true.
// The cv::Mat overload is compiled where OpenCV's C++ headers exist; mrgingham::ImageView works
// everywhere (see find_chessboard_corners.hh). The file variant (find_blobs.cc:48-64) needs
// cv::imread and is therefore only provided where OpenCV's highgui/imgcodecs headers exist.
#pragma once

#include "find_chessboard_corners.hh"

#if defined(MRGINGHAM_B200_HAVE_OPENCV) && defined(__has_include)
#  if __has_include(<opencv2/highgui/highgui.hpp>)
#    include <opencv2/highgui/highgui.hpp>
#    define MRGINGHAM_B200_HAVE_OPENCV_IMREAD 1
#  endif
#endif

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
