// TEST INFRASTRUCTURE ONLY -- never shipped, never linked into the product library.
//
// Minimal stand-in for the handful of OpenCV C++ names that the reference's
// find_chessboard_corners.cc touches, so that file can be compiled UNMODIFIED from
// /root/reference into oracle/_ref/ (OpenCV's C++ headers are not installed in this image;
// only the cv2 Python module is). Everything here is plumbing except cv::resize(), which is
// the one piece of third-party arithmetic on the path: it implements the exact integer model
// of cv::resize(..., INTER_LINEAR) for power-of-two down-scaling that tests/test_pyramid_model.py
// pins bit-for-bit against cv2 4.13.0 (the in-container OpenCV).
#pragma once

#include <assert.h>     // the real OpenCV headers pull it in; mrgingham.cc:83 relies on that
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define CV_8U   0
#define CV_16S  3
#define CV_8UC1 0

namespace cv
{
    struct Size
    {
        int width, height;
        Size(int w=0, int h=0) : width(w), height(h) {}
    };

    // Reference-counting is irrelevant for the way the reference uses Mat (locals only), so an
    // owning buffer + a non-owning view is all that is needed.
    class Mat
    {
        std::vector<uint8_t> _own;
        int _type;
    public:
        int      rows, cols;
        uint8_t* data;
        size_t   step;

        Mat() : _type(CV_8U), rows(0), cols(0), data(NULL), step(0) {}
        // non-owning view over caller memory: cv::Mat(rows, cols, type, data, step)
        Mat(int r, int c, int type, void* d, size_t s) :
            _type(type), rows(r), cols(c), data((uint8_t*)d), step(s) {}
        Mat(const Mat& m) { *this = m; }
        Mat& operator=(const Mat& m)
        {
            _type = m._type; rows = m.rows; cols = m.cols; step = m.step;
            if(!m._own.empty()) { _own = m._own; data = _own.data(); }
            else                { _own.clear();  data = m.data; }
            return *this;
        }

        void create(int r, int c, int type)
        {
            _type = type; rows = r; cols = c;
            const size_t elsize = (type == CV_16S) ? 2 : 1;
            step = (size_t)c * elsize;
            _own.assign((size_t)r * step, 0);
            data = _own.data();
        }
        static Mat zeros(Size s, int type)
        {
            Mat m;
            m.create(s.height, s.width, type);
            return m;
        }
        int  type()         const { return _type; }
        bool isContinuous() const
        {
            const size_t elsize = (_type == CV_16S) ? 2 : 1;
            return rows <= 1 || step == (size_t)cols * elsize;
        }
    };

    enum { INTER_LINEAR = 1 };
    enum { NORM_MINMAX  = 32 };
    enum { IMREAD_GRAYSCALE = 0, IMREAD_IGNORE_ORIENTATION = 128 };

    // The debug-only calls: no-ops here. The oracle never runs the reference with debug=true.
    inline bool imwrite(const std::string&, const Mat&) { return true; }
    inline Mat  imread (const std::string&, int)        { return Mat(); }
    inline void normalize(const Mat&, Mat&, double, double, int) {}

    static inline int _cvRound(double v)
    {
        // round-half-to-even, as cvRound() (lrint in the default rounding mode)
        return (int)__builtin_nearbyint(v);
    }
    // rint((a+b)/2.0) with ties to even, on integers
    static inline int _rint_half(int s) { int q = s >> 1; return q + ((s & 1) & (q & 1)); }

    // cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR) for fx == fy == 1/2^L, 8-bit single
    // channel. Model (validated against cv2 4.13.0 over 3420 (size,level) cases):
    //   B = 2^L; ow = cvRound(W/B); oh = cvRound(H/B)
    //   out[dy][dx] = (I[y0][x0] + I[y0][x1] + I[y1][x0] + I[y1][x1] + 2) >> 2
    //   x0 = min(B*dx + B/2 - 1, W-1), x1 = min(x0+1, W-1), same in y
    // except for L == 1 where OpenCV takes its 2x2 INTER_AREA fast path and a trailing PARTIAL
    // cell (W or H == 3 mod 4) is the round-half-even mean of the pixels that exist.
    inline void resize(const Mat& src, Mat& dst, Size, double fx, double fy, int /*interp*/)
    {
        int L = 0;
        while((1.0 / (double)(1 << L)) > fx && L < 30) L++;
        (void)fy;
        const int B = 1 << L, W = src.cols, H = src.rows;
        const int ow = _cvRound((double)W / B), oh = _cvRound((double)H / B);
        dst.create(oh, ow, CV_8U);
        for(int dy = 0; dy < oh; dy++)
        {
            int y0 = B*dy + B/2 - 1; if(y0 > H-1) y0 = H-1;
            int y1 = y0 + 1;         if(y1 > H-1) y1 = H-1;
            const uint8_t* r0 = src.data + (size_t)y0 * src.step;
            const uint8_t* r1 = src.data + (size_t)y1 * src.step;
            uint8_t* o = dst.data + (size_t)dy * dst.step;
            for(int dx = 0; dx < ow; dx++)
            {
                int x0 = B*dx + B/2 - 1; if(x0 > W-1) x0 = W-1;
                int x1 = x0 + 1;         if(x1 > W-1) x1 = W-1;
                o[dx] = (uint8_t)((r0[x0] + r0[x1] + r1[x0] + r1[x1] + 2) >> 2);
            }
        }
        if(L == 1)
        {
            const bool px = 2*ow > W, py = 2*oh > H;
            if(px)
                for(int dy = 0; dy < oh; dy++)
                {
                    int ya = 2*dy, yb = ya+1 > H-1 ? H-1 : ya+1;
                    dst.data[(size_t)dy*dst.step + ow-1] =
                        (uint8_t)_rint_half(src.data[(size_t)ya*src.step + W-1] +
                                            src.data[(size_t)yb*src.step + W-1]);
                }
            if(py)
                for(int dx = 0; dx < ow; dx++)
                {
                    int xa = 2*dx, xb = xa+1 > W-1 ? W-1 : xa+1;
                    dst.data[(size_t)(oh-1)*dst.step + dx] =
                        (uint8_t)_rint_half(src.data[(size_t)(H-1)*src.step + xa] +
                                            src.data[(size_t)(H-1)*src.step + xb]);
                }
            if(px && py)
                dst.data[(size_t)(oh-1)*dst.step + ow-1] = src.data[(size_t)(H-1)*src.step + W-1];
        }
    }
}
