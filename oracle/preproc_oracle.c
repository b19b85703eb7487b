/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the reference CLI's optional preprocessing (SURVEY.md
 * section 8f, row F2): mrgingham-from-image.cc:71-80 with --clahe does
 *     cv::normalize(image, image, 0, 255, NORM_MINMAX);  clahe->apply(image, image1);   // clipLimit 8 (:43-44)
 * before the blur (:106-111, restated in blob_oracle.c). Both are OpenCV (third party, version not
 * pinned by the reference); restated here from OpenCV's published algorithm and PINNED against the
 * in-container cv2 4.13.0 by tests/test_preproc.py (random images, divisible and non-divisible sizes).
 * One machine dependence is inherited from OpenCV: its 8u->8u convertTo multiplies and adds with one
 * rounding (fused multiply-add) on CPUs that have FMA3, which is what cv2 does here and on the GPU box.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define EXPORT __attribute__((visibility("default")))

/* cv::normalize(src, dst, 0, 255, NORM_MINMAX) for CV_8U: scale = 255 * (1 / (max - min)) (0 if max == min),
 * shift = -min * scale in double; then convertTo with float alpha/beta: rint(fmaf(v, alpha, beta)), saturated */
EXPORT void preproc_oracle_normalize_lut(const uint8_t* in, int w, int h, int stride, uint8_t lut[256])
{
    int mn = 255, mx = 0;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { const int v = in[(size_t)y * stride + x]; if (v < mn) mn = v; if (v > mx) mx = v; }
    const double smin = mn, smax = mx;
    const double scale = 255.0 * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, b = (float)shift;
    for (int v = 0; v < 256; v++)
    {
        const float r = fmaf((float)v, a, b);
        long q = lrintf(r);
        lut[v] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
}

static int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}

/* cv::CLAHE::apply for CV_8U with clipLimit and an 8x8 tile grid (OpenCV imgproc/clahe.cpp) */
EXPORT void preproc_oracle_clahe(const uint8_t* in, int w, int h, int stride, double clip_limit, uint8_t* out /* dense [h][w] */)
{
    const int tilesX = 8, tilesY = 8, histSize = 256;
    int we = w, he = h;
    if (w % tilesX || h % tilesY) { we = w + tilesX - (w % tilesX); he = h + tilesY - (h % tilesY); }   /* BORDER_REFLECT_101 padding */
    const int tw = we / tilesX, th = he / tilesY, total = tw * th;
    const float lutScale = (float)(histSize - 1) / total;
    int clipLimit = 0;
    if (clip_limit > 0.0) { clipLimit = (int)(clip_limit * total / histSize); if (clipLimit < 1) clipLimit = 1; }
    uint8_t* lut = (uint8_t*)malloc((size_t)tilesX * tilesY * histSize);
    for (int ty = 0; ty < tilesY; ty++)
        for (int tx = 0; tx < tilesX; tx++)
        {
            int hist[256]; memset(hist, 0, sizeof(hist));
            for (int y = ty * th; y < (ty + 1) * th; y++)
                for (int x = tx * tw; x < (tx + 1) * tw; x++)
                    hist[in[(size_t)reflect101(y, h) * stride + reflect101(x, w)]]++;
            if (clipLimit > 0)
            {
                int clipped = 0;
                for (int i = 0; i < histSize; i++) if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
                const int redistBatch = clipped / histSize;
                int residual = clipped - redistBatch * histSize;
                for (int i = 0; i < histSize; i++) hist[i] += redistBatch;
                if (residual != 0)
                {
                    int residualStep = histSize / residual; if (residualStep < 1) residualStep = 1;
                    for (int i = 0; i < histSize && residual > 0; i += residualStep, residual--) hist[i]++;
                }
            }
            int sum = 0;
            uint8_t* tl = lut + (size_t)(ty * tilesX + tx) * histSize;
            for (int i = 0; i < histSize; i++)
            {
                sum += hist[i];
                long q = lrintf((float)sum * lutScale);
                tl[i] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
            }
        }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; y++)
    {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > tilesY - 1) ty2 = tilesY - 1;
        for (int x = 0; x < w; x++)
        {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > tilesX - 1) tx2 = tilesX - 1;
            const int v = in[(size_t)y * stride + x];
            const float p1 = lut[(size_t)(ty1 * tilesX + tx1) * histSize + v], p2 = lut[(size_t)(ty1 * tilesX + tx2) * histSize + v];
            const float q1 = lut[(size_t)(ty2 * tilesX + tx1) * histSize + v], q2 = lut[(size_t)(ty2 * tilesX + tx2) * histSize + v];
            const float res = (p1 * xa1 + p2 * xa) * ya1 + (q1 * xa1 + q2 * xa) * ya;
            long q = lrintf(res);
            out[(size_t)y * w + x] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
        }
    }
    free(lut);
}

/* the CLI's --clahe chain: normalize, then CLAHE with clip limit 8 (mrgingham-from-image.cc:43-44, :77-78) */
EXPORT void preproc_oracle_normalize_clahe(const uint8_t* in, int w, int h, int stride, uint8_t* out)
{
    uint8_t lut[256];
    preproc_oracle_normalize_lut(in, w, h, stride, lut);
    uint8_t* tmp = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) tmp[(size_t)y * w + x] = lut[in[(size_t)y * stride + x]];
    preproc_oracle_clahe(tmp, w, h, w, 8.0, out);
    free(tmp);
}

/* ---- 16-bit input (mrgingham-from-image.cc:83-93): normalize to [0,65535] + CLAHE on 16 bits (with --clahe),
 * then image0.convertTo(image1, CV_8U, 255./65535.). Same OpenCV arithmetic, PINNED against cv2 4.13.0 by
 * tests/test_preproc16.py. ---- */

/* cv::Mat::convertTo(CV_8U, 255./65535.) of a CV_16U image: rint((float)v * (float)alpha), saturated */
EXPORT void preproc_oracle_convert16to8(const uint16_t* in, int w, int h, int stride_elems, uint8_t* out /* dense */)
{
    const float a = (float)(255. / 65535.);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            long q = lrintf(fmaf((float)in[(size_t)y * stride_elems + x], a, 0.0f));
            out[(size_t)y * w + x] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
        }
}

/* cv::normalize(src, dst, 0, 65535, NORM_MINMAX) for CV_16U */
EXPORT void preproc_oracle_normalize16(const uint16_t* in, int w, int h, int stride_elems, uint16_t* out /* dense */)
{
    int mn = 65535, mx = 0;
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) { const int v = in[(size_t)y * stride_elems + x]; if (v < mn) mn = v; if (v > mx) mx = v; }
    const double smin = mn, smax = mx;
    const double scale = 65535.0 * (smax - smin > DBL_EPSILON ? 1. / (smax - smin) : 0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, b = (float)shift;
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            long q = lrintf(fmaf((float)in[(size_t)y * stride_elems + x], a, b));
            out[(size_t)y * w + x] = (uint16_t)(q < 0 ? 0 : q > 65535 ? 65535 : q);
        }
}

/* cv::CLAHE::apply for CV_16U: the 8-bit algorithm with 65536 bins */
EXPORT void preproc_oracle_clahe16(const uint16_t* in, int w, int h, int stride_elems, double clip_limit, uint16_t* out /* dense */)
{
    const int tilesX = 8, tilesY = 8, histSize = 65536;
    int we = w, he = h;
    if (w % tilesX || h % tilesY) { we = w + tilesX - (w % tilesX); he = h + tilesY - (h % tilesY); }
    const int tw = we / tilesX, th = he / tilesY, total = tw * th;
    const float lutScale = (float)(histSize - 1) / total;
    int clipLimit = 0;
    if (clip_limit > 0.0) { clipLimit = (int)(clip_limit * total / histSize); if (clipLimit < 1) clipLimit = 1; }
    uint16_t* lut = (uint16_t*)malloc((size_t)tilesX * tilesY * histSize * sizeof(uint16_t));
    int* hist = (int*)malloc(sizeof(int) * histSize);
    for (int ty = 0; ty < tilesY; ty++)
        for (int tx = 0; tx < tilesX; tx++)
        {
            memset(hist, 0, sizeof(int) * histSize);
            for (int y = ty * th; y < (ty + 1) * th; y++)
                for (int x = tx * tw; x < (tx + 1) * tw; x++)
                    hist[in[(size_t)reflect101(y, h) * stride_elems + reflect101(x, w)]]++;
            if (clipLimit > 0)
            {
                int clipped = 0;
                for (int i = 0; i < histSize; i++) if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
                const int redistBatch = clipped / histSize;
                int residual = clipped - redistBatch * histSize;
                for (int i = 0; i < histSize; i++) hist[i] += redistBatch;
                if (residual != 0)
                {
                    int residualStep = histSize / residual; if (residualStep < 1) residualStep = 1;
                    for (int i = 0; i < histSize && residual > 0; i += residualStep, residual--) hist[i]++;
                }
            }
            int sum = 0;
            uint16_t* tl = lut + (size_t)(ty * tilesX + tx) * histSize;
            for (int i = 0; i < histSize; i++)
            {
                sum += hist[i];
                long q = lrintf((float)sum * lutScale);
                tl[i] = (uint16_t)(q < 0 ? 0 : q > 65535 ? 65535 : q);
            }
        }
    const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
    for (int y = 0; y < h; y++)
    {
        const float tyf = y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        const float ya = tyf - ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > tilesY - 1) ty2 = tilesY - 1;
        for (int x = 0; x < w; x++)
        {
            const float txf = x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            const float xa = txf - tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > tilesX - 1) tx2 = tilesX - 1;
            const int v = in[(size_t)y * stride_elems + x];
            const float p1 = lut[(size_t)(ty1 * tilesX + tx1) * histSize + v], p2 = lut[(size_t)(ty1 * tilesX + tx2) * histSize + v];
            const float q1 = lut[(size_t)(ty2 * tilesX + tx1) * histSize + v], q2 = lut[(size_t)(ty2 * tilesX + tx2) * histSize + v];
            const float res = (p1 * xa1 + p2 * xa) * ya1 + (q1 * xa1 + q2 * xa) * ya;
            long q = lrintf(res);
            out[(size_t)y * w + x] = (uint16_t)(q < 0 ? 0 : q > 65535 ? 65535 : q);
        }
    }
    free(hist); free(lut);
}

/* the CLI's chain for a 16-bit image: [normalize + CLAHE(8) if clahe] then convertTo 8 bits */
EXPORT void preproc_oracle_chain16(const uint16_t* in, int w, int h, int stride_elems, int clahe, uint8_t* out /* dense */)
{
    if (!clahe) { preproc_oracle_convert16to8(in, w, h, stride_elems, out); return; }
    uint16_t* a = (uint16_t*)malloc((size_t)w * h * 2), *b = (uint16_t*)malloc((size_t)w * h * 2);
    preproc_oracle_normalize16(in, w, h, stride_elems, a);
    preproc_oracle_clahe16(a, w, h, w, 8.0, b);
    preproc_oracle_convert16to8(b, w, h, w, out);
    free(a); free(b);
}
