/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the blob path.
 *
 * The reference's blob detector is find_blobs_from_image_array() (find_blobs.cc:14-46): a
 * cv::SimpleBlobDetector with minArea=20, maxArea=80000, minDistBetweenBlobs=5, blobColor=0
 * (find_blobs.cc:18-22), everything else OpenCV's defaults, followed by
 * PointInt((int)(pt.x*1000 + 0.5), ...) on the float32 keypoints (find_blobs.cc:40-41).
 *
 * The arithmetic lives in OpenCV, a third-party dependency that is NOT under /root/reference and
 * whose version the reference does not pin (Makefile:29-30 takes whatever pkg-config yields). This
 * file restates the published algorithm of OpenCV's SimpleBlobDetector::detect / findContours
 * (Suzuki-Abe border following, RETR_LIST + CHAIN_APPROX_NONE) / contour moments / convex hull in
 * plain C. PARITY PINNED against the in-container stand-in cv2 4.13.0: tests/test_blob_oracle.py
 * compares contours point-for-point with cv2.findContours on random binaries and the final point
 * lists with the fixtures tests/golden/blobs_v1.npz that tests/golden/make_blob_golden.py generated
 * from cv2.SimpleBlobDetector with the reference's parameters.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

#define EXPORT __attribute__((visibility("default")))

typedef struct { int x, y; } pt_t;

typedef struct
{
    pt_t* pts;   int npts, cappts;
    int*  start; int ncont, capcont;      /* start[i] .. start[i+1] */
} contours_t;

static void push_pt(contours_t* c, int x, int y)
{
    if (c->npts == c->cappts) { c->cappts = c->cappts ? 2*c->cappts : 4096; c->pts = (pt_t*)realloc(c->pts, sizeof(pt_t)*c->cappts); }
    c->pts[c->npts].x = x; c->pts[c->npts].y = y; c->npts++;
}
static void begin_contour(contours_t* c)
{
    if (c->ncont + 2 > c->capcont) { c->capcont = c->capcont ? 2*c->capcont : 1024; c->start = (int*)realloc(c->start, sizeof(int)*c->capcont); }
    c->start[c->ncont] = c->npts;
}
static void end_contour(contours_t* c) { c->ncont++; c->start[c->ncont] = c->npts; }

/* Border following in discovery order. bin: nonzero = foreground. The image is used with a
 * one-pixel zero frame around it, as cv::findContours does, so borders may run along the image
 * edge. Pixel states: 0 background, 1 untouched foreground, 2 visited, -126 visited with the
 * "right neighbour examined and zero" flag (Suzuki-Abe 1985, appendix; border number fixed to 2
 * because RETR_LIST keeps no hierarchy). */
static void find_contours(const uint8_t* bin, int w, int h, int stride, contours_t* out)
{
    const int W = w + 2, H = h + 2;
    int8_t* img = (int8_t*)calloc((size_t)W * H, 1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
            img[(size_t)(y + 1) * W + x + 1] = bin[(size_t)y * stride + x] ? 1 : 0;
    /* directions: 0 = east, then counter-clockwise on the screen (y grows downwards): NE, N, NW, W, SW, S, SE */
    const int dxs[8] = { 1, 1, 0, -1, -1, -1, 0, 1 };
    const int dys[8] = { 0, -1, -1, -1, 0, 1, 1, 1 };
    int delta[16];
    for (int k = 0; k < 16; k++) delta[k] = dys[k & 7] * W + dxs[k & 7];

    for (int y = 1; y < H - 1; y++)
    {
        int8_t* row = img + (size_t)y * W;
        int prev = 0;
        for (int x = 1; x < W - 1; x++)
        {
            int p = row[x];
            if (p == prev) continue;
            int is_hole = 0;
            if (!(prev == 0 && p == 1))
            {
                if (p != 0 || prev < 1) { prev = p; continue; }
                is_hole = 1;
            }
            /* trace the border that starts at (x - is_hole, y) */
            int8_t* i0 = row + x - is_hole;
            int px = x - is_hole, py = y;
            begin_contour(out);
            int s_end = is_hole ? 0 : 4, s = s_end;
            int8_t* i1;
            do { s = (s - 1) & 7; i1 = i0 + delta[s]; } while (*i1 == 0 && s != s_end);
            if (s == s_end)
            {
                *i0 = (int8_t)(2 | -128);              /* isolated pixel */
                push_pt(out, px - 1, py - 1);
            }
            else
            {
                int8_t* i3 = i0;
                for (;;)
                {
                    s_end = s;
                    int8_t* i4;
                    for (;;) { i4 = i3 + delta[++s]; if (*i4 != 0) break; }
                    s &= 7;
                    if ((unsigned)(s - 1) < (unsigned)s_end) *i3 = (int8_t)(2 | -128);
                    else if (*i3 == 1) *i3 = 2;
                    push_pt(out, px - 1, py - 1);
                    px += dxs[s]; py += dys[s];
                    if (i4 == i0 && i3 == i1) break;
                    i3 = i4;
                    s = (s + 4) & 7;
                }
            }
            end_contour(out);
            prev = row[x];
        }
    }
    free(img);
}

/* cv2.findContours(RETR_LIST, CHAIN_APPROX_NONE) for tests: contours in OpenCV's order (the reverse
 * of discovery). xy = [total][2] int32, lens = [ncont]. Returns the number of contours, or -1 - n
 * if a capacity is too small. */
EXPORT int blob_oracle_find_contours(const uint8_t* bin, int w, int h, int stride, int32_t* xy, int max_pts, int32_t* lens, int max_cont)
{
    contours_t c; memset(&c, 0, sizeof(c));
    find_contours(bin, w, h, stride, &c);
    int ret = c.ncont;
    if (c.npts > max_pts || c.ncont > max_cont) ret = -1 - c.ncont;
    else
    {
        int o = 0;
        for (int i = c.ncont - 1; i >= 0; i--)
        {
            lens[c.ncont - 1 - i] = c.start[i + 1] - c.start[i];
            for (int k = c.start[i]; k < c.start[i + 1]; k++) { xy[2*o] = c.pts[k].x; xy[2*o + 1] = c.pts[k].y; o++; }
        }
    }
    free(c.pts); free(c.start);
    return ret;
}

typedef struct { double x, y, radius, confidence; } center_t;

static int cmp_pt(const void* a, const void* b)
{
    const pt_t* p = (const pt_t*)a; const pt_t* q = (const pt_t*)b;
    if (p->x != q->x) return p->x < q->x ? -1 : 1;
    return p->y < q->y ? -1 : (p->y > q->y);
}
static int cmp_dbl(const void* a, const void* b) { const double p = *(const double*)a, q = *(const double*)b; return p < q ? -1 : (p > q); }
static long long cross(pt_t o, pt_t a, pt_t b) { return (long long)(a.x - o.x) * (b.y - o.y) - (long long)(a.y - o.y) * (b.x - o.x); }

/* twice the area of the convex hull of the points (monotone chain) */
static long long hull_area2(const pt_t* pts, int n)
{
    pt_t* s = (pt_t*)malloc(sizeof(pt_t) * n);
    memcpy(s, pts, sizeof(pt_t) * n);
    qsort(s, n, sizeof(pt_t), cmp_pt);
    pt_t* hl = (pt_t*)malloc(sizeof(pt_t) * (2 * n + 2));
    int k = 0;
    for (int i = 0; i < n; i++) { while (k >= 2 && cross(hl[k-2], hl[k-1], s[i]) <= 0) k--; hl[k++] = s[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; i--) { while (k >= t && cross(hl[k-2], hl[k-1], s[i]) <= 0) k--; hl[k++] = s[i]; }
    if (k > 1) k--;
    long long a2 = 0;
    for (int i = 0; i < k; i++) { const pt_t p = hl[i], q = hl[(i + 1) % k]; a2 += (long long)p.x * q.y - (long long)q.x * p.y; }
    free(s); free(hl);
    return a2 < 0 ? -a2 : a2;
}

static double cv_round_half_even(double v) { return nearbyint(v); }   /* cvRound: default rounding mode */

/* SimpleBlobDetector::findBlobs for one binary image (foreground = gray > thresh) */
static int find_blobs_one(const uint8_t* gray, int w, int h, int stride, int thresh, center_t** out, int* nout, int* capout)
{
    uint8_t* bin = (uint8_t*)malloc((size_t)w * h);
    for (int y = 0; y < h; y++) for (int x = 0; x < w; x++) bin[(size_t)y * w + x] = gray[(size_t)y * stride + x] > thresh;
    contours_t c; memset(&c, 0, sizeof(c));
    find_contours(bin, w, h, w, &c);
    for (int ci = c.ncont - 1; ci >= 0; ci--)            /* OpenCV hands the contours over in reverse discovery order */
    {
        const pt_t* p = c.pts + c.start[ci];
        const int n = c.start[ci + 1] - c.start[ci];
        /* contour moments (Green's theorem over the closed polygon); integer-valued, exact in double */
        double a00 = 0, a10 = 0, a01 = 0, a20 = 0, a11 = 0, a02 = 0;
        double xi_1 = p[n - 1].x, yi_1 = p[n - 1].y;
        for (int i = 0; i < n; i++)
        {
            const double xi = p[i].x, yi = p[i].y, xi2 = xi * xi, yi2 = yi * yi;
            const double dxy = xi_1 * yi - xi * yi_1, xii_1 = xi_1 + xi, yii_1 = yi_1 + yi;
            a00 += dxy; a10 += dxy * xii_1; a01 += dxy * yii_1;
            a20 += dxy * (xi_1 * xii_1 + xi2);
            a11 += dxy * (xi_1 * (yii_1 + yi_1) + xi * (yii_1 + yi));
            a02 += dxy * (yi_1 * yii_1 + yi2);
            xi_1 = xi; yi_1 = yi;
        }
        double m00 = 0, m10 = 0, m01 = 0, m20 = 0, m11 = 0, m02 = 0;
        if (fabs(a00) > FLT_EPSILON)
        {
            double db1_2, db1_6, db1_12, db1_24;
            if (a00 > 0) { db1_2 = 0.5; db1_6 = 0.16666666666666666666666666666667; db1_12 = 0.083333333333333333333333333333333; db1_24 = 0.041666666666666666666666666666667; }
            else { db1_2 = -0.5; db1_6 = -0.16666666666666666666666666666667; db1_12 = -0.083333333333333333333333333333333; db1_24 = -0.041666666666666666666666666666667; }
            m00 = a00 * db1_2; m10 = a10 * db1_6; m01 = a01 * db1_6; m20 = a20 * db1_12; m11 = a11 * db1_24; m02 = a02 * db1_12;
        }
        double mu20 = 0, mu11 = 0, mu02 = 0;
        if (fabs(m00) > DBL_EPSILON)
        {
            const double inv_m00 = 1. / m00, cx = m10 * inv_m00, cy = m01 * inv_m00;
            mu20 = m20 - m10 * cx; mu11 = m11 - m10 * cy; mu02 = m02 - m01 * cy;
        }
        center_t ctr; ctr.confidence = 1;
        /* filterByArea */
        { const double area = m00; if (area < 20.0f || area >= 80000.0f) continue; }
        /* filterByInertia (min 0.1f, max FLT_MAX) */
        {
            const double denominator = sqrt(pow(2 * mu11, 2) + pow(mu20 - mu02, 2));
            const double eps = 1e-2;
            double ratio;
            if (denominator > eps)
            {
                const double cosmin = (mu20 - mu02) / denominator, sinmin = 2 * mu11 / denominator;
                const double cosmax = -cosmin, sinmax = -sinmin;
                const double imin = 0.5 * (mu20 + mu02) - 0.5 * (mu20 - mu02) * cosmin - mu11 * sinmin;
                const double imax = 0.5 * (mu20 + mu02) - 0.5 * (mu20 - mu02) * cosmax - mu11 * sinmax;
                ratio = imin / imax;
            }
            else ratio = 1;
            if (ratio < 0.1f || ratio >= FLT_MAX) continue;
            ctr.confidence = ratio * ratio;
        }
        /* filterByConvexity (min 0.95f) */
        {
            const double area = m00, hullArea = 0.5 * (double)hull_area2(p, n);
            if (fabs(hullArea) < DBL_EPSILON) continue;
            const double ratio = area / hullArea;
            if (ratio < 0.95f || ratio >= FLT_MAX) continue;
        }
        if (m00 == 0.0) continue;
        ctr.x = m10 / m00; ctr.y = m01 / m00;
        /* filterByColor: blobColor = 0 */
        {
            const int ry = (int)cv_round_half_even(ctr.y), rx = (int)cv_round_half_even(ctr.x);
            if (bin[(size_t)ry * w + rx] != 0) continue;
        }
        /* radius = median distance of the contour points to the centre */
        {
            double* d = (double*)malloc(sizeof(double) * n);
            for (int i = 0; i < n; i++) { const double dx = ctr.x - p[i].x, dy = ctr.y - p[i].y; d[i] = sqrt(dx * dx + dy * dy); }
            qsort(d, n, sizeof(double), cmp_dbl);
            ctr.radius = (d[(n - 1) / 2] + d[n / 2]) / 2.;
            free(d);
        }
        if (*nout == *capout) { *capout = *capout ? 2 * *capout : 256; *out = (center_t*)realloc(*out, sizeof(center_t) * *capout); }
        (*out)[(*nout)++] = ctr;
    }
    free(c.pts); free(c.start); free(bin);
    return 0;
}

/* for tests of the per-threshold stage: centres (x, y, radius, confidence) of one threshold */
EXPORT int blob_oracle_centers(const uint8_t* gray, int w, int h, int stride, int thresh, double* out4, int max_out)
{
    center_t* c = NULL; int n = 0, cap = 0;
    find_blobs_one(gray, w, h, stride, thresh, &c, &n, &cap);
    const int ret = n <= max_out ? n : -1 - n;
    for (int i = 0; i < n && i < max_out; i++) { out4[4*i] = c[i].x; out4[4*i+1] = c[i].y; out4[4*i+2] = c[i].radius; out4[4*i+3] = c[i].confidence; }
    free(c);
    return ret;
}

typedef struct { center_t* v; int n, cap; } group_t;

/* find_blobs_from_image_array (find_blobs.cc:14-46): xy_scaled = [n][2] int32, coordinates * 1000.
 * Returns the number of points, or -1 - n if max_points is too small. */
EXPORT int blob_oracle_find_blobs(const uint8_t* gray, int w, int h, int stride, int32_t* xy_scaled, int max_points)
{
    group_t* groups = NULL; int ngroups = 0, capgroups = 0;
    for (double thresh = 50; thresh < 220; thresh += 10)
    {
        center_t* cur = NULL; int ncur = 0, capcur = 0;
        find_blobs_one(gray, w, h, stride, (int)floor(thresh), &cur, &ncur, &capcur);
        const int nold = ngroups;
        for (int i = 0; i < ncur; i++)
        {
            int is_new = 1;
            for (int j = 0; j < nold; j++)
            {
                group_t* g = &groups[j];
                const center_t* mid = &g->v[g->n / 2];
                const double dx = mid->x - cur[i].x, dy = mid->y - cur[i].y, dist = sqrt(dx * dx + dy * dy);
                is_new = dist >= 5.0f && dist >= mid->radius && dist >= cur[i].radius;
                if (!is_new)
                {
                    if (g->n == g->cap) { g->cap *= 2; g->v = (center_t*)realloc(g->v, sizeof(center_t) * g->cap); }
                    g->v[g->n++] = cur[i];
                    int k = g->n - 1;
                    while (k > 0 && cur[i].radius < g->v[k - 1].radius) { g->v[k] = g->v[k - 1]; k--; }
                    g->v[k] = cur[i];
                    break;
                }
            }
            if (is_new)
            {
                if (ngroups == capgroups) { capgroups = capgroups ? 2 * capgroups : 256; groups = (group_t*)realloc(groups, sizeof(group_t) * capgroups); }
                group_t* g = &groups[ngroups++];
                g->cap = 8; g->n = 1; g->v = (center_t*)malloc(sizeof(center_t) * g->cap); g->v[0] = cur[i];
            }
        }
        free(cur);
    }
    int n = 0;
    for (int i = 0; i < ngroups; i++)
    {
        group_t* g = &groups[i];
        if (g->n >= 2)
        {
            double sx = 0, sy = 0, normalizer = 0;
            for (int j = 0; j < g->n; j++) { sx += g->v[j].confidence * g->v[j].x; sy += g->v[j].confidence * g->v[j].y; normalizer += g->v[j].confidence; }
            const double inv = 1. / normalizer;
            sx *= inv; sy *= inv;
            const float fx = (float)sx, fy = (float)sy;              /* KeyPoint::pt is Point2f */
            if (n < max_points)
            {
                /* find_blobs.cc:40-41: float * int -> float product, + 0.5 in double, truncation */
                const float px = fx * 1000, py = fy * 1000;
                xy_scaled[2*n] = (int)(px + 0.5); xy_scaled[2*n + 1] = (int)(py + 0.5);
            }
            n++;
        }
        free(g->v);
    }
    free(groups);
    return n <= max_points ? n : -1 - n;
}


/* ------------------------------------------------------------------------------------------------
 * Preprocessing the reference CLI applies before the detector by default:
 * cv::blur(image, image, Size(1+2R, 1+2R)) with R = 1 (mrgingham-from-image.cc:106-111, default
 * --blur 1 at :222). OpenCV's normalised box filter on 8-bit data with BORDER_REFLECT_101 is, for
 * odd k*k, exactly out = floor((sum + (k*k-1)/2) / (k*k)) (no ties exist, so the float rounding
 * inside OpenCV cannot matter). PINNED against cv2.blur 4.13.0 by tests/test_blur.py.
 * ------------------------------------------------------------------------------------------------ */
static int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) { if (i < 0) i = -i; else i = 2 * (n - 1) - i; }
    return i;
}
EXPORT void blob_oracle_box_blur(const uint8_t* in, int w, int h, int stride, int radius, uint8_t* out /* dense [h][w] */)
{
    const int k2 = (2 * radius + 1) * (2 * radius + 1);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            int sum = 0;
            for (int dy = -radius; dy <= radius; dy++)
                for (int dx = -radius; dx <= radius; dx++)
                    sum += in[(size_t)reflect101(y + dy, h) * stride + reflect101(x + dx, w)];
            out[(size_t)y * w + x] = (uint8_t)((sum + (k2 - 1) / 2) / k2);
        }
}
