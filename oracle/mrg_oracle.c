/*
  TEST INFRASTRUCTURE ONLY -- the parity oracle. Never shipped, never linked into or called by the
  product library (mrgingham_b200/). Only tests/, __graft_entry__.smoke() and bench.py's CPU
  baseline leg may load this.

  A plain-C, CPU restatement of the reference's per-image corner-detection hot path, written from
  the behaviour of the reference (file:line in /root/reference cited per function). Parity status:
  PINNED -- tests/test_oracle_vs_ref.py checks every function here bit-for-bit against the
  reference itself (oracle/_ref, compiled from /root/reference by oracle/Makefile) and
  tests/golden/ holds vectors generated from that reference build (the reference ships no golden
  vectors of its own for this path).

  Third-party arithmetic on the path: OpenCV cv::resize(INTER_LINEAR) (call site
  find_chessboard_corners.cc:450; OpenCV version not pinned by the reference, in-container
  stand-in cv2 4.13.0). Restated here as the integer model "N1" and pinned against cv2 by
  tests/test_pyramid_model.py.
*/
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdbool.h>
#include <math.h>
#include <stddef.h>

#define API __attribute__((visibility("default")))

/* detector constants: find_chessboard_corners.cc:18,22,27,29,38,39 and the margin at :564 */
enum {
    PEAK_MIN        = 120,   /* component peak must exceed this          */
    RESP_MIN        = 15,    /* a pixel must exceed this to be a member  */
    COMPONENT_MIN_N = 2,
    VAR_WINDOW_R    = 10,    /* 21x21 window                             */
    VAR_MIN         = 400,   /* stdev 20                                 */
    CHESS_MARGIN    = 7
};

/* ------------------------------------------------------------------------------------------
   ChESS response, ring radius 5.  Follows ChESS.c:62-105.
   Writes only 7 <= x < w-7, 7 <= y < h-7; everything else in `response` is left untouched.
   ------------------------------------------------------------------------------------------ */
static const int8_t ring_dx[16] = { 2, 0,-2,-4,-5,-5,-5,-4,-2, 0, 2, 4, 5, 5, 5, 4};
static const int8_t ring_dy[16] = {-5,-5,-5,-4,-2, 0, 2, 4, 5, 5, 5, 4, 2, 0,-2,-4};

API void oracle_chess_response_5(int16_t* response, const uint8_t* image, int w, int h, int stride)
{
    for(int y = 7; y < h-7; y++)
    {
        const uint8_t* c = image + (size_t)y*stride;
        for(int x = 7; x < w-7; x++)
        {
            int s[16];
            for(int k = 0; k < 16; k++)
                s[k] = c[x + ring_dx[k] + ring_dy[k]*(ptrdiff_t)stride];

            /* horizontal 3-sample local mean scaled to 16 samples; C integer division truncates
               (operands are non-negative). Held in uint16 by the reference; max 4080, no wrap */
            const int local_mean = (c[x-1] + c[x] + c[x+1]) * 16 / 3;

            int sum = 0, diff = 0, mean = 0;
            for(int i = 0; i < 4; i++)
            {
                const int a = s[i], b = s[i+4], cc = s[i+8], d = s[i+12];
                sum  += abs(a - b + cc - d);
                diff += abs(a - cc) + abs(b - d);
                mean += a + b + cc + d;
            }
            response[x + (size_t)y*w] = (int16_t)(sum - diff - abs(mean - local_mean));
        }
    }
}

/* ------------------------------------------------------------------------------------------
   Pyramid level: what cv::resize(in, out, Size(), 1/2^L, 1/2^L, INTER_LINEAR) produces
   (find_chessboard_corners.cc:449-450), model N1 (see header). out may be NULL to query the size.
   ------------------------------------------------------------------------------------------ */
static int round_half_even_div(int n, int d) /* rint(n/d) for n >= 0, d a power of two */
{
    int q = n / d, r = n % d;
    if(2*r > d || (2*r == d && (q & 1))) q++;
    return q;
}
static int rint_half(int s) { int q = s >> 1; return q + ((s & 1) & (q & 1)); }

API int oracle_pyramid(const uint8_t* in, int rows, int cols, int stride, int level,
                       uint8_t* out, int* orows, int* ocols)
{
    if(level < 0 || level > 10) return -1;
    const int B = 1 << level;
    const int ow = round_half_even_div(cols, B), oh = round_half_even_div(rows, B);
    *orows = oh; *ocols = ow;
    if(out == NULL) return 0;
    if(level == 0)
    {
        for(int y = 0; y < rows; y++) memcpy(out + (size_t)y*cols, in + (size_t)y*stride, cols);
        return 0;
    }
    for(int dy = 0; dy < oh; dy++)
    {
        int y0 = B*dy + B/2 - 1; if(y0 > rows-1) y0 = rows-1;
        int y1 = y0 + 1;         if(y1 > rows-1) y1 = rows-1;
        for(int dx = 0; dx < ow; dx++)
        {
            int x0 = B*dx + B/2 - 1; if(x0 > cols-1) x0 = cols-1;
            int x1 = x0 + 1;         if(x1 > cols-1) x1 = cols-1;
            const int v = in[(size_t)y0*stride + x0] + in[(size_t)y0*stride + x1] +
                          in[(size_t)y1*stride + x0] + in[(size_t)y1*stride + x1];
            out[(size_t)dy*ow + dx] = (uint8_t)((v + 2) >> 2);
        }
    }
    if(level == 1)
    {
        /* OpenCV's exact-2x path averages only the pixels that exist in a trailing partial cell
           and rounds that float mean half-to-even */
        const bool px = 2*ow > cols, py = 2*oh > rows;
        if(px)
            for(int dy = 0; dy < oh; dy++)
            {
                const int ya = 2*dy, yb = (ya+1 > rows-1) ? rows-1 : ya+1;
                out[(size_t)dy*ow + ow-1] =
                    (uint8_t)rint_half(in[(size_t)ya*stride + cols-1] + in[(size_t)yb*stride + cols-1]);
            }
        if(py)
            for(int dx = 0; dx < ow; dx++)
            {
                const int xa = 2*dx, xb = (xa+1 > cols-1) ? cols-1 : xa+1;
                out[(size_t)(oh-1)*ow + dx] =
                    (uint8_t)rint_half(in[(size_t)(rows-1)*stride + xa] + in[(size_t)(rows-1)*stride + xb]);
            }
        if(px && py)
            out[(size_t)(oh-1)*ow + ow-1] = in[(size_t)(rows-1)*stride + cols-1];
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
   21x21 intensity-variance gate at the component's peak.  find_chessboard_corners.cc:50-88
   ------------------------------------------------------------------------------------------ */
static bool variance_gate(int x, int y, int w, int h, const uint8_t* img)
{
    if(x - VAR_WINDOW_R < 0 || x + VAR_WINDOW_R >= w ||
       y - VAR_WINDOW_R < 0 || y + VAR_WINDOW_R >= h)
        return false;
    const int n = (2*VAR_WINDOW_R+1)*(2*VAR_WINDOW_R+1);
    int32_t sum = 0;
    for(int v = y-VAR_WINDOW_R; v <= y+VAR_WINDOW_R; v++)
        for(int u = x-VAR_WINDOW_R; u <= x+VAR_WINDOW_R; u++)
            sum += img[u + (size_t)v*w];
    const int32_t mean = sum / n;
    int32_t ssd = 0;
    for(int v = y-VAR_WINDOW_R; v <= y+VAR_WINDOW_R; v++)
        for(int u = x-VAR_WINDOW_R; u <= x+VAR_WINDOW_R; u++)
        {
            const int32_t e = (int32_t)img[u + (size_t)v*w] - mean;
            ssd += e*e;
        }
    return ssd / n > VAR_MIN;
}

/* ------------------------------------------------------------------------------------------
   One connected component.  find_chessboard_corners.cc:159-267.
   LIFO traversal; a popped pixel is a member iff it is inside the image, its response is > 15
   and > (running max >> 4); non-members are zeroed and NOT expanded. Members are accumulated,
   zeroed, and their neighbours pushed in the order +x, -x, +y, -y; a neighbour in the 7-pixel
   border poisons the component (checked before the response), a neighbour with response <= 0 is
   skipped. Accept iff !poisoned && N >= 2 && max > 120 && variance gate at the peak.
   ------------------------------------------------------------------------------------------ */
typedef struct { int16_t x, y; } pix_t;
typedef struct { pix_t* p; int n, cap; } pixstack_t;

static void stack_push(pixstack_t* s, int x, int y)
{
    if(s->n == s->cap)
    {
        s->cap = s->cap ? 2*s->cap : 256;
        s->p = (pix_t*)realloc(s->p, (size_t)s->cap * sizeof(pix_t));
    }
    s->p[s->n].x = (int16_t)x; s->p[s->n].y = (int16_t)y; s->n++;
}

typedef struct
{
    uint64_t swx, swy, sw;
    int n, peak_x, peak_y, peak;
    bool poisoned;
} comp_t;

static void try_push(pixstack_t* s, comp_t* c, int x, int y, int w, int h, const int16_t* r)
{
    if(x < CHESS_MARGIN || x >= w-CHESS_MARGIN || y < CHESS_MARGIN || y >= h-CHESS_MARGIN)
    {
        c->poisoned = true;
        return;
    }
    if(r[x + (size_t)y*w] > 0) stack_push(s, x, y);
}

static bool grow_component(double* cx, double* cy, pixstack_t* s, int w, int h, int16_t* r,
                           const uint8_t* img)
{
    comp_t c; memset(&c, 0, sizeof(c));
    while(s->n > 0)
    {
        s->n--;
        const int x = s->p[s->n].x, y = s->p[s->n].y;
        bool member = false;
        int v = 0;
        if(x >= 0 && x < w && y >= 0 && y < h)
        {
            v = r[x + (size_t)y*w];
            member = v > RESP_MIN && v > (int)((uint16_t)c.peak >> 4);
        }
        if(!member)
        {
            /* the reference writes d[x+y*w]=0 even for an out-of-image (x,y); such coordinates can
               only come from refinement seeds, which are bounds-checked before being pushed, so the
               write is always in range in practice. Guard it here. */
            if(x >= 0 && x < w && y >= 0 && y < h) r[x + (size_t)y*w] = 0;
            continue;
        }
        if(v > c.peak) { c.peak = v; c.peak_x = x; c.peak_y = y; }
        c.swx += (uint64_t)(v*x); c.swy += (uint64_t)(v*y); c.sw += (uint64_t)v; c.n++;
        r[x + (size_t)y*w] = 0;
        try_push(s, &c, x+1, y,   w, h, r);
        try_push(s, &c, x-1, y,   w, h, r);
        try_push(s, &c, x,   y+1, w, h, r);
        try_push(s, &c, x,   y-1, w, h, r);
    }
    if(c.poisoned || c.n < COMPONENT_MIN_N || c.peak <= PEAK_MIN) return false;
    if(!variance_gate(c.peak_x, c.peak_y, w, h, img)) return false;
    *cx = (double)c.swx / (double)c.sw;
    *cy = (double)c.swy / (double)c.sw;
    return true;
}

/* pixel-centre-preserving rescale, find_chessboard_corners.cc:269-280 */
static double rescale(double p, double scale) { return (p + 0.5)*scale - 0.5; }

/* level image + clamped response; returns 0 on the reference's error paths
   (find_chessboard_corners.cc:433-441, :461-466) */
typedef struct { uint8_t* img; bool img_owned; int16_t* resp; int w, h; } level_t;

static bool level_prepare(level_t* lv, const uint8_t* image, int rows, int cols, int stride, int level)
{
    memset(lv, 0, sizeof(*lv));
    if(level < 0 || level > 10) return false;
    if(level == 0)
    {
        if(rows > 1 && stride != cols) return false; /* "I can only handle continuous arrays" */
        lv->img = (uint8_t*)image; lv->w = cols; lv->h = rows;
    }
    else
    {
        int oh, ow;
        oracle_pyramid(image, rows, cols, stride, level, NULL, &oh, &ow);
        lv->img = (uint8_t*)malloc((size_t)oh*ow + 1); lv->img_owned = true;
        oracle_pyramid(image, rows, cols, stride, level, lv->img, &oh, &ow);
        lv->w = ow; lv->h = oh;
    }
    /* zero-initialised response (:506), ChESS (:511), negatives clamped to 0 (:527-529) */
    lv->resp = (int16_t*)calloc((size_t)lv->w*lv->h + 1, sizeof(int16_t));
    oracle_chess_response_5(lv->resp, lv->img, lv->w, lv->h, lv->w);
    for(size_t i = 0; i < (size_t)lv->w*lv->h; i++) if(lv->resp[i] < 0) lv->resp[i] = 0;
    return true;
}
static void level_release(level_t* lv)
{
    if(lv->img_owned) free(lv->img);
    free(lv->resp);
}

/* ------------------------------------------------------------------------------------------
   find_chessboard_corners_from_image_array(): find_chessboard_corners.cc:568-587 -> :481-565
   -> find branch of process_connected_components :330-355.
   Returns the number of points found (>= 0); writes min(N,cap) points, each
   (int)(0.5 + coord*1000) in full-resolution pixels.
   If xy_double != NULL also writes the un-quantised full-resolution doubles.
   ------------------------------------------------------------------------------------------ */
API int oracle_find_corners(const uint8_t* image, int rows, int cols, int stride, int level,
                            int* xy_out, double* xy_double, int cap)
{
    level_t lv;
    if(!level_prepare(&lv, image, rows, cols, stride, level)) return 0;
    const int w = lv.w, h = lv.h;
    const double scale = (double)(uint16_t)(1U << level);
    pixstack_t st = {0};
    int n = 0;
    for(int y = CHESS_MARGIN+1; y < h-CHESS_MARGIN-1; y++)
        for(int x = CHESS_MARGIN+1; x < w-CHESS_MARGIN-1; x++)
        {
            if(lv.resp[x + (size_t)y*w] <= RESP_MIN) continue;
            st.n = 0; stack_push(&st, x, y);
            double cx, cy;
            if(!grow_component(&cx, &cy, &st, w, h, lv.resp, lv.img)) continue;
            const double fx = rescale(cx, scale), fy = rescale(cy, scale);
            if(n < cap)
            {
                if(xy_out)    { xy_out[2*n] = (int)(0.5 + fx*1000); xy_out[2*n+1] = (int)(0.5 + fy*1000); }
                if(xy_double) { xy_double[2*n] = fx; xy_double[2*n+1] = fy; }
            }
            n++;
        }
    free(st.p);
    level_release(&lv);
    return n;
}

/* ------------------------------------------------------------------------------------------
   refine_chessboard_corners_from_image_array(): find_chessboard_corners.cc:591-619 -> refine
   branch :356-397. Points whose levels[i] == level+1 are re-seeded from the 3x3 around their
   rounded position at this level (dx outer, dy inner), in point order, on one shared response.
   Returns the number of points refined; updates xy (full-res doubles) and levels in place.
   ------------------------------------------------------------------------------------------ */
API int oracle_refine_corners(const uint8_t* image, int rows, int cols, int stride, int level,
                              double* xy, signed char* levels, int npoints)
{
    level_t lv;
    if(!level_prepare(&lv, image, rows, cols, stride, level)) return 0;
    const int w = lv.w, h = lv.h;
    const double scale = (double)(uint16_t)(1U << level);
    pixstack_t st = {0};
    int nrefined = 0;
    for(int i = 0; i < npoints; i++)
    {
        if(levels[i] != level+1) continue;
        const int x = (int)(rescale(xy[2*i],   1.0/scale) + 0.5);
        const int y = (int)(rescale(xy[2*i+1], 1.0/scale) + 0.5);
        st.n = 0;
        for(int dx = -1; dx <= 1; dx++)
            for(int dy = -1; dy <= 1; dy++)
            {
                const int u = x+dx, v = y+dy;
                if(u < 0 || u >= w || v < 0 || v >= h) continue;
                if(lv.resp[u + (size_t)v*w] > RESP_MIN) stack_push(&st, u, v);
            }
        double cx, cy;
        if(!grow_component(&cx, &cy, &st, w, h, lv.resp, lv.img)) continue;
        xy[2*i]   = rescale(cx, scale);
        xy[2*i+1] = rescale(cy, scale);
        levels[i] = (signed char)level;
        nrefined++;
    }
    free(st.p);
    level_release(&lv);
    return nrefined;
}
