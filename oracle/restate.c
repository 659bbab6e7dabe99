/* CPU restatement of the OpenCV-resident stages of the Polychase analyze path.
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into the product library.  Built by
 * oracle/build.py (gcc) into oracle/_build/liboracle_restate.so and used by tests/ to
 * check the CUDA kernels, after being pinned itself against cv2 4.13.0
 * (tests/test_oracle_vs_cv2.py).
 *
 * The reference does not own this arithmetic: it calls OpenCV
 *   cv::cvtColor RGB2GRAY            /root/reference/cpp/opticalflow.cc:259,298
 *   cv::buildOpticalFlowPyramid      /root/reference/cpp/opticalflow.cc:184-186
 *   cv::calcOpticalFlowPyrLK         /root/reference/cpp/opticalflow.cc:119-125
 *   cv::cornerMinEigenVal            /root/reference/cpp/feature_detection/gftt.cc:35-36
 * (opencv4 4.11.x via vcpkg tag 2025.06.13; source not in /root/reference).  The
 * formulas below restate OpenCV 4.x's published algorithms.
 *
 * Compile with -ffp-contract=off: every fused multiply-add below is explicit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline int reflect101(int p, int len) {
    /* cv::borderInterpolate(p, len, BORDER_REFLECT_101) */
    if ((unsigned)p < (unsigned)len) return p;
    if (len == 1) return 0;
    do {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    } while ((unsigned)p >= (unsigned)len);
    return p;
}

/* A.1: g = (R*9798 + G*19235 + B*3735 + 2^14) >> 15 */
void orc_rgb2gray(const uint8_t* rgb, int w, int h, size_t stride, uint8_t* gray) {
    for (int y = 0; y < h; y++) {
        const uint8_t* s = rgb + (size_t)y * stride;
        uint8_t* d = gray + (size_t)y * w;
        for (int x = 0; x < w; x++)
            d[x] = (uint8_t)((s[3 * x] * 9798 + s[3 * x + 1] * 19235 + s[3 * x + 2] * 3735 + (1 << 14)) >> 15);
    }
}

/* A.2: pyrDown, 5x5 binomial, REFLECT_101, out = (sum + 128) >> 8, size (w+1)/2 */
void orc_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst) {
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    int* tmp = (int*)malloc(sizeof(int) * (size_t)dw * h);
    for (int y = 0; y < h; y++) {
        const uint8_t* s = src + (size_t)y * w;
        for (int x = 0; x < dw; x++) {
            const int c = 2 * x;
            tmp[(size_t)y * dw + x] = s[reflect101(c - 2, w)] + 4 * s[reflect101(c - 1, w)] + 6 * s[c] +
                                      4 * s[reflect101(c + 1, w)] + s[reflect101(c + 2, w)];
        }
    }
    for (int y = 0; y < dh; y++) {
        const int c = 2 * y;
        const int* r0 = tmp + (size_t)reflect101(c - 2, h) * dw;
        const int* r1 = tmp + (size_t)reflect101(c - 1, h) * dw;
        const int* r2 = tmp + (size_t)c * dw;
        const int* r3 = tmp + (size_t)reflect101(c + 1, h) * dw;
        const int* r4 = tmp + (size_t)reflect101(c + 2, h) * dw;
        for (int x = 0; x < dw; x++)
            dst[(size_t)y * dw + x] = (uint8_t)((r0[x] + 4 * r1[x] + 6 * r2[x] + 4 * r3[x] + r4[x] + 128) >> 8);
    }
    free(tmp);
}

/* A.2: Scharr derivative image (int16 x2 interleaved dx,dy), REFLECT_101 */
void orc_scharr(const uint8_t* src, int w, int h, int16_t* d) {
    for (int y = 0; y < h; y++) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * w;
        const uint8_t* r1 = src + (size_t)y * w;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            const int t0m = 3 * (r0[xm] + r2[xm]) + 10 * r1[xm];
            const int t0p = 3 * (r0[xp] + r2[xp]) + 10 * r1[xp];
            const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
            d[((size_t)y * w + x) * 2 + 0] = (int16_t)(t0p - t0m);
            d[((size_t)y * w + x) * 2 + 1] = (int16_t)(3 * (t1m + t1p) + 10 * t1c);
        }
    }
}

/* A.4: cornerMinEigenVal(u8, blockSize=3, ksize=3), REFLECT_101.
 * mode bit0: 1 = AVX2-dispatched Sobel op order (what cv2 / the reference wheel run on
 *            x86-64: FMA everywhere except the w%32 tail columns of the row-smoothing
 *            filter, which are plain mul+add), 0 = plain C++ op order (cv::setUseOptimized(false)).
 * mode bit1: 0 = 3x3 box sums exactly as OpenCV's ColumnSum<double,float> forms them: a
 *            running double sum per column, SUM += row[y+1]; out = (float)SUM;
 *            SUM -= row[y-1], carried from the top of the image (its rounding history
 *            shows up in ~2 pixels per million), 1 = each 3x3 sum formed independently
 *            (the correctly rounded sum; what the CUDA kernel computes). */
void orc_min_eig(const uint8_t* src, int w, int h, int mode, float* eig) {
    const float s = (float)(1.0 / (4.0 * 3.0 * 255.0));
    const float s2 = 2.0f * s;
    const int fma_mode = mode & 1, indep_box = (mode >> 1) & 1;
    /* measured against cv2 4.13.0: the row filter's vector body covers 32 columns per
     * step and its w%32 tail is plain mul+add; the column filter is fused everywhere. */
    const int wvec = fma_mode ? (w / 32) * 32 : 0;
    const size_t n = (size_t)w * h;
    float* rx = (float*)malloc(sizeof(float) * n);   /* p[x+1]-p[x-1] */
    float* rs = (float*)malloc(sizeof(float) * n);   /* row-smoothed */
    float* cov = (float*)malloc(sizeof(float) * n * 3);
    double* rsum = (double*)malloc(sizeof(double) * n * 3);
    for (int y = 0; y < h; y++) {
        const uint8_t* r = src + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            const float pm = (float)r[reflect101(x - 1, w)], pc = (float)r[x], pp = (float)r[reflect101(x + 1, w)];
            rx[(size_t)y * w + x] = pp - pm;
            if (x < wvec)
                rs[(size_t)y * w + x] = fmaf(pp, s, fmaf(pc, s2, pm * s));
            else
                rs[(size_t)y * w + x] = (pm * s + pc * s2) + pp * s;
        }
    }
    for (int y = 0; y < h; y++) {
        const size_t ym = (size_t)reflect101(y - 1, h) * w, yc = (size_t)y * w, yp = (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            float dx;
            if (fma_mode)
                dx = fmaf(rx[ym + x] + rx[yp + x], s, rx[yc + x] * s2);
            else
                dx = (rx[ym + x] + rx[yp + x]) * s + rx[yc + x] * s2;
            const float dy = rs[yp + x] - rs[ym + x];
            cov[(yc + x) * 3 + 0] = dx * dx;
            cov[(yc + x) * 3 + 1] = dx * dy;
            cov[(yc + x) * 3 + 2] = dy * dy;
        }
    }
    /* RowSum<float,double>, ksize 3: D = S[x-1] + S[x] + S[x+1] (left to right) */
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const size_t xm = (size_t)y * w + reflect101(x - 1, w), xc = (size_t)y * w + x,
                         xp = (size_t)y * w + reflect101(x + 1, w);
            for (int c = 0; c < 3; c++)
                rsum[xc * 3 + c] = ((double)cov[xm * 3 + c] + (double)cov[xc * 3 + c]) + (double)cov[xp * 3 + c];
        }
    double* SUM = (double*)malloc(sizeof(double) * (size_t)w * 3);
    if (!indep_box) {
        const size_t r0 = (size_t)reflect101(-1, h) * w;
        for (size_t i = 0; i < (size_t)w * 3; i++) SUM[i] = (0.0 + rsum[r0 * 3 + i]) + rsum[i];
    }
    for (int y = 0; y < h; y++) {
        const size_t ym = (size_t)reflect101(y - 1, h) * w, yc = (size_t)y * w, yp = (size_t)reflect101(y + 1, h) * w;
        for (int x = 0; x < w; x++) {
            float box[3];
            for (int c = 0; c < 3; c++) {
                if (indep_box) {
                    box[c] = (float)((rsum[(ym + x) * 3 + c] + rsum[(yc + x) * 3 + c]) + rsum[(yp + x) * 3 + c]);
                } else {
                    const double s0 = SUM[(size_t)x * 3 + c] + rsum[(yp + x) * 3 + c];
                    box[c] = (float)s0;
                    SUM[(size_t)x * 3 + c] = s0 - rsum[(ym + x) * 3 + c];
                }
            }
            const float a = box[0] * 0.5f, b = box[1], c = box[2] * 0.5f;
            const float t = a - c;
            const float tt = t * t;
            const float bb = b * b;
            eig[(size_t)y * w + x] = (a + c) - sqrtf(tt + bb);
        }
    }
    free(rx); free(rs); free(cov); free(rsum); free(SUM);
}

/* ---- A.3: pyramidal Lucas-Kanade, one point at a time ------------------------- */

typedef struct {
    const uint8_t* img;     /* unpadded level */
    const int16_t* deriv;   /* unpadded Scharr derivs (only needed for image 1) */
    int w, h;
} orc_level;

static inline int pixI(const orc_level* L, int x, int y) {
    /* level padded by winSize with REFLECT_101 */
    return L->img[(size_t)reflect101(y, L->h) * L->w + reflect101(x, L->w)];
}
static inline int derI(const orc_level* L, int x, int y, int c) {
    /* derivative image padded with zeros */
    if ((unsigned)x >= (unsigned)L->w || (unsigned)y >= (unsigned)L->h) return 0;
    return L->deriv[((size_t)y * L->w + x) * 2 + c];
}
static inline int cv_round(float v) { return (int)lrintf(v); } /* round-half-even */
static inline int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

/* Track n points through `nlevels` levels (index 0 = full resolution).
 * win = window size (square), iters/eps = TermCriteria, min_eig = minEigThreshold.
 * Outputs next (x,y), status, err exactly as cv::calcOpticalFlowPyrLK lays them out. */
/* Optional trace (scheduling studies of the CUDA kernel, scripts/lk_sched_sim.py): when set, element
 * [pi * nlevels + level] receives the number of window passes the point ran at that level. */
int* orc_lk_iter_trace = 0;
void orc_lk_set_iter_trace(int* p) { orc_lk_iter_trace = p; }

void orc_lk(const orc_level* L1, const orc_level* L2, int nlevels, const float* pts, int n,
            int win, int iters, double eps, double min_eig_thr, float* next, uint8_t* status, float* err) {
    const int W_BITS = 14;
    const float FLT_SCALE = 1.f / (1 << 20);
    const float halfw = (win - 1) * 0.5f;
    int16_t* Iw = (int16_t*)malloc(sizeof(int16_t) * win * win);
    int16_t* dIw = (int16_t*)malloc(sizeof(int16_t) * win * win * 2);
    if (iters < 0) iters = 0; if (iters > 100) iters = 100;
    if (eps < 0) eps = 0; if (eps > 10) eps = 10;
    const double eps2 = eps * eps;
    for (int i = 0; i < n; i++) { status[i] = 1; err[i] = 0; }
    for (int level = nlevels - 1; level >= 0; level--) {
        const orc_level* A = &L1[level];
        const orc_level* B = &L2[level];
        for (int pi = 0; pi < n; pi++) {
            float px = pts[2 * pi] * (float)(1. / (1 << level));
            float py = pts[2 * pi + 1] * (float)(1. / (1 << level));
            float nx, ny;
            if (level == nlevels - 1) { nx = px; ny = py; }
            else { nx = next[2 * pi] * 2.f; ny = next[2 * pi + 1] * 2.f; }
            next[2 * pi] = nx; next[2 * pi + 1] = ny;
            px -= halfw; py -= halfw;
            int ipx = (int)floorf(px), ipy = (int)floorf(py);
            if (ipx < -win || ipx >= A->w || ipy < -win || ipy >= A->h) {
                if (level == 0) { status[pi] = 0; err[pi] = 0; }
                continue;
            }
            float a = px - ipx, b = py - ipy;
            int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
            int iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
            int iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
            int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            /* Accumulation order of OpenCV's CV_SIMD128 path (lkpyramid.cpp): columns
             * 0..nvec-1 go, 8 at a time, into four float lanes (lane k takes column k then
             * column k+4 of each group, products rounded then added); the remaining columns
             * are added sequentially into a scalar; total = scalar + ((q0+q2)+(q1+q3)). */
            const int nvec = (win / 8) * 8;
            float iA11 = 0, iA12 = 0, iA22 = 0;
            float q11[4] = {0, 0, 0, 0}, q12[4] = {0, 0, 0, 0}, q22[4] = {0, 0, 0, 0};
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const int X = ipx + x, Y = ipy + y;
                    int ival = descale(pixI(A, X, Y) * iw00 + pixI(A, X + 1, Y) * iw01 +
                                       pixI(A, X, Y + 1) * iw10 + pixI(A, X + 1, Y + 1) * iw11, W_BITS - 5);
                    int ixval = descale(derI(A, X, Y, 0) * iw00 + derI(A, X + 1, Y, 0) * iw01 +
                                        derI(A, X, Y + 1, 0) * iw10 + derI(A, X + 1, Y + 1, 0) * iw11, W_BITS);
                    int iyval = descale(derI(A, X, Y, 1) * iw00 + derI(A, X + 1, Y, 1) * iw01 +
                                        derI(A, X, Y + 1, 1) * iw10 + derI(A, X + 1, Y + 1, 1) * iw11, W_BITS);
                    Iw[y * win + x] = (int16_t)ival;
                    dIw[(y * win + x) * 2] = (int16_t)ixval;
                    dIw[(y * win + x) * 2 + 1] = (int16_t)iyval;
                    if (x < nvec) {
                        const int k = x & 3;
                        const float fx = (float)ixval, fy = (float)iyval;
                        q22[k] = fy * fy + q22[k];
                        q12[k] = fx * fy + q12[k];
                        q11[k] = fx * fx + q11[k];
                    } else {
                        iA11 += (float)(ixval * ixval);
                        iA12 += (float)(ixval * iyval);
                        iA22 += (float)(iyval * iyval);
                    }
                }
            iA11 += (q11[0] + q11[2]) + (q11[1] + q11[3]);
            iA12 += (q12[0] + q12[2]) + (q12[1] + q12[3]);
            iA22 += (q22[0] + q22[2]) + (q22[1] + q22[3]);
            float A11 = iA11 * FLT_SCALE, A12 = iA12 * FLT_SCALE, A22 = iA22 * FLT_SCALE;
            float D = A11 * A22 - A12 * A12;
            float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
            if (minEig < min_eig_thr || D < 1.1920929e-07f) {
                if (level == 0) status[pi] = 0;
                continue;
            }
            D = 1.f / D;
            nx -= halfw; ny -= halfw;
            float pdx = 0, pdy = 0;
            int j;
            for (j = 0; j < iters; j++) {
                int inx = (int)floorf(nx), iny = (int)floorf(ny);
                if (inx < -win || inx >= B->w || iny < -win || iny >= B->h) {
                    if (level == 0) status[pi] = 0;
                    break;
                }
                a = nx - inx; b = ny - iny;
                iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                /* SIMD128 order: per row and group of 8 columns, lane pairs (c, c+4) are
                 * summed exactly in int32 (pmaddwd), converted to float and accumulated into
                 * X[c]; tail columns go sequentially into a scalar; total = scalar +
                 * ((X0+X2)+(X1+X3)). */
                float ib1 = 0, ib2 = 0;
                float qx[4] = {0, 0, 0, 0}, qy[4] = {0, 0, 0, 0};
                for (int y = 0; y < win; y++) {
                    int dif[64];
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        dif[x] = descale(pixI(B, X, Y) * iw00 + pixI(B, X + 1, Y) * iw01 +
                                         pixI(B, X, Y + 1) * iw10 + pixI(B, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                 Iw[y * win + x];
                    }
                    int x = 0;
                    for (; x + 8 <= nvec; x += 8)
                        for (int c = 0; c < 4; c++) {
                            const int p0 = y * win + x + c, p1 = p0 + 4;
                            qx[c] += (float)(dif[x + c] * dIw[p0 * 2] + dif[x + c + 4] * dIw[p1 * 2]);
                            qy[c] += (float)(dif[x + c] * dIw[p0 * 2 + 1] + dif[x + c + 4] * dIw[p1 * 2 + 1]);
                        }
                    for (; x < win; x++) {
                        ib1 += (float)(dif[x] * dIw[(y * win + x) * 2]);
                        ib2 += (float)(dif[x] * dIw[(y * win + x) * 2 + 1]);
                    }
                }
                ib1 += (qx[0] + qx[2]) + (qx[1] + qx[3]);
                ib2 += (qy[0] + qy[2]) + (qy[1] + qy[3]);
                float b1 = ib1 * FLT_SCALE, b2 = ib2 * FLT_SCALE;
                float dx = (float)((A12 * b2 - A22 * b1) * D);
                float dy = (float)((A12 * b1 - A11 * b2) * D);
                nx += dx; ny += dy;
                next[2 * pi] = nx + halfw; next[2 * pi + 1] = ny + halfw;
                if ((double)dx * dx + (double)dy * dy <= eps2) break;
                if (j > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                    next[2 * pi] -= dx * 0.5f; next[2 * pi + 1] -= dy * 0.5f;
                    break;
                }
                pdx = dx; pdy = dy;
            }
            if (orc_lk_iter_trace) orc_lk_iter_trace[pi * nlevels + level] = j < iters ? j + 1 : iters;
            if (status[pi] && level == 0) {
                float fx = next[2 * pi] - halfw, fy = next[2 * pi + 1] - halfw;
                int inx = (int)floorf(fx), iny = (int)floorf(fy);
                if (inx < -win || inx >= B->w || iny < -win || iny >= B->h) {
                    status[pi] = 0;
                    continue;
                }
                a = fx - inx; b = fy - iny;
                iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
                iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
                iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
                iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
                float errval = 0.f;
                for (int y = 0; y < win; y++)
                    for (int x = 0; x < win; x++) {
                        const int X = inx + x, Y = iny + y;
                        int diff = descale(pixI(B, X, Y) * iw00 + pixI(B, X + 1, Y) * iw01 +
                                           pixI(B, X, Y + 1) * iw10 + pixI(B, X + 1, Y + 1) * iw11, W_BITS - 5) -
                                   Iw[y * win + x];
                        errval += (float)abs(diff);
                    }
                err[pi] = errval * 1.f / (32 * win * win);
            }
        }
    }
    free(Iw); free(dIw);
}

/* ---- The reference's own detector logic around the eig map ---------------------------
 * /root/reference/cpp/feature_detection/gftt.cc:38-192: per-grid-cell max and TOZERO
 * threshold (maxVal*quality formed in double, compared as float, value kept iff > thresh),
 * 3x3 dilate + equality on interior pixels, sort by (value desc, address desc), greedy
 * min-distance suppression on a bucket grid with cell = cvRound(min_distance), early exit
 * at max_corners.  `eig` is modified in place like the reference does.  Returns the number
 * of corners written to out_xy (x,y float pairs), or -1 if cap is too small. */
typedef struct { float v; int addr; } orc_cand;
static int cand_cmp(const void* a, const void* b) {
    const orc_cand* p = (const orc_cand*)a; const orc_cand* q = (const orc_cand*)b;
    if (p->v > q->v) return -1;
    if (p->v < q->v) return 1;
    return (p->addr > q->addr) ? -1 : (p->addr < q->addr) ? 1 : 0;
}
int orc_gftt_select(float* eig, int w, int h, double quality, double min_distance, int max_corners,
                    int grid_rows, int grid_cols, float* out_xy, int cap) {
    if (grid_rows < 1) grid_rows = 1;
    if (grid_cols < 1) grid_cols = 1;
    const int bh = (h + grid_rows - 1) / grid_rows, bw = (w + grid_cols - 1) / grid_cols;
    for (int gy = 0; gy < grid_rows; gy++)
        for (int gx = 0; gx < grid_cols; gx++) {
            const int y0 = gy * bh, x0 = gx * bw;
            const int y1 = y0 + bh < h ? y0 + bh : h, x1 = x0 + bw < w ? x0 + bw : w;
            if (y1 <= y0 || x1 <= x0) continue;
            double maxVal = eig[(size_t)y0 * w + x0];
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++)
                    if (eig[(size_t)y * w + x] > maxVal) maxVal = eig[(size_t)y * w + x];
            const float thr = (float)(maxVal * quality);
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++) {
                    float* p = &eig[(size_t)y * w + x];
                    *p = (*p > thr) ? *p : 0.f;
                }
        }
    orc_cand* cands = (orc_cand*)malloc(sizeof(orc_cand) * (size_t)w * h);
    size_t total = 0;
    for (int y = 1; y < h - 1; y++)
        for (int x = 1; x < w - 1; x++) {
            const float v = eig[(size_t)y * w + x];
            if (v == 0.f) continue;
            float m = v;
            for (int dy = -1; dy <= 1; dy++)
                for (int dx = -1; dx <= 1; dx++) {
                    const float n = eig[(size_t)(y + dy) * w + x + dx];
                    if (n > m) m = n;
                }
            if (v == m) { cands[total].v = v; cands[total].addr = y * w + x; total++; }
        }
    qsort(cands, total, sizeof(orc_cand), cand_cmp);
    int n = 0;
    if (min_distance >= 1) {
        const int cell = (int)lrint(min_distance);
        const int gw = (w + cell - 1) / cell, gh = (h + cell - 1) / cell;
        /* bucket grid: singly linked lists of kept corners per cell */
        int* head = (int*)malloc(sizeof(int) * (size_t)gw * gh);
        int* nextp = (int*)malloc(sizeof(int) * (total ? total : 1));
        int* kx = (int*)malloc(sizeof(int) * (total ? total : 1));
        int* ky = (int*)malloc(sizeof(int) * (total ? total : 1));
        for (size_t i = 0; i < (size_t)gw * gh; i++) head[i] = -1;
        const double md2 = min_distance * min_distance;
        for (size_t i = 0; i < total; i++) {
            const int y = cands[i].addr / w, x = cands[i].addr - y * w;
            const int xc = x / cell, yc = y / cell;
            const int x1 = xc - 1 > 0 ? xc - 1 : 0, y1 = yc - 1 > 0 ? yc - 1 : 0;
            const int x2 = xc + 1 < gw - 1 ? xc + 1 : gw - 1, y2 = yc + 1 < gh - 1 ? yc + 1 : gh - 1;
            int good = 1;
            for (int yy = y1; yy <= y2 && good; yy++)
                for (int xx = x1; xx <= x2 && good; xx++)
                    for (int j = head[yy * gw + xx]; j >= 0; j = nextp[j]) {
                        const float dx = (float)x - (float)kx[j], dy = (float)y - (float)ky[j];
                        if ((double)(dx * dx + dy * dy) < md2) { good = 0; break; }
                    }
            if (good) {
                if (n >= cap) { n = -1; break; }
                kx[n] = x; ky[n] = y;
                nextp[n] = head[yc * gw + xc];
                head[yc * gw + xc] = n;
                out_xy[2 * n] = (float)x; out_xy[2 * n + 1] = (float)y;
                n++;
                if (max_corners > 0 && n == max_corners) break;
            }
        }
        free(head); free(nextp); free(kx); free(ky);
    } else {
        for (size_t i = 0; i < total; i++) {
            if (n >= cap) { n = -1; break; }
            const int y = cands[i].addr / w, x = cands[i].addr - y * w;
            out_xy[2 * n] = (float)x; out_xy[2 * n + 1] = (float)y;
            n++;
            if (max_corners > 0 && n == max_corners) break;
        }
    }
    free(cands);
    return n;
}
