/*
 * TEST INFRASTRUCTURE ONLY - see yolo2_oracle.h.  Plain-C restatement of the reference's
 * YOLOv2 accelerator datapath (parity PINNED against oracle/_ref, tests/test_oracle_vs_ref.py).
 * Citations are relative to /root/reference.
 */
#include "yolo2_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MIN(a, b) ((a) < (b) ? (a) : (b))

int orc_align8(int w) { return (w + 7) & ~7; } /* yolo2_accel.cpp:89-94 */

/* core_compute.cpp:49-62 (shift magnitude clamped to 30), :86-94, :108-113. */
int64_t orc_round_shift(int64_t v, int shift)
{
    if (shift > 0) {
        int mag = shift > 30 ? 30 : shift;
        return (v + ((int64_t)1 << (mag - 1))) >> mag;
    }
    if (shift < 0) {
        int mag = -shift > 30 ? 30 : -shift;
        return (int64_t)((uint64_t)v << mag); /* v << mag without signed-shift UB */
    }
    return v;
}

static inline int64_t clamp16(int64_t v)
{
    if (v > 32767) return 32767;
    if (v < -32768) return -32768;
    return v;
}

/* Element offset of w[m][c][tap] inside one layer of the reorganised blob
 * (producer yolov2_weight_gen.cpp:43-66, consumer core_io.cpp:154-198). */
static inline size_t reorg_woff(int m, int c, int tap, int ifm, int ofm, int k2, int TM, int TN)
{
    int m0 = (m / TM) * TM, n0 = (c / TN) * TN;
    int tmm = ORC_MIN(TM, ofm - m0), tnn = ORC_MIN(TN, ifm - n0);
    return (size_t)m0 * ifm * k2 + (size_t)tmm * n0 * k2 + ((size_t)tap * tmm + (m - m0)) * tnn + (c - n0);
}

/* Tile parameters of the reference BUILD (scripts/hw_params_gen.py --tn/--tm -> hls/core/params.hpp): Tn is the rounding
 * group of the int16 accumulator (core_scheduler.cpp:45), Tm the weight block.  Used by the net-level drivers below. */
static int g_tn = 4, g_tm = 32;
void orc_set_tile_params(int tn, int tm) { g_tn = tn > 0 ? tn : 4; g_tm = tm > 0 ? tm : 32; }

int orc_conv_i16(const int16_t *in, int16_t *out, const int16_t *w_reorg, const int16_t *bias,
                 int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                 int pad, int is_nl, int TM, int TN, int qw, int qa_in, int qa_out, int qb)
{
    if (ifm <= 0 || ofm <= 0 || ksize <= 0 || kstride <= 0 || TM <= 0 || TN <= 0) return -1;
    const int iwa = orc_align8(iw), owa = orc_align8(ow);
    const int k2 = ksize * ksize;
    const int shift_out = qa_in + qw - qa_out;  /* core_compute.cpp:49 */
    const int shift_bias = qb - qa_out;         /* core_compute.cpp:50 */
    const int groups = (ifm + TN - 1) / TN;     /* core_scheduler.cpp:45 */
    const size_t opix = (size_t)oh * ow;
    int err = 0;

#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m < ofm; ++m) {
        int64_t *acc = (int64_t *)malloc(sizeof(int64_t) * opix);
        if (!acc) { err = -2; continue; }
        const int64_t base = orc_round_shift((int64_t)bias[m], shift_bias); /* not saturated: :86-94 */
        for (size_t p = 0; p < opix; ++p) acc[p] = base;
        for (int g = 0; g < groups; ++g) {
            const int n0 = g * TN, tnn = ORC_MIN(TN, ifm - n0);
            for (int i = 0; i < ksize; ++i)
                for (int j = 0; j < ksize; ++j) {
                    int32_t wv[64];
                    for (int t = 0; t < tnn && t < 64; ++t)
                        wv[t] = w_reorg[reorg_woff(m, n0 + t, i * ksize + j, ifm, ofm, k2, TM, TN)];
                    const int first = (g == 0 && i == 0 && j == 0);
                    for (int y = 0; y < oh; ++y) {
                        const int iy = y * kstride + i - pad;
                        const int row_ok = (iy >= 0 && iy < ih);
                        int64_t *arow = acc + (size_t)y * ow;
                        if (!row_ok && !first) continue; /* P = 0 and acc already in range */
                        for (int x = 0; x < ow; ++x) {
                            const int ix = x * kstride + j - pad;
                            int64_t P = 0;
                            if (row_ok && ix >= 0 && ix < iw) { /* core_io.cpp:53-70 */
                                const int16_t *px = in + ((size_t)n0 * ih + iy) * iwa + ix;
                                for (int t = 0; t < tnn; ++t) /* core_compute.cpp:100-106 */
                                    P += (int64_t)(wv[t] * (int32_t)px[(size_t)t * ih * iwa]);
                            }
                            arow[x] = clamp16(arow[x] + orc_round_shift(P, shift_out)); /* :108-118 */
                        }
                    }
                }
        }
        int16_t *orow = out + (size_t)m * oh * owa;
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                int32_t v = (int32_t)acc[(size_t)y * ow + x];
                if (is_nl && v < 0) v = v / 10; /* core_compute.cpp:193-198, C division */
                orow[(size_t)y * owa + x] = (int16_t)clamp16(v);
            }
        free(acc);
    }
    return err;
}

int orc_conv_f32(const float *in, float *out, const float *w_reorg, const float *bias,
                 int ifm, int ofm, int ksize, int kstride, int iw, int ih, int ow, int oh,
                 int pad, int is_nl, int TM, int TN)
{
    if (ifm <= 0 || ofm <= 0 || ksize <= 0 || kstride <= 0 || TM <= 0 || TN <= 0) return -1;
    const int iwa = orc_align8(iw), owa = orc_align8(ow);
    const int k2 = ksize * ksize;
    const int groups = (ifm + TN - 1) / TN;
    const size_t opix = (size_t)oh * ow;
    int err = 0;

#pragma omp parallel for schedule(dynamic, 1)
    for (int m = 0; m < ofm; ++m) {
        float *acc = (float *)malloc(sizeof(float) * opix);
        if (!acc) { err = -2; continue; }
        for (size_t p = 0; p < opix; ++p) acc[p] = bias[m]; /* core_compute.cpp:151-152 */
        for (int g = 0; g < groups; ++g) {
            const int n0 = g * TN, tnn = ORC_MIN(TN, ifm - n0);
            for (int i = 0; i < ksize; ++i)
                for (int j = 0; j < ksize; ++j) {
                    float wv[64];
                    for (int t = 0; t < tnn && t < 64; ++t)
                        wv[t] = w_reorg[reorg_woff(m, n0 + t, i * ksize + j, ifm, ofm, k2, TM, TN)];
                    for (int y = 0; y < oh; ++y) {
                        const int iy = y * kstride + i - pad;
                        if (iy < 0 || iy >= ih) continue;
                        float *arow = acc + (size_t)y * ow;
                        for (int x = 0; x < ow; ++x) {
                            const int ix = x * kstride + j - pad;
                            if (ix < 0 || ix >= iw) continue;
                            const float *px = in + ((size_t)n0 * ih + iy) * iwa + ix;
                            float ps = 0.0f; /* core_compute.cpp:159-168 */
                            for (int t = 0; t < tnn; ++t) {
                                float mul = wv[t] * px[(size_t)t * ih * iwa];
                                ps += mul;
                            }
                            arow[x] = arow[x] + ps;
                        }
                    }
                }
        }
        float *orow = out + (size_t)m * oh * owa;
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                float v = acc[(size_t)y * ow + x];
                if (v < 0.0f && is_nl) v = v * 0.1f; /* core_compute.cpp:200-204 */
                orow[(size_t)y * owa + x] = v;
            }
        free(acc);
    }
    return err;
}

int orc_maxpool_i16(const int16_t *in, int16_t *out, int ch, int ksize, int kstride,
                    int iw, int ih, int ow, int oh)
{
    if (ksize != 2) return -1; /* the store is hard-wired to i==1&&j==1, core_compute.cpp:299-300 */
    const int iwa = orc_align8(iw), owa = orc_align8(ow);
#pragma omp parallel for
    for (int c = 0; c < ch; ++c)
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                int16_t best = -32768; /* core_compute.cpp:289-291 */
                for (int i = 0; i < ksize; ++i)
                    for (int j = 0; j < ksize; ++j) {
                        int iy = y * kstride + i, ix = x * kstride + j; /* Padding forced 0 */
                        int16_t v = (iy < ih && ix < iw) ? in[((size_t)c * ih + iy) * iwa + ix]
                                                         : (int16_t)-32768; /* core_io.cpp:96-99 */
                        if (v > best) best = v;
                    }
                out[((size_t)c * oh + y) * owa + x] = best;
            }
    return 0;
}

int orc_maxpool_f32(const float *in, float *out, int ch, int ksize, int kstride,
                    int iw, int ih, int ow, int oh)
{
    if (ksize != 2) return -1;
    const int iwa = orc_align8(iw), owa = orc_align8(ow);
#pragma omp parallel for
    for (int c = 0; c < ch; ++c)
        for (int y = 0; y < oh; ++y)
            for (int x = 0; x < ow; ++x) {
                float best = -1024 * 1024; /* core_compute.cpp:292-293 */
                for (int i = 0; i < ksize; ++i)
                    for (int j = 0; j < ksize; ++j) {
                        int iy = y * kstride + i, ix = x * kstride + j;
                        float v = (iy < ih && ix < iw) ? in[((size_t)c * ih + iy) * iwa + ix]
                                                       : (float)(-1024 * 1024); /* core_io.cpp:100-102 */
                        if (v > best) best = v;
                    }
                out[((size_t)c * oh + y) * owa + x] = best;
            }
    return 0;
}

int orc_reorg_hls_i16(const int16_t *in, int16_t *out, int ch, int TM, int iw, int ih, int ow, int oh)
{
    /* core_scheduler.cpp:88-112 loads channels m.. as tile channel 0.., reorg_yolo2 reads tile
     * channel 0 only and write_back stores TM_MIN output channels at m (yolo2_accel.cpp:127-169). */
    if (TM <= 0) return -1;
    const int iwa = orc_align8(iw), owa = orc_align8(ow);
    for (int m = 0; m < ch; m += TM) {
        int tmm = ORC_MIN(TM, ch - m);
        for (int q = 0; q < tmm && q < 4; ++q) {
            int ky = q >> 1, kx = q & 1;
            for (int y = 0; y < oh; ++y)
                for (int x = 0; x < ow; ++x) {
                    int iy = 2 * y + ky, ix = 2 * x + kx;
                    int16_t v = (iy < ih && ix < iw) ? in[((size_t)m * ih + iy) * iwa + ix] : 0;
                    out[((size_t)(m + q) * oh + y) * owa + x] = v;
                }
        }
    }
    return 0;
}

void orc_weight_reorg(const void *w, void *w_reorg, int ifm, int ofm, int ksize, int Tm, int Tn, int elem)
{
    /* yolov2_weight_gen.cpp:43-66 : per (m-tile, n-tile) block, order [tap][tm][tn]. */
    const int k2 = ksize * ksize;
    const unsigned char *src = (const unsigned char *)w;
    unsigned char *dst = (unsigned char *)w_reorg;
    size_t off = 0;
    for (int m = 0; m < ofm; m += Tm) {
        int tmm = ORC_MIN(Tm, ofm - m);
        for (int n = 0; n < ifm; n += Tn) {
            int tnn = ORC_MIN(Tn, ifm - n);
            for (int tk = 0; tk < k2; ++tk)
                for (int tm = 0; tm < tmm; ++tm)
                    for (int tn = 0; tn < tnn; ++tn) {
                        size_t s = ((size_t)(m + tm) * ifm + (n + tn)) * k2 + tk;
                        memcpy(dst + off * elem, src + s * elem, (size_t)elem);
                        ++off;
                    }
        }
    }
}

void orc_quantize_input(const float *in, int16_t *out, size_t count, int q_in)
{
    const float scale = ldexpf(1.0f, q_in); /* yolo2_model.cpp:262 */
    for (size_t i = 0; i < count; ++i) {
        float v = in[i] * scale;
        if (v > 32767.f) v = 32767.f;
        if (v < -32768.f) v = -32768.f;
        long long q = llroundf(v); /* std::llround(float): half away from zero, :268 */
        if (q > 32767) q = 32767;
        if (q < -32768) q = -32768;
        out[i] = (int16_t)q;
    }
}

/* Flat-memory reorg of yolo2_model.cpp:112-129 called as reorg_cpu(x, W, H*C/4, 4, 2, out) (:373). */
static void reorg_flat_index(int c, int h, int w, size_t f, size_t *src)
{
    const size_t hc = (size_t)h * c / 4; /* 416 in the reference */
    size_t i = f % w, jj = f / w, j = jj % hc, k = jj / hc;
    *src = (2 * i + (k & 1)) + 2 * (size_t)w * (2 * j + (k >> 1));
}

void orc_reorg_driver_i16(const int16_t *in, int16_t *out, int c, int h, int w, int shift)
{
    const int wa = orc_align8(w), ow = w / 2, oh = h / 2, oc = 4 * c, owa = orc_align8(ow);
    const size_t total = (size_t)c * h * w;
    memset(out, 0, sizeof(int16_t) * (size_t)oc * oh * owa); /* :375 */
    for (size_t f = 0; f < total; ++f) {
        size_t s;
        reorg_flat_index(c, h, w, f, &s);
        /* compact source index s -> padded source address (:371-372) */
        size_t srow = s / w, scol = s % w;
        int32_t v = in[srow * wa + scol];
        if (shift > 0) v >>= shift; /* :386-391, arithmetic, no rounding */
        v = (int32_t)clamp16(v);
        /* compact destination index f -> padded destination address (:376-377) */
        size_t drow = f / ow, dcol = f % ow;
        out[drow * owa + dcol] = (int16_t)v;
    }
}

void orc_reorg_driver_f32(const float *in, float *out, int c, int h, int w)
{
    const int wa = orc_align8(w), ow = w / 2, oh = h / 2, oc = 4 * c, owa = orc_align8(ow);
    const size_t total = (size_t)c * h * w;
    memset(out, 0, sizeof(float) * (size_t)oc * oh * owa);
    for (size_t f = 0; f < total; ++f) {
        size_t s;
        reorg_flat_index(c, h, w, f, &s);
        size_t srow = s / w, scol = s % w, drow = f / ow, dcol = f % ow;
        out[drow * owa + dcol] = in[srow * wa + scol];
    }
}

void orc_region_strip_dequant_i16(const int16_t *in, float *out, int ch, int h, int w, int q)
{
    const int wa = orc_align8(w);
    const float scale = ldexpf(1.0f, -q); /* yolo2_model.cpp:417 */
    for (size_t r = 0; r < (size_t)ch * h; ++r)
        for (int x = 0; x < w; ++x) out[r * w + x] = (float)in[r * wa + x] * scale;
}

void orc_region_strip_f32(const float *in, float *out, int ch, int h, int w)
{
    const int wa = orc_align8(w);
    for (size_t r = 0; r < (size_t)ch * h; ++r)
        for (int x = 0; x < w; ++x) out[r * w + x] = in[r * wa + x];
}

static inline float logistic_f(float x) { return (float)(1. / (1. + exp((double)(-x)))); } /* yolo_math.cpp:19 */

void orc_region_forward(const float *in_f, float *out, int w, int h, int n, int classes,
                        int coords, int softmax, int background)
{
    const int wh = w * h, per = coords + 1 + classes;
    memcpy(out, in_f, sizeof(float) * (size_t)n * per * wh); /* yolo_region.cpp:125 */
    for (int a = 0; a < n; ++a) {
        float *base = out + (size_t)a * per * wh;
        for (int t = 0; t < 2 * wh; ++t) base[t] = logistic_f(base[t]);               /* :129-130 */
        if (!background)
            for (int t = 0; t < wh; ++t) base[coords * wh + t] = logistic_f(base[coords * wh + t]); /* :131-132 */
    }
    if (softmax) { /* :136-139 -> yolo_math.cpp:226-250, temp = 1, stride = w*h */
        const int nc = classes + background;
        const int first = coords + !background;
        for (int a = 0; a < n; ++a)
            for (int loc = 0; loc < wh; ++loc) {
                const float *ip = in_f + ((size_t)a * per + first) * wh + loc;
                float *op = out + ((size_t)a * per + first) * wh + loc;
                float sum = 0, largest = -FLT_MAX;
                for (int i = 0; i < nc; ++i)
                    if (ip[(size_t)i * wh] > largest) largest = ip[(size_t)i * wh];
                for (int i = 0; i < nc; ++i) {
                    float arg = ip[(size_t)i * wh] / 1.0f - largest / 1.0f;
                    float e = (float)exp((double)arg);
                    sum += e;
                    op[(size_t)i * wh] = e;
                }
                for (int i = 0; i < nc; ++i) op[(size_t)i * wh] /= sum;
            }
    }
}

typedef struct { float x, y, w, h; } orc_box;
typedef struct { orc_box bbox; int classes; float *prob; float *mask; float objectness; int sort_class; } orc_det;

static int nms_cmp(const void *pa, const void *pb) /* yolo_post.cpp:7-20 */
{
    orc_det a = *(const orc_det *)pa, b = *(const orc_det *)pb;
    float diff;
    if (b.sort_class >= 0) diff = a.prob[b.sort_class] - b.prob[b.sort_class];
    else diff = a.objectness - b.objectness;
    if (diff < 0) return 1;
    else if (diff > 0) return -1;
    return 0;
}

static float overlap1(float x1, float w1, float x2, float w2) /* yolo_post.cpp:22-31 */
{
    float l1 = x1 - w1 / 2, l2 = x2 - w2 / 2;
    float left = l1 > l2 ? l1 : l2;
    float r1 = x1 + w1 / 2, r2 = x2 + w2 / 2;
    float right = r1 < r2 ? r1 : r2;
    return right - left;
}

static float box_iou1(orc_box a, orc_box b) /* yolo_post.cpp:33-52 */
{
    float w = overlap1(a.x, a.w, b.x, b.w), h = overlap1(a.y, a.h, b.y, b.h);
    float inter = (w < 0 || h < 0) ? 0 : w * h;
    float uni = a.w * a.h + b.w * b.h - inter;
    return inter / uni;
}

int orc_region_boxes_nms(const float *region, int lw, int lh, int n, int classes,
                         const float *anchors, int im_w, int im_h, int net_w, int net_h,
                         float thresh, float nms, float *boxes, float *probs, float *objectness)
{
    const int wh = lw * lh, total = wh * n, per = 5 + classes, coords = 4;
    orc_det *dets = (orc_det *)calloc((size_t)total, sizeof(orc_det));
    float *pstore = (float *)calloc((size_t)total * classes, sizeof(float));
    for (int i = 0; i < total; ++i) { dets[i].prob = pstore + (size_t)i * classes; dets[i].classes = classes; }

    int count = 0; /* yolo_region.cpp:169-193: compacting scan, cell-major then anchor */
    for (int i = 0; i < wh; ++i) {
        int row = i / lw, col = i % lw;
        for (int a = 0; a < n; ++a) {
            const float *cell = region + (size_t)a * per * wh + i;
            float obj = cell[(size_t)coords * wh];
            if (obj <= thresh) continue;
            orc_box b; /* get_region_box, yolo_region.cpp:18-26 */
            b.x = (col + cell[0]) / lw;
            b.y = (row + cell[(size_t)1 * wh]) / lh;
            b.w = expf(cell[(size_t)2 * wh]) * anchors[2 * a] / lw;
            b.h = expf(cell[(size_t)3 * wh]) * anchors[2 * a + 1] / lh;
            dets[count].bbox = b;
            dets[count].objectness = obj;
            for (int j = 0; j < classes; ++j) {
                float p = obj * cell[(size_t)(coords + 1 + j) * wh];
                dets[count].prob[j] = (p > thresh) ? p : 0;
            }
            ++count;
        }
    }
    { /* correct_region_boxes(dets, count, ..., relative=1), yolo_region.cpp:28-53 */
        int new_w, new_h;
        if (((float)net_w / im_w) < ((float)net_h / im_h)) { new_w = net_w; new_h = (im_h * net_w) / im_w; }
        else { new_h = net_h; new_w = (im_w * net_h) / im_h; }
        for (int i = 0; i < count; ++i) {
            orc_box b = dets[i].bbox;
            b.x = (b.x - (net_w - new_w) / 2. / net_w) / ((float)new_w / net_w);
            b.y = (b.y - (net_h - new_h) / 2. / net_h) / ((float)new_h / net_h);
            b.w *= (float)net_w / new_w;
            b.h *= (float)net_h / new_h;
            dets[i].bbox = b;
        }
    }
    if (nms > 0.0f) { /* do_nms_sort over ALL w*h*n entries, yolo_post.cpp:54-85 */
        int k = total - 1;
        for (int i = 0; i <= k; ++i)
            if (dets[i].objectness == 0) {
                orc_det sw = dets[i]; dets[i] = dets[k]; dets[k] = sw;
                --k; --i;
            }
        int live = k + 1;
        for (int c = 0; c < classes; ++c) {
            for (int i = 0; i < live; ++i) dets[i].sort_class = c;
            qsort(dets, (size_t)live, sizeof(orc_det), nms_cmp);
            for (int i = 0; i < live; ++i) {
                if (dets[i].prob[c] == 0) continue;
                orc_box a = dets[i].bbox;
                for (int j = i + 1; j < live; ++j)
                    if (box_iou1(a, dets[j].bbox) > nms) dets[j].prob[c] = 0;
            }
        }
    }
    for (int i = 0; i < total; ++i) {
        boxes[4 * i + 0] = dets[i].bbox.x; boxes[4 * i + 1] = dets[i].bbox.y;
        boxes[4 * i + 2] = dets[i].bbox.w; boxes[4 * i + 3] = dets[i].bbox.h;
        objectness[i] = dets[i].objectness;
        memcpy(probs + (size_t)i * classes, dets[i].prob, sizeof(float) * classes);
    }
    free(pstore);
    free(dets);
    return total;
}

/* ---- generalised driver (yolo2_model.cpp:229-446) -------------------------------------- */

static size_t fm_elems(int c, int h, int w) { return (size_t)c * h * orc_align8(w); }

/* Index of the conv whose Qa is remembered for the concat (reference: `i == 24`, :332-334):
 * the non-reorg input of the first multi-input route that follows a reorg. */
static int find_skip_layer(const orc_layer *L, int n)
{
    for (int i = 0; i < n; ++i)
        if (L[i].type == ORC_ROUTE && L[i].n_inputs >= 2)
            for (int a = 0; a < L[i].n_inputs; ++a)
                if (L[L[i].inputs[a]].type == ORC_REORG)
                    for (int b = 0; b < L[i].n_inputs; ++b)
                        if (b != a) return L[i].inputs[b];
    return -1;
}

#define NET_FORWARD_BODY(T, IS_I16)                                                                   \
    T **buf = (T **)calloc((size_t)n_layers, sizeof(T *));                                            \
    int rc = 0, conv_idx = 0;                                                                         \
    size_t woff = 0, boff = 0;                                                                        \
    const orc_layer *L0 = &layers[0];                                                                 \
    T *input = (T *)calloc(fm_elems(L0->c, L0->h, L0->w) + 64, sizeof(T));

int orc_net_forward_i16(const orc_layer *layers, int n_layers, const float *frame,
                        const int16_t *w_reorg, const int16_t *bias,
                        const int32_t *weight_q, const int32_t *bias_q, const int32_t *act_q, int n_act_q,
                        int16_t **dump, float *region_out)
{
    NET_FORWARD_BODY(int16_t, 1)
    if (n_act_q <= 0) { free(buf); free(input); return -10; } /* :258-260 */
    int current_qa = act_q[0], route_q = 0, pending_route_q = -1;
    const int skip_layer = find_skip_layer(layers, n_layers);
    {   /* quantise into the aligned layout; the reference image width is already a multiple of 8 */
        const int wa = orc_align8(L0->w);
        int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)L0->c * L0->h * L0->w);
        orc_quantize_input(frame, tmp, (size_t)L0->c * L0->h * L0->w, act_q[0]);
        for (size_t r = 0; r < (size_t)L0->c * L0->h; ++r) memcpy(input + r * wa, tmp + r * L0->w, sizeof(int16_t) * L0->w);
        free(tmp);
    }
    for (int i = 0; i < n_layers && rc == 0; ++i) {
        const orc_layer *l = &layers[i];
        const int16_t *src = (i == 0) ? input : buf[i - 1];
        if (l->type != ORC_ROUTE && l->type != ORC_REGION)
            buf[i] = (int16_t *)calloc(fm_elems(l->out_c, l->out_h, l->out_w) + 64, sizeof(int16_t));
        switch (l->type) {
        case ORC_CONV: {
            int TM = ORC_MIN(l->n, g_tm), TN = ORC_MIN(l->c, g_tn); /* :307-308 */
            int qa_in = (conv_idx < n_act_q) ? act_q[conv_idx] : current_qa;         /* :314 */
            int qa_out = (conv_idx + 1 < n_act_q) ? act_q[conv_idx + 1] : qa_in;     /* :315 */
            if (pending_route_q >= 0) qa_in = pending_route_q;                       /* :318-320 */
            rc = orc_conv_i16(src, buf[i], w_reorg + woff, bias + boff, l->c, l->n, l->size, l->stride,
                              l->w, l->h, l->out_w, l->out_h, l->pad, l->leaky, TM, TN,
                              weight_q[conv_idx], qa_in, qa_out, bias_q[conv_idx]);
            woff += (size_t)l->c * l->n * l->size * l->size;
            boff += (size_t)l->n;
            current_qa = qa_out;
            if (i == skip_layer) route_q = current_qa; /* :332-334 */
            pending_route_q = -1;
            ++conv_idx;
            break;
        }
        case ORC_MAXPOOL:
            rc = orc_maxpool_i16(src, buf[i], l->c, l->size, l->stride, l->w, l->h, l->out_w, l->out_h);
            break;
        case ORC_REORG: {
            int shift = 0;
            if (route_q > 0) { /* :379-399 */
                int target = ORC_MIN(route_q, current_qa);
                shift = current_qa - target;
                if (shift != 0) current_qa = target;
                pending_route_q = current_qa;
            }
            orc_reorg_driver_i16(src, buf[i], l->c, l->h, l->w, shift);
            break;
        }
        case ORC_ROUTE: { /* no-op by arena placement in the reference (:404-405): concat in input order */
            size_t tot = 0;
            for (int a = 0; a < l->n_inputs; ++a) {
                const orc_layer *s = &layers[l->inputs[a]];
                tot += fm_elems(s->out_c, s->out_h, s->out_w);
            }
            buf[i] = (int16_t *)calloc(tot + 64, sizeof(int16_t));
            size_t o = 0;
            for (int a = 0; a < l->n_inputs; ++a) {
                const orc_layer *s = &layers[l->inputs[a]];
                size_t e = fm_elems(s->out_c, s->out_h, s->out_w);
                memcpy(buf[i] + o, buf[l->inputs[a]], e * sizeof(int16_t));
                o += e;
            }
            break;
        }
        case ORC_REGION: {
            const int ch = l->n * (l->coords + 1 + l->classes);
            float *rf = (float *)malloc(sizeof(float) * (size_t)ch * l->h * l->w);
            orc_region_strip_dequant_i16(src, rf, ch, l->h, l->w, current_qa);
            if (region_out) orc_region_forward(rf, region_out, l->w, l->h, l->n, l->classes, l->coords, l->softmax, l->background);
            free(rf);
            break;
        }
        default: rc = -20;
        }
        if (dump && dump[i] && buf[i]) {
            size_t e = (l->type == ORC_ROUTE) ? 0 : fm_elems(l->out_c, l->out_h, l->out_w);
            memcpy(dump[i], buf[i], e * sizeof(int16_t));
        }
    }
    for (int i = 0; i < n_layers; ++i) free(buf[i]);
    free(buf); free(input);
    return rc;
}

int orc_net_forward_f32(const orc_layer *layers, int n_layers, const float *frame,
                        const float *w_reorg, const float *bias, float **dump, float *region_out)
{
    NET_FORWARD_BODY(float, 0)
    {
        const int wa = orc_align8(L0->w);
        for (size_t r = 0; r < (size_t)L0->c * L0->h; ++r) memcpy(input + r * wa, frame + r * L0->w, sizeof(float) * L0->w);
    }
    (void)conv_idx;
    for (int i = 0; i < n_layers && rc == 0; ++i) {
        const orc_layer *l = &layers[i];
        const float *src = (i == 0) ? input : buf[i - 1];
        if (l->type != ORC_ROUTE && l->type != ORC_REGION)
            buf[i] = (float *)calloc(fm_elems(l->out_c, l->out_h, l->out_w) + 64, sizeof(float));
        switch (l->type) {
        case ORC_CONV: {
            int TM = ORC_MIN(l->n, g_tm), TN = ORC_MIN(l->c, g_tn);
            rc = orc_conv_f32(src, buf[i], w_reorg + woff, bias + boff, l->c, l->n, l->size, l->stride,
                              l->w, l->h, l->out_w, l->out_h, l->pad, l->leaky, TM, TN);
            woff += (size_t)l->c * l->n * l->size * l->size;
            boff += (size_t)l->n;
            break;
        }
        case ORC_MAXPOOL:
            rc = orc_maxpool_f32(src, buf[i], l->c, l->size, l->stride, l->w, l->h, l->out_w, l->out_h);
            break;
        case ORC_REORG:
            orc_reorg_driver_f32(src, buf[i], l->c, l->h, l->w);
            break;
        case ORC_ROUTE: {
            size_t tot = 0;
            for (int a = 0; a < l->n_inputs; ++a) {
                const orc_layer *s = &layers[l->inputs[a]];
                tot += fm_elems(s->out_c, s->out_h, s->out_w);
            }
            buf[i] = (float *)calloc(tot + 64, sizeof(float));
            size_t o = 0;
            for (int a = 0; a < l->n_inputs; ++a) {
                const orc_layer *s = &layers[l->inputs[a]];
                size_t e = fm_elems(s->out_c, s->out_h, s->out_w);
                memcpy(buf[i] + o, buf[l->inputs[a]], e * sizeof(float));
                o += e;
            }
            break;
        }
        case ORC_REGION: {
            const int ch = l->n * (l->coords + 1 + l->classes);
            float *rf = (float *)malloc(sizeof(float) * (size_t)ch * l->h * l->w);
            orc_region_strip_f32(src, rf, ch, l->h, l->w);
            if (region_out) orc_region_forward(rf, region_out, l->w, l->h, l->n, l->classes, l->coords, l->softmax, l->background);
            free(rf);
            break;
        }
        default: rc = -20;
        }
        if (dump && dump[i] && buf[i]) {
            size_t e = (l->type == ORC_ROUTE) ? 0 : fm_elems(l->out_c, l->out_h, l->out_w);
            memcpy(dump[i], buf[i], e * sizeof(float));
        }
    }
    for (int i = 0; i < n_layers; ++i) free(buf[i]);
    free(buf); free(input);
    return rc;
}

/* ---- image front-end ---------------------------------------------------------------------------
 * load_image_stb's u8 -> float conversion (src/core/yolo_image.cpp:178-187), resize_image (:84-127:
 * the two-pass bilinear, `part` along x first, then along y with set_pixel + add_pixel) and
 * letterbox_image / fill_image / embed_image (:129-165).  Plain float arithmetic, one rounding per
 * operation (built without -march like the reference, so no FMA contraction). */
int orc_letterbox_u8(const unsigned char *hwc, int iw, int ih, int ic, float *out, int net_w, int net_h)
{
    if (!hwc || !out || iw <= 0 || ih <= 0 || ic <= 0 || net_w <= 0 || net_h <= 0) return -1;
    int new_w, new_h;
    if (((float)net_w / iw) < ((float)net_h / ih)) { new_w = net_w; new_h = (ih * net_w) / iw; }
    else { new_h = net_h; new_w = (iw * net_h) / ih; }
    float *im = (float *)malloc(sizeof(float) * (size_t)iw * ih * ic);
    float *part = (float *)malloc(sizeof(float) * (size_t)new_w * ih * ic);
    float *res = (float *)malloc(sizeof(float) * (size_t)new_w * new_h * ic);
    if (!im || !part || !res) { free(im); free(part); free(res); return -2; }
    for (int k = 0; k < ic; ++k)
        for (int j = 0; j < ih; ++j)
            for (int i = 0; i < iw; ++i) im[i + iw * j + iw * ih * k] = (float)hwc[k + ic * i + ic * iw * j] / 255.;
    const float w_scale = (float)(iw - 1) / (new_w - 1);
    const float h_scale = (float)(ih - 1) / (new_h - 1);
    for (int k = 0; k < ic; ++k)
        for (int r = 0; r < ih; ++r)
            for (int c = 0; c < new_w; ++c) {
                float val;
                if (c == new_w - 1 || iw == 1) val = im[(iw - 1) + iw * r + iw * ih * k];
                else {
                    float sx = c * w_scale;
                    int ix = (int)sx;
                    float dx = sx - ix;
                    val = (1 - dx) * im[ix + iw * r + iw * ih * k] + dx * im[ix + 1 + iw * r + iw * ih * k];
                }
                part[c + new_w * r + new_w * ih * k] = val;
            }
    for (int k = 0; k < ic; ++k)
        for (int r = 0; r < new_h; ++r) {
            float sy = r * h_scale;
            int iy = (int)sy;
            float dy = sy - iy;
            for (int c = 0; c < new_w; ++c) res[c + new_w * r + new_w * new_h * k] = (1 - dy) * part[c + new_w * iy + new_w * ih * k];
            if (r == new_h - 1 || ih == 1) continue;
            for (int c = 0; c < new_w; ++c) res[c + new_w * r + new_w * new_h * k] += dy * part[c + new_w * (iy + 1) + new_w * ih * k];
        }
    for (size_t i = 0; i < (size_t)net_w * net_h * ic; ++i) out[i] = .5;
    const int ox = (net_w - new_w) / 2, oy = (net_h - new_h) / 2;
    for (int k = 0; k < ic; ++k)
        for (int y = 0; y < new_h; ++y)
            for (int x = 0; x < new_w; ++x) out[(ox + x) + net_w * (oy + y) + net_w * net_h * k] = res[x + new_w * y + new_w * new_h * k];
    free(im); free(part); free(res);
    return 0;
}

/* the host libm's exp on an array: what the reference's logistic_activate / softmax evaluate (src/core/yolo_math.cpp:19,234) */
void orc_libm_exp(const double *x, double *y, long n)
{
    for (long i = 0; i < n; ++i) y[i] = exp(x[i]);
}
