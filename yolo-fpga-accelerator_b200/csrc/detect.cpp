// Host-side detection tail of the YOLOv2 pipeline: candidate boxes from the region tensor,
// letterbox un-mapping and per-class greedy NMS.  Mirrors the observable behaviour of the
// reference's get_region_detections / correct_region_boxes (src/core/yolo_region.cpp:18-53,
// 169-195) and do_nms_sort (src/core/yolo_post.cpp:54-85): same float arithmetic (expf for box
// sizes, float IoU), same candidate scan order (cell-major, then anchor), same qsort on an order that
// is carried from class to class (ties between equal probabilities resolve as in the reference).
// Output: the candidates with objectness > thresh in scan order (the reference's list BEFORE its
// qsort re-orders it); the return value is their count.  Stays on the host
// because it is <0.1 % of the frame time and shares libm with the reference.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <numeric>
#include <vector>

#include "../../include/yolo2cuda.h"

namespace {

struct Box { float x, y, w, h; };

float overlap_1d(float c1, float w1, float c2, float w2)
{
    float l1 = c1 - w1 / 2, l2 = c2 - w2 / 2;
    float left = l1 > l2 ? l1 : l2;
    float r1 = c1 + w1 / 2, r2 = c2 + w2 / 2;
    float right = r1 < r2 ? r1 : r2;
    return right - left;
}

float iou(const Box &a, const Box &b)
{
    float w = overlap_1d(a.x, a.w, b.x, b.w), h = overlap_1d(a.y, a.h, b.y, b.h);
    float inter = (w < 0 || h < 0) ? 0 : w * h;
    float uni = a.w * a.h + b.w * b.h - inter;
    return inter / uni;
}

struct SortKey { const float *pr; int classes, cls; };

// nms_comparator's outcomes (yolo_post.cpp:7-20): descending probability of the class being processed, 0 on ties
int by_class_prob_desc(const void *pa, const void *pb, void *ctx)
{
    const SortKey *k = static_cast<const SortKey *>(ctx);
    const float diff = k->pr[(size_t)(*static_cast<const int *>(pa)) * k->classes + k->cls] -
                       k->pr[(size_t)(*static_cast<const int *>(pb)) * k->classes + k->cls];
    return diff < 0 ? 1 : diff > 0 ? -1 : 0;
}

}  // namespace

extern "C" int yolo2cuda_region_detections(const float *region, int lw, int lh, int n, int classes, const float *anchors,
                                           int im_w, int im_h, int net_w, int net_h, float thresh, float nms,
                                           float *boxes, float *probs, float *objectness)
{
    if (!region || !anchors || !boxes || !probs || !objectness) return YOLO2CUDA_ERROR;
    if (lw <= 0 || lh <= 0 || n <= 0 || classes <= 0 || im_w <= 0 || im_h <= 0 || net_w <= 0 || net_h <= 0) return YOLO2CUDA_ERROR;
    const int wh = lw * lh, per = 5 + classes;
    std::vector<Box> bb;
    std::vector<float> obj;
    std::vector<float> pr;  // [count][classes]
    for (int cell = 0; cell < wh; ++cell) {
        const int row = cell / lw, col = cell % lw;
        for (int a = 0; a < n; ++a) {
            const float *e = region + (size_t)a * per * wh + cell;
            const float o = e[(size_t)4 * wh];
            if (o <= thresh) continue;
            Box b;
            b.x = (col + e[0]) / lw;
            b.y = (row + e[(size_t)wh]) / lh;
            b.w = std::exp(e[(size_t)2 * wh]) * anchors[2 * a] / lw;      // float exp, like the reference
            b.h = std::exp(e[(size_t)3 * wh]) * anchors[2 * a + 1] / lh;
            bb.push_back(b);
            obj.push_back(o);
            for (int j = 0; j < classes; ++j) {
                float p = o * e[(size_t)(5 + j) * wh];
                pr.push_back(p > thresh ? p : 0.0f);
            }
        }
    }
    const int count = (int)bb.size();
    // undo the letterbox (relative coordinates)
    int new_w, new_h;
    if (((float)net_w / im_w) < ((float)net_h / im_h)) { new_w = net_w; new_h = (im_h * net_w) / im_w; }
    else { new_h = net_h; new_w = (im_w * net_h) / im_h; }
    for (Box &b : bb) {
        b.x = (b.x - (net_w - new_w) / 2. / net_w) / ((float)new_w / net_w);
        b.y = (b.y - (net_h - new_h) / 2. / net_h) / ((float)new_h / net_h);
        b.w *= (float)net_w / new_w;
        b.h *= (float)net_h / new_h;
    }
    if (nms > 0.0f) {
        // do_nms_sort (yolo_post.cpp:54-85) sorts ONE array in place class after class: the order a class's qsort starts from is
        // the order the previous class left behind, and with equal probabilities that order decides which of two overlapping
        // boxes survives.  So the candidate order is carried across classes here too, and the sort is the C library's qsort
        // with the reference comparator's outcomes (yolo_post.cpp:7-20), i.e. the same permutation on the same libc.
        std::vector<int> order(count);
        std::iota(order.begin(), order.end(), 0);
        int live = count;
        for (int i = 0, k = count - 1; i <= k; ++i)          // entries with objectness == 0 go to the back (yolo_post.cpp:57-67)
            if (obj[order[i]] == 0) {
                std::swap(order[i], order[k]);
                --k; --i;
                live = k + 1;
            }
        SortKey key{pr.data(), classes, 0};
        for (int k = 0; k < classes; ++k) {
            key.cls = k;
            qsort_r(order.data(), (size_t)live, sizeof(int), by_class_prob_desc, &key);
            for (int i = 0; i < live; ++i) {
                const int di = order[i];
                if (pr[(size_t)di * classes + k] == 0) continue;
                for (int j = i + 1; j < live; ++j) {
                    const int dj = order[j];
                    if (iou(bb[di], bb[dj]) > nms) pr[(size_t)dj * classes + k] = 0;
                }
            }
        }
    }
    for (int i = 0; i < count; ++i) {
        boxes[4 * i + 0] = bb[i].x; boxes[4 * i + 1] = bb[i].y; boxes[4 * i + 2] = bb[i].w; boxes[4 * i + 3] = bb[i].h;
        objectness[i] = obj[i];
        std::copy(pr.begin() + (size_t)i * classes, pr.begin() + (size_t)(i + 1) * classes, probs + (size_t)i * classes);
    }
    return count;
}
