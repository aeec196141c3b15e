// `--backend cuda` for the reference host (see yolov2_cuda_ps.hpp).  Thin: translate the reference's
// `network` into the yolo2cuda_layer_desc table, load the reference's weight files exactly as
// load_weights does (hls/models/yolov2/yolo2_model.cpp:158-227, incl. the odd-length pad rule), call the
// C ABI.  No arithmetic happens here and there is no CPU fallback: errors throw std::runtime_error like
// the reference's loaders do (caught in main, src/models/yolov2/yolov2_main.cpp:340-346).
#include "yolov2_cuda_ps.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include <core/yolo.h>          // the reference's own header
#include <core/precision.hpp>

#include "../../include/yolo2cuda.h"

namespace {

template <typename T>
std::vector<T> read_binary(const std::string &path, bool required = true)
{
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp) {
        if (!required) return {};
        throw std::runtime_error("Failed to open file: " + path);
    }
    std::fseek(fp, 0, SEEK_END);
    long sz = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    if (sz < 0 || sz % sizeof(T) != 0) {
        std::fclose(fp);
        throw std::runtime_error("Invalid size for file: " + path);
    }
    std::vector<T> buf(sz / sizeof(T));
    size_t rd = std::fread(buf.data(), sizeof(T), buf.size(), fp);
    std::fclose(fp);
    if (rd != buf.size()) throw std::runtime_error("Short read: " + path);
    return buf;
}

std::vector<yolo2cuda_layer_desc> describe(const network *net)
{
    std::vector<yolo2cuda_layer_desc> d(net->n);
    for (int i = 0; i < net->n; ++i) {
        const layer &l = net->layers[i];
        yolo2cuda_layer_desc &o = d[i];
        std::memset(&o, 0, sizeof(o));
        o.c = l.c; o.h = l.h; o.w = l.w;
        o.out_c = l.out_c; o.out_h = l.out_h; o.out_w = l.out_w;
        o.size = l.size; o.stride = l.stride; o.pad = l.pad;
        switch (l.type) {
        case CONVOLUTIONAL:
            o.type = YOLO2CUDA_CONV;
            o.n = l.n;
            o.leaky = l.activation == LEAKY ? 1 : 0;   // yolo2_model.cpp:324
            o.batch_normalize = l.batch_normalize;
            break;
        case MAXPOOL:
            o.type = YOLO2CUDA_MAXPOOL;
            o.n = l.c;
            break;
        case REORG:
            o.type = YOLO2CUDA_REORG;
            break;
        case ROUTE:
            o.type = YOLO2CUDA_ROUTE;
            o.n_inputs = l.n;
            if (l.n > 4) throw std::runtime_error("route with more than 4 inputs is not on the accelerator path");
            for (int a = 0; a < l.n; ++a) o.inputs[a] = l.input_layers[a];
            o.c = o.out_c; o.h = o.out_h; o.w = o.out_w;
            break;
        case REGION:
            o.type = YOLO2CUDA_REGION;
            o.n = l.n; o.classes = l.classes; o.coords = l.coords; o.softmax = l.softmax; o.background = l.background;
            for (int a = 0; a < 2 * l.n && a < 32; ++a) o.anchors[a] = l.biases ? l.biases[a] : 0.5f;
            break;
        default:
            throw std::runtime_error("layer " + std::to_string(i) + ": type not on the accelerator path");
        }
    }
    return d;
}

struct Pack {
    std::vector<int16_t> w16, b16;
    std::vector<float> w32, b32;
    std::vector<int32_t> wq, bq, aq;
};

Pack load_pack(const std::vector<yolo2cuda_layer_desc> &d, Precision precision)
{
    Pack p;
    std::vector<size_t> wl, bl;
    size_t ew = 0, eb = 0;
    for (const auto &l : d)
        if (l.type == YOLO2CUDA_CONV) {
            wl.push_back((size_t)l.c * l.n * l.size * l.size);
            bl.push_back((size_t)l.n);
            ew += wl.back();
            eb += bl.back();
        }
    if (precision == Precision::FP32) {
        p.w32 = read_binary<float>("weights/weights_reorg.bin");
        p.b32 = read_binary<float>("weights/bias.bin");
        if (p.w32.size() < ew) throw std::runtime_error("weights file too small");
        if (p.b32.size() < eb) throw std::runtime_error("bias file too small");
        return p;
    }
    auto w = read_binary<int16_t>("weights/weights_reorg_int16.bin");
    auto b = read_binary<int16_t>("weights/bias_int16.bin");
    if (w.size() < ew) throw std::runtime_error("weights file too small");
    if (b.size() < eb) throw std::runtime_error("bias file too small");
    p.wq = read_binary<int32_t>("weights/weight_int16_Q.bin");
    p.bq = read_binary<int32_t>("weights/bias_int16_Q.bin");
    if (p.wq.size() < wl.size() || p.bq.size() < wl.size()) throw std::runtime_error("Q tables too small for conv layers");
    p.aq = read_binary<int32_t>("weights/iofm_Q.bin", false);
    if (p.aq.empty()) throw std::runtime_error("Activation Q table (iofm_Q.bin) is required for int16 inference.");
    p.w16.resize(ew);
    p.b16.resize(eb);
    size_t wf = 0, wo = 0, bf = 0, bo = 0;
    for (size_t li = 0; li < wl.size(); ++li) {  // strip the pad element after odd-length layers (yolo2_model.cpp:216-223)
        if (wf + wl[li] > w.size()) throw std::runtime_error("int16 weight truncated at layer " + std::to_string(li));
        if (bf + bl[li] > b.size()) throw std::runtime_error("int16 bias truncated at layer " + std::to_string(li));
        std::memcpy(p.w16.data() + wo, w.data() + wf, wl[li] * sizeof(int16_t));
        std::memcpy(p.b16.data() + bo, b.data() + bf, bl[li] * sizeof(int16_t));
        wf += wl[li] + (wl[li] & 1); wo += wl[li];
        bf += bl[li] + (bl[li] & 1); bo += bl[li];
    }
    return p;
}

void check(yolo2cuda_ctx *ctx, int rc, const char *what)
{
    if (rc != YOLO2CUDA_SUCCESS)
        throw std::runtime_error(std::string(what) + ": " + (ctx ? yolo2cuda_last_error(ctx) : "no CUDA device (the cuda backend has no CPU fallback)"));
}

}  // namespace

void yolov2_cuda_ps_batch(network *net, const float *frames, int batch, float *region_out, Precision precision)
{
    const auto desc = describe(net);
    const Pack pk = load_pack(desc, precision);
    yolo2cuda_ctx *ctx = nullptr;
    const bool i16 = precision == Precision::INT16;
    check(nullptr, yolo2cuda_create(&ctx, 0, i16 ? YOLO2CUDA_PRECISION_INT16 : YOLO2CUDA_PRECISION_FP32), "yolo2cuda_create");
    yolo2cuda_net *n = nullptr;
    try {
        check(ctx, yolo2cuda_net_create(ctx, desc.data(), (int)desc.size(), batch < 64 ? batch : 64, &n), "yolo2cuda_net_create");
        if (i16)
            check(ctx, yolo2cuda_net_load_weights(n, pk.w16.data(), pk.w16.size(), pk.b16.data(), pk.b16.size(), pk.wq.data(),
                                                  pk.bq.data(), (int)std::min(pk.wq.size(), pk.bq.size()), pk.aq.data(), (int)pk.aq.size()),
                  "yolo2cuda_net_load_weights");
        else
            check(ctx, yolo2cuda_net_load_weights(n, pk.w32.data(), pk.w32.size(), pk.b32.data(), pk.b32.size(), nullptr, nullptr, 0, nullptr, 0),
                  "yolo2cuda_net_load_weights");
        check(ctx, yolo2cuda_net_forward_host(n, frames, batch, region_out), "yolo2cuda_net_forward_host");
    } catch (...) {
        if (n) yolo2cuda_net_destroy(n);
        yolo2cuda_destroy(ctx);
        throw;
    }
    yolo2cuda_net_destroy(n);
    yolo2cuda_destroy(ctx);
}

void yolov2_cuda_ps(network *net, const float *input, Precision precision)
{
    layer last = net->layers[net->n - 1];
    yolov2_cuda_ps_batch(net, input, 1, last.output, precision);   // same side effect as forward_region_layer(l, ...)
}
