// `--backend cuda` for the reference host (see yolov2_cuda_ps.hpp).  Thin: translate the reference's
// `network` into the yolo2cuda_layer_desc table, read the reference's weight files (the file formats
// are the contract: hls/models/yolov2/yolo2_model.cpp:158-227, incl. the odd-length filler rule), call the
// C ABI.  No arithmetic happens here and there is no CPU fallback: errors throw std::runtime_error like
// the reference's loaders do (caught in main, src/models/yolov2/yolov2_main.cpp:340-346).
#include "yolov2_cuda_ps.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include <core/yolo.h>          // the reference's own header
#include <core/precision.hpp>

#include "../../include/yolo2cuda.h"

namespace {

// Whole-file loader for the reference's flat little-endian .bin tables (weights/*.bin next to the cwd, the paths of
// hls/models/yolov2/yolo2_model.cpp:171-193).  The file length must be a whole number of elements.
template <typename T>
struct BinTable {
    std::vector<T> v;
    bool found = false;

    static BinTable open(const std::string &path, bool optional = false)
    {
        BinTable t;
        std::ifstream in(path, std::ios::binary | std::ios::ate);
        if (!in) {
            if (optional) return t;
            throw std::runtime_error("cuda backend: cannot open " + path);
        }
        const std::streamoff bytes = in.tellg();
        if (bytes < 0 || bytes % (std::streamoff)sizeof(T))
            throw std::runtime_error("cuda backend: " + path + " is not a whole number of " + std::to_string(sizeof(T)) + "-byte entries");
        t.v.resize((size_t)bytes / sizeof(T));
        in.seekg(0);
        in.read(reinterpret_cast<char *>(t.v.data()), bytes);
        if (in.gcount() != bytes) throw std::runtime_error("cuda backend: " + path + " ended early");
        t.found = true;
        return t;
    }
};

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

// The five tables of one precision, as the C ABI wants them (conv layers back to back, no padding).
struct Pack {
    std::vector<int16_t> w16, b16;
    std::vector<float> w32, b32;
    std::vector<int32_t> wq, bq, aq;
};

// The int16 files carry ONE filler element after every layer whose element count is odd (the quantiser keeps each layer
// 4-byte aligned; the reference loader steps over it, yolo2_model.cpp:216-223).  Returns the layers packed back to back.
std::vector<int16_t> drop_layer_fillers(const std::vector<int16_t> &file, const std::vector<size_t> &layer_len, const char *what)
{
    size_t total = 0;
    for (size_t n : layer_len) total += n;
    std::vector<int16_t> packed;
    packed.reserve(total);
    size_t at = 0;
    for (size_t li = 0; li < layer_len.size(); ++li) {
        const size_t n = layer_len[li];
        if (at + n > file.size())
            throw std::runtime_error(std::string("cuda backend: int16 ") + what + " file ends inside conv layer " + std::to_string(li));
        packed.insert(packed.end(), file.begin() + at, file.begin() + at + n);
        at += n + (n % 2);
    }
    return packed;
}

Pack load_pack(const std::vector<yolo2cuda_layer_desc> &d, Precision precision)
{
    std::vector<size_t> w_len, b_len;
    size_t w_total = 0, b_total = 0;
    for (const auto &l : d) {
        if (l.type != YOLO2CUDA_CONV) continue;
        w_len.push_back((size_t)l.c * l.n * l.size * l.size);
        b_len.push_back((size_t)l.n);
        w_total += w_len.back();
        b_total += b_len.back();
    }
    Pack p;
    if (precision == Precision::FP32) {
        p.w32 = BinTable<float>::open("weights/weights_reorg.bin").v;
        p.b32 = BinTable<float>::open("weights/bias.bin").v;
        if (p.w32.size() < w_total || p.b32.size() < b_total)
            throw std::runtime_error("cuda backend: fp32 weight/bias files hold fewer values than the cfg's conv layers need");
        return p;
    }
    const auto wfile = BinTable<int16_t>::open("weights/weights_reorg_int16.bin");
    const auto bfile = BinTable<int16_t>::open("weights/bias_int16.bin");
    if (wfile.v.size() < w_total || bfile.v.size() < b_total)
        throw std::runtime_error("cuda backend: int16 weight/bias files hold fewer values than the cfg's conv layers need");
    p.wq = BinTable<int32_t>::open("weights/weight_int16_Q.bin").v;
    p.bq = BinTable<int32_t>::open("weights/bias_int16_Q.bin").v;
    if (p.wq.size() < w_len.size() || p.bq.size() < w_len.size())
        throw std::runtime_error("cuda backend: weight/bias Q tables have fewer entries than there are conv layers");
    // the reference treats iofm_Q.bin as optional at load time and then refuses to run int16 without it (yolo2_model.cpp:190-196,258-260)
    const auto aq = BinTable<int32_t>::open("weights/iofm_Q.bin", /*optional=*/true);
    if (!aq.found || aq.v.empty()) throw std::runtime_error("cuda backend: int16 inference needs the activation Q table weights/iofm_Q.bin");
    p.aq = aq.v;
    p.w16 = drop_layer_fillers(wfile.v, w_len, "weight");
    p.b16 = drop_layer_fillers(bfile.v, b_len, "bias");
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
