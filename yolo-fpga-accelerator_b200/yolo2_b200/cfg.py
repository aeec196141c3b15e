"""Darknet cfg -> shape-propagated layer table for the accelerator path.

Mirrors what the reference's parse_network_cfg produces for the section types the YOLOv2 driver
handles (src/core/yolo_net.cpp:218-291; per-section parsers src/core/yolo_layers.cpp:90-117
convolutional, :119-160 route, :188-240 region, :272-287 reorg, :312-326 maxpool).
Only the fields yolov2_hls_ps reads (hls/models/yolov2/yolo2_model.cpp:294-446) are kept.
"""
from dataclasses import dataclass, field
from typing import List

from . import _capi

CONV, MAXPOOL, REORG, ROUTE, REGION = _capi.CONV, _capi.MAXPOOL, _capi.REORG, _capi.ROUTE, _capi.REGION
_TYPE_NAMES = {"convolutional": CONV, "conv": CONV, "maxpool": MAXPOOL, "max": MAXPOOL, "reorg": REORG,
               "route": ROUTE, "region": REGION}


@dataclass
class Layer:
    type: int
    c: int = 0
    h: int = 0
    w: int = 0
    out_c: int = 0
    out_h: int = 0
    out_w: int = 0
    n: int = 0
    size: int = 0
    stride: int = 0
    pad: int = 0
    leaky: int = 0
    batch_normalize: int = 0
    inputs: List[int] = field(default_factory=list)
    classes: int = 0
    coords: int = 0
    softmax: int = 0
    background: int = 0
    anchors: List[float] = field(default_factory=list)

    @property
    def outputs(self):
        return self.out_c * self.out_h * self.out_w


@dataclass
class Network:
    w: int
    h: int
    c: int
    layers: List[Layer]

    @property
    def conv_layers(self):
        return [l for l in self.layers if l.type == CONV]

    def weight_counts(self):
        """Per-conv (weights, biases) element counts = the reference's weight_offsets/beta_offsets
        tables (hls/models/yolov2/model_config.cpp:4-10), derived instead of hard-coded."""
        return [(l.c * l.n * l.size * l.size, l.n) for l in self.conv_layers]


def _sections(text):
    secs, cur = [], None
    for raw in text.splitlines():
        line = raw.strip()
        if not line or line[0] in "#;":
            continue
        if line.startswith("["):
            cur = (line.strip("[]").strip().lower(), {})
            secs.append(cur)
        elif "=" in line and cur is not None:
            k, v = line.split("=", 1)
            cur[1][k.strip()] = v.strip()
    return secs


def parse_network_cfg(path_or_text, width=None, height=None) -> Network:
    """Parses a cfg file (or cfg text). `width`/`height` override [net] (e.g. 608x608)."""
    text = path_or_text
    if "\n" not in path_or_text:
        with open(path_or_text) as f:
            text = f.read()
    secs = _sections(text)
    if not secs or secs[0][0] not in ("net", "network"):
        raise ValueError("First section must be [net] or [network]")  # yolo_net.cpp:225
    net_opts = secs[0][1]
    W = int(width or net_opts.get("width", 0))
    H = int(height or net_opts.get("height", 0))
    Cc = int(net_opts.get("channels", 0))
    if not (W and H and Cc):
        raise ValueError("No input parameters supplied")
    layers: List[Layer] = []
    c, h, w = Cc, H, W
    for idx, (name, o) in enumerate(secs[1:]):
        if name not in _TYPE_NAMES:
            raise ValueError(f"Type not recognized or not on the accelerator path: {name}")
        t = _TYPE_NAMES[name]
        if t == CONV:
            n = int(o.get("filters", 1)); size = int(o.get("size", 1)); stride = int(o.get("stride", 1))
            pad = int(o.get("pad", 0)); padding = int(o.get("padding", 0))
            if pad:
                padding = size // 2                       # yolo_layers.cpp:98
            act = o.get("activation", "logistic").lower()
            if act not in ("leaky", "linear"):
                raise ValueError(f"activation {act} is not supported by the accelerator (leaky/linear only)")
            l = Layer(CONV, c=c, h=h, w=w, n=n, size=size, stride=stride, pad=padding,
                      leaky=int(act == "leaky"), batch_normalize=int(o.get("batch_normalize", 0)))
            l.out_c = n
            l.out_h = (h + 2 * padding - size) // stride + 1
            l.out_w = (w + 2 * padding - size) // stride + 1
        elif t == MAXPOOL:
            stride = int(o.get("stride", 1)); size = int(o.get("size", stride))
            padding = int(o.get("padding", size - 1))      # yolo_layers.cpp:314-316
            l = Layer(MAXPOOL, c=c, h=h, w=w, n=c, size=size, stride=stride, pad=padding)
            l.out_c = c
            l.out_h = (h + padding - size) // stride + 1   # yolo_layers.cpp:299-300
            l.out_w = (w + padding - size) // stride + 1
        elif t == REORG:
            stride = int(o.get("stride", 1))
            if int(o.get("reverse", 0)) or int(o.get("flatten", 0)) or int(o.get("extra", 0)):
                raise ValueError("reorg reverse/flatten/extra are not on the accelerator path")
            l = Layer(REORG, c=c, h=h, w=w, stride=stride)
            l.out_c, l.out_h, l.out_w = c * stride * stride, h // stride, w // stride
        elif t == ROUTE:
            srcs = [int(s) for s in o["layers"].split(",")]
            srcs = [s if s >= 0 else idx + s for s in srcs]  # yolo_layers.cpp:134
            first = layers[srcs[0]]
            l = Layer(ROUTE, inputs=srcs)
            l.out_h, l.out_w = first.out_h, first.out_w
            l.out_c = sum(layers[s].out_c for s in srcs)
            for s in srcs[1:]:
                if layers[s].out_h != first.out_h or layers[s].out_w != first.out_w:
                    raise ValueError("route inputs differ in size")
            l.c, l.h, l.w = l.out_c, l.out_h, l.out_w
        else:  # REGION
            coords = int(o.get("coords", 4)); classes = int(o.get("classes", 20)); num = int(o.get("num", 1))
            l = Layer(REGION, c=c, h=h, w=w, n=num, classes=classes, coords=coords,
                      softmax=int(o.get("softmax", 0)), background=int(o.get("background", 0)))
            l.out_c, l.out_h, l.out_w = c, h, w
            if num * (classes + coords + 1) != c:
                raise ValueError("region layer: channel count does not match num*(classes+coords+1)")
            anchors = [float(a) for a in o.get("anchors", "").split(",") if a.strip()]
            l.anchors = (anchors + [0.5] * (2 * num))[: 2 * num] if anchors else [0.5] * (2 * num)
        layers.append(l)
        c, h, w = l.out_c, l.out_h, l.out_w
    return Network(W, H, Cc, layers)


def to_desc_array(net: Network):
    """-> ctypes array of yolo2cuda_layer_desc (include/yolo2cuda.h)."""
    arr = (_capi.LayerDesc * len(net.layers))()
    for d, l in zip(arr, net.layers):
        d.type = l.type
        d.c, d.h, d.w = l.c, l.h, l.w
        d.out_c, d.out_h, d.out_w = l.out_c, l.out_h, l.out_w
        d.n, d.size, d.stride, d.pad = l.n, l.size, l.stride, l.pad
        d.leaky, d.batch_normalize = l.leaky, l.batch_normalize
        d.n_inputs = len(l.inputs)
        for i, s in enumerate(l.inputs[:4]):
            d.inputs[i] = s
        d.classes, d.coords, d.softmax, d.background = l.classes, l.coords, l.softmax, l.background
        for i, a in enumerate(l.anchors[:32]):
            d.anchors[i] = a
    return arr


# The reference's own cfg for COCO (config/yolov2.cfg) restated structurally so that tests,
# bench.py and smoke() do not need /root/reference at run time.  VOC differs only in the head.
def yolov2_cfg_text(width=416, height=416, classes=80, anchors=None, channel_div=1):
    """channel_div > 1 thins every hidden layer (filters // channel_div, at least 4): the same 32-section
    topology and spatial sizes at a fraction of the work, for quick parity tests."""
    anchors = anchors or ("0.57273, 0.677385, 1.87446, 2.06253, 3.33843, 5.47434, 7.88282, 3.52778, 9.77052, 9.16828"
                          if classes == 80 else
                          "1.3221, 1.73145, 3.19275, 4.00944, 5.05587, 8.09892, 9.47112, 4.84053, 11.2364, 10.0071")

    def conv(f, s, act="leaky", bn=1):
        if act == "leaky" and channel_div > 1:
            f = max(4, f // channel_div)
        return f"[convolutional]\n{'batch_normalize=1' if bn else ''}\nfilters={f}\nsize={s}\nstride=1\npad=1\nactivation={act}\n"
    mp = "[maxpool]\nsize=2\nstride=2\n"
    t = f"[net]\nwidth={width}\nheight={height}\nchannels=3\n"
    t += conv(32, 3) + mp + conv(64, 3) + mp + conv(128, 3) + conv(64, 1) + conv(128, 3) + mp
    t += conv(256, 3) + conv(128, 1) + conv(256, 3) + mp
    t += conv(512, 3) + conv(256, 1) + conv(512, 3) + conv(256, 1) + conv(512, 3) + mp
    t += conv(1024, 3) + conv(512, 1) + conv(1024, 3) + conv(512, 1) + conv(1024, 3)
    t += conv(1024, 3) + conv(1024, 3)
    t += "[route]\nlayers=-9\n" + conv(64, 1) + "[reorg]\nstride=2\n" + "[route]\nlayers=-1,-4\n"
    t += conv(1024, 3) + conv(5 * (classes + 5), 1, act="linear", bn=0)
    t += f"[region]\nanchors = {anchors}\nbias_match=1\nclasses={classes}\ncoords=4\nnum=5\nsoftmax=1\n"
    return t
