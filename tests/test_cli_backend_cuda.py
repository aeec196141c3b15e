"""BASELINE config #1 plumbing: the reference's OWN CLI (src/models/yolov2/yolov2_main.cpp) built with the
four-line `--backend cuda` integration (oracle/Makefile target ref-detect) must print the
same region tensor and the same detections as its `--backend hls` on the same image and weights."""
import os
import subprocess

import numpy as np
import pytest

from yolo2_b200 import cfg as ycfg, weights as yw

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "yolov2_detect_cuda")


def _fixture(d):
    from PIL import Image
    net = ycfg.parse_network_cfg(ycfg.yolov2_cfg_text(416, 416, 80))
    pack = yw.synth_pack(net, "int16", seed=1, table="stress")
    yw.save_reference_files(pack, net, os.path.join(d, "weights"))
    with open(os.path.join(d, "net.cfg"), "w") as f:
        f.write(ycfg.yolov2_cfg_text(416, 416, 80))
    with open(os.path.join(d, "names.txt"), "w") as f:
        f.write("\n".join(f"class{i}" for i in range(80)) + "\n")
    os.makedirs(os.path.join(d, "data", "labels"))
    glyph = Image.new("RGB", (8, 12), (255, 255, 255))
    for ch in range(32, 127):                      # load_alphabet() exits without them (yolo_image.cpp:170-174,207-221)
        for s in range(8):
            glyph.save(os.path.join(d, "data", "labels", f"{ch}_{s}.png"))
    rng = np.random.default_rng(3)
    Image.fromarray(rng.integers(0, 256, (360, 500, 3), dtype=np.uint8)).save(os.path.join(d, "frame.png"))


def _run(d, backend):
    env = dict(os.environ, YOLO2_DUMP_REGION=os.path.join(d, f"region_{backend}.txt"),
               YOLO2_DUMP_REGION_RAW=os.path.join(d, f"raw_{backend}.txt"))
    r = subprocess.run([BIN, "--cfg", "net.cfg", "--names", "names.txt", "--input", "frame.png", "--backend", backend,
                        "--precision", "int16", "--thresh", "0.55", "--output", os.path.join(d, f"out_{backend}")],
                       cwd=d, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    dets = sorted(l for l in r.stdout.splitlines() if l.startswith("class") and l.rstrip().endswith("%"))
    return open(env["YOLO2_DUMP_REGION"]).read(), dets, r.stdout


@pytest.mark.gpu
@pytest.mark.slow
def test_reference_cli_backend_cuda_equals_backend_hls(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/yolov2_detect_cuda not built (make -C oracle ref-detect, needs /root/reference)")
    d = str(tmp_path)
    _fixture(d)
    region_cuda, dets_cuda, out_cuda = _run(d, "cuda")
    region_hls, dets_hls, _ = _run(d, "hls")
    assert "Predicted in" in out_cuda
    assert region_cuda == region_hls                  # layers[31].output, %.9g per float: identical text
    assert len(region_cuda.splitlines()) == 71825
    assert dets_cuda == dets_hls and len(dets_hls) > 0


def test_reference_cli_rejects_unknown_backend(tmp_path):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/yolov2_detect_cuda not built")
    r = subprocess.run([BIN, "--backend", "tpu"], cwd=str(tmp_path), capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Unsupported backend" in r.stderr
