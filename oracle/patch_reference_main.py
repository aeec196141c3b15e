#!/usr/bin/env python3
"""TEST INFRASTRUCTURE - writes oracle/_ref/yolov2_main_cuda.cpp: the reference's own CLI
(src/models/yolov2/yolov2_main.cpp) with the four-line `--backend cuda` integration of INTEGRATION.md
applied.  The reference source is read where it lies; the patched copy only ever exists under the
git-ignored oracle/_ref/."""
import sys

src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()


def sub(old, new):
    global s
    assert s.count(old) == 1, f"anchor not found exactly once: {old!r}"
    s = s.replace(old, new)


sub("#include <api.hpp>", "#include <api.hpp>\n#include \"yolov2_cuda_ps.hpp\"  // --backend cuda")
sub("enum class Backend { Hls, Cpu } backend = Backend::Hls;", "enum class Backend { Hls, Cpu, Cuda } backend = Backend::Hls;")
sub("""            } else if (backend_val == "cpu") {
                cfg.backend = AppConfig::Backend::Cpu;""", """            } else if (backend_val == "cpu") {
                cfg.backend = AppConfig::Backend::Cpu;
            } else if (backend_val == "cuda") {
                cfg.backend = AppConfig::Backend::Cuda;""")
sub("""        case AppConfig::Backend::Cpu:""", """        case AppConfig::Backend::Cuda:
            yolov2_cuda_ps(net_guard.ptr, sized.img.data, cfg.precision);
            break;
        case AppConfig::Backend::Cpu:""")
open(dst, "w").write(s)
