"""Development aid: lower one of the example IR kernels, write the CUDA source and run nvcc -Xptxas -v on it (registers, spills).
usage: python tools/lower_ptxas.py {raytracing|path_tracer|c3_trace|tiled_path_tracer} [outdir]   (LC_B200_LOWERING=direct|wavefront is honoured)"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import luisa_compute_rs_b200 as lc  # noqa: E402
from luisa_compute_rs_b200 import examples_ir  # noqa: E402


def main():
    which = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 else "/tmp/lk"
    os.makedirs(out, exist_ok=True)
    lib = lc._abi.load_library()
    if which == "raytracing":
        k = examples_ir.raytracing_kernel(1, 2, 64, 64)
    elif which == "path_tracer":
        k = examples_ir.path_tracer_kernel(1, 2, polynomial_sincos=True)
    elif which == "c3_trace":
        k = examples_ir.trace_buffer_kernel()
    elif which == "tiled_path_tracer":
        k = examples_ir.tiled_path_tracer_kernel(1, 2)
    else:
        raise SystemExit("unknown kernel " + which)
    src = C.string_at(lib.lc_b200_ir_lower_source(C.addressof(k.km))).decode()
    tag = os.environ.get("LC_B200_LOWERING", "auto")
    path = os.path.join(out, f"{which}_{tag}.cu")
    open(path, "w").write(src)
    csrc = os.path.join(ROOT, "luisa-compute-rs_b200", "csrc")
    cmd = ["/usr/local/cuda/bin/nvcc", "-arch=sm_100a", "-std=c++17", "-I", csrc, "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-lineinfo", "-diag-suppress=177",
           "-Xptxas", "-v", "-cubin", "-o", path.replace(".cu", ".cubin"), path]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stderr[-3000:])
    print("source:", path)


if __name__ == "__main__":
    main()
