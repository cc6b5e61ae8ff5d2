"""Writes tests/golden/device_math_reference.npz from the compiled reference headers (development container only):

    make -C oracle ref_device_math && python tests/golden/make_device_math_golden.py

Every array is what the reference CPU backend's own device_math.h returns for the operands of tests/device_math_cases.py."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import device_math_cases as cases  # noqa: E402

REF_LIB = os.path.join(HERE, "..", "..", "oracle", "_ref", "libref_device_math.so")


def compute(lib_path=REF_LIB):
    lib = C.CDLL(lib_path)
    P = C.c_void_p
    d = cases.inputs()
    out = {}

    def ptr(a):
        return a.ctypes.data_as(P)

    def run(fn, name, args, dtype, rows=cases.N):
        res = np.zeros((rows, 4), dtype)
        f = getattr(lib, fn)
        f.restype = C.c_int
        rc = f(name.encode(), *[ptr(a) for a in args], ptr(res), C.c_size_t(cases.N))
        if rc != 0:
            raise RuntimeError("%s: unknown function %s" % (fn, name))
        return res

    for n in cases.F4_UNARY:
        out["f1_" + n] = run("ref_f4_unary", n, [d["fa"]], np.float32)
    for n in cases.F4_BINARY:
        out["f2_" + n] = run("ref_f4_binary", n, [d["fa"], d["fb"]], np.float32)
    for n in cases.F4_TERNARY:
        out["f3_" + n] = run("ref_f4_ternary", n, [d["fa"], d["fb"], d["fc"]], np.float32)
    for n in cases.F3_GEOMETRY:
        out["g3_" + n] = run("ref_f3_geometry", n, [d["fa"], d["fb"], d["fc"]], np.float32)
    for n in cases.MAT3:
        out["m3_" + n] = run("ref_mat3", n, [d["ga"], d["gb"], d["gc"]], np.float32, rows=3 * cases.N)
    for n in cases.U4:
        out["u4_" + n] = run("ref_u4", n, [d["ua"], d["us"] if n in ("Shl", "Shr") else d["ub"]], np.uint32)
    for n in cases.I4:
        out["i4_" + n] = run("ref_i4", n, [d["ia"], d["is_"] if n == "Shr" else d["ib"]], np.int32)
    for n in cases.F4_TO_U4:
        src = np.abs(d["fcast"]) if n == "CastU32" else d["fcast"] if n == "CastI32" else d["fa"]   # negative float -> unsigned is undefined in C++
        out["fu_" + n] = run("ref_f4_to_u4", n, [np.ascontiguousarray(src)], np.uint32)
    return out


def texture_inputs():
    rng = np.random.default_rng(45)
    w, h, n = 13, 9, 1024
    img = rng.uniform(0, 1, (h, w, 4)).astype(np.float32)
    uv = rng.uniform(-1.5, 2.5, (n, 2)).astype(np.float32)
    uv[:8] = [[0, 0], [1, 1], [0.5, 0.5], [-1, 2], [0.999999, 0.000001], [1.0, 0.0], [2.0, -2.0], [0.25, 0.75]]
    return img, uv


def compute_texture(lib_path=REF_LIB):
    """cpu_texture.h's lc_texture_2d_sample for 2 filters x 4 address modes: key tex2d_<filter>_<address>."""
    lib = C.CDLL(lib_path)
    img, uv = texture_inputs()
    res = {}
    for filt in (0, 1):
        for address in range(4):
            out = np.zeros((uv.shape[0], 4), np.float32)
            lib.ref_texture2d_sample(img.ctypes.data_as(C.c_void_p), img.shape[1], img.shape[0], uv.ctypes.data_as(C.c_void_p), C.c_size_t(uv.shape[0]),
                                     filt, address, out.ctypes.data_as(C.c_void_p))
            res["tex2d_%d_%d" % (filt, address)] = out
    return res


def texture3d_inputs():
    rng = np.random.default_rng(46)
    w, h, d, n = 7, 5, 6, 768
    vol = rng.uniform(0, 1, (d, h, w, 4)).astype(np.float32)
    uvw = rng.uniform(-1.5, 2.5, (n, 4)).astype(np.float32)
    uvw[:6, :3] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [-1, 2, 0.25], [0.999999, 0.000001, 1.0], [2.0, -2.0, 0.75]]
    uvw[:, 3] = 0
    return vol, uvw


def compute_texture3d(lib_path=REF_LIB):
    """cpu_texture.h's lc_texture_3d_sample for 2 filters x 4 address modes: key tex3d_<filter>_<address>."""
    lib = C.CDLL(lib_path)
    vol, uvw = texture3d_inputs()
    res = {}
    for filt in (0, 1):
        for address in range(4):
            out = np.zeros((uvw.shape[0], 4), np.float32)
            lib.ref_texture3d_sample(vol.ctypes.data_as(C.c_void_p), vol.shape[2], vol.shape[1], vol.shape[0], uvw.ctypes.data_as(C.c_void_p), C.c_size_t(uvw.shape[0]),
                                     filt, address, out.ctypes.data_as(C.c_void_p))
            res["tex3d_%d_%d" % (filt, address)] = out
    return res


# format name (luisa-compute-rs_b200/runtime.py PIXEL_FORMATS) -> (LCPixelStorage, log2 bytes per pixel, numpy dtype, channels)
PIXEL_CASES = {"Rgba8Unorm": (2, 2, np.uint8, 4), "R8Unorm": (0, 0, np.uint8, 1), "Rgba16Unorm": (5, 3, np.uint16, 4), "Rgba16f": (11, 3, np.float16, 4),
               "R32f": (12, 2, np.float32, 1), "Rgba32f": (14, 4, np.float32, 4)}
PIXEL_W = 96


def pixel_values():
    rng = np.random.default_rng(0xF1E1D)
    v = rng.uniform(-0.2, 1.2, (PIXEL_W, 4)).astype(np.float32)
    ties = np.array([0.0, -0.0, 1.0, -0.5, 1.5, 0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255, 0.5 / 65535, 1.5 / 65535, 65534.5 / 65535,
                     0.49999 / 255, 0.50001 / 255, 70000.0, -70000.0, 6.0e-8, 1.0e-5, 0.333333, 0.1, 2049.0 / 2048.0, 1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 100.0], np.float32)
    v.reshape(-1)[:ties.size] = ties
    return v


def compute_pixels(lib_path=REF_LIB):
    """cpu_texture.h write -> stored pixel -> read for the storages of PIXEL_CASES: px_<format>_raw (stored texels), px_<format>_back."""
    lib = C.CDLL(lib_path)
    v = pixel_values()
    res = {}
    for name, (storage, shift, dtype, ch) in PIXEL_CASES.items():
        raw = np.zeros(PIXEL_W << shift, np.uint8)
        back = np.zeros((PIXEL_W, 4), np.float32)
        lib.ref_texture2d_write_read(storage, shift, v.ctypes.data_as(C.c_void_p), PIXEL_W, raw.ctypes.data_as(C.c_void_p), back.ctypes.data_as(C.c_void_p))
        res["px_%s_raw" % name] = raw.view(dtype).reshape(PIXEL_W, ch) if ch > 1 else raw.view(dtype).reshape(PIXEL_W)
        res["px_%s_back" % name] = back
    return res


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "pixel_conversion_reference.npz"), **compute_pixels())
    np.savez_compressed(os.path.join(HERE, "texture_sample_reference.npz"), **compute_texture(), **compute_texture3d())
    res = compute()
    np.savez_compressed(os.path.join(HERE, "device_math_reference.npz"), **res)
    print("wrote %d arrays" % len(res))
