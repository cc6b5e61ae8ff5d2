"""CPU tests of the drop-in boundary: the library loads, exports every symbol include/lc_b200_api.h declares, and the
ctypes mirror has the sizes the reference's cbindgen header produces.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import luisa_compute_rs_b200 as lc

abi = lc._abi
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = abi.load_library()
    header = open(os.path.join(ROOT, "include", "lc_b200_api.h")).read()
    declared = set(re.findall(r"^LCB_EXPORT[^;(]*?\b(\w+)\s*\(", header, re.M))
    assert declared == set(abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.lc_b200_version()


def test_lib_interface_table_is_populated():
    iface = abi.load_library().luisa_compute_lib_interface()
    for f in ("set_logger_callback", "create_context", "destroy_context", "create_device", "free_string"):
        assert C.cast(getattr(iface, f), C.c_void_p).value
    ctx = iface.create_context(b".")
    iface.destroy_context(ctx)


def test_struct_sizes_match_reference_header():
    # sizes measured from LC/include/luisa/rust/api_types.h with gcc
    assert C.sizeof(abi.Command) == 88
    assert C.sizeof(abi.CmdMeshBuild) == 80
    assert C.sizeof(abi.CmdAccelBuild) == 40
    assert C.sizeof(abi.AccelModification) == 72
    assert C.sizeof(abi.AccelOption) == 8
    assert C.sizeof(abi.DeviceInterface) == 288
    assert C.sizeof(abi.LibInterface) == 48
    assert C.sizeof(abi.CreatedBuffer) == 32
    assert lc.Ray.itemsize == 32 and lc.SurfaceHit.itemsize == 24 and lc.Index.itemsize == 12


def test_make_ir_type_blocks_have_the_ir_layout():
    lib = abi.load_library()
    p = lib.lc_b200_make_ir_type(12, 4)       # -> &CArc<Type>: pointer to {inner*}
    inner = C.cast(p, C.POINTER(C.c_void_p))[0]
    type_ptr = C.cast(inner, C.POINTER(C.c_void_p))[0]   # CArcSharedBlock.ptr at offset 0
    tag = C.cast(type_ptr, C.POINTER(C.c_int32))[0]
    assert tag == 5  # Type::Struct
    size = C.cast(type_ptr + 8 + 24 + 8, C.POINTER(C.c_size_t))[0]  # union@8: fields slice (24) | alignment | size
    assert size == 12
    v = lib.lc_b200_make_ir_type(0, 0)
    inner = C.cast(v, C.POINTER(C.c_void_p))[0]
    assert C.cast(C.cast(inner, C.POINTER(C.c_void_p))[0], C.POINTER(C.c_int32))[0] == 0  # Type::Void


def test_affine_packing_matches_into_affine3x4():
    import numpy as np
    m = np.arange(16, dtype=np.float32).reshape(4, 4)
    assert lc.affine_from_mat4(m).tolist() == list(range(12))


def test_foreign_device_names_are_forwarded_to_the_stock_library(tmp_path):
    """Drop-in under the frontend's fixed file name (liblc-api.so): device names other than "b200" go to the original backend library
    (liblc-api-orig.so next to ours, or LC_B200_FORWARD_LIB).  A stand-in library records the call; runs in a subprocess because the
    forwarding target is resolved once per process."""
    import subprocess, sys, textwrap
    src = tmp_path / "fake_backend.c"
    src.write_text(textwrap.dedent('''
        #include <stdint.h>
        #include <string.h>
        #include "lc_b200_api.h"
        static char seen[64];
        static lcb_context create_context(const char *dir) { lcb_context c = {42}; return c; }
        static void destroy_context(lcb_context c) {}
        static void set_logger(void (*cb)(lcb_logger_message)) {}
        static void free_string(char *s) {}
        static uint32_t warp(lcb_device d) { return 7; }
        static char *query(lcb_device d, const char *name) { return seen; }
        static lcb_device_interface create_device(lcb_context c, const char *name, const char *json) {
            lcb_device_interface t; memset(&t, 0, sizeof(t));
            t.device.id = c.id * 100 + 1; t.compute_warp_size = warp; t.query = query;
            strncpy(seen, name, sizeof(seen) - 1);
            return t;
        }
        lcb_lib_interface luisa_compute_lib_interface(void) {
            lcb_lib_interface l; memset(&l, 0, sizeof(l));
            l.set_logger_callback = set_logger; l.create_context = create_context; l.destroy_context = destroy_context; l.create_device = create_device; l.free_string = free_string;
            return l;
        }
    '''))
    so = tmp_path / "libfake.so"
    subprocess.run(["gcc", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", str(so), str(src)], check=True)
    code = textwrap.dedent(f'''
        import ctypes as C, sys
        sys.path.insert(0, {ROOT!r})
        import luisa_compute_rs_b200 as lc
        abi = lc._abi
        iface = abi.load_library().luisa_compute_lib_interface()
        ctx = iface.create_context(b".")
        dev = iface.create_device(ctx, b"cpu", None)
        assert dev.device.id == 4201, dev.device.id
        assert dev.compute_warp_size(dev.device) == 7
        assert C.string_at(dev.query(dev.device, b"device_name")) == b"cpu"
        print("forwarded")
    ''')
    env = dict(os.environ, LC_B200_FORWARD_LIB=str(so))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "forwarded" in r.stdout, r.stdout + r.stderr
