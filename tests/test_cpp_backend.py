"""The C++-ABI backend module: liblc-backend-b200.so = csrc/cpp_backend.cpp, the adapter from
luisa::compute::DeviceInterface (LC/include/luisa/runtime/rhi/device_interface.h:99-220) to the C table of include/lc_b200_api.h.

It is built against the reference's own C++ headers, and the host program that exercises it (oracle/cpp_host/cpp_host_check.cpp) runs
on the reference's own Context / DynamicModule / CommandList code compiled where it lies — both in the development container only
(`make -C luisa-compute-rs_b200/csrc cpp_backend`, `make -C oracle cpp_host`; __graft_entry__.build() does it).  The built files
travel to the GPU box under oracle/_ref/cpp_host/ (a self-contained LuisaCompute runtime directory); these tests only run them."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODULE = os.path.join(ROOT, "luisa-compute-rs_b200", "lib", "liblc-backend-b200.so")
HOST_DIR = os.path.join(ROOT, "oracle", "_ref", "cpp_host")
HOST = os.path.join(HOST_DIR, "lc_cpp_host_check")

needs_module = pytest.mark.skipif(not os.path.exists(MODULE), reason="liblc-backend-b200.so is built only where the reference headers are present")
needs_host = pytest.mark.skipif(not os.path.exists(HOST), reason="oracle/_ref/cpp_host is built only where the reference sources are present")


def _dynamic_symbols(path, defined):
    out = subprocess.run(["nm", "-D", "--defined-only" if defined else "--undefined-only", path], capture_output=True, text=True, check=True).stdout
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


@needs_module
def test_module_exports_what_context_binds():
    # LC/src/runtime/context.cpp:85-87 binds exactly these three, by their C names
    syms = _dynamic_symbols(MODULE, defined=True)
    assert {"create", "destroy", "backend_device_names"} <= syms
    # and it reaches the device only through the drop-in symbol of liblc_b200.so
    assert "luisa_compute_lib_interface" in _dynamic_symbols(MODULE, defined=False)
    needed = subprocess.run(["readelf", "-d", MODULE], capture_output=True, text=True, check=True).stdout
    assert "liblc_b200.so" in needed and "$ORIGIN" in needed


def _gpu_present():
    import torch
    return torch.cuda.is_available()


@needs_host
def test_reference_context_loads_the_module_and_there_is_no_cpu_fallback():
    """Without a GPU the reference's Context still discovers `b200` in its runtime directory, loads the module, asks it for its device
    names and calls create(); the device then refuses to exist — through the host program's logger — instead of falling back."""
    if _gpu_present():
        pytest.skip("a GPU is present: covered by the gpu test")
    r = subprocess.run([HOST], cwd=HOST_DIR, capture_output=True, text=True, timeout=120)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    assert "no CPU fallback" in out and "liblc-backend-b200.so :: create" in out, out[-2000:]


@pytest.mark.gpu
@needs_host
def test_cpp_host_program_matches_the_oracle_on_b200():
    """Context::create_device("b200") -> DeviceInterface virtuals -> CommandList{uploads, MeshBuildCommand, AccelBuildCommand, callback}
    -> batch trace -> BufferDownloadCommand (then a DSL kernel: create_shader + ShaderDispatchCommand), twice (FORCE_BUILD, then PREFER_UPDATE + instance made invisible): closest hits (inst, prim,
    bary, t) and any-hit flags bit-identical to the CPU oracle."""
    r = subprocess.run([HOST], cwd=HOST_DIR, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "cpp_host_check ok" in r.stdout
    assert "round 1:" in r.stdout and "(0 on instance 1)" in r.stdout
    # create_shader(const ir::KernelModule *) + ShaderDispatchCommand{buffer, uniform} packed by the reference's own ComputeDispatchCmdEncoder
    assert "shader: reference-built ir::KernelModule" in r.stdout and "777 of 1000 elements updated" in r.stdout
    assert "textures: Float4 and RGBA8 images round-trip" in r.stdout   # TextureUpload / TextureDownload / BindlessArrayUpdate commands
