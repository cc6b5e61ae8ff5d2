// ir_ref_module.cpp — TEST INFRASTRUCTURE (development container only: needs the reference tree).
// Builds an ir::KernelModule with the REFERENCE's own C++ definitions (LC/include/luisa/rust/ir.hpp, compiled where it lies) —
// the records the Rust frontend hands to create_shader have exactly this layout — and gives it to liblc_b200.so:
//   kernel(buffer<float> a, uint n):  i = dispatch_id().x;  if (i < n) a[i] = a[i] * 2 + 1;
// Prints the CUDA source the library lowers it to and the result of its NVRTC compile check.  Complements ir_layout_check.cpp
// (which compares layouts key by key) with an executable end-to-end check that does not go through the Python IR builder.
#include <cstdio>
#include <dlfcn.h>
#include "ir_ref_build.hpp"

int main(int argc, char **argv) {
    const char *lib_path = argc > 1 ? argv[1] : "../luisa-compute-rs_b200/lib/liblc_b200.so";
    void *lib = dlopen(lib_path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "dlopen %s: %s\n", lib_path, dlerror()); return 2; }
    auto lower = (char *(*)(const void *))dlsym(lib, "lc_b200_ir_lower_source");
    auto check = (int (*)(const void *, bool, char **))dlsym(lib, "lc_b200_shader_compile_check");
    if (!lower || !check) { fprintf(stderr, "missing symbols\n"); return 2; }

    luisa::compute::ir::KernelModule &km = *ir_ref::build_axpy_module();

    char *src = lower(&km);
    printf("%s\n", src);
    char *log = nullptr;
    const int rc = check(&km, false, &log);
    printf("// compile_check rc=%d %s\n", rc, log ? log : "");
    return rc;
}
