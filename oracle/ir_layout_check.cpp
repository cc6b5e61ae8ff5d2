// ir_layout_check.cpp — TEST INFRASTRUCTURE (never linked into the product).
// Pins luisa-compute-rs_b200/csrc/ir_layout.h against the reference's own cbindgen header
//   /root/reference/luisa_compute_sys/LuisaCompute/include/luisa/rust/ir.hpp (+ ir_common.h),
// compiled where it lies (oracle/Makefile target `ir_layout`; only the development container has the reference tree).
// For every key of ir_layout_keys.inc it compares sizeof / offsetof / discriminant on both sides, prints the
// reference-side table as JSON (committed as tests/golden/ir_layout_reference.json) and exits 1 on any difference.
#include <cstdio>
#include <cstring>
#include "luisa/rust/ir.hpp"
#include "../luisa-compute-rs_b200/csrc/ir_layout.h"

namespace R = luisa::compute::ir;
namespace M = lcb::ir;

int main() {
    int bad = 0, n = 0;
    printf("{\n");
#define K(key, RV, MV)                                                                                       \
    {                                                                                                        \
        const size_t r = (size_t)(RV), m = (size_t)(MV);                                                     \
        printf("%s  \"%s\": %zu", n++ ? ",\n" : "", key, r);                                                 \
        if (r != m) { fprintf(stderr, "MISMATCH %s: reference %zu, ir_layout.h %zu\n", key, r, m); bad++; }  \
    }
#include "../luisa-compute-rs_b200/csrc/ir_layout_keys.inc"
#undef K
    printf("\n}\n");
    fprintf(stderr, "%d keys, %d mismatches\n", n, bad);
    return bad ? 1 : 0;
}
