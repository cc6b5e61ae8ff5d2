// Test infrastructure (development container only).  The C++ host check links a handful of lc-core / lc-runtime sources compiled
// where they lie in the reference tree; the few symbols below belong to libraries that cannot be built here — lc-ast (the whole
// DSL front end) and lc-ir (Rust: luisa_compute_ir) — and sit on paths the check never takes.  Reaching one is a test bug.
#include <cstdio>
#include <cstdlib>
#include <luisa/ast/type.h>
#include <luisa/ast/function.h>
#include <luisa/ir/ast2ir.h>

[[noreturn]] static void unreachable_stub(const char *what) {
    std::fprintf(stderr, "cpp_host_check: %s is outside the check (needs lc-ast / lc-ir)\n", what);
    std::abort();
}

namespace luisa::compute {
bool Type::is_resource() const noexcept { unreachable_stub("Type::is_resource"); }
size_t Type::size() const noexcept { unreachable_stub("Type::size"); }
luisa::span<const Function::Binding> Function::bound_arguments() const noexcept { unreachable_stub("Function::bound_arguments"); }
luisa::shared_ptr<ir::CArc<ir::KernelModule>> AST2IR::build_kernel(Function) noexcept { unreachable_stub("AST2IR::build_kernel"); }
ir::CArc<ir::Type> AST2IR::build_type(const Type *) noexcept { unreachable_stub("AST2IR::build_type"); }
}// namespace luisa::compute
