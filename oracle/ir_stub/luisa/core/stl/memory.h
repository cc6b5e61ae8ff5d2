// Stub of the one LuisaCompute core header that LC/include/luisa/rust/ir_common.h includes, so that the layout checker
// (oracle/ir_layout_check.cpp) can compile the reference's ir.hpp in place without the rest of the C++ tree.
#pragma once
#include <cstddef>
namespace luisa {
template <class T> T *allocate_with_allocator(size_t n);
template <class T> void deallocate_with_allocator(T *p);
}  // namespace luisa
