// ir_ref_build.hpp — TEST INFRASTRUCTURE (development container only: needs the reference tree).
// An ir::KernelModule built with the REFERENCE's own C++ definitions (LC/include/luisa/rust/ir.hpp, compiled where it lies):
//   kernel(buffer<float> a, uint n):  i = dispatch_id().x;  if (i < n) a[i] = a[i] * 2 + 1;
// Shared by ir_ref_module.cpp (lowering + NVRTC compile check, no GPU) and cpp_host/cpp_host_check.cpp (create_shader + ShaderDispatchCommand
// through the C++ DeviceInterface on B200).
#pragma once
#include <cstdlib>
#include <cstring>
#include <vector>
#include "luisa/rust/ir.hpp"

namespace ir_ref {
using namespace luisa::compute::ir;

template <class T> inline CArc<T> arc(const T &v) {
    auto *blk = new CArcSharedBlock<T>{new T(v), {1}, nullptr};
    return CArc<T>{blk};
}
template <class T> inline CBoxedSlice<T> slice(const std::vector<T> &v) {
    T *p = v.empty() ? nullptr : (T *)malloc(sizeof(T) * v.size());
    if (p) memcpy(p, v.data(), sizeof(T) * v.size());
    return CBoxedSlice<T>{p, v.size(), nullptr};
}
inline CArc<Type> prim(Primitive p) { Type t{}; t.tag = Type::Tag::Primitive; t.primitive._0 = p; return arc(t); }
inline CArc<Type> vec(Primitive p, uint32_t n) {
    Type t{}; t.tag = Type::Tag::Vector; t.vector._0.element.tag = VectorElementType::Tag::Scalar; t.vector._0.element.scalar._0 = p; t.vector._0.length = n; return arc(t);
}
inline CArc<Type> void_ty() { Type t{}; t.tag = Type::Tag::Void; return arc(t); }

inline NodeRef new_node(CArc<Type> ty, const Instruction &ins) { return NodeRef{(size_t) new Node{ty, INVALID_REF, INVALID_REF, arc(ins)}}; }
inline Node *get(NodeRef r) { return (Node *)r._0; }

struct Block {
    BasicBlock *bb;
    Block() {
        Instruction inv{}; inv.tag = Instruction::Tag::Invalid;
        NodeRef first = new_node(void_ty(), inv), last = new_node(void_ty(), inv);
        get(first)->next = last; get(last)->prev = first;
        bb = new BasicBlock{first, last};
    }
    NodeRef append(NodeRef n) {
        Node *last = get(bb->last), *prev = get(last->prev);
        get(n)->prev = last->prev; get(n)->next = bb->last; prev->next = n; last->prev = n;
        return n;
    }
    NodeRef call(Func::Tag f, std::vector<NodeRef> args, CArc<Type> ty) {
        Instruction ins{}; ins.tag = Instruction::Tag::Call; ins.call._0.tag = f; ins.call._1 = slice(args);
        return append(new_node(ty, ins));
    }
    NodeRef const_u32(uint32_t v, CArc<Type> ty) { Instruction ins{}; ins.tag = Instruction::Tag::Const; ins.const_._0.tag = Const::Tag::Uint32; ins.const_._0.uint32._0 = v; return append(new_node(ty, ins)); }
    NodeRef const_f32(float v, CArc<Type> ty) { Instruction ins{}; ins.tag = Instruction::Tag::Const; ins.const_._0.tag = Const::Tag::Float32; ins.const_._0.float32._0 = v; return append(new_node(ty, ins)); }
    Pooled<BasicBlock> pooled() const { return Pooled<BasicBlock>{bb}; }
};


inline KernelModule *build_axpy_module() {
    auto f32 = prim(Primitive::Float32), u32 = prim(Primitive::Uint32), b = prim(Primitive::Bool), u3 = vec(Primitive::Uint32, 3), vd = void_ty();
    Instruction buf{}; buf.tag = Instruction::Tag::Buffer;
    Instruction uni{}; uni.tag = Instruction::Tag::Uniform;
    NodeRef a = new_node(f32, buf), n = new_node(u32, uni);

    Block then_b, else_b, entry;
    NodeRef id = entry.call(Func::Tag::DispatchId, {}, u3);
    NodeRef zero = entry.const_u32(0, u32);
    NodeRef i = entry.call(Func::Tag::ExtractElement, {id, zero}, u32);
    NodeRef cond = entry.call(Func::Tag::Lt, {i, n}, b);
    NodeRef x = then_b.call(Func::Tag::BufferRead, {a, i}, f32);
    NodeRef y = then_b.call(Func::Tag::Add, {then_b.call(Func::Tag::Mul, {x, then_b.const_f32(2.0f, f32)}, f32), then_b.const_f32(1.0f, f32)}, f32);
    then_b.call(Func::Tag::BufferWrite, {a, i, y}, vd);
    Instruction iff{}; iff.tag = Instruction::Tag::If; iff.if_.cond = cond; iff.if_.true_branch = then_b.pooled(); iff.if_.false_branch = else_b.pooled();
    entry.append(new_node(vd, iff));

    KernelModule &km = *new KernelModule{};
    km.module.kind = ModuleKind::Kernel; km.module.entry = entry.pooled(); km.module.flags = ModuleFlags_NONE;
    km.captures = slice(std::vector<Capture>{});
    km.args = slice(std::vector<NodeRef>{a, n});
    km.shared = slice(std::vector<NodeRef>{});
    km.cpu_custom_ops = CBoxedSlice<CArc<CpuCustomOp>>{nullptr, 0, nullptr};
    km.block_size[0] = 128; km.block_size[1] = 1; km.block_size[2] = 1;

    return &km;
}
}  // namespace ir_ref
