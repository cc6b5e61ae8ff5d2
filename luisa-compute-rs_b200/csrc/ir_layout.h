// ir_layout.h — read-only C++ view of the frontend's #[repr(C)] SSA IR, as handed to create_shader through
// LCKernelModule.ptr (a `*const ir::KernelModule`, luisa_compute_backend/src/proxy.rs:188-193) and to create_buffer
// (`&CArc<ir::Type>`).  Written from the Rust definitions (luisa_compute_ir/src/ir.rs: Type :100-200, Func :605-949,
// Const :951-983, Instruction :1168-1262, Node :1264-1272, BasicBlock :1347-1351, Module/KernelModule :1840-1990,
// containers in luisa_compute_ir/src/lib.rs) and pinned against cbindgen's C++ rendering of the same types
// (LC/include/luisa/rust/ir.hpp, ir_common.h) by oracle/ir_layout_check.cpp: sizeof / offsetof / discriminant of
// everything below is compared with the reference header at build time in the development container, and the table is
// committed as tests/golden/ir_layout_reference.json for machines without the reference tree.
//
// A #[repr(C)] enum with payload is { c_int tag; union of payload structs }, each payload aligned naturally.
#pragma once
#include <cstddef>
#include <cstdint>

namespace lcb {
namespace ir {

template <class T> struct ArcBlock { T *ptr; size_t ref_count; void (*destructor)(ArcBlock<T> *); };
template <class T> struct Arc {
    ArcBlock<T> *inner;
    const T *get() const { return inner ? inner->ptr : nullptr; }
    const T *operator->() const { return inner->ptr; }
};
template <class T> struct Slice {
    T *ptr; size_t len; void (*destructor)(T *, size_t);
    const T *begin() const { return ptr; }
    const T *end() const { return ptr + len; }
    const T &operator[](size_t i) const { return ptr[i]; }
};
template <class T> struct Pool { T *ptr; };

struct Node;
typedef size_t NodeRef;  // address of a Node inside the module's pools; 0 = INVALID_REF (ir.rs:590, :1602-1605)
inline const Node *node(NodeRef r) { return reinterpret_cast<const Node *>(r); }

struct BasicBlock { NodeRef first, last; };  // two sentinel nodes; the content is first->next .. last->prev (ir.rs:1401-1406)

enum ModuleKind : int32_t { MK_Block, MK_Function, MK_Kernel };
enum Primitive : int32_t { P_Bool, P_Int8, P_Uint8, P_Int16, P_Uint16, P_Int32, P_Uint32, P_Int64, P_Uint64, P_Float16, P_Float32, P_Float64 };

struct ModulePools;
struct Module {
    int32_t kind;
    Pool<BasicBlock> entry;
    uint32_t flags;
    uint32_t curve_basis_set;
    Arc<ModulePools> pools;
};

struct VectorType;
struct VectorElementType {
    enum : int32_t { Scalar, Vector };
    int32_t tag;
    union { int32_t scalar; Arc<VectorType> vector; };
};
struct VectorType { VectorElementType element; uint32_t length; };
struct MatrixType { VectorElementType element; uint32_t dimension; };
struct Type;
struct StructType { Slice<Arc<Type>> fields; size_t alignment; size_t size; };
struct ArrayType { Arc<Type> element; size_t length; };
struct Type {
    enum : int32_t { Void, UserData, Primitive, Vector, Matrix, Struct, Array, Opaque };
    int32_t tag;
    union {
        int32_t primitive;
        VectorType vector;
        MatrixType matrix;
        StructType struct_;
        ArrayType array;
        Slice<uint8_t> opaque;
    };
};

struct BufferBinding { uint64_t handle; uint64_t offset; size_t size; };
struct TextureBinding { uint64_t handle; uint32_t level; };
struct Binding {
    enum : int32_t { Buffer, Texture, BindlessArray, Accel };
    int32_t tag;
    union { BufferBinding buffer; TextureBinding texture; uint64_t bindless_array; uint64_t accel; };
};
struct Capture { NodeRef node; Binding binding; };

struct CpuCustomOp { uint8_t *data; void (*func)(uint8_t *, uint8_t *); void (*destructor)(uint8_t *); Arc<Type> arg_type; };

struct CallableModule {
    Module module;
    Arc<Type> ret_type;
    Slice<NodeRef> args;
    Slice<Capture> captures;
    Slice<Arc<CpuCustomOp>> cpu_custom_ops;
    Arc<ModulePools> pools;
};
struct KernelModule {
    Module module;
    Slice<Capture> captures;
    Slice<NodeRef> args;
    Slice<NodeRef> shared;
    Slice<Arc<CpuCustomOp>> cpu_custom_ops;
    uint32_t block_size[3];
    Arc<ModulePools> pools;
};

struct Func {
    enum Tag : int32_t {
#define X(name) name,
#include "ir_funcs.inc"
#undef X
        COUNT
    };
    int32_t tag;
    union {
        Slice<uint8_t> message;           // Unreachable / Assert / External
        Arc<ir::CallableModule> callable;    // Callable(CallableModuleRef)
        Arc<ir::CpuCustomOp> cpu_custom_op;
    };
};

struct Const {
    enum : int32_t { Zero, One, Bool, Int8, Uint8, Int16, Uint16, Int32, Uint32, Int64, Uint64, Float16, Float32, Float64, Generic };
    int32_t tag;
    union {
        Arc<Type> type;  // Zero / One
        bool b; int8_t i8; uint8_t u8; int16_t i16; uint16_t u16; int32_t i32; uint32_t u32; int64_t i64; uint64_t u64;
        uint16_t f16_bits; float f32; double f64;
        struct { Slice<uint8_t> bytes; Arc<Type> type; } generic;
    };
};

struct PhiIncoming { NodeRef value; Pool<BasicBlock> block; };
struct SwitchCase { int32_t value; Pool<BasicBlock> block; };
struct UserData { uint64_t tag; const uint8_t *data; bool (*eq)(const uint8_t *, const uint8_t *); };

struct Instruction {
    enum : int32_t {
        Buffer, Bindless, Texture2D, Texture3D, Accel, Shared, Uniform, Local, Argument, UserData, Invalid, Const, Update, Call, Phi, Return,
        Loop, GenericLoop, Break, Continue, If, Switch, AdScope, RayQuery, Print, AdDetach, Comment
    };
    int32_t tag;
    union {
        struct { NodeRef init; } local;
        struct { bool by_value; } argument;
        Arc<ir::UserData> user_data;
        ir::Const const_;
        struct { NodeRef var; NodeRef value; } update;
        struct { Func func; Slice<NodeRef> args; } call;
        Slice<PhiIncoming> phi;
        NodeRef return_;
        struct { Pool<BasicBlock> body; NodeRef cond; } loop;
        struct { Pool<BasicBlock> prepare; NodeRef cond; Pool<BasicBlock> body; Pool<BasicBlock> update; } generic_loop;
        struct { NodeRef cond; Pool<BasicBlock> true_branch; Pool<BasicBlock> false_branch; } if_;
        struct { NodeRef value; Pool<BasicBlock> default_; Slice<SwitchCase> cases; } switch_;
        struct { Pool<BasicBlock> body; bool forward; size_t n_forward_grads; } ad_scope;
        struct { NodeRef ray_query; Pool<BasicBlock> on_triangle_hit; Pool<BasicBlock> on_procedural_hit; } ray_query;
        struct { Slice<uint8_t> fmt; Slice<NodeRef> args; } print;
        Pool<BasicBlock> ad_detach;
        Slice<uint8_t> comment;
    };
};

struct Node {
    Arc<Type> type_;
    NodeRef next;
    NodeRef prev;
    Arc<Instruction> instruction;
};

static_assert(sizeof(ArcBlock<int>) == 24 && sizeof(Slice<int>) == 24, "container layouts (ir_common.h:28-39, ir.hpp:165-170)");
static_assert(sizeof(Module) == 32 && sizeof(Type) == 48 && sizeof(Binding) == 32 && sizeof(Capture) == 40, "ir.hpp layouts");
static_assert(sizeof(Func) == 32 && sizeof(Const) == 40 && sizeof(Instruction) == 64 && sizeof(Node) == 32, "ir.hpp layouts");
static_assert(sizeof(KernelModule) == 152 && sizeof(CallableModule) == 120, "ir.hpp layouts");

}  // namespace ir
}  // namespace lcb
