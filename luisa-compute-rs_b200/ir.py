"""Host-side construction of the frontend's SSA IR, in memory, in the #[repr(C)] layout `create_shader` receives.

There is no Rust toolchain in this image, so the role of the `luisa_compute` DSL + `luisa_compute_ir::IrBuilder`
(luisa_compute_ir/src/ir.rs:2290-2700; C API `luisa_compute_ir_build_*`, LC/include/luisa/rust/ir.hpp:889-990) is played by this
module: it allocates `Node` / `Instruction` / `Type` / `BasicBlock` / `KernelModule` records with ctypes, byte-compatible with
what the Rust frontend would hand to `DeviceInterface::create_shader` (`LCKernelModule.ptr`, proxy.rs:188-193).  The layouts
are the ones of csrc/ir_layout.h and are pinned against the reference's cbindgen header by tests/test_ir_layout.py.

Two layers:
  * `Module` — raw builder: types are interned, `call/const/local/update/if_/generic_loop/...` append nodes to the current
    basic block exactly like IrBuilder::call / const_ / local / update / if_ / generic_loop (ir.rs:2330-2650).
  * `Value` — operator sugar (``a + b``, ``v.x``, ``cond.select(t, f)``) in the spirit of `Expr<T>` / `Var<T>`
    (luisa_compute/src/lang/types): every operator appends one `Instruction::Call` with the `Func` the Rust DSL would emit.
"""
import ctypes as C
import struct as _struct

# --------------------------------------------------------------------------------------------------------------------------
# discriminants (order = ABI; csrc/ir_funcs.inc is the single source for Func)
# --------------------------------------------------------------------------------------------------------------------------
import os as _os
import re as _re

_here = _os.path.dirname(_os.path.abspath(__file__))
FUNC_NAMES = _re.findall(r"X\((\w+)\)", "".join(l for l in open(_os.path.join(_here, "csrc", "ir_funcs.inc")) if not l.startswith("//")))


class Func:
    pass


for _i, _n in enumerate(FUNC_NAMES):
    setattr(Func, _n, _i)

PRIMS = ["bool", "i8", "u8", "i16", "u16", "i32", "u32", "i64", "u64", "f16", "f32", "f64"]
PRIM_SIZE = [1, 1, 1, 2, 2, 4, 4, 8, 8, 2, 4, 8]
PRIM_FMT = ["?", "b", "B", "h", "H", "i", "I", "q", "Q", "e", "f", "d"]
T_VOID, T_USERDATA, T_PRIMITIVE, T_VECTOR, T_MATRIX, T_STRUCT, T_ARRAY, T_OPAQUE = range(8)
(I_BUFFER, I_BINDLESS, I_TEX2D, I_TEX3D, I_ACCEL, I_SHARED, I_UNIFORM, I_LOCAL, I_ARGUMENT, I_USERDATA, I_INVALID, I_CONST, I_UPDATE, I_CALL, I_PHI,
 I_RETURN, I_LOOP, I_GENERIC_LOOP, I_BREAK, I_CONTINUE, I_IF, I_SWITCH, I_ADSCOPE, I_RAYQUERY, I_PRINT, I_ADDETACH, I_COMMENT) = range(27)
(C_ZERO, C_ONE, C_BOOL, C_I8, C_U8, C_I16, C_U16, C_I32, C_U32, C_I64, C_U64, C_F16, C_F32, C_F64, C_GENERIC) = range(15)
B_BUFFER, B_TEXTURE, B_BINDLESS, B_ACCEL = range(4)

# --------------------------------------------------------------------------------------------------------------------------
# ctypes mirror of csrc/ir_layout.h
# --------------------------------------------------------------------------------------------------------------------------
vp, sz = C.c_void_p, C.c_size_t


class ArcBlock(C.Structure):
    _fields_ = [("ptr", vp), ("ref_count", sz), ("destructor", vp)]


class Slice(C.Structure):
    _fields_ = [("ptr", vp), ("len", sz), ("destructor", vp)]


class BasicBlock(C.Structure):
    _fields_ = [("first", sz), ("last", sz)]


class ModuleS(C.Structure):
    _fields_ = [("kind", C.c_int32), ("entry", vp), ("flags", C.c_uint32), ("curve_basis_set", C.c_uint32), ("pools", vp)]


class _VecElemU(C.Union):
    _fields_ = [("scalar", C.c_int32), ("vector", vp)]


class VectorElementType(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _VecElemU)]


class VectorType(C.Structure):
    _fields_ = [("element", VectorElementType), ("length", C.c_uint32)]


class StructType(C.Structure):
    _fields_ = [("fields", Slice), ("alignment", sz), ("size", sz)]


class ArrayType(C.Structure):
    _fields_ = [("element", vp), ("length", sz)]


class _TypeU(C.Union):
    _fields_ = [("primitive", C.c_int32), ("vector", VectorType), ("matrix", VectorType), ("struct_", StructType), ("array", ArrayType), ("opaque", Slice)]


class TypeS(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _TypeU)]


class BufferBinding(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("offset", C.c_uint64), ("size", sz)]


class TextureBinding(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("level", C.c_uint32)]


class _BindingU(C.Union):
    _fields_ = [("buffer", BufferBinding), ("texture", TextureBinding), ("bindless_array", C.c_uint64), ("accel", C.c_uint64)]


class Binding(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _BindingU)]


class Capture(C.Structure):
    _fields_ = [("node", sz), ("binding", Binding)]


class CallableModuleS(C.Structure):
    _fields_ = [("module", ModuleS), ("ret_type", vp), ("args", Slice), ("captures", Slice), ("cpu_custom_ops", Slice), ("pools", vp)]


class KernelModuleS(C.Structure):
    _fields_ = [("module", ModuleS), ("captures", Slice), ("args", Slice), ("shared", Slice), ("cpu_custom_ops", Slice), ("block_size", C.c_uint32 * 3), ("pools", vp)]


class _FuncU(C.Union):
    _fields_ = [("message", Slice), ("callable", vp), ("cpu_custom_op", vp)]


class FuncS(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _FuncU)]


class _GenericConst(C.Structure):
    _fields_ = [("bytes", Slice), ("type", vp)]


class _ConstU(C.Union):
    _fields_ = [("type", vp), ("b", C.c_bool), ("i8", C.c_int8), ("u8", C.c_uint8), ("i16", C.c_int16), ("u16", C.c_uint16), ("i32", C.c_int32), ("u32", C.c_uint32),
                ("i64", C.c_int64), ("u64", C.c_uint64), ("f32", C.c_float), ("f64", C.c_double), ("generic", _GenericConst)]


class ConstS(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _ConstU)]


class PhiIncoming(C.Structure):
    _fields_ = [("value", sz), ("block", vp)]


class SwitchCase(C.Structure):
    _fields_ = [("value", C.c_int32), ("block", vp)]


class _Local(C.Structure):
    _fields_ = [("init", sz)]


class _Argument(C.Structure):
    _fields_ = [("by_value", C.c_bool)]


class _Update(C.Structure):
    _fields_ = [("var", sz), ("value", sz)]


class _Call(C.Structure):
    _fields_ = [("func", FuncS), ("args", Slice)]


class _Loop(C.Structure):
    _fields_ = [("body", vp), ("cond", sz)]


class _GenericLoop(C.Structure):
    _fields_ = [("prepare", vp), ("cond", sz), ("body", vp), ("update", vp)]


class _If(C.Structure):
    _fields_ = [("cond", sz), ("true_branch", vp), ("false_branch", vp)]


class _Switch(C.Structure):
    _fields_ = [("value", sz), ("default_", vp), ("cases", Slice)]


class _RayQuery(C.Structure):
    _fields_ = [("ray_query", sz), ("on_triangle_hit", vp), ("on_procedural_hit", vp)]


class _Print(C.Structure):
    _fields_ = [("fmt", Slice), ("args", Slice)]


class _InstrU(C.Union):
    _fields_ = [("local", _Local), ("argument", _Argument), ("const_", ConstS), ("update", _Update), ("call", _Call), ("phi", Slice), ("return_", sz), ("loop", _Loop),
                ("generic_loop", _GenericLoop), ("if_", _If), ("switch_", _Switch), ("ray_query", _RayQuery), ("print", _Print), ("comment", Slice)]


class InstructionS(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _InstrU)]


class NodeS(C.Structure):
    _fields_ = [("type_", vp), ("next", sz), ("prev", sz), ("instruction", vp)]


LAYOUT = {  # compared with tests/golden/ir_layout_reference.json
    "sizeof.CArcSharedBlock": C.sizeof(ArcBlock), "sizeof.CBoxedSlice": C.sizeof(Slice), "sizeof.BasicBlock": C.sizeof(BasicBlock), "sizeof.Module": C.sizeof(ModuleS),
    "sizeof.VectorElementType": C.sizeof(VectorElementType), "sizeof.VectorType": C.sizeof(VectorType), "sizeof.StructType": C.sizeof(StructType),
    "sizeof.ArrayType": C.sizeof(ArrayType), "sizeof.Type": C.sizeof(TypeS), "sizeof.Binding": C.sizeof(Binding), "sizeof.Capture": C.sizeof(Capture),
    "sizeof.CallableModule": C.sizeof(CallableModuleS), "sizeof.KernelModule": C.sizeof(KernelModuleS), "sizeof.Func": C.sizeof(FuncS), "sizeof.Const": C.sizeof(ConstS),
    "sizeof.PhiIncoming": C.sizeof(PhiIncoming), "sizeof.SwitchCase": C.sizeof(SwitchCase), "sizeof.Instruction": C.sizeof(InstructionS), "sizeof.Node": C.sizeof(NodeS),
    "offsetof.Type.vector": TypeS.u.offset, "offsetof.Binding.buffer": Binding.u.offset, "offsetof.Capture.binding": Capture.binding.offset,
    "offsetof.KernelModule.captures": KernelModuleS.captures.offset, "offsetof.KernelModule.args": KernelModuleS.args.offset,
    "offsetof.KernelModule.block_size": KernelModuleS.block_size.offset, "offsetof.KernelModule.pools": KernelModuleS.pools.offset,
    "offsetof.CallableModule.ret_type": CallableModuleS.ret_type.offset, "offsetof.CallableModule.args": CallableModuleS.args.offset,
    "offsetof.Func.callable": FuncS.u.offset, "offsetof.Const.float32": ConstS.u.offset, "offsetof.Const.generic.type": ConstS.u.offset + _GenericConst.type.offset,
    "offsetof.Instruction.call.func": InstructionS.u.offset + _Call.func.offset, "offsetof.Instruction.call.args": InstructionS.u.offset + _Call.args.offset,
    "offsetof.Instruction.generic_loop.update": InstructionS.u.offset + _GenericLoop.update.offset, "offsetof.Instruction.if.false_branch": InstructionS.u.offset + _If.false_branch.offset,
    "offsetof.Instruction.switch.cases": InstructionS.u.offset + _Switch.cases.offset,
    "offsetof.Instruction.ray_query.on_procedural_hit": InstructionS.u.offset + _RayQuery.on_procedural_hit.offset, "offsetof.Node.instruction": NodeS.instruction.offset,
}


# --------------------------------------------------------------------------------------------------------------------------
# types
# --------------------------------------------------------------------------------------------------------------------------
class Ty:
    """An interned ir::Type: `arc` is the address of its CArcSharedBlock (what a CArc<Type> field holds)."""

    def __init__(self, key, rec, arc, size, align):
        self.key, self.rec, self.arc, self.size, self.align = key, rec, arc, size, align
        self.kind = rec.tag
        self.fields = []  # struct
        self.element = None  # vector / matrix / array
        self.length = 0

    def __repr__(self):
        return "Ty(%s)" % (self.key,)

    @property
    def is_vector(self):
        return self.kind == T_VECTOR

    @property
    def is_float(self):
        return (self.kind == T_PRIMITIVE and self.rec.u.primitive >= 9) or (self.kind in (T_VECTOR, T_MATRIX) and self.element.is_float)


class Module:
    """Owns every record of one kernel (the role of ModulePools + the type context)."""

    def __init__(self):
        self._keep = []
        self._types = {}
        self.void = self._intern(("void",), lambda t: setattr(t, "tag", T_VOID), 0, 1)
        for i, n in enumerate(PRIMS):
            def init(t, i=i):
                t.tag = T_PRIMITIVE
                t.u.primitive = i
            setattr(self, n, self._intern(("prim", n), init, PRIM_SIZE[i], PRIM_SIZE[i]))
        self._blocks = []  # stack of (BasicBlock, last real node ref)

    # -- memory ---------------------------------------------------------------------------------------------------------
    def keep(self, obj):
        self._keep.append(obj)
        return obj

    def arc(self, obj):
        blk = self.keep(ArcBlock(C.addressof(obj), 1, None))
        return C.addressof(blk)

    def slice(self, ctype, items):
        if not items:
            return Slice(None, 0, None)
        arr = self.keep((ctype * len(items))(*items))
        return Slice(C.addressof(arr), len(items), None)

    def bytes_slice(self, data):
        return self.slice(C.c_uint8, list(data))

    # -- types ----------------------------------------------------------------------------------------------------------
    def _intern(self, key, init, size, align):
        if key in self._types:
            return self._types[key]
        rec = self.keep(TypeS())
        init(rec)
        ty = Ty(key, rec, self.arc(rec), size, align)
        self._types[key] = ty
        return ty

    def vector(self, elem, n):
        assert elem.kind == T_PRIMITIVE and 2 <= n <= 4

        def init(t):
            t.tag = T_VECTOR
            t.u.vector.element.tag = 0
            t.u.vector.element.u.scalar = elem.rec.u.primitive
            t.u.vector.length = n
        ty = self._intern(("vec", elem.key, n), init, elem.size * (4 if n == 3 else n), min(elem.size * (4 if n == 3 else n), 16))
        ty.element, ty.length = elem, n
        return ty

    def matrix(self, n):
        def init(t):
            t.tag = T_MATRIX
            t.u.matrix.element.tag = 0
            t.u.matrix.element.u.scalar = 10
            t.u.matrix.length = n
        ty = self._intern(("mat", n), init, 4 * (2 if n == 2 else 4) * n, min(4 * (4 if n == 3 else n), 16))
        ty.element, ty.length = self.f32, n
        return ty

    def array(self, elem, n):
        def init(t):
            t.tag = T_ARRAY
            t.u.array.element = elem.arc
            t.u.array.length = n
        ty = self._intern(("array", elem.key, n), init, elem.size * n, elem.align)
        ty.element, ty.length = elem, n
        return ty

    def struct(self, fields, align=None):
        """#[repr(C)] struct: fields at naturally aligned offsets; `align` raises the struct alignment (e.g. Ray is align(16))."""
        off, a = 0, 1
        for f in fields:
            off = (off + f.align - 1) // f.align * f.align + f.size
            a = max(a, f.align)
        a = max(a, align or 1)
        size = (off + a - 1) // a * a

        def init(t):
            t.tag = T_STRUCT
            t.u.struct_.fields = self.slice(vp, [f.arc for f in fields])
            t.u.struct_.alignment = a
            t.u.struct_.size = size
        ty = self._intern(("struct", tuple(f.key for f in fields), a), init, size, a)
        ty.fields = list(fields)
        return ty

    def opaque(self, name):
        """Type::Opaque(name): "LC_RayQueryAll" / "LC_RayQueryAny" are the RayQuery objects (rtx.rs:672-700)."""
        def init(t):
            t.tag = T_OPAQUE
            t.u.opaque = self.bytes_slice(name.encode())
        return self._intern(("opaque", name), init, 0, 1)

    # shorthand for the common vector types
    def __getattr__(self, name):
        m = _re.fullmatch(r"(bool|i8|u8|i16|u16|i32|u32|i64|u64|f16|f32|f64)([234])", name)
        if m:
            return self.vector(getattr(self, m.group(1)), int(m.group(2)))
        raise AttributeError(name)

    # -- blocks ---------------------------------------------------------------------------------------------------------
    def _new_node(self, ty, instr):
        n = self.keep(NodeS(ty.arc, 0, 0, self.arc(instr)))
        return C.addressof(n)

    def _raw_instr(self, tag):
        ins = self.keep(InstructionS())
        ins.tag = tag
        return ins

    def begin_block(self):
        first = self._new_node(self.void, self._raw_instr(I_INVALID))
        last = self._new_node(self.void, self._raw_instr(I_INVALID))
        NodeS.from_address(first).next = last
        NodeS.from_address(last).prev = first
        bb = self.keep(BasicBlock(first, last))
        self._blocks.append(bb)
        return bb

    def end_block(self):
        return self._blocks.pop()

    def block(self, fn):
        """Builds a basic block from the nodes `fn()` appends; returns (Pooled<BasicBlock> address, fn's result)."""
        bb = self.begin_block()
        r = fn() if fn else None
        self.end_block()
        return C.addressof(bb), r

    def append(self, ty, instr):
        ref = self._new_node(ty, instr)
        bb = self._blocks[-1]
        last = NodeS.from_address(bb.last)
        prev = NodeS.from_address(last.prev)
        me = NodeS.from_address(ref)
        me.prev, me.next = last.prev, bb.last
        prev.next = ref
        last.prev = ref
        return ref

    def detached(self, ty, tag):
        """A node that lives outside every block: kernel arguments, captures, callable parameters."""
        return self._new_node(ty, self._raw_instr(tag))

    # -- instructions (IrBuilder, ir.rs:2330-2650) ------------------------------------------------------------------------
    def call(self, func, args, ret, payload=None):
        ins = self._raw_instr(I_CALL)
        ins.u.call.func.tag = func
        if payload is not None:
            if isinstance(payload, (bytes, str)):
                ins.u.call.func.u.message = self.bytes_slice(payload.encode() if isinstance(payload, str) else payload)
            else:
                ins.u.call.func.u.callable = payload
        ins.u.call.args = self.slice(sz, [v.ref if isinstance(v, Value) else v for v in args])
        return Value(self, self.append(ret, ins), ret)

    def const(self, ty, value):
        ins = self._raw_instr(I_CONST)
        c = ins.u.const_
        if ty.kind == T_PRIMITIVE:
            p = ty.rec.u.primitive
            c.tag = [C_BOOL, C_I8, C_U8, C_I16, C_U16, C_I32, C_U32, C_I64, C_U64, C_F16, C_F32, C_F64][p]
            if p == 9:  # Const::Float16(c_half{bits})
                c.u.u16 = _struct.unpack("<H", _struct.pack("<e", value))[0]
            else:
                setattr(c.u, ["b", "i8", "u8", "i16", "u16", "i32", "u32", "i64", "u64", None, "f32", "f64"][p], value)
        else:
            data = self.pack(ty, value)
            c.tag = C_GENERIC
            c.u.generic.bytes = self.bytes_slice(data)
            c.u.generic.type = ty.arc
        return Value(self, self.append(ty, ins), ty)

    def zero(self, ty):
        ins = self._raw_instr(I_CONST)
        ins.u.const_.tag = C_ZERO
        ins.u.const_.u.type = ty.arc
        return Value(self, self.append(ty, ins), ty)

    def one(self, ty):
        ins = self._raw_instr(I_CONST)
        ins.u.const_.tag = C_ONE
        ins.u.const_.u.type = ty.arc
        return Value(self, self.append(ty, ins), ty)

    def pack(self, ty, value):
        """Host bytes of a constant of type `ty` (Const::Generic payload)."""
        if ty.kind == T_PRIMITIVE:
            return _struct.pack("<" + PRIM_FMT[ty.rec.u.primitive], value)
        if ty.kind == T_VECTOR:
            return b"".join(self.pack(ty.element, v) for v in value).ljust(ty.size, b"\0")
        if ty.kind == T_MATRIX:
            col = self.vector(self.f32, ty.length)
            return b"".join(self.pack(col, c) for c in value)
        if ty.kind == T_ARRAY:
            return b"".join(self.pack(ty.element, v) for v in value)
        if ty.kind == T_STRUCT:
            out = b""
            for f, v in zip(ty.fields, value):
                out = out.ljust((len(out) + f.align - 1) // f.align * f.align, b"\0") + self.pack(f, v)
            return out.ljust(ty.size, b"\0")
        raise TypeError(ty)

    def local(self, init):
        ins = self._raw_instr(I_LOCAL)
        ins.u.local.init = init.ref
        return Var(self, self.append(init.ty, ins), init.ty)

    def local_zero(self, ty):
        return self.local(self.zero(ty))

    def update(self, var, value):
        ins = self._raw_instr(I_UPDATE)
        ins.u.update.var, ins.u.update.value = var.ref, value.ref
        self.append(self.void, ins)

    def phi(self, ty, incomings):
        """incomings: [(value, block address)]"""
        ins = self._raw_instr(I_PHI)
        ins.u.phi = self.slice(PhiIncoming, [PhiIncoming(v.ref, b) for v, b in incomings])
        return Value(self, self.append(ty, ins), ty)

    def if_(self, cond, then_fn, else_fn=None):
        tb, tr = self.block(then_fn)
        fb, fr = self.block(else_fn)
        ins = self._raw_instr(I_IF)
        ins.u.if_.cond, ins.u.if_.true_branch, ins.u.if_.false_branch = cond.ref, tb, fb
        self.append(self.void, ins)
        return (tb, tr), (fb, fr)

    def if_phi(self, cond, then_fn, else_fn):
        """`if cond { a } else { b }` as an expression: both closures return a Value; the result is a phi (control_flow.rs:100-170)."""
        (tb, tv), (fb, fv) = self.if_(cond, then_fn, else_fn)
        return self.phi(tv.ty, [(tv, tb), (fv, fb)])

    def generic_loop(self, prepare_fn, body_fn, update_fn=None):
        """prepare_fn returns the loop condition (a bool Value computed inside the prepare block)."""
        pb, cond = self.block(prepare_fn)
        bb, _ = self.block(body_fn)
        ub, _ = self.block(update_fn)
        ins = self._raw_instr(I_GENERIC_LOOP)
        g = ins.u.generic_loop
        g.prepare, g.cond, g.body, g.update = pb, cond.ref, bb, ub
        self.append(self.void, ins)

    def loop(self, body_fn):
        """do { body } while (cond): body_fn returns the condition."""
        bb, cond = self.block(body_fn)
        ins = self._raw_instr(I_LOOP)
        ins.u.loop.body, ins.u.loop.cond = bb, cond.ref
        self.append(self.void, ins)

    def switch(self, value, cases, default_fn=None):
        blocks = [(v, self.block(fn)[0]) for v, fn in cases]
        db, _ = self.block(default_fn)
        ins = self._raw_instr(I_SWITCH)
        ins.u.switch_.value, ins.u.switch_.default_ = value.ref, db
        ins.u.switch_.cases = self.slice(SwitchCase, [SwitchCase(v, b) for v, b in blocks])
        self.append(self.void, ins)

    def ray_query(self, rq, on_triangle_fn, on_procedural_fn=None):
        """Instruction::RayQuery (ir.rs:1251-1255): traverse; the two blocks are the candidate callbacks."""
        tb, _ = self.block(on_triangle_fn)
        pb, _ = self.block(on_procedural_fn)
        ins = self._raw_instr(I_RAYQUERY)
        ins.u.ray_query.ray_query, ins.u.ray_query.on_triangle_hit, ins.u.ray_query.on_procedural_hit = rq.ref, tb, pb
        self.append(self.void, ins)

    def break_(self):
        self.append(self.void, self._raw_instr(I_BREAK))

    def continue_(self):
        self.append(self.void, self._raw_instr(I_CONTINUE))

    def return_(self, value=None):
        ins = self._raw_instr(I_RETURN)
        ins.u.return_ = value.ref if value is not None else 0
        self.append(self.void, ins)

    def comment(self, text):
        ins = self._raw_instr(I_COMMENT)
        ins.u.comment = self.bytes_slice(text.encode())
        self.append(self.void, ins)

    # -- sugar --------------------------------------------------------------------------------------------------------------
    def lit(self, ty, v):
        return v if isinstance(v, Value) else self.const(ty, v)

    def f(self, v):
        return self.const(self.f32, float(v))

    def u(self, v):
        return self.const(self.u32, int(v))

    def i(self, v):
        return self.const(self.i32, int(v))

    def b(self, v):
        return self.const(self.bool, bool(v))

    def vec(self, ty, *comps):
        """Func::Vec2/3/4 from scalars (or a splat, Func::Vec, from one)."""
        comps = [self.lit(ty.element, c) for c in comps]
        if len(comps) == 1:
            return self.call(Func.Vec, comps, ty)
        assert len(comps) == ty.length
        return self.call([None, None, Func.Vec2, Func.Vec3, Func.Vec4][ty.length], comps, ty)

    def make_struct(self, ty, *fields):
        return self.call(Func.Struct, list(fields), ty)

    def dispatch_id(self):
        return self.call(Func.DispatchId, [], self.u323)

    def thread_id(self):
        return self.call(Func.ThreadId, [], self.u323)

    def block_id(self):
        return self.call(Func.BlockId, [], self.u323)

    def dispatch_size(self):
        return self.call(Func.DispatchSize, [], self.u323)

    def min(self, a, b):
        return self.call(Func.Min, [a, b], a.ty)

    def max(self, a, b):
        return self.call(Func.Max, [a, b], a.ty)


def _bool_of(m, ty):
    return m.vector(m.bool, ty.length) if ty.kind == T_VECTOR else m.bool


class Value:
    """An SSA value (`Expr<T>`): operators append Call nodes to the module's current block."""

    def __init__(self, m, ref, ty):
        self.m, self.ref, self.ty = m, ref, ty

    def _coerce(self, o):
        if isinstance(o, Value):
            return o
        base = self.ty.element if self.ty.kind == T_VECTOR else self.ty
        return self.m.const(base, o)

    def _bin(self, func, o, swap=False, cmp=False):
        o = self._coerce(o)
        a, b = (o, self) if swap else (self, o)
        wide = a.ty if (a.ty.kind in (T_VECTOR, T_MATRIX) or b.ty.kind == T_PRIMITIVE) else b.ty
        if a.ty.kind == T_MATRIX and b.ty.kind == T_VECTOR and func == Func.Mul:
            wide = b.ty
        ret = _bool_of(self.m, wide) if cmp else wide
        return self.m.call(func, [a, b], ret)

    def __add__(self, o): return self._bin(Func.Add, o)
    def __radd__(self, o): return self._bin(Func.Add, o, True)
    def __sub__(self, o): return self._bin(Func.Sub, o)
    def __rsub__(self, o): return self._bin(Func.Sub, o, True)
    def __mul__(self, o): return self._bin(Func.Mul, o)
    def __rmul__(self, o): return self._bin(Func.Mul, o, True)
    def __truediv__(self, o): return self._bin(Func.Div, o)
    def __rtruediv__(self, o): return self._bin(Func.Div, o, True)
    def __mod__(self, o): return self._bin(Func.Rem, o)
    def __and__(self, o): return self._bin(Func.BitAnd, o)
    def __or__(self, o): return self._bin(Func.BitOr, o)
    def __xor__(self, o): return self._bin(Func.BitXor, o)
    def __lshift__(self, o): return self._bin(Func.Shl, o)
    def __rshift__(self, o): return self._bin(Func.Shr, o)
    def __neg__(self): return self.m.call(Func.Neg, [self], self.ty)
    def __invert__(self): return self.m.call(Func.BitNot, [self], self.ty)
    def not_(self): return self.m.call(Func.Not, [self], self.ty)
    def eq(self, o): return self._bin(Func.Eq, o, cmp=True)
    def ne(self, o): return self._bin(Func.Ne, o, cmp=True)
    def lt(self, o): return self._bin(Func.Lt, o, cmp=True)
    def le(self, o): return self._bin(Func.Le, o, cmp=True)
    def gt(self, o): return self._bin(Func.Gt, o, cmp=True)
    def ge(self, o): return self._bin(Func.Ge, o, cmp=True)

    def unary(self, func, ret=None):
        return self.m.call(func, [self], ret or self.ty)

    def cast(self, ty):
        return self.m.call(Func.Cast, [self], ty)

    def bitcast(self, ty):
        return self.m.call(Func.Bitcast, [self], ty)

    def select(self, t, f):
        """self is the condition: Func::Select(cond, t, f)"""
        return self.m.call(Func.Select, [self, t, f], t.ty)

    def extract(self, *idx):
        ty = self.ty
        refs = []
        for i in idx:
            if ty.kind == T_STRUCT:
                refs.append(self.m.const(self.m.u32, i)); ty = ty.fields[i]
            elif ty.kind == T_MATRIX:
                refs.append(self.m.lit(self.m.u32, i)); ty = self.m.vector(self.m.f32, ty.length)
            else:
                refs.append(self.m.lit(self.m.u32, i)); ty = ty.element
        return self.m.call(Func.ExtractElement, [self] + refs, ty)

    def insert(self, value, *idx):
        refs = [self.m.lit(self.m.u32, i) for i in idx]
        return self.m.call(Func.InsertElement, [self, value] + refs, self.ty)

    def permute(self, *idx):
        ty = self.m.vector(self.ty.element, len(idx))
        return self.m.call(Func.Permute, [self] + [self.m.const(self.m.u32, i) for i in idx], ty)

    x = property(lambda s: s.extract(0))
    y = property(lambda s: s.extract(1))
    z = property(lambda s: s.extract(2))
    w = property(lambda s: s.extract(3))

    def dot(self, o): return self.m.call(Func.Dot, [self, o], self.ty.element)
    def cross(self, o): return self.m.call(Func.Cross, [self, o], self.ty)
    def normalize(self): return self.unary(Func.Normalize)
    def length(self): return self.unary(Func.Length, self.ty.element)
    def length_squared(self): return self.unary(Func.LengthSquared, self.ty.element)
    def sqrt(self): return self.unary(Func.Sqrt)
    def abs(self): return self.unary(Func.Abs)
    def sin(self): return self.unary(Func.Sin)
    def cos(self): return self.unary(Func.Cos)
    def floor(self): return self.unary(Func.Floor)
    def is_nan(self): return self.unary(Func.IsNan, _bool_of(self.m, self.ty))
    def any(self): return self.unary(Func.Any, self.m.bool)
    def all(self): return self.unary(Func.All, self.m.bool)
    def clamp(self, lo, hi): return self.m.call(Func.Clamp, [self, lo, hi], self.ty)
    def fma(self, b, c): return self.m.call(Func.Fma, [self, b, c], self.ty)


class Var(Value):
    """A mutable local (`Var<T>`): the node is the address; reads are Func::Load, writes Instruction::Update."""

    def load(self):
        return self.m.call(Func.Load, [self], self.ty)

    def store(self, v):
        self.m.update(self, self._coerce(v) if not isinstance(v, Value) else v)

    def gep(self, *idx):
        ty = self.ty
        refs = []
        for i in idx:
            if ty.kind == T_STRUCT:
                refs.append(self.m.const(self.m.u32, i)); ty = ty.fields[i]
            else:
                refs.append(self.m.lit(self.m.u32, i)); ty = ty.element
        v = self.m.call(Func.GetElementPtr, [self] + refs, ty)
        return Var(self.m, v.ref, ty)


# --------------------------------------------------------------------------------------------------------------------------
# kernels and callables
# --------------------------------------------------------------------------------------------------------------------------
class Resource(Value):
    """A buffer / texture / bindless array / accel parameter node; `ty` is the element (texel) type where one exists."""

    def read(self, index):
        m = self.m
        return m.call(Func.BufferRead, [self, m.lit(m.u32, index)], self.ty)

    def write(self, index, value):
        m = self.m
        m.call(Func.BufferWrite, [self, m.lit(m.u32, index), value], m.void)

    def size(self):
        return self.m.call(Func.BufferSize, [self], self.m.u32)

    def tex_read(self, coord):
        return self.m.call(Func.Texture2dRead, [self, coord], self.ty)

    def tex_write(self, coord, value):
        self.m.call(Func.Texture2dWrite, [self, coord, value], self.m.void)

    def bindless_buffer_read(self, slot, index, ty):
        m = self.m
        return m.call(Func.BindlessBufferRead, [self, m.lit(m.u32, slot), m.lit(m.u32, index)], ty)

    def trace_closest(self, ray, mask, hit_ty):
        return self.m.call(Func.RayTracingTraceClosest, [self, ray, self.m.lit(self.m.u32, mask)], hit_ty)

    def trace_any(self, ray, mask):
        return self.m.call(Func.RayTracingTraceAny, [self, ray, self.m.lit(self.m.u32, mask)], self.m.bool)

    def query(self, ray, mask, any=False):
        """`accel.traverse(ray, opts)` / `traverse_any` (rtx.rs:871-902): the RayQuery object."""
        m = self.m
        return m.call(Func.RayTracingQueryAny if any else Func.RayTracingQueryAll, [self, ray, m.lit(m.u32, mask)], m.opaque("LC_RayQueryAny" if any else "LC_RayQueryAll"))

    def atomic_fetch_add(self, index, value):
        m = self.m
        return m.call(Func.AtomicFetchAdd, [self, m.lit(m.u32, index), value], value.ty)


class KernelBuilder(Module):
    """Builds one ir::KernelModule.  Usage:
        k = KernelBuilder(block_size=(16, 16, 1))
        img = k.arg_tex2d(k.f324); n = k.arg_uniform(k.u32); heap = k.capture_bindless(handle)
        k.body(lambda: ...)          # appends nodes
        km = k.finish()              # address of the KernelModule (LCKernelModule.ptr)
    """

    def __init__(self, block_size=(64, 1, 1), curve_bases=0):
        super().__init__()
        self.block_size = block_size
        self.curve_bases = curve_bases  # CurveBasisSet bits (AccelTraceOptions::curve_bases, rtx.rs:659-670): 1 linear, 2 B-spline, 4 Catmull-Rom, 8 Bezier
        self._args, self._captures, self._shared = [], [], []
        self._entry = None

    def _arg(self, tag, ty):
        r = Resource(self, self.detached(ty, tag), ty)
        self._args.append(r.ref)
        return r

    def arg_buffer(self, elem_ty): return self._arg(I_BUFFER, elem_ty)
    def arg_tex2d(self, texel_ty): return self._arg(I_TEX2D, texel_ty)
    def arg_tex3d(self, texel_ty): return self._arg(I_TEX3D, texel_ty)
    def arg_bindless(self): return self._arg(I_BINDLESS, self.void)
    def arg_accel(self): return self._arg(I_ACCEL, self.void)

    def arg_uniform(self, ty):
        v = Value(self, self.detached(ty, I_UNIFORM), ty)
        self._args.append(v.ref)
        return v

    def _capture(self, tag, ty, binding):
        r = Resource(self, self.detached(ty, tag), ty)
        self._captures.append(Capture(r.ref, binding))
        return r

    def capture_buffer(self, elem_ty, handle, offset, size):
        b = Binding(); b.tag = B_BUFFER; b.u.buffer = BufferBinding(handle, offset, size)
        return self._capture(I_BUFFER, elem_ty, b)

    def capture_tex2d(self, texel_ty, handle, level=0):
        b = Binding(); b.tag = B_TEXTURE; b.u.texture = TextureBinding(handle, level)
        return self._capture(I_TEX2D, texel_ty, b)

    def capture_bindless(self, handle):
        b = Binding(); b.tag = B_BINDLESS; b.u.bindless_array = handle
        return self._capture(I_BINDLESS, self.void, b)

    def capture_accel(self, handle):
        b = Binding(); b.tag = B_ACCEL; b.u.accel = handle
        return self._capture(I_ACCEL, self.void, b)

    def shared(self, array_ty):
        v = Var(self, self.detached(array_ty, I_SHARED), array_ty)
        self._shared.append(v.ref)
        return v

    def body(self, fn):
        self._entry, _ = self.block(fn)

    def finish(self):
        assert self._entry is not None, "call body() first"
        km = self.keep(KernelModuleS())
        km.module.kind = 2
        km.module.entry = self._entry
        km.module.flags = 0
        km.module.curve_basis_set = self.curve_bases
        km.captures = self.slice(Capture, self._captures)
        km.args = self.slice(sz, self._args)
        km.shared = self.slice(sz, self._shared)
        km.cpu_custom_ops = Slice(None, 0, None)
        km.block_size = (C.c_uint32 * 3)(*self.block_size)
        self.km = km
        return C.addressof(km)

    # -- callables (Callable::new_static, lang/functions.rs) -------------------------------------------------------------------
    def callable(self, params, ret_ty, fn):
        """params: [(type, by_value)] — by-reference parameters arrive as `Var`s.  fn(*params) must end with k.return_(...).
        Returns a function that emits the Func::Callable call."""
        nodes = []
        for ty, by_value in params:
            ins = self._raw_instr(I_ARGUMENT)
            ins.u.argument.by_value = by_value
            ref = self._new_node(ty, ins)
            nodes.append((Value if by_value else Var)(self, ref, ty))
        entry, _ = self.block(lambda: fn(*nodes))
        cm = self.keep(CallableModuleS())
        cm.module.kind = 1
        cm.module.entry = entry
        cm.ret_type = ret_ty.arc
        cm.args = self.slice(sz, [n.ref for n in nodes])
        cm.captures = Slice(None, 0, None)
        cm.cpu_custom_ops = Slice(None, 0, None)
        arc = self.arc(cm)

        def invoke(*args):
            return self.call(Func.Callable, list(args), ret_ty, payload=arc)
        return invoke
