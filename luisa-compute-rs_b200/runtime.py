"""Host-side mirror of the reference frontend's runtime objects for the ray-tracing path.

Names and argument meaning follow luisa_compute/src/runtime.rs and src/lib.rs:
`Context::new` / `create_device` (lib.rs:159-221), `Device::{create_buffer, create_mesh,
create_accel, create_stream, default_stream}` (runtime.rs:313-797), `Stream::scope().submit`
(runtime.rs:1001-1075), `BufferView::{copy_from_async, copy_to_async}`.  Every operation is
expressed as the same `api::Command` the Rust frontend would emit and goes through
`DeviceInterface.dispatch`; nothing here computes on the CPU.
"""
import ctypes as C
import threading

import numpy as np

from . import _abi as abi


class LuisaError(RuntimeError):
    pass


class Context:
    """`Context::new(lib_path)` (lib.rs:163-188): loads the backend library and creates a context."""

    def __init__(self, lib_path=None):
        self.lib = abi.load_library(lib_path)
        self.iface = self.lib.luisa_compute_lib_interface()
        self._logger = None
        self.handle = self.iface.create_context(b".")

    def set_logger(self, fn):
        """fn(target, level, message) — api_types LoggerMessage (lib.rs:753-766)."""
        def trampoline(msg):
            fn((msg.target or b"").decode(), (msg.level or b"").decode(), (msg.message or b"").decode())
        self._logger = abi.LoggerCallback(trampoline)
        self.iface.set_logger_callback(self._logger)

    def create_device(self, name="b200", config=None):
        return Device(self, name, config)


class Device:
    """`Context::create_device(name)` (lib.rs:199-220): device interface table + default stream."""

    def __init__(self, ctx, name, config=None):
        self.ctx = ctx
        self.lib = ctx.lib
        cfg = None if config is None else config.encode()
        self.iface = ctx.iface.create_device(ctx.handle, name.encode(), cfg)
        self.handle = self.iface.device
        self._default_stream = None
        self._closed = False
        self._type_cache = {}

    # -- resources ---------------------------------------------------------------------------
    def _ir_type(self, stride, align):
        key = (stride, align)
        if key not in self._type_cache:
            self._type_cache[key] = C.c_void_p(self.lib.lc_b200_make_ir_type(stride, align))
        return self._type_cache[key]

    def create_buffer(self, count, stride, align=4, ext_mem=None):
        """`Device::create_buffer::<T>(count)`: `count` elements of `stride` bytes, zero-initialised."""
        info = self.iface.create_buffer(self.handle, self._ir_type(stride, align), count, ext_mem)
        if info.total_size_bytes != count * stride:
            raise LuisaError("backend reported an unexpected buffer size")
        return Buffer(self, info, count, stride)

    def create_byte_buffer(self, size):
        info = self.iface.create_buffer(self.handle, self._ir_type(0, 0), size, None)
        return Buffer(self, info, size, 1)

    def create_buffer_from_array(self, arr):
        """`Device::create_buffer_from_slice`: element = one row of `arr` (or one item of a structured array)."""
        arr = np.ascontiguousarray(arr)
        count = arr.shape[0]
        stride = int(np.prod(arr.shape[1:], dtype=np.int64)) * arr.dtype.itemsize
        buf = self.create_buffer(count, stride, min(stride & -stride, 16) if stride else 4)
        buf.view().copy_from(arr)
        return buf

    def wrap_device_memory(self, ptr, count, stride, align=4):
        """Buffer over externally owned device memory (`ext_mem` of create_buffer), e.g. a torch tensor's data_ptr()."""
        return self.create_buffer(count, stride, align, ext_mem=C.c_void_p(ptr))

    def create_stream(self, tag=0):
        return Stream(self, tag)

    def default_stream(self):
        if self._default_stream is None:
            self._default_stream = Stream(self, 0)
        return self._default_stream

    def create_mesh(self, vertex_view, index_view, option=None):
        from .rtx import Mesh
        return Mesh(self, vertex_view, index_view, option)

    def create_accel(self, option=None):
        from .rtx import Accel
        return Accel(self, option)

    def create_curve(self, basis, cp_view, seg_view, option=None):
        from .rtx import Curve
        return Curve(self, basis, cp_view, seg_view, option)

    def create_procedural_primitive(self, aabb_view, option=None):
        from .rtx import ProceduralPrimitive
        return ProceduralPrimitive(self, aabb_view, option)

    def create_event(self):
        return Event(self)

    def create_tex2d(self, format, width, height):
        """`Device::create_tex2d::<T>(storage, w, h, mips=1)` (runtime.rs:529-570)."""
        return Texture(self, format, 2, width, height, 1)

    def create_tex3d(self, format, width, height, depth):
        return Texture(self, format, 3, width, height, depth)

    def create_bindless_array(self, slots):
        """`Device::create_bindless_array(slots)` (runtime.rs:572-600)."""
        return BindlessArray(self, slots)

    def create_shader(self, kernel_module_ptr, fast_math=False, name=None, keep=None):
        """`Device::create_kernel` -> DeviceInterface::create_shader(LCKernelModule{ptr}, &ShaderOption) (runtime.rs:1226-1260).
        `kernel_module_ptr`: address of an ir::KernelModule (ir.KernelBuilder.finish()); `keep`: the builder that owns it."""
        return Shader(self, kernel_module_ptr, fast_math, name, keep)

    def query(self, name):
        p = self.iface.query(self.handle, name.encode())
        if not p:
            return None
        s = C.string_at(p).decode()
        self.ctx.iface.free_string(p)
        return s or None

    def name(self):
        return self.query("device_name")

    def ordinal(self):
        return self.lib.lc_b200_device_ordinal(self.handle)

    def close(self):
        if self._closed:
            return
        self._closed = True
        if self._default_stream is not None:
            self._default_stream.destroy()
            self._default_stream = None
        self.iface.destroy_device(self.iface)


class Buffer:
    def __init__(self, device, info, count, stride):
        self.device = device
        self.handle = abi.Handle(info.resource.handle)
        self.native = info.resource.native_handle
        self.count = count
        self.stride = stride
        self.size_bytes = info.total_size_bytes
        self._alive = True

    def view(self, start=0, count=None):
        """`Buffer::view(range)` in elements."""
        if count is None:
            count = self.count - start
        if start < 0 or start + count > self.count:
            raise LuisaError("buffer view out of range")
        return BufferView(self, start * self.stride, count * self.stride)

    def device_ptr(self):
        return self.device.lib.lc_b200_buffer_native(self.device.handle, self.handle)

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_buffer(self.device.handle, self.handle)
        self._alive = False


class BufferView:
    def __init__(self, buffer, offset, size):
        self.buffer = buffer
        self.offset = offset
        self.size = size

    def copy_from_async(self, arr):
        """`BufferView::copy_from_async(&[T])` -> api::Command::BufferUpload."""
        arr = np.ascontiguousarray(arr)
        if arr.nbytes != self.size:
            raise LuisaError(f"copy_from: {arr.nbytes} bytes into a {self.size}-byte view")
        cmd = abi.Command()
        cmd.tag = abi.CMD_BUFFER_UPLOAD
        cmd.u.buffer_upload = abi.CmdBufferUpload(self.buffer.handle, self.offset, self.size, arr.ctypes.data)
        return HostCommand(cmd, keep=[arr, self.buffer])

    def copy_to_async(self, arr):
        """`BufferView::copy_to_async(&mut [T])` -> api::Command::BufferDownload."""
        if not arr.flags["C_CONTIGUOUS"] or not arr.flags["WRITEABLE"]:
            raise LuisaError("copy_to needs a writable C-contiguous array")
        if arr.nbytes != self.size:
            raise LuisaError(f"copy_to: {self.size}-byte view into {arr.nbytes} bytes")
        cmd = abi.Command()
        cmd.tag = abi.CMD_BUFFER_DOWNLOAD
        cmd.u.buffer_download = abi.CmdBufferDownload(self.buffer.handle, self.offset, self.size, arr.ctypes.data)
        return HostCommand(cmd, keep=[arr, self.buffer])

    def copy_to_buffer_async(self, dst_view):
        cmd = abi.Command()
        cmd.tag = abi.CMD_BUFFER_COPY
        cmd.u.buffer_copy = abi.CmdBufferCopy(self.buffer.handle, self.offset, dst_view.buffer.handle, dst_view.offset, self.size)
        return HostCommand(cmd, keep=[self.buffer, dst_view.buffer])

    def copy_from(self, arr):
        s = self.buffer.device.default_stream()
        s.submit([self.copy_from_async(arr)])
        s.synchronize()

    def copy_to(self, arr):
        s = self.buffer.device.default_stream()
        s.submit([self.copy_to_async(arr)])
        s.synchronize()

    def to_numpy(self, dtype):
        out = np.empty(self.size // np.dtype(dtype).itemsize, dtype=dtype)
        self.copy_to(out)
        return out


PIXEL_FORMATS = {  # api_types PixelFormat discriminants (lib.rs:310-362) -> (numpy dtype, channels)
    "R8Uint": (1, np.uint8, 1), "R8Unorm": (2, np.uint8, 1), "Rg8Unorm": (5, np.uint8, 2), "Rgba8Uint": (7, np.uint8, 4), "Rgba8Unorm": (8, np.uint8, 4),
    "R16Uint": (10, np.uint16, 1), "R16Unorm": (11, np.uint16, 1), "Rgba16Unorm": (17, np.uint16, 4),
    "R32Sint": (18, np.int32, 1), "R32Uint": (19, np.uint32, 1), "Rg32Uint": (21, np.uint32, 2), "Rgba32Sint": (22, np.int32, 4), "Rgba32Uint": (23, np.uint32, 4),
    "R16f": (24, np.float16, 1), "Rgba16f": (26, np.float16, 4), "R32f": (27, np.float32, 1), "Rg32f": (28, np.float32, 2), "Rgba32f": (29, np.float32, 4),
}
_FORMAT_TO_STORAGE = [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 10, 11, 12, 13, 14]


class Texture:
    """`Tex2d<T>` / `Tex3d<T>` (runtime.rs:529-570): one mip level, row-major texels on upload / download."""

    def __init__(self, device, format, dim, width, height, depth):
        self.device = device
        self.format_id, self.dtype, self.channels = PIXEL_FORMATS[format]
        self.storage = _FORMAT_TO_STORAGE[self.format_id]
        self.dim, self.width, self.height, self.depth = dim, width, height, depth
        info = device.iface.create_texture(device.handle, self.format_id, dim, width, height, depth, 1, False, False)
        self.handle = abi.Handle(info.handle)
        self._alive = True

    @property
    def shape(self):
        return ((self.depth,) if self.dim == 3 else ()) + (self.height, self.width) + ((self.channels,) if self.channels > 1 else ())

    def _transfer(self, tag, arr):
        if arr.dtype != self.dtype or arr.shape != self.shape or not arr.flags["C_CONTIGUOUS"]:
            raise LuisaError(f"texture transfer needs a C-contiguous {np.dtype(self.dtype).name} array of shape {self.shape}")
        cmd = abi.Command()
        cmd.tag = tag
        t = abi.CmdTextureTransfer(self.handle, self.storage, 0, (C.c_uint32 * 3)(self.width, self.height, self.depth), arr.ctypes.data)
        if tag == abi.CMD_TEXTURE_UPLOAD:
            cmd.u.texture_upload = t
        else:
            cmd.u.texture_download = t
        return HostCommand(cmd, keep=[arr, self])

    def copy_from_async(self, arr):
        return self._transfer(abi.CMD_TEXTURE_UPLOAD, np.ascontiguousarray(arr))

    def copy_to_async(self, arr):
        return self._transfer(abi.CMD_TEXTURE_DOWNLOAD, arr)

    def copy_from(self, arr):
        s = self.device.default_stream()
        s.submit([self.copy_from_async(arr)])
        s.synchronize()

    def to_numpy(self):
        out = np.empty(self.shape, dtype=self.dtype)
        s = self.device.default_stream()
        s.submit([self.copy_to_async(out)])
        s.synchronize()
        return out

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_texture(self.device.handle, self.handle)
        self._alive = False


class BindlessArray:
    """`BindlessArray` (runtime.rs:572-700): emplace_*_async queue modifications, update_async emits the command."""

    def __init__(self, device, slots):
        self.device = device
        self.slots = slots
        self.handle = abi.Handle(device.iface.create_bindless_array(device.handle, slots).handle)
        self._mods = {}
        self._keep = {}
        self._alive = True

    def _mod(self, slot):
        if slot not in self._mods:
            m = abi.BindlessModification()
            m.slot = slot
            self._mods[slot] = m
        return self._mods[slot]

    def emplace_buffer_async(self, slot, buffer, offset=0):
        m = self._mod(slot)
        m.buffer = abi.BindlessBufferUpdate(abi.BINDLESS_EMPLACE, buffer.handle, offset)
        self._keep[("b", slot)] = buffer

    def emplace_tex2d_async(self, slot, texture, filter=0, address=0):
        """`emplace_tex2d_async(slot, tex, Sampler{filter, address})`: filter 0 point / 1-3 linear, address 0 edge / 1 repeat / 2 mirror / 3 zero."""
        m = self._mod(slot)
        m.tex2d = abi.BindlessTextureUpdate(abi.BINDLESS_EMPLACE, texture.handle, abi.Sampler(filter, address))
        self._keep[("t2", slot)] = texture

    def emplace_tex3d_async(self, slot, texture, filter=0, address=0):
        m = self._mod(slot)
        m.tex3d = abi.BindlessTextureUpdate(abi.BINDLESS_EMPLACE, texture.handle, abi.Sampler(filter, address))
        self._keep[("t3", slot)] = texture

    def remove_buffer_async(self, slot):
        self._mod(slot).buffer = abi.BindlessBufferUpdate(abi.BINDLESS_REMOVE, abi.Handle(0xFFFFFFFFFFFFFFFF), 0)
        self._keep.pop(("b", slot), None)

    def update_async(self):
        mods = list(self._mods.values())
        self._mods = {}
        arr = (abi.BindlessModification * max(len(mods), 1))(*mods)
        cmd = abi.Command()
        cmd.tag = abi.CMD_BINDLESS_UPDATE
        cmd.u.bindless_update = abi.CmdBindlessUpdate(self.handle, arr, len(mods))
        return HostCommand(cmd, keep=[arr, self])

    def update(self):
        s = self.device.default_stream()
        s.submit([self.update_async()])
        s.synchronize()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_bindless_array(self.device.handle, self.handle)
        self._alive = False


class Shader:
    """A compiled kernel (`RawKernel`, runtime.rs:1262-1330): `dispatch_async(dispatch_size, *args)` emits api::Command::ShaderDispatch
    with one Argument per kernel argument, in order (KernelArg::encode, runtime.rs:1332-1500)."""

    def __init__(self, device, kernel_module_ptr, fast_math=False, name=None, keep=None):
        self.device = device
        self._keep = keep
        opt = abi.ShaderOption(True, fast_math, False, False, False, 0, name.encode() if name else None, None)
        info = device.iface.create_shader(device.handle, abi.KernelModule(kernel_module_ptr), C.byref(opt))
        self.handle = abi.Handle(info.resource.handle)
        self.block_size = tuple(info.block_size)
        self._alive = True

    def dispatch_async(self, dispatch_size, *args):
        from .rtx import Accel
        n = len(args)
        arr = (abi.Argument * max(n, 1))()
        keep = [arr, self]
        for i, a in enumerate(args):
            if isinstance(a, BufferView):
                arr[i].tag = abi.ARG_BUFFER
                arr[i].u.buffer = abi._ArgBuffer(a.buffer.handle, a.offset, a.size)
            elif isinstance(a, Buffer):
                arr[i].tag = abi.ARG_BUFFER
                arr[i].u.buffer = abi._ArgBuffer(a.handle, 0, a.size_bytes)
            elif isinstance(a, Texture):
                arr[i].tag = abi.ARG_TEXTURE
                arr[i].u.texture = abi._ArgTexture(a.handle, 0)
            elif isinstance(a, BindlessArray):
                arr[i].tag = abi.ARG_BINDLESS
                arr[i].u.bindless = a.handle
            elif isinstance(a, Accel):
                arr[i].tag = abi.ARG_ACCEL
                arr[i].u.accel = a.handle
            else:  # uniform: a numpy scalar / array / bytes holding the value in the IR type's layout
                data = np.ascontiguousarray(a) if not isinstance(a, (bytes, bytearray)) else np.frombuffer(bytes(a), dtype=np.uint8)
                arr[i].tag = abi.ARG_UNIFORM
                arr[i].u.uniform = abi._ArgUniform(data.ctypes.data, data.nbytes)
                keep.append(data)
                continue
            keep.append(a)
        ds = tuple(dispatch_size) + (1,) * (3 - len(dispatch_size))
        cmd = abi.Command()
        cmd.tag = abi.CMD_SHADER_DISPATCH
        cmd.u.shader_dispatch = abi.CmdShaderDispatch(self.handle, (C.c_uint32 * 3)(*ds), arr, n)
        return HostCommand(cmd, keep=keep)

    def dispatch(self, dispatch_size, *args):
        s = self.device.default_stream()
        s.submit([self.dispatch_async(dispatch_size, *args)])
        s.synchronize()

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_shader(self.device.handle, self.handle)
        self._alive = False


class HostCommand:
    """A frontend `Command`: the api::Command plus the resources it keeps alive (ResourceTracker, lib.rs:223-270)."""

    def __init__(self, cmd, keep=()):
        self.cmd = cmd
        self.keep = list(keep)


class Stream:
    """`Stream` + `Scope::submit` (runtime.rs:1001-1075): in-order; completion callback fires once per submit."""

    def __init__(self, device, tag=0):
        self.device = device
        info = device.iface.create_stream(device.handle, tag)
        self.handle = abi.Handle(info.handle)
        self.native = info.native_handle
        self._lock = threading.Lock()
        self._pending = {}
        self._next = 1
        self._alive = True

        def on_done(ctx):
            key = C.cast(ctx, C.c_void_p).value
            with self._lock:
                entry = self._pending.pop(key, None)
            if entry and entry[1]:
                entry[1]()
        self._cb = abi.DispatchCallback(on_done)

    def submit(self, commands, callback=None):
        commands = list(commands)
        n = len(commands)
        arr = (abi.Command * max(n, 1))()
        for i, c in enumerate(commands):
            arr[i] = c.cmd
        with self._lock:
            key = self._next
            self._next += 1
            self._pending[key] = (commands, callback)  # CommandCallbackCtx: keeps args alive until the callback
        self.device.iface.dispatch(self.device.handle, self.handle, abi.CommandList(arr, n), self._cb, C.cast(C.c_void_p(key), abi.u8p))

    def synchronize(self):
        self.device.iface.synchronize_stream(self.device.handle, self.handle)

    def cuda_stream(self):
        """The cudaStream_t behind this stream (for torch.cuda.ExternalStream / event timing)."""
        return self.device.lib.lc_b200_stream_native(self.device.handle, self.handle)

    def destroy(self):
        if self._alive and not self.device._closed:
            self.device.iface.destroy_stream(self.device.handle, self.handle)
        self._alive = False


class Event:
    """Timeline event (runtime.rs `Event`; backend semantics cpu/resource.rs:10-44)."""

    def __init__(self, device):
        self.device = device
        self.handle = abi.Handle(device.iface.create_event(device.handle).handle)

    def signal(self, stream, value):
        self.device.iface.signal_event(self.device.handle, self.handle, stream.handle, value)

    def wait(self, stream, value):
        self.device.iface.wait_event(self.device.handle, self.handle, stream.handle, value)

    def synchronize(self, value):
        self.device.iface.synchronize_event(self.device.handle, self.handle, value)

    def is_completed(self, value):
        return bool(self.device.iface.is_event_completed(self.device.handle, self.handle, value))

    def destroy(self):
        self.device.iface.destroy_event(self.device.handle, self.handle)
