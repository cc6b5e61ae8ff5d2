"""ctypes view of include/lc_b200_api.h and the loader of liblc_b200.so.

The Python host layer plays the role of the Rust frontend's `ProxyBackend`
(luisa_compute_backend/src/proxy.rs:25-285): it dlopens the library, calls the single
symbol `luisa_compute_lib_interface` and from then on only goes through the two
function-pointer tables — the same path the unmodified Rust frontend would take.

There is no fallback: if the CUDA library is missing, import fails with instructions
to build it (python -c "import __graft_entry__ as g; g.build()").
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liblc_b200.so")

u8p = C.POINTER(C.c_uint8)


class Handle(C.Structure):
    _fields_ = [("id", C.c_uint64)]


class Created(C.Structure):
    _fields_ = [("handle", C.c_uint64), ("native_handle", C.c_void_p)]


class CreatedBuffer(C.Structure):
    _fields_ = [("resource", Created), ("element_stride", C.c_size_t), ("total_size_bytes", C.c_size_t)]


class CreatedShader(C.Structure):
    _fields_ = [("resource", Created), ("block_size", C.c_uint32 * 3)]


class CreatedSwapchain(C.Structure):
    _fields_ = [("resource", Created), ("storage", C.c_int32)]


class AccelOption(C.Structure):
    """api_types AccelOption (lib.rs:204-220); defaults = FastTrace, compaction on, update off."""
    _fields_ = [("hint", C.c_int32), ("allow_compaction", C.c_bool), ("allow_update", C.c_bool)]

    def __init__(self, hint=0, allow_compaction=True, allow_update=False):
        super().__init__(hint, allow_compaction, allow_update)


class AccelModification(C.Structure):
    _fields_ = [("index", C.c_uint32), ("user_id", C.c_uint32), ("flags", C.c_uint32), ("visibility", C.c_uint32),
                ("mesh", C.c_uint64), ("affine", C.c_float * 12)]


class CmdBufferUpload(C.Structure):
    _fields_ = [("buffer", Handle), ("offset", C.c_size_t), ("size", C.c_size_t), ("data", C.c_void_p)]


class CmdBufferDownload(C.Structure):
    _fields_ = [("buffer", Handle), ("offset", C.c_size_t), ("size", C.c_size_t), ("data", C.c_void_p)]


class CmdBufferCopy(C.Structure):
    _fields_ = [("src", Handle), ("src_offset", C.c_size_t), ("dst", Handle), ("dst_offset", C.c_size_t), ("size", C.c_size_t)]


class CmdMeshBuild(C.Structure):
    _fields_ = [("mesh", Handle), ("request", C.c_int32), ("vertex_buffer", Handle), ("vertex_buffer_offset", C.c_size_t),
                ("vertex_buffer_size", C.c_size_t), ("vertex_stride", C.c_size_t), ("index_buffer", Handle),
                ("index_buffer_offset", C.c_size_t), ("index_buffer_size", C.c_size_t), ("index_stride", C.c_size_t)]


class CmdProceduralBuild(C.Structure):
    """ProceduralPrimitiveBuildCommand (api_types:633-641)"""
    _fields_ = [("handle", Handle), ("request", C.c_int32), ("aabb_buffer", Handle), ("aabb_offset", C.c_size_t), ("aabb_count", C.c_size_t)]


class CmdCurveBuild(C.Structure):
    """CurveBuildCommand (api_types:618-631); basis = CurveBasis (api_types:196-202)"""
    _fields_ = [("curve", Handle), ("request", C.c_int32), ("basis", C.c_int32), ("cp_count", C.c_size_t), ("seg_count", C.c_size_t),
                ("cp_buffer", Handle), ("cp_offset", C.c_size_t), ("cp_stride", C.c_size_t), ("seg_buffer", Handle), ("seg_offset", C.c_size_t)]


class CmdAccelBuild(C.Structure):
    _fields_ = [("accel", Handle), ("request", C.c_int32), ("instance_count", C.c_uint32),
                ("modifications", C.POINTER(AccelModification)), ("modifications_count", C.c_size_t),
                ("update_instance_buffer_only", C.c_bool)]


class CmdBufferTexture(C.Structure):
    """BufferToTextureCopyCommand / TextureToBufferCopyCommand (api_types:516-541)"""
    _fields_ = [("buffer", Handle), ("buffer_offset", C.c_size_t), ("texture", Handle), ("storage", C.c_int32), ("level", C.c_uint32), ("size", C.c_uint32 * 3)]


class CmdTextureTransfer(C.Structure):
    """TextureUploadCommand / TextureDownloadCommand (api_types:543-566)"""
    _fields_ = [("texture", Handle), ("storage", C.c_int32), ("level", C.c_uint32), ("size", C.c_uint32 * 3), ("data", C.c_void_p)]


class CmdTextureCopy(C.Structure):
    _fields_ = [("storage", C.c_int32), ("src", Handle), ("dst", Handle), ("size", C.c_uint32 * 3), ("src_level", C.c_uint32), ("dst_level", C.c_uint32)]


class _ArgBuffer(C.Structure):
    _fields_ = [("buffer", Handle), ("offset", C.c_size_t), ("size", C.c_size_t)]


class _ArgTexture(C.Structure):
    _fields_ = [("texture", Handle), ("level", C.c_uint32)]


class _ArgUniform(C.Structure):
    _fields_ = [("data", C.c_void_p), ("size", C.c_size_t)]


class _ArgUnion(C.Union):
    _fields_ = [("buffer", _ArgBuffer), ("texture", _ArgTexture), ("uniform", _ArgUniform), ("bindless", Handle), ("accel", Handle)]


class Argument(C.Structure):
    """api_types Argument (lib.rs:454-484)"""
    _fields_ = [("tag", C.c_int32), ("u", _ArgUnion)]


ARG_BUFFER, ARG_TEXTURE, ARG_UNIFORM, ARG_BINDLESS, ARG_ACCEL = range(5)


class CmdShaderDispatch(C.Structure):
    _fields_ = [("shader", Handle), ("dispatch_size", C.c_uint32 * 3), ("args", C.POINTER(Argument)), ("args_count", C.c_size_t)]


class Sampler(C.Structure):
    _fields_ = [("filter", C.c_int32), ("address", C.c_int32)]


class BindlessBufferUpdate(C.Structure):
    _fields_ = [("op", C.c_int32), ("handle", Handle), ("offset", C.c_size_t)]


class BindlessTextureUpdate(C.Structure):
    _fields_ = [("op", C.c_int32), ("handle", Handle), ("sampler", Sampler)]


class BindlessModification(C.Structure):
    """BindlessArrayUpdateModification (api_types:697-704)"""
    _fields_ = [("slot", C.c_size_t), ("buffer", BindlessBufferUpdate), ("tex2d", BindlessTextureUpdate), ("tex3d", BindlessTextureUpdate)]


BINDLESS_NONE, BINDLESS_EMPLACE, BINDLESS_REMOVE = 0, 1, 2


class CmdBindlessUpdate(C.Structure):
    _fields_ = [("handle", Handle), ("modifications", C.POINTER(BindlessModification)), ("modifications_count", C.c_size_t)]


class ShaderOption(C.Structure):
    """api_types ShaderOption (lib.rs:828-846)"""
    _fields_ = [("enable_cache", C.c_bool), ("enable_fast_math", C.c_bool), ("enable_debug_info", C.c_bool), ("compile_only", C.c_bool), ("time_trace", C.c_bool),
                ("max_registers", C.c_uint32), ("name", C.c_char_p), ("native_include", C.c_char_p)]


class _CmdUnion(C.Union):
    _fields_ = [("buffer_upload", CmdBufferUpload), ("buffer_download", CmdBufferDownload), ("buffer_copy", CmdBufferCopy),
                ("buffer_to_texture", CmdBufferTexture), ("texture_to_buffer", CmdBufferTexture), ("texture_upload", CmdTextureTransfer),
                ("texture_download", CmdTextureTransfer), ("texture_copy", CmdTextureCopy), ("shader_dispatch", CmdShaderDispatch),
                ("bindless_update", CmdBindlessUpdate), ("procedural_build", CmdProceduralBuild), ("curve_build", CmdCurveBuild),
                ("mesh_build", CmdMeshBuild), ("accel_build", CmdAccelBuild), ("_raw", C.c_uint8 * 80)]


class Command(C.Structure):
    _fields_ = [("tag", C.c_int32), ("u", _CmdUnion)]


assert C.sizeof(Command) == 88 and Command.u.offset == 8
assert C.sizeof(AccelModification) == 72 and C.sizeof(AccelOption) == 8
assert C.sizeof(Argument) == 32 and C.sizeof(BindlessModification) == 80 and C.sizeof(CmdShaderDispatch) == 40

CMD_BUFFER_UPLOAD, CMD_BUFFER_DOWNLOAD, CMD_BUFFER_COPY = 0, 1, 2
CMD_BUFFER_TO_TEXTURE, CMD_TEXTURE_TO_BUFFER, CMD_TEXTURE_DOWNLOAD, CMD_TEXTURE_COPY = 3, 4, 6, 7
CMD_TEXTURE_UPLOAD, CMD_SHADER_DISPATCH, CMD_MESH_BUILD, CMD_CURVE_BUILD, CMD_PROCEDURAL_BUILD, CMD_ACCEL_BUILD, CMD_BINDLESS_UPDATE = 5, 8, 9, 10, 11, 12, 13

MOD_PRIMITIVE, MOD_TRANSFORM, MOD_OPAQUE_ON, MOD_OPAQUE_OFF, MOD_VISIBILITY, MOD_USER_ID = 1, 2, 4, 8, 16, 32


class CommandList(C.Structure):
    _fields_ = [("commands", C.POINTER(Command)), ("commands_count", C.c_size_t)]


class LoggerMessage(C.Structure):
    _fields_ = [("target", C.c_char_p), ("level", C.c_char_p), ("message", C.c_char_p)]


class PinnedMemoryExt(C.Structure):
    _fields_ = [("data", C.c_void_p), ("pin_host_memory", C.c_void_p), ("allocate_pinned_memory", C.c_void_p)]


class DenoiserExt(C.Structure):
    _fields_ = [("data", C.c_void_p), ("create", C.c_void_p), ("init", C.c_void_p), ("execute", C.c_void_p), ("destroy", C.c_void_p)]


class KernelModule(C.Structure):
    _fields_ = [("ptr", C.c_uint64)]


DispatchCallback = C.CFUNCTYPE(None, u8p)
LoggerCallback = C.CFUNCTYPE(None, LoggerMessage)


class DeviceInterface(C.Structure):
    pass


F = C.CFUNCTYPE
DeviceInterface._fields_ = [
    ("device", Handle),
    ("destroy_device", F(None, DeviceInterface)),
    ("create_buffer", F(CreatedBuffer, Handle, C.c_void_p, C.c_size_t, C.c_void_p)),
    ("destroy_buffer", F(None, Handle, Handle)),
    ("create_texture", F(Created, Handle, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_bool, C.c_bool)),
    ("native_handle", F(C.c_void_p, Handle)),
    ("compute_warp_size", F(C.c_uint32, Handle)),
    ("destroy_texture", F(None, Handle, Handle)),
    ("create_bindless_array", F(Created, Handle, C.c_size_t)),
    ("destroy_bindless_array", F(None, Handle, Handle)),
    ("create_stream", F(Created, Handle, C.c_int32)),
    ("destroy_stream", F(None, Handle, Handle)),
    ("synchronize_stream", F(None, Handle, Handle)),
    ("dispatch", F(None, Handle, Handle, CommandList, DispatchCallback, u8p)),
    ("create_swapchain", F(CreatedSwapchain, Handle, C.c_void_p, Handle)),
    ("present_display_in_stream", F(None, Handle, Handle, Handle, Handle)),
    ("destroy_swapchain", F(None, Handle, Handle)),
    ("create_shader", F(CreatedShader, Handle, KernelModule, C.c_void_p)),
    ("destroy_shader", F(None, Handle, Handle)),
    ("create_event", F(Created, Handle)),
    ("destroy_event", F(None, Handle, Handle)),
    ("signal_event", F(None, Handle, Handle, Handle, C.c_uint64)),
    ("synchronize_event", F(None, Handle, Handle, C.c_uint64)),
    ("wait_event", F(None, Handle, Handle, Handle, C.c_uint64)),
    ("is_event_completed", F(C.c_bool, Handle, Handle, C.c_uint64)),
    ("create_mesh", F(Created, Handle, C.POINTER(AccelOption))),
    ("destroy_mesh", F(None, Handle, Handle)),
    ("create_curve", F(Created, Handle, C.POINTER(AccelOption))),
    ("destroy_curve", F(None, Handle, Handle)),
    ("create_procedural_primitive", F(Created, Handle, C.POINTER(AccelOption))),
    ("destroy_procedural_primitive", F(None, Handle, Handle)),
    ("create_accel", F(Created, Handle, C.POINTER(AccelOption))),
    ("destroy_accel", F(None, Handle, Handle)),
    ("query", F(C.c_void_p, Handle, C.c_char_p)),
    ("pinned_memory_ext", F(PinnedMemoryExt, Handle)),
    ("denoiser_ext", F(DenoiserExt, Handle)),
]
assert C.sizeof(DeviceInterface) == 8 + 35 * 8


class LibInterface(C.Structure):
    _fields_ = [
        ("inner", C.c_void_p),
        ("set_logger_callback", F(None, LoggerCallback)),
        ("create_context", F(Handle, C.c_char_p)),
        ("destroy_context", F(None, Handle)),
        ("create_device", F(DeviceInterface, Handle, C.c_char_p, C.c_char_p)),
        ("free_string", F(None, C.c_void_p)),
    ]


class BuildStats(C.Structure):
    _fields_ = [("primitive_count", C.c_uint64), ("wide_node_count", C.c_uint64), ("packed_tri_count", C.c_uint64), ("bvh_bytes", C.c_uint64),
                ("max_depth", C.c_uint32), ("was_refit", C.c_uint32), ("build_ms", C.c_float), ("builder", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("_")}


class TraceCounters(C.Structure):
    _fields_ = [("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("rays", C.c_uint64), ("instance_entries", C.c_uint64)]


# every symbol include/lc_b200_api.h declares with LCB_EXPORT
EXPORTED_SYMBOLS = [
    "luisa_compute_lib_interface", "lc_b200_trace_closest", "lc_b200_trace_any", "lc_b200_trace_closest_host",
    "lc_b200_trace_any_host", "lc_b200_instance_transform", "lc_b200_instance_user_id", "lc_b200_instance_visibility_mask",
    "lc_b200_mesh_stats", "lc_b200_accel_stats", "lc_b200_trace_closest_counted", "lc_b200_stream_native",
    "lc_b200_buffer_native", "lc_b200_device_ordinal", "lc_b200_kernel_launch_count", "lc_b200_version", "lc_b200_make_ir_type",
    "lc_b200_ray_query", "lc_b200_example_path_tracer", "lc_b200_ir_lower_source", "lc_b200_shader_compile_check", "lc_b200_ir_layout_json", "lc_b200_set_builder", "lc_b200_set_lowering",
]


class PathTracerArgs(C.Structure):
    """lcb_path_tracer_args"""
    _fields_ = [("accel", Handle), ("vertex_heap", C.POINTER(Handle)), ("index_heap", C.POINTER(Handle)), ("heap_size", C.c_uint32),
                ("image", Handle), ("seed_image", Handle), ("width", C.c_uint32), ("height", C.c_uint32), ("spp_per_dispatch", C.c_uint32),
                ("max_depth", C.c_uint32), ("tan_half_fov", C.c_float)]


class CandidateFilter(C.Structure):
    """lcb_candidate_filter: the candidate hook of the batch RayQuery entry point."""
    _fields_ = [("kind", C.c_int32), ("radius", C.c_float), ("bits", Handle), ("first_bit", Handle)]


FILTER_COMMIT_ALL, FILTER_BARY_DISC, FILTER_PRIM_BITS, FILTER_REJECT_ALL = 0, 1, 2, 3

_lib = None


def load_library(path=None):
    """dlopen liblc_b200.so and declare the native entry points.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("LC_B200_LIB", LIB_PATH)
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} is missing: the B200 device has no CPU fallback. Build it with "
            "`make -C luisa-compute-rs_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`.")
    lib = C.CDLL(p)
    lib.luisa_compute_lib_interface.restype = LibInterface
    lib.luisa_compute_lib_interface.argtypes = []
    H = Handle
    lib.lc_b200_trace_closest.argtypes = [H, H, H, H, C.c_size_t, H, C.c_size_t, C.c_uint64, C.c_uint32]
    lib.lc_b200_trace_closest.restype = None
    lib.lc_b200_trace_any.argtypes = [H, H, H, H, C.c_size_t, H, C.c_size_t, C.c_uint64, C.c_uint32]
    lib.lc_b200_trace_any.restype = None
    lib.lc_b200_ray_query.argtypes = [H, H, H, H, C.c_size_t, H, C.c_size_t, C.c_uint64, C.c_uint32, C.c_bool, C.POINTER(CandidateFilter)]
    lib.lc_b200_ray_query.restype = None
    lib.lc_b200_example_path_tracer.argtypes = [H, H, C.POINTER(PathTracerArgs), C.POINTER(C.c_uint64)]
    lib.lc_b200_example_path_tracer.restype = None
    lib.lc_b200_trace_closest_counted.argtypes = [H, H, H, H, C.c_size_t, H, C.c_size_t, C.c_uint64, C.c_uint32, C.POINTER(TraceCounters)]
    lib.lc_b200_trace_closest_counted.restype = None
    lib.lc_b200_trace_closest_host.argtypes = [H, H, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
    lib.lc_b200_trace_closest_host.restype = None
    lib.lc_b200_trace_any_host.argtypes = [H, H, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32]
    lib.lc_b200_trace_any_host.restype = None
    lib.lc_b200_instance_transform.argtypes = [H, H, C.c_uint32, C.POINTER(C.c_float)]
    lib.lc_b200_instance_transform.restype = None
    lib.lc_b200_instance_user_id.argtypes = [H, H, C.c_uint32]
    lib.lc_b200_instance_user_id.restype = C.c_uint32
    lib.lc_b200_instance_visibility_mask.argtypes = [H, H, C.c_uint32]
    lib.lc_b200_instance_visibility_mask.restype = C.c_uint32
    lib.lc_b200_mesh_stats.argtypes = [H, H, C.POINTER(BuildStats)]
    lib.lc_b200_mesh_stats.restype = None
    lib.lc_b200_accel_stats.argtypes = [H, H, C.POINTER(BuildStats)]
    lib.lc_b200_accel_stats.restype = None
    lib.lc_b200_stream_native.argtypes = [H, H]
    lib.lc_b200_stream_native.restype = C.c_void_p
    lib.lc_b200_buffer_native.argtypes = [H, H]
    lib.lc_b200_buffer_native.restype = C.c_void_p
    lib.lc_b200_device_ordinal.argtypes = [H]
    lib.lc_b200_device_ordinal.restype = C.c_int
    lib.lc_b200_kernel_launch_count.argtypes = []
    lib.lc_b200_kernel_launch_count.restype = C.c_uint64
    lib.lc_b200_version.argtypes = []
    lib.lc_b200_version.restype = C.c_char_p
    lib.lc_b200_make_ir_type.argtypes = [C.c_size_t, C.c_size_t]
    lib.lc_b200_make_ir_type.restype = C.c_void_p
    lib.lc_b200_ir_lower_source.argtypes = [C.c_void_p]
    lib.lc_b200_ir_lower_source.restype = C.c_void_p
    lib.lc_b200_shader_compile_check.argtypes = [C.c_void_p, C.c_bool, C.POINTER(C.c_void_p)]
    lib.lc_b200_shader_compile_check.restype = C.c_int
    lib.lc_b200_set_builder.argtypes = [C.c_int]
    lib.lc_b200_set_builder.restype = C.c_int
    lib.lc_b200_set_lowering.argtypes = [C.c_int]
    lib.lc_b200_set_lowering.restype = C.c_int
    lib.lc_b200_ir_layout_json.argtypes = []
    lib.lc_b200_ir_layout_json.restype = C.c_char_p
    if path is None:
        _lib = lib
    return lib
