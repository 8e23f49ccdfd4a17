"""ctypes binding of libipoke_b200.so (the C ABI declared in include/ipoke_b200.h).

There is no CPU fallback: importing this module without the built library raises, and every entry point raises
RuntimeError with the library's message when the native call fails.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libipoke_b200.so")

IPK_MAX_LEVELS = 32
IPK_MAX_DEC = 8
PREC = {"fp32_simt": 0, "fp32": 1, "fp32_split": 1, "bf16": 2}
DT_F32, DT_I64, DT_U8 = 0, 1, 2


class FlowConfig(ctypes.Structure):
    _fields_ = [("flow_in_channels", ctypes.c_int32), ("flow_mid_channels", ctypes.c_int32), ("h_channels", ctypes.c_int32),
                ("n_levels", ctypes.c_int32), ("num_steps", ctypes.c_int32 * IPK_MAX_LEVELS), ("factor", ctypes.c_int32),
                ("kernel_h", ctypes.c_int32), ("kernel_w", ctypes.c_int32), ("precision", ctypes.c_int32),
                ("max_batch", ctypes.c_int32)]


class FsConfig(ctypes.Structure):
    _fields_ = [("z_dim", ctypes.c_int32), ("spatial", ctypes.c_int32), ("n_gru_layers", ctypes.c_int32),
                ("n_dec", ctypes.c_int32), ("dec_channels", ctypes.c_int32 * IPK_MAX_DEC), ("precision", ctypes.c_int32),
                ("max_batch", ctypes.c_int32), ("max_frames", ctypes.c_int32), ("chunk_videos", ctypes.c_int32)]


class EncConfig(ctypes.Structure):
    _fields_ = [("z_dim", ctypes.c_int32), ("img_size", ctypes.c_int32), ("max_frames", ctypes.c_int32), ("full_seq", ctypes.c_int32),
                ("n_channels", ctypes.c_int32), ("channels", ctypes.c_int32 * IPK_MAX_DEC), ("min_spatial_size", ctypes.c_int32),
                ("max_batch", ctypes.c_int32), ("precision", ctypes.c_int32)]


class CencConfig(ctypes.Structure):
    _fields_ = [("nf_in", ctypes.c_int32), ("nf_max", ctypes.c_int32), ("spatial", ctypes.c_int32), ("min_spatial_size", ctypes.c_int32),
                ("n_stages", ctypes.c_int32), ("max_batch", ctypes.c_int32)]


class I3dConfig(ctypes.Structure):
    _fields_ = [("num_classes", ctypes.c_int32), ("max_batch", ctypes.c_int32), ("max_frames", ctypes.c_int32), ("precision", ctypes.c_int32)]


EXPORTS = [
    "ipk_version", "ipk_last_error", "ipk_launch_count", "ipk_launch_count_reset", "ipk_prof_enable", "ipk_prof_report",
    "ipk_flow_create", "ipk_flow_set_tensor", "ipk_flow_data_init", "ipk_flow_finalize", "ipk_flow_reverse", "ipk_flow_forward", "ipk_flow_destroy",
    "ipk_fs_create", "ipk_fs_set_tensor", "ipk_fs_finalize", "ipk_fs_decode", "ipk_fs_gru_step", "ipk_fs_gen", "ipk_fs_destroy",
    "ipk_cenc_create", "ipk_cenc_set_tensor", "ipk_cenc_finalize", "ipk_cenc_forward", "ipk_cenc_destroy",
    "ipk_i3d_create", "ipk_i3d_set_tensor", "ipk_i3d_finalize", "ipk_i3d_forward", "ipk_i3d_destroy", "ipk_i3d_preprocess",
    "ipk_enc_create", "ipk_enc_set_tensor", "ipk_enc_finalize", "ipk_enc_forward", "ipk_enc_destroy",
    "ipk_sample", "ipk_sample_host", "ipk_sample_host_u8", "ipk_frames_to_u8", "ipk_test_gemm", "ipk_test_conv3x3", "ipk_test_convT3x3", "ipk_test_conv3d", "ipk_tc_trace_enable", "ipk_tc_trace_read", "ipk_test_tc_plan",
    "ipk_flowtrain_create", "ipk_flowtrain_set_tensor", "ipk_flowtrain_finalize", "ipk_flowtrain_step", "ipk_flowtrain_forward", "ipk_flowtrain_backward", "ipk_flowtrain_destroy", "ipk_adam_step",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"ipoke_b200: native library not built ({LIB_PATH}); run `make` or __graft_entry__.build(). "
                           "There is no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, cp = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_char_p
    L.ipk_version.restype = ctypes.c_int
    L.ipk_last_error.restype = cp
    L.ipk_launch_count.restype = i64
    L.ipk_launch_count_reset.restype = None
    L.ipk_prof_enable.argtypes = [ctypes.c_int]
    L.ipk_prof_enable.restype = None
    L.ipk_prof_report.argtypes = [ctypes.c_char_p, ctypes.c_int]
    L.ipk_prof_report.restype = ctypes.c_int
    L.ipk_flow_create.argtypes = [ctypes.POINTER(FlowConfig), ctypes.POINTER(vp)]
    L.ipk_flow_set_tensor.argtypes = [vp, cp, vp, i64, ctypes.c_int]
    L.ipk_flow_finalize.argtypes = [vp, vp]
    L.ipk_flow_data_init.argtypes = [vp, vp, i32, vp]
    L.ipk_flow_reverse.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ipk_flow_forward.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.ipk_flow_destroy.argtypes = [vp]
    L.ipk_fs_create.argtypes = [ctypes.POINTER(FsConfig), ctypes.POINTER(vp)]
    L.ipk_fs_set_tensor.argtypes = [vp, cp, vp, i64, ctypes.c_int]
    L.ipk_fs_finalize.argtypes = [vp, vp]
    L.ipk_fs_decode.argtypes = [vp, vp, vp, vp, i32, i32, vp]
    L.ipk_fs_gru_step.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ipk_fs_gen.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ipk_fs_destroy.argtypes = [vp]
    L.ipk_cenc_create.argtypes = [ctypes.POINTER(CencConfig), ctypes.POINTER(vp)]
    L.ipk_cenc_set_tensor.argtypes = [vp, cp, vp, i64, ctypes.c_int]
    L.ipk_cenc_finalize.argtypes = [vp, vp]
    L.ipk_cenc_forward.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ipk_cenc_destroy.argtypes = [vp]
    L.ipk_enc_create.argtypes = [ctypes.POINTER(EncConfig), ctypes.POINTER(vp)]
    L.ipk_enc_set_tensor.argtypes = [vp, cp, vp, i64, ctypes.c_int]
    L.ipk_enc_finalize.argtypes = [vp, vp]
    L.ipk_enc_forward.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.ipk_enc_destroy.argtypes = [vp]
    L.ipk_i3d_create.argtypes = [ctypes.POINTER(I3dConfig), ctypes.POINTER(vp)]
    L.ipk_i3d_set_tensor.argtypes = [vp, cp, vp, i64, ctypes.c_int]
    L.ipk_i3d_finalize.argtypes = [vp, vp]
    L.ipk_i3d_forward.argtypes = [vp, vp, vp, i32, i32, vp]
    L.ipk_i3d_destroy.argtypes = [vp]
    L.ipk_i3d_preprocess.argtypes = [vp, vp, i64, i32, vp]
    L.ipk_sample.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.ipk_sample_host.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.ipk_sample_host_u8.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.ipk_frames_to_u8.argtypes = [vp, vp, i64, i32, vp]
    L.ipk_test_gemm.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    L.ipk_test_conv3x3.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.ipk_test_convT3x3.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.ipk_test_conv3d.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    f32 = ctypes.c_float
    L.ipk_flowtrain_create.argtypes = [ctypes.POINTER(FlowConfig), ctypes.POINTER(vp)]
    L.ipk_flowtrain_set_tensor.argtypes = [vp, cp, vp, vp, i64, ctypes.c_int]
    L.ipk_flowtrain_finalize.argtypes = [vp, vp]
    L.ipk_flowtrain_step.argtypes = [vp, vp, vp, vp, vp, vp, i32, vp]
    L.ipk_flowtrain_forward.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.ipk_flowtrain_backward.argtypes = [vp, vp, vp, vp, i32, vp]
    L.ipk_flowtrain_destroy.argtypes = [vp]
    f64 = ctypes.c_double
    L.ipk_adam_step.argtypes = [vp, vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, i32, f32, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name.startswith(("ipk_flow_", "ipk_fs_", "ipk_enc_", "ipk_cenc_", "ipk_sample", "ipk_frames_", "ipk_test_", "ipk_tc_", "ipk_flowtrain_", "ipk_adam_", "ipk_i3d_")):
            fn.restype = ctypes.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"ipoke_b200: {what} failed (status {rc}): {lib().ipk_last_error().decode()}")


def precision_code(p):
    if isinstance(p, int):
        return p
    if p not in PREC:
        raise ValueError(f"unknown precision {p!r}; use one of {sorted(PREC)}")
    return PREC[p]


def current_stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(t):
    import torch
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.int64:
        return DT_I64
    if t.dtype == torch.uint8:
        return DT_U8
    raise TypeError(f"unsupported tensor dtype {t.dtype}")


def launch_count():
    return int(lib().ipk_launch_count())


def launch_count_reset():
    lib().ipk_launch_count_reset()


def prof_enable(on=True):
    lib().ipk_prof_enable(1 if on else 0)


def prof_report():
    """{tag: (count, total_ms)} of the phases recorded since prof_enable(True); synchronises the device."""
    L = lib()
    n = L.ipk_prof_report(None, 0)
    buf = ctypes.create_string_buffer(n + 16)
    L.ipk_prof_report(buf, n + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        tag, cnt, ms = line.split()
        out[tag] = (int(cnt), float(ms))
    return out


def tensors_key(tensors):
    """Cheap identity of a list of parameters / buffers for plan caching: the sum of their version counters (bumped by every in-place
    update) and the xor of their storage addresses (a `.data` swap or a re-allocation does not bump the version)."""
    ver = ptr = 0
    for q in tensors:
        ver += q._version
        ptr ^= q.data_ptr()
    return ver, ptr
