"""ctypes view of include/nd_b200.h plus the in-tree nvcc build of libnd_b200.so.

There is no CPU fallback: if the shared library (or a CUDA device) is missing, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("ND_B200_LIB") or os.path.join(_PKG, "libnd_b200.so")   # override: A/B builds while tuning
SOURCES = [os.path.join(_PKG, "csrc", "nd_b200.cu")]
HEADERS = [os.path.join(_PKG, "csrc", "nd_b200_kernels.cuh"), os.path.join(_ROOT, "include", "nd_b200.h")] + \
          [os.path.join(_PKG, "csrc", n) for n in ("nd_b200_launch.inc", "nd_b200_custom.inc", "nd_b200_build.inc", "nd_b200_comm.inc")]

ABI_VERSION = 5
OK, EINVAL, EUNSUPPORTED, ECUDA, ENOMEM, ETIMEOUT = range(6)

# registry ids (include/nd_b200.h)
V_DIFFUSION, V_KURAMOTO_FIRST, V_KURAMOTO_SECOND, V_KURAMOTO_SECOND_BENCH, V_SWING_DQ = range(5)
E_DIFFUSION, E_DIFFUSION_NOP, E_KURAMOTO, E_LINE_DQ, E_DIFFUSION_ODE, E_RELAX_ODE, E_DIFFUSION_FID, E_LOOPBACK = range(8)
ANTISYMMETRIC, SYMMETRIC, DIRECTED, FIDUCIAL = range(4)
CUSTOM_KIND_BASE = 1000
FLAG_NO_EXPORT = 1
FLAG_HOST_ONLY = 2
FLAG_ROW_RANGE = 4

i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


class VBatch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("pdim", C.c_int32), ("outdim", C.c_int32),
                ("count", C.c_int64), ("indices", i64p), ("state_first", C.c_int64), ("p_first", C.c_int64),
                ("out_first", C.c_int64), ("aggr_first", C.c_int64), ("extdim", C.c_int32), ("reserved", C.c_int32),
                ("ext_src", i64p)]


class EBatch(C.Structure):
    _fields_ = [("kind", C.c_int32), ("coupling", C.c_int32), ("dim", C.c_int32), ("pdim", C.c_int32),
                ("outdim_src", C.c_int32), ("outdim_dst", C.c_int32), ("count", C.c_int64), ("indices", i64p),
                ("state_first", C.c_int64), ("p_first", C.c_int64), ("out_first", C.c_int64),
                ("gbuf_first", C.c_int64), ("mask_src_first", C.c_int32), ("mask_dst_first", C.c_int32),
                ("extdim", C.c_int32), ("reserved", C.c_int32), ("ext_src", i64p)]


class CustomKind(C.Structure):
    _fields_ = [("kind", C.c_int32), ("role", C.c_int32), ("dim", C.c_int32), ("pdim", C.c_int32), ("outdim", C.c_int32),
                ("two_sided", C.c_int32), ("f_body", C.c_char_p), ("g_body", C.c_char_p), ("extdim", C.c_int32),
                ("g_ff", C.c_int32)]


class Desc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("nv", C.c_int64), ("ne", C.c_int64),
                ("edge_src", i64p), ("edge_dst", i64p), ("vdepth", C.c_int32), ("edepth", C.c_int32),
                ("n_vbatches", C.c_int32), ("n_ebatches", C.c_int32), ("vbatches", C.POINTER(VBatch)),
                ("ebatches", C.POINTER(EBatch)), ("lastidx_dynamic", C.c_int64), ("lastidx_p", C.c_int64),
                ("lastidx_out", C.c_int64), ("lastidx_aggr", C.c_int64), ("row_begin", C.c_int64),
                ("row_end", C.c_int64), ("long_row_threshold", C.c_int32), ("flags", C.c_int32),
                ("gather_offset", i64p), ("gather_len", C.c_int64), ("n_custom", C.c_int32), ("reserved", C.c_int32),
                ("custom", C.POINTER(CustomKind))]


EXPORTED_SYMBOLS = [
    "nd_b200_create", "nd_b200_destroy", "nd_b200_last_error", "nd_b200_abi_version", "nd_b200_rhs",
    "nd_b200_rhs_host", "nd_b200_get_buffers", "nd_b200_aggregate", "nd_b200_rk4", "nd_b200_export_sizes", "nd_b200_export_tables",
    "nd_b200_launch_count", "nd_b200_set_timing", "nd_b200_timings", "nd_b200_host_alloc", "nd_b200_host_free",
    "nd_b200_comm_create", "nd_b200_comm_export", "nd_b200_comm_open_peer", "nd_b200_comm_set_send", "nd_b200_rhs_local", "nd_b200_custom_source", "nd_b200_create_from_edgelist", "nd_b200_rhs_exchange", "nd_b200_comm_status",
    "nd_b200_comm_last_error", "nd_b200_comm_destroy", "nd_b200_export_jag_sizes", "nd_b200_export_jag", "nd_b200_pack_params", "nd_b200_rk4_exchange", "nd_b200_kernel_name",
]
IPC_HANDLE_BYTES = 64


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 engine cannot be built (there is no CPU fallback)")


def write_embedded_header() -> str:
    """csrc/nd_b200_kernels_embed.inc: nd_b200_kernels.cuh as a list of C++ raw string literals.  The engine hands this
    text to NVRTC when a network brings user-supplied component kinds, so the run-time compiled kernels are the same
    templates as the precompiled ones."""
    src = open(HEADERS[0]).read()
    delim = "NDB200SRC"
    assert ")" + delim + '"' not in src
    out = os.path.join(_PKG, "csrc", "nd_b200_kernels_embed.inc")
    lines, chunk, size = src.splitlines(keepends=True), [], 0
    parts = []
    for ln in lines:
        chunk.append(ln)
        size += len(ln)
        if size > 8000:
            parts.append("".join(chunk))
            chunk, size = [], 0
    if chunk:
        parts.append("".join(chunk))
    with open(out, "w") as f:
        f.write("// generated by _cabi.write_embedded_header() from nd_b200_kernels.cuh -- do not edit\n")
        for part in parts:
            f.write('R"' + delim + "(" + part + ")" + delim + '",\n')
    return out


def build(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> networkdynamics.jl_b200/libnd_b200.so (in-tree)."""
    newest = max(os.path.getmtime(f) for f in SOURCES + HEADERS)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= newest:
        return LIB_PATH
    write_embedded_header()
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc_path())), "lib64")
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-fmad=false", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(_ROOT, "include"),
           "-L", cuda_lib, "-lnvrtc", "-Xlinker", "-rpath=" + cuda_lib,
           "-o", LIB_PATH] + SOURCES
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc wrapper without a complete tool chain
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd, env=env)
    return LIB_PATH


_lib = None


def lib():
    """Load libnd_b200.so; fail loudly when it is absent (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the B200 engine has no CPU fallback)")
    _lib = bind(C.CDLL(LIB_PATH))
    return _lib


def bind(L):
    """declare the signatures of include/nd_b200.h on a loaded library object"""
    dp = C.c_void_p  # raw device (or host) addresses
    L.nd_b200_create.restype = C.c_int
    L.nd_b200_create.argtypes = [C.POINTER(Desc), C.POINTER(C.c_void_p)]
    L.nd_b200_create_from_edgelist.restype = C.c_int
    L.nd_b200_create_from_edgelist.argtypes = [C.c_int32, C.c_int64, C.c_int64, i64p, i64p, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int64, C.c_int64, C.c_int32, i64p, C.c_int64, C.POINTER(C.c_void_p)]
    L.nd_b200_destroy.restype = None
    L.nd_b200_destroy.argtypes = [C.c_void_p]
    L.nd_b200_last_error.restype = C.c_char_p
    L.nd_b200_last_error.argtypes = [C.c_void_p]
    L.nd_b200_abi_version.restype = C.c_int
    L.nd_b200_rhs.restype = C.c_int
    L.nd_b200_rhs.argtypes = [C.c_void_p, dp, dp, dp, C.c_double, C.c_void_p]
    L.nd_b200_rhs_host.restype = C.c_int
    L.nd_b200_rhs_host.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
    L.nd_b200_aggregate.restype = C.c_int
    L.nd_b200_aggregate.argtypes = [C.c_void_p, dp, dp, C.c_void_p]
    L.nd_b200_get_buffers.restype = C.c_int
    L.nd_b200_get_buffers.argtypes = [C.c_void_p, dp, dp, dp, dp, C.c_double, C.c_void_p]
    L.nd_b200_pack_params.restype = C.c_int
    L.nd_b200_pack_params.argtypes = [C.c_void_p, dp, C.c_void_p]
    L.nd_b200_rk4.restype = C.c_int
    L.nd_b200_rk4.argtypes = [C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int64, C.c_void_p]
    L.nd_b200_export_sizes.restype = C.c_int
    L.nd_b200_export_sizes.argtypes = [C.c_void_p, i64p]
    L.nd_b200_export_tables.restype = C.c_int
    L.nd_b200_export_tables.argtypes = [C.c_void_p, i64p, i64p, i64p, i32p]
    L.nd_b200_export_jag_sizes.restype = C.c_int
    L.nd_b200_export_jag_sizes.argtypes = [C.c_void_p, i64p]
    L.nd_b200_export_jag.restype = C.c_int
    L.nd_b200_export_jag.argtypes = [C.c_void_p, i32p, C.POINTER(C.c_uint16), i32p, i32p]
    L.nd_b200_custom_source.restype = C.c_char_p
    L.nd_b200_custom_source.argtypes = [C.c_void_p]
    L.nd_b200_kernel_name.restype = C.c_char_p
    L.nd_b200_kernel_name.argtypes = [C.c_void_p]
    L.nd_b200_launch_count.restype = C.c_int64
    L.nd_b200_launch_count.argtypes = [C.c_void_p]
    L.nd_b200_set_timing.restype = C.c_int
    L.nd_b200_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.nd_b200_timings.restype = C.c_int
    L.nd_b200_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), i64p]
    L.nd_b200_host_alloc.restype = C.c_void_p
    L.nd_b200_host_alloc.argtypes = [C.c_int64]
    L.nd_b200_host_free.restype = None
    L.nd_b200_host_free.argtypes = [C.c_void_p]
    L.nd_b200_comm_create.restype = C.c_int
    L.nd_b200_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.POINTER(C.c_void_p)]
    L.nd_b200_comm_set_send.restype = C.c_int
    L.nd_b200_comm_set_send.argtypes = [C.c_void_p, C.c_int32, i64p, C.c_int64, C.c_int64]
    L.nd_b200_comm_export.restype = C.c_int
    L.nd_b200_comm_export.argtypes = [C.c_void_p, C.c_char_p]
    L.nd_b200_comm_open_peer.restype = C.c_int
    L.nd_b200_comm_open_peer.argtypes = [C.c_void_p, C.c_int32, C.c_char_p]
    L.nd_b200_rhs_exchange.restype = C.c_int
    L.nd_b200_rhs_exchange.argtypes = [C.c_void_p, C.c_void_p, dp, dp, dp, C.c_double, C.c_void_p]
    L.nd_b200_rk4_exchange.restype = C.c_int
    L.nd_b200_rk4_exchange.argtypes = [C.c_void_p, C.c_void_p, dp, dp, C.c_double, C.c_double, C.c_int64, C.c_void_p]
    L.nd_b200_rhs_local.restype = C.c_int
    L.nd_b200_rhs_local.argtypes = [C.c_void_p, C.c_void_p, dp, dp, dp, C.c_double, C.c_void_p]
    L.nd_b200_comm_status.restype = C.c_int
    L.nd_b200_comm_status.argtypes = [C.c_void_p, i32p]
    L.nd_b200_comm_last_error.restype = C.c_char_p
    L.nd_b200_comm_last_error.argtypes = [C.c_void_p]
    L.nd_b200_comm_destroy.restype = None
    L.nd_b200_comm_destroy.argtypes = [C.c_void_p]
    if L.nd_b200_abi_version() != ABI_VERSION:
        raise RuntimeError("libnd_b200.so ABI version mismatch; rebuild")
    return L
