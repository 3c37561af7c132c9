"""ctypes binding of libmli_b200.so (include/mli_b200.h).  PyTorch is only the allocator / stream provider here:
tensors are passed as raw device pointers, the stream as ``torch.cuda.current_stream().cuda_stream``.

There is NO CPU fallback: if the library is missing it must be built (``python -m mli_nerf_b200.build``); if no
sm_100 device is current every compute entry point returns MLI_ENODEV and this module raises.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmli_b200.so")

MLI_MAX_LEVELS = 32
ACT_NONE, ACT_RELU, ACT_SOFTPLUS100, ACT_SIGMOID = 0, 1, 2, 3
PREC_FP32, PREC_BF16 = 0, 1
MODE_RGB, MODE_RGB_R_S, MODE_RGB_R, MODE_R_S, MODE_R_S_RE = 0, 1, 2, 3, 4
MODE_BY_NAME = {None: MODE_RGB, "rgb": MODE_RGB, "rgb_r_s": MODE_RGB_R_S, "rgb_r": MODE_RGB_R, "r_s": MODE_R_S,
                "r_s_re": MODE_R_S_RE}
LOSS_NAMES = ("total", "render", "eikonal", "curvature", "intrinsic", "regularize_re", "mse")


class Level(C.Structure):
    _fields_ = [("scale", C.c_float), ("res", C.c_uint32), ("size", C.c_uint32), ("offset", C.c_uint32),
                ("hashed", C.c_uint32)]


class Grid(C.Structure):
    _fields_ = [("n_levels", C.c_uint32), ("feat", C.c_uint32), ("active_levels", C.c_uint32),
                ("n_entries", C.c_uint32), ("level", Level * MLI_MAX_LEVELS)]


class CompositeCfg(C.Structure):
    _fields_ = [("N", C.c_int32), ("mode", C.c_int32), ("white_bg", C.c_int32), ("eval_extras", C.c_int32),
                ("anneal_ratio", C.c_float), ("anneal_dev", C.c_void_p)]


class LossCfg(C.Structure):
    _fields_ = [("w_render", C.c_float), ("w_eikonal", C.c_float), ("w_curvature", C.c_float),
                ("w_intrinsic", C.c_float), ("w_regularize_re", C.c_float), ("range_sha", C.c_float * 2),
                ("range_vis", C.c_float * 2), ("factor_ref", C.c_float), ("factor_sha", C.c_float),
                ("factor_negative", C.c_float), ("factor_positive", C.c_float), ("exponent_positive", C.c_float),
                ("has_intrinsic", C.c_int32), ("weights_dev", C.c_void_p)]


class TnReduceJob(C.Structure):
    """mli_tn_reduce_job_t (a deferred split-K reduction of mli_tc_wgrad_defer)."""
    _fields_ = [("part", C.c_void_p), ("cs_part", C.c_void_p), ("out", C.c_void_p), ("colsum", C.c_void_p),
                ("ldo", C.c_int64), ("out_batch_stride", C.c_int64), ("colsum_batch_stride", C.c_int64),
                ("S", C.c_int32), ("rows", C.c_int32), ("cols", C.c_int32), ("batch", C.c_int32),
                ("transpose", C.c_int32), ("reserved", C.c_int32)]


class WnDesc(C.Structure):
    """mli_wn_desc_t (batched weight_norm pack / unpack)."""
    _fields_ = [("v", C.c_void_p), ("g", C.c_void_p), ("col_map", C.c_void_p),
                ("Wp", C.c_void_p), ("tcl", C.c_void_p), ("tclt", C.c_void_p * 2),
                ("dWp", C.c_void_p), ("dv", C.c_void_p), ("dg", C.c_void_p),
                ("ldw", C.c_int64),
                ("N", C.c_int32), ("K", C.c_int32), ("row_off", C.c_int32),
                ("tcl_tile", C.c_int32), ("tcl_chunks", C.c_int32), ("tcl_lo", C.c_int32),
                ("tclt_c0", C.c_int32 * 2), ("tclt_c1", C.c_int32 * 2), ("tclt_tile", C.c_int32 * 2),
                ("tclt_chunks", C.c_int32 * 2), ("tclt_row_off", C.c_int32 * 2), ("tclt_col_off", C.c_int32 * 2),
                ("row_begin", C.c_int32),
                ("tcl2", C.c_void_p), ("tcl2_tile", C.c_int32), ("tcl2_chunks", C.c_int32)]


ADAMW_MAX_TENSORS = 64


class AdamwDesc(C.Structure):
    """mli_adamw_desc_t (one tensor of a batched AdamW launch)."""
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int64)]


# Signature table derived from include/mli_b200.h itself (single source of truth for the ABI):
#   p = device pointer (tensor / None / int), i = int32, l = int64, f = float, d = double, u = uint32,
#   s = stream, h / H = HOST int32 / float array (parameter names starting with host_).
HEADER_PATH = os.path.join(_HERE, "..", "include", "mli_b200.h")


def _parse_header(path=HEADER_PATH):
    import re
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(int|int64_t)\s+(mli_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), " ".join(m.group(3).split())
        if ret != "int" or args in ("void", ""):
            continue
        sig = ""
        for a in args.split(","):
            a = a.strip()
            pname = a.split()[-1].lstrip("*")
            if "*" in a:
                if pname == "stream":
                    sig += "s"
                elif pname.startswith("host_"):
                    sig += "h" if "int32_t" in a else "H"
                else:
                    sig += "p"
            else:
                ty = a.replace("const", "").split()[0]
                sig += {"int64_t": "l", "int32_t": "i", "uint32_t": "u", "float": "f", "double": "d", "int": "i"}[ty]
        sigs[name] = sig
    return sigs


_SIGS = {k: v for k, v in _parse_header().items() if k not in ("mli_abi_version", "mli_device_ok", "mli_grid_init",
                                                                "mli_set_sm_limit", "mli_peer_alloc", "mli_peer_open", "mli_l2_info")}
HOST_ONLY = {"mli_enable_peer_access", "mli_peer_close", "mli_peer_free"}  # management calls: no stream argument
_CTYPE = {"p": C.c_void_p, "i": C.c_int32, "l": C.c_int64, "f": C.c_float, "d": C.c_double, "u": C.c_uint32,
          "s": C.c_void_p, "h": C.c_void_p, "H": C.c_void_p}

_lib = None


class MliError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MliError(f"{LIB_PATH} not found: build it with `python -m mli_nerf_b200.build` "
                       "(hand-written sm_100a CUDA; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.mli_last_error.restype = C.c_char_p
    lib.mli_abi_version.restype = C.c_int
    lib.mli_device_ok.restype = C.c_int
    lib.mli_set_sm_limit.argtypes = [C.c_int32]
    lib.mli_set_sm_limit.restype = C.c_int
    lib.mli_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]
    lib.mli_peer_alloc.restype = C.c_int
    lib.mli_peer_open.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
    lib.mli_peer_open.restype = C.c_int
    lib.mli_grid_init.argtypes = [C.POINTER(Grid), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float]
    lib.mli_grid_init.restype = C.c_int
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = [_CTYPE[c] for c in sig]
        fn.restype = C.c_int
    for name in ("mli_linear_wgrad_ws_bytes", "mli_rowdot_bwd_ws_bytes", "mli_losses_ws_bytes", "mli_tc_wgrad_ws_bytes",
                 "mli_tc_colsum_ws_bytes", "mli_tc_sdf_trunk_bwd_ws_bytes"):
        getattr(lib, name).restype = C.c_int64
    lib.mli_tc_wgrad_ws_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32]
    lib.mli_tc_colsum_ws_bytes.argtypes = [C.c_int64, C.c_int32]
    lib.mli_tc_sdf_trunk_bwd_ws_bytes.argtypes = [C.c_int64]
    lib.mli_linear_wgrad_ws_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32]
    lib.mli_rowdot_bwd_ws_bytes.argtypes = [C.c_int64, C.c_int32, C.c_int32]
    lib.mli_losses_ws_bytes.argtypes = [C.c_int64, C.c_int64]
    _lib = lib
    return lib


def _raise(code, name):
    msg = (load().mli_last_error() or b"").decode()
    if code == -1 and "Only support 4 or 6 taps" in msg:
        raise ValueError(msg)  # same exception type/message as modules.py:177
    if code in (-1,):
        raise ValueError(f"{name}: {msg}")
    if code == -4:
        raise NotImplementedError(f"{name}: {msg}")
    raise MliError(f"{name} failed ({code}): {msg}")


def _ptr(x):
    if x is None:
        return None
    if isinstance(x, torch.Tensor):
        if not x.is_cuda:
            raise MliError("libmli_b200 takes CUDA tensors only (no CPU fallback)")
        return x.data_ptr()
    if isinstance(x, C.Structure):
        return C.addressof(x)
    return int(x)


def stream_ptr():
    if not torch.cuda.is_available():
        raise MliError("libmli_b200 needs a B200-class CUDA device: the render path has no CPU fallback")
    return torch.cuda.current_stream().cuda_stream


# ---- optional per-entry-point timing (CUDA events on the launching stream) used by bench.py ---------------------
_PROFILE = None
LAUNCH_COUNT = 0


_WORK_FN = None
_WORK = None


def profile_begin(work_fn=None):
    """work_fn(name, args) -> (flops, bytes) or None: the caller's accounting of the algorithmic work of one call
    (bench.py's per-kernel roofline); summed per entry point, read back with profile_work()."""
    global _PROFILE, _WORK_FN, _WORK
    _PROFILE, _WORK_FN, _WORK = {}, work_fn, {}


def profile_work():
    """-> {entry point: [flops, bytes]} accumulated since profile_begin(work_fn)."""
    return dict(_WORK or {})


def profile_end():
    """-> {entry point: (calls, total ms)}; synchronises."""
    global _PROFILE
    prof, _PROFILE = _PROFILE, None
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (prof or {}).items()}


def _n_launches(name, args):
    """How many of our kernels one call launches (for bench.py's gpu_launches claim)."""
    if name == "mli_tc_wgrad":
        return 3 if args[16] is not None else 2  # TN GEMM + split-K reduce (+ bias-gradient reduce when colsum_L is given)
    if name == "mli_tc_wgrad_reduce_batch":
        return (args[1] + 15) // 16
    if name in ("mli_linear_wgrad", "mli_tc_colsum"):
        return 2
    if name == "mli_tc_sdf_trunk_bwd":
        return 2 if args[9] is not None else 1
    if name == "mli_rowdot_bwd":
        return (1 if args[10] is not None else 0) + (2 if args[14] is not None else 0)
    if name == "mli_composite_bwd":
        return 1 + (1 if args[18] is not None else 0)
    if name in ("mli_weightnorm_pack_batch", "mli_weightnorm_unpack_grad_batch"):
        return 1
    if name in ("mli_copy_async", "mli_enable_peer_access", "mli_set_l2_window"):
        return 0  # copy-engine transfer / host call: no kernel
    if name == "mli_losses_fwd_bwd":
        return 3 + (1 if args[0].has_intrinsic else 0)
    return 1


def call(name, *args):
    """Invoke an entry point; tensors -> device pointers, 's' slot filled with the current stream when omitted."""
    global LAUNCH_COUNT
    if _PROFILE is not None:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _call(name, *args)
        b.record()
        _PROFILE.setdefault(name, []).append((a, b))
        if _WORK_FN is not None:
            w = _WORK_FN(name, args)
            if w is not None:
                acc = _WORK.setdefault(name, [0.0, 0.0])
                acc[0] += w[0]
                acc[1] += w[1]
    else:
        _call(name, *args)
    LAUNCH_COUNT += _n_launches(name, args)


_FN = {}  # entry point name -> (ctypes function, signature string): resolved once, the call path is hot in eager mode


def _call(name, *args):
    ent = _FN.get(name)
    if ent is None:
        ent = _FN[name] = (getattr(load(), name), _SIGS[name])
    fn, sig = ent
    n = len(sig)
    if len(args) == n - 1 and sig[-1] == "s":
        args = args + (None,)  # stream slot, filled below once the device is known
    elif len(args) != n:
        raise TypeError(f"{name}: expected {n} arguments, got {len(args)}")
    # One pass: convert, and check that every tensor of the call lives on ONE GPU -- the launch goes to that GPU's current
    # stream whatever the caller's current device is (a model on cuda:1 driven while the current device is cuda:0).
    conv, keep, dev = [None] * n, None, -1
    for k in range(n):
        c, a = sig[k], args[k]
        if c == "p":
            if a is None:
                continue
            if isinstance(a, torch.Tensor):
                d = a.device
                if d.type != "cuda":
                    raise MliError("libmli_b200 takes CUDA tensors only (no CPU fallback)")
                if dev != d.index:
                    if dev >= 0:
                        raise MliError(f"{name}: tensors on different devices (cuda:{dev} and cuda:{d.index})")
                    dev = d.index
                conv[k] = a.data_ptr()
            elif isinstance(a, C.Structure):
                conv[k] = C.addressof(a)
            else:
                conv[k] = int(a)
        elif c == "s":
            conv[k] = a  # resolved below
        elif c == "h" or c == "H":
            if a is not None:
                arr = ((C.c_int32 if c == "h" else C.c_float) * len(a))(*a)
                keep = (keep or []) + [arr]
                conv[k] = C.addressof(arr)
        elif c == "f" or c == "d":
            conv[k] = float(a)
        else:
            conv[k] = int(a)
    if dev >= 0 and dev != torch.cuda.current_device():
        with torch.cuda.device(dev):
            return _call(name, *args[:-1]) if args[-1] is None and sig[-1] == "s" else _call(name, *args)
    if sig[-1] == "s" and conv[-1] is None:
        conv[-1] = stream_ptr()
    code = fn(*conv)
    if code != 0:
        _raise(code, name)


def make_grid(n_levels, feat, log2_hashmap_size, base_resolution, per_level_scale):
    g = Grid()
    code = load().mli_grid_init(C.byref(g), n_levels, feat, log2_hashmap_size, base_resolution, per_level_scale)
    if code != 0:
        _raise(code, "mli_grid_init")
    return g


def device_ok():
    return bool(load().mli_device_ok())


class _RawCudaBuffer:
    """A raw device pointer seen by torch as a 1-D fp32 tensor (torch.as_tensor over __cuda_array_interface__)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def peer_alloc(n_floats, device):
    """-> (fp32 tensor over a fresh cudaMalloc allocation, its CUDA IPC handle as bytes, raw pointer)."""
    ptr, handle = C.c_void_p(), C.create_string_buffer(64)
    with torch.cuda.device(device):
        code = load().mli_peer_alloc(4 * int(n_floats), C.byref(ptr), handle)
    if code != 0:
        _raise(code, "mli_peer_alloc")
    t = torch.as_tensor(_RawCudaBuffer(ptr.value, n_floats), device=torch.device(device))
    return t, bytes(handle.raw), ptr.value


def peer_open(handle, device):
    """Map a peer rank's buffer (IPC handle bytes) into `device`'s address space -> raw pointer."""
    ptr, buf = C.c_void_p(), C.create_string_buffer(handle, 64)
    with torch.cuda.device(device):
        code = load().mli_peer_open(buf, C.byref(ptr))
    if code != 0:
        _raise(code, "mli_peer_open")
    return ptr.value


def l2_info():
    """-> (L2 bytes, maximum persisting carve-out, maximum access-policy window) of the current device."""
    out = (C.c_int32 * 3)()
    code = load().mli_l2_info(C.byref(out, 0), C.byref(out, 4), C.byref(out, 8))
    if code != 0:
        _raise(code, "mli_l2_info")
    return int(out[0]), int(out[1]), int(out[2])


def set_sm_limit(n_sms):
    code = load().mli_set_sm_limit(int(n_sms))
    if code != 0:
        _raise(code, "mli_set_sm_limit")
