"""ctypes binding of libsln_b200.so (C ABI declared in include/sln_b200.h).

There is no CPU fallback: if the shared library cannot be loaded (and cannot be built with nvcc), every hot-path
call raises.  The library is built in-tree (``sln_b200/libsln_b200.so``) so that it travels with the repo snapshot.
"""
import ctypes
import os
import shutil
import subprocess
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libsln_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG_DIR), "include")

SOURCES = ["runtime.cu", "vae_engine.cu", "raster.cu", "spade.cu", "collate.cu", "refine_loss.cu", "scene.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

_lock = threading.Lock()
_lib = None


def _sources():
    return [os.path.join(_CSRC, s) for s in SOURCES if os.path.exists(os.path.join(_CSRC, s))]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(_CSRC, f) for f in os.listdir(_CSRC)] + [os.path.join(INCLUDE_DIR, "sln_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _obj_stale(obj, src):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    deps = [src, os.path.join(INCLUDE_DIR, "sln_b200.h")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into libsln_b200.so (nvcc cross-compiles without a GPU).  Each translation
    unit is compiled to an object file in parallel (build/ is git-ignored), then linked."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("sln_b200: nvcc not found and %s is missing/stale; cannot build the CUDA library" % LIB_PATH)
    objdir = os.path.join(_PKG_DIR, "build")
    os.makedirs(objdir, exist_ok=True)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + os.environ.get("SLN_NVCC_EXTRA", "").split()   # e.g. -DSLN_TC_TRACE (tuning builds)
    jobs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if force or _obj_stale(obj, src):
            cmd = [nvcc] + cflags + ["-I", INCLUDE_DIR, "-c", "-o", obj, src]
            jobs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    errs = []
    for cmd, pr in jobs:
        out, err = pr.communicate()
        if pr.returncode != 0:
            errs.append("%s\n%s" % (" ".join(cmd), err[-8000:]))
        elif verbose:
            print(err)
    if errs:
        raise RuntimeError("sln_b200: nvcc failed\n" + "\n".join(errs))
    objs = [os.path.join(objdir, os.path.basename(s)[:-3] + ".o") for s in _sources()]
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", tmp] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("sln_b200: link failed\n%s\n%s" % (" ".join(cmd), res.stderr[-8000:]))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


class VaeDesc(ctypes.Structure):
    """Mirror of ``sln_vae_desc`` (include/sln_b200.h)."""
    _fields_ = [
        ("embedding_dim", ctypes.c_int32), ("n_layers", ctypes.c_int32), ("recurrent", ctypes.c_int32),
        ("norm", ctypes.c_int32), ("training", ctypes.c_int32), ("box_dim", ctypes.c_int32),
        ("n_angle", ctypes.c_int32), ("num_objs", ctypes.c_int32), ("num_preds", ctypes.c_int32),
        ("num_attrs", ctypes.c_int32), ("bn_eps", ctypes.c_float), ("bn_momentum", ctypes.c_float),
        ("gconv_dim_override", ctypes.c_int32), ("gconv_hidden_override", ctypes.c_int32),
        ("packed_weights", ctypes.c_void_p), ("bn_sync", ctypes.c_void_p), ("graph_ws", ctypes.c_void_p),
    ]


BN_SYNC_MAX_WORLD = 8


class BnSync(ctypes.Structure):
    """Mirror of ``sln_bn_sync`` (include/sln_b200.h): the device-resident table of the in-kernel SyncBatchNorm exchange."""
    _fields_ = [("world", ctypes.c_int32), ("rank", ctypes.c_int32), ("recv", ctypes.c_void_p * BN_SYNC_MAX_WORLD),
                ("flag", ctypes.c_void_p * BN_SYNC_MAX_WORLD), ("use", ctypes.c_void_p)]


_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_I32 = ctypes.c_int32
_SZ = ctypes.c_size_t
_F = ctypes.c_float
_DESC = ctypes.POINTER(VaeDesc)

# name -> (restype, argtypes); one entry per symbol declared in include/sln_b200.h
SIGNATURES = {
    "sln_version": (ctypes.c_int, []),
    "sln_last_error": (ctypes.c_char_p, []),
    "sln_launch_count": (_I64, []),
    "sln_prof_enable": (ctypes.c_int, [ctypes.c_int]),
    "sln_prof_read": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_I64)]),
    "sln_vae_num_params": (ctypes.c_int, [_DESC]),
    "sln_vae_num_bn": (ctypes.c_int, [_DESC]),
    "sln_vae_packed_bytes": (_SZ, [_DESC]),
    "sln_vae_pack_weights": (ctypes.c_int, [_DESC, _P, _P, _SZ, _P]),
    "sln_vae_workspace_bytes": (_SZ, [_DESC, _I64, _I64, ctypes.c_int]),
    "sln_vae_index_flag_offset": (_I64, [_DESC, _I64, _I64, ctypes.c_int]),
    "sln_bn_sync_recv_bytes": (_SZ, [_I32]),
    "sln_bn_sync_flag_bytes": (_SZ, []),
    "sln_vae_encoder_fwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "sln_vae_encoder_bwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, _P, _I64, _I64, _P, _SZ, _P]),
    "sln_vae_decoder_fwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "sln_vae_decoder_bwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, ctypes.c_int, _P, _I64, _I64, _P, _SZ, _P]),
    "sln_gconv_layer_fwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "sln_gconv_layer_bwd": (ctypes.c_int, [_DESC, _P, _P, _P, _P, _P, _P, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "sln_gconv_pool_workspace_bytes": (_SZ, [_I64, _I64]),
    "sln_csr_build": (ctypes.c_int, [_P, _I64, _I64, _I64, _P, _SZ, _P]),
    "sln_gconv_pool_fwd": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    "sln_csr_pointers": (ctypes.c_int, [_P, _I64, _I64, ctypes.POINTER(_P), ctypes.POINTER(_P)]),
    "sln_set_engine": (ctypes.c_int, [ctypes.c_int]),
    "sln_scene_assemble_workspace_bytes": (_SZ, [_I64]),
    "sln_scene_assemble_fwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _P, _I64, _P, _I64, _P, _P, _P, _I64, _P, _P, _F, _P, _P, _P, _P, _SZ, _P]),
    "sln_scene_assemble_bwd": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _P, _P, _P, _P, _P, _P, _SZ, _I32, _F, _P, _P, _P]),
    "sln_composite_workspace_bytes": (_SZ, [_I32]),
    "sln_composite_fwd": (ctypes.c_int, [_P, _P, _I32, _I64, _I32, _P, _I32, _P, _I32, _P, _P, _P, _SZ, _P]),
    "sln_composite_bwd": (ctypes.c_int, [_P, _P, _P, _I32, _I64, _P, _I32, _P, _I32, _P, _P, _P, _P, _SZ, _P]),
    "sln_refine_loss_workspace_bytes": (_SZ, [_I32, _P, _I32, _I32]),
    "sln_refine_loss": (ctypes.c_int, [_P, _I32, _P, _I32, _I32, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "sln_collate_layout": (ctypes.c_int, [_I64, _I64, _I64, _I32, ctypes.POINTER(_I64)]),
    "sln_collate_finish": (ctypes.c_int, [_P, _SZ, _I64, _I64, _I64, _I32, _P, _P, _P, _P, _P]),
    "sln_contract": (ctypes.c_int, [_P, _I64, _I32, _P, _I64, _I32, _P, _I64, _I64, _I64, _I64, _I32, _I32, _P]),
    "sln_raster_workspace_bytes": (_SZ, [_I64, _I64, _I32]),
    "sln_raster_setup": (ctypes.c_int, [_P, _I64, _P, _I64, _I32, _P, _P, _P, _F, _I32, _P, _SZ, _P]),
    "sln_raster_face_arrays": (ctypes.c_int, [_P, _I64, _I64, _I32, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P)]),
    "sln_raster_forward": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _F, _F, _P, _P, _P, _P]),
    "sln_raster_forward2": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P]),
    "sln_raster_texture_sample": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _P, _I32, _F, _P, _P, _P, _P, _P]),
    "sln_raster_backward_rgb": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _F, _P, _P, _P, _P, _P]),
    "sln_raster_backward_depth": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "sln_raster_vertex_grad": (ctypes.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _P, _P, _F, _P, _P, _P, _P]),
    "sln_scene_classes_fwd": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _I32, _F, _P, _P, _P, _P, _I32, _P, _P, _P]),
    "sln_scene_classes_bwd": (ctypes.c_int, [_P, _I64, _I64, _I32, _I32, _F, _P, _P, _I32, _P, _P, _P, _P]),
    "sln_reparam_fwd": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P]),
    "sln_reparam_bwd": (ctypes.c_int, [_P, _P, _P, _I64, _P, _P, _P]),
    "sln_vae_loss": (ctypes.c_int, [_P, _P, _I32, _P, _P, _I32, _P, _P, _I32, _F, _I64, _P, _P, _P, _I32, _P, _P, _P, _SZ, _P]),
    "sln_adam_step": (ctypes.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _F, _P, _I32, _P]),
    "sln_vae_loss_dyn": (ctypes.c_int, [_P, _P, _I32, _P, _P, _I32, _P, _P, _I32, _P, _I64, _P, _P, _P, _I32, _P, _P, _P, _SZ, _P]),
    "sln_adam_step_dyn": (ctypes.c_int, [_P, _P, _P, _P, _I64, _P, _F, _F, _F, _F, _F, _P, _I32, _P, _P]),
    "sln_packed_weights_bytes": (_SZ, [_I64, _I64]),
    "sln_pack_weights": (ctypes.c_int, [_P, _I64, _I64, _P, _P]),
    "sln_spade_conv": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _I32, _I32, _P, _P, _P, _I64, _P, _P]),
    "sln_conv2d_nhwc": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _I32, _I32, _I32, _P, _P, _P, _I64, _P, _P]),
    "sln_spade_modulate_ex": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _I32, _I32, _I32, _F, _P, _P]),
    "sln_instnorm_stats": (ctypes.c_int, [_P, _I64, _I64, _I64, _F, _P, _P, _P]),
    "sln_norm_act": (ctypes.c_int, [_P, _I64, _I64, _I64, _P, _P, _I32, _I32, _I32, _P, _P]),
    "sln_seg_resize_nhwc": (ctypes.c_int, [_P, _I64, _I32, _I32, _I32, _I64, _I64, _I32, _P, _P]),
    "sln_spade_modulate": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _I32, _P, _P, _P, _F, _P, _P]),
    "sln_spade_ln_stats": (ctypes.c_int, [_P, _I64, _I64, _F, _P, _P, _P, _P]),
    "sln_spade_seg_features": (ctypes.c_int, [_P, _I64, _I32, _I32, _I32, _I64, _I64, _P, _P, _I32, _P, _P]),
    "sln_spade_upsample2x": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _I32, _P, _P]),
    "sln_spade_se_residual": (ctypes.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, _I32, _P, _SZ, _P, _P]),
    "sln_spade_to_rgb": (ctypes.c_int, [_P, _I64, _I64, _I64, _I64, _P, _P, _I32, _I32, _F, _P, _P, _P]),
}


def declared_symbols():
    """Symbols declared in include/sln_b200.h (parsed from the header, used by the CPU-side ABI test)."""
    import re
    with open(os.path.join(INCLUDE_DIR, "sln_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sln_[a-z0-9_]+)\s*\(", text)))


def load():
    """Load (building if needed) the CUDA library.  Raises RuntimeError when it is unavailable — no fallback."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("SLN_LIB_PATH") or LIB_PATH      # SLN_LIB_PATH: an instrumented tuning build (tools/build_trace.sh)
        if path == LIB_PATH and _stale():
            build()
        try:
            lib = ctypes.CDLL(path)
        except OSError as e:
            raise RuntimeError("sln_b200: cannot load %s (%s); the hot path has no CPU/PyTorch fallback" % (path, e))
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(lib, name):
                continue  # optional components (raster/spade) may be absent in a partial build; checked by tests
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.sln_version() != 2:
            raise RuntimeError("sln_b200: ABI version mismatch (library %d, binding 2)" % lib.sln_version())
        eng = os.environ.get("SLN_ENGINE")
        if eng in ("0", "1"):      # 0 = FP32 SIMT tiles, 1 = tcgen05 3xTF32 tiles (default) for the MLP contractions
            lib.sln_set_engine(int(eng))
        _lib = lib
        return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sln_last_error()
        raise RuntimeError("sln_b200 %s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t):
    """Device pointer of a tensor (or None) as an int for ctypes."""
    return None if t is None else t.data_ptr()


def cur_stream(device=None):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def ptr_array(items):
    """Host array of device pointers (void*[]) from tensors / ints / None."""
    arr = (ctypes.c_void_p * max(len(items), 1))()
    for i, it in enumerate(items):
        if it is None:
            arr[i] = None
        elif isinstance(it, int):
            arr[i] = it
        else:
            arr[i] = it.data_ptr()
    return arr
