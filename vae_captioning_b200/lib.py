"""ctypes binding of libvaecap.so -- the C-ABI boundary (include/vaecap.h).

There is deliberately no fallback: if the shared library is missing or a call fails, the
caller gets an exception. PyTorch tensors are used by callers only as device-memory containers;
everything that crosses this boundary is a raw pointer, a size or a scalar.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvaecap.so")

_lib = None


class VaecapError(RuntimeError):
    pass


_STATUS = {-1: "VC_E_ARG", -2: "VC_E_SHAPE", -3: "VC_E_CUDA", -4: "VC_E_NCCL", -5: "VC_E_STATE", -6: "VC_E_NOMEM"}


def load():
    """Loads (once) and returns the ctypes handle of libvaecap.so."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VaecapError(
            "libvaecap.so not found at %s -- build it with `python -m vae_captioning_b200.build` "
            "(there is no CPU fallback for the hot path)" % LIB_PATH)
    _prefer_bundled_nccl()
    lib = ctypes.CDLL(LIB_PATH)
    lib.vc_last_error.restype = ctypes.c_char_p
    lib.vc_abi_version.restype = ctypes.c_int
    _lib = lib
    return lib


def _prefer_bundled_nccl():
    """libvaecap binds libnccl with dlopen at the first vc_comm_* call. Inside a PyTorch process that must be the copy
    torch links against (site-packages/nvidia/nccl): two different libnccl.so.2 in one process clash on symbols."""
    if os.environ.get("VC_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["VC_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def check(status):
    """Converts a C-ABI status into the reference's error behaviour (Python exceptions)."""
    if status == 0:
        return
    msg = load().vc_last_error().decode("utf-8", "replace")
    name = _STATUS.get(status, str(status))
    if status in (-1, -2):
        raise ValueError("%s: %s" % (name, msg))
    raise VaecapError("%s: %s" % (name, msg))


def ptr(t):
    """Raw device/host pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _RawCudaBuffer(object):
    """Exposes a library-owned device allocation through __cuda_array_interface__ so torch can alias it."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def alias_tensor(ptr, count, dtype, device_index):
    """A torch tensor (container only) aliasing `count` elements at device pointer `ptr`; used to hand the
    library's flat gradient buffer to torch.distributed's NCCL all-reduce without a copy."""
    import torch
    typestr = {torch.float32: "<f4", torch.int32: "<i4", torch.bfloat16: "<u2"}[dtype]
    t = torch.as_tensor(_RawCudaBuffer(ptr, count, typestr), device=torch.device("cuda", device_index))
    if dtype == torch.bfloat16:
        t = t.view(torch.bfloat16)
    return t
