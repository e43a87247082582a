"""ctypes binding of libair_b200.so (the C ABI declared in include/air_b200.h).

There is deliberately NO fallback: if the library is missing or a symbol cannot be resolved the
import fails loudly, and every wrapper raises on a non-zero status.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libair_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "air_b200.h")

_lib = None


class AirError(RuntimeError):
    pass


def _declarations(header=HEADER):
    """[(return type, name, [parameter declarations])] of every entry point declared in the public header."""
    with open(header) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = []
    for m in re.finditer(r"\b(int|long long|const char\s*\*)\s*(air_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(3).replace("\n", " ").split(",")]
        out.append((m.group(1), m.group(2), [] if params in (["void"], [""]) else params))
    return out


def declared_symbols(header=HEADER):
    """Names of every `int air_*(...)` entry point declared in the public header."""
    return [name for _, name, _ in _declarations(header)]


_SCALARS = {"int": ctypes.c_int, "unsigned": ctypes.c_uint, "long long": ctypes.c_longlong, "unsigned long long": ctypes.c_ulonglong,
            "float": ctypes.c_float, "double": ctypes.c_double, "air_stream_t": ctypes.c_void_p}


def _ctype_of(param):
    """ctypes type of one C parameter declaration: every pointer (device or host) travels as c_void_p."""
    if "*" in param:
        return ctypes.c_char_p if re.match(r"const char\s*\*\s*\w+$", param) else ctypes.c_void_p
    base = param.rsplit(" ", 1)[0].replace("const ", "").strip()
    return _SCALARS[base]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AirError("libair_b200.so is not built (%s); run `python -m asvspoof2021_air_b200.build`. "
                           "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        for ret, name, params in _declarations():          # restype AND argtypes from the header: a wrong Python
            fn = getattr(_lib, name)                       # argument raises ctypes.ArgumentError instead of being
            fn.restype = (ctypes.c_longlong if ret == "long long" else ctypes.c_char_p if "char" in ret
                          else ctypes.c_int)                                            # truncated to a 32-bit int
            fn.argtypes = [_ctype_of(p) for p in params]
    return _lib


LAUNCHES = [0]        # kernels launched through the C ABI by this process (bench.py reports it)


def check(status, what, n=1):
    LAUNCHES[0] += n
    if status != 0:
        kind = "argument error" if status < 0 else "CUDA error"
        try:
            detail = lib().air_last_error_string().decode()
        except Exception:                                   # never mask the original failure
            detail = ""
        raise AirError("%s failed: %s %d (%s)" % (what, kind, status, detail))


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


I = ctypes.c_int
LL = ctypes.c_longlong
F = ctypes.c_float
D = ctypes.c_double
