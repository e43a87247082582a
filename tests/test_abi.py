"""CPU tests of the C-ABI boundary: the library loads and exports every declared symbol."""
import ctypes

from asvspoof2021_air_b200 import _lib


def test_library_exports_every_declared_symbol():
    names = _lib.declared_symbols()
    assert "air_lfcc_fwd" in names and "air_version" in names
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.air_version() >= 100
    assert lib.air_lfcc_table_floats() > 0


def test_argument_errors_are_reported_without_a_gpu():
    lib = _lib.lib()
    # null pointers -> AIR_ERR_ARG before any CUDA call
    st = lib.air_lfcc_fwd(None, _lib.LL(0), None, 0, 0, None, None, _lib.LL(0), _lib.LL(0), _lib.LL(0),
                          0, 0, 0, 0, None, None, ctypes.c_float(0.97), 0, None)
    assert st == -1
