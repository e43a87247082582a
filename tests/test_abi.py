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


def test_det_and_audio_entry_points_validate_arguments_without_a_gpu():
    lib = _lib.lib()
    n = ctypes.c_longlong(0)
    assert lib.air_det_workspace_bytes(_lib.LL(0), ctypes.byref(n)) == -1
    assert lib.air_det_workspace_bytes(_lib.LL(71237), None) == -1
    assert lib.air_det_workspace_bytes(_lib.LL(71237), ctypes.byref(n)) == 0
    # two key arrays + two class arrays + per-tile histograms: a little over 18 bytes per trial
    assert 18 * 71237 < n.value < 24 * 71237 and n.value % 256 == 0
    assert lib.air_det_launches(_lib.LL(71237), 0) == 1 + 3 * 5 + 4 and lib.air_det_launches(_lib.LL(71237), 1) == 1 + 3 * 8 + 4
    args = (None, _lib.LL(0), None, _lib.LL(0), 0, _lib.D(0), _lib.D(0), 0, None, _lib.LL(0), None, None, None, None, None, None)
    assert lib.air_det_curve_f32(*args) == -1 and lib.air_det_curve_f64(*args) == -1      # no scores, no workspace
    assert lib.air_det_threshold_counts_f32(None, _lib.LL(5), _lib.D(0), None, None) == -1
    assert lib.air_audio_info(None, None, None, None, None) == -1
    assert lib.air_audio_info(b"/nonexistent/x.flac", None, None, None, None) == -3
    assert lib.air_audio_decode_f32(b"/nonexistent/x.flac", None, _lib.LL(0), None, None, 0) == -1
    assert lib.air_audio_decode_batch_f32(None, 0, None, _lib.LL(0), None, None, None, 1, 0) == -1
