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


def test_last_error_string_names_the_rejecting_check():
    """air_last_error_string (SURVEY.md section 8b minimum export set): per-thread record of the last non-zero status."""
    import threading
    lib = _lib.lib()
    assert "air_last_error_string" in _lib.declared_symbols() and "air_status_string" in _lib.declared_symbols()
    assert lib.air_status_string(0) == b"AIR_OK" and lib.air_status_string(-2).startswith(b"AIR_ERR_UNSUPPORTED")
    seen = {}
    t = threading.Thread(target=lambda: seen.setdefault("fresh", lib.air_last_error_string()))
    t.start(); t.join()
    assert seen["fresh"] == b"no error recorded on this thread"
    st = lib.air_lfcc_fwd(None, _lib.LL(0), None, 0, 0, None, None, _lib.LL(0), _lib.LL(0), _lib.LL(0),
                          0, 0, 0, 0, None, None, ctypes.c_float(0.97), 0, None)
    msg = lib.air_last_error_string().decode()
    assert st == -1 and msg.startswith("status -1 at lfcc.cu:") and "AIR_ERR_ARG" in msg
    try:
        _lib.check(st, "air_lfcc_fwd")
    except _lib.AirError as e:
        assert "lfcc.cu:" in str(e)
    else:
        raise AssertionError("check() must raise")


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


def _header_signatures():
    """name -> list of parameter declarations, from include/air_b200.h."""
    import re
    with open(_lib.HEADER) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    sigs = {}
    for m in re.finditer(r"\b(?:int|long long|const char\s*\*)\s*(air_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = [p.strip() for p in m.group(2).replace("\n", " ").split(",")]
        sigs[m.group(1)] = [] if params == ["void"] or params == [""] else params
    return sigs


def test_every_ctypes_call_site_matches_the_header():
    """Static check of the binding layer: every `<lib>.air_xxx(...)` call in the package passes as many arguments as the
    header declares, wraps 64-bit integers / floats / doubles in the matching ctypes type (a bare Python number would
    be passed as a 32-bit int), and passes pointers as c_void_p / byref / None.  Covers wrappers no test can run here."""
    import ast
    import os
    sigs = _header_signatures()
    assert len(sigs) >= 70 and sigs["air_version"] == []
    pkg = os.path.dirname(_lib.__file__)
    wrap_of = {"long long": {"LL", "c_longlong"}, "unsigned long long": {"c_ulonglong"}, "float": {"F", "c_float"},
               "double": {"D", "c_double"}}
    all_wrappers = {w for ws in wrap_of.values() for w in ws} | {"c_int", "c_void_p", "c_char_p"}
    checked, problems = 0, []
    for fn in sorted(os.listdir(pkg)):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(pkg, fn)).read())

        def entry_points(expr):
            """air_* names an expression can evaluate to: `lib.air_x`, or `lib.air_x_f32 if cond else lib.air_x` (the
            bf16 / float-storage twins share one argument list)."""
            if isinstance(expr, ast.Attribute) and expr.attr.startswith("air_"):
                return [expr.attr]
            if isinstance(expr, ast.IfExp):
                a, b = entry_points(expr.body), entry_points(expr.orelse)
                return a + b if a and b else []
            return []

        calls = []                                   # (call node, [entry point names])
        for scope in ast.walk(tree):
            if not isinstance(scope, (ast.FunctionDef, ast.Module)):
                continue
            env = {}
            for node in (ast.walk(scope) if isinstance(scope, ast.FunctionDef) else ()):   # `fn = <entry point expr>` ... `fn(...)`
                if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name):
                    names = entry_points(node.value)
                    if names:
                        env[node.targets[0].id] = names
            for node in ast.walk(scope):
                if not isinstance(node, ast.Call):
                    continue
                names = env.get(node.func.id, []) if isinstance(node.func, ast.Name) else entry_points(node.func)
                if names:
                    calls.append((node, names))
        seen = set()
        for node, name in ((n, nm) for n, nms in calls for nm in nms):
            if (id(node), name) in seen:
                continue
            seen.add((id(node), name))
            if name not in sigs:
                problems.append("%s:%d calls undeclared %s" % (fn, node.lineno, name))
                continue
            if any(isinstance(a, ast.Starred) for a in node.args):
                continue
            checked += 1
            params = sigs[name]
            if len(node.args) != len(params):
                problems.append("%s:%d %s passes %d arguments, header declares %d" % (fn, node.lineno, name, len(node.args), len(params)))
                continue
            for a, p in zip(node.args, params):
                base = p.rsplit(" ", 1)[0].replace("const ", "").strip() if "*" not in p else "ptr"
                callee = a.func.attr if isinstance(a, ast.Call) and isinstance(a.func, ast.Attribute) else \
                    a.func.id if isinstance(a, ast.Call) and isinstance(a.func, ast.Name) else None
                if base in wrap_of:
                    if callee not in wrap_of[base]:
                        problems.append("%s:%d %s: `%s` is not wrapped as %s" % (fn, node.lineno, name, p, sorted(wrap_of[base])[0]))
                elif callee in all_wrappers and not (base == "ptr" and callee in ("c_void_p", "c_char_p")):
                    # argtypes come from the header (_lib.lib()): a c_longlong handed to an `int` parameter would raise
                    problems.append("%s:%d %s: `%s` receives a %s" % (fn, node.lineno, name, p, callee))
    assert checked >= 60, checked
    assert not problems, "\n".join(problems)
