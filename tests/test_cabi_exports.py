"""CPU: the C-ABI library builds, loads and exports every symbol include/gpmdm_b200.h declares.
No compute call is made (there is no GPU here and the library has no CPU path)."""
import ctypes
import os
import re

from gpmdm_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "gpmdm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gpmdm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for required in ("gpmdm_pf_transition_f64", "gpmdm_pf_propagate_f64", "gpmdm_pf_observe_f64",
                     "gpmdm_pf_normalize_f64", "gpmdm_pf_cdf_f64", "gpmdm_pf_resample_f64",
                     "gpmdm_pf_summaries_f64", "gpmdm_kernel_build_f64", "gpmdm_kernel_grad_f64"):
        assert required in names


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/gpmdm_b200.h but not exported"
    assert handle.gpmdm_abi_version() == _cabi.ABI_VERSION
    assert "#define GPMDM_ABI_VERSION %d" % _cabi.ABI_VERSION in open(os.path.join(ROOT, "include", "gpmdm_b200.h")).read()


def test_binding_table_matches_header():
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared_symbols()
    lib = _cabi.lib()
    assert lib.gpmdm_workspace_bytes(1 << 20, 8) > 0


def test_invalid_arguments_are_reported_not_crashed():
    lib = _cabi.lib()
    rc = lib.gpmdm_pf_transition_f64(None, None, None, 16, 2, None, None)
    assert rc < 0 and b"null" in lib.gpmdm_last_error()
    rc = lib.gpmdm_pf_cdf_f64(None, 0, 0, None, None, None)
    assert rc < 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gpmdm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
