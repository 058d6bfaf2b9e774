"""CPU: the C-ABI library loads, exports every symbol include/maxpro_b200.h declares, and the
host-side mirror of the reference's Python surface validates arguments like the pyo3 wrappers.
No compute call needs a GPU here; on a GPU-less box compute calls must FAIL LOUDLY."""
import os
import re
import subprocess

import numpy as np
import pytest

from maxpro_b200 import MaxproError, _lib, engine, maxpro

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "maxpro_b200.h")).read()
    return sorted(set(re.findall(r"MAXPRO_API[^;(]*?\b(maxpro_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == sorted(_lib.EXPORTS)


def test_library_exports_every_declared_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (maxpro_[a-z0-9_]+)", out))
    assert set(header_symbols()) <= exported
    L = _lib.lib()
    for name in header_symbols():
        assert hasattr(L, name)
    assert b"maxpro_b200" in L.maxpro_version()


def test_no_oracle_in_product_path():
    """The product package must not import, link or call anything under oracle/."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "maxpro_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "import oracle" not in text and "from oracle" not in text, f
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_fails_loudly_without_gpu():
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(MaxproError) as e:
        engine.maxpro_criterion([[0.0, 1.0], [1.0, 0.0]])
    assert e.value.code == _lib.ERR_NO_DEVICE
    with pytest.raises(MaxproError):
        engine.build_lhd(10, 2, 100, _lib.METRIC_MAXPRO, 1)


def test_reduce_best_is_the_reference_fold():
    """src/lib.rs:76-98: strict win keeps the left element, so equal scores keep the later index."""
    assert engine.reduce_best(_lib.METRIC_MAXPRO, [3.0, 1.0, 1.0, 2.0], [10, 40, 20, 30]) == (1.0, 40)
    assert engine.reduce_best(_lib.METRIC_MAXIMIN, [0.3, 0.9, 0.9, 0.2], [1, 2, 7, 9]) == (0.9, 7)
    s, i = engine.reduce_best(_lib.METRIC_MAXPRO, [], [])
    assert s == np.inf and i == 2 ** 64 - 1  # identity (lib.rs:77-83)
    # shard records with "no candidate" markers are ignored
    assert engine.reduce_best(_lib.METRIC_MAXPRO, [np.inf, 5.0], [2 ** 64 - 1, 3]) == (5.0, 3)


def test_l2_distance_host():
    assert engine.l2_distance([0.0, 1.0], [1.0, 0.0]) == 2.0 ** 0.5
    with pytest.raises(ValueError):
        engine.l2_distance([0.0, 1.0], [1.0])
    with pytest.raises(MaxproError):
        engine.l2_distance([], [])


def test_python_surface_validation_messages():
    """ValueError texts of the pyo3 wrappers."""
    cases = [
        (lambda: maxpro.generate_lhd(0, 2, 1), "n_samples must be a positive, nonzero integer"),
        (lambda: maxpro.generate_lhd(5, 0, 1), "n_dim must be a positive, nonzero integer"),
        (lambda: maxpro.build_lhd(0, 2, 10, "maxpro", 1), "n_samples must be positive and nonzero"),
        (lambda: maxpro.build_lhd(5, 0, 10, "maxpro", 1), "n_dim must be positive and nonzero"),
        (lambda: maxpro.build_lhd(5, 2, 0, "maxpro", 1), "n_iterations must be positive and nonzero"),
        (lambda: maxpro.build_lhd(5, 2, 10, "MaxPro", 1), "Metric must be 'maxpro' or 'maximin'."),
        (lambda: maxpro.maxpro_criterion([]), "Design cannot be empty"),
        (lambda: maxpro.maximin_criterion([]), "Design cannot be empty"),
        (lambda: maxpro.anneal_lhd([[0.1, 0.2]], 0, 1.0, 0.99, "maxpro", True, True, 1), "n_iterations must be positive and nonzero"),
        (lambda: maxpro.anneal_lhd([[0.1, 0.2]], 5, 0.0, 0.99, "maxpro", True, True, 1), "initial_temp must be positive"),
        (lambda: maxpro.anneal_lhd([[0.1, 0.2]], 5, 1.0, 0.99, "nope", True, True, 1),
         "Unknown metric: 'nope'. Available metrics are 'maxpro' and 'maximin'."),
        (lambda: maxpro.order_design([[0.1, 0.2]], "nope"),
         "Unknown metric: 'nope'. Available metrics are 'maxpro' and 'maximin'."),
    ]
    for fn, msg in cases:
        with pytest.raises(ValueError) as e:
            fn()
        assert str(e.value) == msg
    assert maxpro.order_design([], "maxpro") == []          # order.rs:163-165
    assert maxpro.order_design([[]], "maxpro") == [[]]      # order.rs:42-46, 140-149
    assert maxpro.anneal_lhd([], 5, 1.0, 0.99, "MAXPRO", True, True, 1) == []  # anneal.rs:69-71, name lower-cased
    with pytest.raises(TypeError):
        maxpro.anneal_lhd([[0.1, 0.2]], 5, 1.0, 0.99, "maxpro", True, 42)  # stale example: seed lands on `swap`
    with pytest.raises(OverflowError):
        maxpro.generate_lhd(-1, 2, 1)


def test_c_abi_argument_errors_before_device():
    L = _lib.lib()
    x = np.zeros((2, 2))
    out = np.zeros((2, 2))
    rc = L.maxpro_build_lhd(0, 2, 10, 0, 1, _lib.dptr(out), None, None)
    assert rc == _lib.ERR_INVALID and L.maxpro_last_error() == b"n_samples must be positive and nonzero"
    rc = L.maxpro_build_lhd(2, 2, 0, 0, 1, _lib.dptr(out), None, None)
    assert rc == _lib.ERR_INVALID and L.maxpro_last_error() == b"n_iterations must be positive and nonzero"
    rc = L.maxpro_anneal_lhd(_lib.dptr(x), 2, 2, 10, -1.0, 0.99, 0, 1, 1, 1, 1, _lib.dptr(out), None, None)
    assert rc == _lib.ERR_INVALID and L.maxpro_last_error() == b"initial_temp must be positive and nonzero"
    assert L.maxpro_generate_lhd(0, 3, 1, None) == _lib.OK            # lhd.rs:98-101
    assert L.maxpro_generate_lhd(3, 0, 1, _lib.dptr(out)) == _lib.ERR_INVALID  # lhd.rs:106
    assert L.maxpro_order_design(None, 0, 3, 0, 1, None, None) == _lib.OK      # order.rs:38-46
    assert L.maxpro_anneal_lhd(None, 0, 3, 5, 1.0, 0.9, 0, 1, 1, 1, 1, None, None, None) == _lib.OK  # anneal.rs:69-71


def test_rust_crate_declares_only_exported_symbols_with_header_arity():
    """rust/ is source only (no cargo in the image): at least its extern "C" block must name
    symbols the library exports, with as many parameters as the header declares."""
    ffi = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()
    block = ffi[ffi.index('unsafe extern "C" {'):]
    block = block[:block.index("\n}")]
    decls = re.findall(r"pub fn (maxpro_[a-z0-9_]+)\s*\(([^)]*)\)", block, flags=re.S)
    assert len(decls) >= 8
    hdr = open(os.path.join(ROOT, "include", "maxpro_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    for name, params in decls:
        assert name in _lib.EXPORTS, name
        m = re.search(r"\b%s\s*\(([^)]*)\)" % name, hdr, flags=re.S)
        assert m, name
        c_params = [p for p in m.group(1).split(",") if p.strip() and p.strip() != "void"]
        rust_params = [p for p in params.split(",") if p.strip()]
        assert len(c_params) == len(rust_params), (name, c_params, rust_params)
    # build.rs compiles exactly the sources the Makefile does
    mk = open(os.path.join(ROOT, "maxpro_b200", "csrc", "Makefile")).read()
    srcs = re.search(r"^SRCS\s*:=\s*(.*)$", mk, flags=re.M).group(1).split()
    build_rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    assert sorted(re.findall(r'"([a-z_]+\.cu)"', build_rs)) == sorted(srcs)


def test_import_maxpro_is_the_reference_module_name():
    """`import maxpro` (pyo3 module name, src/lib.rs:166-176) resolves to the engine's mirror."""
    import maxpro as ref_named
    for fn in ("generate_lhd", "build_lhd", "maxpro_criterion", "maximin_criterion", "anneal_lhd", "order_design"):
        assert getattr(ref_named, fn) is getattr(maxpro, fn)
