"""The C-ABI library loads and exports every symbol the public headers declare;
without a GPU the batched path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import turtle_b200 as tb
from turtle_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for header in ("turtle.h", "turtle_b200.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"TURTLE_API[^;(]*?\b(turtle_\w+)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_path():
    names = declared_symbols()
    assert len(names) >= 70
    for must in ("turtle_stepper_step", "turtle_stepper_add_layer", "turtle_map_elevation",
                 "turtle_ecef_to_geodetic", "turtle_stepper_trace_batch",
                 "turtle_stepper_step_batch", "turtle_map_elevation_batch",
                 "turtle_ecef_to_geodetic_batch", "turtle_stepper_freeze"):
        assert must in names


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert missing == []
    # and the ctypes table binds each of them
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_headers_are_plain_c99_and_cxx(tmp_path):
    """Both headers compile as pedantic C99 and as C++11, and the record layouts have the
    sizes the Python mirror assumes."""
    import subprocess
    src = ('#include "turtle.h"\n#include "turtle_b200.h"\n'
           'int main(void) { return ((sizeof(struct turtle_trace_result) == 96) && '
           '(sizeof(struct turtle_trace_crossing) == 16) && '
           '(sizeof(struct turtle_trace_rule) == 32)) ? 0 : 1; }\n')
    for compiler, std, lang in (("gcc", "-std=c99", "c"), ("g++", "-std=c++11", "c++")):
        exe = str(tmp_path / ("hdr_" + lang.replace("+", "x")))
        subprocess.run([compiler, std, "-Wall", "-Wextra", "-pedantic", "-Werror",
                        "-I" + os.path.join(ROOT, "include"), "-x", lang, "-", "-o", exe],
                       input=src, text=True, check=True)
        assert subprocess.run([exe]).returncode == 0


def test_no_torch_types_in_the_abi():
    for header in ("turtle.h", "turtle_b200.h"):
        text = open(os.path.join(ROOT, "include", header)).read()
        assert "torch" not in text.lower() and "at::" not in text


def test_error_function_names():
    fn = C.cast(_lib.lib.turtle_stepper_step, C.c_void_p)
    assert _lib.lib.turtle_error_function(fn) == b"turtle_stepper_step"
    fn = C.cast(_lib.lib.turtle_stepper_trace_batch, C.c_void_p)
    assert _lib.lib.turtle_error_function(fn) == b"turtle_stepper_trace_batch"


def test_error_message_format():
    """`{ function [#code], file:line } message`, ref: error.c:108-138 and the
    expectations of tests/test-turtle.c:496-507."""
    with pytest.raises(tb.TurtleError) as e:
        tb.Map(10, 10, (0, 1), (0, 1), (0, 1), "nothing")
    assert e.value.code == 4  # TURTLE_RETURN_BAD_PROJECTION
    assert re.match(r"\{ turtle_map_create \[#4\], src/turtle/projection.c:[0-9]+ \} "
                    r"invalid projection `nothing'", str(e.value))
    with pytest.raises(tb.TurtleError) as e:
        tb.Map(0, 10, (0, 1), (0, 1), (0, 1), None)
    assert e.value.code == 6
    assert re.match(r"\{ turtle_map_create \[#6\], src/turtle/map.c:[0-9]+ \} "
                    r"invalid input parameter\(s\)", str(e.value))
    with pytest.raises(tb.TurtleError) as e:
        tb.Map(path="nothing")
    assert e.value.code == 2
    assert "no valid format for file `nothing'" in str(e.value)
    # the read-only formats refuse a dump from their own source file (grd.c:52-56, ...)
    m = tb.Map(3, 3, (0, 1), (0, 1), (0, 1), None)
    for ext in ("grd", "asc", "hgt"):
        with pytest.raises(tb.TurtleError) as e:
            m.dump("/tmp/never_written." + ext)
        assert re.match(r"\{ turtle_map_dump \[#3\], src/turtle/io/%s.c:[0-9]+ \} invalid write "
                        r"format for file `/tmp/never_written.%s'" % (ext, ext), str(e.value))


def test_walk_batch_checks_its_arguments_first():
    """turtle_stepper_walk_batch[_device]: no step or no directions is a domain error, raised
    before anything is touched (no plan, no device needed)."""
    d = np.zeros((1, 1, 3))
    for name, extra in (("turtle_stepper_walk_batch", ()),
                        ("turtle_stepper_walk_batch_device", (None,))):
        call = getattr(_lib.lib, name)
        for n_steps, direction in ((0, d.ctypes.data), (3, None)):
            with pytest.raises(tb.TurtleError) as e:
                tb.api._check(call(None, None, 1, n_steps, None, direction, None, None, None,
                                   None, None, None, *extra))
            assert e.value.code == 6  # TURTLE_RETURN_DOMAIN_ERROR
            assert name + " [#6]" in str(e.value)


def test_batch_path_fails_loudly_without_gpu(has_gpu):
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(tb.TurtleError) as e:
        tb.ecef_to_geodetic_batch(np.zeros((4, 3)))
    assert e.value.code == 7  # TURTLE_RETURN_LIBRARY_ERROR
    assert "no CUDA device" in str(e.value) and "no CPU fallback" in str(e.value)
    s = tb.Stepper()
    s.add_flat(0.)
    with pytest.raises(tb.TurtleError) as e:
        s.freeze(0)
    assert "no CUDA device" in str(e.value)
