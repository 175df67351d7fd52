"""A C program written against include/turtle.h + include/turtle_b200.h that makes the move
INTEGRATION.md section 2 describes (freeze, a fan with two columns back, explicit rays into
records, a multi-step walk with device states): strict C99 (-Wall -Wextra -Werror), links
against the library, does its set-up on the host and -- on a machine without a CUDA device --
is stopped by the FIRST batched call through the error handler: the batched path has no CPU
fallback."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "batched_caller.c")


def test_batched_c_caller(tmp_path, has_gpu):
    exe = str(tmp_path / "batched_caller")
    libdir = os.path.join(ROOT, "turtle_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1",
                           "-I" + os.path.join(ROOT, "include"), SRC, "-o", exe, "-L" + libdir,
                           "-lturtle_b200", "-Wl,-rpath," + libdir, "-lm"])
    r = subprocess.run([exe, "64", "32"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=300)
    assert r.stdout.startswith("station 4459355.543 220829.535 4540681.088\n")
    if has_gpu:
        assert r.returncode == 0, r.stderr
        assert [l.split()[0] for l in r.stdout.splitlines()] == ["station", "fan", "rays", "walk"]
    else:
        assert r.returncode == 3
        assert "turtle_stepper_freeze [#7]" in r.stderr and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_batched_c_caller_on_the_gpu(tmp_path, has_gpu):
    """The same program on the B200 box: all four stages run."""
    assert has_gpu
    test_batched_c_caller(tmp_path, has_gpu)
