"""The kernel profiles the rooflines read (profiles/r02_kernels.json, written by
tools/kernel_profiles.py from ncu captures) describe the kernels of THIS build: same kernel,
same registers. bench.py / tools/bench_configs.py refuse a stale entry at run time through
turtle_b200_kernel_info; this is the same check without a GPU, from the SASS of the built
library (cuobjdump -res-usage)."""
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "turtle_b200", "libturtle_b200.so")


def canonical(name):
    """`trace_kernel<0,0,6,1,0,0,0>` from a demangled name, whoever printed it (ncu or
    c++filt): no namespaces, no argument list, booleans as digits."""
    m = re.search(r"(\w+_kernel)(<[^>]*>)?\s*\(", name)
    if m is None:
        return None
    args = (m.group(2) or "").replace("tb::", "").replace("false", "0").replace("true", "1")
    args = re.sub(r"\([^)]*\)", "", args)  # casts of enumerators
    return m.group(1) + args.replace(" ", "")


def built_kernels():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], stdout=subprocess.PIPE, text=True,
                         check=True).stdout
    table, symbol = {}, None
    for line in out.splitlines():
        m = re.search(r"Function (\w+):", line)
        if m:
            symbol = m.group(1)
            continue
        m = re.search(r"REG:(\d+)", line)
        if m and symbol:
            table[symbol] = int(m.group(1))
            symbol = None
    names = subprocess.run(["c++filt"], input="\n".join(table), stdout=subprocess.PIPE,
                           text=True, check=True).stdout.splitlines()
    return {canonical(n): r for n, r in zip(names, table.values()) if canonical(n)}


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None,
                    reason="needs the CUDA binary utilities")
def test_profiles_describe_the_built_kernels():
    if not os.path.exists(LIB):
        pytest.skip("library not built (tests/test_abi.py says so loudly)")
    profiles = json.load(open(os.path.join(ROOT, "profiles", "r02_kernels.json")))
    built = built_kernels()
    assert len(built) > 40
    for name, entry in profiles.items():
        key = canonical(entry["kernel"])
        assert key in built, (name, key)
        assert built[key] == entry["registers"], \
            "%s: captured with %d registers, built with %d -- capture it again " \
            "(tools/capture_profiles.sh)" % (name, entry["registers"], built[key])
        assert os.path.exists(os.path.join(ROOT, "profiles", entry["source"]))


def test_canonical_names():
    assert canonical("void <unnamed>::trace_kernel<0, 0, 6, 1, 0, 0, 0>(Geometry, <unnamed>::"
                     "TraceArgs, <unnamed>::ExtraOf<T6, T7>::type)") == "trace_kernel<0,0,6,1,0,0,0>"
    assert canonical("void (anonymous namespace)::trace_kernel<false, false, 6, 1, false, false, "
                     "0>(tb::Geometry, (anonymous namespace)::TraceArgs, (anonymous namespace)::"
                     "ExtraOf<false, 0>::type)") == "trace_kernel<0,0,6,1,0,0,0>"
    assert canonical("void (anonymous namespace)::map_elevation_kernel<tb::NodesPacked>(tb::MapDesc,"
                     " unsigned long long)") == "map_elevation_kernel<NodesPacked>"
    assert canonical("(anonymous namespace)::to_geodetic_kernel(unsigned long long, double const*)") \
        == "to_geodetic_kernel"
