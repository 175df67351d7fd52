"""turtle_map_resample: a projected local map filled from a tile
stack by one kernel, against the node-by-node flow of examples/example-projection.c:88-104
run through the REFERENCE's own scalar calls (oracle/_ref: turtle_map_node ->
turtle_projection_unproject -> turtle_stack_elevation -> turtle_map_fill)."""
import numpy as np
import pytest

import turtle_b200 as tb
from oracle import harness as H

pytestmark = pytest.mark.gpu


def _reference_flow(stack_dir, tag, n, x, y, z):
    """-> (node elevations [n, n] as the reference stores them, nodes without data)."""
    ref = H.Driver(H.best_oracle())
    st = ref.stack_create(stack_dir)
    blank = ref.map_create(n, n, x, y, z, tag, np.full(n * n, z[0]))
    iy, ix = np.divmod(np.arange(n * n), n)
    nx_, ny_, _ = ref.map_node(blank, ix, iy)
    la, lo = ref.project(tag, nx_, ny_, inverse=True) if tag else (ny_, nx_)
    zs, inside = ref.stack_elevation(st, la, lo)
    filled = ref.map_create(n, n, x, y, z, tag, np.where(inside == 1, zs, z[0]))
    _, _, got = ref.map_node(filled, ix, iy)
    return got.reshape(n, n), int((inside == 0).sum())


def _nodes(m):
    info, _ = m.meta()
    return np.array([[m.node(ix, iy)[2] for ix in range(info.nx)] for iy in range(info.ny)])


@pytest.mark.parametrize("tag", ["Lambert 93", "UTM 31N", None])
def test_resample_matches_the_reference_flow(small_stack, tag):
    stack = tb.Stack(small_stack)
    stepper = tb.Stepper(range=0.)
    stepper.add_stack(stack, 0.)
    plan = stepper.freeze(0)
    n = 61
    if tag is None:
        x, y = (2.40, 2.46), (45.50, 45.56)
        proj = None
    else:
        proj = tb.Projection(tag)
        cx, cy = proj.project(45.53, 2.43)
        x, y = (cx - 2400., cx + 2400.), (cy - 2400., cy + 2400.)
    a = tb.Map(n, n, x, y, (-10., 3100.), tag)
    assert a.resample(plan, 0) == 0
    zb, missing = _reference_flow(small_stack, tag, n, x, y, (-10., 3100.))
    assert missing == 0
    za = _nodes(a)
    quantum = 3110. / 65535
    # the device's inverse projection differs from glibc's by a few ulp: at most one node
    # in a few thousand lands on the other side of a rounding boundary of the 16-bit scale
    assert np.abs(za - zb).max() <= quantum * 1.0001
    assert (za != zb).mean() < 2e-3
    assert za.std() > 1.  # real terrain, not a constant


def test_resample_counts_nodes_without_data(small_stack):
    """A map that sticks out of the stack (and over its missing tile N46E003): those nodes
    are left untouched and counted."""
    stack = tb.Stack(small_stack)
    stepper = tb.Stepper(range=0.)
    stepper.add_stack(stack, 0.)
    plan = stepper.freeze(0)
    m = tb.Map(41, 41, (2.8, 3.2), (45.8, 46.2), (-10., 3100.), None)
    missing = m.resample(plan, 0)
    _, want = _reference_flow(small_stack, None, 41, (2.8, 3.2), (45.8, 46.2), (-10., 3100.))
    assert missing == want and 0 < missing < 41 * 41
    # values outside of the map's z span are an error, as with turtle_map_fill
    with pytest.raises(tb.TurtleError, match="outside of map span"):
        tb.Map(11, 11, (2.4, 2.5), (45.4, 45.5), (5000., 6000.), None).resample(plan, 0)
    with pytest.raises(tb.TurtleError, match="invalid layer"):
        m.resample(plan, 3)
