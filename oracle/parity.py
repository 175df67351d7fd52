"""Parity report of two sets of trace records -- TEST INFRASTRUCTURE (numpy only).

north_star: "per-ray medium/layer index and step count match the reference, bit-exact
except for rays grazing within the stepper's resolution (counted and reported). Exit
positions and path lengths must agree within 1e-9 relative / 1 mm."

`report(ref, got)` holds EVERY field the north star names to that tolerance -- all
`length[m]`, `total`, the exit `position` and `altitude` -- and says for each how many rays
are over it, the maximum and the 99-th percentile. Nothing is dropped; the rays are only
sorted in two classes, because the path has two kinds of end points:

  located    the quantity ends on a BOUNDARY the stepper bisects to 1e-8 m
             (stepper.c:832-864): lengths in media the ray has left, and everything about
             a ray that stopped by leaving the data (status DOMAIN). Well conditioned:
             held to 1 mm / 1e-9 strictly.
  threshold  the ray was stopped by the caller's altitude / length / step-count rule in
             the middle of a medium: its last position, total and final-medium length are
             wherever the optimistic steps ds = slope * |altitude - ground| happened to
             land. A perturbation d of one sample is carried RELATIVE to the height above
             the ground from then on (each step multiplies both by the same factor), so
             1e-8 m at the first centimetre-sized step off a boundary is metres after
             100 km. No implementation that rounds any operation differently can meet
             1 mm there: the reference itself does not when it is compiled with FMA
             contraction (oracle/_ref/libturtle_ref_fma.so). That self-difference is the
             ROUNDING-NOISE FLOOR of the path; `against_floor` holds the GPU to a small
             multiple of it, field by field, instead of exempting the class.
"""
import numpy as np

REL, ABS = 1e-9, 1e-3
STATUS_DOMAIN = 1
FIELDS = ("length0", "length1", "length2", "length3", "total", "position", "altitude")


def discrete_equal(ref, got):
    """Rays whose discrete outcome is identical: steps, stop status, final indices,
    number and sequence (hash) of media."""
    ok = np.ones(len(ref), dtype=bool)
    for f in ("n_steps", "status", "medium_hash", "n_changes"):
        ok &= ref[f] == got[f]
    return ok & (ref["index"] == got["index"]).all(1)


def _excess(ref, got, scale=None):
    """|got - ref| and the north-star tolerance max(1 mm, 1e-9 |ref|) per element."""
    d = np.abs(got - ref)
    tol = np.maximum(ABS, REL * np.abs(ref if scale is None else scale))
    return d, tol


def _stats(d, tol, mask):
    d, tol = d[mask], tol[mask]
    if d.size == 0:
        return {"n": 0, "over": 0, "max": 0.0, "p99": 0.0, "p50": 0.0}
    return {"n": int(d.size), "over": int((d > tol).sum()), "max": float(d.max()),
            "p99": float(np.percentile(d, 99)), "p50": float(np.percentile(d, 50))}


def deltas(ref, got):
    """Per ray and field: (|difference|, tolerance). Positions: the largest coordinate
    difference against 1 mm / 1e-9 of the geocentric distance."""
    out = {}
    for m in range(ref["length"].shape[1]):
        out["length%d" % m] = _excess(ref["length"][:, m], got["length"][:, m])
    out["total"] = _excess(ref["total"], got["total"])
    dp = np.abs(got["position"] - ref["position"]).max(1)
    out["position"] = (dp, np.maximum(ABS, REL * np.abs(ref["position"]).max(1)))
    out["altitude"] = _excess(ref["altitude"], got["altitude"])
    return out


def report(ref, got):
    """-> dict: discrete outcome counts + per class and field {n, over, max, p99, p50}."""
    n = len(ref)
    same = discrete_equal(ref, got)
    dl = deltas(ref, got)
    final = np.clip(ref["index"][:, 0], 0, ref["length"].shape[1] - 1)
    domain = ref["status"] == STATUS_DOMAIN
    located, threshold = {}, {}
    for f in FIELDS:
        d, tol = dl[f]
        if f.startswith("length"):
            left = final != int(f[-1])  # the ray is not in that medium any more
            located[f] = _stats(d, tol, same & (left | domain))
            threshold[f] = _stats(d, tol, same & ~left & ~domain)
        else:
            located[f] = _stats(d, tol, same & domain)
            threshold[f] = _stats(d, tol, same & ~domain)
    rays_located = same & domain
    rays_threshold = same & ~domain
    any_over_located = np.zeros(n, dtype=bool)
    any_over_threshold = np.zeros(n, dtype=bool)
    for f in FIELDS:
        d, tol = dl[f]
        over = d > tol
        if f.startswith("length"):
            left = final != int(f[-1])
            any_over_located |= same & (left | domain) & over
            any_over_threshold |= same & ~left & ~domain & over
        else:
            any_over_located |= rays_located & over
            any_over_threshold |= rays_threshold & over
    bit = (np.frombuffer(ref.tobytes(), dtype=np.uint8).reshape(n, -1) ==
           np.frombuffer(got.tobytes(), dtype=np.uint8).reshape(n, -1)).all(1) if n else \
        np.zeros(0, dtype=bool)
    return {"rays": int(n), "discrete_mismatch": int((~same).sum()),
            "bit_identical": int(bit.sum()),
            "rays_domain_exit": int(rays_located.sum()),
            "rays_threshold_exit": int(rays_threshold.sum()),
            "located_rays_over": int(any_over_located.sum()),
            "threshold_rays_over": int(any_over_threshold.sum()),
            "located": located, "threshold": threshold}


def against_floor(rep, floor, k=8.0):
    """The GPU-vs-reference report `rep` against the noise floor `floor` (reference vs the
    reference with FMA contraction, SAME rays). Returns a list of violations (empty = ok):

      * discrete outcome and located quantities: at most k x the floor's count (+ 1 ray
        per 50 000, so that a floor of zero on a small sample is not a trap), and no
        located quantity farther than k x the floor's maximum (or 1 mm);
      * threshold quantities: count over tolerance, p99 and maximum each within k x the
        floor's (the two perturbations are different random walks through the same
        amplifier: same scale, not the same rays)."""
    bad = []
    slack = 1 + rep["rays"] // 50000
    if rep["discrete_mismatch"] > k * floor["discrete_mismatch"] + slack:
        bad.append("discrete mismatches %d > %g x floor %d" % (
            rep["discrete_mismatch"], k, floor["discrete_mismatch"]))
    for cls in ("located", "threshold"):
        for f in FIELDS:
            a, b = rep[cls][f], floor[cls][f]
            if a["over"] > k * b["over"] + slack:
                bad.append("%s.%s: %d rays over tolerance > %g x floor %d" % (
                    cls, f, a["over"], k, b["over"]))
            if a["max"] > k * max(b["max"], ABS):
                bad.append("%s.%s: max %.3g m > %g x floor %.3g m" % (
                    cls, f, a["max"], k, b["max"]))
            if cls == "threshold" and a["p99"] > k * max(b["p99"], ABS):
                bad.append("%s.%s: p99 %.3g m > %g x floor %.3g m" % (
                    cls, f, a["p99"], k, b["p99"]))
    return bad


def table(rep, floor=None):
    """Plain-text table of a report (next to its floor when given)."""
    rows = ["rays %d | discrete mismatch %d%s | bit-identical %d | domain exits %d, threshold "
            "exits %d" % (rep["rays"], rep["discrete_mismatch"],
                          (" (floor %d)" % floor["discrete_mismatch"]) if floor else "",
                          rep["bit_identical"], rep["rays_domain_exit"],
                          rep["rays_threshold_exit"])]
    head = "%-10s %-9s %9s %8s %10s %10s" % ("class", "field", "n", ">tol", "p99 [m]", "max [m]")
    if floor:
        head += " | %8s %10s %10s" % ("fl >tol", "fl p99", "fl max")
    rows.append(head)
    for cls in ("located", "threshold"):
        for f in FIELDS:
            a = rep[cls][f]
            if a["n"] == 0:
                continue
            row = "%-10s %-9s %9d %8d %10.3g %10.3g" % (cls, f, a["n"], a["over"], a["p99"], a["max"])
            if floor:
                b = floor[cls][f]
                row += " | %8d %10.3g %10.3g" % (b["over"], b["p99"], b["max"])
            rows.append(row)
    return "\n".join(rows)
