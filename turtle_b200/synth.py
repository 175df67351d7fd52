"""Synthetic SRTM-shaped DEMs and ray sets of the BASELINE.json configurations.

The DEM is a deterministic fBm value noise on an integer-hash lattice (SURVEY.md
section 8d): splitmix64(seed ^ octave * K ^ ix << 32 ^ iy), 6 octaves, base wavelength
512 nodes, smoothstep interpolation, mapped to [0, 3000] m and rounded to int16.
Tiles are cut from ONE global lattice (node pitch 1 arc-second, origin at a given
integer degree corner), so shared edge rows / columns of neighbouring tiles are
identical, like real SRTM. The same bytes feed the oracle and the GPU.

Everything here is host-side input generation (numpy); nothing is on the hot path.
"""
import os

import numpy as np

SEED = 0x7075727474  # "turtle"
OCTAVES = 6
BASE_WAVELENGTH = 512
Z_MAX = 3000.0
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _lattice(octave, lx, ly, seed):
    """Hash values in [0, 1) at integer lattice points (lx[None, :], ly[:, None])."""
    with np.errstate(over="ignore"):
        key = (np.uint64(seed) ^ (np.uint64(octave) * np.uint64(0x9E3779B97F4A7C15))
               ^ (lx.astype(np.uint64)[None, :] << np.uint64(32))
               ^ (ly.astype(np.uint64)[:, None] & np.uint64(0xFFFFFFFF)))
    return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def fbm_grid(gx, gy, seed=SEED):
    """fBm in [0, 1] on the tensor grid gy[:, None] x gx[None, :] (float node coords).

    Coordinates may be fractional (used to resample the same terrain on a projected
    local map). Lattice indices are offset to stay positive."""
    gx = np.asarray(gx, dtype=np.float64) + 65536.0
    gy = np.asarray(gy, dtype=np.float64) + 65536.0
    out = np.zeros((len(gy), len(gx)))
    norm = 0.0
    for o in range(OCTAVES):
        w = BASE_WAVELENGTH / (1 << o)
        amp = 0.5 ** o
        ux, uy = gx / w, gy / w
        ix, iy = np.floor(ux).astype(np.int64), np.floor(uy).astype(np.int64)
        fx, fy = ux - ix, uy - iy
        sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)
        # unique lattice columns / rows, then gather
        lx = np.arange(ix.min(), ix.max() + 2)
        ly = np.arange(iy.min(), iy.max() + 2)
        h = _lattice(o, lx, ly, seed)
        jx, jy = ix - lx[0], iy - ly[0]
        h00 = h[np.ix_(jy, jx)]
        h10 = h[np.ix_(jy, jx + 1)]
        h01 = h[np.ix_(jy + 1, jx)]
        h11 = h[np.ix_(jy + 1, jx + 1)]
        top = h00 + (h10 - h00) * sx[None, :]
        bot = h01 + (h11 - h01) * sx[None, :]
        out += amp * (top + (bot - top) * sy[:, None])
        norm += amp
    return out / norm


def fbm_points(gx, gy, seed=SEED):
    """fBm in [0, 1] at scattered points (gx[i], gy[i])."""
    gx = np.asarray(gx, dtype=np.float64) + 65536.0
    gy = np.asarray(gy, dtype=np.float64) + 65536.0
    out = np.zeros(gx.shape)
    norm = 0.0
    for o in range(OCTAVES):
        w = BASE_WAVELENGTH / (1 << o)
        amp = 0.5 ** o
        ux, uy = gx / w, gy / w
        ix, iy = np.floor(ux).astype(np.int64), np.floor(uy).astype(np.int64)
        fx, fy = ux - ix, uy - iy
        sx, sy = fx * fx * (3 - 2 * fx), fy * fy * (3 - 2 * fy)

        def hv(kx, ky):
            with np.errstate(over="ignore"):
                key = (np.uint64(seed) ^ (np.uint64(o) * np.uint64(0x9E3779B97F4A7C15))
                       ^ (kx.astype(np.uint64) << np.uint64(32))
                       ^ (ky.astype(np.uint64) & np.uint64(0xFFFFFFFF)))
            return (splitmix64(key) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        h00, h10 = hv(ix, iy), hv(ix + 1, iy)
        h01, h11 = hv(ix, iy + 1), hv(ix + 1, iy + 1)
        top = h00 + (h10 - h00) * sx
        bot = h01 + (h11 - h01) * sx
        out += amp * (top + (bot - top) * sy)
        norm += amp
    return out / norm


def elevation_grid(gx, gy, seed=SEED):
    """Integer elevations (int16) of the synthetic terrain on a tensor grid."""
    return np.rint(fbm_grid(gx, gy, seed) * Z_MAX).astype(np.int16)


def tile_nodes(lat, lon, lat0, lon0, n=3601, seed=SEED):
    """int16[n, n] nodes of the 1 x 1 degree tile at (lat, lon), rows SOUTH first.
    (lat0, lon0) is the corner where global node coordinates are 0."""
    step = n - 1
    gx = (lon - lon0) * step + np.arange(n)
    gy = (lat - lat0) * step + np.arange(n)
    return elevation_grid(gx, gy, seed)


def hgt_name(lat, lon):
    return "%s%02d%s%03d.hgt" % ("N" if lat >= 0 else "S", abs(lat),
                                 "E" if lon >= 0 else "W", abs(lon))


def write_hgt_stack(directory, lat0, lon0, nlat, nlon, n=3601, seed=SEED, skip=()):
    """Write nlat x nlon `.hgt` tiles (big-endian int16, rows NORTH first,
    ref: src/turtle/io/hgt.c:127-147). n = 3601 (SRTMGL1 name) or 1201.
    `skip` lists (lat, lon) tiles to leave missing."""
    os.makedirs(directory, exist_ok=True)
    paths = []
    for j in range(nlat):
        for i in range(nlon):
            la, lo = lat0 + j, lon0 + i
            if (la, lo) in skip:
                continue
            name = hgt_name(la, lo)
            if n != 3601:
                name = name[:-4] + ".SRTMGL3.hgt"
            path = os.path.join(directory, name)
            if not os.path.exists(path):
                nodes = tile_nodes(la, lo, lat0, lon0, n, seed)
                nodes[::-1].astype(">i2").tofile(path + ".tmp")
                os.replace(path + ".tmp", path)
            paths.append(path)
    return paths


# ---- plain numpy geodesy for INPUT generation only (not a parity reference) -------

WGS84_A, WGS84_E = 6378137.0, 0.081819190842622


def np_ecef_from_geodetic(lat, lon, h):
    la, lo = np.radians(lat), np.radians(lon)
    s, c = np.sin(la), np.cos(la)
    R = WGS84_A / np.sqrt(1 - WGS84_E ** 2 * s * s)
    return np.stack([(R + h) * c * np.cos(lo), (R + h) * c * np.sin(lo),
                     (R * (1 - WGS84_E ** 2) + h) * s], axis=-1)


def np_from_horizontal(lat, lon, az, el):
    la, lo, az, el = np.radians(lat), np.radians(lon), np.radians(az), np.radians(el)
    sl, cl, sp, cp = np.sin(lo), np.cos(lo), np.sin(la), np.cos(la)
    e = np.stack([-sl, cl, np.zeros_like(sl)], -1)
    n = np.stack([-cl * sp, -sl * sp, cp], -1)
    u = np.stack([cl * cp, sl * cp, sp], -1)
    ce = np.cos(el)
    return ((ce * np.sin(az))[..., None] * e + (ce * np.cos(az))[..., None] * n
            + np.sin(el)[..., None] * u)


def fan_angles(r, n_az, n_el, el_min=0.5, el_max=30.0, part=0, parts=1, bundle=32):
    """(azimuth, elevation) of ray index r of the muography fan of config 2:
    az = 360 (i + (part + 1/2) / parts) / n_az, el = el_min + (el_max - el_min) (j + 1/2) / n_el.

    Ray order: elevation BANDS of `bundle` consecutive elevations, lowest band first;
    within a band azimuth by azimuth; within an azimuth the `bundle` elevations:
    r = (band * n_az + i) * bundle + k, j = band * bundle + k. Grazing rays take the most
    steps, so the longest rays are handed out first, and the `bundle` = 32 rays of a
    warp share one vertical plane: they walk the same ground track and gather the same
    DEM nodes (bundle = 1 is plain elevation-major order).
    `part` of `parts` interleaves the azimuths of several GPUs: together the parts form
    ONE fan of parts * n_az azimuths."""
    r = np.asarray(r, dtype=np.int64)
    band, rem = r // (n_az * bundle), r % (n_az * bundle)
    i, k = rem // bundle, rem % bundle
    j = band * bundle + k
    return (360.0 * (i + (part + 0.5) / parts) / n_az,
            el_min + (el_max - el_min) * (j + 0.5) / n_el)


def fan_directions(lat, lon, n_az, n_el, el_min=0.5, el_max=30.0, first=0, count=None,
                   part=0, parts=1, bundle=32):
    """ECEF directions of rays [first, first + count) of the fan (see fan_angles)."""
    total = n_az * n_el
    if count is None:
        count = total - first
    r = np.arange(first, first + count, dtype=np.int64)
    az, el = fan_angles(r, n_az, n_el, el_min, el_max, part, parts, bundle)
    return np_from_horizontal(np.full(count, lat), np.full(count, lon), az, el)


def golden_fan(n, el_min=0.5, el_max=30.0):
    """Config 1 fan: az = 360 (i + 1/2) / n, el = el_min + (el_max - el_min) frac(i phi)."""
    i = np.arange(n, dtype=np.float64)
    az = 360.0 * (i + 0.5) / n
    el = el_min + (el_max - el_min) * np.mod(i * 0.6180339887, 1.0)
    return az, el


def random_unit(n, seed, index=None):
    """Isotropic unit vectors from a counter-based splitmix64 stream (`index`: only those
    elements of the n-element stream, e.g. the strided shard of one rank)."""
    k = np.arange(n, dtype=np.uint64) if index is None else np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u1 = splitmix64(k * np.uint64(2) + np.uint64(seed) * np.uint64(0x1000003))
        u2 = splitmix64(k * np.uint64(2) + np.uint64(1) + np.uint64(seed) * np.uint64(0x1000003))
    a = (u1 >> np.uint64(11)).astype(np.float64) / (1 << 53)
    b = (u2 >> np.uint64(11)).astype(np.float64) / (1 << 53)
    cz = 2 * a - 1
    sz = np.sqrt(np.maximum(0.0, 1 - cz * cz))
    ph = 2 * np.pi * b
    return np.stack([sz * np.cos(ph), sz * np.sin(ph), cz], -1)


def random_uniform(n, seed, stream=0, index=None):
    k = np.arange(n, dtype=np.uint64) if index is None else np.asarray(index, dtype=np.uint64)
    with np.errstate(over="ignore"):
        u = splitmix64(k + (np.uint64(seed) << np.uint64(20)) + (np.uint64(stream) << np.uint64(44)))
    return (u >> np.uint64(11)).astype(np.float64) / (1 << 53)
