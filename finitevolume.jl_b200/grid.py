"""Host-side grid helpers (kept on the host, as north_star asks: "grid.jl still building
neighbor lists on the host").  numpy restatement of src/grid.jl with identical ordering:
node index i3 + n3*(i2-1) + n3*n2*(i1-1) (:60), faces emitted per node in the order
+x, +y, +z (:91-105), half cell widths at the two boundary indices (:73-86).

`regulargrid` can also emit only the faces a slab of x-planes owns or touches, which is
how each rank of a slab-partitioned run builds its part without ever holding the global
list (48 GiB of Pair{Int64,Int64} at 1024^3)."""
from __future__ import annotations

import numpy as np


def _axis(lo, hi, n):
    # range(lo; stop=hi, length=n); dx = xs[2] - xs[1]  (src/grid.jl:62-67)
    step = (hi - lo) / (n - 1)
    return (lo + 1 * step) - lo


def grid_sizes(ns):
    n1, n2, n3 = (int(v) for v in ns)
    N = n1 * n2 * n3
    return N, 3 * N - n1 * n2 - n1 * n3 - n2 * n3


def regulargrid(mins, maxs, ns, want_coords=True, planes=None, out=None):
    """src/grid.jl:56-110 -> (coords[3,N], neighbors[F,2], areasoverlengths[F], volumes[N]).

    planes=(p_lo, p_hi) (1-based inclusive x-plane range) restricts the output to the faces
    emitted by nodes of planes p_lo-1..p_hi (i.e. every face with an endpoint in planes
    p_lo..p_hi), in global face order, and volumes to planes p_lo..p_hi; coords are skipped.
    `out` may supply preallocated (e.g. pinned) arrays {"neighbors","areasoverlengths"}."""
    if len(mins) != len(maxs) or len(mins) != len(ns):
        raise AssertionError("mins, maxs, ns must have equal lengths")
    if len(mins) != 3:
        raise ValueError("only 3 dimensions supported")
    n1, n2, n3 = (int(v) for v in ns)
    dx, dy, dz = _axis(mins[0], maxs[0], n1), _axis(mins[1], maxs[1], n2), _axis(mins[2], maxs[2], n3)
    wy = np.full(n2, dy); wy[0] *= 0.5; wy[-1] *= 0.5
    wz = np.full(n3, dz); wz[0] *= 0.5; wz[-1] *= 0.5
    wx = np.full(n1, dx); wx[0] *= 0.5; wx[-1] *= 0.5
    if planes is None:
        e_lo, e_hi, v_lo, v_hi = 1, n1, 1, n1
    else:
        v_lo, v_hi = int(planes[0]), int(planes[1])
        e_lo, e_hi = max(1, v_lo - 1), v_hi  # plane v_lo-1 emits the +x faces into plane v_lo
    # faces emitted per plane i1: (i1<n1)*n2*n3 + (n2-1)*n3 + n2*(n3-1); only +x of plane v_lo-1 counts
    per_plane_full = (n2 - 1) * n3 + n2 * (n3 - 1)
    F = 0
    for i1 in range(e_lo, e_hi + 1):
        if planes is not None and i1 == v_lo - 1:
            F += n2 * n3
        else:
            F += per_plane_full + (n2 * n3 if i1 < n1 else 0)
    nb = out["neighbors"] if out else np.empty((F, 2), np.int64)
    aol = out["areasoverlengths"] if out else np.empty(F, np.float64)
    assert nb.shape[0] >= F and aol.shape[0] >= F
    vol = np.empty((v_hi - v_lo + 1) * n2 * n3, np.float64)
    # per-plane templates (i2, i3 grids flattened with i3 fastest)
    i2g, i3g = np.meshgrid(np.arange(1, n2 + 1), np.arange(1, n3 + 1), indexing="ij")
    i2f, i3f = i2g.ravel(), i3g.ravel()
    WY, WZ = wy[i2f - 1], wz[i3f - 1]
    hasy, hasz = i2f < n2, i3f < n3
    lin_in_plane = i3f + n3 * (i2f - 1)
    pos = 0
    for i1 in range(e_lo, e_hi + 1):
        lin = lin_in_plane + n3 * n2 * (i1 - 1)
        hasx = np.full(lin.shape, i1 < n1)
        only_x = planes is not None and i1 == v_lo - 1
        mask = np.stack([hasx, np.zeros_like(hasy) if only_x else hasy, np.zeros_like(hasz) if only_x else hasz], axis=1)
        other = np.stack([lin + n3 * n2, lin + n3, lin + 1], axis=1)
        a = np.stack([WY * WZ / dx, wx[i1 - 1] * WZ / dy, wx[i1 - 1] * WY / dz], axis=1)
        m = mask.ravel()
        k = int(m.sum())
        nb[pos:pos + k, 0] = np.repeat(lin, 3)[m]
        nb[pos:pos + k, 1] = other.ravel()[m]
        aol[pos:pos + k] = a.ravel()[m]
        pos += k
        if v_lo <= i1 <= v_hi:
            o = (i1 - v_lo) * n2 * n3
            vol[o:o + n2 * n3] = wx[i1 - 1] * WY * WZ
    assert pos == F
    coords = None
    if want_coords and planes is None:
        xs = np.array([mins[0] + i * ((maxs[0] - mins[0]) / (n1 - 1)) for i in range(n1)]); xs[-1] = maxs[0]
        ys = np.array([mins[1] + i * ((maxs[1] - mins[1]) / (n2 - 1)) for i in range(n2)]); ys[-1] = maxs[1]
        zs = np.array([mins[2] + i * ((maxs[2] - mins[2]) / (n3 - 1)) for i in range(n3)]); zs[-1] = maxs[2]
        coords = np.empty((3, n1 * n2 * n3))
        coords[0] = np.repeat(xs, n2 * n3)
        coords[1] = np.tile(np.repeat(ys, n3), n1)
        coords[2] = np.tile(zs, n1 * n2)
    return coords, nb[:F], aol[:F], vol


def nodehycos2neighborhycos(neighbors, nodehycos, logtransformhyco=False):
    """src/grid.jl:14-33.  nodehycos is shaped (n3,n2,n1) in Julia's column-major layout, i.e.
    its memory order is the node order; a flat node-ordered vector is accepted as well.
    Geometric mean, or the arithmetic mean of logs when logtransformhyco."""
    nb = np.asarray(neighbors, np.int64).reshape(-1, 2)
    k = np.asarray(nodehycos, np.float64)
    flat = k.reshape(-1, order="F") if k.ndim == 3 else k.reshape(-1)
    a, b = flat[nb[:, 0] - 1], flat[nb[:, 1] - 1]
    return 0.5 * (a + b) if logtransformhyco else np.sqrt(a * b)
