"""Device implementation shared by coordinates.py / coordinates_1d.py / coordinates_3d.py."""
from . import _host, _lib


def shift_op(op, q, axis, shift=0, d=1.0):
    """op 0: np.roll(q, shift, axis)   1: (q + np.roll(q, shift, axis)) / 2   2: (np.roll(q, -1, axis) - q) / d
    axis counts from the fastest dimension: 0 = i, 1 = j, 2 = k (gcm_shift_op, include/gcm_b200.h)."""
    fam = _host.Family(q)
    t = _host.dev(q)
    assert t.dim() >= axis + 1, "array has no axis %d" % axis
    shp = list(t.shape)
    while len(shp) < 3:
        shp.insert(0, 1)
    n0, n1 = shp[-1], shp[-2]
    n2 = 1
    for s in shp[:-2]:
        n2 *= s
    if axis == 2 and len(shp) > 3:
        raise ValueError("k shifts need a 3-D [k, j, i] array")
    out = _host.empty(t.shape)
    _lib.check(_lib.lib().gcm_shift_op(op, _host.ptr(t), _host.ptr(out), n2, n1, n0, axis, shift, float(_host.scalar(d)),
                                       _lib.stream()), "gcm_shift_op")
    return fam.out(out)
