"""[k, j, i] shift / half-average / gradient helpers, mirror of the reference `coordinates_3d`
(coordinates_3d.py:32-98).  All shifts are periodic (np.roll) in all three axes, as in the reference."""
from ._shift import shift_op


def ipj(q): return shift_op(0, q, 0, -1)     # value at i+1   coordinates_3d.py:32
def imj(q): return shift_op(0, q, 0, 1)      # value at i-1   :39
def ijp(q): return shift_op(0, q, 1, -1)     # value at j+1   :43
def ijm(q): return shift_op(0, q, 1, 1)      # value at j-1   :47
def imjp(q): return imj(ijp(q))              # :51
def kp(q): return shift_op(0, q, 2, -1)      # value at k+1   :55
def km(q): return shift_op(0, q, 2, 1)       # value at k-1   :59
def kph(q): return shift_op(1, q, 2, -1)     # :63
def kmh(q): return shift_op(1, q, 2, 1)      # :67
def iph(q): return shift_op(1, q, 0, -1)     # :71
def imh(q): return shift_op(1, q, 0, 1)      # :75
def jph(q): return shift_op(1, q, 1, -1)     # :79
def jmh(q): return shift_op(1, q, 1, 1)      # :83
def gradi(q, dx): return shift_op(2, q, 0, d=dx)   # (ipj(q) - q) / dx   :87
def gradj(q, dy): return shift_op(2, q, 1, d=dy)   # (ijp(q) - q) / dy   :94
