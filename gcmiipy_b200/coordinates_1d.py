"""[i] helpers, mirror of the reference `coordinates_1d` (coordinates_1d.py:25-53)."""
from ._shift import shift_op


def ip(q): return shift_op(0, q, 0, -1)          # :25
def im(q): return shift_op(0, q, 0, 1)           # :29
def iph(q): return shift_op(1, q, 0, -1)         # :33
def imh(q): return shift_op(1, q, 0, 1)          # :37
def gradh(q_i, dx): return shift_op(2, q_i, 0, d=dx)   # (ip(q) - q) / dx   :49


def div(q_h, dx):
    """(q_h - im(q_h)) / dx   (coordinates_1d.py:41) = gradh evaluated one cell to the left."""
    return im(gradh(q_h, dx))
