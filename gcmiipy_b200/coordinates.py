"""[j, i] shift / half-average / gradient helpers, mirror of the reference `coordinates`
(coordinates.py:29-79); direction pinned by the reference's test_matsumo.py:9-21."""
from ._shift import shift_op


def ipj(q): return shift_op(0, q, 0, -1)     # coordinates.py:29
def imj(q): return shift_op(0, q, 0, 1)
def ijp(q): return shift_op(0, q, 1, -1)
def ijm(q): return shift_op(0, q, 1, 1)
def imjp(q): return imj(ijp(q))              # coordinates.py:48
def iph(q): return shift_op(1, q, 0, -1)
def imh(q): return shift_op(1, q, 0, 1)
def jph(q): return shift_op(1, q, 1, -1)
def jmh(q): return shift_op(1, q, 1, 1)
def gradi(q, dx): return shift_op(2, q, 0, d=dx)
def gradj(q, dy): return shift_op(2, q, 1, d=dy)
