"""No-op stand-in for matplotlib (TEST INFRASTRUCTURE ONLY): the reference imports
`matplotlib.pyplot` at module scope (e.g. matsuno_c_grid.py:8) but the hot path never plots."""
