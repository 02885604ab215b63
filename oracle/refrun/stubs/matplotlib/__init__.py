"""TEST INFRASTRUCTURE -- empty stand-in so that ``import matplotlib.pyplot`` at the top of the reference module
(/root/reference/mpopt/mpopt.py:25) succeeds; plotting is out of scope; calls are absorbed (see pyplot.py)."""
