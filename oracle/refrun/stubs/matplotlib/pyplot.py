"""See the package docstring: nothing of the reference's hot path draws."""


def __getattr__(name):
    raise NotImplementedError(f"matplotlib.pyplot.{name}: plotting is out of scope of the reference run")
