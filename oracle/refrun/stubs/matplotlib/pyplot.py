"""See the package docstring: nothing of the reference's hot path draws.  The reference's h-adaptive loop calls its
plotting helpers unconditionally (mpopt.py:2541), so every pyplot call returns an object that absorbs whatever is done
to it."""


class _Null:
    def __getattr__(self, name):
        return self

    def __call__(self, *a, **k):
        return self

    def __getitem__(self, k):
        return self

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0


_NULL = _Null()


def subplots(*a, **k):
    return _NULL, _NULL


def __getattr__(name):
    return _NULL
