"""Import stub for lmdb (no LMDB data is read here)."""


class Environment:  # pragma: no cover
    pass


class Transaction:  # pragma: no cover
    pass


def open(*a, **k):  # pragma: no cover
    raise RuntimeError("lmdb is not available in this image")
