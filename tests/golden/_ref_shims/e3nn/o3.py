"""Import stub: the reference's sezm descriptor imports these names at module level; nothing here calls them."""


class FromS2Grid:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("e3nn is not available in this image")


class ToS2Grid(FromS2Grid):  # pragma: no cover
    pass


def spherical_harmonics(*a, **k):  # pragma: no cover
    raise RuntimeError("e3nn is not available in this image")
