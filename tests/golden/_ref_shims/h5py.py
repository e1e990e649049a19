"""Import stub (the reference's deepmd.utils.path imports h5py at module level; nothing here opens an HDF5 file)."""


class File:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("h5py is not available in this image")


class Group:  # pragma: no cover
    pass


class Dataset:  # pragma: no cover
    pass
