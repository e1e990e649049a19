"""Minimal stand-in for the `array_api_compat` package (absent from this image, no network): NumPy >= 2 implements the
array-API functions in its main namespace, which is all the reference's NumPy backend needs for a CPU evaluation."""
import numpy as np


def array_namespace(*xs, **kw):
    return np


get_namespace = array_namespace


def device(x):
    return "cpu"


def to_device(x, device, **kw):
    return x


def size(x):
    return int(np.size(x))


def is_numpy_array(x):
    return isinstance(x, np.ndarray)


def is_array_api_obj(x):
    return isinstance(x, (np.ndarray, np.generic))


def is_torch_array(x):
    return False


def is_jax_array(x):
    return False


def is_jax_namespace(xp):
    return False


def is_torch_namespace(xp):
    return False


def is_numpy_namespace(xp):
    return xp is np
