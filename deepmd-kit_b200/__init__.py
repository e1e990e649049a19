"""deepmd-kit_b200: B200-native (sm_100a) implementation of DeePMD-kit's compressed se_e2_a /
se_atten force-evaluation hot path behind the reference's own operator surface.

The directory name is not a Python identifier; import it through ``__graft_entry__.load_package()``
(which registers it as ``deepmd_kit_b200``).  Importing the package loads ``lib/libdpb200.so`` and
defines ``torch.ops.deepmd.*``; it fails loudly when the CUDA library has not been built.
"""
from . import _lib

_lib.lib()  # fail loudly if the native library is missing

from . import ops  # noqa: E402

ops.register_torch_ops()

__all__ = ["ops"]
