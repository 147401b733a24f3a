"""pytorch_empirical-mvm_b200: B200-native (sm_100a) Video-Swin 3D hot path of EmpiricalMVM/VIOLETv2.

Drop-in for the reference's ``visbackbone/video_swin.py`` (see ``video_swin.py`` here); kernels live in
``csrc/`` behind the C ABI of ``include/vsw.h`` and are loaded through ctypes (``_lib``).
The directory name contains a '-', so import it with ``importlib.import_module("pytorch_empirical-mvm_b200")``.
"""
from . import _lib, dp, functional  # noqa: F401
from .video_swin import *  # noqa: F401,F403
from .video_swin import __all__ as _vs_all
from .enc_video import EncVideo  # noqa: F401  (reference model.py:7-78)
from . import mvm  # noqa: F401  (reference main_pretrain.py:309-362, 508-524)

__all__ = list(_vs_all) + ["EncVideo", "mvm", "functional", "_lib", "dp"]
__version__ = "0.1.0"
