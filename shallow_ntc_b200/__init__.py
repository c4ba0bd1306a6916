"""shallow_ntc_b200 -- B200-native decode path of mandt-lab/shallow-ntc.

Host side: Python + ctypes over the C ABI in ``include/sntc.h`` (``libsntc.so``: hand-written sm_100a
CUDA).  No PyTorch, no TensorFlow, no CPU fallback.  ``transforms.class_builder`` mirrors the
reference's transform registry; ``models.Model.decompress`` is the fused decode entry.
"""
from .transforms import class_builder, ClassBuilder  # noqa: F401
from .models import Model, FactorizedModel, CONFIGS, build_config  # noqa: F401
from .tensors import Context, DeviceArray, as_tensor  # noqa: F401
from .pipeline import DecodePipeline  # noqa: F401
from .codec import EntropyCoder  # noqa: F401
from . import codec, eval_lib  # noqa: F401
from .lpips import Lpips  # noqa: F401
from ._lib import SntcError, EXPORTED_SYMBOLS, LIB_PATH  # noqa: F401

__all__ = ["class_builder", "ClassBuilder", "Model", "FactorizedModel", "CONFIGS", "build_config", "Context",
           "DeviceArray", "as_tensor", "DecodePipeline", "EntropyCoder", "Lpips", "codec", "eval_lib", "SntcError", "EXPORTED_SYMBOLS", "LIB_PATH"]
