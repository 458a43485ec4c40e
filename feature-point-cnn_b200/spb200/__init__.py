"""spb200 - B200-native SuperPoint / MagicPoint inference (Python side).

Mirrors the reference's Python entry points (python/src/superpoint.py, inferencewrapper.py,
netutils.py, settings.py) on top of the C ABI in libspb200.so.  PyTorch is used for device
tensors and streams only.
"""
from .settings import SuperPointSettings          # noqa: F401
from .engine import Engine, Spb200Error           # noqa: F401
from .superpoint import SuperPoint                # noqa: F401
from .inferencewrapper import InferenceWrapper    # noqa: F401
from .netutils import get_points, get_descriptors, restore_prob_map, heatmap_from_logits   # noqa: F401
from . import ops                                  # noqa: F401  (registers torch.ops.spb200.*)
