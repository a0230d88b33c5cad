"""B200-native micrograph denoiser: the inference path of Jeffrey-Ede/AI-CV-Automation-Elect-Micr
(atrous Xception encoder-decoder + crop-tile / normalise / stitch wrapper) on hand-written sm_100a
CUDA behind a C ABI.  See DESIGN.md."""
from .denoiser import Denoiser, scale0to1  # noqa: F401
from .engine import Engine  # noqa: F401
from . import weights  # noqa: F401
from . import sharding  # noqa: F401
from . import tfckpt  # noqa: F401
from . import micrograph_io  # noqa: F401
from . import quality  # noqa: F401
