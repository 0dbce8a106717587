"""polarcub_b200 -- B200-native batched polar-code engine (drop-in for polarcub's encode/decode entry points).

Host side: Python classes mirroring the reference's `BinaryPolarEncoderDecoder` / `QaryPolarEncoderDecoder`
(same constructor, `encode` / `decode` / `listDecode` signatures and return conventions) plus batched
array entry points.  Device side: hand-written sm_100a CUDA kernels behind the C-ABI in
include/polarcub_b200.h (polarcub_b200/csrc).  PyTorch is used only for device memory and streams.
"""
from . import engine  # noqa: F401
from ._lib import PolarcubError  # noqa: F401
from .BinaryPolarEncoderDecoder import BinaryPolarEncoderDecoder, polarTransformOfBits  # noqa: F401
from .simulation import encodeDecodeSimulation, genieEncodeDecodeSimulation, frozenSetFromTVAndPe, readFrozenSetFromFile  # noqa: F401
from . import Guardbands, CollectionOfBinaryTrellises, construction  # noqa: F401
from .construction import calcFrozenSet_degradingUpgrading, calcTVAndPe_degradingUpgrading  # noqa: F401
from .QaryPolarEncoderDecoder import QaryPolarEncoderDecoder, polarTransformOfQudits, ProbResult, irSimulation  # noqa: F401
from . import results_csv, channels  # noqa: F401

__all__ = ["BinaryPolarEncoderDecoder", "QaryPolarEncoderDecoder", "polarTransformOfBits", "polarTransformOfQudits",
           "ProbResult", "PolarcubError", "engine", "construction", "calcFrozenSet_degradingUpgrading", "calcTVAndPe_degradingUpgrading"]
