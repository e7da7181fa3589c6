"""abcnet_b200 -- B200-native (sm_100a) implementation of the ABC-Net hot path.

Public surface (mirrors what the reference's scripts use):
  UNet(in_channels, heads)      drop-in for ``from unet import UNet`` (reference src/unet.py:77-119)
  PeakDecoder / records_to_lists  heat-map decoding (reference src/img2smiles.py:62-193)
  NativeUNet                    the eval forward behind the one-call C entry (abc_unet_create / abc_unet_forward_infer)
  InferGraph                    eval forward + decode of a fixed batch shape as one replayable CUDA graph
  SparseHeadsPipeline           opt-in fused inference + decode with the class heads evaluated at peaks only (bit-identical records)
  assemble_molblocks            native host assembly records -> MOL-block text (src/img2smiles.py:183-318, generate_smiles.py:18-105)
  HeatmapLoss                   fused training losses (reference src/train.py:95-137)
  TrainStep / make_optimizer    one whole training iteration (reference src/train.py:94-141), CUDA-graph replayable
  TargetRasteriser / parse_labels  dense training targets stamped on the GPU from label strings (src/utils.py:83-228)
  FusedAdam                     the reference's Adam (src/train.py:55) as one launch over all parameters
All compute goes through ``libabcnet_b200.so`` (``include/abcnet_b200.h``); there is no CPU fallback.
"""
from ._lib import LIB_PATH, launch_count, lib  # noqa: F401  (raises ImportError if the library is not built)
from .decode import PeakDecoder, assemble_molblocks, records_to_lists  # noqa: F401
from .infer_graph import InferGraph  # noqa: F401
from .loss import HeatmapLoss  # noqa: F401
from .native import NativeUNet  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .sparse import SparseHeadsPipeline  # noqa: F401
from .targets import TargetRasteriser, parse_labels  # noqa: F401
from .train_step import TrainStep, make_optimizer  # noqa: F401
from .unet import UNet  # noqa: F401

__all__ = ["UNet", "HeatmapLoss", "TrainStep", "make_optimizer", "FusedAdam", "PeakDecoder", "InferGraph", "NativeUNet", "records_to_lists", "assemble_molblocks", "SparseHeadsPipeline", "TargetRasteriser", "parse_labels", "launch_count", "LIB_PATH"]
