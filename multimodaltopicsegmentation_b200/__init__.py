"""multimodaltopicsegmentation_b200 -- B200-native (sm_100a) hot path of Ighina/MultimodalTopicSegmentation.

The package mirrors the reference's Python interface for the segmentation-model path
(`TextSegmenter`, the segmenter classes of models/CRF.py, the `EncoderDataset` batch layout) and runs
its arithmetic in hand-written CUDA kernels behind the C ABI of include/mts_b200.h.
"""
from . import _lib, ops, results_io  # noqa: F401
from .EncoderDataset import AudioPortionDataset, AudioPortionDatasetInference, DevicePrefetcher, to_device  # noqa: F401
from .lightning_model import TextSegmenter  # noqa: F401
from .load_datasets_precomputed import (ResidentDataset, cross_validation_split, load_dataset_for_inference,  # noqa: F401
                                        load_dataset_from_precomputed)
from .metrics import compute_Pk, compute_window_diff, get_boundaries  # noqa: F401
from .modules import CRF, RNN, BiLSTM, BiLSTMLateFusion, BiLSTMLateFusionCrf, BiRnnCrf  # noqa: F401
from .recurrent_longformer import RecurrentLongformer, RecurrentLongformerBlock  # noqa: F401

__version__ = "0.1.0"
