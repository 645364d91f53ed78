"""cns_b200: B200-native (sm_100a) implementation of the CNS hot path.

Correspondence-graph construction (`GraphGenerator`, `Midend`) and the recurrent GNN
servo policy (`GraphVS`) of hhcaz/CNS behind the reference's own Python API; the
arithmetic runs in `libcns_sm100.so` (hand-written CUDA, C ABI in include/cns_b200.h).
"""
from .data import Batch, GraphData
from .graph_gen import GraphGenerator
from .midend import CameraIntrinsic, Correspondence, Midend
from .model import GraphVS
from .servo import BatchedServo
from .stop_policy import BatchedPixelStop, ErrorHoldingStopPolicy, PixelStopPolicy, align_matches

__all__ = ["Batch", "GraphData", "GraphGenerator", "CameraIntrinsic", "Correspondence", "Midend", "GraphVS", "BatchedServo",
           "BatchedPixelStop", "ErrorHoldingStopPolicy", "PixelStopPolicy", "align_matches"]
