"""pixflow-b200: B200-native (sm_100a) dense bidirectional optical flow + flow-guided blend behind the
OpticalFlowInterface / NovelViewGenerator interface of MungoMeng/Panorama-OpticalFlow.

The product is the CUDA library csrc/ -> libpixflow_b200.so (C-ABI: include/pixflow_b200.h); this package
is its Python host-side mirror.  Importing it does not need a GPU; creating an engine does.
"""
from .api import (DirectionHint, NovelViewGenerator, NovelViewGeneratorAsymmetricFlow, NovelViewUtil,  # noqa: F401
                  OpticalFlowInterface, PixFlow, PixFlowError, Stitchtools, four_input_frontend, makeOpticalFlowByName,
                  stitch_iteration)

__version__ = "0.1.0"
