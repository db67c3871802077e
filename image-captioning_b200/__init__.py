"""B200-native per-RoI captioning path: PyramidROIAlign -> RoI head -> inject-LSTM decoder.

Host side is Python/PyTorch (device memory, streams, torch.distributed); all arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI of ``include/dcap.h`` (``libdcap.so``).
The directory is named ``image-captioning_b200``; import it as ``image_captioning_b200``.
"""
from . import _lib                                            # noqa: F401
from .roi_align import (PyramidROIAlign, pyramid_roi_align, fpn_levels, pyramid_roi_align_backward,   # noqa: F401
                        pyramid_roi_align_autograd)
from . import synth  # noqa: F401
from .text_model import (DenseCapConfig, build_lstm_model, build_model, RoiCaptionModel,   # noqa: F401
                         InjectModelV2, Adam, roi_caption_loss)
from .postprocess import refine_generations, caption_text    # noqa: F401
from .proposals import ProposalLayer, ProposalConfig, generate_pyramid_anchors, normalize_boxes   # noqa: F401
from . import parallel    # noqa: F401
from . import data    # noqa: F401
from . import callbacks    # noqa: F401
