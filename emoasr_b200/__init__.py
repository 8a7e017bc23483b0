"""emoasr_b200 -- B200 (sm_100a) implementation of emoASR's sequence-loss hot path.

Host side mirrors the reference's Python surface (asr/criteria.py-style loss modules, the
RNNTDecoder / CTCDecoder classes of asr/modeling/decoders, a warp_rnnt-compatible entry point);
all arithmetic is done by hand-written CUDA kernels behind the C ABI in include/emoasr_b200.h.
"""
from . import _lib  # noqa: F401
from .functional import (ctc_forced_align, ctc_head_loss, ctc_loss, rnnt_forced_align, rnnt_joint_loss, rnnt_joint_loss_from_outputs,  # noqa: F401
                         rnnt_joint_outputs, rnnt_loss)
from .criteria import (CTCForcedAligner, CTCHeadLoss, CTCLoss, RNNTAlignDistillLoss, RNNTForcedAligner, RNNTJointFullLoss,  # noqa: F401
                       RNNTJointLoss,
                       RNNTLoss, RNNTWordDistillLoss)

__version__ = "0.1.0"
