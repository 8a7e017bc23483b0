"""Module seam: put ``emoasr_b200/compat`` on sys.path and the reference's
``import warp_rnnt`` (asr/modeling/decoders/rnn_transducer.py:14) resolves to the B200 kernels.
Exposes exactly what the reference touches: ``rnnt_loss`` (:106-115) and ``__version__`` (:65)."""
from emoasr_b200.functional import rnnt_loss  # noqa: F401

__version__ = "emoasr_b200-compat-0.1.0"
