"""A stand-in for the absent third-party ``warp_rnnt`` module so that the UNMODIFIED reference can
be imported on CPU in the build container (ORACLE - test infrastructure).

The reference imports ``warp_rnnt`` at module top (asr/modeling/decoders/rnn_transducer.py:14) and
calls ``warp_rnnt.rnnt_loss`` once (:106-115) and ``warp_rnnt.__version__`` (:65).  Use::

    from oracle import warp_rnnt_shim; warp_rnnt_shim.install()

before importing ``asr.modeling.*`` from /root/reference.
"""
import sys
import types

from .torch_path import rnnt_loss_from_log_probs


def install():
    mod = types.ModuleType("warp_rnnt")
    mod.rnnt_loss = rnnt_loss_from_log_probs
    mod.__version__ = "oracle-shim(torchaudio.rnnt_loss, fused_log_softmax=False)"
    sys.modules["warp_rnnt"] = mod
    return mod
