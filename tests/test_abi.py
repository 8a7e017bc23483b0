"""CPU: the C-ABI library builds/loads and exports every symbol include/emoasr_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from emoasr_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "emoasr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emo_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.EXPORTS)


def test_library_exports_every_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.emo_abi_version() == _lib.ABI_VERSION
    assert isinstance(lib.emo_last_error_string(), bytes)


def test_workspace_query_is_host_only():
    n = _lib.workspace_bytes(_lib.OP_RNNT_JOINT_BWD, _lib.PREC_FP32, 4, 50, 21, 64, 128)
    assert n > 0
    assert _lib.workspace_bytes(_lib.OP_RNNT_JOINT_FWD, _lib.PREC_FP32, 0, 50, 21, 64, 128) == 0


def test_bad_arguments_are_reported_not_crashed():
    lib = _lib.load()
    rc = lib.emo_rnnt_lattice_fwd_bwd(None, None, None, 1, 1, 1, None, None, None, None, None)
    assert rc == 1
    assert b"null" in lib.emo_last_error_string()


def test_ops_refuse_cpu_tensors():
    import pytest
    import torch
    import emoasr_b200 as E

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.ctc_loss(torch.zeros(1, 2, 3), torch.zeros(1, 1, dtype=torch.long), torch.tensor([2]), torch.tensor([1]))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.rnnt_loss(torch.zeros(1, 2, 2, 3), torch.zeros(1, 1, dtype=torch.int32), torch.tensor([2]), torch.tensor([1]))


def test_shape_support_query():
    """Host-only: the tensor-core mode accepts any vocabulary size (the reference's 10872 / 9798 are not multiples of
    32: padded inside the workspace) and any joint size up to 512 (the C ABI itself wants J % 128 == 0; the Python layer
    zero-pads other sizes); fp32 mode accepts everything."""
    import emoasr_b200.functional as F
    assert F.joint_supported("bf16", 8, 249, 61, 512, 10872)
    assert F.joint_supported("bf16", 8, 249, 61, 256, 9798)
    assert not F.joint_supported("bf16", 8, 249, 61, 640, 1024)
    assert F.joint_supported("bf16", 8, 249, 61, 320, 1024)            # run as J = 384 (functional._pad_hidden)
    assert not _lib.load().emo_rnnt_joint_supported(_lib.PREC_BF16, 8, 249, 61, 320, 1024)   # the raw entry point
    assert F.ctc_head_supported(8, 249, 144, 5000, 60) and not _lib.load().emo_ctc_head_supported(8, 249, 144, 5000, 60)
    assert F.joint_supported("fp32", 8, 249, 61, 640, 1000)
    # the padded vocabulary needs a slightly larger workspace than the next smaller multiple of 32
    a = _lib.workspace_bytes(_lib.OP_RNNT_JOINT_BWD, _lib.PREC_BF16, 2, 20, 8, 128, 96)
    b = _lib.workspace_bytes(_lib.OP_RNNT_JOINT_BWD, _lib.PREC_BF16, 2, 20, 8, 128, 100)
    assert b > a
