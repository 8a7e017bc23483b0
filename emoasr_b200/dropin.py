"""Seams that make the UNMODIFIED reference run on the B200 kernels.

    from emoasr_b200 import dropin
    dropin.install(reference_root="/path/to/emoASR", precision="bf16")
    # ... then build asr.modeling.asr.ASR(params) / run asr/train_asr.py as usual

or as a launcher:   python -m emoasr_b200.dropin --reference /path/to/emoASR -- asr/train_asr.py -conf X.yaml

S1 module seam  ``warp_rnnt`` -> emoasr_b200.compat.warp_rnnt (mandatory: the import is at module top).
S2 class seam   asr.modeling.asr.RNNTDecoder / CTCDecoder and
                asr.modeling.decoders.rnn_transducer.CTCDecoder are rebound to subclasses of the
                reference's own classes whose forward() is the fused one.
S3 attribute    every CTCDecoder gets ``ctc_loss_fn`` = emoasr_b200.criteria.CTCLoss and, where the reference built
                one (distillation), ``forced_aligner`` = emoasr_b200.criteria.CTCForcedAligner.
"""
import os
import runpy
import sys


def install(reference_root=None, precision="bf16"):
    compat = os.path.join(os.path.dirname(os.path.abspath(__file__)), "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)                       # S1
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    import asr.modeling.asr as ref_asr                   # noqa: E402  (the reference)
    import asr.modeling.decoders.ctc as ref_ctc
    import asr.modeling.decoders.rnn_transducer as ref_rnnt

    from .criteria import CTCForcedAligner, CTCLoss
    from .decoders import FusedCTCForward, FusedRNNTForward, FusedRNNTSearch

    if getattr(ref_asr, "_emoasr_b200_installed", False):
        return ref_asr

    class CTCDecoder(FusedCTCForward, ref_ctc.CTCDecoder):
        fused_precision = precision

        def __init__(self, params):
            ref_ctc.CTCDecoder.__init__(self, params)
            self.ctc_loss_fn = CTCLoss(blank=self.blank_id, reduction="sum", zero_infinity=True)  # S3
            if hasattr(self, "forced_aligner"):                                                   # S3 (ctc.py:67,85)
                self.forced_aligner = CTCForcedAligner(blank_id=self.blank_id)

    class RNNTDecoder(FusedRNNTForward, FusedRNNTSearch, ref_rnnt.RNNTDecoder):
        fused_precision = precision

        def __init__(self, params, phase="train"):
            ref_rnnt.RNNTDecoder.__init__(self, params, phase)

    ref_rnnt.CTCDecoder = CTCDecoder                     # S2 (aux CTC built at rnn_transducer.py:62)
    ref_asr.CTCDecoder = CTCDecoder                      # S2 (asr.py:38)
    ref_asr.RNNTDecoder = RNNTDecoder                    # S2 (asr.py:40)
    ref_asr._emoasr_b200_installed = True
    return ref_asr


def main(argv=None):
    import argparse

    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="root of an emoASR checkout")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("script", help="reference script, e.g. asr/train_asr.py")
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    install(ns.reference, ns.precision)
    script = ns.script if os.path.isabs(ns.script) else os.path.join(ns.reference, ns.script)
    sys.argv = [script] + [a for a in ns.args if a != "--"]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
