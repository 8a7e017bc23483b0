"""CPU oracle for the emoASR sequence-loss hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``emoasr_b200/`` imports
``oracle``; the product path fails loudly when the CUDA library is missing.

Contents
--------
rnnt_dp.py     fp64 numpy restatement of the RNN-T joint + transducer loss and its
               analytic gradients (follows asr/modeling/decoders/rnn_transducer.py:97-117,
               147-156 and the recursion of asr/modeling/decoders/rnnt_aligner.py:49-83,121-152).
ctc_dp.py      fp64 numpy restatement of Linear -> log_softmax -> CTCLoss(sum, zero_infinity)/B
               (asr/modeling/decoders/ctc.py:103-115; blank-extended path as
               asr/modeling/decoders/ctc_aligner.py:19-22).
ctc_align.py   float32 numpy restatement of the CTC forced aligner
               (asr/modeling/decoders/ctc_aligner.py:138-221), pinned by ref_ctc_forced_align.npz.
torch_path.py  the reference's own op sequence in torch fp32 on CPU
               (joint -> log_softmax -> rnnt_loss ; Linear -> log_softmax -> nn.CTCLoss).
               ``warp_rnnt`` (1ytic/warp-rnnt, version unpinned by the reference, CUDA-only,
               absent from the image) is replaced by torchaudio's CPU ``rnnt_loss`` with
               ``fused_log_softmax=False`` which has the same sparse-gradient contract.
lattice.c      plain-C fp64 restatement of both lattices (built by ``make -C oracle``).
warp_rnnt_shim.py / gen_golden.py
               import the *unmodified* reference from /root/reference on CPU and write the
               golden vectors committed under tests/golden/.

Parity pinning: the reference ships no tests and no golden vectors for this path
 -- "parity unpinned" by the reference itself.  The pins used here are
(1) outputs of the reference's own modules executed in the build container
(tests/golden/ref_*.npz, made by oracle/gen_golden.py), (2) the upstream warp-transducer
known-answer vector (cost 4.495666), (3) agreement between the fp64 DP, torch CTCLoss and
torchaudio rnnt_loss.
"""
