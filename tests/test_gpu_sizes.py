"""GPU parity at the BASELINE sizes and in the corners the small random cases do not reach: cfg-4 lattice and
joint, cfg-1 CTC, peaked logits, the drop-in decoders in bf16 mode against the reference's goldens, the phone /
intermediate CTC heads, run-to-run determinism.  All through the C ABI (emoasr_b200.functional / decoders).
"""
from collections import namedtuple

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

LOSS_RTOL, GRAD_RTOL = 1e-5, 1e-4            # fp32 mode (north star)
BF16_LOSS_RTOL, BF16_GRAD_RTOL = 2e-3, 2e-2  # stated bf16-operand tolerance

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def dev():
    return torch.device("cuda:0")


def T_(a, dtype=None):
    t = torch.as_tensor(np.asarray(a)).to(dev())
    return t.to(dtype) if dtype is not None else t


# ---------------------------------------------------------------- cfg 4: T=1000, U=400, V=4096
def test_lattice_cfg4_vs_c_oracle():
    """cfg-4 lattice size (T=1000, U1=401: 13 warps per lattice CTA, 1400 diagonals) against the plain-C
    restatement (oracle/lattice.c, fp64), full and ragged lengths."""
    import emoasr_b200 as E
    from oracle import clattice
    B, T, U = 2, 1000, 400
    rng = np.random.default_rng(4)
    # log-probs of a plausible joint: blank ~ -0.3 .. -2, label ~ -1 .. -6
    lp2 = np.stack([-(rng.random((B, T, U + 1)) * 1.7 + 0.3), -(rng.random((B, T, U + 1)) * 5 + 1)], -1).astype(np.float32)
    tl, ul = np.array([T, 733]), np.array([U, 257])
    x = T_(lp2).requires_grad_()
    costs = E.rnnt_loss(x, None, T_(tl), T_(ul), gather=True)
    costs.sum().backward()
    g = -x.grad.cpu().numpy()
    for b in range(B):
        cost_ref, g_ref = clattice.rnnt_lattice(lp2[b], tl[b], ul[b])
        assert abs(float(costs[b]) - cost_ref) <= LOSS_RTOL * abs(cost_ref)
        # fp32 log-domain posteriors: alpha+beta-ll cancels numbers of magnitude ~3e3 (ulp 2.4e-4)
        assert rel_err(g[b], g_ref) < 2e-3
        assert np.all(g[b, tl[b]:] == 0) and np.all(g[b, :, ul[b] + 1:] == 0)


def test_joint_cfg4_shape_vs_fp32_mode():
    """One cfg-4-shaped joint step (T=1000, U=400, V=4096, J=512; B=2, second utterance shorter): loss and all
    four gradients of the tensor-core path against the fp32 mode (itself pinned to the reference at 1e-5 / 1e-4)."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(44)
    B, T, U, V, J = 2, 1000, 400, 4096, 512
    enc = torch.randn(B, T, J, generator=gen).to(dev())
    dec_ = torch.randn(B, U + 1, J, generator=gen).to(dev())
    w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev())
    bo = (0.1 * torch.randn(V, generator=gen)).to(dev())
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    tl, ul = torch.tensor([T, 611], device=dev()), torch.tensor([U, 333], device=dev())
    out = {}
    for prec in ("fp32", "bf16"):
        te = [t.clone().requires_grad_() for t in (enc, dec_, w, bo)]
        loss = E.rnnt_joint_loss(*te, ys, tl, ul, blank=0, reduction="mean", precision=prec)
        loss.backward()
        out[prec] = (float(loss), [t.grad for t in te])
    assert abs(out["bf16"][0] - out["fp32"][0]) <= BF16_LOSS_RTOL * abs(out["fp32"][0])
    for got, ref, k in zip(out["bf16"][1], out["fp32"][1], ("d_enc", "d_dec", "d_w_out", "d_b_out")):
        assert float((got - ref).norm() / ref.norm()) < BF16_GRAD_RTOL, k
    d_enc, d_dec = out["bf16"][1][:2]
    assert float(d_enc[1, 611:].abs().sum()) == 0.0 and float(d_dec[1, 334:].abs().sum()) == 0.0


# ---------------------------------------------------------------- cfg 1: CTC with V = 10872
def test_ctc_cfg1_vs_torch_fp64():
    import emoasr_b200 as E
    B, T, V, U = 8, 249, 10872, 60
    gen = torch.Generator().manual_seed(1)
    logits = torch.randn(B, T, V, generator=gen).to(dev()).requires_grad_()
    ys = torch.randint(4, V, (B, U), generator=gen)
    tl = torch.tensor([249, 249, 230, 201, 180, 150, 97, 61])
    ul = torch.tensor([60, 41, 60, 33, 12, 55, 40, 60])
    nll = E.ctc_loss(logits, ys, tl, ul, blank=0)
    (nll.sum() / B).backward()
    def torch_ctc(dtype):
        x = logits.detach().to(dtype).requires_grad_()
        l = torch.nn.functional.ctc_loss(x.transpose(0, 1).log_softmax(2), ys.to(dev()), tl.to(dev()), ul.to(dev()),
                                         blank=0, reduction="none", zero_infinity=True)
        (l.sum() / B).backward()
        return l.detach(), x.grad
    ref, ref_g = torch_ctc(torch.float64)
    _, t32_g = torch_ctc(torch.float32)
    assert torch.allclose(nll.double(), ref, rtol=LOSS_RTOL, atol=1e-3)
    # at T = 249 every fp32 log-domain lattice (torch's own CUDA kernel included) sits at a few 1e-4 of the fp64
    # gradient: alpha + beta - nll cancels numbers of magnitude ~2e3.  Bar: 1e-4, or no worse than 2x torch fp32.
    err_ours = float((logits.grad.double() - ref_g).norm() / ref_g.norm())
    err_torch32 = float((t32_g.double() - ref_g).norm() / ref_g.norm())
    assert err_ours < max(GRAD_RTOL, 2 * err_torch32), (err_ours, err_torch32)
    assert float(logits.grad[7, 61:].abs().sum()) == 0.0


# ---------------------------------------------------------------- peaked logits (the regime of a trained model)
def test_joint_bf16_peaked_logits():
    """w_out scaled x8: |z| reaches ~40, softmax rows are close to one-hot.  Exercises exp2 of large negative
    arguments and the exact blank / label entries of dz."""
    import emoasr_b200 as E
    from oracle import rnnt_dp
    B, T, U, V, J = 3, 40, 15, 1024, 512
    rng = np.random.default_rng(8)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    enc, dec_ = f(B, T, J), f(B, U + 1, J)
    w_out, b_out = f(V, J) * (16.0 / np.sqrt(J)), f(V) * 0.5
    ys = rng.integers(1, V, (B, U))
    tl, ul = np.array([T, 31, 17]), np.array([U, 15, 4])
    eye = np.eye(J, dtype=np.float32)
    _, _, _, z = rnnt_dp.joint_logits(enc, dec_, eye, np.zeros(J), eye, np.zeros(J), w_out, b_out)
    assert np.abs(z).max() > 30
    r = rnnt_dp.joint_loss_and_grads(enc, dec_, eye, np.zeros(J), eye, np.zeros(J), w_out, b_out, ys, tl, ul)
    te = [T_(a).requires_grad_() for a in (enc, dec_, w_out, b_out)]
    loss = E.rnnt_joint_loss(*te, T_(ys), T_(tl), T_(ul), blank=0, reduction="mean", precision="bf16")
    loss.backward()
    # |z| ~ 40 with 8-bit-mantissa operands: the logits themselves carry ~0.1 absolute error, i.e. ~1e-3 of a
    # loss of ~1e3 -- still inside the stated tolerance
    assert abs(float(loss) - r["loss"]) <= BF16_LOSS_RTOL * abs(r["loss"])
    for t, k in zip(te, ["d_enc_proj", "d_dec_proj", "d_w_out", "d_b_out"]):
        assert np.isfinite(t.grad.cpu().numpy()).all()
        assert rel_err(t.grad.cpu().numpy(), r[k]) < 4 * BF16_GRAD_RTOL, k   # one-hot rows amplify logit noise


# ---------------------------------------------------------------- drop-in decoders, bf16 mode, reference goldens
def _params_from_golden(g, keys):
    d = {k: g["hp." + k].item() for k in keys}
    return namedtuple("Params", d.keys())(**d)


RNNT_KEYS = ["dec_num_layers", "dec_hidden_size", "embedding_size", "joint_hidden_size", "enc_hidden_size",
             "vocab_size", "eos_id", "blank_id", "mtl_ctc_weight", "kd_weight", "dropout_emb_rate",
             "dropout_dec_rate"]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["ref_rnnt_tcshape_ragged", "ref_rnnt_tcshape_auxctc", "ref_rnnt_tcfull_auxctc"])
def test_rnnt_decoder_tensor_core_shape_vs_reference_golden(name, precision):
    """Goldens of the UNMODIFIED reference at a shape the tensor-core kernels accept (J=128, V=64): the drop-in
    decoder must reproduce them at 1e-5 / 1e-4 in fp32 mode and at the stated bf16 tolerance in bf16 mode (the
    default of dropin.install())."""
    from emoasr_b200.decoders import RNNTDecoder
    g = load_golden(name)
    dec = RNNTDecoder(_params_from_golden(g, RNNT_KEYS), phase="test")
    dec.fused_precision = precision
    dec.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    dec = dec.to(dev()).train()
    eouts = T_(g["eouts"]).requires_grad_()
    loss, loss_dict, logits = dec(eouts, T_(g["elens"]), None, T_(g["ys"]), T_(g["ylens"]), T_(g["ys_in"]), T_(g["ys_out"]))
    loss.backward()
    ltol, gtol = (LOSS_RTOL, GRAD_RTOL) if precision == "fp32" else (BF16_LOSS_RTOL, BF16_GRAD_RTOL)
    assert logits is None
    assert abs(float(loss) - float(g["loss_total"])) <= ltol * abs(float(g["loss_total"]))
    assert abs(float(loss_dict["loss_rnnt"]) - float(g["lossdict.loss_rnnt"])) <= ltol * abs(float(g["lossdict.loss_rnnt"]))
    assert rel_err(eouts.grad.cpu().numpy(), g["grad_eouts"]) < gtol
    for k, v in dec.named_parameters():
        ref = g["grad." + k]
        if ref.size == 0:
            continue
        assert rel_err(v.grad.cpu().numpy(), ref) < gtol, k


@pytest.mark.parametrize("name", ["ref_ctc_phone_final", "ref_ctc_phone_hie_inter"])
def test_ctc_decoder_phone_and_inter_heads_vs_reference_golden(name):
    """Phone CTC on the final / intermediate layer (ctc.py:129-148) and intermediate CTC (ctc.py:150-170)."""
    from emoasr_b200.decoders import CTCDecoder
    g = load_golden(name)
    keys = ["enc_hidden_size", "vocab_size", "eos_id", "blank_id", "kd_weight", "mtl_phone_ctc_weight",
            "hie_mtl_phone", "phone_vocab_size", "mtl_inter_ctc_weight"]
    dec = CTCDecoder(_params_from_golden(g, keys))
    dec.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    dec = dec.to(dev())
    eouts, eouts_inter = T_(g["eouts"]).requires_grad_(), T_(g["eouts_inter"]).requires_grad_()
    loss, loss_dict, logits = dec(eouts, T_(g["elens"]), eouts_inter, T_(g["ys"]), T_(g["ylens"]), None, None, None,
                                  T_(g["ps"]), T_(g["plens"]))
    loss.backward()
    assert abs(float(loss) - float(g["loss_total"])) <= LOSS_RTOL * abs(float(g["loss_total"]))
    ref_keys = sorted(k[len("lossdict."):] for k in g if k.startswith("lossdict."))
    assert sorted(loss_dict) == ref_keys
    for k in ref_keys:
        assert abs(float(loss_dict[k]) - float(g["lossdict." + k])) <= LOSS_RTOL * abs(float(g["lossdict." + k])), k
    assert rel_err(eouts.grad.cpu().numpy(), g["grad_eouts"]) < GRAD_RTOL
    if g["grad_eouts_inter"].size:
        assert rel_err(eouts_inter.grad.cpu().numpy(), g["grad_eouts_inter"]) < GRAD_RTOL
    for k, v in dec.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), g["grad." + k]) < GRAD_RTOL, k


# ---------------------------------------------------------------- run-to-run determinism
def test_joint_bf16_repeatability():
    """50 repeats of the same step: the forward outputs (cost) are bit-identical; the gradients are sums over
    CTAs combined with red.global.add in arrival order, so they may differ in the last bits only (<= 1e-6
    relative).  A race between roles of the ring kernel would show up here as a large or growing difference."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(3)
    B, T, U, V, J = 6, 120, 40, 1024, 512
    enc = torch.randn(B, T, J, generator=gen).to(dev())
    dec_ = torch.randn(B, U + 1, J, generator=gen).to(dev())
    w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev())
    bo = (0.1 * torch.randn(V, generator=gen)).to(dev())
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    tl = torch.tensor([120, 120, 101, 77, 50, 9], device=dev())
    ul = torch.tensor([40, 33, 40, 12, 0, 7], device=dev())
    ref = None
    for it in range(50):
        te = [t.clone().requires_grad_() for t in (enc, dec_, w, bo)]
        costs = E.rnnt_joint_loss(*te, ys, tl, ul, blank=0, precision="bf16")
        costs.mean().backward()
        cur = (costs.detach().clone(), [t.grad.clone() for t in te])
        if ref is None:
            ref = cur
            continue
        assert torch.equal(cur[0], ref[0]), it
        for a, b_, k in zip(cur[1], ref[1], ("d_enc", "d_dec", "d_w_out", "d_b_out")):
            assert float((a - b_).norm() / b_.norm()) <= 1e-6, (it, k)


@pytest.mark.timeout(180, method="thread")
def test_two_steps_on_two_streams_at_once():
    """The ring kernel's roles wait for each other, so all of its CTAs must be resident together; two instances
    enqueued on two streams at the same time must not be interleaved on the SMs (cooperative launch).  Results equal
    the sequential runs; a deadlock would trip the timeout."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(0)
    B, T, U, V, J = 4, 120, 40, 1024, 512
    tl, ul = torch.full((B,), T, device=dev()), torch.full((B,), U, device=dev())

    def make():
        return [torch.randn(B, T, J, generator=gen).to(dev()), torch.randn(B, U + 1, J, generator=gen).to(dev()),
                (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev()), torch.zeros(V, device=dev()),
                torch.randint(1, V, (B, U), generator=gen).to(dev())]

    def step(d):
        te = [t.clone().requires_grad_() for t in d[:4]]
        loss = E.rnnt_joint_loss(*te, d[4], tl, ul, blank=0, reduction="mean", precision="bf16")
        loss.backward()
        return [loss.detach()] + [t.grad for t in te]

    a, b = make(), make()
    ref_a, ref_b = step(a), step(b)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    for _ in range(10):
        with torch.cuda.stream(s1):
            ra = step(a)
        with torch.cuda.stream(s2):
            rb = step(b)
        torch.cuda.synchronize()
        for got, ref in ((ra, ref_a), (rb, ref_b)):
            for x, y in zip(got, ref):
                assert float((x - y).norm() / y.norm().clamp_min(1e-30)) < 1e-5


def test_forward_and_backward_keep_nothing_of_size_N_x_V():
    """Peak device memory of a whole training step stays far below one N x V tensor (even at 2 bytes per entry):
    the logits are neither cached by the forward nor formed by the backward."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(5)
    B, T, U, V, J = 8, 200, 60, 4096, 512
    enc = torch.randn(B, T, J, generator=gen).to(dev()).requires_grad_()
    dec_ = torch.randn(B, U + 1, J, generator=gen).to(dev()).requires_grad_()
    w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev()).requires_grad_()
    bo = torch.zeros(V, device=dev(), requires_grad=True)
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    tl, ul = torch.full((B,), T, device=dev()), torch.full((B,), U, device=dev())
    zbytes = B * T * (U + 1) * V * 2
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    loss = E.rnnt_joint_loss(enc, dec_, w, bo, ys, tl, ul, reduction="mean", precision="bf16")
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss)
    assert torch.cuda.max_memory_allocated() - base < zbytes // 2


# ---------------------------------------------------------------- fused CTC head (Linear + log_softmax + CTC)
HEAD_SHAPES = [
    # B, T, He, V, U
    (3, 40, 128, 64, 9),       # one vocab chunk
    (4, 61, 256, 1000, 17),    # vocabulary not a multiple of 32, several chunks
    (2, 300, 256, 5000, 80),   # cfg-2 head, more than one 256-frame pair tile per utterance
    (2, 37, 512, 2048, 12),    # widest supported hidden size
    (3, 50, 144, 300, 11),     # encoder width not a multiple of 128: zero-padded to 256 by functional.ctc_head_loss
]


@pytest.mark.parametrize("B,T,He,V,U", HEAD_SHAPES)
def test_ctc_head_vs_oracle(B, T, He, V, U):
    """Fused head (tensor-core path) against the reference's op sequence Linear -> log_softmax -> nn.CTCLoss in
    fp64 on the CPU (oracle/torch_path.py), ragged lengths, repeated labels, at the stated bf16 tolerance."""
    import emoasr_b200 as E
    from oracle import torch_path
    gen = torch.Generator().manual_seed(B * 100 + V)
    eouts = torch.randn(B, T, He, generator=gen)
    w = torch.randn(V, He, generator=gen) * (2.0 / He ** 0.5)
    bias = torch.randn(V, generator=gen) * 0.5
    ys = torch.randint(1, V, (B, U), generator=gen)
    ys[:, 3] = ys[:, 2]                                   # a repeated label (needs the blank between)
    tl = torch.randint(max(2 * U + 2, T // 2), T + 1, (B,), generator=gen); tl[0] = T
    ul = torch.randint(1, U + 1, (B,), generator=gen); ul[0] = U
    ref_in = [t.double().requires_grad_() for t in (eouts, w, bias)]
    ref = torch_path.ctc_head_loss(*ref_in, ys, tl, ul, blank=0)
    ref.backward()
    te = [t.to(dev()).requires_grad_() for t in (eouts, w, bias)]
    nll = E.ctc_head_loss(*te, ys.to(dev()), tl.to(dev()), ul.to(dev()), blank=0)
    loss = nll.sum() / B
    loss.backward()
    assert abs(float(loss) - float(ref)) <= BF16_LOSS_RTOL * abs(float(ref))
    for got, want, k in zip(te, ref_in, ("d_eouts", "d_weight", "d_bias")):
        assert rel_err(got.grad.cpu().numpy(), want.grad.numpy()) < BF16_GRAD_RTOL, k
    b = int(torch.argmin(tl))
    assert float(te[0].grad[b, int(tl[b]):].abs().sum()) == 0.0       # padded frames


def test_ctc_head_infeasible_and_matches_unfused():
    """zero_infinity: an utterance whose labels do not fit its frames contributes 0 loss and 0 gradient; the fused
    head agrees with the unfused route (cuBLAS Linear + emo_ctc_fwd/bwd in fp32) on the rest."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(9)
    B, T, He, V, U = 4, 50, 256, 512, 20
    eouts = torch.randn(B, T, He, generator=gen).to(dev())
    lin = torch.nn.Linear(He, V).to(dev())
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    tl = torch.tensor([50, 12, 50, 33], device=dev())     # utterance 1: 12 frames < 20 labels
    ul = torch.tensor([20, 20, 0, 7], device=dev())       # utterance 2: empty transcript
    out = {}
    for name in ("fused", "unfused"):
        lin.zero_grad(set_to_none=True)
        x = eouts.clone().requires_grad_()
        if name == "fused":
            nll = E.ctc_head_loss(x, lin.weight, lin.bias, ys, tl, ul, blank=0)
        else:
            nll = E.ctc_loss(lin(x), ys, tl, ul, blank=0)
        (nll.sum() / B).backward()
        out[name] = (nll.detach().clone(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone())
    f, u = out["fused"], out["unfused"]
    assert float(f[0][1]) == 0.0 and float(f[1][1].abs().sum()) == 0.0
    assert torch.allclose(f[0], u[0], rtol=BF16_LOSS_RTOL, atol=1e-3)
    for a, b_, k in zip(f[1:], u[1:], ("d_eouts", "d_weight", "d_bias")):
        assert float((a - b_).norm() / b_.norm()) < BF16_GRAD_RTOL, k


@pytest.mark.parametrize("name", ["ref_ctc_tchead_ragged", "ref_ctc_tchead_phone_hie_inter"])
def test_ctc_decoder_fused_head_vs_reference_golden(name):
    """The drop-in CTCDecoder with fused_precision="bf16" (every head fused from eouts where the shape allows)
    against the goldens of the unmodified reference, at the stated bf16 tolerance."""
    from emoasr_b200.decoders import CTCDecoder
    g = load_golden(name)
    keys = ["enc_hidden_size", "vocab_size", "eos_id", "blank_id", "kd_weight", "mtl_phone_ctc_weight",
            "hie_mtl_phone", "phone_vocab_size", "mtl_inter_ctc_weight"]
    keys = [k for k in keys if "hp." + k in g]
    dec = CTCDecoder(_params_from_golden(g, keys))
    dec.fused_precision, dec.fused_head_min_frames = "bf16", 0
    dec.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    dec = dec.to(dev())
    eouts = T_(g["eouts"]).requires_grad_()
    inter = T_(g["eouts_inter"]).requires_grad_() if "eouts_inter" in g and g["eouts_inter"].size else None
    ps = T_(g["ps"]) if "ps" in g else None
    plens = T_(g["plens"]) if "plens" in g else None
    loss, loss_dict, logits = dec(eouts, T_(g["elens"]), inter, T_(g["ys"]), T_(g["ylens"]), None, None, None, ps, plens)
    loss.backward()
    assert logits is None                      # the fused head ran: no (B,T,V) tensor exists
    assert abs(float(loss) - float(g["loss_total"])) <= BF16_LOSS_RTOL * abs(float(g["loss_total"]))
    assert rel_err(eouts.grad.cpu().numpy(), g["grad_eouts"]) < BF16_GRAD_RTOL
    for k, v in dec.named_parameters():
        assert rel_err(v.grad.cpu().numpy(), g["grad." + k]) < BF16_GRAD_RTOL, k


# ---------------------------------------------------------------- decode-time joint / greedy search
def test_joint_step_vs_dense_joint():
    """emo_rnnt_joint_step (one launch per search step) against the dense joint restricted to single cells, fp32:
    logits and argmax for N rows, an enc-row index (greedy: row = b*T + t), one frame against a beam of rows, and
    more rows than one pass of the kernel holds (N > 32)."""
    from emoasr_b200 import functional as F
    gen = torch.Generator().manual_seed(21)
    for N, J, V in ((5, 64, 131), (32, 512, 1024), (70, 256, 1000)):
        rows = 3 * N
        enc = torch.randn(rows, J, generator=gen).to(dev())
        dec_ = torch.randn(N, J, generator=gen).to(dev())
        w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev())
        b = torch.randn(V, generator=gen).to(dev())
        idx = torch.randint(0, rows, (N,), generator=gen).int().to(dev())
        logits, tok = F.joint_step(enc, dec_, w, b, enc_row=idx, want_logits=True, want_token=True)
        ref = torch.tanh(enc[idx.long()].double() + dec_.double()) @ w.double().t() + b.double()
        assert float((logits.double() - ref).abs().max()) < 1e-4
        assert torch.equal(tok, ref.argmax(-1))
        # token only, workspace reused across calls (the kernel leaves it zeroed)
        ws = F.step_workspace(N, dev())
        for _ in range(3):
            _, tok2 = F.joint_step(enc, dec_, w, b, enc_row=idx, want_logits=False, want_token=True, ws=ws)
            assert torch.equal(tok2, tok)
        assert int(ws.sum()) == 0


def test_greedy_search_vs_reference_golden():
    """Batched on-device greedy search against hypotheses AND per-step alignments of the unmodified reference's
    _greedy (rnn_transducer.py:194-240): utterances that emit up to the max_seq_len break, an all-blank one, T_b = 1."""
    from emoasr_b200.decoders import RNNTDecoder
    g = load_golden("ref_rnnt_greedy")
    dec = RNNTDecoder(_params_from_golden(g, RNNT_KEYS), phase="test")
    dec.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    dec = dec.to(dev()).eval()
    hyps, scores, logits, aligns = dec._greedy(T_(g["eouts"]), T_(g["elens"]))
    for b in range(int(g["n"])):
        assert hyps[b] == g[f"hyp.{b}"].tolist(), b
        assert aligns[b] == g[f"align.{b}"].tolist(), b
    # the reference's decode() contract (rnn_transducer.py:327-347)
    out = dec.decode(T_(g["eouts"]), T_(g["elens"]), beam_width=1)
    assert out[0] == hyps and out[1:] == (None, None, None)
    # joint() on single cells (what the reference's own search loops call) runs through the step kernel
    e, d = T_(g["eouts"])[:1, 3:4], torch.randn(4, 1, dec.w_dec.in_features, device=dev())
    with torch.no_grad():
        fused = dec.joint(e, d)
        dense = dec._dense_joint(e, d)
    assert fused.shape == dense.shape and float((fused - dense).abs().max()) < 1e-4


# ---------------------------------------------------------------- forced aligner + distillation on the fused lattice
def test_forced_aligner_vs_reference_smoke_fixture():
    """emoasr_b200.RNNTForcedAligner (dense log-probs in, like rnnt_aligner.py:155-198) against what the unmodified
    Numba aligner returned on its own smoke input (rnnt_aligner.py:201-208)."""
    import emoasr_b200 as E
    g = load_golden("ref_rnnt_aligner_smoke")
    al = E.RNNTForcedAligner(blank_id=0)(T_(g["log_probs"]), T_(g["T"]), T_(g["labels"]), T_(g["U"]))
    assert al.dtype == torch.int32 and np.array_equal(al.cpu().numpy(), g["aligns"])


KD_KEYS = RNNT_KEYS + ["kd_type", "reduce_main_loss_kd"]


@pytest.mark.parametrize("name", ["ref_rnnt_kd_word", "ref_rnnt_kd_align"])
def test_rnnt_decoder_distillation_vs_reference_golden(name):
    """kd_weight > 0 (rnn_transducer.py:127-141) WITHOUT the dense logits: the word loss from the fused joint's
    differentiable per-cell lse, the align loss from the forced alignment computed on the loss's own lattice.
    fp32 mode against the unmodified reference: loss 1e-5, gradients 1e-4."""
    from emoasr_b200.decoders import RNNTDecoder
    g = load_golden(name)
    dec = RNNTDecoder(_params_from_golden(g, KD_KEYS), phase="train")
    dec.fused_precision = "fp32"
    dec.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("param.")})
    dec = dec.to(dev()).train()
    eouts = T_(g["eouts"]).requires_grad_()
    loss, loss_dict, logits = dec(eouts, T_(g["elens"]), None, T_(g["ys"]), T_(g["ylens"]), T_(g["ys_in"]),
                                  T_(g["ys_out"]), T_(g["soft_labels"]))
    loss.backward()
    assert logits is None
    for k in ("loss_rnnt", "loss_kd", "loss_total"):
        assert abs(float(loss_dict[k]) - float(g["lossdict." + k])) <= LOSS_RTOL * abs(float(g["lossdict." + k])), k
    assert rel_err(eouts.grad.cpu().numpy(), g["grad_eouts"]) < GRAD_RTOL
    for k, v in dec.named_parameters():
        ref = g["grad." + k]
        if ref.size == 0:
            continue
        assert rel_err(v.grad.cpu().numpy(), ref) < GRAD_RTOL, k
    if "aligns" in g:
        from emoasr_b200 import functional as F
        with torch.no_grad():
            douts, _ = dec.recurrency(T_(g["ys_in"]), None)
            _, _, al = F.rnnt_joint_outputs(dec.w_enc(T_(g["eouts"])), dec.w_dec(douts), dec.output.weight,
                                            dec.output.bias, T_(g["ys"]), T_(g["elens"]), T_(g["ylens"]),
                                            precision="fp32", aligns=True)
        assert np.array_equal(al.cpu().numpy(), g["aligns"])


def test_lse_output_gradient_bf16_vs_fp32_mode():
    """The tensor-core backward with a gradient on the lse output (grad_lse * softmax(z) added to dz in the ring
    kernel's epilogue) against the fp32 mode, plus a finite-difference-free identity: with cost weight 0 and
    grad_lse = 1 on valid cells, d_b_out = sum over valid cells of softmax(z) -> sums to the number of valid cells."""
    from emoasr_b200 import functional as F
    gen = torch.Generator().manual_seed(31)
    B, T, U, V, J = 3, 30, 11, 512, 256
    enc = torch.randn(B, T, J, generator=gen).to(dev())
    dec_ = torch.randn(B, U + 1, J, generator=gen).to(dev())
    w = (torch.randn(V, J, generator=gen) / J ** 0.5).to(dev())
    bo = (0.1 * torch.randn(V, generator=gen)).to(dev())
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    tl, ul = torch.tensor([30, 21, 9], device=dev()), torch.tensor([11, 6, 11], device=dev())
    wgt = torch.rand(B, T, U + 1, generator=gen).to(dev())
    out = {}
    for prec in ("fp32", "bf16"):
        te = [t.clone().requires_grad_() for t in (enc, dec_, w, bo)]
        costs, lse, _ = F.rnnt_joint_outputs(*te, ys, tl, ul, precision=prec)
        (costs.mean() + (wgt * lse).sum() / 50).backward()
        out[prec] = [t.grad for t in te]
    for a, b_, k in zip(out["bf16"], out["fp32"], ("d_enc", "d_dec", "d_w_out", "d_b_out")):
        assert float((a - b_).norm() / b_.norm()) < BF16_GRAD_RTOL, k
    te = [t.clone().requires_grad_() for t in (enc, dec_, w, bo)]
    costs, lse, _ = F.rnnt_joint_outputs(*te, ys, tl, ul, precision="bf16")
    lse.sum().backward()
    n_valid = float((tl * (ul + 1)).sum())
    assert abs(float(te[3].grad.sum()) - n_valid) < 2e-2 * n_valid


# ---------------------------------------------------------------- projections folded into the library
FULL_SHAPES = [
    # B, T, U, He, Hd, J, V
    (2, 20, 7, 64, 48, 128, 96),        # one output tile everywhere, K tails (48 = 0.75 K block)
    (3, 40, 15, 256, 512, 512, 1024),   # cfg-3 layer sizes: two 256-column tiles in the forward projections
    (2, 150, 30, 144, 320, 256, 1000),  # hidden sizes that are not multiples of 64, odd vocabulary, split-K weight grads
]


@pytest.mark.parametrize("B,T,U,He,Hd,J,V", FULL_SHAPES)
def test_joint_from_outputs_vs_oracle(B, T, U, He, Hd, J, V):
    """emo_rnnt_joint_full_fwd / _bwd (w_enc / w_dec projections, their weight and bias gradients and all casts inside
    the library) against the fp64 oracle of the whole joint: loss and all eight gradients."""
    import emoasr_b200 as E
    from oracle import rnnt_dp
    rng = np.random.default_rng(B * 31 + V)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    eouts, douts = f(B, T, He), np.tanh(f(B, U + 1, Hd))
    w_enc, b_enc = f(J, He) / np.sqrt(He), f(J) * 0.1
    w_dec, b_dec = f(J, Hd) / np.sqrt(Hd), f(J) * 0.1
    w_out, b_out = f(V, J) * (2.0 / np.sqrt(J)), f(V) * 0.5
    ys = rng.integers(1, V, (B, U))
    tl = rng.integers(1, T + 1, B); tl[0] = T
    ul = rng.integers(0, U + 1, B); ul[0] = U
    r = rnnt_dp.joint_loss_and_grads(eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out, ys, tl, ul)
    te = [T_(a).requires_grad_() for a in (eouts, douts, w_enc, b_enc, w_dec, b_dec, w_out, b_out)]
    loss = E.rnnt_joint_loss_from_outputs(*te, T_(ys), T_(tl), T_(ul), blank=0, reduction="mean")
    loss.backward()
    assert abs(float(loss) - r["loss"]) <= BF16_LOSS_RTOL * abs(r["loss"])
    for t, k in zip(te, ["d_eouts", "d_douts", "d_w_enc", "d_b_enc", "d_w_dec", "d_b_dec", "d_w_out", "d_b_out"]):
        assert rel_err(t.grad.cpu().numpy(), r[k]) < BF16_GRAD_RTOL, k


def test_folded_path_full_size_cfg3_vs_fp32_mode():
    """The configuration bench.py times by default -- BASELINE cfg 3 at full size (B=32, T=250, U=100, V=1024, He=256,
    Hd=512, J=512) through emo_rnnt_joint_full_fwd / _bwd -- against the fp32 mode (fp32 cuBLAS projections around the
    fp32 joint kernels, itself pinned to the reference at 1e-5 / 1e-4): loss and all eight gradients within the stated
    bf16 tolerance, exact zeros for padded frames / labels."""
    import emoasr_b200 as E
    gen = torch.Generator().manual_seed(3)
    B, T, U, He, Hd, J, V = 32, 250, 100, 256, 512, 512, 1024
    eouts = torch.randn(B, T, He, generator=gen).to(dev())
    douts = torch.tanh(torch.randn(B, U + 1, Hd, generator=gen)).to(dev())
    torch.manual_seed(1234)
    lin = [torch.nn.Linear(He, J), torch.nn.Linear(Hd, J), torch.nn.Linear(J, V)]
    ps = [p.detach().to(dev()) for l in lin for p in (l.weight, l.bias)]      # w_enc b_enc w_dec b_dec w_out b_out
    ys = torch.randint(1, V, (B, U), generator=gen).to(dev())
    r = torch.linspace(1.0, 0.6, B)
    tl, ul = (T * r).long().to(dev()), (U * r).long().to(dev())
    a = [t.clone().requires_grad_() for t in [eouts, douts] + ps]
    loss = E.rnnt_joint_loss_from_outputs(*a, ys, tl, ul, blank=0, reduction="mean")
    loss.backward()
    b = [t.clone().requires_grad_() for t in [eouts, douts] + ps]
    enc_proj = torch.nn.functional.linear(b[0], b[2], b[3])
    dec_proj = torch.nn.functional.linear(b[1], b[4], b[5])
    loss32 = E.rnnt_joint_loss(enc_proj, dec_proj, b[6], b[7], ys, tl, ul, blank=0, reduction="mean", precision="fp32")
    loss32.backward()
    assert abs(float(loss) - float(loss32)) <= BF16_LOSS_RTOL * abs(float(loss32))
    names = ["d_eouts", "d_douts", "d_w_enc", "d_b_enc", "d_w_dec", "d_b_dec", "d_w_out", "d_b_out"]
    for got, ref, k in zip(a, b, names):
        assert float((got.grad - ref.grad).norm() / ref.grad.norm()) < BF16_GRAD_RTOL, k
    last = B - 1
    assert float(a[0].grad[last, int(tl[last]):].abs().sum()) == 0.0
    assert float(a[1].grad[last, int(ul[last]) + 1:].abs().sum()) == 0.0


# ---------------------------------------------------------------- CTC forced aligner (ctc_aligner.py:138-221)
def test_ctc_forced_aligner_vs_reference_golden():
    """emo_ctc_align against alignments of the UNMODIFIED reference aligner (integer output: exact)."""
    import emoasr_b200 as E
    g = load_golden("ref_ctc_forced_align")
    aligner = E.CTCForcedAligner(blank_id=0)
    for i in range(int(g["n_cases"])):
        lp = T_(g[f"c{i}_log_probs"])
        keep = lp.clone()
        got = aligner(lp, T_(g[f"c{i}_elens"]), T_(g[f"c{i}_ys"]), T_(g[f"c{i}_ylens"]))
        assert got.dtype == torch.int64 and got.shape == lp.shape[:2]
        assert torch.equal(lp, keep)                       # the argument is left alone
        assert np.array_equal(got.cpu().numpy(), g[f"c{i}_aligns"]), f"case {i}"


def test_ctc_forced_aligner_cfg2_size_vs_oracle():
    """B=8 utterances of the cfg-2 shape (T=374, V=5000, U<=80) against the numpy restatement; picks may differ from
    it only where the two best reachable states tie to float32 rounding, which a random input does not produce."""
    import emoasr_b200 as E
    from oracle import ctc_align
    gen = torch.Generator().manual_seed(5)
    B, T, V, U = 8, 374, 5000, 80
    lp = torch.log_softmax(torch.randn(B, T, V, generator=gen) * 2.0, dim=-1)
    ys = torch.randint(1, V, (B, U), generator=gen)
    ys[0, 10:20] = 7                                       # a run of repeated labels
    elens = torch.tensor([374, 374, 300, 251, 200, 170, 161, 90])
    ylens = torch.tensor([80, 41, 80, 60, 80, 3, 80, 0])
    want = ctc_align.ctc_forced_align(lp.numpy(), elens.numpy(), ys.numpy(), ylens.numpy(), blank=0)
    got = E.ctc_forced_align(lp.to(dev()), ys.to(dev()), elens.to(dev()), ylens.to(dev()), blank=0).cpu().numpy()
    assert np.array_equal(got, want)
    for b in range(B):                                     # collapsing the walk gives back the labels
        x, u = int(elens[b]), int(ylens[b])
        seq = got[b, :x]
        labels = [int(v) for k, v in enumerate(seq) if v != 0 and (k == 0 or v != seq[k - 1] or False)]
        if b != 0:                                         # utterance 0 has repeated labels separated by blanks
            assert labels == [int(v) for v in ys[b, :u]]
        assert (got[b, x:] == 0).all()


def test_ctc_forced_aligner_edges():
    """empty label matrix, an utterance without frames, labels outside the vocabulary are clamped, too long a label
    matrix is refused (2*Umax+1 > 1024 threads), CPU tensors are refused."""
    import emoasr_b200 as E
    from oracle import ctc_align
    gen = torch.Generator().manual_seed(9)
    lp = torch.log_softmax(torch.randn(3, 12, 6, generator=gen), dim=-1)
    elens = torch.tensor([12, 0, 5])
    got = E.ctc_forced_align(lp.to(dev()), torch.zeros(3, 0, dtype=torch.long, device=dev()), elens.to(dev()),
                             torch.zeros(3, dtype=torch.long, device=dev()), blank=0).cpu().numpy()
    assert (got == 0).all()                                      # only blanks can be aligned
    ys = torch.tensor([[1, 2, 3], [4, 5, 1], [2, 2, 2]])
    ylens = torch.tensor([3, 2, 3])
    want = ctc_align.ctc_forced_align(lp.numpy(), elens.numpy(), ys.numpy(), ylens.numpy(), blank=0)
    got = E.ctc_forced_align(lp.to(dev()), ys.to(dev()), elens.to(dev()), ylens.to(dev()), blank=0).cpu().numpy()
    assert np.array_equal(got, want) and (got[1] == 0).all()
    with pytest.raises(RuntimeError, match="1024"):
        E.ctc_forced_align(lp.to(dev()), torch.ones(3, 600, dtype=torch.long, device=dev()), elens.to(dev()),
                           ylens.to(dev()), blank=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.ctc_forced_align(lp, ys, elens, ylens)
