"""The UNMODIFIED reference on the B200 through the drop-in seams (asr/train_asr.py:35-98,
asr/modeling/asr.py:53-67): a Conformer-encoder RNN-Transducer with an auxiliary CTC head is built by the
reference's own ASR class, trained by the reference's own train_step / main loop on a synthetic corpus
(tools/synth_corpus.py), with emoasr_b200.dropin.install() routing both losses to the CUDA kernels.

The reference's sources are not part of this repository: they are looked up in $EMOASR_REFERENCE,
baseline/_ref/emoASR (staged by tools/stage_reference.py; travels to the GPU box) or /root/reference.
"""
import os
import re
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(ROOT, "tools"))


def _reference_root():
    for cand in (os.environ.get("EMOASR_REFERENCE"), os.path.join(ROOT, "baseline", "_ref", "emoASR"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "asr", "train_asr.py")):
            return cand
    pytest.skip("reference sources not found (run tools/stage_reference.py where /root/reference exists)")


def _git_cwd(path):
    # train_asr.py:209 logs the commit hash of the repository the process runs in
    subprocess.check_call(["git", "init", "-q", path])
    subprocess.check_call(["git", "-C", path, "-c", "user.email=t@t", "-c", "user.name=t", "commit", "-q",
                           "--allow-empty", "-m", "init"])


LOSS_RE = re.compile(r"step =\s+(\d+) /\s+\d+ .*loss_rnnt: ([\d.]+) loss_ctc: ([\d.]+) loss_total: ([\d.]+)")


def _run_script(ref, conf, cwd, launcher_args, env_extra):
    env = dict(os.environ, CONDA_DEFAULT_ENV="none", PYTHONPATH=ROOT, **env_extra)
    cmd = [sys.executable] + launcher_args + [os.path.join(ref, "asr", "train_asr.py") if not launcher_args else "asr/train_asr.py",
                                              "-conf", conf, "--debug", "--num_workers", "0"]
    p = subprocess.run(cmd, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)
    log = p.stdout + p.stderr
    assert p.returncode == 0, log[-3000:]
    assert "ERROR occurs in training" not in log, log[-3000:]
    return [tuple(float(x) for x in m.groups()[1:]) for m in LOSS_RE.finditer(log)], log


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unmodified_train_script_through_launcher(tmp_path, precision):
    """python -m emoasr_b200.dropin --reference REF -- asr/train_asr.py -conf X.yaml : one epoch (5 steps) +
    validation of the reference's own script on the GPU.  The logged losses of every step are compared with the
    same script run on the CPU with the oracle's warp_rnnt shim: both start from the same seed (train_asr.py:31-32),
    so the whole trajectory must agree (fp32 to the 3 printed decimals of the first step, bf16 to 2e-3)."""
    import synth_corpus
    ref = _reference_root()
    root = str(tmp_path / "corpus")
    conf = synth_corpus.write(root)
    _git_cwd(str(tmp_path))
    gpu, log = _run_script(ref, conf, str(tmp_path),
                           ["-m", "emoasr_b200.dropin", "--reference", ref, "--precision", precision, "--"], {})
    assert len(gpu) == 5, log[-3000:]
    assert "valid WER" in log
    # CPU run of the same script: oracle shim as `warp_rnnt`, no CUDA device
    shim = tmp_path / "cpu_run.py"
    shim.write_text(
        "import runpy, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from oracle import warp_rnnt_shim\n"
        "warp_rnnt_shim.install()\n"
        f"sys.path.insert(0, {ref!r})\n"
        f"sys.argv = [{os.path.join(ref, 'asr', 'train_asr.py')!r}] + sys.argv[1:]\n"
        "runpy.run_path(sys.argv[0], run_name='__main__')\n")
    env = dict(os.environ, CONDA_DEFAULT_ENV="none", PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, str(shim), "-conf", conf, "--debug", "--num_workers", "0"], cwd=str(tmp_path),
                       env=env, capture_output=True, text=True, timeout=600)
    cpu = [tuple(float(x) for x in m.groups()[1:]) for m in LOSS_RE.finditer(p.stdout + p.stderr)]
    assert len(cpu) == 5, (p.stdout + p.stderr)[-3000:]
    tol0 = 2e-3 if precision == "fp32" else 0.2      # absolute, on values ~80 printed with 3 decimals
    for a, b in zip(gpu[0], cpu[0]):
        assert abs(a - b) <= tol0, (gpu[0], cpu[0])
    for g, c in zip(gpu, cpu):                        # later steps: Adam amplifies rounding differences a little
        assert abs(g[2] - c[2]) <= (2e-3 if precision == "fp32" else 1e-2) * c[2], (gpu, cpu)


def test_train_step_fp32_matches_cpu_reference_to_1e5():
    """dropin.install() -> ASR(params) -> the reference's train_step(), three steps; step 1 (identical weights)
    must equal the reference on the CPU (oracle shim) to 1e-5 relative in every loss_dict entry."""
    import synth_corpus
    import tempfile
    ref = _reference_root()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from oracle import warp_rnnt_shim
    warp_rnnt_shim.install()                 # `import warp_rnnt` of the reference's CPU model
    sys.path.insert(0, ref)
    with tempfile.TemporaryDirectory() as tmp:
        conf = synth_corpus.write(os.path.join(tmp, "corpus"))
        from utils.configure import load_config
        from asr.datasets import ASRDataset
        from asr.optimizers import ScheduledOptimizer
        import asr.modeling.asr as ref_asr
        params = load_config(conf)
        torch.manual_seed(0)
        cpu_model = ref_asr.ASR(params)      # the reference's own classes
        from emoasr_b200 import dropin
        dropin.install(ref, precision="fp32")
        gpu_model = ref_asr.ASR(params)      # same class, decoders rebound to the fused subclasses
        gpu_model.load_state_dict(cpu_model.state_dict())
        assert type(gpu_model.decoder).__mro__[1].__name__ == "FusedRNNTForward"
        gpu_model.cuda().train()
        cpu_model.train()
        ds = ASRDataset(params, params.train_path, phase="train")
        batches = [ds.collate_fn([ds[i] for i in range(k, k + 4)]) for k in (0, 4, 8)]
        # train_step is defined in a script that parses argv at import: load it as a module without running main
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_train_asr", os.path.join(ref, "asr", "train_asr.py"))
        ta = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ta)
        outs = {}
        for name, model, device in (("cpu", cpu_model, torch.device("cpu")), ("gpu", gpu_model, torch.device("cuda"))):
            opt = ScheduledOptimizer(torch.optim.Adam(model.parameters(), lr=0, weight_decay=params.weight_decay), params)
            outs[name] = [ta.train_step(model, opt, data, params, device) for data in batches]
    for k in ("loss_rnnt", "loss_ctc", "loss_total"):
        assert abs(outs["gpu"][0][k] - outs["cpu"][0][k]) <= 1e-5 * abs(outs["cpu"][0][k]), (k, outs["gpu"][0], outs["cpu"][0])
    for g, c in zip(outs["gpu"][1:], outs["cpu"][1:]):
        assert abs(g["loss_total"] - c["loss_total"]) <= 1e-3 * abs(c["loss_total"]), (g, c)


def test_ctc_distillation_through_the_seams_matches_the_reference_classes():
    """CTC knowledge distillation (ctc.py:117-127): the reference's forward with soft labels, once with its own
    nn.CTCLoss + Python forced aligner, once with the attribute seams (CUDA CTC loss, one-launch forced aligner,
    ctc_aligner.py:138-221) -- same loss_dict and gradients."""
    from collections import namedtuple
    ref = _reference_root()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sys.path.insert(0, ref)
    from emoasr_b200 import dropin
    from emoasr_b200.criteria import CTCForcedAligner
    ref_asr = dropin.install(ref, precision="fp32")
    import asr.modeling.decoders.ctc as ref_ctc
    base = dict(enc_hidden_size=32, vocab_size=47, blank_id=0, eos_id=2, kd_weight=0.4, lsm_prob=0.1,
                reduce_main_loss_kd=False, mtl_phone_ctc_weight=0, mtl_inter_ctc_weight=0)
    p = namedtuple("Params", base.keys())(**base)
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    plain = ref_ctc.CTCDecoder(p).to(dev)
    fused = ref_asr.CTCDecoder(p).to(dev)
    fused.load_state_dict(plain.state_dict())
    assert isinstance(fused.forced_aligner, CTCForcedAligner)
    B, T, U = 4, 31, 7
    g = torch.Generator().manual_seed(4)
    eouts = torch.randn(B, T, 32, generator=g).to(dev)
    elens = torch.tensor([31, 26, 19, 31], device=dev)
    ys = torch.randint(4, 47, (B, U), generator=g)     # labels stay on the host: the reference's to_onehot (criteria.py:5-6)
    ylens = torch.tensor([7, 5, 7, 2])                # indexes a CPU identity matrix with them
    soft = torch.softmax(torch.randn(B, U, 47, generator=g), dim=-1).to(dev)
    outs = {}
    for name, dec in (("plain", plain), ("fused", fused)):
        x = eouts.clone().requires_grad_()
        loss, loss_dict, _ = dec(x, elens, None, ys, ylens, None, None, soft)
        loss.backward()
        outs[name] = ({k: float(v) for k, v in loss_dict.items()}, x.grad.clone(), dec.output.weight.grad.clone())
    for k, v in outs["plain"][0].items():
        assert abs(outs["fused"][0][k] - v) <= 1e-5 * abs(v), (k, outs)
    for a, b in zip(outs["fused"][1:], outs["plain"][1:]):
        assert float((a - b).norm() / b.norm()) < 1e-4
