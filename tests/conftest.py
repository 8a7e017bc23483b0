import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.fixture
def golden():
    return load_golden


def pred_net_douts(g):
    """Prediction-network output for a golden RNN-T case, rebuilt with torch from the stored
    state_dict (Embedding -> LSTM x L, dropout 0; rnn_transducer.py:158-192)."""
    L = int(g["hp.dec_num_layers"])
    emb = torch.from_numpy(g["param.embed.weight"])
    x = torch.nn.functional.embedding(torch.from_numpy(g["ys_in"]), emb)
    for l in range(L):
        H = int(g["hp.dec_hidden_size"])
        rnn = torch.nn.LSTM(x.size(-1), H, 1, batch_first=True)
        rnn.load_state_dict({k[len(f"param.rnns.{l}."):]: torch.from_numpy(v) for k, v in g.items()
                             if k.startswith(f"param.rnns.{l}.")})
        x, _ = rnn(x)
    return x.detach()


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
