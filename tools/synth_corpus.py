"""Writes a tiny synthetic corpus + config in the formats the UNMODIFIED reference trains from
(asr/datasets.py:24-186: TSV with feat_path/utt_id/token_id/text/xlen/ylen, one .npy of (frames, 80) log-mel per
utterance, vocab.txt "token id"; utils/configure.py: one YAML), for a Conformer-encoder RNN-Transducer
(asr/modeling/decoders/rnn_transducer.py:25-62 reads the dec_*/joint_*/mtl_ctc_weight keys).

    python tools/synth_corpus.py <out_dir> [--utts 24] [--vocab 64] [--joint 128]

Used by tests/test_gpu_reference_train.py; sizes are small, the joint shape (J % 128 == 0, V % 32 == 0) is one the
tensor-core kernels accept.
"""
import argparse
import os

import numpy as np

CONF = """encoder_type: "conformer"
decoder_type: "rnn_transducer"
lr_schedule_type: "noam"
input_layer: "conv2d"
feat_dim: 80
num_framestacks: 1
spec_augment: false
enc_hidden_size: {He}
enc_num_attention_heads: 4
enc_num_layers: 2
enc_intermediate_size: 128
pos_encode_type: "rel"
dec_num_layers: 1
dec_hidden_size: 64
embedding_size: 32
joint_hidden_size: {J}
dropout_emb_rate: 0.0
dropout_dec_rate: 0.0
blank_id: 0
eos_id: 2
vocab_path: "{root}/vocab.txt"
vocab_size: {V}
train_path: "{root}/train.tsv"
dev_path: "{root}/dev.tsv"
test_path: "{root}/dev.tsv"
train_data_shuffle: false
model_path: ""
optim_path: ""
startep: 0
log_step: 1
save_step: 1
batch_size: {B}
max_xlens_batch: 30000
max_ylens_batch: 3000
num_epochs: 1
learning_rate: 1.0
num_warmup_steps: 100
clip_grad_norm: 5.0
dropout_enc_rate: 0.0
dropout_attn_rate: 0.0
weight_decay: 0.000001
accum_grad: 1
lsm_prob: 0
kd_weight: 0
mtl_ctc_weight: {ctc_w}
beam_width: 0
len_weight: 0
decode_ctc_weight: 0
lm_weight: 0
"""


def write(root, utts=24, V=64, J=128, He=64, B=4, ctc_w=0.3, seed=0):
    os.makedirs(os.path.join(root, "feats"), exist_ok=True)
    rng = np.random.default_rng(seed)
    with open(os.path.join(root, "vocab.txt"), "w") as f:
        for i, tok in enumerate(["<pad>", "<unk>", "<eos>", "<pad2>"] + [f"t{i}" for i in range(4, V)]):
            f.write(f"{tok} {i}\n")
    rows = []
    for i in range(utts):
        xlen = int(rng.integers(60, 121))                 # frames; T = ((xlen-1)//2-1)//2 after conv2d x4
        ylen = int(rng.integers(3, 9))
        x = rng.standard_normal((xlen, 80)).astype(np.float32)
        path = os.path.join(root, "feats", f"utt{i:03d}.npy")
        np.save(path, x)
        y = rng.choice(np.concatenate([[1], np.arange(4, V)]), ylen)
        rows.append((path, f"utt{i:03d}", " ".join(map(str, y)), " ".join(f"t{t}" for t in y), xlen, ylen))
    rows.sort(key=lambda r: r[4])                          # the reference expects length-sorted data
    for name, part in (("train.tsv", rows[: utts - B]), ("dev.tsv", rows[utts - B:])):
        with open(os.path.join(root, name), "w") as f:
            f.write("feat_path\tutt_id\ttoken_id\ttext\txlen\tylen\n")
            for r in part:
                f.write("\t".join(map(str, r)) + "\n")
    conf = os.path.join(root, "rnnt_conformer.yaml")
    with open(conf, "w") as f:
        f.write(CONF.format(root=root, V=V, J=J, He=He, B=B, ctc_w=ctc_w))
    return conf


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("out")
    ap.add_argument("--utts", type=int, default=24)
    ap.add_argument("--vocab", type=int, default=64)
    ap.add_argument("--joint", type=int, default=128)
    a = ap.parse_args()
    print(write(os.path.abspath(a.out), a.utts, a.vocab, a.joint))
