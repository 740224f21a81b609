"""Data path: the device batcher against the oracle restatement of the reference
dataset, the oracle against the REAL reference dataset (when /root/reference is present) and against the committed
golden batch it produced.  The batcher is device-agnostic tensor plumbing, so its logic runs on CPU tensors here;
tests/test_gpu_data.py repeats it on cuda:0 and feeds a train step."""
import os
import random

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "ref_dataset_golden.npz")
SPKRS = ["SF1", "SM1", "TF1", "TM1", "TM2"]
TENSOR_KEYS = ["in_feats", "out_feats", "lcf0", "cv_lcf0", "uv", "org_h", "cv_h", "org_h_onehot", "cv_h_onehot",
               "encoder_mask", "decoder_mask", "cycle_encoder_mask", "cycle_decoder_mask", "flen"]
BATCH_LEN = 120
IDX = [3, 0, 7, 5, 11, 2]


def oracle_batch(seed=7):
    from oracle import dataset_port as dp

    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    random.seed(seed)
    return dp.collate([dp.get_item(utts[i], SPKRS, scaler, BATCH_LEN) for i in IDX]), utts, scaler


def product_batch(device, seed=7):
    from crank_b200.data import DeviceBatcher, UtteranceStore
    from oracle import dataset_port as dp

    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    store = UtteranceStore(utts, SPKRS, scaler, device=device)
    random.seed(seed)
    # the reference draws (target speaker, crop start) per sample in this order; DeviceBatcher.draw does the same
    return DeviceBatcher(store, BATCH_LEN).make_batch(IDX)


def compare(prod, orc):
    for k in TENSOR_KEYS:
        a, b = prod[k].cpu().numpy(), np.asarray(orc[k])
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if a.dtype == np.float32:
            assert b.dtype in (np.float32, np.float64), (k, b.dtype)
            assert np.array_equal(a, b.astype(np.float32)), f"{k}: max diff {np.abs(a - b).max()}"
        else:
            assert a.dtype == b.dtype, (k, a.dtype, b.dtype)
            assert np.array_equal(a, b), k
    for k in ("flbl", "org_spkr_name", "cv_spkr_name"):
        assert list(prod[k]) == list(orc[k]), k


def test_device_batcher_matches_oracle_bit_exactly_on_cpu_tensors():
    orc, _, _ = oracle_batch()
    prod = product_batch("cpu")
    compare(prod, orc)
    # crop and pad were both exercised, padded frames carry the ignore index / zeros
    flen = orc["flen"]
    assert (flen > BATCH_LEN).any() and (flen < BATCH_LEN).any()
    assert (prod["org_h"] == -100).any() and not prod["encoder_mask"].all()


def test_oracle_matches_committed_reference_golden():
    orc, _, _ = oracle_batch()
    g = np.load(GOLD, allow_pickle=False)
    for k in TENSOR_KEYS:
        b = g[k]
        a = np.asarray(orc[k])
        assert a.shape == b.shape, k
        assert np.array_equal(a.astype(b.dtype), b), k
    assert [str(s) for s in g["cv_spkr_name"]] == list(orc["cv_spkr_name"])


def test_oracle_matches_real_reference_dataset():
    from oracle import refshim

    if not refshim.available():
        pytest.skip("reference checkout not present on this box")
    from tests.golden.make_dataset_golden import reference_batch

    ref = reference_batch()
    orc, _, _ = oracle_batch()
    for k in TENSOR_KEYS:
        a, b = np.asarray(orc[k]), ref[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a.astype(b.dtype), b), k
    assert list(orc["cv_spkr_name"]) == list(ref["cv_spkr_name"])


def test_epoch_iterator_covers_the_corpus_once():
    from crank_b200.data import DeviceBatcher, UtteranceStore
    from oracle import dataset_port as dp

    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    store = UtteranceStore(utts, SPKRS, scaler, device="cpu")
    random.seed(0)
    seen = []
    for batch in DeviceBatcher(store, BATCH_LEN).epoch(4):
        assert batch["in_feats"].shape == (4, BATCH_LEN, 80)
        seen += batch["flbl"]
    assert sorted(seen) == sorted(u["flbl"] for u in utts)


def test_device_dataloader_token_budget_for_eval():
    """utils.py:85-88: eval / reconstruction batches hold whole utterances under the training token budget."""
    from crank_b200.data import get_device_dataloader
    from oracle import dataset_port as dp

    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    conf = {"batch_len": 100, "batch_size": 8, "input_feat_type": "mlfb", "ignore_scaler": []}
    corpora = {"train": utts[:8], "dev": utts[8:10], "eval": utts[4:]}
    dl = get_device_dataloader(conf, corpora, SPKRS, scaler, flag="train", device="cpu")
    assert set(dl) == {"spkrs", "train", "dev", "eval"} and dl["spkrs"]["TM2"] == 4
    b = next(iter(dl["train"]))
    assert b["in_feats"].shape == (8, 100, 80)
    de = get_device_dataloader(conf, corpora, SPKRS, scaler, flag="eval", device="cpu")
    longest = max(u["mlfb"].shape[0] for u in corpora["eval"])
    batches = list(de["eval"])
    assert all(x["in_feats"].shape[1] == longest for x in batches)
    assert batches[0]["in_feats"].shape[0] == max(800 // longest, 1)
    assert sum(x["in_feats"].shape[0] for x in batches) == len(corpora["eval"])
    # whole utterances: nothing cropped, every valid frame kept
    for x in batches:
        assert torch.equal(x["encoder_mask"].sum(dim=(1, 2)), torch.minimum(x["flen"], torch.tensor(longest)))
