"""Generates tests/golden/ref_dataset_golden.npz by running the REAL reference dataset
(crank/net/trainer/dataset.py, unmodified, through oracle/refshim.py) on the seeded in-memory corpus of
oracle/dataset_port.make_corpus; `read_feature` (the HDF5 reader, dataset.py:223-229) is served from memory.
Run in the build container only:  python -m tests.golden.make_dataset_golden"""
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SPKRS = ["SF1", "SM1", "TF1", "TM1", "TM2"]
BATCH_LEN = 120
IDX = [3, 0, 7, 5, 11, 2]


def reference_batch(seed=7):
    from oracle import dataset_port as dp
    from oracle import refshim

    ds = refshim.ref("crank.net.trainer.dataset")
    utts, scaler = dp.make_corpus(12, SPKRS, seed=1)
    paths = {f"/mem/{u['spkr']}/{u['flbl'].split('/')[1]}.h5": u for u in utts}

    def read_feature(h5f, ext="mlfb"):
        a = np.asarray(paths[str(h5f)][ext], dtype=np.float64)
        return a[:, np.newaxis] if a.ndim == 1 else a

    ds.read_feature = read_feature

    # dataset.py:111 compares the feature ARRAY with the string "excit"; numpy < 1.25 (the reference's era)
    # evaluated that to a scalar False, numpy 2 raises "truth value is ambiguous".  Emulate the old semantics
    # on the one array that reaches the comparison; the reference source stays untouched.
    class _OldEq(np.ndarray):
        def __eq__(self, other):
            return False if isinstance(other, str) else np.ndarray.__eq__(self, other)

        __hash__ = None

    orig_transform = ds.BaseDataset._transform

    def _transform(self, sample):
        sample = orig_transform(self, sample)
        k = self.conf["output_feat_type"]
        sample[k] = np.asarray(sample[k]).view(_OldEq)
        return sample

    ds.BaseDataset._transform = _transform
    conf = {"batch_len": BATCH_LEN, "input_feat_type": "mlfb", "output_feat_type": "mlfb", "use_raw": False,
            "cache_dataset": False, "ignore_scaler": [], "use_mcep_0th": False, "spec_augment": False,
            "feature": {"fftl": 1024, "hop_size": 128}}
    scp = {"train": {"feats": {k: k for k in paths}, "spkrs": SPKRS}}
    dataset = ds.BaseDataset(conf, scp, scaler, phase="train")
    random.seed(seed)
    samples = [dataset[i] for i in IDX]
    out = {}
    for k in samples[0]:
        v0 = samples[0][k]
        if isinstance(v0, np.ndarray):
            out[k] = np.stack([np.asarray(s[k]) for s in samples])
        elif isinstance(v0, (int, np.integer)):
            out[k] = np.asarray([s[k] for s in samples], dtype=np.int64)
        else:
            out[k] = [s[k] for s in samples]
    return out


if __name__ == "__main__":
    b = reference_batch()
    arrays = {k: v for k, v in b.items() if isinstance(v, np.ndarray)}
    arrays["cv_spkr_name"] = np.asarray(b["cv_spkr_name"])
    arrays["org_spkr_name"] = np.asarray(b["org_spkr_name"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_dataset_golden.npz"), **arrays)
    print({k: (v.shape, v.dtype) for k, v in arrays.items()})
