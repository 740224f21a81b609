"""ORACLE (test infrastructure, NOT product code): numpy restatement of the reference's per-sample data path,
`BaseDataset.__getitem__` (crank/net/trainer/dataset.py:58-203) + the default collate, for in-memory utterances.

Pinned: tests/test_cpu_data.py runs the REAL `crank.net.trainer.dataset.BaseDataset` (through oracle/refshim.py,
with `read_feature` served from memory instead of HDF5) on the same seeded corpus and requires identical batches;
the same comparison is committed as tests/golden/ref_dataset_golden.npz for boxes without the reference.
"""
import random

import numpy as np


def _padding(x, dlen, batch_len, value=0.0, p=0):
    """dataset.py:239-259"""
    if dlen >= 0:
        actual = batch_len - x.shape[0]
        if actual != 0:
            if x.ndim == 2:
                x = np.concatenate([x, np.ones((actual, x.shape[1])) * value])
            else:
                x = np.concatenate([x, np.ones((actual)) * value])
        else:
            return x
    else:
        x = x[p : p + batch_len]
    if type(value) == bool:
        return x.astype(bool)
    if isinstance(value, int):
        return x.astype(np.int64)
    return x.astype(np.float32)


def get_item(utt, spkrs, scaler, batch_len, feat_type="mlfb"):
    """One sample; consumes Python `random` exactly like the reference (target speaker, then crop start)."""
    spkrdict = dict(zip(spkrs, range(len(spkrs))))
    S = len(spkrs)

    def col(a):
        a = np.asarray(a, dtype=np.float64)
        return a[:, None] if a.ndim == 1 else a

    s = {feat_type: col(utt[feat_type]), "lcf0": col(utt["lcf0"]), "uv": col(utt["uv"])}
    org = utt["spkr"]
    cv = random.choice([k for k in spkrdict.keys() if k != org])                      # :85-87
    flen = s[feat_type].shape[0]
    s["mask"] = np.ones(flen, dtype=bool)[:, None]
    for name, key in ((org, "org"), (cv, "cv")):                                       # :152-156
        num = int(spkrdict[name])
        s[f"{key}_h"] = (np.ones(flen) * num).astype(np.int64)
        oh = np.zeros((flen, S), dtype=np.float32)
        oh[:, num] = 1
        s[f"{key}_h_onehot"] = oh
    s["cv_lcf0"] = (s["lcf0"] - scaler[org]["lcf0"].mean_) / np.sqrt(scaler[org]["lcf0"].var_) * np.sqrt(
        scaler[cv]["lcf0"].var_) + scaler[cv]["lcf0"].mean_                          # :290-293, on the RAW lcf0
    for k in (feat_type, "lcf0"):                                                       # :146-150 (uv is never scaled)
        s[k] = scaler[k].transform(s[k])
    diff = batch_len - flen                                                             # :158-190
    p = random.choice(range(0, abs(diff))) if diff < 0 else 0
    for k, v in list(s.items()):
        if k == "mask":
            s[k] = _padding(v, diff, batch_len, value=False, p=p)
        elif k in ("org_h", "cv_h"):
            s[k] = _padding(v, diff, batch_len, value=-100, p=p)
        else:
            s[k] = _padding(v, diff, batch_len, value=0.0, p=p)
    for ed in ("encoder_mask", "decoder_mask", "cycle_encoder_mask", "cycle_decoder_mask"):
        s[ed] = np.copy(s["mask"])
    del s["mask"]
    s["in_feats"] = s[feat_type].copy()
    s["out_feats"] = s[feat_type].copy()
    del s[feat_type]
    s["flen"] = flen
    s["org_spkr_name"], s["cv_spkr_name"] = org, cv
    s["flbl"] = utt.get("flbl", "")
    return s


def collate(samples):
    out = {}
    for k in samples[0]:
        v0 = samples[0][k]
        if isinstance(v0, np.ndarray):
            out[k] = np.stack([s[k] for s in samples])
        elif isinstance(v0, (int, np.integer)):
            out[k] = np.asarray([s[k] for s in samples], dtype=np.int64)
        else:
            out[k] = [s[k] for s in samples]
    return out


def make_corpus(n_utts, spkrs, seed=0, dim=80, min_len=60, max_len=260):
    """Seeded in-memory corpus + fitted scalers (global for the features / lcf0, per speaker for lcf0)."""
    from sklearn.preprocessing import StandardScaler

    rng = np.random.RandomState(seed)
    utts = []
    for i in range(n_utts):
        n = int(rng.randint(min_len, max_len))
        spk = spkrs[i % len(spkrs)]
        base = 4.5 + 0.3 * (i % len(spkrs))
        utts.append({"mlfb": rng.randn(n, dim) * 2.0 + 0.5, "lcf0": base + 0.2 * rng.randn(n),
                     "uv": (rng.rand(n) < 0.7).astype(np.float64), "spkr": spk, "flbl": f"{spk}/utt{i}"})
    scaler = {"mlfb": StandardScaler().fit(np.concatenate([u["mlfb"] for u in utts])),
              "lcf0": StandardScaler().fit(np.concatenate([u["lcf0"][:, None] for u in utts]))}
    for s in spkrs:
        scaler[s] = {"lcf0": StandardScaler().fit(np.concatenate([u["lcf0"][:, None] for u in utts if u["spkr"] == s]))}
    return utts, scaler
