"""Device-resident replacement of the reference's training data path.

What it replaces: `BaseDataset.__getitem__` + the default DataLoader collate
(crank/net/trainer/dataset.py:58-203, 239-293) -- per utterance: HDF5 read, StandardScaler transform, speaker
codes / one-hots, log-F0 conversion to the target speaker, random crop or zero padding to `batch_len`, four mask
copies -- executed in Python worker processes per sample, then stacked and copied to the GPU every step.  At
>1 M frames/s per GPU that path cannot feed the trainers.

Here the corpus is uploaded ONCE: every utterance's frame features live in HBM as one packed (sum N_i, D) panel
(VCC2020: ~1k utterances x ~700 frames x 82 floats = 0.2 GB of 180 GB).  A step's batch is assembled on the device
from per-utterance (offset, length, crop start, speaker ids): gather / crop / pad / mask / one-hot / F0 conversion
are a handful of batched tensor ops over (B, T) index grids -- no per-sample Python, no host->device copy of
features.  Host-side randomness (target speaker, crop start) is drawn with Python's `random` in the reference's
order (`random.choice` of the target speaker per sample, dataset.py:85-87, then `random.choice` of the crop start
for utterances longer than batch_len, dataset.py:161-162), so a seeded run selects the same crops.

The batch dict follows the reference schema.  Reference behaviours kept on purpose:
  * `cv_lcf0` is converted from the RAW log-F0 with the per-speaker statistics and is NOT standardised afterwards
    (it is not in `self.features`, dataset.py:36-44,146-150), while `lcf0` is;
  * `uv` (and `cap`) are never scaled; padded frames are 0 / False / -100 (dataset.py:168-190);
  * `in_feats` and `out_feats` are separate tensors (the trainers may modify one).
This module is host plumbing on torch tensors (device-agnostic, so its logic is unit-tested on CPU tensors too);
the train step it feeds has no CPU path.
"""

import random

import numpy as np
import torch

IGNORE_INDEX = -100


class UtteranceStore:
    """Packed corpus: features (sum N, D) standardised once, raw log-F0 / uv, per-utterance offset / length / speaker."""

    def __init__(self, utterances, spkrs, scaler, feat_type="mlfb", device="cuda", ignore_scaler=()):
        """utterances: list of dicts {"<feat_type>": (N, D), "lcf0": (N,) or (N,1), "uv": (N,) or (N,1),
        "spkr": name, "flbl": label}; spkrs: ordered speaker names (scp["train"]["spkrs"], dataset.py:32);
        scaler: {"<feat_type>": StandardScaler, "lcf0": StandardScaler, spkr: {"lcf0": StandardScaler}} as
        loaded by the reference's train.py (joblib scaler file)."""
        self.spkrs = list(spkrs)
        self.spkrdict = {s: i for i, s in enumerate(self.spkrs)}
        self.n_spkrs = len(self.spkrs)
        self.device = torch.device(device)
        self.feat_type = feat_type
        lens = [int(np.asarray(u[feat_type]).shape[0]) for u in utterances]
        self.lengths = lens
        self.offsets = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64) if lens else np.zeros(0, np.int64)
        self.flbl = [u.get("flbl", f"{u['spkr']}/utt{i}") for i, u in enumerate(utterances)]
        self.spkr_names = [u["spkr"] for u in utterances]
        self.spkr_ids = [self.spkrdict[s] for s in self.spkr_names]

        def col(a):
            a = np.asarray(a, dtype=np.float64)
            return a[:, None] if a.ndim == 1 else a

        feats = np.concatenate([col(u[feat_type]) for u in utterances], axis=0)
        lcf0 = np.concatenate([col(u["lcf0"]) for u in utterances], axis=0)
        uv = np.concatenate([col(u["uv"]) for u in utterances], axis=0)
        if scaler is not None and feat_type not in ignore_scaler:
            feats = scaler[feat_type].transform(feats)            # dataset.py:146-150
        self.feats = torch.from_numpy(feats.astype(np.float32)).to(self.device)
        self.lcf0_raw = torch.from_numpy(lcf0).to(self.device)     # float64: the conversion below is done in double
        self.uv = torch.from_numpy(uv.astype(np.float32)).to(self.device)
        if scaler is not None:
            g = scaler["lcf0"]
            self.lcf0_global = (float(g.mean_[0]), float(g.scale_[0])) if "lcf0" not in ignore_scaler else (0.0, 1.0)
            mean = [float(scaler[s]["lcf0"].mean_[0]) for s in self.spkrs]
            std = [float(np.sqrt(scaler[s]["lcf0"].var_[0])) for s in self.spkrs]
        else:
            self.lcf0_global = (0.0, 1.0)
            mean, std = [0.0] * self.n_spkrs, [1.0] * self.n_spkrs
        self.spk_mean = torch.tensor(mean, dtype=torch.float64, device=self.device)
        self.spk_std = torch.tensor(std, dtype=torch.float64, device=self.device)

    def __len__(self):
        return len(self.lengths)

    @property
    def n_frames(self):
        return int(self.feats.shape[0])


class DeviceBatcher:
    """Assembles reference-schema batches on the device from an UtteranceStore."""

    def __init__(self, store, batch_len):
        self.store = store
        self.batch_len = int(batch_len)

    def draw(self, indices):
        """Host randomness in the reference's per-sample order: (target speaker ids, crop starts)."""
        st = self.store
        cv, crop = [], []
        for i in indices:
            org = st.spkr_names[i]
            cv_name = random.choice([s for s in st.spkrdict.keys() if s != org])   # dataset.py:85-87
            cv.append(st.spkrdict[cv_name])
            diff = self.batch_len - st.lengths[i]
            crop.append(random.choice(range(0, abs(diff))) if diff < 0 else 0)     # dataset.py:161-162
        return cv, crop

    def make_batch(self, indices, cv_spkrs=None, crops=None):
        st, T, dev = self.store, self.batch_len, self.store.device
        indices = [int(i) for i in indices]
        if cv_spkrs is None or crops is None:
            d_cv, d_crop = self.draw(indices)
            cv_spkrs = d_cv if cv_spkrs is None else cv_spkrs
            crops = d_crop if crops is None else crops
        B = len(indices)
        lens = torch.tensor([st.lengths[i] for i in indices], device=dev)
        offs = torch.tensor([int(st.offsets[i]) for i in indices], device=dev)
        p = torch.tensor([int(c) for c in crops], device=dev)
        org = torch.tensor([st.spkr_ids[i] for i in indices], device=dev)
        cv = torch.tensor([int(c) for c in cv_spkrs], device=dev)
        t = torch.arange(T, device=dev)[None, :]                              # (1, T)
        n_valid = torch.minimum(lens - p, torch.full_like(lens, T))           # frames kept of each utterance
        valid = t < n_valid[:, None]                                           # (B, T)
        src = (offs + p)[:, None] + torch.minimum(t, (n_valid - 1).clamp_min(0)[:, None])   # clamped: padded frames re-read a valid row
        vf = valid[..., None]
        feats = torch.where(vf, st.feats[src], torch.zeros((), device=dev))
        lraw = st.lcf0_raw[src]                                                # (B, T, 1) float64
        gm, gs = st.lcf0_global
        lcf0 = torch.where(vf, ((lraw - gm) / gs).float(), torch.zeros((), device=dev))
        # convert_f0 (dataset.py:290-293): (lcf0 - mean_org) / std_org * std_cv + mean_cv on the raw log-F0
        m_o, s_o = st.spk_mean[org][:, None, None], st.spk_std[org][:, None, None]
        m_c, s_c = st.spk_mean[cv][:, None, None], st.spk_std[cv][:, None, None]
        cv_lcf0 = torch.where(vf, ((lraw - m_o) / s_o * s_c + m_c).float(), torch.zeros((), device=dev))
        uv = torch.where(vf, st.uv[src], torch.zeros((), device=dev))
        ign = torch.full((), IGNORE_INDEX, device=dev, dtype=torch.int64)
        org_h = torch.where(valid, org[:, None].expand(B, T), ign)
        cv_h = torch.where(valid, cv[:, None].expand(B, T), ign)
        onehot = torch.nn.functional.one_hot(org, st.n_spkrs).float()[:, None, :] * vf
        cv_onehot = torch.nn.functional.one_hot(cv, st.n_spkrs).float()[:, None, :] * vf
        mask = vf.clone()
        return {
            "in_feats": feats,
            "out_feats": feats.clone(),
            "lcf0": lcf0,
            "cv_lcf0": cv_lcf0,
            "uv": uv,
            "org_h": org_h,
            "cv_h": cv_h,
            "org_h_onehot": onehot.contiguous(),
            "cv_h_onehot": cv_onehot.contiguous(),
            "encoder_mask": mask,
            "decoder_mask": mask.clone(),
            "cycle_encoder_mask": mask.clone(),
            "cycle_decoder_mask": mask.clone(),
            "flen": torch.tensor([st.lengths[i] for i in indices], dtype=torch.int64),
            "flbl": [st.flbl[i] for i in indices],
            "org_spkr_name": [st.spkr_names[i] for i in indices],
            "cv_spkr_name": [st.spkrs[int(c)] for c in cv_spkrs],
        }

    def epoch(self, batch_size, shuffle=True, drop_last=True):
        """Yield batches over the corpus (the DataLoader(shuffle=True) role, trainer/utils.py:77-100)."""
        order = list(range(len(self.store)))
        if shuffle:
            random.shuffle(order)
        for b0 in range(0, len(order), batch_size):
            idx = order[b0 : b0 + batch_size]
            if len(idx) < batch_size and drop_last:
                break
            yield self.make_batch(idx)


class DeviceLoader:
    """Iterable with the DataLoader role (len / iteration over one epoch; no drop_last, like the reference's loaders)."""

    def __init__(self, batcher, batch_size, shuffle):
        self.batcher, self.batch_size, self.shuffle = batcher, max(int(batch_size), 1), shuffle

    def __len__(self):
        n = len(self.batcher.store)
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        return self.batcher.epoch(self.batch_size, shuffle=self.shuffle, drop_last=False)


def get_device_dataloader(conf, corpora, spkrs, scaler, flag="train", device="cuda"):
    """Device-resident counterpart of `get_dataloader` (crank/net/trainer/utils.py:77-109): the same
    {"spkrs", "train", "dev", "eval"} dict the trainers take.  corpora: {"train"/"dev"/"eval": [utterance dicts]}.
    For `reconstruction` / `eval` the reference converts its (batch_len x batch_size) token budget into
    whole-utterance batches: batch_len = longest utterance, batch_size = tokens // batch_len (utils.py:85-88)."""
    batch_len, batch_size = int(conf["batch_len"]), int(conf["batch_size"])
    if flag in ("reconstruction", "eval"):
        pool = corpora.get("eval", []) if flag == "eval" else corpora.get("train", []) + corpora.get("dev", [])
        ft = conf["input_feat_type"]
        token_size = batch_len * batch_size
        batch_len = max(int(np.asarray(u[ft]).shape[0]) for u in pool)
        batch_size = max(token_size // batch_len, 1)
    out = {"spkrs": {s: i for i, s in enumerate(spkrs)}}
    for phase, shuffle in (("train", True), ("dev", True), ("eval", False)):
        if corpora.get(phase):
            store = UtteranceStore(corpora[phase], spkrs, scaler, feat_type=conf["input_feat_type"], device=device,
                                   ignore_scaler=conf.get("ignore_scaler", ()))
            out[phase] = DeviceLoader(DeviceBatcher(store, batch_len), batch_size, shuffle)
    return out
