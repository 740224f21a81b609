"""Synthetic mini-batches with the exact dict schema the reference's dataset emits.

Mirrors the output of `BaseDataset.__getitem__` + default collate
(crank/net/trainer/dataset.py:58-203; schema listed in SURVEY.md section 8b): there is no dataset
on the build/GPU boxes, so benches and parity tests feed this instead.

Tensors (B = utterances, T = batch_len, S = #speakers):
  in_feats, out_feats (B,T,80) f32 ~ N(0,1)   (features are StandardScaler-normalised)
  lcf0, cv_lcf0 (B,T,1) f32 ~ N(0,1);  uv (B,T,1) f32 ~ Bernoulli(0.7)
  org_h, cv_h (B,T) int64 (pad -100);  org_h_onehot, cv_h_onehot (B,T,S) f32
  encoder_mask, decoder_mask, cycle_encoder_mask, cycle_decoder_mask (B,T,1) bool
  flen (B,) int64; flbl / org_spkr_name / cv_spkr_name list[str]
`ragged=True` is "variant R": flen ~ U[T/2, T], masks False / labels -100 / features 0 beyond flen.
"""

import torch


def make_batch(B, T, n_spkrs, dim=80, seed=0, ragged=False, device="cpu", spkr_offset=0):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, T, dim, generator=g)
    lcf0 = torch.randn(B, T, 1, generator=g)
    cv_lcf0 = torch.randn(B, T, 1, generator=g)
    uv = (torch.rand(B, T, 1, generator=g) < 0.7).float()
    spk = (torch.randint(0, n_spkrs, (B,), generator=g) + spkr_offset) % n_spkrs
    cv_spk = (spk + 1) % n_spkrs
    if ragged:
        flen = torch.randint(T // 2, T + 1, (B,), generator=g)
    else:
        flen = torch.full((B,), T, dtype=torch.int64)
    t = torch.arange(T)[None, :]
    valid = t < flen[:, None]  # (B,T)
    org_h = torch.where(valid, spk[:, None].expand(B, T), torch.full((B, T), -100))
    cv_h = torch.where(valid, cv_spk[:, None].expand(B, T), torch.full((B, T), -100))
    onehot = torch.nn.functional.one_hot(spk, n_spkrs).float()[:, None, :] * valid[..., None]
    cv_onehot = (
        torch.nn.functional.one_hot(cv_spk, n_spkrs).float()[:, None, :] * valid[..., None]
    )
    vf = valid[..., None].float()
    feats = feats * vf
    mask = valid[..., None].clone()
    batch = {
        "in_feats": feats.clone(),
        "out_feats": feats.clone(),
        "lcf0": lcf0 * vf,
        "cv_lcf0": cv_lcf0 * vf,
        "uv": uv * vf,
        "org_h": org_h.long(),
        "cv_h": cv_h.long(),
        "org_h_onehot": onehot.contiguous(),
        "cv_h_onehot": cv_onehot.contiguous(),
        "encoder_mask": mask.clone(),
        "decoder_mask": mask.clone(),
        "cycle_encoder_mask": mask.clone(),
        "cycle_decoder_mask": mask.clone(),
        "flen": flen,
        "flbl": [f"spk{int(s)}/utt{n}" for n, s in enumerate(spk)],
        "org_spkr_name": [f"spk{int(s)}" for s in spk],
        "cv_spkr_name": [f"spk{int(s)}" for s in cv_spk],
    }
    if device != "cpu":
        batch = to_device(batch, device)
    return batch


def to_device(batch, device, non_blocking=False):
    """crank/utils/utils.py:201-207 (tensor values only)."""
    out = {}
    for k, v in batch.items():
        out[k] = v.to(device, non_blocking=non_blocking) if isinstance(v, torch.Tensor) else v
    return out


def clone_batch(batch):
    return {k: (v.clone() if isinstance(v, torch.Tensor) else list(v)) for k, v in batch.items()}


def spkr_dict(n_spkrs):
    return {f"spk{i}": i for i in range(n_spkrs)}
