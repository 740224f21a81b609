"""N ranks x batch b  ==  1 device x batch N*b, with the REAL kernels (round-1 verdict weak item 3 / next-round item 7).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 profiles/dp_equiv.py

Every rank builds identical models.  Rank 0 first runs two LSGAN steps on the whole ragged batch (N*b utterances) with
data parallelism off; then all ranks run the same two steps data-parallel on their own b utterances (gradient all-reduce,
VQ-EMA statistics all-reduce before the normalisation, exact ragged-mask means, side-stream overlap on).  Compared on
rank 0: every loss key of both steps (<= 1e-4 relative) and the movement of every parameter / EMA buffer.
Writes gpurun_out/dp_equiv.json.
"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist


class W:
    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def build(kind, S, dev):
    from crank_b200.conf import vcc2020_conf
    from crank_b200.net.trainer import TrainerWrapper, get_criterion, get_model, get_optimizer, get_scheduler
    from crank_b200.synthetic import spkr_dict

    conf = vcc2020_conf(trainer_type=kind, n_steps_gan_start=-1, n_steps_cycle_start=-1, discriminator_dropout=0.0)
    random.seed(1234)
    np.random.seed(1234)
    torch.manual_seed(1234)
    pm = get_model(conf, S, device=dev)
    opt = get_optimizer(conf, pm)
    P = TrainerWrapper(kind, model=pm, optimizer=opt, criterion=get_criterion(conf), dataloader={"spkrs": spkr_dict(S)},
                       writer={"train": W(), "dev": W()}, expdir="/tmp/exp", conf=conf, feat_conf=conf["feature"],
                       scheduler=get_scheduler(conf, opt), scaler=None, resume=0, device=dev, n_jobs=1)
    P.tqdm.close()
    return P


def state(P):
    return {f"{k}.{n}": v.detach().double().cpu() for k, m in P.model.items() for n, v in m.state_dict().items()
            if v.dtype.is_floating_point}


def main():
    from crank_b200.net import _dp
    from crank_b200.synthetic import make_batch, to_device

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    kind, S, b, T, STEPS = os.environ.get("DP_KIND", "lsgan"), 14, 8, 300, 2
    whole = [make_batch(world * b, T, S, seed=70 + i, ragged=True) for i in range(STEPS)]

    def part(batch, r):
        return {k: (v[r * b:(r + 1) * b] if isinstance(v, (torch.Tensor, list)) else v) for k, v in batch.items()}

    ref_losses, ref_state, init = None, None, None
    if rank == 0:
        P1 = build(kind, S, dev)
        init = state(P1)
        ref_losses = []
        ref_state0 = None
        for i in range(STEPS):
            random.seed(100 + i)
            ref_losses.append(P1.train(to_device(whole[i], dev), "train"))
            if i == 0:
                ref_state0 = state(P1)
        ref_state = state(P1)
    dist.barrier()
    _dp.enable()
    P = build(kind, S, dev)
    dp_losses = []
    dp_state0 = None
    for i in range(STEPS):
        random.seed(100 + i)
        dp_losses.append(P.train(to_device(part(whole[i], rank), dev), "train"))
        if i == 0:
            _dp.flush()
            torch.cuda.synchronize()
            dp_state0 = state(P)
    _dp.flush()
    torch.cuda.synchronize()
    dp_state = state(P)
    ok = True
    if rank == 0:
        worst_loss = (0.0, "")
        for i in range(STEPS):
            for k, ref in ref_losses[i].items():
                err = abs(dp_losses[i][k] - ref) / max(abs(ref), 1e-12) if ref != 0 else abs(dp_losses[i][k])
                worst_loss = max(worst_loss, (err, f"step{i}.{k}"))
        worst_rms, worst_buf = (0.0, ""), (0.0, "")
        for k, v in ref_state.items():
            mo, mp = v - init[k], dp_state[k] - init[k]
            if mo.abs().max().item() < 1e-12:
                continue
            rms = ((mp - mo).pow(2).mean().sqrt() / mo.pow(2).mean().sqrt()).item()
            if "ema_" in k or "embedding.weight" in k:
                worst_buf = max(worst_buf, (((dp_state[k] - v).abs().max() / v.abs().max().clamp_min(1e-30)).item(), k))
            else:
                worst_rms = max(worst_rms, (rms, k))
        # after the FIRST step no VQ index can have flipped (both runs quantise with identical codebooks): the movement of every
        # tensor must agree to summation-order noise there; later steps may amplify near-tie flips of a barely trained codebook
        worst_rms0 = (0.0, "")
        for k, v in ref_state0.items():
            mo, mp = v - init[k], dp_state0[k] - init[k]
            if mo.abs().max().item() < 1e-12 or "ema_" in k or "embedding.weight" in k:
                continue
            worst_rms0 = max(worst_rms0, (((mp - mo).pow(2).mean().sqrt() / mo.pow(2).mean().sqrt()).item(), k))
        ok = worst_loss[0] <= 1e-4 and worst_rms[0] <= 3e-2 and worst_buf[0] <= 1e-3
        out = {"what": f"{world} ranks x {b} utterances == 1 device x {world * b} utterances, {kind}, T={T}, ragged masks, "
                       f"{STEPS} steps, 3xTF32 kernels, side-stream overlap {'on' if _dp._state['overlap'] else 'off'}",
               "worst_loss_rel_err": worst_loss, "worst_parameter_movement_rms_err": worst_rms,
               "worst_codebook_or_ema_buffer_rel_err": worst_buf, "pass": bool(ok),
               "worst_parameter_movement_rms_err_after_step0": worst_rms0, "pass_step0": bool(worst_rms0[0] <= 1e-3),
               "losses_step1_ref": ref_losses[-1], "losses_step1_dp": dp_losses[-1]}
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(out, open(f"gpurun_out/dp_equiv_{kind}_{world}gpu.json", "w"), indent=1)
        print(json.dumps({k: out[k] for k in ("what", "worst_loss_rel_err", "worst_parameter_movement_rms_err",
                                              "worst_codebook_or_ema_buffer_rel_err", "pass",
                                              "worst_parameter_movement_rms_err_after_step0", "pass_step0")}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
