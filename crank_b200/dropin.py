"""Recipe-level drop-in: make an unmodified crank checkout import this package's classes.

`crank/bin/train.py:24-37` builds everything from five imports (parallel_wavegan.models, crank.net.module.vqvae2,
crank.net.module.spkradv, crank.net.trainer.TrainerWrapper, crank.net.trainer.utils.get_*).  `install()` registers
this package's modules under those names, so `python -m crank.bin.train ...` of the
Kaldi-style recipes runs the sm_100a kernels with no source change in crank:

    # sitecustomize.py next to the recipe's path.sh (or `python -c "import crank_b200.dropin as d; d.install()" ...`)
    import crank_b200.dropin
    crank_b200.dropin.install()

What stays crank's: dataset / DataLoader (HDF5 I/O), yaml loading, scp handling, checkpoint naming, the CLI.
"""
import importlib
import sys

TRAINER_NAMES = ("TrainerWrapper", "BaseTrainer", "VQVAETrainer", "LSGANTrainer", "CycleGANTrainer", "StarGANTrainer")


def install(patch_trainers=True):
    """Idempotent.  Must run before `crank.bin.train` (or anything that imports crank.net.*) is imported."""
    from . import parallel_wavegan as pw_pkg
    from .net import trainer as tr
    from .net.module import loss, mlfb, spkradv, vqvae2
    from .net.trainer import utils as tu
    from .parallel_wavegan import models as pwg

    if "parallel_wavegan" not in sys.modules:
        sys.modules["parallel_wavegan"] = pw_pkg
    sys.modules["parallel_wavegan"].models = pwg
    sys.modules["parallel_wavegan.models"] = pwg            # vqvae2.py:17, spkradv.py:14, train.py:24-27
    sys.modules["crank.net.module.vqvae2"] = vqvae2          # train.py:29
    sys.modules["crank.net.module.spkradv"] = spkradv        # train.py:28
    sys.modules["crank.net.module.loss"] = loss              # trainer/utils.py:14
    sys.modules["crank.net.module.mlfb"] = mlfb              # vqvae2.py:19
    if not patch_trainers:
        return
    # trainers: same names, same TrainerWrapper(trainer_type, **ka) keyword set (train.py:211-226);
    # crank's own get_dataloader (HDF5 I/O) is kept
    ref_tr = importlib.import_module("crank.net.trainer")
    for name in TRAINER_NAMES:
        if hasattr(tr, name):
            setattr(ref_tr, name, getattr(tr, name))
    ref_tu = importlib.import_module("crank.net.trainer.utils")
    ref_tu.get_criterion, ref_tu.get_optimizer, ref_tu.get_scheduler = tu.get_criterion, tu.get_optimizer, tu.get_scheduler
