from .basetrainer import BaseTrainer, TrainerWrapper  # noqa
from .trainer_vqvae import VQVAETrainer  # noqa
from .trainer_lsgan import LSGANTrainer  # noqa
from .trainer_cyclegan import CycleGANTrainer  # noqa
from .trainer_stargan import StarGANTrainer  # noqa
from .utils import get_criterion, get_optimizer, get_scheduler, get_model  # noqa
