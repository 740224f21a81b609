from .vqvae2 import VQVAE2, Quantizer  # noqa
from .spkradv import SpeakerAdversarialNetwork  # noqa
from .loss import CustomFeatureLoss, STFTLoss, MultiSizeSTFTLoss  # noqa
