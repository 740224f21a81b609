"""Speaker-adversarial classifier on the encoder outputs.

API mirror of crank/net/module/spkradv.py:20-81.  The gradient-reversal layer
(GradientReversalFunction :63-72, backward = -scale * g) is folded into the input-gradient
epilogue of the classifier's first conv (crk_convstack_bwd dx_scale), so no extra pass exists.
"""

import torch
import torch.nn as nn

from ...parallel_wavegan.models import ParallelWaveGANDiscriminator


class SpeakerAdversarialNetwork(nn.Module):
    def __init__(self, conf, spkr_size=0):
        super().__init__()
        self.conf = conf
        self.spkr_size = spkr_size
        self._construct_net()

    def forward(self, x, detach=False):
        """x: list of (B,T,D_n) encoder outputs -> (B,T,spkr_size)."""
        x = torch.cat(x, dim=-1)
        if detach:
            x = x.detach()
        return self.classifier.forward_cl(x, grad_scale=-self.grl.scale_value)

    def _construct_net(self):
        self.grl = GradientReversalLayer(scale=self.conf["spkradv_lambda"])
        self.classifier = ParallelWaveGANDiscriminator(
            in_channels=sum(self.conf["emb_dim"][: self.conf["n_vq_stacks"]]),
            out_channels=self.spkr_size,
            kernel_size=self.conf["spkradv_kernel_size"],
            layers=self.conf["n_spkradv_layers"],
            conv_channels=64,
            dilation_factor=1,
            nonlinear_activation="LeakyReLU",
            nonlinear_activation_params={"negative_slope": 0.2},
            bias=True,
            use_weight_norm=True,
        )


class GradientReversalFunction(torch.autograd.Function):
    """Stand-alone GRL (identity forward, -scale*g backward) for callers that want the layer itself."""

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = float(scale)
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * (-ctx.scale), None


class GradientReversalLayer(nn.Module):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = torch.tensor(scale)
        self.scale_value = float(scale)

    def forward(self, x):
        return GradientReversalFunction.apply(x, self.scale_value)
