"""MNASNet backbone with the module surface of the reference's ``src/models/mnasnet.py`` (ConvBlock :37-62,
SepConv :64-103, MBConv_block :105-137, MBConv :139-173, Mnasnet :175-213), executed by hand-written
sm_100a CUDA kernels through ``mnb200.Engine`` -- there is no torch-operator forward in here.

What is kept identical to the reference (checked by tests/test_structure.py against the oracle, which is
pinned to the live reference): class names and constructor signatures, the ``state_dict`` key list / shapes /
aliasing (shared-weight repeated blocks, mnasnet.py:76-85,162-164), ``modules()`` order, and -- because the
parameter containers are created in the same order with the same initialisers -- bit-identical weights for
a given ``torch.manual_seed``.

What is different: the modules are parameter containers plus a tag (``_mnb``) that tells the engine how to
lower them (conv block / chain / residual block); ``forward`` of ANY of them lowers the subtree to a
kernel program (NHWC, BN statistics fused into the producing conv, BN-apply+ReLU fused into the consumer,
see DESIGN.md) and runs it on the current CUDA stream with full autograd support.
"""
import torch
import torch.nn as nn
from torch.nn import init

from mnb200.engine import run_module

default_activation = nn.ReLU            # mnasnet.py:9 (ReLU6 is commented out in the reference)

__all__ = ['ConvBlock', 'SepConv', 'MBConv_block', 'MBConv', 'Mnasnet', 'MnasNet', '_InvertedResidual']

# (in, out, channel_factor, layers, kernel, reduce): the six MBConv stages, mnasnet.py:181-192
_STAGES = ((16, 24, 3, 3, 3, True), (24, 40, 3, 3, 5, True), (40, 80, 6, 3, 5, True),
           (80, 96, 6, 2, 3, False), (96, 192, 6, 4, 5, True), (192, 320, 6, 1, 3, False))


class _Lowered(nn.Module):
    """Base: forward() runs the CUDA program the engine builds for this subtree."""
    _mnb = None

    def forward(self, input):
        return run_module(self, input)


class ConvBlock(_Lowered):
    """conv(+bias) -> BatchNorm2d(eps 1e-5, momentum 0.1) -> ReLU.  `momentum` is accepted and ignored, as in
    the reference (mnasnet.py:46,55)."""
    _mnb = 'convblock'

    def __init__(self, in_, out_, kernel_size=3, stride=1, padding=0, groups=1,
                 activation=default_activation, momentum=0.1):
        super().__init__()
        if activation is not nn.ReLU:
            raise ValueError("only nn.ReLU is lowered (the reference never uses anything else)")
        self.conv = nn.Conv2d(in_, out_, kernel_size=kernel_size, stride=stride, padding=padding,
                              groups=groups, bias=True)
        self.bn = nn.BatchNorm2d(out_)
        self.activation = activation(inplace=True)


def _dw_pw(cin, cout, k, stride):
    return [ConvBlock(cin, cin, kernel_size=k, stride=stride, padding=k // 2, groups=cin),
            ConvBlock(cin, cout, kernel_size=1, stride=1)]


class SepConv(_Lowered):
    _mnb = 'chain'

    def __init__(self, in_channels, out_channels, kernel_size=3, reduce=False, repeat=0):
        super().__init__()
        stride = 2 if reduce else 1
        # the reference builds the repeated pair even when repeat == 0 (and drops it): keep the RNG draws
        rep = _dw_pw(in_channels, in_channels, kernel_size, stride)
        last = _dw_pw(in_channels, out_channels, kernel_size, stride)
        self.sequence = nn.Sequential(*(rep * repeat + last))


class MBConv_block(_Lowered):
    """x + [1x1 expand, depthwise kxk, 1x1 project](x); every conv is a full ConvBlock and the skip is added
    after the last ReLU (mnasnet.py:131-133)."""
    _mnb = 'resblock'

    def __init__(self, in_channels, channel_factor, kernel_size=3):
        super().__init__()
        self.in_channels = in_channels
        mid = in_channels * channel_factor
        self.sequence = nn.Sequential(
            ConvBlock(in_channels, mid, kernel_size=1, stride=1),
            ConvBlock(mid, mid, kernel_size=kernel_size, stride=1, padding=kernel_size // 2, groups=mid),
            ConvBlock(mid, in_channels, kernel_size=1, stride=1))


class MBConv(_Lowered):
    _mnb = 'chain'

    def __init__(self, in_channels, out_channels, channel_factor, layers, kernel_size=3, reduce=True,
                 cut_channels_first=True):
        super().__init__()
        transition = ConvBlock(in_channels, out_channels, kernel_size=3, stride=2 if reduce else 1, padding=1)
        # ONE block object repeated `layers` times: weights are shared and BN buffers are updated once per
        # application (mnasnet.py:162-164)
        block = MBConv_block(out_channels if cut_channels_first else in_channels, channel_factor, kernel_size)
        seq = [transition] + [block] * layers
        self.sequence = nn.Sequential(*(seq if cut_channels_first else seq[::-1]))


class Mnasnet(_Lowered):
    _mnb = 'chain'

    def __init__(self, cut_channels_first=True):
        super().__init__()
        mods = [ConvBlock(3, 32, kernel_size=3, stride=2, padding=1), SepConv(32, 16, kernel_size=3)]
        for cin, cout, f, layers, k, reduce in _STAGES:
            mods.append(MBConv(cin, cout, channel_factor=f, layers=layers, kernel_size=k, reduce=reduce,
                               cut_channels_first=cut_channels_first))
        self.features = nn.Sequential(*mods)
        self.init_params()

    @property
    def sequence(self):          # lowering view: a chain over `features`
        return self.features

    def init_params(self):
        """Kaiming-normal(fan_out) convs, unit BN, N(0, 1e-3) Linear -- mnasnet.py:197-209."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                init.kaiming_normal_(m.weight, mode='fan_out')
                if m.bias is not None:
                    init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                init.constant_(m.weight, 1)
                init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                init.normal_(m.weight, std=0.001)
                if m.bias is not None:
                    init.constant_(m.bias, 0)


_InvertedResidual = MBConv_block     # name used by BASELINE.json's north_star


def MnasNet(num_classes=1000, classifier_config=512):
    """north_star alias: FineTuneModelPool(load_model('mnasnet'), 'mnasnet', num_classes, str(cfg))."""
    from models.classifiers import FineTuneModelPool, load_model
    return FineTuneModelPool(load_model('mnasnet'), 'mnasnet', num_classes, str(classifier_config))
