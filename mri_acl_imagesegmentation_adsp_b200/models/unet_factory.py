"""Consumer of the fused input stage (BASELINE.json configs[3]).

``build_unet`` keeps the signature of ``REF/src/models/unet_factory.py:4-32``.  The reference builds
``segmentation_models_pytorch.Unet("resnet34", in_channels=1, classes=1)``; that package is not installed
here (and ``runs/fastmri_unet/best.pt`` is missing from the reference, ``.MISSING_LARGE_BLOBS:1``), so the
same topology is written out in plain PyTorch: ResNet-34 encoder (stem stride 2, max-pool, stages of
3/4/6/3 basic blocks with 64/128/256/512 channels), five decoder blocks with channels 256/128/64/32/16
(nearest 2x upsample, concatenate the skip, two conv3x3-BN-ReLU), 3x3 segmentation head.  It is the dense
convolution part of the application -- library kernels (cuDNN), not part of the hand-written hot path; it
exists so that configs[3] can be measured end to end and the tensor contract of the stage is exercised.
Weights are seeded random (``data: synthetic`` in every number that involves it).
"""
from __future__ import annotations

import torch
from torch import nn
import torch.nn.functional as F


class _BasicBlock(nn.Module):
    def __init__(self, cin: int, cout: int, stride: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        y = F.relu(self.bn1(self.conv1(x)), inplace=True)
        y = self.bn2(self.conv2(y))
        return F.relu(y + idt, inplace=True)


def _stage(cin: int, cout: int, n: int, stride: int) -> nn.Sequential:
    return nn.Sequential(*[_BasicBlock(cin if i == 0 else cout, cout, stride if i == 0 else 1) for i in range(n)])


class _DecoderBlock(nn.Module):
    def __init__(self, cin: int, cskip: int, cout: int):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(cin + cskip, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
        self.conv2 = nn.Sequential(nn.Conv2d(cout, cout, 3, 1, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class _ResNet34Encoder(nn.Module):
    """Module names of smp's ``ResNetEncoder`` (torchvision ``ResNet`` without ``fc`` / ``avgpool``):
    ``conv1, bn1, relu, maxpool, layer1..layer4``."""

    def __init__(self, in_ch: int):
        super().__init__()
        self.conv1 = nn.Conv2d(in_ch, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = _stage(64, 64, 3, 1)
        self.layer2 = _stage(64, 128, 4, 2)
        self.layer3 = _stage(128, 256, 6, 2)
        self.layer4 = _stage(256, 512, 3, 2)

    def forward(self, x):
        f0 = self.relu(self.bn1(self.conv1(x)))
        f1 = self.layer1(self.maxpool(f0))
        f2 = self.layer2(f1)
        f3 = self.layer3(f2)
        f4 = self.layer4(f3)
        return f0, f1, f2, f3, f4


class _UnetDecoder(nn.Module):
    """smp's ``UnetDecoder``: ``blocks`` = five ``DecoderBlock`` s (``conv1`` / ``conv2`` = conv-BN-ReLU sequentials)."""

    def __init__(self, enc, dec):
        super().__init__()
        skips = [enc[3], enc[2], enc[1], enc[0], 0]
        cins = [enc[4]] + list(dec[:-1])
        self.blocks = nn.ModuleList([_DecoderBlock(ci, cs, co) for ci, cs, co in zip(cins, skips, dec)])

    def forward(self, feats):
        f0, f1, f2, f3, f4 = feats
        y = f4
        for blk, sk in zip(self.blocks, (f3, f2, f1, f0, None)):
            y = blk(y, sk)
        return y


class ResNet34UNet(nn.Module):
    """``smp.Unet("resnet34")`` topology: encoder features at strides 2, 4, 8, 16, 32 with 64, 64, 128, 256, 512
    channels; decoder 256-128-64-32-16; logits at the input resolution.  The module hierarchy carries smp's names
    (``encoder.conv1 / bn1 / layer1..4``, ``decoder.blocks.N.conv1.0 / .1``, ``segmentation_head.0``), so a raw
    ``state_dict`` the reference saves (``REF/src/train/engine.py:264,279``, ``REF/src/train/train_unet.py:227``:
    ``best.pt`` / ``epoch_XXX.pt``) loads with ``load_state_dict`` unchanged."""
    encoder_channels = (64, 64, 128, 256, 512)
    decoder_channels = (256, 128, 64, 32, 16)

    def __init__(self, in_ch: int = 1, classes: int = 1):
        super().__init__()
        self.encoder = _ResNet34Encoder(in_ch)
        self.decoder = _UnetDecoder(self.encoder_channels, self.decoder_channels)
        # smp's SegmentationHead: Sequential(conv3x3, upsampling = Identity, activation = Identity)
        self.segmentation_head = nn.Sequential(nn.Conv2d(self.decoder_channels[-1], classes, 3, 1, 1), nn.Identity(), nn.Identity())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.shape[-1] % 32 or x.shape[-2] % 32:
            raise ValueError(f"input height and width must be divisible by 32, got {tuple(x.shape[-2:])}")
        return self.segmentation_head(self.decoder(self.encoder(x)))


def build_unet(model: str = "unet", encoder: str = "resnet34", encoder_weights: str = "none", in_ch: int = 1,
               classes: int = 1, **kw) -> nn.Module:
    """Same arguments as ``REF/src/models/unet_factory.py:4-32``; ``ValueError`` for what is not available here
    (other encoders, pretrained weights, U-Net++)."""
    if str(encoder_weights).lower() not in ("none", "null"):
        raise ValueError("pretrained encoder weights are not available offline; use encoder_weights='none'")
    if model.lower() != "unet":
        raise ValueError(f"Unsupported model: {model}")
    if encoder.lower() != "resnet34":
        raise ValueError(f"Unsupported encoder: {encoder}")
    if kw:
        raise ValueError(f"unsupported options: {sorted(kw)}")
    return ResNet34UNet(in_ch, classes)
