"""configs[3]: the fused stage feeding the U-Net consumer.  CPU: topology / signature of ``build_unet``;
GPU: segmentation-input parity (the tensor the network sees is the oracle's normalised image) and the driver."""
import numpy as np
import pytest
import torch

from mri_acl_imagesegmentation_adsp_b200 import synth
from mri_acl_imagesegmentation_adsp_b200.models.unet_factory import ResNet34UNet, build_unet
from oracle import recon_oracle as O


def test_build_unet_signature_and_shapes():
    torch.manual_seed(0)
    net = build_unet("unet", "resnet34", "none", in_ch=1, classes=1)
    assert isinstance(net, ResNet34UNet)
    n_params = sum(p.numel() for p in net.parameters())
    assert 24.0e6 < n_params < 25.0e6          # smp.Unet(resnet34, in_channels=1): 24.4 M parameters
    net.eval()
    with torch.no_grad():
        y = net(torch.randn(2, 1, 64, 96))
    assert y.shape == (2, 1, 64, 96) and torch.isfinite(y).all()
    for bad in (dict(model="unetpp"), dict(encoder="resnet50"), dict(encoder_weights="imagenet")):
        with pytest.raises(ValueError):
            build_unet(**bad)
    with pytest.raises(ValueError):
        net(torch.randn(1, 1, 60, 64))


@pytest.mark.gpu
def test_segmentation_input_parity_and_driver():
    from mri_acl_imagesegmentation_adsp_b200.infer.segment import segment_kspace
    k_np = synth.phantom_kspace(synth.KNEE_SHAPE, 3)[None].repeat(3, 0)
    k_np[1] = synth.gaussian_kspace(synth.KNEE_SHAPE, 4)
    m = synth.knee_mask()
    torch.manual_seed(1)
    net = build_unet().cuda()
    out = segment_kspace(net, torch.from_numpy(k_np).cuda(), m, amp=True, net_batch=2)
    assert out["input"].shape == (3, 1, 320, 320) and out["input"].dtype == torch.float32 and out["input"].is_contiguous()
    assert out["logits"].shape == (3, 1, 320, 320) and out["mask"].dtype == torch.bool
    for s in range(3):
        want, _, _ = O.knee_chain_numpy(k_np[s], m, synth.CROP, "instance")
        assert O.rel_l2(out["input"][s, 0].cpu().numpy(), want) <= 1e-5      # the network sees the reference's image
    ref = net(out["input"]).float()                                           # fp32 forward of the same weights
    assert (torch.sigmoid(ref) > 0.5).eq(out["mask"]).float().mean() > 0.99
