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


def _smp_unet_resnet34_keys(in_ch=1, classes=1):
    """state_dict layout of ``smp.Unet("resnet34", in_channels=1, classes=1)`` written out from its module tree
    (``ResNetEncoder`` = torchvision ResNet-34 minus fc; ``UnetDecoder.blocks[i].conv{1,2}`` = Sequential(conv, bn, relu);
    ``SegmentationHead`` = Sequential(conv, Identity, Identity)) -- the package itself is not installed here."""
    keys = {}
    def bn(prefix, c):
        for n in ("weight", "bias", "running_mean", "running_var"):
            keys[f"{prefix}.{n}"] = (c,)
        keys[f"{prefix}.num_batches_tracked"] = ()
    keys["encoder.conv1.weight"] = (64, in_ch, 7, 7)
    bn("encoder.bn1", 64)
    cin = 64
    for li, (cout, n, stride) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)), start=1):
        for b in range(n):
            p = f"encoder.layer{li}.{b}"
            keys[f"{p}.conv1.weight"] = (cout, cin if b == 0 else cout, 3, 3)
            bn(f"{p}.bn1", cout)
            keys[f"{p}.conv2.weight"] = (cout, cout, 3, 3)
            bn(f"{p}.bn2", cout)
            if b == 0 and (stride != 1 or cin != cout):
                keys[f"{p}.downsample.0.weight"] = (cout, cin, 1, 1)
                bn(f"{p}.downsample.1", cout)
        cin = cout
    enc, dec = (64, 64, 128, 256, 512), (256, 128, 64, 32, 16)
    cins, skips = [enc[4]] + list(dec[:-1]), [enc[3], enc[2], enc[1], enc[0], 0]
    for i, (ci, cs, co) in enumerate(zip(cins, skips, dec)):
        keys[f"decoder.blocks.{i}.conv1.0.weight"] = (co, ci + cs, 3, 3)
        bn(f"decoder.blocks.{i}.conv1.1", co)
        keys[f"decoder.blocks.{i}.conv2.0.weight"] = (co, co, 3, 3)
        bn(f"decoder.blocks.{i}.conv2.1", co)
    keys["segmentation_head.0.weight"] = (classes, 16, 3, 3)
    keys["segmentation_head.0.bias"] = (classes,)
    return keys


def test_state_dict_layout_is_smp_unet():
    """Reference checkpoints are raw smp.Unet state_dicts (REF/src/train/engine.py:264,279): same keys, same shapes."""
    sd = build_unet().state_dict()
    want = _smp_unet_resnet34_keys()
    assert set(sd) == set(want)
    for k, shp in want.items():
        assert tuple(sd[k].shape) == shp, k
    other = build_unet()
    other.load_state_dict({k: torch.zeros(s) if s else torch.tensor(0) for k, s in want.items()})   # strict


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
