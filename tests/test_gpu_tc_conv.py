"""GPU: the tcgen05 3xTF32 implicit-GEMM convolution against (a) the fp32 CUDA-core kernel of the same
library and (b) a plain PyTorch fp32 conv2d of the same folded layer, layer by layer."""
import numpy as np
import pytest
import torch

from conftest import golden_cfg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CIN = [64, 64, 64, 64, 128, 128, 128, 128]
NAMES = [("inc.conv.conv.3", "inc.conv.conv.4"), ("down1.mpconv.1.conv.0", "down1.mpconv.1.conv.1"),
         ("down1.mpconv.1.conv.3", "down1.mpconv.1.conv.4"), ("down2.mpconv.1.conv.0", "down2.mpconv.1.conv.1"),
         ("down2.mpconv.1.conv.3", "down2.mpconv.1.conv.4"), ("down3.mpconv.1.conv.0", "down3.mpconv.1.conv.1"),
         ("down3.mpconv.1.conv.3", "down3.mpconv.1.conv.4")]


@pytest.fixture(scope="module")
def model():
    from image_matching_b200 import Matching, synth
    cfg = golden_cfg(max_kp=128)
    sp, sg = synth.superpoint_weights(3, 128), synth.superglue_weights(3, 128)
    m = Matching({"superpoint": dict(cfg["superpoint"], weights=None),
                  "superglue": dict(cfg["superglue"], weights="")}).eval()
    m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sp.items()})
    m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
    return m.to(DEV), sp


def _torch_ref(sp, layer, x):
    from oracle.matching_oracle import fold_bn
    if layer < 7:
        conv, bn = NAMES[layer]
        w, b = fold_bn(sp[conv + ".weight"], sp[conv + ".bias"], sp, bn)
    else:
        wa, ba = fold_bn(sp["convPa.weight"], sp["convPa.bias"], sp, "bnPa")
        wd, bd = fold_bn(sp["convDa.weight"], sp["convDa.bias"], sp, "bnDa")
        w, b = np.concatenate([wa, wd]), np.concatenate([ba, bd])
    # float64 reference without cuDNN (its first fp64 call can stall for minutes on a cold box):
    # unfold + one matmul
    xd = x.double()
    n, c, H, W = xd.shape
    cols = torch.nn.functional.unfold(xd, 3, padding=1)                        # (n, c*9, H*W)
    wd = torch.from_numpy(w).to(DEV).double().reshape(w.shape[0], -1)
    y = torch.relu(wd @ cols + torch.from_numpy(b).to(DEV).double()[None, :, None]).reshape(n, -1, H, W)
    if layer in (0, 2, 4):
        y = y[:, :, :H // 2 * 2, :W // 2 * 2].reshape(n, -1, H // 2, 2, W // 2, 2).amax((3, 5))
    return y


@pytest.mark.parametrize("layer", range(8))
@pytest.mark.parametrize("shape", [(2, 48, 64), (1, 37, 53), (3, 16, 16)])
def test_tc_conv_matches_fp32(model, layer, shape):
    from image_matching_b200 import stages
    m, sp = model
    n, H, W = shape
    g = torch.Generator(device=DEV).manual_seed(100 * layer + H)
    x = torch.relu(torch.randn((n, CIN[layer], H, W), device=DEV, generator=g)) * 3.0
    ref = _torch_ref(sp, layer, x)
    simt = stages.debug_conv_layer(m, layer, False, x)
    tc = stages.debug_conv_layer(m, layer, True, x)
    scale = float(ref.abs().max())
    e_simt = float((simt.double() - ref).abs().max()) / scale
    e_tc = float((tc.double() - ref).abs().max()) / scale
    bias = float((tc.double() - ref).mean()) / scale
    print(f"layer {layer} {shape}: rel err simt {e_simt:.2e}  tc(fp16x3) {e_tc:.2e} (mean signed {bias:+.2e})")
    assert e_simt < 2e-6
    assert e_tc < 6e-6          # fp32-class: the fp16 2-term split keeps ~24 mantissa bits per operand (plain TF32: ~5e-4)


def test_fused_stem_is_bit_identical_to_two_kernels():
    """First conv fused into the second conv's operand producer (tc_conv.cu FUSE1) vs conv1_direct + tc_conv3x3:
    same fp32 operation order, so semi / desc must agree bit for bit (odd sizes exercise the image border)."""
    import os
    from image_matching_b200 import Matching, stages, synth
    from conftest import golden_cfg, real_superpoint_weights

    def model():
        cfg = golden_cfg()
        m = Matching({"superpoint": dict(cfg["superpoint"], weights=None),
                      "superglue": dict(cfg["superglue"], weights="")}).eval()
        m.superpoint.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in real_superpoint_weights().items()})
        sg = synth.superglue_weights(0, 128)
        m.superglue.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sg.items()})
        return m.to("cuda:0")

    for (H, W) in [(480, 640), (136, 200)]:
        a, b = synth.make_pair(3, H, W)
        img = torch.from_numpy(np.stack([a, b])[:, None]).cuda()
        fused = stages.superpoint_dense(model(), img)
        os.environ["B200M_STEM_IMPL"] = "unfused"
        try:
            two = stages.superpoint_dense(model(), img)
        finally:
            os.environ.pop("B200M_STEM_IMPL", None)
        for x, y in zip(fused, two):
            assert torch.equal(x, y), float((x - y).abs().max())
