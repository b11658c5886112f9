"""RRDBNet (RealESRGAN x4) through the C ABI against the CPU oracle restatement (oracle/rrdb.py): the network
at two depths, the uint8 in / out formats, and the RealESRGANer.enhance pre / post-processing of the reference's upscale()."""
import numpy as np
import pytest
import torch

from oracle import rrdb as O

pytestmark = pytest.mark.gpu


def make_pair(num_block, seed=0):
    from maua_b200.super.image.models.realesrgan import RRDBNet

    onet = O.make(num_block=num_block, seed=seed)
    net = RRDBNet(num_block=num_block)
    net.load_state_dict(onet.state_dict(), strict=True)
    return onet, net


@pytest.mark.parametrize("num_block,hw", [(1, (40, 52)), (3, (64, 64)), (6, (33, 47))])
def test_rrdbnet_matches_oracle(cuda, num_block, hw):
    onet, net = make_pair(num_block)
    torch.manual_seed(1)
    x = torch.rand(2, 3, *hw)
    ref = onet(x)
    out = net(x.to(cuda)).cpu()
    assert out.shape == ref.shape == (2, 3, 4 * hw[0], 4 * hw[1])
    err = float((out.clamp(0, 1) - ref.clamp(0, 1)).abs().max())
    rel = float((out - ref).abs().max() / ref.abs().max())
    print(f"RRDBNet x{num_block} {hw}: max-abs pixel error {err:.3e}, relative {rel:.3e}, |out| max {float(ref.abs().max()):.2f}")
    assert err <= 1e-3 and rel <= 2e-3          # fp16 activations / weights, fp32 accumulation
    # uint8 in / out: the fused pipeline's formats
    x8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    ref8 = (onet(x8.permute(0, 3, 1, 2).float() / 255).clamp(0, 1) * 255).round()
    out8 = net(x8.to(cuda), out_fmt="u8").cpu()
    assert out8.shape == (2, 4 * hw[0], 4 * hw[1], 3) and out8.dtype == torch.uint8
    assert float((out8.permute(0, 3, 1, 2).float() - ref8).abs().max()) <= 1
    assert net.last_launch_count() == 2 + 18 * num_block + 7 + 1


def test_enhance_and_upscale_mirror_the_reference_entry_points(cuda, tmp_path):
    from maua_b200.super.image.models import realesrgan as R

    onet, net = make_pair(2, seed=3)
    ckpt = tmp_path / "RealESRGAN_test.pth"
    torch.save({"params_ema": onet.state_dict()}, ckpt)
    model = R.RealESRGANer(scale=4, model_path=str(ckpt), model=R.RRDBNet(num_block=2), tile=0, half=True)
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, size=(37, 45, 3)).astype(np.uint8)
    got, mode = model.enhance(img)
    want = O.enhance(onet, img)
    assert got.shape == want.shape == (148, 180, 3) and got.dtype == np.uint8 and mode == "RGB"
    assert int(np.abs(got.astype(np.int16) - want.astype(np.int16)).max()) <= 1
    # upscale(): tensors in [0, 1] -> float [1, 3, 4H, 4W] (realesrgan.py:44-49)
    t = torch.rand(1, 3, 24, 20)
    (large,) = list(R.upscale([t], model))
    assert large.shape == (1, 3, 96, 80) and float(large.min()) >= 0 and float(large.max()) <= 1
    want2 = torch.from_numpy(O.enhance(onet, t.squeeze().permute(1, 2, 0).mul(255).numpy())).permute(2, 0, 1)[None].float() / 255
    assert float((large - want2).abs().max()) <= 1.01 / 255
    with pytest.raises(FileNotFoundError):
        R.load_model("x4plus")
