"""Layer geometry: C library == oracle restatement == the reference's own layer_multipliers table
(maua/GAN/wrappers/stylegan3.py:15-19 -- the one thing the reference pins about the absent network)."""
import pytest

from maua_b200.GAN.networks import stylegan3 as N
from maua_b200.GAN.wrappers.stylegan3 import layer_multipliers
from oracle import sg3 as O

KEYS = ["name", "is_torgb", "is_critically_sampled", "use_fp16", "in_channels", "out_channels", "in_size", "out_size",
        "in_sampling_rate", "out_sampling_rate", "tmp_sampling_rate", "conv_kernel", "up", "down", "up_taps", "down_taps",
        "down_radial"]


@pytest.mark.parametrize("res", [256, 512, 1024])
@pytest.mark.parametrize("config", ["T", "R"])
def test_geometry_matches_oracle(res, config):
    kw = dict(O.SG3_R_KWARGS) if config == "R" else {}
    og = O.sg3_geometry(img_resolution=res, **kw)
    cg = N.sg3_geometry(N.sg3_cfg(img_resolution=res, **kw))
    assert cg["input"] == pytest.approx(og["input"])
    assert len(cg["layers"]) == len(og["layers"]) == 15
    for a, b in zip(cg["layers"], og["layers"]):
        for k in KEYS:
            assert a[k] == b[k], (a["name"], k)
        assert [a["pad_lo"], a["pad_hi"]] == b["padding"][:2]
        for k in ["in_cutoff", "out_cutoff", "in_half_width", "out_half_width"]:
            assert a[k] == pytest.approx(b[k], rel=1e-12)


@pytest.mark.parametrize("res", [256, 512, 1024])
def test_reference_layer_multipliers(res):
    """layer k of the reference table = input (k=0) or output of layer k-1: img_resolution / (size - 20)."""
    geo = N.sg3_geometry(N.sg3_cfg(img_resolution=res))
    sizes = [geo["input"]["size"]] + [g["out_size"] for g in geo["layers"]]
    for k, mult in layer_multipliers[res].items():
        if k >= len(sizes):
            continue
        size = sizes[k]
        if size == res:  # last layers are emitted without margin
            assert mult == 1
        else:
            assert res / (size - 20) == mult, (k, size)


def test_t1024_table_matches_survey_appendix():
    names = [g["name"] for g in N.sg3_geometry(N.sg3_cfg())["layers"]]
    assert names == ["L0_36_512", "L1_36_512", "L2_52_512", "L3_52_512", "L4_84_512", "L5_148_512", "L6_148_512",
                     "L7_276_323", "L8_276_203", "L9_532_128", "L10_1044_81", "L11_1044_51", "L12_1044_32",
                     "L13_1024_32", "L14_1024_3"]
