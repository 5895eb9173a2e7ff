"""CPU: oracle/preprocess.py (Pillow's 8-bit bicubic resample restated) against Pillow itself and against the
product's host image processor."""
import numpy as np
import pytest
from PIL import Image

from oracle import preprocess as op


@pytest.mark.parametrize("h,w", [(480, 640), (384, 384), (720, 1280), (97, 131)])
def test_resize_matches_pillow_bit_exact(h, w):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 3] = (np.linspace(0, 255, w)[None, :, None]).astype(np.uint8)   # smooth band + noise
    want = np.asarray(Image.fromarray(img).resize((384, 384), resample=Image.BICUBIC))
    got = op.resize_bicubic_u8(img, 384, 384)
    assert np.array_equal(got, want)


def test_preprocess_matches_host_processor():
    from streamvln_b200.vision import SigLipImageProcessor

    img = np.random.default_rng(1).integers(0, 256, size=(480, 640, 3), dtype=np.uint8)
    want = SigLipImageProcessor().preprocess(images=Image.fromarray(img), return_tensors="pt")["pixel_values"][0].numpy()
    assert np.array_equal(op.preprocess(img), want)
    assert want.shape == (3, 384, 384) and want.min() >= -1 and want.max() <= 1
