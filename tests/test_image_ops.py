"""C++ image pipeline vs Pillow / NumPy (reference: rust/sglang-mm is tested against PIL for exactness)."""
import numpy as np
import pytest
from PIL import Image

from sglang_b200 import sglb_rt
from sglang_b200.multimodal import processor as P

pytestmark = pytest.mark.skipif(not hasattr(sglb_rt, "image_resize"), reason="native runtime not built")


@pytest.mark.parametrize("shape", [(37, 53, 24, 24), (480, 640, 336, 448), (100, 80, 224, 179), (300, 300, 300, 150), (17, 400, 9, 31)])
def test_resize_is_pillow_exact(shape):
    h, w, oh, ow = shape
    img = np.random.default_rng(h * w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    for f, pf in ((3, Image.BICUBIC), (2, Image.BILINEAR)):
        assert np.array_equal(sglb_rt.image_resize(img, oh, ow, f), np.asarray(Image.fromarray(img).resize((ow, oh), pf)))


def test_clip_preprocess_batch_matches_pillow_path():
    rng = np.random.default_rng(1)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((50, 70, 3), (336, 336, 3), (400, 123, 3), (31, 31, 3))]
    got = P.clip_preprocess_batch(imgs, 48, threads=3)
    ref = np.stack([P.clip_preprocess(a, 48) for a in imgs])
    assert got.shape == (4, 3, 48, 48)
    np.testing.assert_allclose(got, ref, atol=2e-6)


def test_patchify_matches_numpy_and_hash_is_content_sensitive():
    chw = np.random.default_rng(2).standard_normal((3, 56, 84)).astype(np.float32)
    C, H, W, p, m, t = 3, 56, 84, 14, 2, 2
    x = np.repeat(chw[:, None], t, 1).reshape(C, t, H // p // m, m, p, W // p // m, m, p)
    ref = np.ascontiguousarray(x.transpose(2, 5, 3, 6, 0, 1, 4, 7)).reshape((H // p) * (W // p), C * t * p * p)
    np.testing.assert_array_equal(sglb_rt.patchify(chw, p, m, t), ref)
    a = np.zeros((8, 8, 3), np.uint8)
    b = a.copy()
    b[7, 7, 2] = 1
    assert P.image_hash(a) != P.image_hash(b) and P.image_hash(a) == P.image_hash(a.copy())
    assert P.image_hash(np.zeros((8, 9, 3), np.uint8)) != P.image_hash(np.zeros((9, 8, 3), np.uint8)[:8])   # length goes into the hash
