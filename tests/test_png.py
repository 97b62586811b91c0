"""Native PNG writer (csrc/png.cu, parallel deflate) against PIL, the library the reference saves its images with
(evalscripts/generate-images-sd.py:45-46).  Host-only code: no GPU needed."""
import numpy as np
import pytest
from PIL import Image

from uce_b200.png import save_png


def _images():
    rng = np.random.default_rng(0)
    yy, xx = np.mgrid[0:512, 0:512]
    smooth = np.stack([(xx // 2) % 256, (yy // 2) % 256, ((xx + yy) // 4) % 256], -1).astype(np.uint8)     # image-like: filters matter
    noise = rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)                                            # incompressible
    tiny = rng.integers(0, 256, (1, 1, 3), dtype=np.uint8)
    odd = rng.integers(0, 256, (17, 5, 3), dtype=np.uint8)
    flat = np.full((100, 300, 3), 7, np.uint8)
    return {"smooth": smooth, "noise": noise, "tiny": tiny, "odd": odd, "flat": flat}


@pytest.mark.parametrize("threads", [1, 2, 7, 0])
def test_png_roundtrip_is_lossless(tmp_path, threads):
    for name, img in _images().items():
        p = tmp_path / f"{name}_{threads}.png"
        save_png(p, img, threads=threads)
        with Image.open(p) as im:
            im.load()
            assert im.mode == "RGB" and im.size == (img.shape[1], img.shape[0])
            assert np.array_equal(np.asarray(im), img), (name, threads)


def test_png_accepts_pil_and_compresses(tmp_path):
    img = _images()["smooth"]
    p, q = tmp_path / "ours.png", tmp_path / "pil.png"
    save_png(p, Image.fromarray(img))                      # what generate_images hands over (decode_latents_to_pil)
    Image.fromarray(img).save(q)
    assert np.array_equal(np.asarray(Image.open(p)), img)
    assert p.stat().st_size < img.size // 4               # filters + deflate at work
    assert p.stat().st_size < 3 * q.stat().st_size        # same ballpark as PIL's single-stream encoder
    for level in (0, 1, 9, 42):
        save_png(p, img, level=level, threads=3)
        assert np.array_equal(np.asarray(Image.open(p)), img)


def test_png_rejects_bad_input(tmp_path):
    with pytest.raises(ValueError):
        save_png(tmp_path / "x.png", np.zeros((4, 4), np.uint8))
    with pytest.raises(ValueError):
        save_png(tmp_path / "x.png", np.zeros((4, 4, 3), np.float32))
    from uce_b200._native import UCEError
    with pytest.raises(UCEError):
        save_png(tmp_path / "no_such_dir" / "x.png", np.zeros((4, 4, 3), np.uint8))


def test_png_through_the_generation_driver_path(tmp_path):
    """Exactly what generate_images() does with a pipeline's decode_latents_to_pil() output (tests/test_unet_gpu.py runs the same
    path on the GPU box): PIL images in, {case}_{num}.png out, read back bit-exact."""
    import torch
    from oracle.fake_pipe import FakeGenPipe
    from oracle import unet_oracle as U
    from uce_b200.unet_spec import tiny_config
    cfg = tiny_config(ch=(64, 128), ctx_dim=64, heads=4, groups=8)
    pipe = FakeGenPipe(cfg, U.random_weights(cfg, seed=5), latent_size=16)
    lat = torch.randn(2, 4, 16, 16, generator=torch.Generator().manual_seed(3))
    for num, im in enumerate(pipe.decode_latents_to_pil(lat)):
        p = tmp_path / f"7_{num}.png"
        save_png(str(p), im)
        assert np.array_equal(np.asarray(Image.open(p)), np.asarray(im))
