"""Test-only stand-in for the `imageio` package (not installed in this image, no network): the one function reference
inference/predict.py:83 calls.  Lives outside the product package; put on PYTHONPATH only by tests/test_dropin_drivers.py."""
import numpy as np


def mimsave(path, frames, fps=4, loop=0, **kwargs):
    from PIL import Image
    imgs = [Image.fromarray(np.asarray(f, dtype=np.uint8)) for f in frames]
    imgs[0].save(path, save_all=True, append_images=imgs[1:], duration=int(1000 / max(fps, 1)), loop=loop)
