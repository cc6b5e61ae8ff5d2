"""Generator of tests/golden/cbox_reference_128.npz — the one output of the ray-tracing hot path that the reference tree itself holds:
/root/reference/luisa_compute/examples/cbox.png, the 1024 x 1024 RGB image examples/path_tracer.rs writes when its window is closed
(path_tracer.rs:518-529: the Byte4 display image, i.e. accumulated radiance / spp through the sRGB display kernel of :465-479).

Development container only (the reference tree does not travel).  The fixture keeps 8 x 8 block statistics, not the image:
  srgb_sum  uint16 [128,128,3]   sum of the 64 8-bit display values of a block (mean = sum / 64)
  lin_mean  float32 [128,128,3]  mean of the 64 values after undoing the display transform per pixel (sRGB -> linear radiance, clipped at 1)
A 128 x 128 render of the same scene has exactly these blocks as its pixel footprints (the kernel jitters inside the pixel), so the
CPU restatement can be compared at low resolution and the device's 1024 x 1024 render after the same block reduction.

usage: python tests/golden/make_cbox_reference.py
"""
import os

import numpy as np
from PIL import Image

SRC = "/root/reference/luisa_compute/examples/cbox.png"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cbox_reference_128.npz")
BLOCK = 8


def srgb_to_linear(s8):
    s = s8.astype(np.float64) / 255.0
    return np.where(s <= 0.04045, s / 12.92, ((s + 0.055) / 1.055) ** 2.4)


def main():
    img = np.asarray(Image.open(SRC).convert("RGB"))
    assert img.shape == (1024, 1024, 3) and img.dtype == np.uint8
    n = 1024 // BLOCK
    blocks = img.reshape(n, BLOCK, n, BLOCK, 3)
    srgb_sum = blocks.astype(np.uint32).sum(axis=(1, 3)).astype(np.uint16)
    lin_mean = srgb_to_linear(img).reshape(n, BLOCK, n, BLOCK, 3).mean(axis=(1, 3)).astype(np.float32)
    np.savez_compressed(OUT, srgb_sum=srgb_sum, lin_mean=lin_mean, block=np.int32(BLOCK), source=np.bytes_(SRC.encode()))
    print(OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
