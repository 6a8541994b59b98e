"""Plane export.  The reference writes deflate-9 TIFFs with tifffile and re-encodes them to tiled pyramidal
JPEG TIFFs with ImageMagick (DigiPathAI/Segmentation.py:333-352); neither tool exists in this image and the
pyramidal writer is scoped as the next row after the hot path.  Until then planes are
written as plain single-level TIFFs through Pillow (float32 'F' mode or uint8 'L' mode), which OpenSlide's
generic-TIFF reader and the viewer's mask layer can open for moderate sizes.
"""
from __future__ import annotations

import numpy as np


def save_plane(path: str, plane) -> None:
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    a = plane.detach().cpu().numpy() if hasattr(plane, "detach") else np.asarray(plane)
    a = np.ascontiguousarray(a)
    if a.dtype == np.uint8:
        im = Image.fromarray(a, mode="L")
    else:
        im = Image.fromarray(a.astype(np.float32), mode="F")
    im.save(path, format="TIFF", compression="tiff_adobe_deflate")
