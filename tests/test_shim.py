"""The import path the reference's viewer uses -- ``from DigiPathAI.Segmentation import getSegmentation``
(DigiPathAI/main_server.py:155, README.md:77) -- resolves to the B200 implementation through ``shim/``, and the call
the viewer makes (keyword arguments of main_server.py:165-169) and the README's call (README.md:79-87) bind."""
import importlib
import inspect
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_viewer_import_path_resolves_to_the_b200_module():
    sys.path.insert(0, os.path.join(ROOT, "shim"))
    try:
        for k in [k for k in sys.modules if k == "DigiPathAI" or k.startswith("DigiPathAI.")]:
            del sys.modules[k]
        mod = importlib.import_module("DigiPathAI.Segmentation")
        from digipathai_b200 import Segmentation as ours
        assert mod.getSegmentation is ours.getSegmentation and mod.get_prediction is ours.get_prediction
        sig = inspect.signature(mod.getSegmentation)
        # main_server.run_segmentation (main_server.py:165-169)
        sig.bind(img_path="a.tiff", mask_path="m.tiff", uncertainty_path="u.tiff", status={}, mode="colon")
        # README.md:79-87
        sig.bind(img_path="a.tiff", patch_size=256, stride_size=128, batch_size=32, quick=True, tta_list=None, crf=False,
                 save_path="m.tiff", status=None)
        # usage/usage2.py:40-53
        sig.bind(img_path="a.tiff", probs_path="p.tiff", mask_path="m.tiff", uncertainty_path="u.tiff", mask_level=-1,
                 model="dense", mode="colon", quick=True, tta_list=None, patch_size=256, stride_size=128, batch_size=32)
        # positional order of the reference definition (Segmentation.py:192-205) for its first nine parameters
        names = list(sig.parameters)[:9]
        assert names == ["img_path", "patch_size", "stride_size", "batch_size", "quick", "tta_list", "crf", "save_path",
                         "status"]
    finally:
        sys.path.remove(os.path.join(ROOT, "shim"))
        for k in [k for k in sys.modules if k == "DigiPathAI" or k.startswith("DigiPathAI.")]:
            del sys.modules[k]
