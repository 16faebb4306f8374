"""Importable alias of the product package.

The package directory is named ``enhancing-3d-point-cloud-segmentation-using-multi-modal-fusion-
with-2d-images_b200`` (not a Python identifier), so ``import mvkpconv_b200`` loads it through
importlib and aliases the module objects.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

PACKAGE_DIR_NAME = "enhancing-3d-point-cloud-segmentation-using-multi-modal-fusion-with-2d-images_b200"

_pkg = importlib.import_module(PACKAGE_DIR_NAME)
sys.modules[__name__] = _pkg
for _k, _v in list(sys.modules.items()):
    if _k.startswith(PACKAGE_DIR_NAME + "."):
        sys.modules[__name__ + _k[len(PACKAGE_DIR_NAME):]] = _v
