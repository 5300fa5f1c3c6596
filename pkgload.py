"""Imports the package directory `pdn-jpegxl_b200/` (a hyphen is not a valid module name) as module `pdn_jpegxl_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "pdn-jpegxl_b200")


def load(build_if_missing=False):
    if "pdn_jpegxl_b200" in sys.modules:
        return sys.modules["pdn_jpegxl_b200"]
    if build_if_missing and not os.path.exists(os.path.join(PKG_DIR, "libJpegXLFileTypeIO_X64.so")):
        load_build().build()
    spec = importlib.util.spec_from_file_location("pdn_jpegxl_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["pdn_jpegxl_b200"] = mod
    try:
        spec.loader.exec_module(mod)
    except Exception:
        del sys.modules["pdn_jpegxl_b200"]
        raise
    return mod


def load_build():
    spec = importlib.util.spec_from_file_location("pdn_jpegxl_b200_build", os.path.join(PKG_DIR, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
