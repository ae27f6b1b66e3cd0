"""Make the package in `hisat-genotype_b200/` importable as `hisatgenotype_b200` (the directory name the
project prescribes contains a hyphen, which Python cannot import directly)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "hisat-genotype_b200")


def load():
    if "hisatgenotype_b200" in sys.modules:
        return sys.modules["hisatgenotype_b200"]
    spec = importlib.util.spec_from_file_location(
        "hisatgenotype_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["hisatgenotype_b200"] = mod
    spec.loader.exec_module(mod)
    return mod
