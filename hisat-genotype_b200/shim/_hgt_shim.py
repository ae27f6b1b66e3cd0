"""Plumbing of the drop-in modules in this directory.

The reference finds its modules through PYTHONPATH (`setup.sh`: PYTHONPATH=$HG_DIR/hisatgenotype_modules; the CLI does
`from hisatgenotype_typing_core import genotyping_locus`, hisatgenotype:36).  Put THIS directory in front of the
reference's hisatgenotype_modules:

    PYTHONPATH=<repo>/hisat-genotype_b200/shim:<reference>/hisatgenotype_modules  hisatgenotype --base hla ...

`hisatgenotype_typing_core` and `hisatgenotype_typing_common` then resolve to the files here.  Each loads the reference
module of the same name from the NEXT place on sys.path (or $HGT_REFERENCE_MODULES), re-exports every name of it and
replaces the two entry points of the hot path: typing() (stage a + EM glue + report) and single_abundance() (EM).
HGT_DISABLE=1 leaves the reference untouched (pure pass-through); HGT_DEVICE picks the GPU.
"""
import importlib.util
import os
import sys

SHIM_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(os.path.dirname(SHIM_DIR))


def disabled():
    return os.environ.get("HGT_DISABLE", "") not in ("", "0")


def reference_modules_dir():
    env = os.environ.get("HGT_REFERENCE_MODULES", "")
    if env:
        return env
    for p in sys.path:
        d = os.path.abspath(p or ".")
        if d == SHIM_DIR:
            continue
        if os.path.exists(os.path.join(d, "hisatgenotype_typing_core.py")) and \
                os.path.exists(os.path.join(d, "hisatgenotype_typing_common.py")):
            return d
    raise ImportError("hisat-genotype_b200 shim: the reference's hisatgenotype_modules directory is not on sys.path "
                      "(put it after %s in PYTHONPATH, or set HGT_REFERENCE_MODULES)" % SHIM_DIR)


def load_reference(name):
    """The reference module `name`, executed under the alias _hgt_ref_<name> (its own file, its own globals)."""
    alias = "_hgt_ref_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    path = os.path.join(reference_modules_dir(), name + ".py")
    spec = importlib.util.spec_from_file_location(alias, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        del sys.modules[alias]
        raise
    return mod


def reexport(ref, into):
    for k, v in vars(ref).items():
        if k.startswith("__") and k.endswith("__"):
            continue
        into[k] = v


def product():
    """The package of this repository (the directory name has a hyphen: it is loaded by path as hisatgenotype_b200)."""
    if REPO_ROOT not in sys.path:
        sys.path.append(REPO_ROOT)
    import _hgt_path
    return _hgt_path.load()
