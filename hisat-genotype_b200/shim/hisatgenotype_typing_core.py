"""Drop-in for the reference's hisatgenotype_modules/hisatgenotype_typing_core.py.

Every name of the reference module is re-exported unchanged, genotyping_locus (core:2278-2309; what the CLI imports,
hisatgenotype:36) included.  typing() (core:249-286) is replaced: the contracted path - graph index, no assembly, not the
genotype-genome / CODIS branches - runs on the GPU (hisat-genotype_b200/typing_core.py -> libhgt), every other call goes to
the reference's own typing() with the same arguments.  The replacement is also installed as the global `typing` of the
reference module, because genotyping_locus calls typing() through its own module globals (core:2582, 2655).
See _hgt_shim.py for the set-up.
"""
import _hgt_shim

_reference = _hgt_shim.load_reference("hisatgenotype_typing_core")
_hgt_shim.reexport(_reference, globals())

if not _hgt_shim.disabled():
    _hgt_shim.product()
    from hisatgenotype_b200 import typing_core as _product  # noqa: E402

    _reference_typing = _reference.typing

    def typing(*args, **kwargs):
        return _product.typing_dispatch(_reference, _reference_typing, *args, **kwargs)

    typing.__doc__ = _product.typing.__doc__
    _reference.typing = typing
