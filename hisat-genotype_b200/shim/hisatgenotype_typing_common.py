"""Drop-in for the reference's hisatgenotype_modules/hisatgenotype_typing_common.py.

Every name of the reference module is re-exported unchanged; single_abundance(Gene_cmpt, remove_low_abundance_allele,
Gene_length) (reference hisatgenotype_typing_common.py:1282-1410) is replaced by the GPU implementation
(hisat-genotype_b200/typing_common.py -> libhgt hgt_em).  The reference's typing() reaches it through the module attribute
`typing_common.single_abundance` (core:1734, 1767, 1789), i.e. through this module.  See _hgt_shim.py for the set-up.
"""
import _hgt_shim

_reference = _hgt_shim.load_reference("hisatgenotype_typing_common")
_hgt_shim.reexport(_reference, globals())

if not _hgt_shim.disabled():
    _hgt_shim.product()
    from hisatgenotype_b200.typing_common import single_abundance  # noqa: E402,F401
