"""Diploid pair model (hgt_pair_em behind typing_common.joint_abundance) against the Python-3 restatement of the legacy
joint_abundance (oracle/joint_oracle.py; parity unpinned by the reference itself): identical ranked pairs,
probabilities within 1e-6 relative, same iteration count."""
import numpy as np
import pytest

import joint_oracle as J
from conftest import GOLDEN_NAMES, load_golden

pytestmark = pytest.mark.gpu


def check(cmpt):
    from hisatgenotype_b200.typing_common import joint_abundance
    ref, it_ref = J.joint_abundance(dict(cmpt), None, True)
    got, it = joint_abundance(dict(cmpt), None, return_iters=True)
    assert [p for p, _ in got] == [p for p, _ in ref]
    for (_, x), (_, y) in zip(got, ref):
        assert x == pytest.approx(y, rel=1e-6, abs=1e-12)
    assert it == it_ref


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_pair_model_on_reference_tables(name):
    g = load_golden(name)
    for cap in g["loci"]:
        check(cap["Gene_cmpt"])
        if cap["Gene_exons_cmpt"]:
            check(cap["Gene_exons_cmpt"])


def test_pair_model_names_that_contain_each_other():
    """`allele in allele_pair` is a substring test in the legacy code: A*01:01 also matches inside A*01:011."""
    rng = np.random.default_rng(5)
    names = ["A*01:01", "A*01:011", "A*01:0", "A*02:01", "A*02:011", "A*11", "A*11:1", "B*07", "B*07:02"]
    cmpt = {}
    for _ in range(40):
        k = rng.integers(1, 5)
        key = "-".join(sorted(set(names[i] for i in rng.choice(len(names), k, replace=False))))
        cmpt[key] = cmpt.get(key, 0) + int(rng.integers(1, 40))
    check(cmpt)


def test_pair_model_synthetic_many_alleles():
    rng = np.random.default_rng(11)
    names = ["L*%02d:%03d" % (i // 20 + 10, i % 20 + 100) for i in range(600)]
    cmpt = {}
    for _ in range(300):
        g0 = int(rng.integers(0, 30)) * 20
        mem = [names[g0 + j] for j in range(20) if rng.random() < 0.4] or [names[g0]]
        key = "-".join(sorted(mem))
        cmpt[key] = cmpt.get(key, 0) + int(rng.integers(1, 30))
    check(cmpt)


def test_pair_model_empty_and_single():
    from hisatgenotype_b200.typing_common import joint_abundance
    assert joint_abundance({}) == {}
    check({"A*01:01": 10})
    check({"A*01:01-A*02:01": 7})
