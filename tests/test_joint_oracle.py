"""The py3 restatement of the legacy joint_abundance (oracle/joint_oracle.py) on hand-checked cases."""
import joint_oracle as J


def test_two_clear_alleles():
    # 50 reads only compatible with x, 50 only with y, 20 with both: the pair (x, y) explains everything
    r = J.joint_abundance({"x1": 50, "y1": 50, "x1-y1": 20})
    assert r[0][0] == "x1-y1" and abs(r[0][1] - 1.0) < 1e-12


def test_initial_mass_and_pruning():
    # masses: m[a] = 30 + 5 = 35, m[b] = 5; pairs: a-a 35, a-b 40, b-b 5; best 40 -> keep mass * 2 > 40: a-a, a-b
    r, it = J.joint_abundance({"a": 30, "a-b": 10}, None, True)
    assert [p for p, _ in r] == ["a-b", "a-a"] or [p for p, _ in r] == ["a-a", "a-b"] or len(r) == 1
    assert abs(sum(x for _, x in r) - 1.0) < 1e-12 and it >= 1


def test_substring_rule():
    # "a1" occurs inside "a10": the class {a1} also feeds every pair that holds a10
    r = dict(J.joint_abundance({"a1": 10, "a10": 10}))
    assert abs(sum(r.values()) - 1.0) < 1e-12
    assert "a1-a10" in r


def test_empty():
    assert J.joint_abundance({}) == {}
