"""Host-side pieces of bench.py that need no GPU: the LPT packing of units onto ranks and the
sub-record launcher's environment (children of a torchrun worker must not look for the agent's store)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_lpt_pack_covers_and_balances():
    import bench
    rng = np.random.default_rng(3)
    costs = (rng.pareto(1.5, size=200) * 100 + 1).tolist()  # long tail, like reads x W over a panel of loci
    for n_bins in (1, 2, 4, 8):
        bins, loads = bench.lpt_pack(costs, n_bins)
        assert sorted(u for b in bins for u in b) == list(range(len(costs)))
        for b, load in zip(bins, loads):
            assert abs(sum(costs[u] for u in b) - load) < 1e-6 * max(load, 1.0)
        # LPT guarantee: the heaviest bin is within 4/3 of the optimum, itself >= max(mean load, largest unit)
        lower = max(sum(costs) / n_bins, max(costs))
        assert max(loads) <= 4.0 / 3.0 * lower + 1e-9
    bins, loads = bench.lpt_pack([], 4)
    assert bins == [[], [], [], []] and loads == [0.0] * 4


def test_subrecord_is_skipped_when_disabled():
    import argparse

    import bench
    args = argparse.Namespace(oversized_sub_reads=0)
    assert bench.oversized_subrecord(args, 0, 1) is None


def test_subrecord_child_environment(monkeypatch):
    """The child gets its own rendezvous port and none of the TORCHELASTIC_* variables (with them init_process_group would
    wait for the parent agent's store on the new port)."""
    import argparse
    import subprocess

    import bench
    seen = {}

    class Done:
        returncode = 0
        stdout = '{"value": 1.0, "config": {"workload": "w"}, "e2e": {"value": 2.0, "unit": "reads/s", "ms_per_step": 3.0}}\n'
        stderr = ""

    def fake_run(cmd, env=None, **kw):
        seen["cmd"], seen["env"] = cmd, env
        return Done()

    monkeypatch.setattr(subprocess, "run", fake_run)
    monkeypatch.setenv("MASTER_PORT", "29500")
    monkeypatch.setenv("TORCHELASTIC_USE_AGENT_STORE", "True")
    monkeypatch.setenv("TORCHELASTIC_RUN_ID", "x")
    sub = bench.oversized_subrecord(argparse.Namespace(oversized_sub_reads=1000), 0, 2)
    assert seen["env"]["MASTER_PORT"] == "29501"
    assert not any(k.startswith("TORCHELASTIC_") for k in seen["env"])
    assert "--workload" in seen["cmd"] and "oversized" in seen["cmd"] and "1000" in seen["cmd"]
    assert sub["value"] == 1.0 and sub["workload"] == "w" and sub["e2e"]["ms_per_step"] == 3.0
    # a rank other than 0 launches the child too (the job is collective) but reports nothing
    assert bench.oversized_subrecord(argparse.Namespace(oversized_sub_reads=1000), 1, 2) is None
