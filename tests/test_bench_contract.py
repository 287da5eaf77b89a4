"""bench.py host logic that needs no GPU: the source hash that ties `roofline.traffic` to the code it was measured
on, and the algorithmic-bytes model of SURVEY.md 8(d)."""
import json
import os
import shutil

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_source_hash_ignores_comments_but_not_code(tmp_path, monkeypatch):
    csrc = os.path.join(ROOT, "parallel_dmd_for_biomolecules_b200", "csrc")
    pkg = tmp_path / "parallel_dmd_for_biomolecules_b200"
    shutil.copytree(csrc, pkg / "csrc", ignore=shutil.ignore_patterns("*.so", "*.o"))
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    h0 = bench.kernel_source_hash()
    monkeypatch.setattr(bench, "ROOT", ROOT)
    assert h0 == bench.kernel_source_hash()  # same files, other place
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    f = pkg / "csrc" / "dmd_math.h"
    text = f.read_text()
    f.write_text("// a comment\n" + text.replace("\n", "\n   ", 3) + "\n/* another\n one */\n")
    assert bench.kernel_source_hash() == h0  # comments and white space do not count
    f.write_text(text + "\nnamespace dmd { inline int added_function() { return 1; } }\n")
    assert bench.kernel_source_hash() != h0  # code does


def test_committed_traffic_stamp_matches_the_sources():
    """profiles/event_loop_traffic.json must have been captured from the code in the tree, or bench.py reports
    roofline.traffic = null"""
    with open(os.path.join(ROOT, "profiles", "event_loop_traffic.json")) as f:
        stamp = json.load(f)
    assert 2000 < stamp["dram_bytes_per_event"] < 20000
    if stamp["source_sha16"] != bench.kernel_source_hash():
        pytest.skip("stale stamp: the kernel sources changed since the capture (tools/evidence_pass.sh + tools/stamp_traffic.py)")


def test_algorithmic_bytes_model():
    # pair event 256 B + 84 B per list entry visited; interval event 72 N + 16 (N + 3)
    n = 1344
    assert bench.algorithmic_bytes(n, 1, 1, 0, 27) == 256 + 27 * 84
    assert bench.algorithmic_bytes(n, 1, 0, 0, 0) == 72 * n + 16 * (n + 3)
    assert bench.algorithmic_bytes(n, 10, 8, 1, 200) == 8 * 256 + 200 * 84 + (64 + 840) + (72 * n + 16 * (n + 3))
