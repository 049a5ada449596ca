"""The bit-exact device arithmetic (slimt_b200/csrc/exact_math.cuh) against the host libm and against the plain
formulations, on the GPU: expf over the non-positive floats, quantize1, the shared-divisor division and the
branch-free sigmoid quotient.  tools/exact_check.cu sweeps these exhaustively (profiles/r1_exact_check.jsonl holds
that run); here its --quick mode runs in a few seconds."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "_bin", "exact_check")


@pytest.mark.gpu
def test_exact_arithmetic_sweeps(gpu_ctx):
    if not os.path.exists(BIN):
        pytest.skip("tools/_bin/exact_check not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = subprocess.run([BIN, "--quick"], capture_output=True, text=True, timeout=600)
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    assert out.returncode == 0, out.stdout[-2000:]
    kinds = {r["check"].split(",")[0].split(" vs ")[0] for r in rows}
    assert {"expf_glibc_nonpos_tab", "quantize1", "div_by_rcp", "sigmoid"} <= kinds, kinds
    for r in rows:
        assert r["err"] == "no error", r
        assert r["mismatches"] == 0, r
    assert rows[0]["inputs"] > 100_000_000


def test_committed_exhaustive_run_is_clean():
    """The exhaustive sweep recorded on a B200 (not re-run here): every line reports zero mismatches."""
    path = os.path.join(ROOT, "profiles", "r1_exact_check.jsonl")
    rows = [json.loads(l) for l in open(path) if l.startswith("{")]
    assert len(rows) >= 28
    assert all(r["mismatches"] == 0 and r["err"] == "no error" for r in rows)
    assert rows[0]["inputs"] == 2139095041  # every float <= 0 up to -inf
