"""CPU: the oracle port against the committed reference-generated golden vectors (tests/golden/)."""
import os

import numpy as np
import pytest

import golden_cases
import sb_testutil as util
from oracle import slimt_oracle as so
from slimt_b200 import synth


@pytest.mark.parametrize("name", sorted(golden_cases.FORWARD_CASES))
def test_forward_golden(name, tmp_path):
    path, tokens, lengths, sl, forced, g = golden_cases.load_case(name, tmp_path)
    out = so.Oracle(synth.read_model(path)).forward(tokens, lengths, shortlist=sl, forced=forced, keep=True)
    align = np.stack([a[:, 0, 0, :] for a in out["attn"]])
    golden_cases.check_against_golden(g, lengths, out["step_tokens"], out["encoder_out"], np.stack(out["logits"]), align)
    assert [len(s) for s in out["sentences"]] == g["sentence_lengths"].tolist()


def test_qmm_golden():
    g = np.load(os.path.join(util.GOLDEN, "qmm_cases.npz"))
    shapes = util.REFERENCE_GEMM_SHAPES + [(24, 256, 1536), (24, 1536, 256)]
    for i, (M, K, N) in enumerate(shapes):
        x, Bt, bias, aq, bq = util.make_qmm_case(100 + i, M, K, N)
        y, qa, _ = so.affine(x, Bt, bias, aq, bq, want=True)
        assert np.array_equal(y, g[f"y_{i}"]), (M, K, N)
        u8 = (qa.astype(np.int16) + 127).astype(np.uint64)
        crc = [int(u8.sum()), int((u8 * (np.arange(u8.size).reshape(u8.shape) % 251 + 1)).sum())]
        assert crc == g[f"qa_crc_{i}"].tolist()
    x, Bt, bias, aq, bq = util.make_qmm_case(200, 16, 256, 4096)
    idx = np.sort(np.random.RandomState(5).choice(4096, 512, replace=False)).astype(np.uint32)
    assert np.array_equal(so.affine(x, Bt, bias, aq, bq, indices=idx), g["y_select"])
