#!/usr/bin/env python
"""Top-1 / top-2 logit margins of the synthetic model on the headline workload (SURVEY.md section 7 iii): how close the
greedy decisions are to a tie, which is what decides whether a non-bit-exact arithmetic can keep the token sequence.

Runs the first N sentences of the headline batch (tiny11, full vocabulary and with the shortlist) through the GPU path
with the per-step logits tap, exact mode and tolerance mode, and prints one JSON object:
  histogram of (top1 - top2) / |top1| over all (sentence, step) decisions still alive, in decades;
  the share of decisions whose margin is below the 1e-3 relative tolerance north_star allows on logits;
  how input-dependent the outputs are (distinct tokens, distinct sentences);
  tolerance-mode agreement on the same decisions (teacher-forced by construction: both modes see their own history,
  so only the first divergence per sentence is counted as a flipped decision).
usage (GPU box): python tools/margin_hist.py [N] > profiles/<tag>_margin_histogram.json      (CPU only: add --oracle)"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slimt_b200 import capi  # noqa: E402


def margins(out, lengths):
    """(relative margin, alive mask) per (step, sentence) from the logits tap."""
    steps = out["steps"]
    rel = np.zeros((steps, len(lengths)), dtype=np.float64)
    for s in range(steps):
        lg = out["logits"][s]
        part = np.partition(lg, -2, axis=1)
        top1, top2 = part[:, -1].astype(np.float64), part[:, -2].astype(np.float64)
        rel[s] = (top1 - top2) / np.maximum(np.abs(top1), 1e-30)
    toks = out["step_tokens"]
    alive = np.ones_like(rel, dtype=bool)
    for b in range(len(lengths)):
        eos = np.nonzero(toks[:, b] == 0)[0]
        if len(eos):
            alive[eos[0] + 1:, b] = False
    return rel, alive


def oracle_main(n):
    """The same histogram from the CPU oracle (bit-identical logits, tests/test_gpu_model.py): no GPU needed, no
    tolerance-mode part."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import slimt_oracle as so
    from slimt_b200 import synth
    import sb_testutil as util
    tmp = tempfile.mkdtemp(prefix="slimt_b200_margin_")
    model_path, _, _, sentences = bench.build_assets(tmp, 0)
    tokens, lengths = util.pad_batch(sentences[:n])
    ref = so.Oracle(synth.read_model(model_path)).forward(tokens, lengths, keep=True)
    out = {"steps": len(ref["step_tokens"]), "logits": ref["logits"], "step_tokens": np.asarray(ref["step_tokens"])}
    rel, alive = margins(out, lengths)
    r = rel[alive]
    edges = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, np.inf]
    toks = out["step_tokens"]
    report = {"workload": f"first {n} sentences of the headline batch (tiny11 random-init seed {bench.MODEL_SEED}), full vocabulary; CPU oracle",
              "edges_relative_margin": [str(e) for e in edges], "decisions": int(alive.sum()),
              "histogram": np.histogram(r, bins=edges)[0].tolist(),
              "quantiles": {q: float(np.quantile(r, float(q))) for q in ("0.001", "0.01", "0.1", "0.5", "0.9")},
              "share_below_rtol_1e-3": float((r < 1e-3).mean()), "exact_ties": int((r == 0).sum()),
              "distinct_tokens": int(len(np.unique(toks[alive]))),
              "distinct_sentences": int(len({tuple(toks[:, b].tolist()) for b in range(n)})),
              "sentences_reaching_eos": int(sum(1 for b in range(n) if (toks[:, b] == 0).any()))}
    for m in (1e-3, 1e-4, 1e-5):
        report[f"sentences_with_every_margin_above_{m:g}"] = float(np.mean([(rel[alive[:, b], b] >= m).all() for b in range(n)]))
    print(json.dumps(report))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 128
    if "--oracle" in sys.argv:
        return oracle_main(n)
    ctx = capi.Context(0)
    tmp = tempfile.mkdtemp(prefix="slimt_b200_margin_")
    model_path, sl_path, shortlist, sentences = bench.build_assets(tmp, 0)
    model = capi.Model(ctx, open(model_path, "rb").read())
    sents = sentences[:n]
    T = max(len(s) for s in sents)
    tok = np.zeros((n, T), dtype=np.uint32)
    lens = np.array([len(s) for s in sents], dtype=np.uint32)
    for r, s in enumerate(sents):
        tok[r, :len(s)] = s
    edges = [0.0, 1e-6, 1e-5, 1e-4, 1e-3, 1e-2, 1e-1, 1.0, np.inf]
    report = {"workload": f"first {n} sentences of the headline batch (tiny11 random-init seed {bench.MODEL_SEED}, {T} tokens), full vocabulary",
              "edges_relative_margin": [str(e) for e in edges]}
    ctx.set_math(False)
    exact = model.forward(tok, lens, want_logits=True)
    rel, alive = margins(exact, lens)
    r = rel[alive]
    report["decisions"] = int(alive.sum())
    report["histogram"] = np.histogram(r, bins=edges)[0].tolist()
    report["quantiles"] = {q: float(np.quantile(r, float(q))) for q in ("0.001", "0.01", "0.1", "0.5", "0.9")}
    report["share_below_rtol_1e-3"] = float((r < 1e-3).mean())
    report["exact_ties"] = int((r == 0).sum())
    toks = exact["step_tokens"]
    report["distinct_tokens"] = int(len(np.unique(toks[alive])))
    report["distinct_sentences"] = int(len({tuple(toks[:, b].tolist()) for b in range(n)}))
    report["sentences_reaching_eos"] = int(sum(1 for b in range(n) if (toks[:, b] == 0).any()))
    # per sentence: probability that no decision of the whole sentence falls below a relative margin m
    for m in (1e-3, 1e-4, 1e-5):
        ok = [(rel[alive[:, b], b] >= m).all() for b in range(n)]
        report[f"sentences_with_every_margin_above_{m:g}"] = float(np.mean(ok))
    ctx.set_math(True)
    fast = model.forward(tok, lens, want_logits=True)
    ctx.set_math(False)
    ft = fast["step_tokens"]
    steps = min(len(ft), len(toks))
    first_div, at_small_margin = 0, 0
    for b in range(n):
        d = np.nonzero(ft[:steps, b] != toks[:steps, b])[0]
        if len(d) and alive[d[0], b]:
            first_div += 1
            at_small_margin += rel[d[0], b] < 1e-3
    report["tolerance_mode"] = {"sentences_diverging": first_div, "first_divergence_at_margin_below_1e-3": int(at_small_margin),
                                "max_rel_logit_error_step0": float(np.max(np.abs(fast["logits"][0] - exact["logits"][0]) /
                                                                         np.maximum(np.abs(exact["logits"][0]), 1e-3)))}
    print(json.dumps(report))


if __name__ == "__main__":
    main()
