"""GPU bring-up diagnostics (not a test): op-level and model-level comparison against the port oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from slimt_b200 import capi, synth
from oracle import slimt_oracle as so

def main():
    ctx = capi.Context(0)
    print(capi.lib().slimt_b200_version().decode(), flush=True)
    rng = np.random.RandomState(0)
    ok_all = True
    for (M, K, N, nidx) in [(8, 256, 256, 0), (128, 256, 256, 0), (200, 256, 1536, 0), (64, 1536, 256, 0),
                            (16, 256, 32000, 1000), (300, 512, 512, 0), (1000, 256, 64, 0)]:
        x = (rng.standard_normal((M, K)) * 1.5).astype(np.float32)
        w = rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)
        bq = np.float32(127 / np.abs(w).max()); Bt = np.clip(np.rint(w * bq), -127, 127).astype(np.int8)
        bias = rng.standard_normal(N).astype(np.float32)
        idx = np.sort(rng.choice(N, nidx, replace=False)).astype(np.uint32) if nidx else None
        aq = 127 / 5.0
        y, qa, acc = ctx.qmm_affine(x, Bt, bias, aq, float(bq), idx, debug=True)
        yo, qao, acco = so.affine(x, Bt, bias, aq, float(bq), indices=idx, want=True)
        e = (np.array_equal(qa, qao), np.array_equal(acc, acco), np.array_equal(y, yo))
        print(f"qmm M{M} K{K} N{N} idx{nidx}: qa={e[0]} acc={e[1]} y={e[2]} maxabs={np.abs(y-yo).max():.3g} "
              f"acc_mismatch={(acc!=acco).mean():.4f}", flush=True)
        if not e[1]:
            bad = np.argwhere(acc != acco)
            print("   first bad", bad[:5].tolist(), acc[tuple(bad[0])], acco[tuple(bad[0])])
        ok_all &= all(e)
    # model
    items = synth.make_params(seed=1234)
    path = "/tmp/sb_tiny.bin"
    synth.write_model(path, items)
    blob = open(path, "rb").read()
    t0 = time.time(); model = capi.Model(ctx, blob); print("model load s", time.time() - t0, model.E, model.F, model.V, flush=True)
    orc = so.Oracle(synth.read_model(path))
    for (B, T) in [(8, 16), (5, 9)]:
        sents = synth.make_sentences(B, (3, T), seed=5); sents[0] = synth.make_sentences(1, T, seed=6)[0]
        tokens = np.zeros((B, T), dtype=np.uint32); lengths = np.array([len(s) for s in sents], dtype=np.uint32)
        for i, s in enumerate(sents): tokens[i, :len(s)] = s
        ref = orc.forward(tokens, lengths, keep=True)
        out = model.forward(tokens, lengths, want_encoder=True, want_logits=True, want_alignment=True)
        valid = np.arange(T)[None, :] < lengths[:, None]
        d = np.abs(out["encoder_out"] - ref["encoder_out"])[valid]
        print(f"B{B} T{T}: encoder_out valid rows exact={np.array_equal(out['encoder_out'][valid], ref['encoder_out'][valid])} maxabs={d.max():.3g}", flush=True)
        n = min(out["steps"], len(ref["step_tokens"]))
        print("  steps gpu/ref", out["steps"], len(ref["step_tokens"]), "target", out["target_tokens"], sum(len(s) for s in ref["sentences"]))
        print("  tokens equal", np.array_equal(out["step_tokens"][:n], ref["step_tokens"][:n]), (out["step_tokens"][:n] == ref["step_tokens"][:n]).mean())
        for s in range(min(n, 3)):
            dl = np.abs(out["logits"][s] - ref["logits"][s])
            print(f"  step{s} logits exact={np.array_equal(out['logits'][s], ref['logits'][s])} maxabs={dl.max():.3g}")
        a_ref = np.stack([a[:, 0, 0, :] for a in ref["attn"][:n]])
        print("  alignment maxabs", np.abs(out["alignment"][:n] - a_ref)[:, valid].max() if n else None)
        out2 = model.forward(tokens, lengths)
        print("  fused-argmax tokens == logits-path tokens:", np.array_equal(out2["step_tokens"], out["step_tokens"]), out2["steps"], out2["target_tokens"])
    print("ALL OP CHECKS", ok_all)

if __name__ == "__main__":
    main()
