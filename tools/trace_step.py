"""Phase timeline of the decoder's row-tile kernels: runs the bench workload once with SLIMT_B200_TRACE set and
prints, per kernel, the median over CTAs of every phase stamp (cycles since the CTA started).
usage (GPU box): python tools/trace_step.py [out.txt]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from slimt_b200 import capi  # noqa: E402

FFN = {0: "start", 1: "setup done", 2: "mma: operand a landed", 3: "mma: Wo issued", 4: "epi: Wo done", 5: "epi: x parked",
       6: "epi: LN1 stats", 7: "epi: y operand ready", 8: "mma: y seen", 9: "epi: FFN2 done", 10: "epi: x2 parked",
       11: "epi: LN2 stats", 12: "epi: z written", 13: "exit", 14: "(sum) mma thread waiting for weight tiles",
       15: "(sum) mma thread waiting for requantised blocks"}
for j in range(12):
    FFN[16 + j] = f"mma: FFN1 block {j} issued"
    FFN[28 + j] = f"mma: FFN2 k-step {j} issued"
    FFN[40 + j] = f"epi: FFN1 block {j} in TMEM"
    FFN[52 + j] = f"epi: block {j} requantised"
FINE = ["entry", "tile landed", "fence", "mma 0", "mma 1", "mma 2", "mma 3", "commit"]
for i, nm in enumerate(FINE):
    FFN[64 + i] = f"fine FFN1 block 5 kb 0: {nm}"
    FFN[72 + i] = f"fine FFN1 block 5 kb 1: {nm}"
    FFN[80 + i] = f"fine FFN2 k-step 3 mb 0: {nm}"
    FFN[88 + i] = f"fine FFN2 k-step 3 mb 1: {nm}"
SSRU = {0: "start", 1: "setup done", 2: "mma: x landed", 3: "mma: Wf,W issued", 14: "epi: state/x rows loaded", 4: "epi: Wf,W done",
        5: "epi: x parked", 6: "epi: LN stats", 7: "epi: h operand ready", 8: "mma: h seen", 9: "mma: Wq issued", 10: "epi: Wq done",
        11: "epi: q written", 13: "exit", 15: "epi: row 0's bookkeeping done", 16: "epi: tile's tokens known",
        17: "epi: x operands written"}
CROSS = {0: "start"}
RC = ["group: q stored", "after group barrier", "K accumulators ready", "K in registers", "scores formed", "softmax done",
      "after mid barrier", "V accumulators ready", "V in registers", "V sums formed", "V emitted"]
for it in range(6):
    for i, nm in enumerate(RC):
        CROSS[8 + 16 * it + i] = f"group {it}: {nm}"


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "trace_step.txt")
    raw = tempfile.mktemp(prefix="sb_trace_")
    ctx = capi.Context(0)
    tmp = tempfile.mkdtemp(prefix="slimt_b200_trace_")
    model_path, sl_path, shortlist, sentences = bench.build_assets(tmp, 0)
    model = capi.Model(ctx, open(model_path, "rb").read())
    sl_bin = open(sl_path, "rb").read()
    tok = np.zeros((len(sentences), bench.SRC_LEN), dtype=np.uint32)
    lens = np.zeros(len(sentences), dtype=np.uint32)
    for r, s in enumerate(sentences):
        tok[r, :len(s)] = s
        lens[r] = len(s)
    sl = capi.shortlist_generate(sl_bin, np.concatenate(sentences), model.V)
    for i in range(3):
        if i == 2:
            os.environ["SLIMT_B200_TRACE"] = raw
        model.forward(tok, lens, shortlist=sl)
    os.environ.pop("SLIMT_B200_TRACE")
    rows = {"ssru": [], "ffn": [], "encffn": [], "cross": [], "out": []}
    for line in open(raw):
        p = line.split()
        rows[p[0]].append([int(x) for x in p[2:]])
    with open(out_path, "w") as f:
        for name, names in (("ssru", SSRU), ("ffn", FFN), ("encffn", FFN), ("cross", CROSS)):
            if not rows[name]:
                continue
            a = np.array(rows[name], dtype=np.float64)
            a[a < 0] = np.nan
            med, lo, hi = np.nanmedian(a, axis=0), np.nanmin(a, axis=0), np.nanmax(a, axis=0)
            f.write(f"== {name}: {len(a)} CTAs, cycles since CTA start (median [min, max])\n")
            for slot in sorted(names, key=lambda k: (med[k], k)):
                if not np.isnan(med[slot]):
                    f.write(f"  {med[slot]:9.0f} [{lo[slot]:8.0f}, {hi[slot]:8.0f}]  {names[slot]}\n")
        if rows["out"]:  # output GEMM: one line per CTA (stamps 1-3 since entry, then counts and accumulated waits)
            a = np.array(rows["out"], dtype=np.int64)
            f.write(f"== out: {len(a)} CTAs; per CTA: released, epilogue done, exit (cycles since entry) | tiles, weight tiles | "
                    "MMA thread waited for weights, accumulator buffer, activations | exact-path strips (warp 4)\n")
            busy = a[:, 3] - a[:, 1]
            f.write(f"  busy (exit - released): min {busy.min()} median {int(np.median(busy))} max {busy.max()}\n")
            for key, label in ((5, "weight tiles"), (4, "tiles")):
                for v in sorted(set(a[:, key].tolist())):
                    sel = a[:, key] == v
                    f.write(f"  {label} = {v}: {int(sel.sum())} CTAs, busy median {int(np.median(busy[sel]))}, waits (weights, buffer, activations) "
                            f"{int(np.median(a[sel, 6]))} {int(np.median(a[sel, 7]))} {int(np.median(a[sel, 8]))}, exact strips {int(np.median(a[sel, 9]))}\n")
            f.write("  busy by CTA: " + " ".join(str(int(b) // 100) for b in busy) + "  (x100 cycles)\n")
            f.write("  buffer waits by CTA: " + " ".join(str(int(b) // 100) for b in a[:, 7]) + "\n")
            f.write("  exact strips by CTA: " + " ".join(str(int(b)) for b in a[:, 9]) + "\n")
            f.write("  exact groups by CTA: " + " ".join(str(int(b)) for b in a[:, 10]) + "\n")
            f.write("  correlation(busy, exact groups) = %.3f\n" % np.corrcoef(busy, a[:, 10])[0, 1])
            order = np.argsort(busy)
            for cta in list(order[:5]) + list(order[-8:]):
                f.write("  cta %3d: released %6d epi %6d exit %6d | %2d tiles %d weight tiles | waits %6d %6d %6d | exact %3d\n" %
                        (cta, a[cta, 1], a[cta, 2], a[cta, 3], a[cta, 4], a[cta, 5], a[cta, 6], a[cta, 7], a[cta, 8], a[cta, 9]))
    print(open(out_path).read())


if __name__ == "__main__":
    main()
