#!/usr/bin/env python
"""Text in, text out on the GPU: how much of a request is host work.

Trains a 32000-piece unigram vocabulary on a seeded synthetic corpus with the sentencepiece wheel (cached in /tmp),
writes the tiny11-shaped synthetic model and shortlist of the headline benchmark, generates 4096 lines of text of about
32 tokens each, and runs tools/_bin/text_bench (the C++ services, include/slimt_b200.hh) for several host thread
counts.  Prints one JSON line per configuration; `python tools/text_bench.py > profiles/<tag>_text_bench.jsonl`."""
import json
import os
import random
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

from slimt_b200 import synth  # noqa: E402
import make_text_golden as g  # noqa: E402

TMP = "/tmp/slimt_b200_text_bench"


def assets(n_lines=4096):
    import sentencepiece as spm
    os.makedirs(TMP, exist_ok=True)
    rng = random.Random(11)
    words = [g.word(rng, g.LATIN, 2, 5) for _ in range(60000)]

    def line(n):
        ws = [words[min(int(rng.paretovariate(1.1)) - 1, len(words) - 1)] if rng.random() < 0.7 else rng.choice(words) for _ in range(n)]
        ws[0] = ws[0].capitalize()
        return " ".join(ws) + rng.choice(".!?")
    vocab = os.path.join(TMP, "spm32k.model")
    if not os.path.exists(vocab):
        corpus = os.path.join(TMP, "corpus.txt")
        with open(corpus, "w") as f:
            for _ in range(200000):
                f.write(line(rng.randint(4, 16)) + "\n")
        spm.SentencePieceTrainer.train(input=corpus, model_prefix=vocab[:-6], vocab_size=32000, model_type="unigram", bos_id=-1,
                                       eos_id=0, unk_id=1, pad_id=-1, num_threads=8, minloglevel=2, input_sentence_size=200000)
    sp = spm.SentencePieceProcessor(model_file=vocab)
    text = os.path.join(TMP, "text.txt")
    rng = random.Random(12)
    lines, tokens = [], 0
    while len(lines) < n_lines:
        s = line(rng.randint(20, 34))
        n = len(sp.encode(s)) + 1
        if 24 <= n <= 40:
            lines.append(s)
            tokens += n
    open(text, "w").write("\n".join(lines) + "\n")
    model = os.path.join(TMP, "tiny11.bin")
    if not os.path.exists(model):
        synth.write_model(model, synth.make_params(synth.TINY, seed=1234))
    sl = os.path.join(TMP, "lex.bin")
    if not os.path.exists(sl):
        fr, offs, lists = synth.make_shortlist(vocab=32000, frequent=100, best=100, seed=7)
        synth.write_shortlist(sl, fr, offs, lists, best=100)
    return model, vocab, sl, text, tokens


def main():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "slimt_b200", "csrc"), "text_bench"])
    devices = sys.argv[1] if len(sys.argv) > 1 else "0"
    n_gpus = len(devices.split(","))
    # one GPU: one 4096-line batch, as the headline; N GPUs: two such batches per replica from ONE request
    model, vocab, sl, text, tokens = assets(4096 if n_gpus == 1 else 4096 * 2 * n_gpus)
    for workers in ((1, 4, 8, 16) if n_gpus == 1 else (4, 8, 16, 32)):
        r = subprocess.run([os.path.join(ROOT, "tools", "_bin", "text_bench"), model, vocab, sl, text, str(workers), "5", str(4096 * 40), devices],
                           capture_output=True, text=True)
        if r.returncode != 0:
            print(json.dumps({"error": r.stderr.strip()[-300:]}))
            return 1
        d = json.loads(r.stdout)
        d["target_tokens_per_s_e2e_text"] = round(d["target_tokens"] / d["e2e_s"])
        d["target_tokens_per_s_word_ids"] = round(d["target_tokens"] / d["words_s"])
        d["source_tokens_per_s_tokenize"] = round(d["source_tokens"] / d["tokenize_s"])
        d["host_share_of_e2e"] = round(1.0 - d["words_s"] / d["e2e_s"], 3)
        d["workload"] = f"tiny11 int8 + shortlist, {d['sentences']} lines of synthetic text (24-40 tokens each), 32000-piece unigram vocabulary, sentence mode"
        print(json.dumps(d), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
