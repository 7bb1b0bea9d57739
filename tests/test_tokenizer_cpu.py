"""The host-side CLIP BPE tokenizer against the installed transformers.CLIPTokenizer (third-party
arithmetic: pinned here on a synthetic vocabulary, since no real vocabulary ships offline)."""
import collections
import json

import pytest
import torch

from sdb200.tokenizer import BOS, EOS, CLIPBPETokenizer, bytes_to_unicode

CORPUS = ("a photograph of an astronaut riding a horse on mars , highly detailed digital painting . "
          "the quick brown fox jumps over the lazy dog's back ; it's 42 degrees and they're happy ! "
          "stable diffusion latent text to image model , trending on artstation 4k 8k uhd café naïve ") * 3


def _train(n_merges=300):
    """A tiny BPE training run (pair counts over the corpus) to get a vocabulary with real multi-level merges."""
    be = bytes_to_unicode()
    words = collections.Counter()
    import regex
    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""")
    for tok in pat.findall(CORPUS.lower()):
        m = "".join(be[b] for b in tok.encode("utf-8"))
        words[tuple(m[:-1]) + (m[-1] + "</w>",)] += 1
    merges = []
    for _ in range(n_merges):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w, w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), _cnt = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append(f"{a} {b}")
        new = collections.Counter()
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += c
        words = new
    chars = list(be.values())
    vocab = {}
    for c in chars:
        vocab[c] = len(vocab)
    for c in chars:
        vocab[c + "</w>"] = len(vocab)
    for m in merges:
        vocab.setdefault("".join(m.split(" ")), len(vocab))
    vocab[BOS] = len(vocab)
    vocab[EOS] = len(vocab)
    return vocab, merges


PROMPTS = [
    "a photograph of an astronaut riding a horse",
    "A  Painting\tof a   VIRUS monster playing guitar!!!",
    "it's the dog's 42nd birthday, they're happy; we'll see (maybe) ...",
    "café naïve — trending on ArtStation, 8K UHD ✨ 🚀",
    "",
    "   ",
    "under_score and-hyphen and 3.14159 and $100 #tag @user",
    "x" * 40 + " " + "horse " * 120,                       # > 77 tokens: truncation keeps BOS + 75 + EOS
    "an astronaut <|endoftext|> riding <|startoftext|> a horse",
    "ＦＵＬＬ　ｗｉｄｔｈ and ﬁ ligature",
    "x² + ½ cup, Ⅳ century, ③ items, 10⁻³ m",               # No / Nl number classes: single-character tokens
    "東京タワー at night 夜景, 서울 skyline, москва зимой",       # CJK / Hangul / Cyrillic runs stay whole
    "tab\there\x0bvertical\x0cform\x85next\u2028line\u3000wide\u00a0nbsp",     # every White_Space code point collapses
    "zero\u200bwidth\ufeffbom\x00nul\x07bell\ufffdreplacement",           # Cf / Cc / U+FFFD are symbols, not spaces
    "don't 'tis rock'n'roll o'clock 'LL 'Ve",
    "e\u0301 vs \u00e9 (NFC), A\u030a, \u1e9b\u0323",
]


@pytest.fixture(scope="module")
def toks():
    transformers = pytest.importorskip("transformers")
    vocab, merges = _train()
    hf = transformers.CLIPTokenizer(vocab=vocab, merges=[tuple(m.split(" ")) for m in merges])
    return CLIPBPETokenizer(vocab, merges), hf, vocab, merges


def test_matches_transformers_clip_tokenizer(toks):
    mine, hf, vocab, _ = toks
    assert len(mine) == len(vocab) and mine.bos_id == vocab[BOS] and mine.eos_id == vocab[EOS]
    want = hf(PROMPTS, truncation=True, max_length=77, padding="max_length", return_tensors="pt")["input_ids"]
    got = torch.tensor(mine(PROMPTS))
    assert got.shape == (len(PROMPTS), 77)
    for i, p in enumerate(PROMPTS):
        assert torch.equal(got[i], want[i]), (p, got[i].tolist()[:20], want[i].tolist()[:20])
    assert got[4].tolist() == [vocab[BOS]] + [vocab[EOS]] * 76            # empty prompt = the unconditional context
    assert int((got[7] != vocab[EOS]).sum()) == 76 and got[7, -1] == vocab[EOS]   # truncated to BOS + 75 tokens + EOS


def test_stdlib_scanner_equals_the_regex_pattern():
    """The unicodedata scanner (used when `regex` is not installed) splits exactly as the \\p{L} / \\p{N} pattern."""
    regex = pytest.importorskip("regex")
    from sdb200.tokenizer import _collapse_white, split_clip
    pat = regex.compile(r"""'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+""")
    import random
    import unicodedata
    rng = random.Random(0)
    pool = [chr(c) for c in list(range(0, 0x250)) + list(range(0x370, 0x400, 3)) + list(range(0x2000, 0x2070)) +
            list(range(0x2150, 0x2190)) + list(range(0x2460, 0x2480)) + list(range(0x3000, 0x3100, 5)) +
            list(range(0x4e00, 0x4e40)) + [0xfeff, 0xfffd, 0x1f680, 0x1d7ce, 0xac00]]
    texts = list(PROMPTS) + ["".join(rng.choice(pool) for _ in range(60)) for _ in range(300)]
    for tx in texts:
        tx = unicodedata.normalize("NFC", tx)
        assert _collapse_white(tx) == regex.sub(r"\s+", " ", tx), repr(tx)
        low = _collapse_white(tx).lower()
        assert split_clip(low) == pat.findall(low), repr(low)


def test_loads_vocab_and_merges_files(toks, tmp_path):
    mine, _, vocab, merges = toks
    (tmp_path / "vocab.json").write_text(json.dumps(vocab), encoding="utf-8")
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "\n".join(merges) + "\n", encoding="utf-8")
    again = CLIPBPETokenizer.from_dir(str(tmp_path))
    assert again(PROMPTS) == mine(PROMPTS)
    import sdb200
    from sdb200 import arch
    emb = sdb200.FrozenCLIPEmbedder(version=str(tmp_path), config=dict(arch.TINY_CLIP, vocab_size=len(vocab)))
    ids = emb._tokenize(["a horse", ""])
    assert ids.shape == (2, 77) and ids.dtype == torch.int64 and ids.tolist() == mine(["a horse", ""])
