"""CLIP BPE tokenizer (sd/pipeline.py:109, sd/inference_demo.ipynb:47) on a synthetic
vocab.json / merges.txt pair: known answers worked out by hand, the batch_encode_plus duck type the reference calls,
and - when transformers is importable - agreement with HuggingFace's CLIPTokenizer built from the same two files."""
import json

import pytest

from pytorch_stable_diffusion_b200.tokenizer import BOS, EOS, CLIPTokenizer, bytes_to_unicode

MERGES = ["h e", "l l", "he ll", "hell o</w>", "w o", "r l", "wo rl", "worl d</w>", "c a", "ca t</w>", "' s</w>",
          "! !</w>"]


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("clipvocab")
    b2u = bytes_to_unicode()
    chars = [b2u[b] for b in range(256)]
    vocab = chars + [c + "</w>" for c in chars]
    for m in MERGES:
        a, b = m.split()
        vocab.append(a + b)
    vocab += [BOS, EOS]
    enc = {t: i for i, t in enumerate(vocab)}
    assert len(enc) == len(vocab)
    (d / "vocab.json").write_text(json.dumps(enc), encoding="utf-8")
    (d / "merges.txt").write_text("#version: 0.2\n" + "\n".join(MERGES) + "\n", encoding="utf-8")
    return str(d / "vocab.json"), str(d / "merges.txt"), enc


def test_bytes_to_unicode_is_a_bijection():
    t = bytes_to_unicode()
    assert len(t) == 256 and len(set(t.values())) == 256
    assert t[ord("a")] == "a" and t[ord(" ")] == "Ġ"


def test_known_answers(files):
    vocab, merges, enc = files
    tok = CLIPTokenizer(vocab, merges)
    assert tok.tokenize("Hello   WORLD") == ["hello</w>", "world</w>"]
    assert tok.tokenize("cat's") == ["cat</w>", "'s</w>"]
    assert tok.tokenize("cats") == ["ca", "t", "s</w>"]                 # "ca t</w>" does not apply inside a word
    assert tok.tokenize("hello!!") == ["hello</w>", "!!</w>"]
    assert tok.tokenize("a1b") == ["a</w>", "1</w>", "b</w>"]           # digits are split one by one
    b2u = bytes_to_unicode()
    e0, e1 = (b2u[b] for b in "é".encode("utf-8"))                       # two UTF-8 bytes, the second carries </w>
    assert tok.tokenize("é") == [e0, e1 + "</w>"]
    ids = tok.encode("hello world")
    assert ids == [enc[BOS], enc["hello</w>"], enc["world</w>"], enc[EOS]]
    assert tok.decode(ids) == "hello world"


def test_batch_encode_plus_is_what_the_pipeline_calls(files):
    vocab, merges, enc = files
    tok = CLIPTokenizer(vocab, merges)
    out = tok.batch_encode_plus(["hello world", ""], padding="max_length", max_length=77)
    assert [len(r) for r in out.input_ids] == [77, 77]
    assert out.input_ids[0][:4] == [enc[BOS], enc["hello</w>"], enc["world</w>"], enc[EOS]]
    assert set(out.input_ids[0][4:]) == {enc[EOS]}                      # padded with <|endoftext|>
    assert out.input_ids[1][:2] == [enc[BOS], enc[EOS]]
    assert out.attention_mask[0] == [1] * 4 + [0] * 73
    long = tok.batch_encode_plus(["hello " * 100], padding="max_length", max_length=77, truncation=True).input_ids[0]
    assert len(long) == 77 and long[-1] == enc[EOS]
    # through pipeline._encode_prompts' access pattern
    ids = tok.batch_encode_plus(["cat"], padding="max_length", max_length=77).input_ids[0]
    assert ids[1] == enc["cat</w>"]


def test_matches_huggingface_clip_tokenizer(files):
    vocab, merges, enc = files
    try:
        import inspect
        from transformers import CLIPTokenizer as HF
        if "vocab_file" in inspect.signature(HF.__init__).parameters:   # transformers 4.x (the reference pins 4.51.3)
            hf = HF(vocab_file=vocab, merges_file=merges)
        else:                                                            # transformers 5.x: in-memory vocab / merges
            hf = HF(vocab=dict(enc), merges=[tuple(m.split()) for m in MERGES])
    except Exception as e:                                               # noqa: BLE001
        pytest.skip(f"transformers CLIPTokenizer unavailable for local files: {e}")
    mine = CLIPTokenizer(vocab, merges)
    texts = ["hello world", "Hello, WORLD!!", "cat's cats 11 cat", "  many   spaces\there ", "héllo é", "a-b_c.d", ""]
    for t in texts:
        try:
            ref = hf(t, padding="max_length", max_length=77)["input_ids"]
        except Exception as e:                                           # noqa: BLE001
            pytest.skip(f"transformers CLIPTokenizer cannot encode here: {e}")
        got = mine.batch_encode_plus([t], padding="max_length", max_length=77).input_ids[0]
        assert got == list(ref), t
