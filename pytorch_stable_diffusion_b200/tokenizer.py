"""Self-contained CLIP byte-level BPE tokenizer with the duck type pipeline.generate expects.

The reference passes a HuggingFace `CLIPTokenizer(vocab.json, merges_file=merges.txt)` (sd/inference_demo.ipynb:47)
and calls exactly one method on it (sd/pipeline.py:109,115,127):

    tokenizer.batch_encode_plus([prompt], padding="max_length", max_length=77).input_ids   # list[list[int]]

This class implements that call from the two vocabulary files alone (no transformers / tokenizers dependency), so
real prompts work wherever the SD-1.5 `vocab.json` / `merges.txt` are available. Algorithm (OpenAI CLIP's
SimpleTokenizer as HuggingFace runs it without ftfy): NFC-normalise, collapse whitespace, lower-case; split with the
CLIP pattern; map every UTF-8 byte of a word to a printable unicode character; append "</w>" to the last one; merge
adjacent symbol pairs in the order of merges.txt until none applies; look the symbols up in vocab.json; wrap in
<|startoftext|> ... <|endoftext|>; pad to max_length with the pad token (<|endoftext|>, as CLIPTokenizer's default).
"""
import functools
import json
import unicodedata

try:                                   # \p{L} / \p{N} classes need the third-party `regex` module
    import regex as _re
    _PATTERN = _re.compile(
        r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+", _re.IGNORECASE)
except ImportError:                    # pragma: no cover - ASCII-exact, approximate outside it
    import re as _re
    _PATTERN = _re.compile(
        r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[^\W\d_]+|\d|(?:[^\s\w]|_)+", _re.IGNORECASE)

BOS, EOS = "<|startoftext|>", "<|endoftext|>"


@functools.lru_cache()
def bytes_to_unicode():
    """The GPT-2 / CLIP byte -> printable unicode character table (256 entries, reversible)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("\xa1"), ord("\xac") + 1)) + \
        list(range(ord("\xae"), ord("\xff") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, (chr(c) for c in cs)))


class BatchEncoding(dict):
    """dict with attribute access: `.input_ids`, `.attention_mask` (what the reference reads is `.input_ids`)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e


class CLIPTokenizer:
    """CLIPTokenizer(vocab_file, merges_file) - same constructor arguments as the class the reference uses."""

    def __init__(self, vocab_file, merges_file, unk_token=EOS, bos_token=BOS, eos_token=EOS, pad_token=EOS):
        with open(vocab_file, encoding="utf-8") as f:
            self.encoder = json.load(f)
        self.decoder = {v: k for k, v in self.encoder.items()}
        with open(merges_file, encoding="utf-8") as f:
            lines = f.read().strip().split("\n")
        if lines and lines[0].startswith("#version"):
            lines = lines[1:]
        merges = [tuple(ln.split()) for ln in lines if ln.strip()]
        self.bpe_ranks = {m: i for i, m in enumerate(merges) if len(m) == 2}
        self.byte_encoder = bytes_to_unicode()
        self.byte_decoder = {v: k for k, v in self.byte_encoder.items()}
        self.unk_token, self.bos_token, self.eos_token, self.pad_token = unk_token, bos_token, eos_token, pad_token
        for t in (bos_token, eos_token, pad_token):
            if t not in self.encoder:
                raise ValueError(f"special token {t!r} is not in {vocab_file}")
        self.bos_token_id, self.eos_token_id = self.encoder[bos_token], self.encoder[eos_token]
        self.pad_token_id = self.encoder[pad_token]
        self.unk_token_id = self.encoder.get(unk_token, self.eos_token_id)
        self.model_max_length = 77
        self._cache = {BOS: BOS, EOS: EOS}

    @property
    def vocab_size(self):
        return len(self.encoder)

    # ---- BPE ------------------------------------------------------------------------------------
    def bpe(self, token):
        """Space-separated merged symbols of one pre-token (already byte-mapped)."""
        if token in self._cache:
            return self._cache[token]
        word = tuple(token[:-1]) + (token[-1] + "</w>",)
        while len(word) > 1:
            pairs = set(zip(word, word[1:]))
            best = min(pairs, key=lambda p: self.bpe_ranks.get(p, float("inf")))
            if best not in self.bpe_ranks:
                break
            first, second = best
            merged, i = [], 0
            while i < len(word):
                if i + 1 < len(word) and word[i] == first and word[i + 1] == second:
                    merged.append(first + second)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = tuple(merged)
        out = " ".join(word)
        self._cache[token] = out
        return out

    @staticmethod
    def _clean(text):
        text = unicodedata.normalize("NFC", text)
        return " ".join(text.split()).strip().lower()

    def tokenize(self, text):
        symbols = []
        for tok in _PATTERN.findall(self._clean(text)):
            if tok in (BOS, EOS):
                symbols.append(tok)
                continue
            mapped = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            symbols.extend(self.bpe(mapped).split(" "))
        return symbols

    def convert_tokens_to_ids(self, tokens):
        return [self.encoder.get(t, self.unk_token_id) for t in tokens]

    def encode(self, text, add_special_tokens=True):
        ids = self.convert_tokens_to_ids(self.tokenize(text))
        return [self.bos_token_id] + ids + [self.eos_token_id] if add_special_tokens else ids

    def decode(self, ids, skip_special_tokens=True):
        special = {self.bos_token_id, self.eos_token_id, self.pad_token_id}
        text = "".join(self.decoder[i] for i in ids if not (skip_special_tokens and i in special))
        words = [bytearray(self.byte_decoder[c] for c in w if c in self.byte_decoder).decode("utf-8", errors="replace")
                 for w in text.split("</w>")]
        return " ".join(words).strip()

    # ---- the call the reference makes (sd/pipeline.py:109) -------------------------------------------
    def batch_encode_plus(self, batch_text, padding=False, max_length=None, truncation=False, **_):
        if isinstance(batch_text, str):
            batch_text = [batch_text]
        if max_length is None:
            max_length = self.model_max_length
        input_ids, masks = [], []
        for text in batch_text:
            ids = self.encode(text)
            if truncation and len(ids) > max_length:
                ids = ids[:max_length - 1] + [self.eos_token_id]
            mask = [1] * len(ids)
            if padding == "max_length" and len(ids) < max_length:
                pad = max_length - len(ids)
                ids = ids + [self.pad_token_id] * pad
                mask = mask + [0] * pad
            input_ids.append(ids)
            masks.append(mask)
        if padding in (True, "longest"):
            width = max(len(i) for i in input_ids)
            masks = [m + [0] * (width - len(m)) for m in masks]
            input_ids = [i + [self.pad_token_id] * (width - len(i)) for i in input_ids]
        return BatchEncoding(input_ids=input_ids, attention_mask=masks)

    __call__ = batch_encode_plus
