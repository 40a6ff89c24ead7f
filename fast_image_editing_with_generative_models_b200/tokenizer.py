"""CLIP byte-level BPE tokenizer (the host half of the text-encoder row).

The reference gets it implicitly: ``StableDiffusionXLControlNetImg2ImgPipeline.from_pretrained`` loads ``tokenizer`` /
``tokenizer_2`` (reference ``src/pipeline.py:128-135,147-153``) and ``encode_prompt`` calls them with
``padding="max_length", max_length=77, truncation=True``.  The algorithm is `transformers`' ``CLIPTokenizer`` (third-party,
models/clip/tokenization_clip.py): NFC -> collapse whitespace -> lowercase; split with the CLIP pattern; map bytes to the
GPT-2 printable alphabet; BPE with an end-of-word suffix ``</w>``; ``<|startoftext|> ... <|endoftext|>``; pad to 77.
The two SDXL tokenizers differ only in the pad token (``<|endoftext|>`` for CLIP-L, ``!`` = id 0 for OpenCLIP bigG).

Pure host code (no CUDA, no oracle import).  ``vocab.json`` / ``merges.txt`` are not shipped (no network here): point
``CLIPBPETokenizer.from_files`` at a downloaded tokenizer folder.  tests/test_tokenizer_cpu.py pins it against
``transformers.CLIPTokenizer`` on a synthetic vocabulary learned in the test.
"""
from __future__ import annotations

import json
import os
import unicodedata
from functools import lru_cache
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import regex

BOS, EOS = "<|startoftext|>", "<|endoftext|>"
_SPLIT = regex.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+")
_WS = regex.compile(r"\s+")


@lru_cache(maxsize=1)
def byte_alphabet() -> Dict[int, str]:
    """byte -> printable unicode character (the GPT-2 table: printable Latin-1 bytes map to themselves, the other 68 to U+0100...)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    table, nxt = {}, 0
    for b in range(256):
        if b in keep:
            table[b] = chr(b)
        else:
            table[b] = chr(256 + nxt)
            nxt += 1
    return table


class CLIPBPETokenizer:
    def __init__(self, vocab: Dict[str, int], merges: Sequence[Tuple[str, str]], pad_token: str = EOS, max_length: int = 77):
        self.vocab = dict(vocab)
        self.ranks = {tuple(m): i for i, m in enumerate(merges)}
        self.bos_id, self.eos_id = self.vocab[BOS], self.vocab[EOS]
        self.unk_id = self.eos_id                                   # unk_token = <|endoftext|>
        self.pad_id = self.vocab[pad_token]
        self.max_length = max_length
        self._cache: Dict[str, List[int]] = {}

    @classmethod
    def from_files(cls, folder: str, pad_token: Optional[str] = None, max_length: int = 77) -> "CLIPBPETokenizer":
        """folder with vocab.json + merges.txt (a diffusers ``tokenizer`` / ``tokenizer_2`` sub-folder)."""
        with open(os.path.join(folder, "vocab.json"), encoding="utf-8") as f:
            vocab = json.load(f)
        with open(os.path.join(folder, "merges.txt"), encoding="utf-8") as f:
            lines = f.read().split("\n")
        merges = [tuple(ln.split()) for ln in lines if ln and not ln.startswith("#version") and len(ln.split()) == 2]
        if pad_token is None:
            pad_token = EOS
            spec = os.path.join(folder, "special_tokens_map.json")
            if os.path.isfile(spec):
                with open(spec, encoding="utf-8") as f:
                    pt = json.load(f).get("pad_token", EOS)
                pad_token = pt["content"] if isinstance(pt, dict) else pt
        return cls(vocab, merges, pad_token, max_length)

    # ---- BPE on one pre-token (already in the byte alphabet)
    def _bpe(self, word: str) -> List[int]:
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts = list(word[:-1]) + [word[-1] + "</w>"]
        while len(parts) > 1:
            best, best_rank = -1, None
            for i in range(len(parts) - 1):
                r = self.ranks.get((parts[i], parts[i + 1]))
                if r is not None and (best_rank is None or r < best_rank):
                    best, best_rank = i, r
            if best_rank is None:
                break
            a, b = parts[best], parts[best + 1]
            merged, i = [], 0
            while i < len(parts):                                   # merge every occurrence of the winning pair, left to right
                if i < len(parts) - 1 and parts[i] == a and parts[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        ids = [self.vocab.get(p, self.unk_id) for p in parts]
        self._cache[word] = ids
        return ids

    def tokenize_ids(self, text: str) -> List[int]:
        """ids of the text alone (no BOS/EOS/padding)."""
        text = _WS.sub(" ", unicodedata.normalize("NFC", text)).lower()
        table = byte_alphabet()
        out: List[int] = []
        for tok in _SPLIT.findall(text):
            if tok == BOS or tok == EOS:
                out.append(self.vocab[tok])
                continue
            out.extend(self._bpe("".join(table[b] for b in tok.encode("utf-8"))))
        return out

    def __call__(self, texts) -> List[List[int]]:
        """``tokenizer(texts, padding="max_length", max_length=77, truncation=True).input_ids``"""
        if isinstance(texts, str):
            texts = [texts]
        rows = []
        for t in texts:
            ids = self.tokenize_ids(t)[: self.max_length - 2]
            row = [self.bos_id] + ids + [self.eos_id]
            rows.append(row + [self.pad_id] * (self.max_length - len(row)))
        return rows
