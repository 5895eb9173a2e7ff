"""Offline stand-in for the Qwen2 tokenizer (no tokenizer files or network are available here).

Implements exactly the surface the reference agent uses (streamvln/streamvln_agent.py:44-46,109-167,
248): `add_tokens`, `convert_tokens_to_ids`, `additional_special_tokens_ids`, `__call__(text).input_ids`,
`apply_chat_template(messages)` with the agent's overridden template
`'<|im_start|>' + role + '\\n' + content + '<|im_end|>' + '\\n'`, and `batch_decode`.  Words map to ids by
a stable hash, so prompts are deterministic fixed token-id arrays.  A real `transformers` tokenizer can
be passed to the agent instead; nothing else depends on this class.
"""
from __future__ import annotations

import re
import zlib
from types import SimpleNamespace
from typing import List

ACTION_GLYPHS = ["STOP", "↑", "←", "→"]


class StubQwenTokenizer:
    def __init__(self, vocab_size: int = 152064):
        self.vocab_size = vocab_size
        if vocab_size >= 151646:
            self.im_start, self.im_end, self.nl = 151644, 151645, 198   # Qwen2 ids
            self._next_added = 151646
        else:
            self.im_start, self.im_end, self.nl = vocab_size - 4, vocab_size - 3, vocab_size - 5
            self._next_added = vocab_size - 2
        self.n_text = min(self.im_start, self.nl) if vocab_size < 151646 else 151643
        self.added = {}
        self.chat_template = None
        self.eos_token_id = self.im_end

    # -- surface used by the agent
    @property
    def additional_special_tokens_ids(self) -> List[int]:
        return [self.im_start, self.im_end]

    def add_tokens(self, tokens, special_tokens: bool = False) -> int:
        n = 0
        for t in tokens:
            if t not in self.added:
                if self._next_added >= self.vocab_size:
                    raise ValueError("vocabulary full")
                self.added[t] = self._next_added
                self._next_added += 1
                n += 1
        return n

    def convert_tokens_to_ids(self, token: str) -> int:
        if token in self.added:
            return self.added[token]
        return {"<|im_start|>": self.im_start, "<|im_end|>": self.im_end}.get(token, self._word(token))

    def _word(self, w: str) -> int:
        return zlib.crc32(w.encode("utf-8")) % self.n_text

    def encode(self, text: str) -> List[int]:
        specials = ["<|im_start|>", "<|im_end|>"] + list(self.added)
        pat = "(" + "|".join(re.escape(s) for s in specials) + r"|\n)"
        ids: List[int] = []
        for part in re.split(pat, text):
            if not part:
                continue
            if part == "\n":
                ids.append(self.nl)
            elif part in specials:
                ids.append(self.convert_tokens_to_ids(part))
            else:
                ids.extend(self._word(w) for w in part.split())
        return ids

    def __call__(self, text: str):
        return SimpleNamespace(input_ids=self.encode(text))

    def apply_chat_template(self, messages) -> List[int]:
        out: List[int] = []
        for m in messages:
            out += self.encode("<|im_start|>" + m["role"] + "\n" + m["content"] + "<|im_end|>" + "\n")
        return out

    def decode_id(self, i: int) -> str:
        if i == self.im_start:
            return "<|im_start|>"
        if i == self.im_end:
            return "<|im_end|>"
        if i == self.nl:
            return "\n"
        for k, v in self.added.items():
            if v == i:
                return k
        return ACTION_GLYPHS[i % 4]  # every ordinary id reads as an action glyph: random weights still navigate

    def batch_decode(self, ids, skip_special_tokens: bool = False) -> List[str]:
        rows = ids.tolist() if hasattr(ids, "tolist") else ids
        return ["".join(self.decode_id(int(i)) for i in row) for row in rows]
