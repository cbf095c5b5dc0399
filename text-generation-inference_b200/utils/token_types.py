"""Per-token results of `Model.generate_token` and their wire form.

The same three records as /root/reference/server/text_generation_server/utils/token_types.py:8-56 (the server and the
chooser construct them by field name): `TopToken` (a candidate with its log-probability, ordered best first),
`TokenInfo` (one generated or input token) and `InputTokens` (the echoed prompt of a request).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

from .. import pb


class _Record:
    """`to_pb()` for a dataclass whose field names are the message's field names."""
    PB: str = ""

    def to_pb(self):
        values = {}
        for name in self.__dataclass_fields__:
            v = getattr(self, name)
            values[name] = [item.to_pb() for item in v] if isinstance(v, list) else v
        return getattr(pb, self.PB)(**values)


@dataclasses.dataclass
class TopToken(_Record):
    PB = "TopToken"
    token_id: int
    logprob: float = 0.0

    def _rank_key(self):
        # higher log-probability first; equal log-probabilities resolve to the LOWER token id, the way torch.argmax
        # resolves ties in greedy decoding (token_types.py:14-18)
        return (self.logprob, -self.token_id)

    def __lt__(self, other):
        return self._rank_key() < other._rank_key()

    def __le__(self, other):
        return self._rank_key() <= other._rank_key()

    def __gt__(self, other):
        return self._rank_key() > other._rank_key()

    def __ge__(self, other):
        return self._rank_key() >= other._rank_key()


@dataclasses.dataclass
class TokenInfo(_Record):
    PB = "Token"
    token_id: int
    request_id: int = 0  # not meaningful for input tokens
    logprob: float = 0.0
    rank: int = 0
    top_tokens: Optional[List[TopToken]] = None


@dataclasses.dataclass
class InputTokens(_Record):
    PB = "InputTokens"
    request_id: int
    tokens: List[TokenInfo]
