"""Token result records returned by `Model.generate_token`.
Mirrors /root/reference/server/text_generation_server/utils/token_types.py:8-56."""
from dataclasses import dataclass
from functools import total_ordering
from typing import List, Optional

from .. import pb as generate_pb2


@dataclass(eq=True)
@total_ordering
class TopToken:
    token_id: int
    logprob: float = 0.0

    def __gt__(self, other):
        # equal logprobs tie-break on the lower token id, like torch.argmax (token_types.py:14-18)
        return self.logprob > other.logprob or (self.logprob == other.logprob and self.token_id < other.token_id)

    def to_pb(self):
        return generate_pb2.TopToken(token_id=self.token_id, logprob=self.logprob)


@dataclass
class TokenInfo:
    token_id: int
    request_id: int = 0
    logprob: float = 0.0
    rank: int = 0
    top_tokens: Optional[List[TopToken]] = None

    def to_pb(self):
        return generate_pb2.Token(request_id=self.request_id, token_id=self.token_id, logprob=self.logprob, rank=self.rank,
                                  top_tokens=None if self.top_tokens is None else [tt.to_pb() for tt in self.top_tokens])


@dataclass
class InputTokens:
    request_id: int
    tokens: List[TokenInfo]

    def to_pb(self):
        return generate_pb2.InputTokens(request_id=self.request_id, tokens=[t.to_pb() for t in self.tokens])
