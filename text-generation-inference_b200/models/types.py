"""The batch contract between the shard server and a model (reference: models/types.py:15-62).

`TextGenerationService` (server.py) keeps the objects returned by `from_pb` in its cache between calls and only ever touches
them through this interface; `Model.generate_token` mutates them in place.
"""
import abc
import dataclasses
from typing import List, Optional, Tuple

import torch

from ..utils.token_types import _Record


@dataclasses.dataclass
class GenerateError(_Record):
    """A per-request failure that does not fail the whole batch (generate.proto:150-153)."""
    PB = "GenerateError"
    request_id: int
    message: str


class Batch(abc.ABC):
    """What the server needs from a batch type; `FlashCausalLMBatch` is the one implementation on this path."""

    @abc.abstractmethod
    def get_id(self) -> int:
        """The router's batch id (cache key, cache.py:15-17)."""

    @abc.abstractmethod
    def __len__(self):
        """Number of live requests."""

    @classmethod
    @abc.abstractmethod
    def from_pb(cls, pb, tokenizer, dtype: torch.dtype, device: torch.device, embeddings_lookup: Optional, prefix_cache: Optional,
                use_position_ids: bool = False) -> Tuple["Batch", List[GenerateError]]:
        """Tokenize a `generate.v1.Batch`; requests that fail validation come back as errors, not exceptions."""

    @classmethod
    @abc.abstractmethod
    def concatenate(cls, batches: List["Batch"]) -> "Batch":
        """Merge cached batches after an add-on prefill (continuous batching); inputs must not be used afterwards."""

    @classmethod
    @abc.abstractmethod
    def prune(cls, batch: "Batch", completed_ids: List[int]) -> Optional["Batch"]:
        """Drop finished requests; None when nothing is left, the same object when nothing finished."""

    def compact(self):
        """Release slack memory before a new prefill (COMPACT_BEFORE_PREFILL, server.py:34); nothing to do for paged KV."""
