"""`Model` ABC: what the shard server drives.

Mirrors /root/reference/server/text_generation_server/models/model.py:34-221 for the attributes the server reads
(`batch_type`, `tokenizer`, `dtype`, `device`, `word_embeddings`, `prefix_cache`, `use_position_ids`,
`context_manager`, `engine`, `config`, `model`; server.py:82,106,131-139,326-332) and `generate_token` /
`get_indices_to_keep` (:169-188).  PT2 compile wrappers (:97-157) and the prompt-prefix cache (:47-90) are out of
scope: `prefix_cache` is None and requests with a prefix_id get a GenerateError.
"""
from __future__ import annotations

import inspect
from abc import ABC, abstractmethod
from typing import List, Optional, Tuple, Type, TypeVar

import torch

from .types import Batch, GenerateError
from ..utils.token_types import InputTokens, TokenInfo

B = TypeVar("B", bound=Batch)


class Model(ABC):
    def __init__(self, engine, dtype: torch.dtype, max_seq_length: Optional[int] = None):
        self.engine = engine
        self.config, self.tokenizer, self.model = engine.get_components()
        self.device = engine.get_device()
        self.dtype = dtype
        if getattr(self.config, "eos_token_id", None) is not None:
            self.tokenizer.model_eos_token_id = self.config.eos_token_id
        self.use_position_ids = "position_ids" in inspect.signature(self.model.forward).parameters
        try:
            self.word_embeddings = self.model.get_input_embeddings()
        except Exception:
            self.word_embeddings = None
        self.prefix_cache = None
        self.context_manager = torch.inference_mode
        self.compiled = False

    @property
    @abstractmethod
    def batch_type(self) -> Type[B]:
        raise NotImplementedError

    @abstractmethod
    def generate_token(self, batch: B, first: bool = False, for_concat: bool = False
                       ) -> Tuple[List[TokenInfo], Optional[List[InputTokens]], List[GenerateError], int]:
        raise NotImplementedError

    @staticmethod
    def get_indices_to_keep(requests, completed_ids: List[int]) -> List[int]:
        """model.py:176-188: both lists ascend by request id (router/src/queue.rs:423-424)."""
        keep = []
        completed = iter(completed_ids)
        next_id = next(completed)
        for i, r in enumerate(requests):
            while next_id is not None and r.id > next_id:
                next_id = next(completed, None)
            if r.id != next_id:
                keep.append(i)
        return keep
