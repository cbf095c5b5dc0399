"""`Batch` ABC and `GenerateError`.  Mirrors /root/reference/server/text_generation_server/models/types.py:15-62."""
from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from .. import pb as generate_pb2


@dataclass
class GenerateError:
    request_id: int
    message: str

    def to_pb(self):
        return generate_pb2.GenerateError(request_id=self.request_id, message=self.message)


class Batch(ABC):
    @abstractmethod
    def get_id(self) -> int:
        raise NotImplementedError

    @abstractmethod
    def __len__(self):
        raise NotImplementedError

    @classmethod
    @abstractmethod
    def from_pb(cls, pb, tokenizer, dtype: torch.dtype, device: torch.device, embeddings_lookup: Optional,
                prefix_cache: Optional, use_position_ids: bool = False) -> Tuple["Batch", List[GenerateError]]:
        raise NotImplementedError

    @classmethod
    @abstractmethod
    def concatenate(cls, batches: List["Batch"]) -> "Batch":
        raise NotImplementedError

    @classmethod
    @abstractmethod
    def prune(cls, batch: "Batch", completed_ids: List[int]) -> Optional["Batch"]:
        raise NotImplementedError

    def compact(self):
        pass
