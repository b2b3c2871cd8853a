"""rust-local-rag_b200 -- B200-native retrieval hot path for rust-local-rag.

Scope: `RagEngine::search` / `search_with_diversity` / `get_embedding_candidates`
(exhaustive cosine scan -> top-k -> greedy MMR).  Layout:

  csrc/        sm_100a CUDA kernels + the C ABI (include/rlr_b200.h) -> librlr_b200.so
  binding.py   ctypes binding of the C ABI
  engine.py    host-side mirror of the reference's RagEngine for this path
  dist.py      one-process-per-GPU row-sharded search (torch.distributed / NCCL plumbing)
"""
from . import _build, binding  # noqa: F401

__all__ = ["_build", "binding", "engine", "dist"]
