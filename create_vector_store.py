"""Drop-in for the reference's ``create_vector_store.py`` (index build, create_vector_store.py:14-83):
``python create_vector_store.py`` reads ``./legal_chunks.json`` and writes ``./data/vector_store/``
(``index.faiss``, ``bm25.pkl``-equivalent ``bm25.npz``, ``metadata.json``, plus the fp16 matrix the
GPU engine maps).  Under ``torchrun --nproc-per-node G`` the embedding loop is data-parallel."""
import os

from legal_rag_engine_b200.engine import create_vector_store  # noqa: F401

if __name__ == "__main__":
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    create_vector_store()
