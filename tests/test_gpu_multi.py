"""Two real GPUs, one process each over NCCL: the sharded search (ShardedSearcher) must equal
the unsharded oracle.  Skipped on single-GPU boxes (run with `gpurun --gpus 2`)."""
import os
import socket
import time

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, B, K_TOP, VOCAB = 60000, 4, 10, 3000
WEIGHTS = [0.5, 0.6, 0.5, 0.6]


def _inputs():
    from legal_rag_engine_b200 import synth
    x = synth.host_vectors(N, seed=71, dup_frac=0.01)
    idx = synth.host_bm25(N, seed=72, vocab=VOCAB)
    q = synth.host_queries(B, seed=73)
    terms, ptr = synth.host_query_terms(B, 8, seed=74, vocab=VOCAB)
    return x, idx, q, terms, ptr


def _worker(rank, world, port, mode, exchange, ret):
    import torch.distributed as dist
    from legal_rag_engine_b200 import sharding
    from legal_rag_engine_b200.device_index import DeviceIndex, FUSION
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x, idx, q, terms, ptr = _inputs()
        lo, hi = sharding.shard_range(N, rank, world)
        sh = idx.shard(lo, hi)
        dev = DeviceIndex(rank, rank, world)
        dev.set_corpus(torch.from_numpy(x[lo:hi]).cuda(), lo)
        dev.set_postings(sh.term_ptr, sh.postings, sh.doc_len, sh.idf, sh.avgdl)
        s = sharding.ShardedSearcher(dev, exchange=exchange)
        c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        # several batches back to back: the exchange slots and flags are reused every second call;
        # a different (reversed) batch in between must not leak into the last one
        dq, dt, dp, dw = c(q), c(terms), c(ptr), c(np.array(WEIGHTS))
        dq2 = c(q[::-1])
        for i in range(5):
            s.search(dq2 if i % 2 else dq, dt, dp, K_TOP, FUSION[mode], dw)
        torch.cuda.synchronize()
        if rank == 1:
            time.sleep(0.05)                                  # ranks drift apart: flags must hold
        outs = s.search(dq, dt, dp, K_TOP, FUSION[mode], dw)
        torch.cuda.synchronize()
        ret[rank] = [t.cpu().numpy() for t in outs]
        dist.barrier()                                        # nobody unmaps while a peer may store
        dev.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("mode", ["linear", "rrf"])
def test_two_gpu_sharded_search_equals_oracle(mode, exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import bm25 as obm25
    from oracle.search import OracleIndex
    x, idx, q, terms, ptr = _inputs()
    csr = obm25.BM25OkapiCSR.from_postings(N, idx.doc_len, idx.term_ptr.astype(np.int64),
                                           idx.postings[:, 0], idx.postings[:, 1])
    lists = [terms[ptr[b]:ptr[b + 1]].tolist() for b in range(B)]
    want = OracleIndex(x, csr).search_batch_vec(q, lists, K_TOP, WEIGHTS, mode)
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, mode, exchange, ret), nprocs=2, join=True)
        r0, r1 = ret[0], ret[1]
    for a, b in zip(r0, r1):
        np.testing.assert_array_equal(a, b)                   # replicated result
    ids, score, sem, kw, status = r0
    assert status.sum() == 0
    for b in range(B):
        assert ids[b][:len(want[b])].tolist() == [r[0] for r in want[b]]
        assert score[b][:len(want[b])].tolist() == [r[1] for r in want[b]]
