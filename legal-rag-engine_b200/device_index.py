"""One GPU shard of the hybrid index: torch owns the device memory, the C ABI
(include/lrx.h) runs the kernels.  Thin, typed wrappers; no arithmetic here."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import LRX_DIM, LRX_FUSE_LINEAR, LRX_FUSE_RRF, RECORD_BYTES

FUSION = {"linear": LRX_FUSE_LINEAR, "rrf": LRX_FUSE_RRF}


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceIndex:
    def __init__(self, device: int = 0, rank: int = 0, world: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("DeviceIndex needs a CUDA device: this engine has no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.rank, self.world = rank, world
        cfg = _lib.lrx_config(device, LRX_DIM, rank, world)
        h = C.c_void_p()
        _lib.check(self.lib.lrx_open(C.byref(cfg), C.byref(h)))
        self.h = h
        self.x = None
        self._q8 = None
        self.n_local = 0
        self.id_base = 0
        self._post = None
        self._capacity = 0
        self._pending = None
        self.use_current_stream()

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.lrx_close(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(rc, self.h)

    def use_current_stream(self):
        """Bind the library to torch's current stream on this device."""
        s = torch.cuda.current_stream(self.device).cuda_stream
        self._ck(self.lib.lrx_set_stream(self.h, C.c_void_p(s)))

    # ------------------------------------------------------------- residency
    def set_corpus(self, x: torch.Tensor, id_base: int = 0, prefilter: Optional[bool] = None):
        """Make the fp16 chunk matrix resident.  `prefilter` (default: on, LRX_DENSE_PREFILTER=0
        turns it off): also build the int8 shadow K2a scans for up to 4 queries at a time
        (388 B/row instead of 768; results unchanged, see include/lrx.h)."""
        assert x.dtype == torch.float16 and x.is_cuda and x.is_contiguous()
        assert x.dim() == 2 and x.shape[1] == LRX_DIM
        self.x, self.n_local, self.id_base = x, int(x.shape[0]), int(id_base)
        self._ck(self.lib.lrx_set_corpus(self.h, _ptr(x), self.n_local, self.id_base, LRX_DIM))
        self._q8 = None
        if prefilter is None:
            prefilter = os.environ.get("LRX_DENSE_PREFILTER", "1") not in ("0", "", "off")
        if prefilter and self.n_local > 0:
            self.build_prefilter()

    def build_prefilter(self):
        nbytes = int(self.lib.lrx_dense_prefilter_bytes(self.n_local))
        buf = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        bounds = (C.c_double * 2)()
        self._ck(self.lib.lrx_build_dense_prefilter(self.h, _ptr(buf), nbytes, bounds))
        self._q8 = (buf, nbytes, (float(bounds[0]), float(bounds[1])))

    def set_prefilter(self, on: bool):
        """Switch the int8 pre-filter of an index that has a shadow (A/B runs, tests)."""
        if on:
            if self._q8 is None:
                self.build_prefilter()
                return
            buf, nbytes, b = self._q8
            self._ck(self.lib.lrx_set_dense_prefilter(self.h, _ptr(buf), nbytes, (C.c_double * 2)(*b)))
        else:
            self._ck(self.lib.lrx_set_dense_prefilter(self.h, C.c_void_p(0), 0, None))

    @property
    def prefilter_bounds(self):
        """(max row error norm, max row norm) of the int8 shadow, or None."""
        return self._q8[2] if self._q8 is not None else None

    def set_postings(self, term_ptr, postings, doc_len, idf, avgdl: float, k1: float = 1.5,
                     b: float = 0.75):
        """Make the BM25 postings resident.  term_ptr int64/uint64 [V+1]; postings
        int32/uint32 [nnz,2] = (local doc, tf); doc_len int32/uint32 [n]; idf float64 [V]
        (GLOBAL statistics) -- torch CUDA tensors or numpy arrays (uploaded).  The 8-byte
        scoring layout {doc, u16 tf, u16 len} is built on the device (index-build step)."""
        def dev(a, dt):
            if isinstance(a, np.ndarray):
                if a.dtype == np.uint64:
                    a = a.view(np.int64)
                elif a.dtype == np.uint32:
                    a = a.view(np.int32)
                a = torch.from_numpy(np.ascontiguousarray(a))
            return a.to(self.device, dtype=dt).contiguous()
        tp = dev(term_ptr, torch.int64)
        po = dev(postings, torch.int32)
        dl = dev(doc_len, torch.int32)
        idf_t = dev(idf, torch.float64)
        nnz = int(tp[-1].item())
        assert po.shape[0] == nnz, (po.shape, nnz)
        if dl.numel() == 0:
            dl = torch.zeros(1, dtype=torch.int32, device=self.device)
        max_len = int(dl.max().item())
        p8 = torch.zeros((nnz + 2, 2), dtype=torch.int32, device=self.device)   # 8 B / posting, padded
        self._ck(self.lib.lrx_bm25_build_postings(self.h, _ptr(po), nnz, _ptr(dl), _ptr(p8)))
        del po
        self._post = (tp, p8, idf_t)
        self._post_args = (nnz, float(avgdl), float(k1), float(b), max_len)
        self.doc_len = dl
        self.n_terms = int(tp.numel() - 1)
        self._ck(self.lib.lrx_set_postings(self.h, _ptr(tp), _ptr(p8), _ptr(idf_t), self.n_terms, nnz,
                                           float(avgdl), float(k1), float(b), max_len))

    def clone_view(self) -> "DeviceIndex":
        """A second handle (own stream binding, workspaces, exchange region) over the SAME resident
        matrix and postings -- no copies: lets two query batches be in flight on two streams, one
        batch's merges and fusion running under the other's scans."""
        other = DeviceIndex(self.device.index, self.rank, self.world)
        if self.x is not None:
            other.set_corpus(self.x, self.id_base, prefilter=False)
            if self._q8 is not None:                        # the same shadow, no second copy
                other._q8 = self._q8
                other.set_prefilter(True)
        if self._post is not None:
            tp, p8, idf_t = self._post
            nnz, avgdl, k1, b, max_len = self._post_args
            other._post, other._post_args = self._post, self._post_args
            other.doc_len, other.n_terms = self.doc_len, self.n_terms
            other._ck(self.lib.lrx_set_postings(other.h, _ptr(tp), _ptr(p8), _ptr(idf_t), self.n_terms,
                                                nnz, avgdl, k1, b, max_len))
        return other

    def _cap(self, q_terms: Optional[torch.Tensor], B: int):
        """Device-pointer calls: the library sizes its per-token tables from a capacity it cannot
        read off the device; the tensor's length is an upper bound of the batch's token count."""
        n = int(q_terms.numel()) if q_terms is not None else 0
        want = n if n > B * _lib.LRX_MAX_QUERY_TERMS else 0          # 0 = the default B * 64
        if want != self._capacity:
            self._ck(self.lib.lrx_set_query_capacity(self.h, want))
            self._capacity = want

    def set_exchange_timeout(self, milliseconds: int):
        self._ck(self.lib.lrx_set_exchange_timeout(self.h, int(milliseconds)))

    # ---------------------------------------------------------------- stages
    def dense_topk(self, q: torch.Tensor, K: int, width: int = 0):
        assert q.dtype == torch.float16 and q.is_cuda and q.is_contiguous()
        B = int(q.shape[0])
        exact = torch.empty((B, K), dtype=torch.float64, device=self.device)
        D = torch.empty((B, K), dtype=torch.float32, device=self.device)
        I = torch.empty((B, K), dtype=torch.int64, device=self.device)
        flags = torch.empty(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.lrx_dense_topk_ex(self.h, _ptr(q), B, K, width, _ptr(exact), _ptr(D),
                                            _ptr(I), _ptr(flags)))
        return exact, D, I, flags

    def dense_topk_batched(self, q: torch.Tensor, K: int, stride: int = 0):
        """K2b: tensor-core path for large batches (B up to 4096).  Same outputs as dense_topk."""
        assert q.dtype == torch.float16 and q.is_cuda and q.is_contiguous()
        B = int(q.shape[0])
        exact = torch.empty((B, K), dtype=torch.float64, device=self.device)
        D = torch.empty((B, K), dtype=torch.float32, device=self.device)
        I = torch.empty((B, K), dtype=torch.int64, device=self.device)
        flags = torch.empty(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.lrx_dense_topk_batched(self.h, _ptr(q), B, K, stride, _ptr(exact), _ptr(D),
                                                 _ptr(I), _ptr(flags)))
        return exact, D, I, flags

    def dense_at(self, q: torch.Tensor, ids: torch.Tensor):
        B, n = int(ids.shape[0]), int(ids.shape[1])
        out = torch.empty((B, n), dtype=torch.float64, device=self.device)
        self._ck(self.lib.lrx_dense_at(self.h, _ptr(q), B, _ptr(ids.contiguous()), n, _ptr(out)))
        return out

    def bm25(self, q_terms: torch.Tensor, q_ptr: torch.Tensor, cand_ids: Optional[torch.Tensor] = None,
             K: int = 0):
        """Returns (cand_scores [B,n] or None, max [B], top_scores [B,K], top_ids [B,K])."""
        B = int(q_ptr.numel() - 1)
        n = int(cand_ids.shape[1]) if cand_ids is not None else 0
        cs = torch.empty((B, n), dtype=torch.float64, device=self.device) if n else None
        mx = torch.empty(B, dtype=torch.float64, device=self.device)
        ts = torch.empty((B, max(K, 1)), dtype=torch.float64, device=self.device)
        ti = torch.empty((B, max(K, 1)), dtype=torch.int64, device=self.device)
        if q_terms.numel() == 0:
            q_terms = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._cap(q_terms, B)
        self._ck(self.lib.lrx_bm25(self.h, _ptr(q_terms), _ptr(q_ptr), B,
                                   _ptr(cand_ids.contiguous() if n else None), n, _ptr(cs), _ptr(mx),
                                   K, _ptr(ts), _ptr(ti)))
        return cs, mx, ts[:, :K], ti[:, :K]

    def search_local(self, q, q_terms, q_ptr, k: int, mode: int, width: int = 0):
        """K2 + K3 on this shard -> (records uint8 [B,2,2k,24], maxbm25 [B], flags [B])."""
        B = int(q.shape[0])
        K = 2 * k
        rec = torch.empty((B, 2, K, RECORD_BYTES), dtype=torch.uint8, device=self.device)
        mx = torch.empty(B, dtype=torch.float64, device=self.device)
        flags = torch.empty(B, dtype=torch.int32, device=self.device)
        if q_terms.numel() == 0:
            q_terms = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._cap(q_terms, B)
        self._ck(self.lib.lrx_search_local(self.h, _ptr(q), _ptr(q_terms), _ptr(q_ptr), B, k, mode,
                                           width, _ptr(rec), _ptr(mx), _ptr(flags)))
        return rec, mx, flags

    def search_finish(self, rec_all, max_all, flags_all, world: int, B: int, k: int, mode: int,
                      weights: torch.Tensor):
        ids = torch.empty((B, k), dtype=torch.int64, device=self.device)
        score = torch.empty((B, k), dtype=torch.float64, device=self.device)
        sem = torch.empty((B, k), dtype=torch.float64, device=self.device)
        kw = torch.empty((B, k), dtype=torch.float64, device=self.device)
        status = torch.empty(B, dtype=torch.int32, device=self.device)
        self._ck(self.lib.lrx_search_finish(self.h, _ptr(rec_all), _ptr(max_all), _ptr(flags_all),
                                            world, B, k, mode, _ptr(weights), _ptr(ids), _ptr(score),
                                            _ptr(sem), _ptr(kw), _ptr(status)))
        return ids, score, sem, kw, status

    # ------------------------------------------------ packed form (one all-gather)
    def packed_bytes(self, B: int, k: int) -> int:
        return int(self.lib.lrx_packed_bytes(B, k))

    def search_local_packed(self, q, q_terms, q_ptr, k: int, mode: int, out: torch.Tensor,
                            width: int = 0):
        """K2 + K3 on this shard into `out` (uint8 [packed_bytes]) -- the all-gather unit."""
        B = int(q.shape[0])
        self._cap(q_terms, B)
        self._ck(self.lib.lrx_search_local_packed(self.h, _ptr(q), _ptr(q_terms), _ptr(q_ptr), B, k,
                                                  mode, width, _ptr(out)))
        return out

    def search_finish_packed(self, packed_all, world: int, B: int, k: int, mode: int, weights, outs):
        """K4 on the gathered [world][packed] buffer; outs = (ids, score, sem, kw, status)."""
        ids, score, sem, kw, status = outs
        self._ck(self.lib.lrx_search_finish_packed(self.h, _ptr(packed_all), world, B, k, mode,
                                                   _ptr(weights), _ptr(ids), _ptr(score), _ptr(sem),
                                                   _ptr(kw), _ptr(status)))
        return outs

    # ------------------------------------ peer-memory exchange (world > 1, one box)
    def exchange_setup(self, B_max: int, k_max: int, group=None) -> bool:
        """Allocate this rank's exchange region, swap IPC handles with the other ranks (one
        all-gather of 64 bytes at start-up) and open theirs: afterwards `search_sharded` needs no
        collective call per query batch.  Returns False -- on EVERY rank, the outcome is agreed by
        an all-reduce -- when some rank could not export or map a region (e.g. no peer access
        between the GPUs); the caller then keeps the NCCL exchange."""
        import torch.distributed as dist
        from ._lib import LRX_IPC_HANDLE_BYTES, LrxError
        ok = 1
        mine = (C.c_ubyte * LRX_IPC_HANDLE_BYTES)()
        try:
            self._ck(self.lib.lrx_exchange_export(self.h, B_max, k_max, C.cast(mine, C.c_void_p)))
        except LrxError:
            ok = 0
        world = dist.get_world_size(group)
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        every = torch.empty(world * LRX_IPC_HANDLE_BYTES, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, t, group=group)
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # did everybody export?
        if int(flag.item()) == 1:
            host = every.cpu().numpy().tobytes()
            buf = (C.c_ubyte * len(host)).from_buffer_copy(host)
            try:
                self._ck(self.lib.lrx_exchange_import(self.h, C.cast(buf, C.c_void_p)))
            except LrxError:
                ok = 0
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)      # ... and map every region?
        return int(flag.item()) == 1       # the all-reduce also orders "mapped" before the first store

    def search_sharded(self, q, q_terms, q_ptr, k: int, mode: int, weights, outs, width: int = 0):
        """K2 + K3 on this shard, block stored into every peer's region (world > 1), K4: the
        replicated result.  One captured launch chain per (shape, buffers), replayed per call."""
        B = int(q.shape[0])
        ids, score, sem, kw, status = outs
        if q_terms.numel() == 0:
            q_terms = torch.zeros(1, dtype=torch.int32, device=self.device)
        self._cap(q_terms, B)
        self._ck(self.lib.lrx_search_sharded(self.h, _ptr(q), _ptr(q_terms), _ptr(q_ptr), _ptr(weights),
                                             B, k, mode, width, _ptr(ids), _ptr(score), _ptr(sem),
                                             _ptr(kw), _ptr(status)))
        return outs

    def alloc_outputs(self, B: int, k: int):
        return (torch.empty((B, k), dtype=torch.int64, device=self.device),
                torch.empty((B, k), dtype=torch.float64, device=self.device),
                torch.empty((B, k), dtype=torch.float64, device=self.device),
                torch.empty((B, k), dtype=torch.float64, device=self.device),
                torch.empty(B, dtype=torch.int32, device=self.device))

    # ----------------------------------------------------- whole search, host
    @staticmethod
    def _host_args(q_fp16, term_lists, weights):
        q_fp16 = np.ascontiguousarray(q_fp16, dtype=np.float16)
        B = q_fp16.shape[0]
        ptr = np.zeros(B + 1, dtype=np.int32)
        for i, t in enumerate(term_lists):
            ptr[i + 1] = ptr[i] + len(t)
        terms = np.fromiter((int(x) for t in term_lists for x in t), dtype=np.int32, count=int(ptr[-1]))
        if terms.size == 0:
            terms = np.zeros(1, dtype=np.int32)
        return q_fp16, B, terms, ptr, np.ascontiguousarray(weights, dtype=np.float64)

    def search_batch_host(self, q_fp16: np.ndarray, term_lists: Sequence[Sequence[int]], k: int,
                          weights: Sequence[float], fusion: str = "linear"):
        """Host buffers in, host buffers out (H2D + kernels + D2H inside the call); on a shard
        (world > 1) every rank makes the same call and gets the replicated result.
        Returns (ids int64 [B,k], score, semantic, keyword float64 [B,k])."""
        self.search_host_begin(q_fp16, term_lists, k, weights, fusion)
        return self.search_host_end()

    def search_host_begin(self, q_fp16: np.ndarray, term_lists: Sequence[Sequence[int]], k: int,
                          weights: Sequence[float], fusion: str = "linear"):
        """Stage + enqueue (H2D, chain, D2H) and return at once; `search_host_end` collects.  With
        two handles over one index (`clone_view`) two batches are in flight."""
        q_fp16, B, terms, ptr, w = self._host_args(q_fp16, term_lists, weights)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.lrx_search_host_begin(self.h, vp(q_fp16), vp(terms), vp(ptr), vp(w), B, k,
                                                FUSION[fusion]))
        self._pending = (B, k)

    def search_host_end(self):
        B, k = self._pending
        self._pending = None
        ids = np.empty((B, k), dtype=np.int64)
        score = np.empty((B, k), dtype=np.float64)
        sem = np.empty((B, k), dtype=np.float64)
        kw = np.empty((B, k), dtype=np.float64)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.lrx_search_host_end(self.h, vp(ids), vp(score), vp(sem), vp(kw)))
        return ids, score, sem, kw

    def search_text_host_begin(self, tok_ids: np.ndarray, tok_lens: np.ndarray, term_lists, k: int,
                               weights: Sequence[float], fusion: str = "linear"):
        """`search_text_host` split in two: stage + enqueue (H2D, encoder, chain, D2H) and return at
        once; `search_host_end` collects.  Several handles over one index keep several batches in
        flight, the encoder of one under the scans of another."""
        tok_ids = np.ascontiguousarray(tok_ids, dtype=np.int32)
        tok_lens = np.ascontiguousarray(tok_lens, dtype=np.int32)
        B, S = tok_ids.shape
        _, _, terms, ptr, w = self._host_args(np.zeros((B, LRX_DIM), np.float16), term_lists, weights)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.lrx_search_text_host_begin(self.h, vp(tok_ids), vp(tok_lens), S, vp(terms), vp(ptr),
                                                     vp(w), B, k, FUSION[fusion]))
        self._pending = (B, k)

    def search_text_host(self, tok_ids: np.ndarray, tok_lens: np.ndarray, term_lists, k: int,
                         weights: Sequence[float], fusion: str = "linear"):
        """The whole of RetrievalEngine.search for tokenised strings: WordPiece ids [B,S] + lens [B]
        and BM25 term-id lists in, fused results out (K1 -> K2 || K3 -> K4 in one call)."""
        tok_ids = np.ascontiguousarray(tok_ids, dtype=np.int32)
        tok_lens = np.ascontiguousarray(tok_lens, dtype=np.int32)
        B, S = tok_ids.shape
        _, _, terms, ptr, w = self._host_args(np.zeros((B, LRX_DIM), np.float16), term_lists, weights)
        ids = np.empty((B, k), dtype=np.int64)
        score, sem, kw = (np.empty((B, k), dtype=np.float64) for _ in range(3))
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        self._ck(self.lib.lrx_search_text_host(self.h, vp(tok_ids), vp(tok_lens), S, vp(terms), vp(ptr),
                                               vp(w), B, k, FUSION[fusion], vp(ids), vp(score), vp(sem),
                                               vp(kw)))
        return ids, score, sem, kw

    # ------------------------------------------------------ K1 stage: GEMM
    def gemm_f16(self, a: torch.Tensor, w: torch.Tensor, epi: int = 3, bias=None, residual=None,
                 gamma=None, beta=None, eps: float = 1e-12):
        """out[M,N] = epi(a[M,K] @ w[N,K]^T) on the tensor cores (lrx_gemm_f16)."""
        assert a.dtype == torch.float16 and w.dtype == torch.float16
        assert a.is_contiguous() and w.is_contiguous()
        M, K = a.shape
        N = w.shape[0]
        out = torch.empty((M, N), dtype=torch.float32 if epi == 3 else torch.float16,
                          device=self.device)
        self._ck(self.lib.lrx_gemm_f16(self.h, _ptr(a), _ptr(w), M, N, K, epi, _ptr(bias),
                                       _ptr(residual), _ptr(gamma), _ptr(beta), float(eps), _ptr(out)))
        return out

    def profile(self, on: bool):
        self._ck(self.lib.lrx_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, which: int):
        """(summed device ms, launches) of kernel `which` (0 dense scan, 1 BM25 scan)."""
        ms, n = C.c_double(), C.c_int64()
        self._ck(self.lib.lrx_profile_read(self.h, which, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def launches(self) -> int:
        return int(self.lib.lrx_launch_count(self.h))
