"""Micro-batching front of ``RetrievalEngine`` for a threaded server.

The reference's handler calls ``engine.query(request.query)`` on the event loop
(``src/server/app.py:109-120``), so requests are served one after the other and each of its 1-4
``RetrievalEngine.search`` calls is a GPU round trip of its own.  Once the handler hands the call to
a worker thread (``await loop.run_in_executor(None, engine.query, request.query)`` -- the one-line
change INTEGRATION.md shows -- or a plain ``def`` handler, which FastAPI runs in its thread pool),
several requests are inside ``search`` at the same time; this front coalesces those concurrent calls
into ONE ``search_batch`` launch chain (one encoder pass, one matrix scan for every <= 4 queries) and
hands each caller its own result list.

    engine = MicroBatchingEngine(RetrievalEngine(store_dir))      # same search() signature
    engine.search("zero fir procedure", k=5, hybrid_weight=0.6)    # from any thread

A request waits at most ``max_wait_ms`` for company (default 0.2 ms: a fraction of one scan of
the corpus).  Calls with different ``k`` / ``fusion`` are batched separately.  The GPU handle is only ever touched by the
dispatcher thread, so the engine's "sequential calls only" contract (include/lrx.h) holds whatever
the server does.
"""
from __future__ import annotations

import threading
import time
from collections import deque
from concurrent.futures import Future
from typing import Deque, List, Optional, Sequence, Tuple


class MicroBatchingEngine:
    def __init__(self, engine, max_batch: int = 64, max_wait_ms: float = 0.2):
        self.engine = engine
        self.max_batch = int(max_batch)
        self.max_wait = float(max_wait_ms) * 1e-3
        self._q: Deque[Tuple[tuple, str, float, Future]] = deque()
        self._cv = threading.Condition()
        self._stop = False
        self.batches = 0                  # dispatched search_batch calls
        self.requests = 0                 # search() calls served
        self.largest_batch = 0
        self._t = threading.Thread(target=self._loop, name="lrx-microbatch", daemon=True)
        self._t.start()

    # ------------------------------------------------------------------ the reference surface
    def search(self, query: str, k: int = 5, hybrid_weight: float = 0.5, fusion: Optional[str] = None):
        fut: Future = Future()
        with self._cv:
            if self._stop:
                raise RuntimeError("engine is closed")
            self._q.append(((int(k), fusion), query, float(hybrid_weight), fut))
            self._cv.notify_all()
        return fut.result()

    def search_batch(self, queries: Sequence[str], k: int = 5, hybrid_weights: Optional[Sequence[float]] = None,
                     fusion: Optional[str] = None):
        """A caller's own batch (the orchestrator fan-out): its queries join the queue together."""
        queries = list(queries)
        weights = list(hybrid_weights) if hybrid_weights is not None else [0.5] * len(queries)
        futs = [Future() for _ in queries]
        with self._cv:
            if self._stop:
                raise RuntimeError("engine is closed")
            for q, w, f in zip(queries, weights, futs):
                self._q.append(((int(k), fusion), q, float(w), f))
            self._cv.notify_all()
        return [f.result() for f in futs]

    def __getattr__(self, name):          # chunks, encode, bm25, ... straight from the engine
        return getattr(self.engine, name)

    # ------------------------------------------------------------------ dispatcher
    def _take(self):
        """Block for the first request, give followers `max_wait` to arrive, return one group of
        requests that share (k, fusion)."""
        with self._cv:
            while not self._q and not self._stop:
                self._cv.wait()
            if not self._q:
                return None
            deadline = time.monotonic() + self.max_wait
            while len(self._q) < self.max_batch and not self._stop:
                left = deadline - time.monotonic()
                if left <= 0:
                    break
                self._cv.wait(left)
            key = self._q[0][0]
            group, rest = [], deque()
            while self._q:
                item = self._q.popleft()
                (group if item[0] == key and len(group) < self.max_batch else rest).append(item)
            self._q = rest
            return key, group

    def _loop(self):
        while True:
            got = self._take()
            if got is None:
                return
            (k, fusion), group = got
            try:
                res = self.engine.search_batch([g[1] for g in group], k, [g[2] for g in group], fusion)
                for g, r in zip(group, res):
                    g[3].set_result(r)
            except BaseException as e:         # every waiter sees the failure (HTTP 500 in the reference)
                for g in group:
                    if not g[3].done():
                        g[3].set_exception(e)
            self.batches += 1
            self.requests += len(group)
            self.largest_batch = max(self.largest_batch, len(group))

    def close(self):
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        self._t.join(timeout=5)
        with self._cv:
            while self._q:
                self._q.popleft()[3].set_exception(RuntimeError("engine is closed"))
        self.engine.close()
