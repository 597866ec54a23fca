"""TokenTap: how the host *observes* tokens of a device-closed decode loop.

With the token loop closed on the device (the finalising shard's step kernel stores the sampled
token straight into the head shard's lane slot) no host is on the token's critical path; the API
still has to see every token (reference api/inference.py:135-212 consumes one TokenResult per
step).  The step kernel therefore also writes (logprob, then token) into a pinned host ring entry
handed to it as ``token_out`` / ``logprob_out``; a watcher thread polls the entries in launch order
per lane -- no stream synchronisation, no D2H copy enqueue -- and turns each into the same
final ``ActivationMessage`` the reference's end shard emits (fit_in_memory.py:168-181).

Entry protocol: the compute thread resets the entry's token word to -1 before the launch; the kernel
writes the logprob, fences at system scope, then writes the token (>= 0, or <= -1000 when a bounded
in-kernel wait timed out: -(1000 + code)).
"""
from __future__ import annotations

import threading
import time
from collections import deque
from typing import Any, Callable, Deque, Dict, Optional, Tuple

import numpy as np
import torch

SENTINEL = -1


class TokenTap:
    def __init__(self, n_lanes: int, depth: int = 512, on_token: Optional[Callable[[Any, int, float], None]] = None):
        self.n_lanes, self.depth = int(n_lanes), int(depth)
        buf = torch.full((self.n_lanes, self.depth, 2), SENTINEL, dtype=torch.int32)
        if torch.cuda.is_available():
            buf = buf.pin_memory()
        self._buf = buf
        self._i32 = buf.numpy()
        self._f32 = self._i32.view(np.float32)
        self._base = buf.data_ptr()
        self._head = [0] * self.n_lanes          # next entry index to hand out, per lane
        self._pending: Dict[int, Deque[Tuple[int, Any]]] = {l: deque() for l in range(self.n_lanes)}
        self._lock = threading.Lock()
        self._wake = threading.Event()
        self.on_token = on_token
        self._thread: Optional[threading.Thread] = None
        self._running = False
        self.delivered = 0

    # -- compute thread ----------------------------------------------------------------
    def post(self, lane: int, info: Any) -> Tuple[int, int]:
        """Reserve the lane's next entry for a step about to be launched; returns the device-visible
        (token_out, logprob_out) addresses."""
        with self._lock:
            q = self._pending[lane]
            if len(q) >= self.depth:
                raise RuntimeError(f"token tap overflow on lane {lane}: {len(q)} steps in flight")
            idx = self._head[lane]
            self._head[lane] = (idx + 1) % self.depth
            self._i32[lane, idx, 0] = SENTINEL
            q.append((idx, info))
        self._wake.set()
        off = ((lane * self.depth) + idx) * 8
        return self._base + off, self._base + off + 4

    def in_flight(self) -> int:
        with self._lock:
            return sum(len(q) for q in self._pending.values())

    # -- watcher -------------------------------------------------------------------------
    def poll_once(self) -> int:
        """Deliver every token that has landed, in launch order per lane; returns how many."""
        n = 0
        for lane in range(self.n_lanes):
            q = self._pending[lane]
            while q:
                idx, info = q[0]
                tok = int(self._i32[lane, idx, 0])
                if tok == SENTINEL:
                    break
                lp = float(self._f32[lane, idx, 1])
                with self._lock:
                    q.popleft()
                n += 1
                self.delivered += 1
                if self.on_token is not None:
                    try:
                        self.on_token(info, tok, lp)
                    except Exception:   # a consumer error must not stop the tap
                        import logging
                        logging.getLogger("dnet").exception("token tap consumer failed")
        return n

    def _run(self) -> None:
        # Never spin: this thread shares the GIL with the compute thread that launches the step kernels, and a
        # Python busy loop here competes with those launches for the interpreter (the 8-shard ring went from
        # 2,268 to 2,937 tok/s over the set of changes that included replacing the spin with this nap; the
        # nap's own share was not isolated).  A 50 us nap between polls costs a token ~25 us of observation
        # latency on average and nothing on the device.
        while self._running:
            self.poll_once()
            if self.in_flight() == 0:
                self._wake.wait(timeout=0.05)
                self._wake.clear()
            else:
                time.sleep(5e-5)

    def start(self) -> None:
        if self._thread is None:
            self._running = True
            self._thread = threading.Thread(target=self._run, name="dnet-token-tap", daemon=True)
            self._thread.start()

    def stop(self) -> None:
        self._running = False
        self._wake.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
            self._thread = None

    def drain(self, timeout_s: float = 30.0) -> bool:
        """Block until every posted step has been delivered (tests / shutdown)."""
        t0 = time.perf_counter()
        while self.in_flight() > 0:
            if self._thread is None:
                self.poll_once()
            if time.perf_counter() - t0 > timeout_s:
                return False
            time.sleep(1e-4)
        return True
