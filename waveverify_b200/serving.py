"""Request batcher for serving.

The reference's web example (examples/web_api_integration.py:260-460) calls `WaveVerify.embed` /
`.detect` once per HTTP request, i.e. with a batch of one clip.  This module is the piece that sits
between such endpoints and the batched GPU path: concurrent requests are queued, a worker thread
collects up to `max_batch` of them (waiting at most `max_wait_s` for stragglers), pads the clips of
one operation to a common length and runs ONE `embed_batch` / `detect_batch` / `locate_batch` call.
The HTTP layer itself is out of scope.

Which requests may share a batch without changing any result: clips of a batch are independent
(tests/test_gpu_parity.py::test_batch_independence_bit_exact) and the networks are causal, so zeros
appended after a clip do not change earlier outputs - PROVIDED the clip length is a multiple of the total
stride (`hop` = 320 samples).  A clip that ends inside a hop is different: the reference zero-pads the
INPUT OF EACH strided conv (modules/conv.py:160-203), whereas a zero-padded waveform produces non-zero
(bias-driven) activations after the clip end, so the last partial hop would differ (measured: up to 3e-4).
Hence: hop-aligned clips are padded to the longest one of the batch; unaligned clips only share a batch
with clips of exactly the same length.  Results are then bit-identical to one call per clip; the
detector's bit decode uses the masked time mean over the clip's own samples (scripts/evaluate.py:471-494),
which equals the unmasked mean of the un-padded clip up to the 1e-8 in its denominator.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future
from dataclasses import dataclass
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

OPS = ("embed", "detect", "locate")


@dataclass
class _Request:
    op: str
    audio: torch.Tensor            # [T] fp32 on the host
    msg: Optional[torch.Tensor]    # [nbits] fp32 (embed only)
    future: Future


class RequestBatcher:
    """`backend` needs `embed_batch(audio[B,1,T], msg[B,nbits]) -> y[B,1,T]`,
    `detect_batch(audio, presence=mask[B,1,T]) -> (bits[B,nbits], conf[B])`,
    `locate_batch(audio) -> mask[B,T]` and a `.device` (waveverify_b200.WaveVerify has exactly these)."""

    def __init__(self, backend, max_batch: int = 64, max_wait_s: float = 0.002, hop: int = 320):
        if max_batch < 1 or hop < 1 or max_wait_s < 0:
            raise ValueError("max_batch >= 1, hop >= 1, max_wait_s >= 0")
        self.backend = backend
        self.max_batch, self.max_wait_s, self.hop = int(max_batch), float(max_wait_s), int(hop)
        self._q: "queue.Queue[Optional[_Request]]" = queue.Queue()
        self._closed = False
        self.batches: List[int] = []          # sizes of the batches that ran (for monitoring / tests)
        self._worker = threading.Thread(target=self._run, name="wv-batcher", daemon=True)
        self._worker.start()

    # ---- client side -----------------------------------------------------------------------------
    def _submit(self, op: str, audio, msg=None) -> Future:
        if self._closed:
            raise RuntimeError("RequestBatcher is closed")
        a = torch.as_tensor(audio, dtype=torch.float32).detach().cpu().reshape(-1)
        if a.numel() == 0:
            raise ValueError("empty audio")
        m = None
        if op == "embed":
            m = torch.as_tensor(msg, dtype=torch.float32).detach().cpu().reshape(-1)
        f: Future = Future()
        self._q.put(_Request(op, a, m, f))
        return f

    def embed(self, audio, msg) -> Future:
        """-> Future of the watermarked clip, np.float32 [T]."""
        return self._submit("embed", audio, msg)

    def detect(self, audio) -> Future:
        """-> Future of (bits np.uint8 [nbits], confidence float)."""
        return self._submit("detect", audio)

    def locate(self, audio) -> Future:
        """-> Future of the presence mask, np.uint8 [T]."""
        return self._submit("locate", audio)

    def close(self) -> None:
        if not self._closed:
            self._closed = True
            self._q.put(None)
            self._worker.join()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- worker ----------------------------------------------------------------------------------
    def _collect(self) -> Optional[List[_Request]]:
        first = self._q.get()
        if first is None:
            return None
        batch = [first]
        import time
        deadline = time.monotonic() + self.max_wait_s
        while len(batch) < self.max_batch:
            left = deadline - time.monotonic()
            try:
                r = self._q.get(timeout=max(left, 0.0)) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if r is None:
                self._q.put(None)          # let the outer loop see the shutdown after this batch
                break
            batch.append(r)
        return batch

    def _run(self) -> None:
        while True:
            batch = self._collect()
            if batch is None:
                return
            groups = {}
            for r in batch:     # hop-aligned clips pad to a common length; others need an exact length match
                n = int(r.audio.numel())
                groups.setdefault((r.op, 0 if n % self.hop == 0 else n), []).append(r)
            for (op, _), reqs in groups.items():
                try:
                    self._execute(op, reqs)
                except Exception as e:  # noqa: BLE001  (a failed batch fails its requests, not the worker)
                    for r in reqs:
                        if not r.future.done():
                            r.future.set_exception(e)

    def _execute(self, op: str, reqs: Sequence[_Request]) -> None:
        dev = self.backend.device
        lens = [int(r.audio.numel()) for r in reqs]
        T = max(lens)
        B = len(reqs)
        host = torch.zeros(B, 1, T, dtype=torch.float32)
        for i, r in enumerate(reqs):
            host[i, 0, :lens[i]] = r.audio
        x = host.to(dev, non_blocking=False)
        self.batches.append(B)
        if op == "embed":
            msg = torch.stack([r.msg for r in reqs]).to(dev)
            y = self.backend.embed_batch(x, msg).cpu().numpy()
            for i, r in enumerate(reqs):
                r.future.set_result(np.ascontiguousarray(y[i, 0, :lens[i]]))
        elif op == "detect":
            presence = torch.zeros(B, 1, T, dtype=torch.uint8)
            for i in range(B):
                presence[i, 0, :lens[i]] = 1
            bits, conf = self.backend.detect_batch(x, presence=presence.to(dev))
            bits, conf = bits.cpu().numpy(), conf.cpu().numpy()
            for i, r in enumerate(reqs):
                r.future.set_result((bits[i].astype(np.uint8), float(conf[i])))
        else:
            mask = self.backend.locate_batch(x).cpu().numpy()
            for i, r in enumerate(reqs):
                r.future.set_result(np.ascontiguousarray(mask[i, :lens[i]].astype(np.uint8)))
