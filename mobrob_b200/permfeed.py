"""Host-side minibatch permutations for PPO.train, produced ahead of the GPU.

[SB3 2.0.0] RolloutBuffer.get draws ``np.random.permutation(n_envs * n_steps)`` on the host once
per epoch (reached from src/mobrob/rl_control/ppo.py:73-74).  At B200 batch sizes that single
legacy-MT19937 stream costs more host time than the whole device iteration, so the permutations
of iteration k+1 are drawn by a pool of worker threads
into pinned memory while the GPU runs iteration k, and copied on a side stream.  Each one is
``mr_host_permutation(seed, stream = (rank, iteration, epoch))`` (csrc/host_perm.cu: a cache-friendly
scatter shuffle on xoshiro256**), i.e. the index sequence is a pure function of the seed and the
position in training, independent of the worker count.  The bit-for-bit SB3 stream
(``np.random`` global state) remains available as PPO(permutation="sb3").
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import torch

from . import _lib


class PermutationFeeder:
    SLOTS = 2

    def __init__(self, n: int, n_epochs: int, device: torch.device, seed: int = 0, rank: int = 0,
                 workers: int | None = None):
        self.n, self.n_epochs, self.device = int(n), int(n_epochs), device
        self.seed, self.rank = int(seed) & 0xFFFFFFFF, int(rank)
        self.lib = _lib.load()
        workers = workers or max(1, min(n_epochs, (os.cpu_count() or 2) - 1))
        self.pool = ThreadPoolExecutor(max_workers=workers, thread_name_prefix="mr-perm")
        self.pinned = [torch.empty((n_epochs, n), dtype=torch.int64).pin_memory() for _ in range(self.SLOTS)]
        self.host = [p.numpy() for p in self.pinned]
        self.dev = [torch.empty((n_epochs, n), dtype=torch.int64, device=device) for _ in range(self.SLOTS)]
        self.copy_stream = torch.cuda.Stream(device)
        self.futures = [None] * self.SLOTS           # per slot: list of futures (one per epoch)
        self.iteration = [None] * self.SLOTS         # which iteration the slot holds / is being filled with
        self.copied = [None] * self.SLOTS            # per slot: list of events, H2D done
        self.released = [None] * self.SLOTS          # per slot: event on the compute stream, readers done
        self.h2d_bytes = 0

    def _draw(self, slot: int, iteration: int, epoch: int) -> None:
        # ctypes releases the GIL for the duration of the call: the workers run in parallel
        stream = (self.rank << 48) ^ (iteration << 16) ^ epoch
        _lib.check(self.lib.mr_host_permutation(self.seed, stream, self.n, self.host[slot][epoch].ctypes.data))

    def prefetch(self, iteration: int) -> None:
        """Start drawing the permutations of `iteration` (no-op if already under way)."""
        slot = iteration % self.SLOTS
        if self.iteration[slot] == iteration:
            return
        if self.copied[slot] is not None:            # the pinned rows are still the source of old copies
            for ev in self.copied[slot]:
                ev.synchronize()
            self.copied[slot] = None
        self.iteration[slot] = iteration
        self.futures[slot] = [self.pool.submit(self._draw, slot, iteration, e) for e in range(self.n_epochs)]

    def stage(self, iteration: int) -> None:
        """Issue the H2D copies of `iteration` on the side stream (waits for the workers)."""
        slot = iteration % self.SLOTS
        self.prefetch(iteration)
        if self.copied[slot] is not None:
            return
        if self.released[slot] is not None:          # device rows may still be read by an older epoch
            self.copy_stream.wait_event(self.released[slot])
            self.released[slot] = None
        events = []
        with torch.cuda.stream(self.copy_stream):
            for e, fut in enumerate(self.futures[slot]):
                fut.result()
                self.dev[slot][e].copy_(self.pinned[slot][e], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append(ev)
                self.h2d_bytes += self.n * 8
        self.copied[slot] = events

    def get(self, iteration: int, epoch: int) -> torch.Tensor:
        """Device permutation of (iteration, epoch); the current stream waits for its copy."""
        slot = iteration % self.SLOTS
        self.stage(iteration)
        torch.cuda.current_stream(self.device).wait_event(self.copied[slot][epoch])
        return self.dev[slot][epoch]

    def release(self, iteration: int) -> None:
        """The current stream has enqueued its last read of `iteration`'s rows."""
        slot = iteration % self.SLOTS
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.released[slot] = ev

    def close(self) -> None:
        self.pool.shutdown(wait=True, cancel_futures=True)
