"""CUDA-graph capture of a whole rasterizer step (forward -> loss -> backward [-> optimiser]).

The reference cannot do this: its forward blocks the host on a D2H copy of `num_rendered` and sizes its buffers from
it (RAST/cuda_rasterizer/rasterizer_impl.cu:282, RAST/rasterize_points.cu:60-92).  Here the forward is pure stream work
once the (tile, Gaussian) pair capacity is fixed (rasterizer.set_static_capacity): camera matrices are read from device
memory, every size that depends on the view lives on the device behind a capacity guard, and the only host-visible
by-product -- the pair / visible counts -- lands in a pinned status slot that `check()` reads afterwards.

    step = GraphedStep(fn)            # fn(): e.g. zero grads, render, loss, backward on STATIC input tensors
    for it in range(n):
        cam_buffer.copy_(next_camera, non_blocking=True)    # inputs change by writing into the captured tensors
        step.replay()                                       # one cudaGraphLaunch: no Python, no host wait per kernel
    step.check()                                            # raises if a replay overflowed the pair capacity

Capacity: measured over `warmup` eager runs of fn (x `headroom`), or given.  A replay whose view has more pairs than
that renders nothing; replay() checks the PREVIOUS replay's counts (free: pinned memory) and check() the last one.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import rasterizer as R


class GraphedStep:
    def __init__(self, fn: Callable[[], object], pair_capacity: Optional[int] = None, warmup: int = 3,
                 headroom: float = 2.0, device=None, pool=None):
        self.idx = torch.cuda.current_device() if device is None else torch.device(device).index
        self.fn = fn
        # Warm-up AND capture run on one side stream: autograd pins every leaf's AccumulateGrad node to the stream it was
        # created on, and a node left over from the default stream (or from another side stream) would make the engine
        # synchronise across streams in the middle of the capture.  fn must not keep the previous iteration's autograd
        # graph alive (drop old outputs before the forward), or those nodes are never re-created.
        side = torch.cuda.Stream(device=self.idx)
        side.wait_stream(torch.cuda.current_stream(self.idx))
        peak = 0
        with torch.cuda.stream(side):
            for _ in range(max(int(warmup), 2)):
                fn()
                peak = max(peak, R._last_pairs.get(self.idx, 0))
        torch.cuda.current_stream(self.idx).wait_stream(side)
        torch.cuda.synchronize(self.idx)
        self.pair_capacity = R._round_cap(pair_capacity if pair_capacity is not None else peak * headroom + 4096)
        self.graph = torch.cuda.CUDAGraph()
        n0 = len((R._static.get(self.idx) or {}).get("graph", []))
        with R.static_capacity(self.pair_capacity, self.idx):
            with torch.cuda.graph(self.graph, pool=pool, stream=side):
                self.outputs = fn()
        self.tickets = list(R._static[self.idx]["graph"][n0:])      # the forwards captured by THIS graph
        self.replays = 0

    def replay(self):
        if self.replays:
            self._check(False)        # the previous replay's counts (advisory while it may still be running)
        self.graph.replay()
        self.replays += 1
        return self.outputs

    def _check(self, final: bool):
        last = {}
        for ticket, cap in self.tickets:
            R._check_ticket(self.idx, ticket, cap, False)
            last = R._last_counts.get(self.idx, {})
        return dict(last)

    def check(self):
        """Synchronises and verifies the most recent replay; returns its counts (num_pairs, num_visible, num_rendered)."""
        torch.cuda.synchronize(self.idx)
        return self._check(True)
