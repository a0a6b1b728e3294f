"""Shared helpers for the multi-process (gloo / NCCL) tests."""
import queue
import socket
import time

import torch.multiprocessing as mp


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def run_workers(target, world: int, args: tuple, timeout: float = 600.0):
    """Spawn `world` processes running target(rank, world, port, *args, q); collect one result per rank from the queue.
    Fails as soon as a worker dies without reporting (instead of waiting out the timeout)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, *args, q)) for r in range(world)]
    for p in procs:
        p.start()
    results, deadline = [], time.monotonic() + timeout
    try:
        while len(results) < world:
            try:
                results.append(q.get(timeout=2.0))
                continue
            except queue.Empty:
                pass
            dead = [p for p in procs if p.exitcode not in (None, 0)]
            if dead:
                raise AssertionError(f"worker(s) exited with {[p.exitcode for p in dead]} before reporting "
                                     "(traceback above)")
            if time.monotonic() > deadline:
                raise AssertionError(f"workers did not report within {timeout} s")
    finally:
        for p in procs:
            p.join(timeout=60 if len(results) == world else 1)
            if p.is_alive():
                p.kill()
    for p in procs:
        assert p.exitcode == 0, f"worker exit code {p.exitcode}"
    return results
