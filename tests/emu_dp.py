"""Single-process emulation of the data-parallel exchange for the tests: ranks run one after the
other; a collective call returns what the same call produced on every rank in the PREVIOUS sweep.
Three sweeps settle every value (operands -> row statistics -> loss), so the last sweep is exactly
what W processes would compute."""
import torch


class EmuDP:
    def __init__(self, rank, world, store, grad_scale=1.0):
        self.rank, self.world, self.store, self.grad_scale = rank, world, store, grad_scale
        self.calls = 0

    def _exchange(self, t):
        idx = self.calls
        self.calls += 1
        mine = self.store.setdefault(self.rank, {})
        mine[idx] = t.detach().clone()
        parts = []
        for r in range(self.world):
            prev = self.store.get(("prev", r), {}).get(idx)
            parts.append(prev if prev is not None else torch.zeros_like(t))
        return parts

    def all_gather_rows(self, t):
        return torch.cat(self._exchange(t.contiguous()), 0)

    def all_reduce_sum(self, t):
        return torch.stack(self._exchange(t.contiguous())).sum(0)

    def all_reduce_max_(self, t):
        t.copy_(torch.stack(self._exchange(t.contiguous())).amax(0))
        return t


def run_emulated(world, fn, sweeps=3):
    """fn(dp, rank) -> result; returns the per-rank results of the last sweep."""
    store, out = {}, None
    for _ in range(sweeps):
        out = []
        for r in range(world):
            out.append(fn(EmuDP(r, world, store), r))
        for r in range(world):
            store[("prev", r)] = store.pop(r, {})
    return out
