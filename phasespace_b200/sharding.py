"""Multi-GPU generation by contiguous event-index ranges.

The reference has no distributed code.  Events are independent and every draw is a pure function of
``(seed, global event index, column)``, so sharding is just arithmetic on index ranges: rank ``r`` of
``W`` generates ``[first_r, first_r + count_r)`` on its own GPU and the concatenation over ranks is
bit-identical to a single-GPU run.  There is no collective on the data path; the only exchange is the
tiny weight-statistics all-reduce (max / sum) used for normalisation and unweighting bounds.
"""

from __future__ import annotations

import torch
import torch.distributed as dist

from ._lib import ForbiddenDecayError


def shard_range(n_events: int, rank: int, world: int) -> tuple[int, int]:
    """``(first, count)`` of rank ``rank``: ranges of ``ceil(n/world)`` events, the last ones shorter."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world: {rank}/{world}")
    per = -(-int(n_events) // world)
    first = min(rank * per, int(n_events))
    return first, min(per, int(n_events) - first)


def _rank_world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def generate_sharded(decay, n_events: int, seed: int, *, boost_to=None, normalize_weights=True, group=None,
                     stats=None, first_event: int = 0, exact: bool = False, err=None):
    """This rank's share of ``decay.generate(n_events, seed=seed)``.

    Returns ``(first, count, weights, weights_max | None, particles)`` with ``particles`` a dict of
    ``(count, 4)`` CUDA tensors.  ``boost_to`` (if per event) is the *full* ``(n_events, 4)`` array or this
    rank's rows.  Outputs stay on their GPUs; gather them only if a consumer really needs one tensor.

    A kinematically forbidden decay raises ``ForbiddenDecayError`` exactly as ``generate`` does (the reference asserts
    "Forbidden decay", ps.py:357-363): statically for all-fixed masses, otherwise from the kernel's error flag, which
    costs one 4-byte read-back -- unless the caller passes its own zeroed ``err`` tensor ``(1,)`` int32 and checks
    it later (several launches can share one flag and one synchronisation).
    """
    rank, world = _rank_world(group)
    first, count = shard_range(n_events, rank, world)
    compiled = decay.compile()
    if compiled.n_user:
        raise ValueError("sharded generation needs in-kernel mass functions (no arbitrary Python callables)")
    decay._mark_generate_called()
    if compiled._static_forbidden is not None and boost_to is None:
        raise ForbiddenDecayError(f"Forbidden decay (particle {compiled._static_forbidden})")
    if boost_to is not None:
        boost_to = torch.as_tensor(boost_to, dtype=torch.float64).reshape(-1, 4)
        if boost_to.shape[0] == n_events and n_events != 1:
            boost_to = boost_to[first:first + count]
        elif boost_to.shape[0] not in (1, count):
            raise ValueError(f"The number of events requested ({n_events}) doesn't match the boost_to input size "
                             f"of {tuple(boost_to.shape)}")
        boost_to = boost_to.to(torch.device("cuda", torch.cuda.current_device())).contiguous()
    deferred = err is not None
    weights, weights_max, slab, err = compiled.launch(
        count, seed, first_event + first, boost_to=boost_to, normalize_weights=normalize_weights, exact=exact,
        stats=stats, err=err)
    if not deferred and count > 0 and (compiled.has_resonances or boost_to is not None) and int(err.item()):
        raise ForbiddenDecayError("Forbidden decay")
    parts = {name: slab[compiled.slots[name]] for name in compiled.names}
    return first, count, weights, weights_max, parts


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Combine per-rank ``[max w, sum w, sum w^2, max w_max]`` vectors (the kernel's epilogue output):
    MAX on entries 0 and 3, SUM on 1 and 2.  Two 16-byte all-reduces: pure latency."""
    out = stats.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        mx = out[[0, 3]].contiguous()
        sm = out[[1, 2]].contiguous()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(sm, op=dist.ReduceOp.SUM, group=group)
        out = torch.stack([mx[0], sm[0], sm[1], mx[1]])
    return out


def allreduce_count(count: int, device=None, group=None) -> int:
    """Global number of (accepted) events."""
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def allreduce_max(t: torch.Tensor, group=None, world: int | None = None) -> torch.Tensor:
    """MAX over ranks of a small tensor (the global maximum weight of a sharded ``GenMultiDecay``, genmultidecay.py:193).
    ``world`` > 1 without an initialised process group is an error: the ranks could not agree on the maximum."""
    if not (dist.is_available() and dist.is_initialized()):
        if world is not None and world > 1:
            raise RuntimeError("a global maximum over ranks needs an initialised torch.distributed process group")
        return t
    if dist.get_world_size(group) == 1:
        return t
    out = t.clone()
    dist.all_reduce(out, op=dist.ReduceOp.MAX, group=group)
    return out


def allreduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """SUM all-reduce of a small tensor (histograms, counters); identity in a single process."""
    out = t.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def generate_histograms_sharded(decay, n_events: int, seed: int, *, group=None, normalize: bool = True, **kwargs):
    """``histogram.generate_histograms`` over all ranks: every rank bins its own contiguous range of event indices on
    its GPU and the per-rank histograms (a few kB) are summed with one all-reduce.  Returns the global histograms on
    every rank, in the same form as ``generate_histograms``."""
    from .histogram import generate_histograms

    rank, world = _rank_world(group)
    first, count = shard_range(n_events, rank, world)
    parts, weight_hist = generate_histograms(decay, count, seed=int(seed), normalize=False, first_event=first, **kwargs)
    names = list(parts)
    flat = torch.cat([parts[name].reshape(-1) for name in names] + [weight_hist.reshape(-1)])
    flat = allreduce_sum(flat, group=group)
    bins = weight_hist.numel()
    rows = flat.view(-1, bins)
    if normalize:
        sums = rows.sum(dim=1, keepdim=True)
        rows = rows / torch.where(sums > 0, sums, torch.ones_like(sums))
    return {name: rows[4 * k:4 * k + 4] for k, name in enumerate(names)}, rows[4 * len(names)]
