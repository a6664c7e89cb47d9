"""Multi-GPU sharding of independent units (families / reads): contiguous blocks per rank, no data-path
collective, host-side gather in input order. One process per GPU under torchrun."""


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of `n_items` owned by `rank` (blocks differ by at most one item)."""
    lo = n_items * rank // world
    hi = n_items * (rank + 1) // world
    return lo, hi


def step_family_range(step, rank, world, families_per_step):
    """bench.py: step k of rank r aligns families [(k * world + r) * F, (k * world + r + 1) * F)."""
    first = (step * world + rank) * families_per_step
    return first, first + families_per_step


def gather_in_order(local_items, group=None):
    """All ranks contribute their block; every rank gets the concatenation in rank (= input) order."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return list(local_items)
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, list(local_items), group=group)
    return [x for part in parts for x in part]


def sharded_msa(fastas, run_local, group=None):
    """Each rank runs `run_local(list_of_fasta) -> list_of_results` on its block; results gathered in input order."""
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(len(fastas), rank, world)
    return gather_in_order(run_local(fastas[lo:hi]), group)
