"""Block partitioning across GPUs.

fqzcomp5 blocks are independent, so N GPUs need no collective on the data path:
block b goes to rank b % world (the order hts_tpool dispatches them,
thread_pool.c:113-164), every rank codes its blocks on its own stream, and the
host gathers results back into dispatch order.
"""


def blocks_of_rank(nblocks, rank, world):
    """Indices of the blocks rank `rank` owns (round-robin)."""
    return list(range(rank, nblocks, world))


def owner(block, world):
    return block % world


def worker_slot(block, ngpu, workers):
    """The library's block calls (csrc/block.cu, b200fqz_*_blocks_multi) run block b on worker
    (b % ngpu, (b // ngpu) % workers): the index of that worker among ngpu * workers.  A caller that gives
    every worker its own output buffer (bench.py's config-5 workload) uses this to pick it: two blocks with
    the same slot are never in flight together."""
    return owner(block, ngpu) * workers + (block // ngpu) % workers


def gather_in_order(per_rank, nblocks, world):
    """per_rank[r] = results of blocks_of_rank(nblocks, r, world), in that order.
    Returns the results in block (dispatch) order, as the reference's ordered
    result queue delivers them."""
    out = [None] * nblocks
    for r in range(world):
        for j, b in enumerate(blocks_of_rank(nblocks, r, world)):
            out[b] = per_rank[r][j]
    assert all(o is not None for o in out)
    return out
