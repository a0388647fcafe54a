"""Host-side sharding of a batch of independent blocks across GPUs (BASELINE config 4).

The path shards naturally: block b goes to rank b mod world, every rank runs its blocks on its
own context (one GPU, one stream), and only tiny per-block results (primary index, digest) are
gathered -- there is no collective on the data path."""
import hashlib


def blocks_for_rank(n_blocks, rank, world):
    """Round-robin assignment: the blocks rank `rank` of `world` processes."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_blocks, world))


def run_batch(n_blocks, rank, world, make_block, bwt_fn):
    """Process this rank's share.  make_block(b) -> uint8 array; bwt_fn(T) -> (U, primary).
    Returns {block: (primary, sha256(U))}."""
    out = {}
    for b in blocks_for_rank(n_blocks, rank, world):
        U, primary = bwt_fn(make_block(b))
        out[b] = (int(primary), hashlib.sha256(U.tobytes()).hexdigest())
    return out


def gather_results(local, dist=None):
    """Merge the per-rank dictionaries on every rank (torch.distributed all_gather_object)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    parts = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError("block %d processed twice" % k)
            merged[k] = v
    return merged
