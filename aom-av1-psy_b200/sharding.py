"""Work partitioning across GPUs.

Two independent axes, no exchange during compute:
  * window-parallel: one ARF window per rank (round robin) -- no collective;
  * slab-parallel (4K): the block rows of one window are split into contiguous
    ranges; every rank needs all whole frames (the search range is +-1023 px, so
    a halo smaller than the frame cannot be bit-exact) and computes only its
    output rows, which are then gathered, plus a 16-byte sum for FRAME_DIFF.
The reference's counterpart is the row job queue of av1/encoder/ethread.c:2062-2189.
"""


def slab_rows(mb_rows, world, rank):
    """Contiguous block-row range [begin, end) of `rank` out of `world`."""
    return (mb_rows * rank) // world, (mb_rows * (rank + 1)) // world


def max_slab_rows(mb_rows, world):
    return max(slab_rows(mb_rows, world, r)[1] - slab_rows(mb_rows, world, r)[0] for r in range(world))


def window_owner(window_index, world):
    """Round-robin owner of an independent window."""
    return window_index % world


def merge_slabs(slabs, mb_rows, world, block_h):
    """Stack per-rank row slabs (each padded to max_slab_rows*block_h rows) into the full plane."""
    import numpy as np
    rows = []
    for r, s in enumerate(slabs):
        b, e = slab_rows(mb_rows, world, r)
        rows.append(s[: (e - b) * block_h])
    return np.concatenate(rows, axis=0)
