"""Host-side helpers for ceremonies in flight (dkgv_share_matrix_enqueue_sharded[_dev] / _settle_sharded[_dev], include/dkgv.h):
how many ctx / stream lanes a GPU gets, how many distinct input copies keep a ring of ceremonies out of the L2, and the decision
`settle` takes from the gathered flag words (mirrors csrc/comm.cu: every rank sees every rank's flags, so all ranks agree)."""

L2_BYTES = 126 << 20


def lanes_for(rows):
    """ctxs (own stream and scratch, ONE shared fixed-base table) per GPU for row blocks of `rows` dealers.  The difference tables
    (ALU pipe, barrier latency) of one ceremony run under the fixed-base multiplications (multiplier pipe) of another.  Measured
    (profiles/r2_pipeline_lanes.md), ms per ceremony with 1 / 2 / 4 lanes: 1024 dealers 5.71 / 5.35 / 5.36, 512: 3.15 / 2.69 / 2.69,
    256: 1.73 / 1.39 / 1.37, 128 (one wave of tables, latency-bound alone): 1.10 / 0.81 / 0.73."""
    return 2 if rows >= 512 else 4


def ring_size(bytes_per_ceremony, l2_bytes=L2_BYTES, cap=64):
    """distinct device copies of the inputs so that a ceremony never finds its verification vectors or shares in the L2:
    the ring holds more than twice the L2"""
    return int(min(cap, max(2, -(-2 * l2_bytes // max(1, int(bytes_per_ceremony))) + 1)))


def settle_needed(flags, world):
    """flags: 2 * world words as dkgv_share_matrix_enqueue_sharded[_dev] leaves them (rank r: [ids are no permutation of 1..n,
    dealers the shortcut could not settle]).  True: the ceremony has to be run again - on EVERY rank, since every rank holds all flags."""
    return any(int(flags[2 * r]) != 0 or int(flags[2 * r + 1]) != 0 for r in range(world))
