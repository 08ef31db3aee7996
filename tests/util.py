"""Shared helpers for the parity tests."""
import numpy as np

P = 0xFFFFFFFF00000001


def splitmix64(seed, n):
    """SplitMix64 stream (SURVEY.md §8(d) synthetic inputs), vectorised."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_felts(seed, shape, canonical=True):
    n = int(np.prod(shape))
    v = splitmix64(seed, n)
    if canonical:
        v = np.where(v >= np.uint64(P), v - np.uint64(P), v)
    return v.reshape(shape)


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def sighash_whitelist_leaves(wl):
    """The 2^tree_height leaves of the reference's sighash-circuit whitelist tree (tests/golden/sighash_whitelist.json):
    the gadget ids in the order BlockSpendCoreConfig::generate_id_permutations emits them
    (city_rollup_common/src/introspection/rollup/introspection.rs:402-431) index the fingerprint table; leaf i is the
    fingerprint of the i-th id in SORTED order (derive(Ord) over num_deposits, num_withdrawals, last_block_num_deposits,
    last_block_num_withdrawals, current_spend_index, :156-163; city_store/src/store/sighash/mod.rs:49-66); every other leaf is
    the zero hash."""
    import numpy as np

    ni, no = wl["max_deposits"] + 1, wl["max_withdrawals"] + 1
    ids = [(nd, nw, lbd, lbw, csi) for lbw in range(no) for lbd in range(ni) for nw in range(no) for nd in range(ni)
           for csi in range(nd + 1)]
    assert len(ids) == len(wl["fingerprints"])
    order = sorted(range(len(ids)), key=lambda i: ids[i])
    leaves = np.zeros((1 << wl["tree_height"], 4), dtype=np.uint64)
    for i, k in enumerate(order):
        leaves[i] = wl["fingerprints"][k]
    return leaves
