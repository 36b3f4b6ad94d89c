"""CPU: the shard-cut planner of the multi-GPU bench (bench_spike.balanced_cuts): a timeline model per shard -- work before the
exchange of expected draw counts, the output branch, phase 1 with windows that widen with the loci in front, the rest -- solved for
equal finishing times.  Off by default in the bench (measured: no gain); kept honest here."""
import bench_spike as bs


def test_equal_cuts_without_balance():
    assert bs.balanced_cuts(4, 4, balance=False) == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert bs.balanced_cuts(1, 1) == [0.0, 1.0]


def test_balanced_cuts_equalise_the_model():
    for n in (2, 4, 8):
        cuts = bs.balanced_cuts(n, n)
        assert len(cuts) == n + 1 and cuts[0] == 0.0 and cuts[-1] == float(n)
        lens = [b - a for a, b in zip(cuts, cuts[1:])]
        assert all(x > 0 for x in lens)
        assert all(a >= b - 1e-9 for a, b in zip(lens, lens[1:])), "later shards simulate wider windows, so they get fewer reads"
        a_all = bs.PRE * max(lens)
        t = [bs.shard_finish(a, b, a_all) for a, b in zip(cuts, cuts[1:])]
        assert max(t) - min(t) < 0.05 * max(t), t
        equal = max(bs.shard_finish(g, g + 1, bs.PRE) for g in range(n))
        assert max(t) <= equal + 1e-6
