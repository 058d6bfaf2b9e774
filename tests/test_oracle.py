"""CPU: pins oracle/ against every golden vector and known-answer test the reference holds
for the hot path, and transcribes the reference's 11 unit tests."""
import math

import numpy as np
import pytest

import oracle
from conftest import lhd_bins_ok, random_lhd


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert oracle.philox4x32_10([0] * 4, [0] * 2) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert oracle.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert oracle.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [
        0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_r_golden_design(r_golden):
    """python/comparison_r.py:126-133 and README.md:89-90."""
    v = oracle.maxpro_criterion(r_golden["design"])
    assert math.isclose(v, r_golden["r_measure"], rel_tol=r_golden["r_rel_tol"])
    assert v == r_golden["rust_value_readme"]  # 95.98505099515626, to the last digit
    assert oracle.maximin_criterion(r_golden["design"]) == 0.06403124237432846  # SURVEY 8c extra vector


def test_pymaxpro_vectors(pymaxpro_golden):
    """Author's NumPy restatement (python/pymaxpro.py:7-62); comparisons.py:19-34 uses np.isclose."""
    for c in pymaxpro_golden["cases"]:
        x = c["design"]
        assert math.isclose(oracle.maxpro_sum(x), c["maxpro_sum"], rel_tol=pymaxpro_golden["rel_tol"])
        assert math.isclose(oracle.maxpro_criterion(x), c["maxpro_criterion"], rel_tol=pymaxpro_golden["rel_tol"])


def test_maxpro_criterion_value_1():  # maxpro_utils.rs:117-124
    assert abs(oracle.maxpro_criterion([[0.0, 1.0], [1.0, 0.0]]) - 1.0) < 1e-9


def test_maxpro_criterion_value_2():  # maxpro_utils.rs:126-133
    assert abs(oracle.maxpro_criterion([[0.0, 2.0], [1.0, 0.0]]) - 0.5) < 1e-9


def test_maximin_criterion_value_1():  # maximin_utils.rs:85-92
    assert abs(oracle.maximin_criterion([[0.0, 1.0], [1.0, 0.0]]) - 2.0 ** 0.5) < 1e-9


def test_maximin_criterion_value_2():  # maximin_utils.rs:94-101
    assert abs(oracle.maximin_criterion([[0.0, 2.0], [1.0, 0.0]]) - 5.0 ** 0.5) < 1e-9


def test_degenerate_shapes():
    assert oracle.maxpro_criterion([[0.3, 0.4]]) == 0.0          # maxpro_utils.rs:75-77
    assert oracle.maximin_criterion([[0.3, 0.4]]) == math.inf    # maximin_utils.rs:57
    with pytest.raises(ValueError):
        oracle.maxpro_criterion(np.empty((0, 2)))


def test_generate_lhd():  # lhd.rs:134-155
    x = oracle.generate_lhd(100, 4, 12345)
    assert x.shape == (100, 4) and lhd_bins_ok(x)
    assert ((x > 0) & (x < 1)).all()


def test_generate_lhd_shapes_and_streams():
    for n, d in [(1, 1), (2, 3), (7, 2), (50, 3), (501, 2), (4096, 1)]:
        assert lhd_bins_ok(oracle.generate_lhd(n, d, 7))
    a, b = oracle.generate_lhd(50, 3, 42), oracle.generate_lhd(50, 3, 43)
    assert not np.array_equal(a, b)
    assert np.array_equal(a, oracle.generate_lhd(50, 3, 42))
    assert oracle.generate_lhd(0, 4, 12345).shape == (0, 4)  # lhd.rs:157-166
    with pytest.raises(ValueError):
        oracle.generate_lhd(10, 0, 12345)                     # lhd.rs:168-177


def test_generate_lhd_is_uniform_permutation():
    # every stratum equally likely in every row position: chi-square-ish sanity on 4000 designs
    n, trials = 8, 4000
    counts = np.zeros((n, n))
    for s in range(trials):
        x = oracle.generate_lhd(n, 1, 1000 + s)
        counts[np.arange(n), np.floor(x[:, 0] * n).astype(int)] += 1
    expected = trials / n
    assert np.abs(counts - expected).max() < 5 * math.sqrt(expected)


@pytest.fixture(params=[1, 2], ids=["LHD-Philox-v1", "LHD-mix-v2"])
def stream_version(request):
    prev = oracle.get_stream_version()
    oracle.set_stream_version(request.param)
    yield request.param
    oracle.set_stream_version(prev)


def _chi2_ok(counts, dof, sigmas=5.0):
    expected = counts.sum() / counts.size
    chi2 = float(((counts - expected) ** 2 / expected).sum())
    return chi2 < dof + sigmas * math.sqrt(2.0 * dof), chi2


def test_design_stream_statistics(stream_version):
    """The engine's own design streams (the reference pins none): 40,000 consecutive candidate streams
    seed + i, as build_lhd draws them, must look like independent uniform Latin hypercubes."""
    n, d, N = 50, 3, 40_000
    x = oracle.generate_batch(n, d, 20240, 0, N)                  # [N, n, d]
    strata = np.floor(x * n).astype(np.int64)
    assert (np.sort(strata, axis=1) == np.arange(n)[None, :, None]).all()      # every column a permutation (lhd.rs:134-155)
    for k in range(d):
        # (row, stratum) table: every stratum equally likely at every row position
        table = np.zeros((n, n))
        np.add.at(table, (np.repeat(np.arange(n), N), strata[:, :, k].T.ravel()), 1.0)
        ok, chi2 = _chi2_ok(table, (n - 1) * (n - 1))
        assert ok, (k, chi2)
        # orderings of adjacent rows: P(stratum[i] < stratum[i+1]) = 1/2 for every i
        less = (strata[:, :-1, k] < strata[:, 1:, k]).sum(axis=0)
        z = (less - N / 2) / math.sqrt(N / 4)
        assert np.abs(z).max() < 5.0 and float((z ** 2).sum()) < (n - 1) + 5 * math.sqrt(2.0 * (n - 1)), (k, z)
        # the differences of adjacent rows are those of a uniform permutation: P(|a - b| = t) = 2 (n - t) / (n (n - 1))
        diff = np.abs(strata[:, :-1, k] - strata[:, 1:, k]).ravel()
        obs = np.bincount(diff, minlength=n)[1:]
        exp = diff.size * 2.0 * (n - np.arange(1, n)) / (n * (n - 1))
        chi2 = float(((obs - exp) ** 2 / exp).sum())
        assert chi2 < (n - 2) + 5 * math.sqrt(2.0 * (n - 2)), (k, chi2)
    # neighbouring candidate streams (seed + i, seed + i + 1) and neighbouring columns are independent
    for a, b in ((strata[:-1, 0, 0], strata[1:, 0, 0]), (strata[:, 0, 0], strata[:, 0, 1]), (strata[:, 7, 1], strata[:, 7, 2])):
        table = np.zeros((n, n))
        np.add.at(table, (a, b), 1.0)
        ok, chi2 = _chi2_ok(table, (n - 1) * (n - 1))
        assert ok, chi2
    # jitter: uniform inside its stratum (Kolmogorov-Smirnov on 10^6 values) and independent of the stratum
    u = (x * n - strata).ravel()[:1_000_000]
    assert 0.0 < u.min() and u.max() < 1.0
    us = np.sort(u)
    ks = max(np.max(np.arange(1, us.size + 1) / us.size - us), np.max(us - np.arange(us.size) / us.size))
    assert ks * math.sqrt(us.size) < 1.95, ks                                   # p ~ 1e-3
    corr = np.corrcoef(u, strata.ravel()[:1_000_000].astype(np.float64))[0, 1]
    assert abs(corr) < 5.0 / math.sqrt(u.size), corr
    # consecutive jitters of one column are not correlated either (a Weyl sequence feeds the v2 mixer)
    uu = (x[:, :, 0] * n - strata[:, :, 0])
    c1 = np.corrcoef(uu[:, :-1].ravel(), uu[:, 1:].ravel())[0, 1]
    assert abs(c1) < 5.0 / math.sqrt(uu[:, 1:].size), c1


def test_design_stream_small_permutations_are_equally_likely(stream_version):
    """n = 4: all 24 permutations of a column equally likely over 48,000 streams (chi-square, 23 degrees of freedom)."""
    n, N = 4, 48_000
    x = oracle.generate_batch(n, 2, 99, 0, N)
    strata = np.floor(x * n).astype(np.int64)
    for k in range(2):
        code = (strata[:, :, k] * (n ** np.arange(n))[None, :]).sum(axis=1)
        _, counts = np.unique(code, return_counts=True)
        assert counts.size == 24
        ok, chi2 = _chi2_ok(counts.astype(np.float64), 23)
        assert ok, chi2


def test_criteria_well_conditioned():  # maxpro_utils.rs:101-115, maximin_utils.rs:69-83 (fewer rounds)
    for s in range(50):
        x = oracle.generate_lhd(100, 5, 12345 + s)
        for v in (oracle.maxpro_criterion(x), oracle.maximin_criterion(x)):
            assert 0.0 <= v < math.inf


def test_score_batch_tie_rule():
    """src/lib.rs:85-97: equal scores keep the LATER index, for both directions."""
    rng = np.random.default_rng(3)
    base = np.stack([random_lhd(rng, 12, 3) for _ in range(6)])
    for metric in (oracle.MAXPRO, oracle.MAXIMIN):
        scores, arg = oracle.score_batch(base, metric)
        best = int(np.argmin(scores) if metric == oracle.MAXPRO else np.argmax(scores))
        assert arg == best
        other = (best + 1) % 6
        dup = np.concatenate([base, base[best:best + 1], base[other:other + 1]])
        scores2, arg2 = oracle.score_batch(dup, metric)
        assert scores2[6] == scores2[best] and arg2 == 6


def test_build_lhd_matches_definition():
    """lib.rs:65-98: candidate i is generate_lhd(seed+i); winner = fold with the tie rule."""
    n, d, iters, seed = 10, 2, 200, 42
    for metric in (oracle.MAXPRO, oracle.MAXIMIN):
        cands = np.stack([oracle.generate_lhd(n, d, seed + i) for i in range(iters)])
        scores, arg = oracle.score_batch(cands, metric)
        design, score, idx = oracle.build_lhd(n, d, iters, metric, seed)
        assert idx == arg and score == scores[arg]
        assert np.array_equal(design, cands[arg])
    none_design, _, _ = oracle.build_lhd(n, d, iters, None, seed)
    assert np.array_equal(none_design, oracle.generate_lhd(n, d, seed))  # lib.rs:46-50


def test_search_range_sharding_invariance():
    n, d, seed, total = 10, 3, 7, 600
    for metric in (oracle.MAXPRO, oracle.MAXIMIN):
        whole = oracle.search_range(n, d, metric, seed, 0, total)
        parts = [oracle.search_range(n, d, metric, seed, a, b) for a, b in [(0, 100), (100, 350), (350, 600)]]
        sc = np.array([p[0] for p in parts])
        best = parts[int(np.argmin(sc) if metric == oracle.MAXPRO else np.argmax(sc))]
        assert whole == best


def test_anneal_never_worse_and_deterministic():
    x = oracle.generate_lhd(15, 3, 5)
    for metric, minimize in ((oracle.MAXPRO, True), (oracle.MAXIMIN, False)):
        crit = oracle.maxpro_criterion if metric == oracle.MAXPRO else oracle.maximin_criterion
        for swap in (True, False):
            out, best = oracle.anneal_lhd(x, 400, 1.0, 0.99, metric, minimize, 11, swap)
            assert crit(out) == best
            assert best <= crit(x) if minimize else best >= crit(x)
            out2, best2 = oracle.anneal_lhd(x, 400, 1.0, 0.99, metric, minimize, 11, swap)
            assert np.array_equal(out, out2) and best == best2
            if swap:  # swap moves keep every column a permutation (README.md:8-9)
                assert lhd_bins_ok(out)


def test_order_empty_and_single():  # order.rs:140-149
    out, perm = oracle.order_design(np.empty((0, 3)), oracle.MAXPRO, True)
    assert out.shape == (0, 3)
    out, perm = oracle.order_design(np.array([[0.2, 0.7]]), oracle.MAXPRO, True)
    assert perm.tolist() == [0]


def test_ordered_criteria_parity():  # order.rs:107-138 (fewer rounds, smaller n)
    for s in range(5):
        x = oracle.generate_lhd(30, 5, 12345 + s)
        out, perm = oracle.order_design(x, oracle.MAXIMIN, False)
        assert sorted(perm.tolist()) == list(range(30))
        assert abs(oracle.maximin_criterion(x) - oracle.maximin_criterion(out)) < 1e-10
        assert abs(oracle.maxpro_criterion(x) - oracle.maxpro_criterion(out)) < 1e-10


def test_order_incremental_equals_faithful():
    """The O(n^2 d) restatement used to check large designs picks the same points as order.rs."""
    for s, (n, d) in enumerate([(12, 2), (25, 3), (40, 5)]):
        x = oracle.generate_lhd(n, d, 99 + s)
        for metric, minimize in ((oracle.MAXPRO, True), (oracle.MAXIMIN, False), (oracle.MAXIMIN, True), (oracle.MAXPRO, False)):
            _, p_ref = oracle.order_design(x, metric, minimize)
            _, p_inc = oracle.order_design(x, metric, minimize, incremental=True)
            assert p_ref.tolist() == p_inc.tolist()
    # centre point first (order.rs:49-59)
    x = oracle.generate_lhd(20, 2, 1)
    _, p = oracle.order_design(x, oracle.MAXPRO, True)
    assert p[0] == int(np.argmin(np.sqrt(((x - 0.5) ** 2).sum(axis=1))))
