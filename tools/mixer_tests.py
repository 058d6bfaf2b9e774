"""Statistical comparison of candidate element mixers for the LHD-mix v2 design stream (maxpro_b200/csrc/philox.cuh).
M9 is the one shipped; M0 the two-round finaliser it replaced.  usage: python tools/mixer_tests.py [M0 M9 ...]"""
import numpy as np, math, sys
M32 = np.uint64(0xFFFFFFFF)
def brev32(x):
    x = ((x >> np.uint64(1)) & np.uint64(0x55555555)) | ((x & np.uint64(0x55555555)) << np.uint64(1))
    x = ((x >> np.uint64(2)) & np.uint64(0x33333333)) | ((x & np.uint64(0x33333333)) << np.uint64(2))
    x = ((x >> np.uint64(4)) & np.uint64(0x0F0F0F0F)) | ((x & np.uint64(0x0F0F0F0F)) << np.uint64(4))
    x = ((x >> np.uint64(8)) & np.uint64(0x00FF00FF)) | ((x & np.uint64(0x00FF00FF)) << np.uint64(8))
    x = ((x >> np.uint64(16)) | (x << np.uint64(16))) & M32
    return x
def rot(x, r): return ((x << np.uint64(r)) | (x >> np.uint64(32 - r))) & M32
def mul(x, c): return (x * np.uint64(c)) & M32
def xs(x, r): return x ^ (x >> np.uint64(r))
MIX = {
 'M0': lambda x: xs(mul(xs(mul(xs(x,16),0x21f0aaad),15),0x735a2d97),15),
 'M1': lambda x: mul(brev32(mul(x,0x21f0aaad)),0x735a2d97),
 'M4': lambda x: mul(xs(mul(xs(x,16),0x21f0aaad),15),0x735a2d97),
 'M5': lambda x: mul(xs(mul(x,0x21f0aaad),15),0x735a2d97),          # mul, xorshift, mul: 4 instrs
 'M6': lambda x: mul(rot(mul(rot(mul(x,0x21f0aaad),16),0x735a2d97),16),0x9E3779B1),  # 5 instrs
 'M7': lambda x: mul(brev32(mul(xs(x,16),0x21f0aaad)),0x735a2d97),  # xs, mul, brev, mul: 5
 'M8': lambda x: xs(mul(brev32(mul(x,0x21f0aaad)),0x735a2d97),16),  # mul brev mul xs: 5
 'M9': lambda x: mul(xs(x,16),0x735a2d97),                          # xorshift, mul: 3 (shipped)
 'M10': lambda x: mul(xs(mul(x,0x21f0aaad),16),0x735a2d97),         # mul, xorshift, mul: 4
}
def gen(mix, N, n, d, rng):
    # random A, B per (candidate, column) stand in for the Philox keys
    A = rng.integers(0, 2**32, size=(N, d), dtype=np.uint64); B = rng.integers(0, 2**32, size=(N, d), dtype=np.uint64) | np.uint64(1)
    perm = np.zeros((N, n, d), dtype=np.int64); jit = np.zeros((N, n, d), dtype=np.uint64)
    a = np.zeros((N, d, n), dtype=np.int64)
    idxN = np.arange(N)[:, None]; idxd = np.arange(d)[None, :]
    for i in range(n):
        w = mix((A + np.uint64(i) * B) & M32)
        P = w * np.uint64(i + 1)
        j = (P >> np.uint64(32)).astype(np.int64); jt = P & M32
        a[idxN, idxd, i] = a[idxN, idxd, j]
        a[idxN, idxd, j] = i
        jit[:, i, :] = jt   # jitter of stratum i
    # a[.., pos] = stratum at row pos
    strata = np.transpose(a, (0, 2, 1))
    # jitter per row = jitter of its stratum
    u = np.take_along_axis(jit, strata, axis=1).astype(np.float64) / 2**32
    return strata, u
def chi2ok(counts, dof, sig=5.0):
    e = counts.sum() / counts.size; c = float(((counts - e) ** 2 / e).sum()); return c, dof + sig * math.sqrt(2 * dof)
def battery(name, N=40000, n=50, d=3, seed=1):
    rng = np.random.default_rng(seed)
    strata, u = gen(MIX[name], N, n, d, rng)
    res = []
    for k in range(d):
        t = np.zeros((n, n)); np.add.at(t, (np.repeat(np.arange(n), N), strata[:, :, k].T.ravel()), 1.0)
        res.append(('pos', k) + chi2ok(t, (n - 1) ** 2))
        less = (strata[:, :-1, k] < strata[:, 1:, k]).sum(axis=0); z = (less - N / 2) / math.sqrt(N / 4)
        res.append(('adjmax', k, float(np.abs(z).max()), 5.0)); res.append(('adjchi', k, float((z ** 2).sum()), (n - 1) + 5 * math.sqrt(2 * (n - 1))))
        diff = np.abs(strata[:, :-1, k] - strata[:, 1:, k]).ravel(); obs = np.bincount(diff, minlength=n)[1:]
        exp = diff.size * 2.0 * (n - np.arange(1, n)) / (n * (n - 1)); res.append(('diff', k, float(((obs - exp) ** 2 / exp).sum()), (n - 2) + 5 * math.sqrt(2 * (n - 2))))
        # triples: P(a<b<c) = 1/6 for rows (i,i+1,i+2)
        s = strata[:, :, k]; inc = ((s[:, :-2] < s[:, 1:-1]) & (s[:, 1:-1] < s[:, 2:])).sum(axis=0); z3 = (inc - N / 6) / math.sqrt(N * 5 / 36)
        res.append(('tripmax', k, float(np.abs(z3).max()), 5.0))
        # the draw index j_i distribution itself for large i: recover from pairs? skip
    uu = u[:, :, 0]; us = np.sort(uu.ravel()[:1000000]); ks = max(np.max(np.arange(1, us.size + 1) / us.size - us), np.max(us - np.arange(us.size) / us.size))
    res.append(('ks', 0, ks * math.sqrt(us.size), 1.95))
    c1 = np.corrcoef(uu[:, :-1].ravel(), uu[:, 1:].ravel())[0, 1]; res.append(('ucorr', 0, abs(c1) * math.sqrt(uu[:, 1:].size), 5.0))
    cs = np.corrcoef(uu.ravel(), strata[:, :, 0].ravel().astype(float))[0, 1]; res.append(('us_corr', 0, abs(cs) * math.sqrt(uu.size), 5.0))
    bad = [r for r in res if r[2] > r[3]]
    print(name, 'FAIL' if bad else 'ok', 'worst ratio %.2f' % max(r[2] / r[3] for r in res), bad[:4])
def avalanche(name, trials=200000):
    rng = np.random.default_rng(5); x = rng.integers(0, 2**32, size=trials, dtype=np.uint64); m = MIX[name]
    y = m(x); worst = 0
    for b in range(32):
        yb = m(x ^ np.uint64(1 << b)); f = y ^ yb
        for ob in (31, 30, 28, 24, 16, 8, 0):
            p = float(((f >> np.uint64(ob)) & np.uint64(1)).mean()); worst = max(worst, abs(p - 0.5))
    print(name, 'avalanche worst |p-0.5| over sampled output bits: %.3f' % worst)
def serial(name, ncol=400000, n=50, bits=6):
    rng = np.random.default_rng(11); m = MIX[name]
    A = rng.integers(0, 2**32, size=ncol, dtype=np.uint64); B = rng.integers(0, 2**32, size=ncol, dtype=np.uint64) | np.uint64(1)
    W = np.stack([m((A + np.uint64(i) * B) & M32) for i in range(n)], axis=1)   # [ncol, n]
    K = 1 << bits; worst = 0; bad = []
    for lag in (1, 2, 3, 4, 8, 16):
        for sh_a, sh_b in ((32 - bits, 32 - bits), (32 - bits, 20), (20, 20), (32 - bits, 12)):
            a = ((W[:, :-lag] >> np.uint64(sh_a)) & np.uint64(K - 1)).ravel().astype(np.int64)
            b = ((W[:, lag:] >> np.uint64(sh_b)) & np.uint64(K - 1)).ravel().astype(np.int64)
            t = np.bincount(a * K + b, minlength=K * K).astype(float); e = t.sum() / t.size
            chi = ((t - e) ** 2 / e).sum(); dof = K * K - 1; z = (chi - dof) / math.sqrt(2 * dof)
            worst = max(worst, z)
            if z > 5: bad.append((lag, sh_a, sh_b, round(z, 1)))
    # triple (w_i, w_{i+1}, w_{i+2}) top 4 bits each
    t3 = (((W[:, :-2] >> np.uint64(28)) << np.uint64(8)) | ((W[:, 1:-1] >> np.uint64(28)) << np.uint64(4)) | (W[:, 2:] >> np.uint64(28))).ravel().astype(np.int64)
    t = np.bincount(t3, minlength=4096).astype(float); e = t.sum() / 4096; z3 = (((t - e) ** 2 / e).sum() - 4095) / math.sqrt(2 * 4095)
    # FY-index pairs at i=48,49 (what the shuffle consumes): j = hi(w*(i+1))
    j48 = ((W[:, 48] * np.uint64(49)) >> np.uint64(32)).astype(np.int64); j49 = ((W[:, 49] * np.uint64(50)) >> np.uint64(32)).astype(np.int64)
    t = np.bincount(j48 * 50 + j49, minlength=49 * 50)[:49 * 50].astype(float); e = t.sum() / t.size; zj = (((t - e) ** 2 / e).sum() - (t.size - 1)) / math.sqrt(2 * (t.size - 1))
    print(f"{name}: worst pair z {worst:.1f}, triple z {z3:.1f}, (j48,j49) z {zj:.1f}", 'FAIL' if bad or z3 > 5 or zj > 5 else 'ok', bad[:6])
if __name__ == '__main__':
    for nm in sys.argv[1:] or MIX:
        battery(nm); avalanche(nm); serial(nm)

