"""Two arithmetic identities the kernels rely on, checked on the CPU (the kernels themselves are pinned bit for bit by the GPU
parity tests; these tests say WHY the rewrites are exact).

* k_prep.cu, arm walk: "max over B, G, R of |a - b| < tau" is evaluated as a SWAR byte test on vabsdiff4(a, b):
  no byte exceeds tau - 1  <=>  ((x + 0x01010101 * (128 - tau)) | x) & 0x80808080 == 0   (byte 3 of x is 0, tau <= 128).
* k_post.cu, region voting: the CSR offsets of the parked votes (exclusive sum of the low-vote counts) and the start of the slice
  that leaks into each high-vote outlier (exclusive running max of the offsets at high-vote outliers) come from ONE scan with the
  operator (sum_a, last_a) o (sum_b, last_b) = (sum_a + sum_b, last_b >= 0 ? sum_a + last_b : last_a)."""
import numpy as np


def test_swar_any_byte_exceeds():
    # every value of every byte matters only through the byte itself and a possible carry into the next one:
    # all 2^16 pairs of adjacent bytes x all positions, plus random triples
    b = np.arange(256, dtype=np.uint32)
    lo, hi = np.meshgrid(b, b, indexing="ij")
    words = [lo | (hi << 8), (lo << 8) | (hi << 16), lo | (hi << 16)]
    rng = np.random.default_rng(7)
    t = rng.integers(0, 256, (1 << 20, 3)).astype(np.uint32)
    words.append(t[:, 0] | (t[:, 1] << 8) | (t[:, 2] << 16))
    for tau in (3, 6, 12, 15, 20, 128):
        c = np.uint32((0x01010101 * (128 - tau)) & 0xFFFFFFFF)
        for w in words:
            w = w.ravel().astype(np.uint32)
            got = ((((w + c) & np.uint32(0xFFFFFFFF)) | w) & np.uint32(0x80808080)) == 0
            mx = np.maximum(np.maximum(w & 0xFF, (w >> 8) & 0xFF), (w >> 16) & 0xFF)
            assert np.array_equal(got, mx < tau), tau


HIGH = -1  # element marker of a high-vote outlier in this restatement


def _comb(a, b):
    return (a[0] + b[0], a[0] + b[1] if b[1] >= 0 else a[1])


def _elem(e):
    return (0, 0) if e == HIGH else (e, -1)


def test_vote_scan_operator_equals_sum_scan_plus_max_scan():
    rng = np.random.default_rng(11)
    for trial in range(200):
        n = int(rng.integers(1, 400))
        kind = rng.integers(0, 4, n)  # 0, 1: nothing; 2: low-vote outlier; 3: high-vote outlier
        elems = [HIGH if k == 3 else (int(rng.integers(1, 21)) if k == 2 else 0) for k in kind]
        # the two-scan definition the kernels used before (k_vote_mark + exclusive max scan)
        low = np.array([0 if e == HIGH else e for e in elems])
        off = np.concatenate(([0], np.cumsum(low)[:-1]))
        mark = np.array([off[i] if elems[i] == HIGH else 0 for i in range(n)])
        start = np.concatenate(([0], np.maximum.accumulate(mark)[:-1]))
        # one left-to-right scan with the operator
        run = (0, -1)
        for i, e in enumerate(elems):
            assert run[0] == off[i] and max(run[1], 0) == start[i], (trial, i)
            run = _comb(run, _elem(e))
        # associativity: any bracketing gives the same total (the kernels combine per thread, per warp, per block)
        vals = [_elem(e) for e in elems]
        while len(vals) > 1:
            j = int(rng.integers(0, len(vals) - 1))
            vals[j:j + 2] = [_comb(vals[j], vals[j + 1])]
        assert vals[0] == run
