"""The closed form of the reference's in-place partition (model/bvh.rs:419-430) that the device builder uses
(bvhtracer_b200/csrc/build_kernels.cu), checked against the sequential loop itself."""
import numpy as np


def sequential_partition(is_left):
    """`while i <= j { if left(a[i]) { i += 1 } else { swap(a[i], a[j]); j -= 1 } }` on the identity arrangement"""
    a = list(range(len(is_left)))
    i, j = 0, len(a) - 1
    while i <= j:
        if is_left[a[i]]:
            i += 1
        else:
            a[i], a[j] = a[j], a[i]
            j -= 1
    return a, i


def closed_form(is_left):
    """front = positions < q, plus q when it holds a right element; H_k = k-th right element of the front (ascending);
    G_k = k-th left element of the back (descending):
        front left -> stays, back right -> p - 1, G_k -> pos(H_k), H_k -> pos(G_{k-1}) - 1 (H_0 -> last)"""
    is_left = np.asarray(is_left, bool)
    n = len(is_left)
    n_left = int(is_left.sum())
    q, last = n_left, n - 1
    pos = np.arange(n)
    front = (pos < q) | ((pos == q) & ~is_left)
    h = pos[front & ~is_left]
    g = pos[~front & is_left][::-1]
    dest = np.empty(n, int)
    dest[front & is_left] = pos[front & is_left]
    dest[~front & ~is_left] = pos[~front & ~is_left] - 1
    for k, p in enumerate(g):
        dest[p] = h[k]
    for k, p in enumerate(h):
        dest[p] = last if k == 0 else g[k - 1] - 1
    out = np.empty(n, int)
    out[dest] = pos
    return list(out), n_left


def test_closed_form_equals_sequential_loop():
    rng = np.random.default_rng(0)
    for _ in range(20000):
        n = int(rng.integers(1, 48))
        p = rng.choice([0.0, 0.05, 0.3, 0.5, 0.8, 0.95, 1.0])
        flags = rng.random(n) < p
        a, i = sequential_partition(flags)
        b, n_left = closed_form(flags)
        assert a == b and i == n_left, (flags, a, b)


def test_closed_form_exhaustive_small():
    for n in range(1, 13):
        for bits in range(1 << n):
            flags = [(bits >> k) & 1 == 1 for k in range(n)]
            assert sequential_partition(flags)[0] == closed_form(flags)[0]
