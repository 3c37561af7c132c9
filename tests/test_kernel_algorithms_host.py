"""Host-side (numpy) models of two algorithmic rewrites made in the CUDA kernels, checked against the plain definition
they replace.  The kernels themselves are compared on the GPU (tests/test_gpu_*.py); these pin the MATH on the CPU box.

* csrc/rays_sampling.cu:stable_rank -- rank of every element of cat(A, B) under a stable ascending sort, by binary search
  in the other run when both runs are sorted (instead of the all-pairs count).
* csrc/hashgrid.cu:encode_rays_bwd_tcl_v2_kernel -- a stencil tap adds what it owes to lattice points it SHARES with the
  centre sample's cell into the centre's eight accumulators (per-axis weights selected by the cell offset) and only
  scatters its remaining corners; must equal the plain "every plane scatters its own eight corners" definition.
"""
import numpy as np


def _stable_rank_all_pairs(v, i):
    x = v[i]
    return int(np.sum((v < x) | ((v == x) & (np.arange(len(v)) < i))))


def _stable_rank_bsearch(v, n, i):
    x = v[i]
    if i < n:
        return i + int(np.searchsorted(v[n:], x, side="left"))     # elements of B strictly below x
    return (i - n) + int(np.searchsorted(v[:n], x, side="right"))  # elements of A not above x


def test_stable_rank_by_binary_search_equals_all_pairs_count():
    rng = np.random.default_rng(0)
    for trial in range(200):
        n, nb = int(rng.integers(1, 40)), int(rng.integers(1, 20))
        # few distinct values -> many ties inside and between the runs
        a = np.sort(rng.integers(0, 12, n).astype(np.float32) * 0.25)
        b = np.sort(rng.integers(0, 12, nb).astype(np.float32) * 0.25)
        v = np.concatenate([a, b])
        ranks = [_stable_rank_bsearch(v, n, i) for i in range(n + nb)]
        assert ranks == [_stable_rank_all_pairs(v, i) for i in range(n + nb)]
        assert sorted(ranks) == list(range(n + nb))                 # a permutation: every slot written once
        out = np.empty_like(v)
        out[ranks] = v
        assert np.array_equal(out, np.sort(v, kind="stable"))


def _cell(scale, x):
    p = scale * x + 0.5
    g = np.floor(p).astype(np.int64)
    return g, (p - g)


def _corner_weights(w):
    out = np.empty(8)
    for c in range(8):
        out[c] = np.prod([w[k] if (c >> k) & 1 else 1.0 - w[k] for k in range(3)])
    return out


def test_lattice_sharing_scatter_equals_plain_per_plane_scatter():
    rng = np.random.default_rng(1)
    scale = 37.0
    for trial in range(300):
        x0 = rng.random(3)
        eps = rng.choice([0.002, 0.02, 0.06])                       # taps inside the cell / one cell over / further
        taps = [x0 + eps * np.array(k) for k in ((1, -1, -1), (-1, -1, 1), (-1, 1, -1), (1, 1, 1))]
        d = rng.standard_normal(5)                                  # one gradient value per plane (delta basis, plane 0 = sum)
        ref = {}
        # plain definition (delta basis): plane 0 scatters d0 * w0(c); tap i scatters d_i * w_i(c) to ITS corners and
        # -d_i * w0(c) to the centre's corners
        g0, w0 = _cell(scale, x0)
        cw0 = _corner_weights(w0)
        for c in range(8):
            key = tuple(g0 + [(c >> k) & 1 for k in range(3)])
            ref[key] = ref.get(key, 0.0) + d[0] * cw0[c]
        for i, xt in enumerate(taps, start=1):
            g, w = _cell(scale, xt)
            cw = _corner_weights(w)
            for c in range(8):
                key = tuple(g + [(c >> k) & 1 for k in range(3)])
                ref[key] = ref.get(key, 0.0) + d[i] * cw[c]
                key0 = tuple(g0 + [(c >> k) & 1 for k in range(3)])
                ref[key0] = ref.get(key0, 0.0) - d[i] * cw0[c]
        # v2 scheme
        got = {}
        agg = d[0] * cw0
        for i, xt in enumerate(taps, start=1):
            g, w = _cell(scale, xt)
            off = g - g0
            fl = np.where(off == 0, 1.0 - w, np.where(off == -1, w, 0.0))   # weight on the centre cell's lower plane
            fu = np.where(off == 0, w, np.where(off == 1, 1.0 - w, 0.0))    # ... upper plane
            for c in range(8):
                shared = np.prod([fu[k] if (c >> k) & 1 else fl[k] for k in range(3)])
                agg[c] += (shared - cw0[c]) * d[i]
            if np.any(off != 0):
                cw = _corner_weights(w)
                for c in range(8):
                    q = off + np.array([(c >> k) & 1 for k in range(3)])
                    if np.all((q == 0) | (q == 1)):
                        continue                                     # a lattice point of the centre's cell: already in agg
                    key = tuple(g + [(c >> k) & 1 for k in range(3)])
                    got[key] = got.get(key, 0.0) + d[i] * cw[c]
        for c in range(8):
            key = tuple(g0 + [(c >> k) & 1 for k in range(3)])
            got[key] = got.get(key, 0.0) + agg[c]
        assert set(got) <= set(ref)
        for key, val in ref.items():
            assert abs(got.get(key, 0.0) - val) < 1e-12, (trial, key)
