"""Invariants of the 16APSK / 32APSK tables of dvbs2rx_b200.apsk (CPU).

Nothing in the reference can pin these tables (it has no APSK demapper, lib/xfecframe_demapper_cb_impl.cc:70-72 throws)
and the standard's text is not available in this build environment: parity is UNPINNED.  What can be checked offline
are the structural properties EN 302 307-1 clause 5.4.3 / 5.4.4 constellations have:
  * 4 + 12 (+ 16) points on concentric rings, equally spaced in phase on each ring, unit average energy;
  * the ring radii ratios are the gamma values of the code rate;
  * the labelling is Gray along each ring (neighbours on a ring differ in exactly one bit) -- except the 16-point
    outer ring of 32APSK, which cannot be (4 + 12 + 16 with 5 bits): there neighbours differ in one or two bits;
  * mirroring the constellation in the I axis or in the Q axis maps it onto itself, and on the 4- and 12-point rings
    it flips exactly one label bit, the same one for every point (the two "sign" bits of the label): the label map is
    symmetric under the four quadrant reflections.  (The 16-point ring of 32APSK has points ON the axes, which are
    their own mirror images, so no bit can flip there.)"""
import numpy as np
import pytest

from dvbs2rx_b200 import apsk


def _rings(pts):
    r = np.hypot(pts[:, 0], pts[:, 1])
    order = np.argsort(r)
    rings, cur = [], [order[0]]
    for i in order[1:]:
        if abs(r[i] - r[cur[-1]]) < 1e-4:
            cur.append(i)
        else:
            rings.append(cur)
            cur = [i]
    rings.append(cur)
    return rings, r


CASES = [("16APSK " + k, apsk.points_16apsk(g), (g,), [4, 12]) for k, g in apsk.GAMMA_16APSK.items()] + \
        [("32APSK " + k, apsk.points_32apsk(*g), g, [4, 12, 16]) for k, g in apsk.GAMMA_32APSK.items()]


@pytest.mark.parametrize("name,pts,gammas,sizes", CASES, ids=[c[0] for c in CASES])
def test_ring_structure_energy_and_ratios(name, pts, gammas, sizes):
    pts = pts.astype(np.float64)
    assert abs((pts ** 2).sum(axis=1).mean() - 1.0) < 1e-6          # unit average energy
    rings, r = _rings(pts)
    assert [len(x) for x in rings] == sizes
    r1 = r[rings[0][0]]
    for ring, g in zip(rings[1:], gammas):
        assert abs(r[ring[0]] / r1 - g) < 1e-5                        # ring ratio of the code rate
    for ring in rings:
        ph = np.sort(np.mod(np.arctan2(pts[ring, 1], pts[ring, 0]), 2 * np.pi))
        gaps = np.diff(np.concatenate([ph, [ph[0] + 2 * np.pi]]))
        assert np.allclose(gaps, 2 * np.pi / len(ring), atol=1e-5)  # equally spaced


@pytest.mark.parametrize("name,pts,gammas,sizes", CASES[:1] + CASES[len(apsk.GAMMA_16APSK):len(apsk.GAMMA_16APSK) + 1],
                         ids=["16APSK", "32APSK"])
def test_gray_along_rings_and_reflection_symmetry(name, pts, gammas, sizes):
    pts = pts.astype(np.float64)
    rings, _ = _rings(pts)
    for ring in rings:
        ph = np.arctan2(pts[ring, 1], pts[ring, 0])
        around = [ring[i] for i in np.argsort(ph)]
        dist = [bin(int(a) ^ int(b)).count("1") for a, b in zip(around, around[1:] + around[:1])]
        if len(ring) == 16:  # 32APSK outer ring: quasi-Gray
            assert max(dist) <= 2 and dist.count(1) >= 8, (name, dist)
        else:
            assert dist == [1] * len(ring), (name, dist)  # Gray along the ring
    # mirror images: one label bit per axis, the same for every point
    def image(pt):
        d = np.hypot(pts[:, 0] - pt[0], pts[:, 1] - pt[1])
        i = int(np.argmin(d))
        assert d[i] < 1e-5
        return i
    inner = [i for ring in rings if len(ring) != 16 for i in ring]
    for flip in (np.array([1.0, -1.0]), np.array([-1.0, 1.0])):
        assert sorted(image(pts[i] * flip) for i in range(len(pts))) == list(range(len(pts)))  # onto itself
        masks = {int(i) ^ image(pts[i] * flip) for i in inner}
        assert len(masks) == 1 and bin(masks.pop()).count("1") == 1, (name, flip)
