"""SURVEY 8(f) rank 3 -- InsideDomain / TargetReached (logic.py:343-387) rest on matplotlib's Path.contains_points, a
third-party routine that is absent here ("parity unpinned"): the oracle restates its crossing-number rule and is checked
against known answers and an independent winding-number implementation on points away from the edges.  CPU only."""
import numpy as np

from crowddynamics_b200 import synthetic as S
from oracle import crowd_oracle as O

L_SHAPE = np.array([(0, 0), (6, 0), (6, 2), (2, 2), (2, 6), (0, 6)], dtype=np.float64)


def _winding(poly, p):
    x, y = p
    w = 0
    for (x0, y0), (x1, y1) in zip(poly, np.roll(poly, -1, 0)):
        if y0 <= y:
            if y1 > y and (x1 - x0) * (y - y0) - (x - x0) * (y1 - y0) > 0:
                w += 1
        elif y1 <= y and (x1 - x0) * (y - y0) - (x - x0) * (y1 - y0) < 0:
            w -= 1
    return w != 0


def test_known_answers():
    sq = [(0, 0), (4, 0), (4, 4), (0, 4), (0, 0)]                # closed ring, as np.asarray(polygon.exterior) gives it
    assert O.point_in_polygon(sq, 2, 2) and O.point_in_polygon(sq, 3.999, 0.001)
    for p in ((5, 2), (-1, 2), (2, 5), (2, -1), (4.001, 4.001)):
        assert not O.point_in_polygon(sq, *p)
    assert O.point_in_polygon(sq[:-1], 2, 2)                     # the closing vertex is optional
    assert O.point_in_polygon(sq[::-1], 2, 2)                    # orientation does not matter
    assert O.point_in_polygon(L_SHAPE, 1, 5) and O.point_in_polygon(L_SHAPE, 5, 1) and not O.point_in_polygon(L_SHAPE, 4, 4)
    assert not O.point_in_polygon([(0, 0), (1, 1)], 0.5, 0.5)    # fewer than 3 vertices: empty


def test_against_winding_number():
    rng = np.random.default_rng(0)
    star = np.array([(np.cos(a) * r, np.sin(a) * r) for a, r in zip(np.linspace(0, 2 * np.pi, 14, endpoint=False), [3, 1] * 7)])
    for poly in (L_SHAPE, star, rng.uniform(-3, 3, (3, 2))):
        pts = rng.uniform(poly.min() - 1, poly.max() + 1, (4000, 2))
        got = np.array([O.point_in_polygon(poly, *p) for p in pts])
        assert (got == np.array([_winding(poly, p) for p in pts])).all()
        assert 0 < got.sum() < len(pts)


def test_inside_domain_and_target_reached_semantics():
    agents, _, side = S.uniform_crowd(400, 'three_circle', density=1.0, seed=2)
    domain = np.array([(0, 0), (side * 0.7, 0), (side * 0.7, side), (0, side)])
    assert agents['active'].all()                                # synthetic crowds start active
    inside = np.array([O.point_in_polygon(domain, *p) for p in agents['position']])
    assert O.inside_domain(agents, domain) == (~inside).sum() > 0    # np.sum(change), logic.py:354-357
    assert (agents['active'] == inside).all()
    assert O.inside_domain(agents, domain) == 0                  # nothing moved
    agents['position'][:10, 0] += side                           # ten agents walk out
    assert O.inside_domain(agents, domain) == inside[:10].sum()
    reached = np.zeros(len(agents), dtype=bool)
    goal = np.array([(0, 0), (3, 0), (3, 3), (0, 3)])
    c0 = O.target_reached(agents, goal, reached)
    assert c0 == reached.sum() > 0
    agents['position'] += 100.0                                  # everybody leaves: reached_by is sticky
    assert O.target_reached(agents, goal, reached) == c0
