"""SURVEY 8(f) rank 4 -- the C oracle's exit detection / herding / leader-follower restatement against golden vectors
produced by the reference's own numba code (tests/golden/generate.py: gen_collective_fixture).  CPU only.

Bar: integer outputs (neighbour tables, targets, leaders, detected exits, flags) bit-exact; directions 1e-13 (libm hypot)."""
import numpy as np
import pytest

from conftest import load_golden, from_raw
from crowddynamics_b200.structures import agent_type_circular, agent_type_three_circle, obstacle_type_linear
from oracle import crowd_oracle as O

DT = {'circular': agent_type_circular, 'three_circle': agent_type_three_circle}
TOL = 1e-13


def _load(model):
    g = load_golden('collective_%s.npz' % model)
    agents = from_raw(g['initial'], DT[model])
    obstacles = np.ascontiguousarray(g['obstacles']).view(obstacle_type_linear).reshape(-1)
    return g, agents, obstacles


def test_line_intersect_known_answers():
    """geom2D.py:38-59 (reference test core/tests/test_geom2D.py:19-30 checks it against shapely's intersects)."""
    assert O.line_intersect((0, 0), (1, 1), (0, 1), (1, 0))
    assert not O.line_intersect((0, 0), (1, 1), (2, 0), (3, 1))           # parallel: d == 0
    assert not O.line_intersect((0, 0), (1, 0), (2, -1), (2, 1))          # beyond the end of the first segment
    assert O.line_intersect((0, 0), (1, 0), (1, -1), (1, 1))              # touching an end point counts (<= 1)
    assert not O.line_intersect((0, 0), (0, 0), (1, -1), (1, 1))          # zero-length: d == 0
    obstacles = np.zeros(2, dtype=obstacle_type_linear)
    obstacles[0]['p0'], obstacles[0]['p1'] = (5, 0), (6, 0)
    obstacles[1]['p0'], obstacles[1]['p1'] = (0.5, -1), (0.5, 1)
    assert O.is_obstacle_between_points((0, 0), (1, 0), obstacles)
    assert not O.is_obstacle_between_points((0, 0), (0.4, 0), obstacles)
    assert not O.is_obstacle_between_points((0, 0), (1, 0), obstacles[:0])


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_herding_relationship(model):
    g, _, _ = _load(model)
    for r, expect in zip(g['relationship_inputs'], g['relationship']):
        for phi, e in zip((np.pi / 2, float(g['phi'])), expect):
            assert O.herding_relationship(r[0:2], r[2:4], r[4:6], r[6:8], phi) == tuple(bool(x) for x in e)


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_exit_detection(model):
    g, agents, obstacles = _load(model)
    detected, has = O.exit_detection(g['center_door'], agents, obstacles, float(g['detection_range']))
    assert (detected == g['detected_exit']).all() and (has == g['has_detected']).all()
    assert 0 < has.sum() < len(agents)                                     # both outcomes are exercised
    # nothing in range / no doors
    d2, h2 = O.exit_detection(g['center_door'], agents, obstacles, 0.0)
    assert (d2 == -1).all() and not h2.any()
    d3, h3 = O.exit_detection(np.zeros((0, 2)), agents, obstacles, 20.0)
    assert (d3 == -1).all() and not h3.any()


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_nearest_neighbors_and_herding(model):
    g, agents, obstacles = _load(model)
    k = int(g['size_nearest_other'])
    nbr = O.find_nearest_neighbors(agents, float(g['sight']), k, obstacles)
    assert (nbr == g['neighbors']).all()
    # independent property: the rows are the k nearest visible agents closer than `sight`
    pos = agents['position']
    rng = np.random.default_rng(0)
    for i in rng.choice(len(agents), 40, replace=False):
        d = np.hypot(*(pos - pos[i]).T)
        cand = [j for j in np.argsort(d) if j != i and d[j] < float(g['sight'])
                and not O.is_obstacle_between_points(pos[i], pos[j], obstacles)][:k]
        assert set(cand) == set(nbr[i][nbr[i] >= 0].tolist())
    direction, has = O.herding_interaction(agents, agents['is_follower'], nbr, 0.15, float(g['phi']))
    assert (has == g['herding_has_direction']).all()
    assert np.abs(direction - g['herding_direction']).max() <= TOL
    assert not has[~agents['is_follower']].any()


@pytest.mark.parametrize('model', ['circular', 'three_circle'])
def test_leader_follower(model):
    g, agents, obstacles = _load(model)
    a = agents.copy()
    d = O.leader_follower_with_herding_interaction(a, obstacles, float(g['sight']), int(g['size_nearest_other']))
    ref = from_raw(g['lfh_after'], DT[model])
    assert (a['target'] == ref['target']).all() and (a['index_leader'] == ref['index_leader']).all()
    assert a.tobytes() == ref.tobytes()
    assert np.abs(d - g['lfh_direction']).max() <= TOL
    assert (a['target'] != agents['target']).any() and (a['index_leader'] != agents['index_leader']).any()
    a = agents.copy()
    d = O.leader_follower_interaction(a, obstacles, 20.0)
    ref = from_raw(g['lf_after'], DT[model])
    assert a.tobytes() == ref.tobytes()
    assert np.abs(d - g['lf_direction']).max() <= TOL
    # leaders are never touched
    lead = agents['is_leader']
    assert (a['target'][lead] == agents['target'][lead]).all()


def test_no_leaders_no_followers():
    g, agents, obstacles = _load('circular')
    a = agents.copy()
    a['is_leader'] = False
    d = O.leader_follower_interaction(a, obstacles, 20.0)
    assert (d == 0).all()
    f = a['is_follower']
    # no visible leader: the remembered leader's target is inherited, the rest go to their familiar exit
    remembered = f & (agents['index_leader'] != -1)
    assert (a['target'][remembered] == agents['target'][agents['index_leader'][remembered]]).all()
    lost = f & ~remembered
    assert (a['target'][lost] == agents['familiar_exit'][lost]).all()
    b = agents.copy()
    b['is_follower'] = False
    d = O.leader_follower_with_herding_interaction(b, obstacles, 10.0, 5)
    assert (d == 0).all() and b.tobytes() == _as_no_followers(agents).tobytes()


def _as_no_followers(agents):
    b = agents.copy()
    b['is_follower'] = False
    return b
