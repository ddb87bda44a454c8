"""Host-side pieces that need no GPU: the logic-tree surface mirrored from the reference (simulation/base.py:27-84,
tests simulation/tests/test_base.py:15-22), the data contract, the synthetic workloads."""
import numpy as np
import pytest

from crowddynamics_b200 import synthetic as S
from crowddynamics_b200.exceptions import InvalidType, CrowdDynamicsException
from crowddynamics_b200.logic import LogicNodeBase, post_order_iter, pre_order_iter
from crowddynamics_b200.structures import (agent_type_circular, agent_type_three_circle, obstacle_type_linear, as_obstacles,
                                           model_of, is_model, MODEL_CIRCULAR, MODEL_THREE_CIRCLE)


class Node(LogicNodeBase):
    def __init__(self, name):
        super().__init__(name=name)
        self.calls = 0

    def update(self):
        self.calls += 1


def test_tree_composition_matches_reference_semantics():
    # examples/simulations.py:123-136: `<<` is left associative and returns self
    reset, inside, integ, fluct, adj, nav, ori, aa, ao = (Node(n) for n in (
        'Reset', 'InsideDomain', 'Integrator', 'Fluctuation', 'Adjusting', 'Navigation', 'Orientation',
        'AgentAgentInteractions', 'AgentObstacleInteractions'))
    tree = reset << inside << (integ << (fluct, adj << (nav, ori), aa, ao))
    assert tree is reset and reset.children == (inside, integ)
    order = [n.name for n in post_order_iter(reset)]
    assert order == ['InsideDomain', 'Fluctuation', 'Navigation', 'Orientation', 'Adjusting', 'AgentAgentInteractions',
                     'AgentObstacleInteractions', 'Integrator', 'Reset']
    assert [n.name for n in pre_order_iter(reset)][0] == 'Reset'
    assert reset['Adjusting'] is adj and nav.root is reset
    with pytest.raises(KeyError):
        reset['Nope']
    for n in post_order_iter(reset):
        n.update()
    assert all(n.calls == 1 for n in pre_order_iter(reset))


def test_inject_before_and_after():
    a, b, c, d = Node('a'), Node('b'), Node('c'), Node('d')
    a << (b, c)
    b.inject_before(d)          # d takes b's place, b becomes d's child
    assert d.parent is a and b.parent is d
    e = Node('e')
    a.inject_after(e)           # e adopts a's children and hangs under a
    assert e.parent is a and a.children == (e,) and set(e.children) == {c, d}


def test_data_contract():
    assert agent_type_circular.itemsize == 228 and agent_type_three_circle.itemsize == 316
    assert agent_type_circular.fields['position'][1] == 92 and agent_type_circular.fields['radius'][1] == 28
    assert agent_type_three_circle.fields['position'][1] == 124 and agent_type_three_circle.fields['orientation'][1] == 260
    assert agent_type_three_circle.fields['std_rand_torque'][1] == 308
    a = np.zeros(3, dtype=agent_type_circular)
    assert model_of(a) == MODEL_CIRCULAR and model_of(np.zeros(1, dtype=agent_type_three_circle)) == MODEL_THREE_CIRCLE
    assert is_model(a, 'circular') and not is_model(a, 'three_circle')
    with pytest.raises(InvalidType):
        model_of(np.zeros(2, dtype=obstacle_type_linear))
    with pytest.raises(InvalidType):
        model_of(np.zeros(2))
    assert issubclass(InvalidType, CrowdDynamicsException)
    obs = S.walls_of_box(0, 0, 2, 3)
    assert as_obstacles(obs).shape == (4, 4) and as_obstacles(None).shape == (0, 4)
    assert (as_obstacles(obs) == as_obstacles(as_obstacles(obs))).all()


def test_synthetic_workloads_are_seeded_and_sane():
    a1, o1, s1 = S.uniform_crowd(500, 'three_circle', density=1.0, seed=3)
    a2, o2, s2 = S.uniform_crowd(500, 'three_circle', density=1.0, seed=3)
    assert (a1.view(np.uint8) == a2.view(np.uint8)).all() and s1 == s2
    # non-overlapping jittered lattice, adult body ranges (conf/body_types.cfg)
    d = np.hypot(*(a1['position'][:, None] - a1['position'][None]).transpose(2, 0, 1)) + np.eye(500) * 10
    assert d.min() > 2 * 0.29
    assert (a1['radius'] >= 0.22).all() and (a1['radius'] <= 0.29).all() and (a1['mass'] >= 65.5).all()
    off = np.stack((np.sin(a1['orientation']), -np.cos(a1['orientation'])), 1) * a1['r_ts'][:, None]
    assert np.allclose(a1['position_ls'], a1['position'] - off)
    ag, ob, fields = S.hallway(seed=0)
    assert len(ag) == 50 and len(ob) == 2 and len(fields) == 2 and set(ag['target']) == {0, 1}
    mg, (U, V) = fields[1]
    assert U.shape == mg.shape == (51, 401) and (U == 1).all() and (V == 0).all()
    assert mg.indicer(np.array([[0.05, 0.19], [39.99, 4.99]])).tolist() == [[0, 1], [399, 49]]
    ar, obr, fr, side = S.room_with_exit(300, 'circular')
    assert len(obr) == 11 and (ar['target'] == 0).all()


# ---- strips: the step schedule of kept block lists is pure host logic (crowddynamics_b200/parallel.py) -------------------------
class _KindRecorder:
    def __init__(self):
        self.kinds = []

    def set_kind(self, kind):
        self.kinds.append(kind)


def _bare_strip(skin, interval):
    from crowddynamics_b200.parallel import StripSimulation
    sim = StripSimulation.__new__(StripSimulation)
    sim.dev = _KindRecorder()
    sim.skin, sim.max_interval, sim.interval = skin, 16, interval
    sim._since, sim._steps, sim._force_rebuild = 0, 0, True
    return sim


def test_strip_step_schedule():
    """cdb_strip_set_kind: 4 = rebuild + migrants, 1 = rebuild, 2 = kept, 3 = kept + migrants (the step before a rebuild)."""
    sim = _bare_strip(0.0, 1)                      # no kept lists at all: classic steps, nothing announced to the device
    for _ in range(3):
        kind, migrate = sim.plan_step()
        assert (kind, migrate) == (0, True)
        sim.end_step(migrate)
    assert sim.dev.kinds == []
    sim = _bare_strip(0.1, 1)                      # interval 1: every step rebuilds and migrates, drift bookkeeping running
    seq = []
    for _ in range(3):
        kind, migrate = sim.plan_step(); seq.append((kind, migrate)); sim.end_step(migrate)
    assert seq == [(4, True)] * 3
    sim = _bare_strip(0.1, 4)
    seq = []
    for _ in range(9):
        kind, migrate = sim.plan_step(); seq.append(kind); sim.end_step(migrate)
    assert seq == [1, 2, 2, 3, 1, 2, 2, 3, 1] and sim.dev.kinds == seq
    # the interval shrinks in the middle of a run of kept steps: the next step is the last one on this list
    sim = _bare_strip(0.1, 8)
    for _ in range(3):
        kind, migrate = sim.plan_step(); sim.end_step(migrate)
    sim.adapt_interval(disp_max=0.05, limit=0.21)          # floor(0.21 / 0.075) = 2
    assert sim.interval == 2
    kind, migrate = sim.plan_step()
    assert (kind, migrate) == (3, True)
    sim.end_step(migrate)
    assert sim.plan_step() == (1, False)
    # adapt cadence and degenerate inputs
    sim = _bare_strip(0.1, 1)
    due = []
    for k in range(1, 70):
        sim._steps = k
        if sim.adapt_due():
            due.append(k)
    assert due == [2, 32, 64]
    for bad in (0.0, float('nan'), float('inf')):
        sim.adapt_interval(bad, 0.21)
        assert sim.interval == 1
    sim.adapt_interval(1e-6, 0.21)
    assert sim.interval == 16                       # capped at max_interval


def test_host_round_trips_runs_every_crowd_and_surfaces_errors():
    """engine.host_round_trips (bench.py's e2e leg): one host thread per crowd, every crowd gets its n updates of
    upload -> step -> download in that order, an exception in one crowd's thread is re-raised in the caller."""
    import threading
    from crowddynamics_b200.engine import host_round_trips

    class FakeDevice:
        def __init__(self, fail_at=None):
            self.log, self.threads, self.fail_at = [], set(), fail_at

        def upload_raw(self, ptr, n):
            self.threads.add(threading.get_ident())
            self.log.append(('up', ptr, n))

        def step(self, k, flags, cell_size, dt_min, dt_max, want_dt=True):
            if self.fail_at is not None and len(self.log) // 3 == self.fail_at:
                raise CrowdDynamicsException('device error')
            self.log.append(('step', k, flags, cell_size, dt_min, dt_max, want_dt))

        def download_raw(self, ptr, n):
            self.log.append(('down', ptr, n))

    devs = [FakeDevice() for _ in range(3)]
    host_round_trips([(d, 1000 + k, 10 * (k + 1)) for k, d in enumerate(devs)], 4, flags=0x7f, cell_size=3.6, dt_min=0.001, dt_max=0.01)
    for k, d in enumerate(devs):
        assert len(d.threads) == 1 and threading.get_ident() not in d.threads
        assert d.log == [('up', 1000 + k, 10 * (k + 1)), ('step', 1, 0x7f, 3.6, 0.001, 0.01, False), ('down', 1000 + k, 10 * (k + 1))] * 4
    bad = FakeDevice(fail_at=2)
    good = FakeDevice()
    with pytest.raises(CrowdDynamicsException):
        host_round_trips([(good, 1, 5), (bad, 2, 5)], 4)
    assert len(good.log) == 12 and len(bad.log) == 7      # the healthy crowd finished, the failing one stopped at its third update
