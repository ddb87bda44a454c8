"""Host logic of the drop-in nodes without a GPU: which transfers ``DeviceState`` issues in strict and resident mode
(strict = upload -> kernel -> download of what the node writes; resident = upload once, refresh on demand).
``DeviceAgents`` is replaced by a recorder, so no library call is made -- this checks the protocol, not the kernels."""
import numpy as np
import pytest

from crowddynamics_b200 import synthetic as S, logic as L, _lib


class Recorder:
    """Stands in for engine.DeviceAgents: records the calls, returns plausible values."""
    instances = []

    def __init__(self, model, capacity=0, device=0, stream=None):
        self.model, self.n, self.calls = model, 0, []
        Recorder.instances.append(self)

    def _rec(self, name, *a):
        self.calls.append((name,) + a)

    def upload(self, agents):
        self.n = len(agents)
        self._rec('upload')

    def upload_fields(self, agents, mask):
        self._rec('upload_fields', mask)

    def host_register(self, agents):
        self._rec('host_register')

    def download(self, agents, mask=_lib.F_ALL_MUTABLE):
        self._rec('download', mask)

    def set_states(self, agents, target=True):
        self._rec('set_states', target)

    def get_states(self, agents):
        self._rec('get_states')

    def set_active(self, active):
        self._rec('set_active')

    def get_active(self):
        self._rec('get_active')
        return np.ones(self.n, dtype=bool)

    def inside_domain(self, want_count=True):
        self._rec('inside_domain')
        return 3

    def target_reached(self, n_polygons, want_counts=True):
        self._rec('target_reached', n_polygons)
        return np.arange(n_polygons) + 7

    def integrate(self, dt_min, dt_max):
        self._rec('integrate')
        return dt_max

    def step(self, n_steps, flags, cell_size, dt_min, dt_max, want_dt=True):
        self._rec('step', n_steps)
        return np.full(n_steps, dt_max)

    def __getattr__(self, name):            # every other node entry point: just record it
        if name.startswith('_'):
            raise AttributeError(name)
        return lambda *a, **k: self._rec(name)

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture
def recorder(monkeypatch):
    Recorder.instances.clear()
    monkeypatch.setattr(L, 'DeviceAgents', Recorder)
    return Recorder


def _sim(mode, upload=None):
    agents, obstacles, doors, side = S.leader_follower_crowd(50, 'circular', seed=1)
    domain = np.array([(0, 0), (side, 0), (side, side), (0, side)])
    sim = L.MultiAgentSimulation(agents, obstacles, (), domain=domain)
    sim.logic = L.Reset(sim, mode=mode, upload=upload) << (
        L.InsideDomain(sim), L.TargetReached(sim, polygons=[domain, None, domain * 0.5]),
        L.Integrator(sim) << (L.Adjusting(sim) << (L.ExitDetection(sim, center_door=doors) << L.LeaderFollowerWithHerding(sim)),
                              L.AgentAgentInteractions(sim)))
    return sim


def test_strict_mode_moves_exactly_what_each_node_touches(recorder):
    sim = _sim('strict', upload='always')
    assert sim.data['inactive'] == 0 and sim.data['target_0'] == 0 and sim.data['target_2'] == 0 and 'target_1' not in sim.data
    sim.update()
    dev, = recorder.instances
    names = dev.names()
    order = [n.name for n in L.post_order_iter(sim.logic.root)]
    assert order == ['InsideDomain', 'TargetReached', 'LeaderFollowerWithHerding', 'ExitDetection', 'Adjusting',
                     'AgentAgentInteractions', 'Integrator', 'Reset']
    assert names.count('upload') == len(order)                         # every node starts from the host array
    # InsideDomain: polygon once, active in, count, active out
    i = names.index('inside_domain')
    assert names[i - 1] == 'set_active' and names[i + 1] == 'get_active' and 'set_polygons' in names[:i]
    assert sim.data['inactive'] == 3
    assert sim.data['target_0'] == 7 and sim.data['target_2'] == 8      # two measured polygons, named by their index
    # the steering nodes send the States fields and fetch back the ones they mutate
    j = names.index('leader_follower_with_herding')
    assert names[j - 2:j] == ['set_states', 'set_obstacles'] and ('download', _lib.F_TARGET_DIRECTION) in dev.calls[j:j + 2]
    assert names[j + 1:j + 3] == ['download', 'get_states']
    k = names.index('exit_detection')
    assert names[k + 1] == 'get_states'                                 # nothing but States fields to publish
    assert ('set_states', False) in dev.calls                            # target travels with the records
    sim.update()
    assert dev.names().count('set_polygons') == 2                       # domain + targets, sent once each
    assert sim.data['inactive'] == 6 and sim.data['iterations'] == 2


def test_strict_mode_uploads_only_dirty_fields(recorder):
    """strict = upload dirty fields -> kernel -> download written fields.  The first node sends the whole array
    (constants stay on the device afterwards, keyed on the array identity), the following ones nothing -- until a host-side
    node declares what it wrote."""
    sim = _sim('strict')
    sim.update()
    dev, = recorder.instances
    names = dev.names()
    assert names.count('upload') == 1 and names.count('host_register') == 1 and 'upload_fields' not in names
    assert names.count('set_states') == 1 and names.count('set_active') == 1
    # every node still publishes exactly what it wrote
    assert ('download', _lib.F_FORCE | _lib.F_TORQUE) in dev.calls and names.count('download') >= 4
    st = sim.logic.state
    assert st.host_dirty == 0 and st.dev_ahead == 0
    # a host-side node (here: a stand-in for the reference's Fluctuation) writes force: only force travels up
    class HostFluctuation:
        def update(self_inner):
            sim.agents.array['force'] += 1.0
    sim.logic['Adjusting'].inject_after(L.HostNode(sim, node=HostFluctuation(), reads=0, writes=_lib.F_FORCE))
    n0 = len(dev.calls)
    sim.update()
    new = dev.calls[n0:]
    assert [c for c in new if c[0] == 'upload'] == [] and [c for c in new if c[0] == 'upload_fields'] == [('upload_fields', _lib.F_FORCE)]
    # the array object was replaced: everything is sent again
    sim.agents.array = sim.agents.array.copy()
    n0 = len(dev.calls)
    sim.update()
    assert [c[0] for c in dev.calls[n0:]].count('upload') == 1


def test_invalidate_never_overwrites_host_edits(recorder):
    """invalidate() used to download first, which overwrote the very edits it was meant to publish."""
    from crowddynamics_b200.exceptions import CrowdDynamicsException
    sim = _sim('resident')
    sim.update()
    dev, = recorder.instances
    st = sim.logic.state
    assert st.dev_ahead and st.dirty_states and st.dirty_active
    n0 = len(dev.calls)
    with pytest.raises(CrowdDynamicsException):
        st.invalidate()                      # the device is ahead: refusing beats silently losing either side
    assert len(dev.calls) == n0              # ... and nothing was downloaded over the host array
    # the supported ways: say what was edited (the rest of the pending fields is still fetched later) ...
    st.invalidate(_lib.F_VELOCITY)
    assert st.host_dirty == _lib.F_VELOCITY and not (st.dev_ahead & _lib.F_VELOCITY) and st.dev_ahead
    st.sync_host()
    assert ('download', st_mask_without(_lib.F_VELOCITY, dev)) in dev.calls
    # ... or sync first, edit, then invalidate everything
    st.invalidate()
    sim.update()
    assert dev.names().count('upload') == 2


def st_mask_without(mask, dev):
    last = [c for c in dev.calls if c[0] == 'download'][-1][1]
    assert not (last & mask)
    return last


def test_a_new_device_gets_walls_fields_and_states_again(recorder):
    """when the agent model changes a new DeviceAgents is created; the keys that say 'already sent' described the
    old device and must be dropped with it."""
    agents, obstacles, side = S.uniform_crowd(30, 'circular', density=1.0, seed=3)
    fields = [S.direction_field(0.5, (0, 0, side, side), 'swirl')]
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    sim.logic = L.FusedStep(sim, seed=11, fluctuation=True)
    sim.update()
    first, = recorder.instances
    assert {'set_obstacles', 'set_navigation_field', 'set_seed'} <= set(first.names())
    three, _, _ = S.uniform_crowd(30, 'three_circle', density=1.0, seed=3)
    sim.agents.array = three                 # another agent model: the state is rebuilt on a new device object
    sim.update()
    second = recorder.instances[-1]
    assert second is not first
    assert {'upload', 'set_obstacles', 'set_navigation_field', 'set_seed'} <= set(second.names())


def test_resident_mode_uploads_once_and_syncs_on_demand(recorder):
    sim = _sim('resident')
    for _ in range(3):
        sim.update()
    dev, = recorder.instances
    names = dev.names()
    assert names.count('upload') == 1 and names.count('set_states') == 1 and names.count('set_active') == 1
    assert 'download' not in names and 'get_states' not in names and 'get_active' not in names
    assert names.count('inside_domain') == 3 and sim.data['inactive'] == 9
    st = sim.logic.state
    assert st.dirty_host and st.dirty_states and st.dirty_active
    st.sync_host()
    assert dev.names()[-3:] == ['download', 'get_states', 'get_active']
    assert not (st.dirty_host or st.dirty_states or st.dirty_active)
    st.sync_host()
    assert dev.names()[-3:] == ['download', 'get_states', 'get_active']  # nothing new to fetch
    # a host-side node edited the array (after the sync above): everything is sent again at the next node
    st.invalidate()
    sim.update()
    names = dev.names()
    assert names.count('upload') == 2 and names.count('set_states') == 2 and names.count('set_active') == 2
    # switching a tree to strict publishes what is pending first
    L.DeviceState.of(sim, 'strict')
    assert st.mode == 'strict' and not st.dirty_host


def test_geometry_arguments_are_validated(recorder):
    from crowddynamics_b200.exceptions import InvalidValue
    agents, obstacles, doors, side = S.leader_follower_crowd(20, 'circular', seed=2)
    sim = L.MultiAgentSimulation(agents, obstacles, ())                  # no domain, integer targets
    with pytest.raises(InvalidValue):
        L.InsideDomain(sim)
    with pytest.raises(InvalidValue):
        L.InsideDomain(sim, domain=[(0, 0), (1, 1)])                     # not a polygon
    node = L.ExitDetection(sim)
    with pytest.raises(InvalidValue):
        node._doors()                                                    # field.targets are indices here, not geometries
    assert L.ExitDetection(sim, center_door=doors)._doors().shape == (2, 2)
    sim.field.targets = [np.array([(0.0, 1.0), (0.0, 3.0)]), np.array([(side, 1.0), (side, 3.0)])]
    assert np.allclose(L.ExitDetection(sim)._doors(), [(0.0, 2.0), (side, 2.0)])      # mean of the coordinates, logic.py:247-248
    tr = L.TargetReached(sim, polygons=[np.array([(0, 0), (1, 0), (1, 1), (0, 0)])])
    assert tr.names == ['target_0'] and sim.data['target_0'] == 0
    assert L.TargetReached(sim).names == []                              # the two door lines (2 vertices) are not polygons: skipped
