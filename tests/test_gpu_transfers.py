"""Field-masked PCIe transfers between the packed host records and the device planes (csrc/transfer_kernels.cuh,
cdb_host_register / cdb_upload_agents_fields / cdb_download_agents_aos with a mask) and the strict-mode protocol built on
them ("upload dirty fields -> kernel -> download written fields")."""
import numpy as np
import pytest

from conftest import rel_err_fields
from crowddynamics_b200 import _lib, logic as L, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR, MODEL_THREE_CIRCLE

pytestmark = pytest.mark.gpu
CELL = 3.6
MODELS = ['circular', 'three_circle']
MASKS = [_lib.F_POSITION, _lib.F_FORCE | _lib.F_TORQUE, _lib.F_POSITION | _lib.F_VELOCITY | _lib.F_FORCE_PREV | _lib.F_SHOULDERS |
         _lib.F_ORIENTATION | _lib.F_ANGULAR_VELOCITY | _lib.F_TORQUE_PREV, _lib.F_ALL_MUTABLE]
FIELDS = {'position': _lib.F_POSITION, 'velocity': _lib.F_VELOCITY, 'target_direction': _lib.F_TARGET_DIRECTION, 'force': _lib.F_FORCE,
          'force_prev': _lib.F_FORCE_PREV, 'position_ls': _lib.F_SHOULDERS, 'position_rs': _lib.F_SHOULDERS,
          'orientation': _lib.F_ORIENTATION, 'angular_velocity': _lib.F_ANGULAR_VELOCITY, 'target_orientation': _lib.F_TARGET_ORIENTATION,
          'torque': _lib.F_TORQUE, 'torque_prev': _lib.F_TORQUE_PREV}


def _mid(model):
    return MODEL_CIRCULAR if model == 'circular' else MODEL_THREE_CIRCLE


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('registered', [False, True])
def test_masked_download_writes_only_the_selected_fields(model, registered):
    agents, obstacles, side = S.uniform_crowd(5000, model, density=1.0, seed=51, overlap_fraction=0.02)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(7, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, CELL, 0.001, 0.01, want_dt=False)     # the planes are re-sorted by now
    full = agents.copy()
    dev.download(full)
    for mask in MASKS:
        got = agents.copy()
        got['active'] = False                   # bytes no mask covers must survive
        if registered:
            dev.host_register(got)
        dev.transfer_stats(reset=True)
        dev.download(got, mask)
        up, down = dev.transfer_stats()
        if registered:
            dev.host_unregister(got)
        for name in agents.dtype.names:
            bit = FIELDS.get(name, 0)
            want = full[name] if bit & mask else (np.zeros_like(got[name]) if name == 'active' else agents[name])
            assert (got[name] == want).all(), (name, hex(mask))
        nbytes = sum(agents.dtype[n].itemsize for n in agents.dtype.names if FIELDS.get(n, 0) & mask)
        if registered:
            assert up == 0 and down == nbytes * len(agents)          # exactly the selected bytes crossed PCIe
        else:
            assert down == agents.dtype.itemsize * len(agents)       # pageable memory: whole records through the bounce buffer
    dev.close()


@pytest.mark.parametrize('model', MODELS)
@pytest.mark.parametrize('registered', [False, True])
def test_masked_upload_into_a_resorted_state(model, registered):
    agents, obstacles, side = S.uniform_crowd(5000, model, density=1.0, seed=52)
    rng = np.random.default_rng(0)
    dev = DeviceAgents(_mid(model))
    dev.upload(agents)
    dev.set_obstacles(obstacles)
    dev.step(5, _lib.STEP_ALL & ~_lib.STEP_NAVIGATION, CELL, 0.001, 0.01, want_dt=False)
    host = agents.copy()
    dev.download(host)
    if registered:
        dev.host_register(host)
    # a host-side node rewrites the velocities and (three-circle) the torque; everything else on the device stays as it is
    host['velocity'] = rng.normal(size=host['velocity'].shape)
    mask = _lib.F_VELOCITY
    if model == 'three_circle':
        host['torque'] = rng.normal(size=len(host))
        mask |= _lib.F_TORQUE
    poison = host.copy()
    poison['position'] += 100.0                 # NOT in the mask: must not reach the device
    dev.transfer_stats(reset=True)
    if registered:
        saved = host['position'].copy()
        host['position'] += 100.0
        dev.upload_fields(host, mask)
        host['position'] = saved
    else:
        dev.upload_fields(poison, mask)
    up, _ = dev.transfer_stats()
    back = agents.copy()
    dev.download(back)
    if registered:
        dev.host_unregister(host)
    dev.close()
    assert rel_err_fields(back, host)[0] == 0
    nbytes = 16 + (8 if model == 'three_circle' else 0)
    assert up == (nbytes if registered else agents.dtype.itemsize) * len(agents)


def test_upload_fields_needs_the_same_crowd():
    from crowddynamics_b200.exceptions import CrowdDynamicsException
    agents, _, _ = S.uniform_crowd(100, 'circular', density=1.0, seed=53)
    dev = DeviceAgents(MODEL_CIRCULAR)
    dev.upload(agents)
    with pytest.raises(CrowdDynamicsException):
        dev.upload_fields(agents[:50].copy(), _lib.F_VELOCITY)
    dev.upload_fields(agents, 0)                # nothing selected: a no-op
    dev.close()


@pytest.mark.parametrize('model', MODELS)
def test_strict_dirty_tree_equals_strict_always_tree(model):
    """The reference's Hallway tree, node by node, for 20 updates with a host-side node in the middle that kicks the agents:
    uploading only the declared dirty fields gives the same arrays as re-sending everything before every node."""
    out = {}
    for policy in ('always', 'dirty'):
        agents, obstacles, fields = S.hallway(seed=3, model=model)
        sim = L.MultiAgentSimulation(agents, obstacles, fields)
        sim.logic = L.hallway_logic(sim, mode='strict')
        sim.logic.state.upload_policy = policy

        class Kick:
            def __init__(self):
                self.k = 0

            def update(self):
                self.k += 1
                sim.agents.array['force'][:, 1] += 5.0 * np.sin(0.3 * self.k)

        sim.logic['Adjusting'].inject_after(L.HostNode(sim, node=Kick(), reads=_lib.F_FORCE, writes=_lib.F_FORCE))
        moved = []
        for _ in range(20):
            sim.logic.state.dev and sim.logic.state.dev.transfer_stats(reset=True)
            sim.update()
            moved.append(sim.logic.state.dev.transfer_stats())
        out[policy] = (agents.copy(), moved)
    assert rel_err_fields(out['dirty'][0], out['always'][0])[0] == 0
    item = out['dirty'][0].dtype.itemsize
    n = len(out['dirty'][0])
    up_d, down_d = out['dirty'][1][-1]
    up_a, down_a = out['always'][1][-1]
    assert up_d == 16 * n                        # steady state: only the kicked force goes up ...
    assert up_a >= 5 * item * n                  # ... where 'always' re-sends the whole array per GPU node
    assert down_d < down_a


@pytest.mark.parametrize('model', MODELS)
def test_host_round_trips_of_independent_crowds_equal_serial_runs(model):
    """engine.host_round_trips: crowds stepped through the host boundary concurrently (one host thread and CUDA stream per
    crowd; bench.py's `e2e` leg) end with the bytes they end with when run one after the other."""
    import torch
    from crowddynamics_b200.engine import host_round_trips
    flags = _lib.STEP_ALL & ~_lib.STEP_NAVIGATION
    crowds, hosts, refs = [], [], []
    for k in range(3):
        agents, obstacles, side = S.uniform_crowd(20000 + 1000 * k, model, density=1.0, seed=70 + k, overlap_fraction=0.01)
        ref = agents.copy()
        dev = DeviceAgents(_mid(model))
        dev.set_obstacles(obstacles)
        for _ in range(4):
            dev.upload(ref); dev.step(1, flags, CELL, 0.001, 0.01, want_dt=False); dev.download(ref, _lib.F_WHOLE_RECORD)
        dev.close()
        refs.append(ref)
        host = torch.empty(agents.nbytes, dtype=torch.uint8).pin_memory()
        host.numpy()[:] = agents.view(np.uint8).reshape(-1)
        dev = DeviceAgents(_mid(model))
        dev.set_obstacles(obstacles)
        crowds.append((dev, host.data_ptr(), len(agents)))
        hosts.append(host)
    host_round_trips(crowds, 4, flags, CELL, 0.001, 0.01)
    for (dev, _, _), host, ref in zip(crowds, hosts, refs):
        dev.close()
        assert np.array_equal(host.numpy(), ref.view(np.uint8).reshape(-1))
