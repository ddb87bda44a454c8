"""Navigation-field construction on the device (cdb_build_navigation_field, csrc/field_kernels.cuh) against the CPU oracle
(oracle/field_oracle.*: fast marching + the reference's direction_map / obstacle_handling), against closed-form distances,
and through the nodes that consume it."""
import numpy as np
import pytest

from crowddynamics_b200 import _lib, logic as L, synthetic as S
from crowddynamics_b200.engine import DeviceAgents
from crowddynamics_b200.structures import MODEL_CIRCULAR
from oracle import crowd_oracle as O, field_oracle as F

pytestmark = pytest.mark.gpu

ROOM = dict(bounds=(0.0, 0.0, 12.0, 9.0), target=[(12.0, 3.5, 12.0, 5.5)],
            walls=[(0.0, 0.0, 12.0, 0.0), (0.0, 9.0, 12.0, 9.0), (0.0, 0.0, 0.0, 9.0), (12.0, 0.0, 12.0, 3.5), (12.0, 5.5, 12.0, 9.0),
                   (6.0, 0.0, 6.0, 5.0), (9.0, 9.0, 9.0, 4.0), (2.0, 6.5, 4.5, 6.5)])
CASES = {
    'free_point': dict(bounds=(0.0, 0.0, 10.0, 8.0), target=[(5.0, 4.0, 5.0, 4.0)], walls=None),
    'free_line': dict(bounds=(-3.0, 1.0, 7.0, 6.0), target=[(7.0, 2.0, 7.0, 4.0)], walls=None),
    'room_with_inner_walls': ROOM,
    'diagonal_walls': dict(bounds=(0.0, 0.0, 9.0, 9.0), target=[(0.0, 4.0, 0.0, 5.0), (4.0, 9.0, 5.0, 9.0)],
                           walls=[(2.0, 1.0, 7.0, 6.0), (1.0, 7.5, 4.0, 5.2)]),
}


def _build(case, step, radius=0.5, strength=0.3):
    dev = DeviceAgents(MODEL_CIRCULAR)
    walls = None if case['walls'] is None else np.array(case['walls'], dtype=np.float64)
    mg, dmap, (U, V) = dev.build_navigation_field(0, case['target'], walls, case['bounds'], step, radius, strength, want_maps=True)
    rounds = dev.last_field_rounds
    dev.close()
    return mg, dmap, U, V, rounds


@pytest.mark.parametrize('name', list(CASES))
@pytest.mark.parametrize('step', [0.1, 0.05])
def test_field_matches_the_oracle(name, step):
    case = CASES[name]
    mg, dmap, U, V, rounds = _build(case, step)
    d_ref, (U_ref, V_ref) = F.navigation_to_target(case['target'], case['walls'], case['bounds'], step, 0.5, 0.3)
    assert dmap.shape == d_ref.shape == mg.shape
    assert (np.isnan(dmap) == np.isnan(d_ref)).all()
    ok = np.isfinite(d_ref)
    # the fast iterative method reaches the fixed point fast marching computes (same update, same arithmetic)
    assert np.abs(dmap[ok] - d_ref[ok]).max() <= 1e-12
    assert (np.isnan(U) == np.isnan(U_ref)).all() and (np.isnan(V) == np.isnan(V_ref)).all()
    okv = np.isfinite(U_ref)
    # directions are normalised differences of nearly equal distances: 1e-12 in the distance is ~1e-10 in the direction
    assert np.abs(U[okv] - U_ref[okv]).max() <= 1e-9 and np.abs(V[okv] - V_ref[okv]).max() <= 1e-9
    assert rounds > 0


def test_closed_form_distance_and_first_order_convergence():
    errs = []
    for step in (0.2, 0.1, 0.05):
        case = dict(bounds=(0.0, 0.0, 10.0, 10.0), target=[(5.0, 5.0, 5.0, 5.0)], walls=None)
        mg, dmap, U, V, _ = _build(case, step)
        ny, nx = dmap.shape
        X, Y = np.meshgrid(step * np.arange(nx), step * np.arange(ny))
        iy, ix = np.argwhere(dmap > 0)[0]
        exact = np.hypot(X - ix * step, Y - iy * step)
        errs.append(np.abs(-dmap - exact).max())
        far = exact > 1.0
        # the field points at the target
        assert (U[far] * (ix * step - X[far]) + V[far] * (iy * step - Y[far]) > 0.95 * exact[far]).all()
    assert errs[0] <= 1.5 * 0.2 and errs[0] > errs[1] > errs[2] and errs[2] < 0.6 * errs[0]


def test_built_field_drives_navigation_like_an_uploaded_one():
    """cdb_build_navigation_field installs the field on the device; sampling it must equal sampling the same maps uploaded
    from the host, and the Navigation node of a tree with geometry targets builds it on first use."""
    step = 0.1
    mg, dmap, U, V, _ = _build(ROOM, step)
    agents, _, _ = S.uniform_crowd(3000, 'circular', density=40.0, seed=61)      # 3000 agents in ~8.7 x 8.7 m
    agents['position'] = agents['position'] * (8.0 / agents['position'].max()) + 0.4
    agents['target'] = 0
    walls = np.array(ROOM['walls'])
    ref = agents.copy()
    O.navigation(ref, [(mg, (U, V))])
    field = L.MultiAgentSimulation.GeometryField(walls, [ROOM['target']], ROOM['bounds'])
    sim = L.MultiAgentSimulation(agents, field=field)
    sim.logic = L.Navigation(sim, step=step, radius=0.5, strength=0.3, mode='strict')
    sim.update()
    changed = (ref['target_direction'] != S.uniform_crowd(3000, 'circular', density=40.0, seed=61)[0]['target_direction']).any(axis=1)
    assert changed.mean() > 0.9
    same = (agents['target_direction'] == ref['target_direction']) | (np.isnan(agents['target_direction']) & np.isnan(ref['target_direction']))
    assert same.all()


def test_large_grid_fixed_point_property():
    """4 M cells (a 200 m room at 0.1 m): no oracle run -- the result must be a fixed point of the upwind update everywhere
    (which is what 'converged' means) and within first-order error of the closed form away from the walls' shadow."""
    step, side = 0.1, 200.0
    case = dict(bounds=(0.0, 0.0, side, side), target=[(side, 95.0, side, 105.0)], walls=[(100.0, 0.0, 100.0, 120.0)])
    mg, dmap, U, V, rounds = _build(case, step)
    T = np.abs(dmap)
    inf = np.where(np.isnan(T), np.inf, T)
    a = np.minimum(np.pad(inf, ((0, 0), (1, 0)), constant_values=np.inf)[:, :-1], np.pad(inf, ((0, 0), (0, 1)), constant_values=np.inf)[:, 1:])
    b = np.minimum(np.pad(inf, ((1, 0), (0, 0)), constant_values=np.inf)[:-1, :], np.pad(inf, ((0, 1), (0, 0)), constant_values=np.inf)[1:, :])
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    with np.errstate(invalid='ignore'):
        upd = np.where(np.isinf(hi) | (hi - lo >= step), lo + step, (a + b + np.sqrt(np.maximum(2 * step * step - (a - b) ** 2, 0.0))) / 2)
    free = np.isfinite(T) & (T > step)          # beyond the frozen band next to the target
    assert np.abs(upd[free] - T[free]).max() <= 1e-12
    ny, nx = T.shape
    X, Y = np.meshgrid(step * np.arange(nx), step * np.arange(ny))
    right = (X > 101.0) & np.isfinite(T)        # the half with a free line of sight to the door
    exact = np.hypot(X - side, Y - np.clip(Y, 95.0, 105.0))
    assert np.abs(T[right] - exact[right]).max() <= 0.6
    assert rounds < 4000
