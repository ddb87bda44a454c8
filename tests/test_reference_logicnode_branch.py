"""The drop-in nodes subclass the reference's own ``LogicNode`` when the crowddynamics package is importable
(crowddynamics_b200/logic.py: ``_HAVE_REFERENCE``).  traitlets / anytree are not installed in this image, so that branch
never ran; here the REAL ``simulation/base.py`` and ``LogicNode`` (logic.py:31-54) of /root/reference are loaded on top of
minimal fakes of those two libraries (tests/fake_reference_env.py) and a whole tree is driven through it, with the device
replaced by a call recorder.  Needs the reference tree: skipped where /root/reference does not exist (the GPU box)."""
import importlib
import os
import sys

import numpy as np
import pytest

import fake_reference_env as env

pytestmark = pytest.mark.skipif(not os.path.isdir(env.REF), reason='reference tree not present')


@pytest.fixture
def ref_logic(monkeypatch):
    from test_host_state_protocol import Recorder
    base, ref = env.install(monkeypatch)
    import crowddynamics_b200.logic as L
    L = importlib.reload(L)
    assert L._HAVE_REFERENCE and L._RefLogicNode is ref.LogicNode
    Recorder.instances.clear()
    monkeypatch.setattr(L, 'DeviceAgents', Recorder)
    yield L, base, ref, Recorder
    monkeypatch.undo()
    importlib.reload(sys.modules['crowddynamics_b200.logic'])        # back to the duck-typed base for the other tests


def test_nodes_are_reference_logic_nodes_and_compose_like_them(ref_logic):
    L, base, ref, Recorder = ref_logic
    from crowddynamics_b200 import synthetic as S
    agents, obstacles, fields = S.hallway(seed=0)
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    tree = L.hallway_logic(sim, mode='resident', dt_min=0.001, dt_max=0.01)
    # isinstance of the reference classes; name trait defaults to the class name (base.py:12-16)
    for node in env.PreOrderIter(tree.root):
        assert isinstance(node, ref.LogicNode) and isinstance(node, base.LogicNodeBase) and isinstance(node, env.HasTraits)
        assert node.name == type(node).__name__ and repr(node) == node.name
        assert node.simulation is sim
    # composition through the reference's own __lshift__ / NodeMixin, lookup through its __getitem__ (base.py:52-84)
    assert [n.name for n in env.PostOrderIter(tree.root)] == ['Navigation', 'Orientation', 'Adjusting', 'AgentAgentInteractions',
                                                               'AgentObstacleInteractions', 'Integrator', 'Reset']
    assert tree['Integrator'].dt_min == 0.001 and tree['Integrator'].dt_max == 0.01      # our params survive the traits __init__
    with pytest.raises(KeyError, match='not in the tree'):
        tree['Nope']
    # inject_before / inject_after are the reference's (base.py:35-45)
    fl = L.Fluctuation(sim, seed=3)
    tree['Adjusting'].inject_before(fl)
    assert fl.parent is tree['Integrator'] and tree['Adjusting'].parent is fl
    # the reference's driver loop (multiagent.py:51-55): post-order over logic.root
    for _ in range(3):
        for node in env.PostOrderIter(tree.root):
            node.update()
    dev, = Recorder.instances
    names = dev.names()
    assert names.count('upload') == 1 and names.count('integrate') == 3 and names.count('fluctuation') == 3
    assert names.index('navigation') < names.index('adjust') < names.index('fluctuation') < names.index('integrate') < names.index('reset')
    assert sim.data['dt'] == 0.01


def test_unknown_keyword_reaches_the_reference_constructor(ref_logic):
    """Anything that is not one of our parameters is passed on to HasTraits.__init__, like for a reference node."""
    L, base, ref, Recorder = ref_logic
    from crowddynamics_b200 import synthetic as S
    agents, obstacles, fields = S.hallway(seed=0)
    sim = L.MultiAgentSimulation(agents, obstacles, fields)
    node = L.Integrator(sim, name='MyIntegrator', dt_max=0.02)
    assert node.name == 'MyIntegrator' and node.dt_max == 0.02
    with pytest.raises(TypeError):
        L.Integrator(sim, no_such_trait=1)


def test_exceptions_are_the_reference_classes(ref_logic):
    import crowddynamics.exceptions as ref_exc
    import crowddynamics_b200.exceptions as ours
    ours = importlib.reload(ours)
    try:
        assert ours.CrowdDynamicsException is ref_exc.CrowdDynamicsException and ours.InvalidType is ref_exc.InvalidType
    finally:
        sys.modules.pop('crowddynamics.exceptions', None)
        importlib.reload(ours)
