"""Minimal stand-ins for the third-party modules the reference's ``simulation/base.py`` needs (traitlets, anytree, dateutil),
so that the REAL ``LogicNodeBase`` / ``LogicNode`` classes of /root/reference can be imported and the branch of
crowddynamics_b200/logic.py that subclasses them (``_HAVE_REFERENCE``) can be executed in this image.
Only what base.py:1-84 and logic.py:31-54 touch is modelled: HasTraits(**kwargs), Unicode / Instance traits, @default,
NodeMixin (parent / children / root), PreOrderIter / PostOrderIter."""
import ast
import collections
import collections.abc
import importlib.util
import os
import sys
import types

REF = os.environ.get('CROWD_REFERENCE', '/root/reference/crowddynamics')


class TraitType:
    def __init__(self, *args, default_value=None, help='', allow_none=False, **kw):
        self.default_value = default_value if default_value is not None else (args[1] if len(args) > 1 else None)
        self.name = None

    def __set_name__(self, owner, name):
        self.name = name

    def __get__(self, obj, cls=None):
        if obj is None:
            return self
        values = obj.__dict__.setdefault('_trait_values', {})
        if self.name not in values:
            for klass in type(obj).__mro__:
                for attr in vars(klass).values():
                    if getattr(attr, '_default_for', None) == self.name:
                        values[self.name] = attr(obj)
                        return values[self.name]
            values[self.name] = self.default_value
        return values[self.name]

    def __set__(self, obj, value):
        obj.__dict__.setdefault('_trait_values', {})[self.name] = value


class HasTraits:
    def __init__(self, *args, **kwargs):
        for key, value in kwargs.items():
            if not isinstance(getattr(type(self), key, None), TraitType):
                raise TypeError('unrecognised trait %r passed to HasTraits.__init__' % key)   # traitlets >= 5 warns / raises
            setattr(self, key, value)
        super().__init__()


def default(name):
    def deco(fn):
        fn._default_for = name
        return fn
    return deco


class NodeMixin:
    @property
    def parent(self):
        return self.__dict__.get('_parent')

    @parent.setter
    def parent(self, node):
        old = self.__dict__.get('_parent')
        if old is not None:
            old.__dict__['_children'].remove(self)
        self.__dict__['_parent'] = node
        if node is not None:
            node.__dict__.setdefault('_children', []).append(self)

    @property
    def children(self):
        return tuple(self.__dict__.get('_children', ()))

    @property
    def root(self):
        node = self
        while node.parent is not None:
            node = node.parent
        return node


def PreOrderIter(node):
    yield node
    for c in node.children:
        yield from PreOrderIter(c)


def PostOrderIter(node):
    for c in node.children:
        yield from PostOrderIter(c)
    yield node


def install(monkeypatch):
    """Registers the fakes plus the reference's real base.py / LogicNode in sys.modules (undone by monkeypatch)."""
    for name in ('Callable', 'Iterable', 'Mapping', 'MutableSequence', 'Generator', 'Collection'):
        if not hasattr(collections, name):                       # removed from `collections` in Python 3.10
            monkeypatch.setattr(collections, name, getattr(collections.abc, name), raising=False)
    tr = types.ModuleType('traitlets.traitlets')
    tr.HasTraits, tr.default = HasTraits, default
    for t in ('Unicode', 'Instance', 'Float', 'Int', 'Bool'):
        setattr(tr, t, type(t, (TraitType,), {}))
    pkg = types.ModuleType('traitlets'); pkg.traitlets = tr; pkg.__path__ = []
    at = types.ModuleType('anytree')
    at.NodeMixin, at.PreOrderIter, at.PostOrderIter = NodeMixin, PreOrderIter, PostOrderIter
    tz = types.ModuleType('dateutil.tz.tz'); tz.tzutc = lambda: None
    dz = types.ModuleType('dateutil.tz'); dz.tz = tz; dz.__path__ = []
    du = types.ModuleType('dateutil'); du.tz = dz; du.__path__ = []
    mods = {'traitlets': pkg, 'traitlets.traitlets': tr, 'anytree': at, 'dateutil': du, 'dateutil.tz': dz, 'dateutil.tz.tz': tz}
    for name in ('crowddynamics', 'crowddynamics.simulation', 'crowddynamics.core'):
        m = types.ModuleType(name); m.__path__ = []
        mods[name] = m
    for k, v in mods.items():
        monkeypatch.setitem(sys.modules, k, v)
    # the reference's real base.py, unmodified
    spec = importlib.util.spec_from_file_location('crowddynamics.simulation.base', os.path.join(REF, 'simulation', 'base.py'))
    base = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, 'crowddynamics.simulation.base', base)
    spec.loader.exec_module(base)
    # the reference's real LogicNode class (logic.py:31-54); the rest of logic.py needs numba kernels, shapely, matplotlib ...
    src = open(os.path.join(REF, 'simulation', 'logic.py')).read()
    node = [n for n in ast.parse(src).body if getattr(n, 'name', None) == 'LogicNode'][0]
    logic = types.ModuleType('crowddynamics.simulation.logic')
    logic.LogicNodeBase = base.LogicNodeBase
    exec(compile('\n'.join(src.splitlines()[node.lineno - 1:node.end_lineno]), 'reference:simulation/logic.py', 'exec'), logic.__dict__)
    monkeypatch.setitem(sys.modules, 'crowddynamics.simulation.logic', logic)
    geometry = types.ModuleType('crowddynamics.core.geometry')
    geometry.geom_to_linear_obstacles = lambda geom: geom
    monkeypatch.setitem(sys.modules, 'crowddynamics.core.geometry', geometry)
    # exceptions.py of the reference is plain Python: load it too, so that our exception classes ARE the reference's
    spec = importlib.util.spec_from_file_location('crowddynamics.exceptions', os.path.join(REF, 'exceptions.py'))
    exc = importlib.util.module_from_spec(spec)
    try:
        spec.loader.exec_module(exc)
        monkeypatch.setitem(sys.modules, 'crowddynamics.exceptions', exc)
    except Exception:
        pass
    return base, logic
