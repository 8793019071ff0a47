"""Test infrastructure ONLY: makes the read-only reference at /root/reference importable
in a container that lacks matplotlib / ruamel.yaml / pydot, so the REAL reference code
(Cells.make_world, Simulator.init_core/init_dynamics/_run_sim_core_loop, sim_toolbox,
finitediff, channels, networks) can be executed to (1) generate the golden fixtures in
tests/golden/ and (2) validate oracle/betse_oracle.py.

Nothing here is on the product path; it is used by tests/golden/make_golden.py and by the
``not gpu`` oracle-validation tests, and only when /root/reference exists (it does not on the
GPU box).  No reference source is copied: the stubs below stand in for *third-party*
plotting / YAML libraries that the numeric code never calls.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

def _find_reference():
    """$BETSE_REFERENCE, else the read-only checkout of the build container, else the copy that
    tools/install_reference.py leaves under baseline/_ref (git-ignored; it travels to the GPU box)."""
    env = os.environ.get("BETSE_REFERENCE")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in ("/root/reference", os.path.join(here, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "betse", "science")):
            return cand
    return "/root/reference"


REF_ROOT = _find_reference()

_STUB_ROOTS = ("matplotlib", "ruamel", "pydot", "mpl_toolkits")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "betse", "science"))


class _StubMeta(type):
    """Metaclass for dummy classes: attribute access on the class yields further dummies."""

    def __getattr__(cls, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _make_dummy(name)

    def __getitem__(cls, key):
        return _make_dummy("item")()


def _make_dummy(name):
    return _StubMeta(name, (object,), {
        "__init__": lambda self, *a, **k: None,
        # decorator semantics: ``@stub.register('x')`` must hand the class back unchanged
        "__call__": lambda self, *a, **k: a[0]
        if (len(a) == 1 and not k and (isinstance(a[0], type) or callable(a[0]))) else self,
        "__getattr__": lambda self, n: (_ for _ in ()).throw(AttributeError(n))
        if n.startswith("__") else _make_dummy(n)(),
        "__iter__": lambda self: iter(()),
        "__len__": lambda self: 0,
        "__enter__": lambda self: self,
        "__exit__": lambda self, *a: False,
        "__getitem__": lambda self, k: _make_dummy("item")(),
        "__setitem__": lambda self, k, v: None,
    })


class _StubModule(types.ModuleType):
    __path__ = []
    __version__ = "3.9.0"

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _make_dummy(name)
        obj.__module__ = self.__name__      # picklable by reference through the stub module
        setattr(self, name, obj)
        return obj


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        return _StubModule(spec.name)

    def exec_module(self, module):
        if module.__name__ == "ruamel.yaml":
            _fill_ruamel_yaml(module)
        if module.__name__ == "ruamel":
            import importlib
            module.yaml = importlib.import_module("ruamel.yaml")
        if module.__name__ == "matplotlib":
            module.rcParams = {}
            module.get_backend = lambda: "agg"
            module.use = lambda *a, **k: None

            class _Colormaps:
                def __getitem__(self, name):
                    import importlib
                    return importlib.import_module("matplotlib.colors").Colormap()

                def __iter__(self):
                    return iter(())

                def register(self, *a, **k):
                    return None
            module.colormaps = _Colormaps()


def _fill_ruamel_yaml(mod):
    """ruamel.yaml stand-in backed by PyYAML (round-trip fidelity is not needed)."""
    import yaml as _pyyaml

    import re

    class _Loader12(_pyyaml.SafeLoader):
        """PyYAML is YAML 1.1; the reference's files are YAML 1.2 (``1e-2`` is a float,
        ``yes/no/on/off`` are strings).  Re-register the core-schema resolvers."""

    _Loader12.yaml_implicit_resolvers = {
        k: [(t, r) for (t, r) in v if t not in ("tag:yaml.org,2002:float", "tag:yaml.org,2002:bool",
                                                 "tag:yaml.org,2002:int")]
        for k, v in _pyyaml.SafeLoader.yaml_implicit_resolvers.items()}
    _Loader12.add_implicit_resolver(
        "tag:yaml.org,2002:bool", re.compile(r"^(?:true|True|TRUE|false|False|FALSE)$"), list("tTfF"))
    _Loader12.add_implicit_resolver(
        "tag:yaml.org,2002:int", re.compile(r"^[-+]?[0-9]+$"), list("-+0123456789"))
    _Loader12.add_implicit_resolver(
        "tag:yaml.org,2002:float",
        re.compile(r"^(?:[-+]?(?:\.[0-9]+|[0-9]+(?:\.[0-9]*)?)(?:[eE][-+]?[0-9]+)?"
                   r"|[-+]?\.(?:inf|Inf|INF)|\.(?:nan|NaN|NAN))$"), list("-+0123456789."))

    class YAML:
        def __init__(self, *a, **k):
            self.default_flow_style = None
            self.representer = _make_dummy("representer")()

        def load(self, stream):
            return _pyyaml.load(stream, Loader=_Loader12)

        def dump(self, data, stream=None):
            return _pyyaml.safe_dump(data, stream)

        def register_class(self, cls):
            return cls

    mod.YAML = YAML
    mod.__version__ = "0.17.0"


_installed = False


def install():
    """Idempotently put the reference on sys.path behind the third-party stubs."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.meta_path.insert(0, _StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


def bypass_science_init():
    """Pre-register a bare ``betse.science`` package so that its ``__init__`` (which
    initialises matplotlib colormaps, logging and the app-metadata singleton) is skipped."""
    install()
    import betse
    if "betse.science" not in sys.modules:
        m = types.ModuleType("betse.science")
        m.__path__ = [os.path.join(REF_ROOT, "betse", "science")]
        sys.modules["betse.science"] = m
        betse.science = m
