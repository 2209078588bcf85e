"""The plugin surface is the reference's, checked against the reference SOURCE by AST (it cannot be
imported: ray / pytorch_lightning are absent).  Runs where /root/reference exists (the build
container); skipped on the GPU box."""
import ast
import inspect
import os

import pytest

REF = "/root/reference/ray_lightning"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference sources not present on this box")


def _cls(path, name):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef) and node.name == name:
            return node
    raise KeyError(name)


def _methods(node):
    return {n.name: n for n in node.body if isinstance(n, ast.FunctionDef)}


def _sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.literal_eval(d) if isinstance(d, ast.Constant) else "<expr>" for d in a.defaults]
    return list(zip(names, defaults)), (a.vararg.arg if a.vararg else None), (a.kwarg.arg if a.kwarg else None), \
        [k.arg for k in a.kwonlyargs]


def _mine(fn):
    sig = inspect.signature(fn)
    pos, var, kw, kwonly = [], None, None, []
    for p in sig.parameters.values():
        if p.kind == p.VAR_POSITIONAL:
            var = p.name
        elif p.kind == p.VAR_KEYWORD:
            kw = p.name
        elif p.kind == p.KEYWORD_ONLY:
            kwonly.append(p.name)
        else:
            pos.append((p.name, None if p.default is p.empty else p.default))
    return pos, var, kw, kwonly


def test_ray_strategy_surface():
    from ray_lightning_b200 import RayStrategy
    ref = _cls("ray_ddp.py", "RayStrategy")
    rm = _methods(ref)
    assert _sig(rm["__init__"]) == _mine(RayStrategy.__init__)
    for name, fn in rm.items():
        assert hasattr(RayStrategy, name), name
        if name != "__init__" and not isinstance(inspect.getattr_static(RayStrategy, name), property):
            assert _sig(fn)[0] == _mine(getattr(RayStrategy, name))[0], name
    assert RayStrategy.strategy_name == "ddp_ray"


def test_sharded_and_horovod_surface():
    from ray_lightning_b200 import HorovodRayStrategy, RayShardedStrategy
    assert RayShardedStrategy.strategy_name == "ddp_sharded_ray"
    ref = _cls("ray_horovod.py", "HorovodRayStrategy")
    rm = _methods(ref)
    assert _sig(rm["__init__"]) == _mine(HorovodRayStrategy.__init__)
    for name in rm:
        assert hasattr(HorovodRayStrategy, name), name
    assert HorovodRayStrategy.strategy_name == "horovod_ray"


def test_launcher_and_executor_surface():
    from ray_lightning_b200.launchers import RayHorovodLauncher, RayLauncher
    from ray_lightning_b200.launchers.utils import _RayExecutorImpl, _RayOutput
    rm = _methods(_cls("launchers/ray_launcher.py", "RayLauncher"))
    for name, fn in rm.items():
        assert hasattr(RayLauncher, name), name
        assert _sig(fn) == _mine(getattr(RayLauncher, name)), name
    ex = _methods(_cls("launchers/utils.py", "RayExecutor"))
    for name, fn in ex.items():
        assert _sig(fn) == _mine(getattr(_RayExecutorImpl, name)), name
    out = _cls("launchers/utils.py", "_RayOutput")
    fields = [n.target.id for n in out.body if isinstance(n, ast.AnnAssign)]
    assert list(_RayOutput._fields) == fields
    hv = _methods(_cls("launchers/ray_horovod_launcher.py", "RayHorovodLauncher"))
    assert _sig(hv["launch"]) == _mine(RayHorovodLauncher.launch)


def test_module_level_names():
    import ray_lightning_b200 as pkg
    from ray_lightning_b200 import session, tune, util
    tree = ast.parse(open(os.path.join(REF, "__init__.py")).read())
    ref_all = next(ast.literal_eval(n.value) for n in tree.body if isinstance(n, ast.Assign) and n.targets[0].id == "__all__")
    assert sorted(pkg.__all__) == sorted(ref_all)
    for mod, path in ((session, "session.py"), (util, "util.py")):
        t = ast.parse(open(os.path.join(REF, path)).read())
        for n in t.body:
            if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name != "DelayedGPUAccelerator":
                assert hasattr(mod, n.name), (path, n.name)
    for name in ("TuneReportCallback", "TuneReportCheckpointCallback", "get_tune_resources", "is_session_enabled", "TUNE_INSTALLED"):
        assert hasattr(tune, name)
