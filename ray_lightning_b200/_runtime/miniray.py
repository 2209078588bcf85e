"""A small single-node actor runtime with the slice of Ray's API that RayLauncher uses.

``ray`` cannot be installed in this image (no network), and the reference's launcher is written
against it: ``ray.remote`` actors created with ``.options(num_cpus, num_gpus, resources)``,
``actor.method.remote()`` futures, ``ray.get / ray.wait / ray.put / ray.kill``,
``ray.get_gpu_ids()``, ``ray.util.get_node_ip_address()`` and ``ray.util.queue.Queue``
(ray_lightning/launchers/ray_launcher.py:71-128,221-250, launchers/utils.py:27-52, util.py:57-70).
This module provides exactly that over ``multiprocessing`` (spawn) + cloudpickle: one OS process
per actor, GPUs handed out by the driver and exported as CUDA_VISIBLE_DEVICES before the actor
starts — the same contract Ray gives the reference.  When the real ``ray`` is importable,
``ray_lightning_b200._compat`` uses it instead and this module is idle.

Out of scope: multi-node, object spilling, fault tolerance, scheduling policies.
"""
import itertools
import multiprocessing as mp
import os
import threading
import time
import traceback
from types import SimpleNamespace

import cloudpickle

_NODE_IP = "127.0.0.1"
_NODE_ID = "b2d0" * 14  # 56 hex chars, like ray's NodeID

_state = SimpleNamespace(initialized=False, resources={}, available={}, gpu_load={}, actors=[], in_actor=False,
                         actor_gpu_ids=[], lock=threading.RLock())


class RayError(Exception):
    pass


class RayActorError(RayError):
    pass


class RayTaskError(RayError):
    def __init__(self, message, cause_repr=None):
        super().__init__(message)
        self.cause_repr = cause_repr


class GetTimeoutError(RayError, TimeoutError):
    pass


# ---- cluster state ---------------------------------------------------------------------------
def init(address=None, num_cpus=None, num_gpus=None, resources=None, ignore_reinit_error=True, **_kw):
    with _state.lock:
        if _state.initialized:
            if ignore_reinit_error:
                return
            raise RuntimeError("miniray.init() called twice")
        if num_cpus is None:
            num_cpus = os.cpu_count() or 1
        if num_gpus is None:
            try:
                import torch
                num_gpus = torch.cuda.device_count()
            except Exception:
                num_gpus = 0
        _state.resources = {"CPU": float(num_cpus), "GPU": float(num_gpus)}
        for k, v in (resources or {}).items():
            _state.resources[k] = float(v)
        _state.available = dict(_state.resources)
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = [s for s in visible.split(",") if s] if visible and visible != "NoDevFiles" else [str(i) for i in range(int(num_gpus))]
        ids = ids[:int(num_gpus)] if len(ids) >= int(num_gpus) else [str(i) for i in range(int(num_gpus))]
        _state.gpu_load = {g: 0.0 for g in ids}
        _state.initialized = True


def is_initialized():
    return _state.initialized


def shutdown():
    with _state.lock:
        for a in list(_state.actors):
            a._terminate()
        _state.actors.clear()
        _state.initialized = False
        _state.resources, _state.available, _state.gpu_load = {}, {}, {}


def cluster_resources():
    return dict(_state.resources)


def available_resources():
    return {k: v for k, v in _state.available.items() if v > 0}


def _reserve(req):
    """Reserve resources for one actor; returns the GPU ids it may use (Ray's bin-packing of
    whole and fractional GPUs: lowest-index GPU with room)."""
    with _state.lock:
        for k, v in req.items():
            if v and _state.available.get(k, 0.0) + 1e-9 < v:
                raise RayError("insufficient resource %s: requested %s, available %s"
                               % (k, v, _state.available.get(k, 0.0)))
        gpu_ids = []
        need = float(req.get("GPU", 0) or 0)
        if need > 0:
            if need >= 1:
                whole = [g for g, load in _state.gpu_load.items() if load == 0.0][:int(need)]
                if len(whole) < int(need):
                    raise RayError("not enough free GPUs for num_gpus=%s" % need)
                for g in whole:
                    _state.gpu_load[g] = 1.0
                gpu_ids = whole
            else:
                for g, load in _state.gpu_load.items():
                    if load + need <= 1.0 + 1e-9:
                        _state.gpu_load[g] = load + need
                        gpu_ids = [g]
                        break
                else:
                    raise RayError("no GPU has %s capacity left" % need)
        for k, v in req.items():
            if v:
                _state.available[k] = _state.available.get(k, 0.0) - v
        return gpu_ids


def _release(req, gpu_ids):
    with _state.lock:
        for k, v in req.items():
            if v:
                _state.available[k] = _state.available.get(k, 0.0) + v
        need = float(req.get("GPU", 0) or 0)
        for g in gpu_ids:
            if g in _state.gpu_load:
                _state.gpu_load[g] = max(0.0, _state.gpu_load[g] - (1.0 if need >= 1 else need))


# ---- object refs -----------------------------------------------------------------------------
class ObjectRef:
    __slots__ = ("_actor", "_call_id", "_value", "_ready", "_error")

    def __init__(self, actor=None, call_id=None, value=None, ready=False):
        self._actor, self._call_id, self._value, self._ready, self._error = actor, call_id, value, ready, None

    def __reduce__(self):
        if self._ready and self._error is None:
            return (put, (self._value,))
        return (_stub, ())

    def _resolve(self, ok, payload):
        if ok:
            self._value = cloudpickle.loads(payload)
        else:
            self._error = payload
        self._ready = True


def put(value):
    return ObjectRef(value=value, ready=True)


def _deref(x):
    return get(x) if isinstance(x, ObjectRef) else x


# ---- actors ----------------------------------------------------------------------------------
def _actor_main(conn, cls_blob, init_blob, env, gpu_ids):
    os.environ.update(env)
    _state.in_actor = True
    _state.actor_gpu_ids = list(gpu_ids)
    _state.initialized = True
    try:
        cls = cloudpickle.loads(cls_blob)
        args, kwargs = cloudpickle.loads(init_blob)
        inst = cls(*args, **kwargs)
        conn.send((0, True, cloudpickle.dumps(None)))
    except BaseException:
        conn.send((0, False, traceback.format_exc()))
        return
    while True:
        try:
            msg = conn.recv()
        except (EOFError, OSError):
            break
        if msg is None:
            break
        call_id, name, blob = msg
        try:
            a, kw = cloudpickle.loads(blob)
            res = getattr(inst, name)(*a, **kw)
            conn.send((call_id, True, cloudpickle.dumps(res)))
        except BaseException:
            try:
                conn.send((call_id, False, traceback.format_exc()))
            except Exception:
                break


class _Method:
    def __init__(self, actor, name):
        self._actor, self._name = actor, name

    def remote(self, *args, **kwargs):
        return self._actor._submit(self._name, args, kwargs)


class _RemoteHandleStub:
    """What an ActorHandle / pending ObjectRef turns into when pickled into another process: actors
    are driven from the driver only (the launcher pickles itself, handles included, to its workers)."""

    def __getattr__(self, name):
        raise RayActorError("actor handles are usable in the driver process only")


def _stub():
    return _RemoteHandleStub()


class ActorHandle:
    def __reduce__(self):
        return (_stub, ())

    def __init__(self, cls, args, kwargs, req, gpu_ids):
        self._req, self._gpu_ids = req, gpu_ids
        self._ids = itertools.count(1)
        self._pending = {}
        self._dead = False
        self._lock = threading.RLock()
        wants_gpu = float(req.get("GPU", 0) or 0) > 0
        # GPU workers get a pristine interpreter (spawn).  CPU-only workers come from a fork server that has
        # torch imported already: same isolation, but the multi-second import is paid once per driver.
        if wants_gpu or os.environ.get("B2D_ACTOR_START", "") == "spawn":
            ctx = mp.get_context("spawn")
        else:
            ctx = mp.get_context("forkserver")
            ctx.set_forkserver_preload(["torch", "cloudpickle", "numpy"])
        self._conn, child = ctx.Pipe(duplex=True)
        env = {"CUDA_VISIBLE_DEVICES": ",".join(gpu_ids)} if wants_gpu else {}
        self._proc = ctx.Process(target=_actor_main, args=(child, cloudpickle.dumps(cls), cloudpickle.dumps((args, kwargs)),
                                                           env, gpu_ids), daemon=True)
        self._proc.start()
        child.close()
        boot = ObjectRef(self, 0)
        self._pending[0] = boot
        get(boot)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return _Method(self, name)

    def _submit(self, name, args, kwargs):
        if self._dead:
            raise RayActorError("actor is dead")
        args = tuple(_deref(a) for a in args)  # top-level ObjectRefs are resolved, like Ray
        kwargs = {k: _deref(v) for k, v in kwargs.items()}
        with self._lock:
            cid = next(self._ids)
            ref = ObjectRef(self, cid)
            self._pending[cid] = ref
            self._conn.send((cid, name, cloudpickle.dumps((args, kwargs))))
        return ref

    def _pump(self, timeout):
        """Move finished calls from the pipe into their refs. Returns False when the actor died."""
        with self._lock:
            try:
                while self._conn.poll(timeout):
                    cid, ok, payload = self._conn.recv()
                    ref = self._pending.pop(cid, None)
                    if ref is not None:
                        ref._resolve(ok, payload)
                    timeout = 0
            except (EOFError, OSError):
                self._fail_all("actor process exited (pipe closed)")
                return False
            if not self._proc.is_alive() and not self._conn.poll(0):
                if self._pending:
                    self._fail_all("actor process died with exit code %s" % self._proc.exitcode)
                return False
        return True

    def _fail_all(self, why):
        self._dead = True
        for ref in self._pending.values():
            ref._error, ref._ready = "RayActorError: " + why, True
        self._pending.clear()

    def _terminate(self):
        if self._dead and not self._proc.is_alive():
            return
        self._dead = True
        try:
            self._conn.send(None)
        except Exception:
            pass
        self._proc.join(timeout=2)
        if self._proc.is_alive():
            self._proc.terminate()
            self._proc.join(timeout=5)
        try:
            self._conn.close()
        except Exception:
            pass
        _release(self._req, self._gpu_ids)


class ActorClass:
    def __init__(self, cls, default_opts=None):
        self._cls = cls
        self._opts = dict(default_opts or {})
        self.__name__ = getattr(cls, "__name__", "Actor")

    def options(self, num_cpus=None, num_gpus=None, resources=None, **_kw):
        o = dict(self._opts)
        if num_cpus is not None:
            o["num_cpus"] = num_cpus
        if num_gpus is not None:
            o["num_gpus"] = num_gpus
        if resources is not None:
            o["resources"] = resources
        return ActorClass(self._cls, o)

    def remote(self, *args, **kwargs):
        if not _state.initialized:
            init()
        req = {"CPU": float(self._opts.get("num_cpus", 1) or 0), "GPU": float(self._opts.get("num_gpus", 0) or 0)}
        for k, v in (self._opts.get("resources") or {}).items():
            req[k] = float(v)
        gpu_ids = _reserve(req)
        try:
            handle = ActorHandle(self._cls, args, kwargs, req, gpu_ids)
        except BaseException:
            _release(req, gpu_ids)
            raise
        _state.actors.append(handle)
        return handle


def remote(*args, **opts):
    if len(args) == 1 and not opts and isinstance(args[0], type):
        return ActorClass(args[0])

    def deco(cls):
        return ActorClass(cls, opts)
    return deco


def kill(actor, no_restart=True):
    actor._terminate()
    if actor in _state.actors:
        _state.actors.remove(actor)


# ---- futures ----------------------------------------------------------------------------------
def _raise(ref):
    msg = ref._error
    if isinstance(msg, str) and msg.startswith("RayActorError"):
        raise RayActorError(msg)
    raise RayTaskError("remote call failed:\n%s" % msg, msg)


def get(refs, timeout=None):
    if isinstance(refs, (list, tuple)):
        return [get(r, timeout=timeout) for r in refs]
    ref = refs
    deadline = None if timeout is None else time.time() + timeout
    while not ref._ready:
        alive = ref._actor._pump(0.05)
        if ref._ready:
            break
        if not alive:
            ref._error, ref._ready = "RayActorError: actor died before returning", True
            break
        if deadline is not None and time.time() > deadline:
            raise GetTimeoutError("get timed out")
    if ref._error is not None:
        _raise(ref)
    return ref._value


def wait(refs, num_returns=1, timeout=None):
    refs = list(refs)
    deadline = None if timeout is None else time.time() + timeout
    while True:
        for r in refs:
            if not r._ready and r._actor is not None:
                r._actor._pump(0)
        ready = [r for r in refs if r._ready]
        if len(ready) >= num_returns or (deadline is not None and time.time() >= deadline):
            ready = ready[:max(num_returns, 0)] if len(ready) > num_returns else ready
            rest = [r for r in refs if r not in ready]
            return ready, rest
        time.sleep(0.002)


# ---- in-actor context -----------------------------------------------------------------------------
def get_gpu_ids():
    return list(_state.actor_gpu_ids)


class _NodeID:
    def hex(self):
        return _NODE_ID


class _RuntimeContext:
    node_id = _NodeID()


def get_runtime_context():
    return _RuntimeContext()


class _Queue:
    """ray.util.queue.Queue stand-in: a manager queue whose items are cloudpickled (closures)."""

    def __init__(self, maxsize=0, actor_options=None):
        self._mgr = mp.get_context("spawn").Manager()
        self._q = self._mgr.Queue(maxsize)

    def __getstate__(self):
        return {"_q": self._q, "_mgr": None}

    def __setstate__(self, st):
        self.__dict__.update(st)

    def put(self, item, block=True, timeout=None):
        self._q.put(cloudpickle.dumps(item), block, timeout)

    def get(self, block=True, timeout=None):
        return cloudpickle.loads(self._q.get(block, timeout))

    def empty(self):
        return self._q.empty()

    def qsize(self):
        return self._q.qsize()

    def shutdown(self):
        if self._mgr is not None:
            self._mgr.shutdown()
            self._mgr = None


util = SimpleNamespace(get_node_ip_address=lambda: _NODE_IP, queue=SimpleNamespace(Queue=_Queue),
                       PublicAPI=lambda *a, **k: (lambda f: f))
exceptions = SimpleNamespace(RayActorError=RayActorError, RayTaskError=RayTaskError, GetTimeoutError=GetTimeoutError)
actor = SimpleNamespace(ActorHandle=ActorHandle)
