"""Sharded data parallelism on libb2d: reduce-scatter to the owner, partitioned optimizer step,
parameter all-gather.

What ``RayShardedStrategy`` (ray_lightning/ray_ddp_sharded.py:12-13) gets from FairScale —
``ShardedDataParallel`` reducing every gradient to its owner and ``OSS`` stepping the owned shard
and broadcasting it — is laid out here B200-first:

* all trainable parameters live in ONE flat fp32 buffer inside the symmetric arena, grouped by
  owner rank; gradients accumulate into a second flat buffer (``param.grad`` are views);
* ``ShardedOptimizer.step()`` is ONE kernel for Adam/AdamW (b2d_sharded_step: stage -> barrier ->
  peer-read reduce of the owned shard -> Adam in registers -> barrier -> peer-read of the other
  shards' new parameters); any other elementwise optimizer runs as reduce-scatter kernel ->
  ``base.step()`` on the owned flat shard -> all-gather kernel;
* optimizer state exists only for the owned shard (the memory saving OSS is used for) and can be
  consolidated to the stock ``torch.optim`` state-dict layout for checkpoints.
"""
from typing import List

import torch

from .partition import flat_layout, partition_parameters


class FlatShards:
    """Flat parameter / gradient buffers of one module + the owner table."""

    def __init__(self, module: torch.nn.Module, comm, rule: str = "fairscale"):
        self.comm = comm
        self.params: List[torch.nn.Parameter] = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("module has no trainable parameters")
        dev = self.params[0].device
        if dev.type != "cuda" and not getattr(comm, "is_test_double", False):
            # a real Communicator cannot exist without CUDA; only the host-logic tests' stand-in gets past here
            raise RuntimeError("FlatShards needs the module on a CUDA device (no CPU fallback)")
        if any(p.dtype != torch.float32 for p in self.params):
            raise ValueError("the sharded path keeps fp32 master parameters; got a non-fp32 parameter")
        self.numels = [p.numel() for p in self.params]
        self.owner = partition_parameters(self.numels, comm.world, rule)
        self.offsets, self.shard_off, self.total = flat_layout(self.numels, self.owner, comm.world)
        self.flat_params = comm.arena_tensor(self.total)      # symmetric: peers read it in the all-gather
        self.flat_params.zero_()
        self.flat_grads = torch.zeros(self.total, device=dev)
        for p, off, n in zip(self.params, self.offsets, self.numels):
            self.flat_params[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_params[off:off + n].view(p.shape)
            p.grad = self.flat_grads[off:off + n].view(p.shape)
        lo, hi = self.shard_off[comm.rank], self.shard_off[comm.rank + 1]
        self.own = slice(lo, hi)

    def rebind_grads(self):
        """Point every ``param.grad`` back at its slice of the flat buffer (after a set_to_none)."""
        for p, off, n in zip(self.params, self.offsets, self.numels):
            if p.grad is None or p.grad.data_ptr() != self.flat_grads[off:off + n].data_ptr():
                p.grad = self.flat_grads[off:off + n].view(p.shape)


def _fusable(opt: torch.optim.Optimizer) -> bool:
    if type(opt) not in (torch.optim.Adam, torch.optim.AdamW) or len(opt.param_groups) != 1:
        return False
    g = opt.param_groups[0]
    return not (g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable")
                or isinstance(g.get("lr"), torch.Tensor))


class ShardedOptimizer(torch.optim.Optimizer):
    """Wraps the user's optimizer the way PL wraps it in FairScale ``OSS``: same hyper-parameters,
    state only for the owned shard, parameters whole again after every ``step()``."""

    def __init__(self, base: torch.optim.Optimizer, shards: FlatShards, wire: str = "bf16", stream=None):
        self.shards, self.comm, self.wire = shards, shards.comm, wire
        self.stream = stream
        self.fused = _fusable(base)
        self._base_cls = type(base)
        group = {k: v for k, v in base.param_groups[0].items() if k != "params"}
        if len(base.param_groups) != 1:
            raise ValueError("the sharded path supports a single parameter group")
        super().__init__(shards.params, group)
        self._steps = 0
        n_own = shards.own.stop - shards.own.start
        dev = shards.flat_params.device
        if self.fused:
            self.exp_avg = torch.zeros(max(n_own, 8), device=dev)
            self.exp_avg_sq = torch.zeros(max(n_own, 8), device=dev)
            self._base = None
        else:
            # any elementwise optimizer, run on the owned shard viewed as one flat parameter
            self._own_param = torch.nn.Parameter(shards.flat_params[shards.own], requires_grad=True)
            self._own_grad = torch.zeros(max(n_own, 8), device=dev)
            self._own_param.grad = self._own_grad[:n_own]
            self._base = self._base_cls([self._own_param], **group)

    def zero_grad(self, set_to_none: bool = False):
        self.shards.flat_grads.zero_()
        self.shards.rebind_grads()

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        sh, g = self.shards, self.param_groups[0]
        self._steps += 1
        on_gpu = sh.flat_params.is_cuda
        cur = torch.cuda.current_stream(sh.flat_params.device) if on_gpu else None
        side = self.stream if self.stream is not None else cur
        if self.fused:
            self.comm.sharded_step_(sh.flat_grads, sh.flat_params, self.exp_avg, self.exp_avg_sq, sh.shard_off,
                                    step=self._steps, lr=float(g["lr"]), betas=tuple(g["betas"]), eps=float(g["eps"]),
                                    weight_decay=float(g["weight_decay"]), adamw=self._base_cls is torch.optim.AdamW,
                                    wire=self.wire, wait_stream=cur, comm_stream=side)
        else:
            for k, v in g.items():  # lr schedulers edit OUR group: mirror it
                if k != "params":
                    self._base.param_groups[0][k] = v
            self.comm.reduce_scatter(sh.flat_grads, self._own_grad, sh.shard_off, wire=self.wire,
                                     wait_stream=cur, comm_stream=side)
            if on_gpu:
                with torch.cuda.stream(side):
                    self._base.step()
            else:
                self._base.step()
            self.comm.allgather_(sh.flat_params, sh.shard_off, wait_stream=side, comm_stream=side)
        if side is not cur:
            cur.wait_stream(side)
        return loss

    # ---- checkpoints: stock torch.optim layout -------------------------------------
    def _gather_full(self, own_vec):
        sh = self.shards
        if not hasattr(self, "_gather_buf"):
            self._gather_buf = self.comm.arena_tensor(sh.total)
        buf = self._gather_buf
        buf[sh.own] = own_vec[:sh.own.stop - sh.own.start]
        if buf.is_cuda:
            torch.cuda.current_stream().synchronize()
        self.comm.allgather_(buf, sh.shard_off)
        if buf.is_cuda:
            torch.cuda.current_stream().synchronize()
        return buf.clone()

    def consolidated_state_dict(self):
        """Collective (every rank must call it): the state dict ``type(base)`` would have produced
        had it stepped all parameters on one device."""
        sh = self.shards
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(sh.params)))
        state = {}
        if self.fused:
            m, v = self._gather_full(self.exp_avg), self._gather_full(self.exp_avg_sq)
            if self._steps > 0:
                for i, (off, n, p) in enumerate(zip(sh.offsets, sh.numels, sh.params)):
                    state[i] = {"step": torch.tensor(float(self._steps)),
                                "exp_avg": m[off:off + n].view(p.shape).cpu(),
                                "exp_avg_sq": v[off:off + n].view(p.shape).cpu()}
        else:
            st = self._base.state.get(self._own_param, {})
            for key, val in st.items():
                if isinstance(val, torch.Tensor) and val.numel() == self._own_param.numel():
                    full = self._gather_full(val.reshape(-1).float())
                    for i, (off, n, p) in enumerate(zip(sh.offsets, sh.numels, sh.params)):
                        state.setdefault(i, {})[key] = full[off:off + n].view(p.shape).cpu()
                else:
                    for i in range(len(sh.params)):
                        state.setdefault(i, {})[key] = val
        return {"state": state, "param_groups": [group]}

    def state_dict(self):
        return self.consolidated_state_dict()

    def load_state_dict(self, sd):
        """Accepts the consolidated layout (written by any previous world size) and keeps the owned slice —
        the resume-with-fewer-workers contract of ray_lightning/tests/test_ddp_sharded.py:118-137."""
        sh = self.shards
        g = dict(sd["param_groups"][0])
        g.pop("params", None)
        self.param_groups[0].update(g)
        st = sd.get("state", {})
        if not st:
            return
        st = {int(k): v for k, v in st.items()}
        lo = sh.own.start
        n_own = sh.own.stop - sh.own.start
        mine = [(i, off, n) for i, (off, n) in enumerate(zip(sh.offsets, sh.numels)) if sh.owner[i] == self.comm.rank]
        if self.fused:
            for i, off, n in mine:
                if i in st:
                    self.exp_avg[off - lo:off - lo + n].copy_(st[i]["exp_avg"].reshape(-1))
                    self.exp_avg_sq[off - lo:off - lo + n].copy_(st[i]["exp_avg_sq"].reshape(-1))
                    self._steps = int(st[i]["step"])
            return
        # any other elementwise optimizer: its state belongs to ONE flat parameter (the owned shard); every
        # full-size per-parameter state tensor is cut to the owned range, scalars (step counts, ...) are taken as is
        if not mine:
            return
        keys = set()
        for i, _, _ in mine:
            if i not in st:
                raise ValueError("optimizer state for parameter %d is missing from the checkpoint" % i)
            keys |= set(st[i].keys())
        own_state = {}
        dev = sh.flat_params.device
        for key in sorted(keys):
            first = st[mine[0][0]][key]
            if isinstance(first, torch.Tensor) and first.numel() == mine[0][2] and first.dim() > 0:
                flat = torch.zeros(max(n_own, 1), device=dev, dtype=torch.float32)
                for i, off, n in mine:
                    v = st[i][key]
                    if not isinstance(v, torch.Tensor) or v.numel() != n:
                        raise ValueError("optimizer state %r of parameter %d has an unexpected shape" % (key, i))
                    flat[off - lo:off - lo + n].copy_(v.reshape(-1))
                own_state[key] = flat[:n_own].view_as(self._own_param)
            else:
                own_state[key] = first.clone() if isinstance(first, torch.Tensor) else first
        self._base.state[self._own_param] = own_state
