"""Generate tests/golden/*.npz by running the reference's real implementation (torch DDP over
gloo, oracle/reference_ddp.py) in THIS container.  Committed together with its outputs so the
fixtures can be regenerated:  python -m oracle.make_golden

Each fixture holds, for one (model, world, ddp_kwargs) case:
  world, n_buckets
  b{i}_param_ids, b{i}_lengths           integer bucket layout (bit-exact contract)
  b{i}_local_r{r}                        rank r's flat fp32 bucket BEFORE the sync
  b{i}_out_{mode}                        the bucket AFTER the sync, mode in default /
                                         allreduce_hook / bf16_compress_hook (identical on all ranks)
  param_numels
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_ddp  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    "small_mlp_w2": dict(world=2, model="small_mlp", batch=4, seed=0,
                         ddp_kwargs=dict(find_unused_parameters=True, bucket_cap_mb=0.004)),
    "small_mlp_w4": dict(world=4, model="small_mlp", batch=4, seed=1,
                         ddp_kwargs=dict(find_unused_parameters=True, bucket_cap_mb=0.004)),
    "mnist_w2": dict(world=2, model="mnist", batch=32, seed=0,
                     ddp_kwargs=dict(find_unused_parameters=True)),
}


def flatten_by_layout(grads, layout):
    out = []
    for b in layout:
        out.append(np.concatenate([grads[i].reshape(-1) for i in b["param_ids"]]))
    return out


def make_case(name, cfg):
    world = cfg["world"]
    rec = reference_ddp.run_grad_sync(world, {k: v for k, v in cfg.items() if k != "world"})
    r0 = rec[0]
    layout = sorted(r0["bf16_compress_hook"]["layout"], key=lambda b: b["index"])
    # every rank must have seen the same bucket layout, and the same reduced values
    for r in range(world):
        lr = sorted(rec[r]["bf16_compress_hook"]["layout"], key=lambda b: b["index"])
        assert [b["param_ids"] for b in lr] == [b["param_ids"] for b in layout]
        for mode in reference_ddp.HOOKS:
            for a, b in zip(rec[r][mode]["grads"], r0[mode]["grads"]):
                assert np.array_equal(a, b), (name, mode, r)
    data = {"world": np.int64(world), "n_buckets": np.int64(len(layout)),
            "param_numels": np.array([g.size for g in r0["local_grads"]], dtype=np.int64)}
    for i, b in enumerate(layout):
        assert b["index"] == i
        data["b%d_param_ids" % i] = np.array(b["param_ids"], dtype=np.int64)
        data["b%d_lengths" % i] = np.array(b["lengths"], dtype=np.int64)
        for r in range(world):
            lr = sorted(rec[r]["bf16_compress_hook"]["layout"], key=lambda x: x["index"])
            flat = lr[i]["local_flat"].astype(np.float32)
            # the bucket the hook received == this rank's local gradients, flattened by the layout
            ref = flatten_by_layout(rec[r]["local_grads"], layout)[i]
            assert np.array_equal(flat, ref), (name, i, r)
            data["b%d_local_r%d" % (i, r)] = flat
        for mode in reference_ddp.HOOKS:
            data["b%d_out_%s" % (i, mode)] = flatten_by_layout(r0[mode]["grads"], layout)[i].astype(np.float32)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **data)
    return path, len(layout)


if __name__ == "__main__":
    for name, cfg in CASES.items():
        path, nb = make_case(name, cfg)
        print("%s: %d buckets -> %s (%d bytes)" % (name, nb, path, os.path.getsize(path)))
