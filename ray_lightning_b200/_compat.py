"""Resolve the two third-party runtimes the reference is written against.

``ray`` and ``pytorch_lightning`` (1.6) are imported when present — the strategies then subclass
the real ``DDPSpawnStrategy`` exactly like the reference (ray_lightning/ray_ddp.py:7,23) — and
replaced by the in-repo stand-ins (``_runtime.miniray`` / ``_runtime.minipl``) otherwise.  This
image has neither package and no network, so the stand-ins are what runs here and on the GPU
box; the real-package branch is kept narrow and is untested (documented in DESIGN.md §7).
"""
import os

USING_REAL_RAY = False
USING_REAL_PL = False

if os.environ.get("B2D_FORCE_MINI_RUNTIME") != "1":
    try:  # pragma: no cover - not installable in this image
        import ray  # noqa: F401
        USING_REAL_RAY = True
    except Exception:
        pass
    try:  # pragma: no cover
        import pytorch_lightning as _pl  # noqa: F401
        from pytorch_lightning.strategies import DDPSpawnStrategy as _probe  # noqa: F401
        USING_REAL_PL = True
    except Exception:
        pass

if not USING_REAL_RAY:
    from ._runtime import miniray as ray  # noqa: F401

if USING_REAL_PL:  # pragma: no cover
    import pytorch_lightning as pl
    from pytorch_lightning import Callback, LightningDataModule, LightningModule, Trainer
    from pytorch_lightning.callbacks import EarlyStopping, ModelCheckpoint
    from pytorch_lightning.strategies import (DDPSpawnShardedStrategy, DDPSpawnStrategy, HorovodStrategy,
                                              ParallelStrategy, Strategy)
    from pytorch_lightning.strategies.launchers import _Launcher
    from pytorch_lightning.utilities.apply_func import apply_to_collection, move_data_to_device
    from pytorch_lightning.utilities.rank_zero import rank_zero_debug, rank_zero_info, rank_zero_only
    from pytorch_lightning.utilities.seed import reset_seed, seed_everything
else:
    from ._runtime import minipl as pl
    from ._runtime.minipl import (Callback, DDPSpawnShardedStrategy, DDPSpawnStrategy, EarlyStopping,  # noqa: F401
                                  HorovodStrategy, LightningDataModule, LightningModule, ModelCheckpoint,
                                  ParallelStrategy, Strategy, Trainer, _Launcher, apply_to_collection,
                                  move_data_to_device, rank_zero_debug, rank_zero_info, rank_zero_only,
                                  reset_seed, seed_everything)
