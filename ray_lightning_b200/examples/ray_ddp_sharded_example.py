"""Sharded training example (role of ray_lightning/examples/ray_ddp_sharded_example.py:16-71): a
transformer LM trained with RayShardedStrategy; a callback reports epoch time and peak CUDA
memory averaged over the workers (the reference's CUDACallback :16-45, including its two scalar
allreduces :33-36 — they run on the control-plane process group, the only collectives this example issues itself).

    python -m ray_lightning_b200.examples.ray_ddp_sharded_example --num-workers 2 --use-gpu
"""
import argparse
import tempfile
import time

import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from ray_lightning_b200 import RayShardedStrategy
from ray_lightning_b200._compat import Callback, LightningModule, Trainer, ray


class TokenData(Dataset):
    def __init__(self, n=256, seq=128, vocab=512, seed=0):
        self.x = torch.randint(0, vocab, (n, seq + 1), generator=torch.Generator().manual_seed(seed))

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i, :-1], self.x[i, 1:]


class TinyGPT(LightningModule):
    def __init__(self, vocab=512, d=256, layers=4, heads=4, seq=128):
        super().__init__()
        self.emb = torch.nn.Embedding(vocab, d)
        self.pos = torch.nn.Parameter(torch.zeros(seq, d))
        layer = torch.nn.TransformerEncoderLayer(d, heads, 4 * d, dropout=0.0, batch_first=True)
        self.blocks = torch.nn.TransformerEncoder(layer, layers)
        self.head = torch.nn.Linear(d, vocab)

    def forward(self, x):
        mask = torch.nn.Transformer.generate_square_subsequent_mask(x.size(1), device=x.device)
        return self.head(self.blocks(self.emb(x) + self.pos[:x.size(1)], mask=mask))

    def training_step(self, batch, batch_idx):
        x, y = batch
        loss = F.cross_entropy(self(x).flatten(0, 1), y.flatten())
        self.log("train_loss", loss)
        return loss

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=1e-3)

    def train_dataloader(self):
        return DataLoader(TokenData(), batch_size=8)


class CUDACallback(Callback):
    def on_train_epoch_start(self, trainer, pl_module):
        if torch.cuda.is_available():
            torch.cuda.reset_peak_memory_stats()
            torch.cuda.synchronize()
        self.t0 = time.time()

    def on_train_epoch_end(self, trainer, pl_module):
        dev = trainer.strategy.root_device
        peak = 0.0
        if dev.type == "cuda":
            torch.cuda.synchronize()
            peak = torch.cuda.max_memory_allocated() / 2 ** 20
        stats = torch.tensor([peak, time.time() - self.t0], dtype=torch.float64, device=dev)
        world = 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.all_reduce(stats[0:1], op=torch.distributed.ReduceOp.SUM)   # peak memory  (reference :33-34)
            torch.distributed.all_reduce(stats[1:2], op=torch.distributed.ReduceOp.SUM)   # epoch time   (reference :35-36)
            world = torch.distributed.get_world_size()
        if trainer.global_rank == 0:
            print("Average Epoch time: %.2f seconds" % (float(stats[1]) / world), flush=True)
            print("Average Peak memory %.2f MiB" % (float(stats[0]) / world), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-workers", type=int, default=2)
    ap.add_argument("--use-gpu", action="store_true")
    ap.add_argument("--num-epochs", type=int, default=1)
    a = ap.parse_args()
    ray.init(num_cpus=max(2, a.num_workers))
    trainer = Trainer(default_root_dir=tempfile.mkdtemp(), max_epochs=a.num_epochs, callbacks=[CUDACallback()],
                      strategy=RayShardedStrategy(num_workers=a.num_workers, use_gpu=a.use_gpu))
    trainer.fit(TinyGPT())
    print({k: float(v) for k, v in trainer.callback_metrics.items()})
    ray.shutdown()
