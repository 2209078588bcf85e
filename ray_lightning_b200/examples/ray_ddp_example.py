"""MNIST-shaped classifier trained with RayStrategy — BASELINE.json config #1.

Same model and hyper-parameters as ray_lightning/examples/ray_ddp_example.py:18-58,167
(784 -> 32 -> 64 -> 10, Adam lr 1e-1 in the example config, batch 32); the dataset is synthetic
(MNIST-shaped, class dependent) because this environment has no network for the download at
ref :24-28.

    python -m ray_lightning_b200.examples.ray_ddp_example --num-workers 2            # CPU / gloo
    python -m ray_lightning_b200.examples.ray_ddp_example --num-workers 2 --use-gpu  # libb2d hook
"""
import argparse
import tempfile

import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from ray_lightning_b200 import RayStrategy
from ray_lightning_b200._compat import LightningModule, Trainer, ray


class SyntheticMNIST(Dataset):
    def __init__(self, n=4096, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.y = torch.randint(0, 10, (n,), generator=g)
        self.x = torch.rand(n, 1, 28, 28, generator=g) * 0.3
        for i, k in enumerate(self.y.tolist()):
            self.x[i, 0, 2 * k:2 * k + 6, 2 * k:2 * k + 6] += 0.7

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class MNISTClassifier(LightningModule):
    def __init__(self, config):
        super().__init__()
        self.lr, self.batch_size = config["lr"], config["batch_size"]
        self.layer_1 = torch.nn.Linear(28 * 28, config["layer_1"])
        self.layer_2 = torch.nn.Linear(config["layer_1"], config["layer_2"])
        self.layer_3 = torch.nn.Linear(config["layer_2"], 10)

    def forward(self, x):
        x = x.view(x.size(0), -1)
        x = torch.relu(self.layer_1(x))
        x = torch.relu(self.layer_2(x))
        return F.log_softmax(self.layer_3(x), dim=1)

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.lr)

    def training_step(self, batch, batch_idx):
        x, y = batch
        logits = self(x)
        loss = F.nll_loss(logits, y)
        self.log("ptl/train_loss", loss)
        self.log("ptl/train_accuracy", (logits.argmax(1) == y).float().mean())
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        logits = self(x)
        return {"val_loss": F.nll_loss(logits, y), "val_accuracy": (logits.argmax(1) == y).float().mean()}

    def validation_epoch_end(self, outputs):
        self.log("ptl/val_loss", torch.stack([o["val_loss"] for o in outputs]).mean())
        self.log("ptl/val_accuracy", torch.stack([o["val_accuracy"] for o in outputs]).mean())

    def train_dataloader(self):
        return DataLoader(SyntheticMNIST(4096, 0), batch_size=self.batch_size)

    def val_dataloader(self):
        return DataLoader(SyntheticMNIST(512, 1), batch_size=self.batch_size)


def train_mnist(config, num_epochs=2, num_workers=2, use_gpu=False, callbacks=None, root=None, **strategy_kwargs):
    model = MNISTClassifier(config)
    trainer = Trainer(default_root_dir=root or tempfile.mkdtemp(), max_epochs=num_epochs, callbacks=callbacks or [],
                      strategy=RayStrategy(num_workers=num_workers, use_gpu=use_gpu, **strategy_kwargs))
    trainer.fit(model)
    return trainer, model


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-workers", type=int, default=2)
    ap.add_argument("--use-gpu", action="store_true")
    ap.add_argument("--smoke-test", action="store_true")
    ap.add_argument("--num-epochs", type=int, default=2)
    a = ap.parse_args()
    ray.init(num_cpus=max(2, a.num_workers))
    config = {"layer_1": 32, "layer_2": 64, "lr": 1e-2, "batch_size": 32}
    trainer, _ = train_mnist(config, num_epochs=1 if a.smoke_test else a.num_epochs, num_workers=a.num_workers,
                             use_gpu=a.use_gpu)
    print({k: float(v) for k, v in trainer.callback_metrics.items()})
    ray.shutdown()
