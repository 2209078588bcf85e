"""Shared models/helpers for the plumbing tests — same shapes and hyper-parameters as the
reference's (ray_lightning/tests/utils.py:16-273), on synthetic data (no dataset downloads)."""
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from ray_lightning_b200._compat import LightningDataModule, LightningModule, Trainer


class RandomDataset(Dataset):
    def __init__(self, size, length, seed=0):
        self.data = torch.randn(length, size, generator=torch.Generator().manual_seed(seed))

    def __getitem__(self, i):
        return self.data[i]

    def __len__(self):
        return len(self.data)


class BoringModel(LightningModule):
    """Linear(32, 2), SGD lr 0.1 + StepLR, 64x32 random data (reference tests/utils.py:28-96)."""

    def __init__(self):
        super().__init__()
        self.layer = torch.nn.Linear(32, 2)
        self.val_epoch = 0

    def forward(self, x):
        return self.layer(x)

    def _loss(self, pred):
        return F.mse_loss(pred, torch.ones_like(pred))

    def training_step(self, batch, batch_idx):
        return {"loss": self._loss(self.layer(batch))}

    def training_epoch_end(self, outputs):
        torch.stack([o["loss"] for o in outputs]).mean()

    def validation_step(self, batch, batch_idx):
        self.layer(batch)
        loss = torch.tensor(1.0)
        self.log("val_loss", loss)
        return {"x": loss}

    def validation_epoch_end(self, outputs):
        torch.stack([o["x"] for o in outputs]).mean()
        self.val_epoch += 1

    def test_step(self, batch, batch_idx):
        return {"y": self._loss(self.layer(batch))}

    def test_epoch_end(self, outputs):
        torch.stack([o["y"] for o in outputs]).mean()

    def configure_optimizers(self):
        opt = torch.optim.SGD(self.layer.parameters(), lr=0.1)
        return [opt], [torch.optim.lr_scheduler.StepLR(opt, step_size=1)]

    def train_dataloader(self):
        return DataLoader(RandomDataset(32, 64, 0))

    def val_dataloader(self):
        return DataLoader(RandomDataset(32, 64, 1))

    def test_dataloader(self):
        return DataLoader(RandomDataset(32, 64, 2))

    def on_save_checkpoint(self, checkpoint):
        checkpoint["val_epoch"] = self.val_epoch

    def on_load_checkpoint(self, checkpoint):
        self.val_epoch = checkpoint["val_epoch"]


class AdamBoringModel(BoringModel):
    def configure_optimizers(self):
        return torch.optim.Adam(self.layer.parameters(), lr=0.05)


class SyntheticMNIST(Dataset):
    """MNIST-shaped, learnable without a download: class k lights up a k-dependent block."""

    def __init__(self, n=2048, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.y = torch.randint(0, 10, (n,), generator=g)
        self.x = torch.rand(n, 1, 28, 28, generator=g) * 0.3
        for i, k in enumerate(self.y.tolist()):
            self.x[i, 0, 2 * k:2 * k + 6, 2 * k:2 * k + 6] += 0.7

    def __len__(self):
        return len(self.y)

    def __getitem__(self, i):
        return self.x[i], self.y[i]


class LightningMNISTClassifier(LightningModule):
    """784 -> l1 -> l2 -> 10, Adam (reference tests/utils.py:99-148)."""

    def __init__(self, config, data_dir=None):
        super().__init__()
        self.lr = config["lr"]
        self.batch_size = config["batch_size"]
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
        loss = F.nll_loss(logits, y.long())
        self.log("ptl/train_loss", loss)
        self.log("ptl/train_accuracy", (logits.argmax(1) == y).float().mean())
        return loss

    def validation_step(self, batch, batch_idx):
        x, y = batch
        logits = self(x)
        return {"val_loss": F.nll_loss(logits, y.long()), "val_accuracy": (logits.argmax(1) == y).float().mean()}

    def validation_epoch_end(self, outputs):
        self.log("ptl/val_loss", torch.stack([o["val_loss"] for o in outputs]).mean())
        self.log("ptl/val_accuracy", torch.stack([o["val_accuracy"] for o in outputs]).mean())


class MNISTDataModule(LightningDataModule):
    def __init__(self, batch_size=32):
        self.batch_size = batch_size

    def setup(self, stage=None):
        self.train, self.val, self.test = SyntheticMNIST(2048, 0), SyntheticMNIST(256, 1), SyntheticMNIST(512, 2)

    def train_dataloader(self):
        return DataLoader(self.train, batch_size=self.batch_size)

    def val_dataloader(self):
        return DataLoader(self.val, batch_size=self.batch_size)

    def test_dataloader(self):
        return DataLoader(self.test, batch_size=self.batch_size)


class XORModel(LightningModule):
    def __init__(self, input_dim=2, output_dim=1):
        super().__init__()
        self.save_hyperparameters()
        self.lin1 = torch.nn.Linear(input_dim, 8)
        self.lin2 = torch.nn.Linear(8, output_dim)

    def forward(self, features):
        return torch.sigmoid(self.lin2(torch.tanh(self.lin1(features.float()))))

    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=0.02)

    def training_step(self, batch, batch_nb):
        return F.binary_cross_entropy(self(batch["x"]), batch["y"].unsqueeze(1).float())

    def validation_step(self, batch, batch_nb):
        loss = F.binary_cross_entropy(self(batch["x"]), batch["y"].unsqueeze(1).float())
        self.log("val_loss", loss, on_step=True)
        self.log("val_bar", torch.tensor(5.678), on_step=True)  # constants, to check the round trip
        return loss

    def validation_epoch_end(self, outputs):
        self.log("avg_val_loss", torch.stack(outputs).mean())
        self.log("val_foo", torch.tensor(1.234))


class XORDataModule(LightningDataModule):
    def train_dataloader(self):
        return iter([{"x": torch.tensor([[0.0, 0.0]]), "y": torch.tensor([0])},
                     {"x": torch.tensor([[1.0, 1.0]]), "y": torch.tensor([0])}])

    def val_dataloader(self):
        return iter([{"x": torch.tensor([[0.0, 1.0]]), "y": torch.tensor([1])},
                     {"x": torch.tensor([[1.0, 0.0]]), "y": torch.tensor([1])}])


def get_trainer(dir, strategy, max_epochs=1, limit_train_batches=10, limit_val_batches=10, callbacks=None,
                checkpoint_callback=True, **trainer_kwargs):
    return Trainer(default_root_dir=dir, callbacks=callbacks or [], strategy=strategy, max_epochs=max_epochs,
                   limit_train_batches=limit_train_batches, limit_val_batches=limit_val_batches,
                   enable_progress_bar=False, checkpoint_callback=checkpoint_callback, **trainer_kwargs)


def train_test(trainer, model):
    """Training must move the weights: ||delta of per-tensor L1 norms|| > 0.1 (reference :236-245)."""
    before = torch.tensor([p.abs().sum() for p in model.parameters()])
    trainer.fit(model)
    after = torch.tensor([p.abs().sum() for p in model.parameters()])
    assert trainer.state.finished, "Trainer failed with %s" % trainer.state
    assert torch.norm(before - after) > 0.1


def load_test(trainer, model):
    trainer.fit(model)
    assert type(model).load_from_checkpoint(trainer.checkpoint_callback.best_model_path) is not None


def predict_test(trainer, model, dm):
    """Accuracy >= 0.5 on the held-out split (reference :256-272)."""
    trainer.fit(model, datamodule=dm)
    model = trainer.lightning_module
    dm.setup(stage="test")
    correct = total = 0
    for x, y in dm.test_dataloader():
        with torch.no_grad():
            correct += int((model(x).cpu().argmax(1) == y).sum())
            total += len(y)
    assert correct / total >= 0.5, "accuracy %.3f < 0.5" % (correct / total)
