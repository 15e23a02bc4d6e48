"""``Predictor``: the downstream consumer of the compressed representations (reference lossyless/predictors.py:38-232),
without PyTorch-Lightning: a small network trained on (decompressed) ``Z`` to predict the task labels -- the "lossless
prediction" half of the paper, and the check that the 16-bit encoder's symbol flips do not move downstream accuracy.

Same constructor (``hparams`` + optional frozen ``featurizer``), ``forward / step / loss / training_step /
validation_step / test_step / configure_optimizers`` and log keys (``loss, acc, err, tasks_{max,std,min,mean,median},
inference_time``).  Plain torch by design: a 512 -> n_classes probe is not on the hot path; the
featurizer in front of it is the B200 path.  ``fit`` is the thin replacement of ``Trainer.fit``.
"""
import torch
import torch.nn as nn

from .architectures import get_Architecture
from .distortions import prediction_loss
from .helpers import NamespaceMap, Timer, append_optimizer_scheduler_

__all__ = ["Predictor"]


class Predictor(nn.Module):
    """Main network for downstream prediction."""

    def __init__(self, hparams, featurizer=None):
        super().__init__()
        self.hparams = hparams if isinstance(hparams, NamespaceMap) else NamespaceMap(hparams)
        self.is_clf = self.hparams.data.target_is_clf
        self.logged, self.global_step, self.current_epoch = {}, 0, 0
        if featurizer is not None:
            featurizer.set_featurize_mode_()  # frozen, not trained with the predictor
            featurizer.stage = "predfeat"
            self.featurizer = featurizer
            pred_in_shape = featurizer.out_shape
        else:
            self.featurizer = nn.Identity()
            pred_in_shape = self.hparams.data.shape
        self.normalizer = nn.Identity()  # image re-normalisation only applies to reconstructions, which are out of scope
        cfg_pred = self.hparams.predictor
        self.predictor = get_Architecture(cfg_pred.arch, **cfg_pred.arch_kwargs)(pred_in_shape, self.hparams.data.target_shape)
        self.stage = self.hparams.stage

    def forward(self, x, is_logits=True, is_return_logs=False):
        with torch.no_grad():
            features = self.normalizer(self.featurizer(x))  # z_hat
        features = features.detach()
        with Timer() as inference_timer:
            Y_pred = self.predictor(features)
        out = Y_pred.softmax(-1) if (not is_logits and self.is_clf) else Y_pred
        if is_return_logs:
            return out, dict(inference_time=inference_timer.duration / Y_pred.size(0))
        return out

    def predict_step(self, batch, batch_idx=0, dataloader_idx=None):
        x, y = batch
        return self(x).cpu(), y.cpu()

    def step(self, batch):
        x, y = batch
        Y_hat, logs = self(x, is_return_logs=True)
        loss, loss_logs = self.loss(Y_hat, y)
        loss = loss.mean()
        logs.update(loss_logs)
        logs["loss"] = loss
        if self.is_clf:
            logs["acc"] = (Y_hat.argmax(dim=-1) == y).float().mean()
            logs["err"] = 1 - logs["acc"]
        return loss, logs

    def loss(self, Y_hat, y):
        """MSE or cross-entropy, plus the per-task aggregates the reference logs."""
        loss = prediction_loss(Y_hat, y, self.is_clf, agg_over_tasks="mean")
        logs = {}
        for agg in ["max", "std", "min", "mean", "median"]:
            logs[f"tasks_{agg}"] = prediction_loss(Y_hat.detach(), y, self.is_clf, agg_over_tasks=agg).mean()
        return loss, logs

    def log_dict(self, d, **kw):
        self.logged.update({k: (v.detach() if torch.is_tensor(v) else v) for k, v in d.items()})

    def training_step(self, batch, batch_idx=0):
        loss, logs = self.step(batch)
        self.log_dict({f"train/{self.stage}/{k}": v for k, v in logs.items()}, sync_dist=True)
        return loss

    def validation_step(self, batch, batch_idx=0):
        loss, logs = self.step(batch)
        self.log_dict({f"val/{self.stage}/{k}": v for k, v in logs.items()}, on_epoch=True, on_step=False, sync_dist=True)
        return loss

    def test_step(self, batch, batch_idx=0):
        loss, logs = self.step(batch)
        self.log_dict({f"test/{self.stage}/{k}": v for k, v in logs.items()}, on_epoch=True, on_step=False, sync_dist=True)
        return loss

    def on_test_epoch_start(self):
        if hasattr(self.featurizer, "on_test_epoch_start"):
            self.featurizer.on_test_epoch_start()

    def configure_optimizers(self):
        optimizers, schedulers = [], []
        append_optimizer_scheduler_(self.hparams.optimizer_pred, self.hparams.get("scheduler_pred"), self.predictor.parameters(),
                                    optimizers, schedulers, name="lr_predictor")
        return optimizers, schedulers

    def fit(self, batches, epochs=1):
        """Lightning-free ``Trainer.fit``: ``epochs`` passes over ``batches`` (an iterable of ``(x, y)``; re-iterated
        per epoch), schedulers stepped per epoch; returns the logged scalars as floats."""
        state = getattr(self, "_trainer_state", None)
        if state is None:
            state = self._trainer_state = self.configure_optimizers()
        optimizers, schedulers = state
        self.train()
        for _ in range(epochs):
            for i, batch in enumerate(batches):
                for opt in optimizers:
                    opt.zero_grad(set_to_none=True)
                    self.training_step(batch, i).backward()
                    opt.step()
                self.global_step += 1
            for s in schedulers:
                s["scheduler"].step()
            self.current_epoch += 1
        self.logged = {k: (float(v) if torch.is_tensor(v) else v) for k, v in self.logged.items()}
        return self.logged

    @torch.no_grad()
    def evaluate(self, batches):
        """Mean test loss / accuracy over ``batches`` (``test_step`` semantics)."""
        self.eval()
        self.on_test_epoch_start()
        n, acc, loss = 0, 0.0, 0.0
        for i, (x, y) in enumerate(batches):
            l, logs = self.step((x, y))
            n += x.shape[0]
            loss += float(l) * x.shape[0]
            acc += float(logs.get("acc", 0.0)) * x.shape[0]
        return dict(loss=loss / max(n, 1), acc=acc / max(n, 1), err=1 - acc / max(n, 1))
