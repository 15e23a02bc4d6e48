"""The one Lightning callback that sits on the training hot path: ``Freezer``
(reference lossyless/callbacks.py:503-531, enabled by config/finetune/freezer.yaml with
``model_name: p_ZlX.mapper`` -- the reference's shipped CLIP recipe, bin/clip/clip_hub.sh).  Same
constructor and hook names; ``LearnableCompressor.fit(callbacks=[...])`` calls the hooks where Lightning's
``BaseFinetuning`` would (before the optimisers are configured)."""
import torch.nn as nn

__all__ = ["Freezer"]


class Freezer:
    """Freeze an entire sub-module of the compressor.

    Parameters
    ----------
    model_name : str
        Name of the module to freeze from the (pl) module. Can use dots, e.g. ``"p_ZlX.mapper"``.
    """

    def __init__(self, model_name):
        self.model_name = model_name.split(".")

    def get_model(self, pl_module):
        model = pl_module
        for name in self.model_name:
            model = getattr(model, name)
        return model

    def freeze_before_training(self, pl_module):
        """``BaseFinetuning.freeze(modules=model, train_bn=False)``: no parameter of the module requires
        grad any more and its normalisation layers run in eval mode.  A CLIP tower that was unfrozen
        (``VisionTransformer.unfreeze_``) goes back to the single-call frozen runner."""
        model = self.get_model(pl_module)
        for m in model.modules():
            if hasattr(m, "freeze_") and hasattr(m, "unfreeze_"):
                m.freeze_()
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.eval()
        for p in model.parameters():
            p.requires_grad_(False)

    def finetune_function(self, pl_module, current_epoch, optimizer, optimizer_idx):
        pass
