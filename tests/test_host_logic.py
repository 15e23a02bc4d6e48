"""CPU: host-side mirror of the reference interface -- hub entry points, state-dict contract,
container helpers, device checks (no kernels are launched)."""
import io
import os
import struct

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_hubconf_surface(native_lib):
    import hubconf

    assert {"clip_compressor_b001", "clip_compressor_b005", "clip_compressor_b01"} <= set(dir(hubconf))
    assert isinstance(hubconf.dependencies, list)
    comp, transform = hubconf.clip_compressor_b005(device="cpu")
    assert comp.z_dim == 512 and comp.side_z_dim == 102 and comp.device == "cpu"
    assert comp.scaling.shape == (512,) and comp.biasing.shape == (512,)
    assert comp.entropy_bottleneck._quantized_cdf.shape == (512, 32)
    assert not comp.training and callable(transform)
    for m in ("compress", "decompress", "get_rate", "compress_dataset", "decompress_dataset", "process_z_in",
              "process_z_out", "make_pickable_", "undo_pickable_", "forward"):
        assert callable(getattr(comp, m))
    # exp(scaling) ranges quoted in SURVEY 8a-2 for b005
    es = comp.scaling.exp()
    assert 4.9 < es.min() < es.max() < 6.5
    # reference behaviour: compressing on CPU is refused
    with pytest.raises(ValueError, match="only implemented on GPU"):
        comp.compress_dataset([], "/tmp/none.bin")
    # CLIP key names: a visual-tower checkpoint loads
    keys = set(comp.clip.state_dict())
    assert {"conv1.weight", "class_embedding", "positional_embedding", "proj", "ln_pre.weight", "ln_post.bias",
            "transformer.resblocks.11.attn.in_proj_weight", "transformer.resblocks.0.attn.out_proj.bias",
            "transformer.resblocks.3.mlp.c_fc.weight", "transformer.resblocks.3.mlp.c_proj.bias",
            "transformer.resblocks.5.ln_1.weight", "transformer.resblocks.5.ln_2.bias"} <= keys
    assert sum(v.numel() for v in comp.clip.state_dict().values()) == 87849216
    # process_z_in / out are inverse affine maps with the reference's shapes
    z = torch.randn(3, 512)
    assert comp.process_z_in(z).shape == (3, 512, 1, 1)
    assert torch.allclose(comp.process_z_out(comp.process_z_in(z)), z, atol=1e-5)


def test_preprocess_transform_matches_clip_recipe():
    from PIL import Image

    from lossyless_b200.clip_vit import CLIP_MEAN, CLIP_STD, clip_preprocess

    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 256, (96, 96, 3), dtype=np.uint8))  # STL10-shaped
    x = clip_preprocess(224)(img)
    assert x.shape == (3, 224, 224) and x.dtype == torch.float32
    lo = (0 - np.array(CLIP_MEAN)) / np.array(CLIP_STD)
    hi = (1 - np.array(CLIP_MEAN)) / np.array(CLIP_STD)
    for c in range(3):
        assert lo[c] - 1e-4 <= x[c].min() and x[c].max() <= hi[c] + 1e-4


def test_container_helpers_roundtrip():
    from lossyless_b200.hub.compressor import read_bytes, read_uints, write_bytes, write_uints

    f = io.BytesIO()
    write_uints(f, (3,))
    for b in (b"abcd", b"", b"12345678"):
        write_uints(f, (len(b),))
        write_bytes(f, b)
    assert f.getvalue()[:4] == struct.pack(">I", 3)
    f.seek(0)
    n = read_uints(f, 1)[0]
    got = [read_bytes(f, read_uints(f, 1)[0]) if True else None for _ in range(n)]
    assert got == [b"abcd", b"", b"12345678"]


def test_vit_requires_cuda(native_lib):
    from lossyless_b200._native import NativeError
    from lossyless_b200.clip_vit import VisionTransformer

    vit = VisionTransformer(224, 32, 768, 2, 12, 512)
    with pytest.raises(NativeError, match="no CPU fallback"):
        vit(torch.zeros(1, 3, 224, 224))


def test_downstream_linear_eval_runs():
    """tools/z_linear_eval.py (README.md:66-74 / utils/Z_linear_eval.py recipe) on separable synthetic Z."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("z_linear_eval", os.path.join(ROOT, "tools", "z_linear_eval.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    centers = rng.normal(size=(5, 512)) * 2
    Y = rng.integers(0, 5, 600)
    Z = (centers[Y] + rng.normal(size=(600, 512))).astype(np.float32)
    res = mod.linear_eval(Z[:400], Y[:400], Z[400:], Y[400:])
    assert res["test/pred/acc"] > 0.95 and abs(res["test/pred/err"] - (1 - res["test/pred/acc"])) < 1e-12


def _banana_hparams():
    from lossyless_b200.helpers import NamespaceMap

    return NamespaceMap(
        stage="feat", data=dict(shape=(2,), mode="distribution", max_steps=50, length=1024000),
        encoder=dict(z_dim=2, fam="deterministic", fam_kwargs={}, arch="mlp", arch_kwargs=dict(hid_dim=32, n_hid_layers=1, activation="Softplus")),
        rate=dict(mode="H_factorized", factor_beta=1, kwargs=dict(z_dim=2, kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))),
        distortion=dict(mode="direct", factor_beta=1, kwargs=dict(z_dim=2, y_shape=(2,), arch="mlp", arch_kwargs=dict(hid_dim=32, n_hid_layers=1, activation="Softplus"))),
        featurizer=dict(is_features=True, loss=dict(beta=0.07, beta_anneal="constant")),
        evaluation=dict(featurizer=dict(is_online=False)),
        optimizer_feat=dict(mode="Adam", kwargs=dict(lr=3e-3)), optimizer_coder=dict(mode="Adam", kwargs=dict(lr=3e-3)),
        scheduler_feat=None, scheduler_coder=None)


def test_checkpoint_layout_roundtrip(tmp_path):
    """save_checkpoint writes Lightning's layout (main.py:516-547: state_dict + hyper_parameters);
    load_from_checkpoint rebuilds the module from it, keyword overrides included."""
    import torch

    from lossyless_b200.learnable_compressors import LearnableCompressor

    torch.manual_seed(0)
    model = LearnableCompressor(_banana_hparams())
    model.global_step = 17
    path = tmp_path / "best_feat.ckpt"
    model.save_checkpoint(path)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert {"state_dict", "hyper_parameters", "global_step", "epoch", "hparams_name"} <= set(raw)
    assert raw["hparams_name"] == "hparams"  # Lightning passes ONE container to __init__(hparams) only with this key
    assert any(k.startswith("p_ZlX.mapper.") for k in raw["state_dict"])
    assert any(k.startswith("rate_estimator.entropy_bottleneck.") for k in raw["state_dict"])
    assert isinstance(raw["hyper_parameters"], dict) and type(raw["hyper_parameters"]["rate"]) is dict
    back = LearnableCompressor.load_from_checkpoint(path, stage="comm")
    assert back.global_step == 17 and back.stage == "comm"
    for (k, a), (_, b) in zip(model.state_dict().items(), back.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(KeyError):
        LearnableCompressor.load_from_checkpoint({"weights": {}})


def test_checkpoint_hparams_are_portable(tmp_path):
    """torch.dtype entries are stringified and an embedded CLIP state dict is dropped from 'hyper_parameters'."""
    import torch

    from lossyless_b200.learnable_compressors import _plain, _unplain

    hp = dict(encoder=dict(arch_kwargs=dict(compute_dtype=torch.bfloat16, clip_state_dict={"w": torch.zeros(3)}, is_frozen=True)), n=(1, 2))
    plain = _plain(hp)
    assert plain == dict(encoder=dict(arch_kwargs=dict(compute_dtype="torch.bfloat16", clip_state_dict=None, is_frozen=True)), n=(1, 2))
    assert _unplain(plain)["encoder"]["arch_kwargs"]["compute_dtype"] is torch.bfloat16


def test_fit_keeps_optimizer_state_and_steps_schedulers():
    """fit() = one epoch per call: optimisers / schedulers are built once, AdamW moments survive across calls, the
    schedulers step once per call and current_epoch advances (they used to be rebuilt on every call)."""
    import torch

    from lossyless_b200.learnable_compressors import LearnableCompressor

    hp = _banana_hparams()
    from lossyless_b200.helpers import NamespaceMap

    hp["scheduler_feat"] = NamespaceMap(modes=["expdecay"], kwargs=dict(expdecay=dict(epochs=10, decay_factor=100)))
    model = LearnableCompressor(hp)
    seen = {}

    def fake_step(self_, batch, batch_idx, optimizer_idx=0):  # a CPU-only loss over every trainable parameter
        ps = [p for g in seen["opt"][optimizer_idx].param_groups for p in g["params"]]
        return sum((p ** 2).sum() for p in ps)

    opts, scheds = model.configure_optimizers()
    model._trainer_state = dict(optimizers=opts, schedulers=scheds)
    seen["opt"] = opts
    model.training_step = fake_step.__get__(model)
    lr0 = opts[0].param_groups[0]["lr"]
    model.fit([None, None])
    assert model.current_epoch == 1 and model.global_step == 2
    st = opts[0].state[opts[0].param_groups[0]["params"][0]]
    assert int(st["step"]) == 2
    model.fit([None])
    assert model._trainer_state["optimizers"][0] is opts[0] and int(st["step"]) == 3 and model.current_epoch == 2
    assert len(scheds) == 1
    assert opts[0].param_groups[0]["lr"] == pytest.approx(lr0 * (1 / 100) ** (2 / 10))  # two epochs of the decay
    model.reset_optimizers_()
    assert model._trainer_state is None


def test_freezer_callback():
    """Freezer("p_ZlX.mapper") (config/finetune/freezer.yaml): the encoder leaves the optimiser's
    parameter set; on a CLIP tower it also switches back to the frozen runner."""
    from lossyless_b200.callbacks import Freezer
    from lossyless_b200.clip_vit import VisionTransformer
    from lossyless_b200.learnable_compressors import LearnableCompressor

    model = LearnableCompressor(_banana_hparams())
    n_all = len(model.get_specific_parameters("main"))
    cb = Freezer("p_ZlX.mapper")
    assert cb.get_model(model) is model.p_ZlX.mapper
    cb.freeze_before_training(model)
    assert all(not p.requires_grad for p in model.p_ZlX.mapper.parameters())
    opts, _ = model.configure_optimizers()
    n_opt = sum(len(g["params"]) for g in opts[0].param_groups)
    assert n_opt == n_all - len(list(model.p_ZlX.mapper.parameters()))

    vit = VisionTransformer(224, 32, 768, 1, 12, 512).unfreeze_()
    assert vit._trainable and vit.proj.dtype == torch.float32 and vit.proj.requires_grad
    holder = torch.nn.Module()
    holder.enc = vit
    Freezer("enc").freeze_before_training(holder)
    assert not vit._trainable and not any(p.requires_grad for p in vit.parameters())


def test_group_host_batches():
    """compress_dataset's grouping of DataLoader batches into encoder passes (hub/compressor.py): order kept,
    no group above the limit, dtype / shape changes and oversized batches close a group, 0 disables it."""
    from lossyless_b200.hub.compressor import group_host_batches

    def mk(n, dtype=torch.float16, hw=8):
        return torch.zeros(n, 3, hw, hw, dtype=dtype)

    sizes = lambda groups: [[int(x.shape[0]) for x in g] for g in groups]
    bs = [mk(100), mk(100), mk(100), mk(60), mk(300), mk(7)]
    assert sizes(group_host_batches(iter(bs), 256)) == [[100, 100], [100, 60], [300], [7]]
    assert sizes(group_host_batches(iter(bs), 0)) == [[100], [100], [100], [60], [300], [7]]
    assert sizes(group_host_batches(iter(bs), 1512)) == [[100, 100, 100, 60, 300, 7]]
    mixed = [mk(10), mk(10, torch.float32), mk(10, torch.float32), mk(10, torch.float32, hw=4), mk(10, torch.uint8)]
    assert sizes(group_host_batches(iter(mixed), 100)) == [[10], [10, 10], [10], [10]]
    flat = [x for g in group_host_batches(iter(bs), 256) for x in g]
    assert all(a is b for a, b in zip(flat, bs))
    assert list(group_host_batches(iter([]), 256)) == []


def test_predictor_probe_learns_on_representations():
    """lossyless/predictors.py Predictor on already-featurised inputs: a linear probe on 512-d Z with
    10 classes trains (cross-entropy down, accuracy up) and logs the reference's keys."""
    import torch

    from lossyless_b200.helpers import NamespaceMap
    from lossyless_b200.predictors import Predictor

    g = torch.Generator().manual_seed(0)
    centers = torch.randn(10, 512, generator=g)
    y = torch.arange(2000) % 10
    Z = centers[y] * 0.5 + torch.randn(2000, 512, generator=g)
    hp = NamespaceMap(stage="pred", data=dict(shape=(512,), target_shape=(10,), target_is_clf=True, balancing_weights={}),
                      predictor=dict(arch="linear", arch_kwargs={}),
                      optimizer_pred=dict(mode="Adam", kwargs=dict(lr=1e-2)), scheduler_pred=None)
    pred = Predictor(hp)
    batches = [(Z[i:i + 250], y[i:i + 250]) for i in range(0, 1500, 250)]
    first = float(pred.step(batches[0])[0])
    logged = pred.fit(batches, epochs=15)
    assert logged["train/pred/loss"] < 0.5 * first
    for k in ("loss", "acc", "err", "tasks_mean", "tasks_max", "inference_time"):
        assert f"train/pred/{k}" in logged, k
    ev = pred.evaluate([(Z[1500:], y[1500:])])
    assert ev["acc"] > 0.9, ev
    assert pred(Z[:4], is_logits=False).sum(-1).allclose(torch.ones(4))


def test_bind_host_to_gpu_reads_sysfs_topology(tmp_path, monkeypatch):
    """parallel.bind_host_to_gpu: PCI address from the device properties -> sysfs numa_node / local_cpulist -> CPU
    affinity restricted to the GPU's node (intersected with what the process may use); missing topology -> None."""
    import os
    import types

    import torch

    from lossyless_b200 import parallel

    assert parallel._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    before = os.sched_getaffinity(0)
    keep = set(sorted(before)[: max(1, len(before) // 2)])
    monkeypatch.setattr(torch.cuda, "get_device_properties",
                        lambda i: types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0))
    dev = tmp_path / "0000:1b:00.0"
    dev.mkdir()
    (dev / "numa_node").write_text("0\n")
    (dev / "local_cpulist").write_text(",".join(str(c) for c in sorted(keep)) + ",4095\n")  # 4095: not ours to use
    try:
        got = parallel.bind_host_to_gpu(0, sysfs=str(tmp_path))
        assert got == {"node": 0, "cpus": len(keep), "pci": "0000:1b:00.0"}
        assert os.sched_getaffinity(0) == keep
    finally:
        os.sched_setaffinity(0, before)
    (dev / "numa_node").write_text("-1\n")
    assert parallel.bind_host_to_gpu(0, sysfs=str(tmp_path)) is None
    assert parallel.bind_host_to_gpu(0, sysfs=str(tmp_path / "missing")) is None
    assert os.sched_getaffinity(0) == before
