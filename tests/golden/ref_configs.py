"""Configurations shared by tests/golden/make_golden_ref.py (which runs the reference on them) and the tests that
replay the resulting fixtures through the oracle and the CUDA path.  Plain data, no reference imports."""
import torch

CONTRASTIVE_KW = dict(temperature=0.01, is_train_temperature=True, is_cosine=True, effective_batch_size=5000,
                      is_already_featurized=True, is_project=True,
                      project_kwargs=dict(mode="mlp", out_shape=0.5, in_shape=512))



def step_hparams(distortion, z_dim=16, x_dim=24):
    if distortion == "lossy_Z":
        dist = dict(mode="lossy_Z", factor_beta=1, kwargs=dict(p_norm=1))
    elif distortion == "contrastive":
        dist = dict(mode="contrastive", factor_beta=0.003,
                    kwargs=dict(is_train_temperature=True, temperature=0.01, effective_batch_size=5000, is_project=True, is_cosine=True,
                                project_kwargs=dict(mode="mlp", out_shape=0.5, in_shape=z_dim)))
    else:
        dist = dict(mode="direct", factor_beta=1,
                    kwargs=dict(z_dim=z_dim, y_shape=(x_dim,), is_classification=False, dataset="banana", is_normalized=False, data_mode="distribution",
                                arch="mlp", arch_kwargs=dict(hid_dim=32, n_hid_layers=1, activation="Softplus")))
    return dict(
        stage="feat", experiment="golden",
        data=dict(shape=(x_dim,), mode="distribution", max_steps=100, length=5000, target_is_clf=False, target_shape=(x_dim,),
                  aux_is_clf=False, aux_shape=(x_dim,)),
        encoder=dict(z_dim=z_dim, fam="deterministic", fam_kwargs={}, arch="mlp", arch_kwargs=dict(hid_dim=32, n_hid_layers=2, activation="Softplus")),
        rate=dict(mode="H_factorized", factor_beta=1,
                  kwargs=dict(z_dim=z_dim, is_endToEnd=(distortion != "lossy_Z"), warmup_k_epoch=0,
                              kwargs_ent_bottleneck=dict(init_scale=10, filters=[3, 3, 3, 3]))),
        distortion=dist,
        featurizer=dict(is_features=True, loss=dict(beta=0.05, beta_anneal="linear")),
        evaluation=dict(featurizer=dict(is_online=False)),
    )



class SeededImages(torch.utils.data.Dataset):
    """N normalised 3x224x224 'images' + a label, regenerated from the index (the tests rebuild the same)."""

    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(1000 + i)
        return torch.randn(3, 224, 224, generator=g), i % 7


# lossyless/predictors.py Predictor on 512-d representations (ref_predictor.npz)
PREDICTOR_CASES = {
    "linear_clf": dict(arch="linear", arch_kwargs={}, is_clf=True, target_shape=[10]),
    "mlp_clf": dict(arch="mlp", arch_kwargs=dict(hid_dim=64, n_hid_layers=2, norm_layer="batch"), is_clf=True, target_shape=[10]),
    "mlp_reg": dict(arch="mlp", arch_kwargs=dict(hid_dim=32, n_hid_layers=1), is_clf=False, target_shape=[3]),
}


def predictor_hparams(case):
    c = PREDICTOR_CASES[case]
    return dict(data=dict(target_is_clf=c["is_clf"], shape=[512], target_shape=c["target_shape"], balancing_weights={}),
                predictor=dict(arch=c["arch"], arch_kwargs=dict(c["arch_kwargs"])), stage="pred",
                optimizer_pred=dict(mode="Adam", kwargs=dict(lr=1e-3)), scheduler_pred=dict(modes=[], kwargs={}))


# config/main.yaml's optimiser / scheduler groups (optimizer/AdamW_lr1e-3_w1e-5, optimizer/AdamW_lr3e-4_w1e-5,
# scheduler/expdecay100, scheduler/unifmultistep100) on top of step_hparams (ref_optim.npz)
def optim_hparams(distortion, scheduler, epochs):
    hp = step_hparams(distortion)
    if scheduler == "expdecay100":
        sch = dict(modes=["expdecay"], kwargs=dict(expdecay=dict(decay_factor=100, epochs=epochs)))
    else:
        sch = dict(modes=["UniformMultiStepLR"], kwargs=dict(UniformMultiStepLR=dict(decay_factor=100, epochs=epochs, k_steps=3)))
    hp.update(optimizer_feat=dict(mode="AdamW", kwargs=dict(lr=1e-3, weight_decay=1e-5)), scheduler_feat=sch,
              optimizer_coder=dict(mode="AdamW", kwargs=dict(lr=3e-4, weight_decay=1e-5)),
              scheduler_coder=dict(modes=list(sch["modes"]), kwargs={k: dict(v) for k, v in sch["kwargs"].items()}))
    return hp

