"""Import the UNMODIFIED reference LightningModules (/root/reference/pkg/models/**) in this container.

The reference classes need packages that are not installed here and cannot be (no network): pytorch_lightning,
torchmetrics, MedicalNet (external clone), tabpfn, matplotlib, seaborn, nibabel.  This module puts minimal stand-ins
for exactly those imports into `sys.modules` and then imports the reference's own files from where they lie, so the
class bodies that run - `__init__`, `forward`, `general_step`, `configure_optimizers`, the truncation / freezing logic,
the TabPFN activation hook + `get_avg_activation` - are the reference's code, byte for byte.

What the stand-ins do (each is logging / plumbing, not path arithmetic, except MedicalNet):
  pytorch_lightning.LightningModule   nn.Module + save_hyperparameters / hparams / log / log_dict, and
                                      `load_from_checkpoint(path, **kw)` resolved from an in-memory registry
                                      {path: (hparams, state_dict)} with Lightning's semantics: cls(hparams, **kw)
                                      followed by a strict load_state_dict.
  torchmetrics(.classification)       metric objects that accept updates and do nothing.
  MedicalNet.model.generate_model     returns (wrapper, params) with wrapper.module = oracle.medicalnet's ResNet:
                                      MedicalNet is a third-party clone absent from /root/reference (SURVEY.md
                                      App. A restates it); this is the one part of the graph that is NOT the
                                      reference's own bytes.
  tabpfn.TabPFNClassifier             a fake classifier whose `model[2].decoder[0]` is an nn.Identity and whose
                                      `predict_proba(x)` pushes [training rows ; x] replicated over the ensemble axis,
                                      shape (training_size + B, ensemble, 1024), through it - the hook, slicing,
                                      averaging and transposition that follow are the reference's.
  matplotlib / seaborn / nibabel / pkg.utils.dataloader / ...data_preparation   empty modules (plotting, file I/O).

Only tools/make_golden_models.py uses this (generation time, build container).  Nothing under tests/, bench.py or the
package imports it: /root/reference does not exist on the GPU box.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECKPOINTS = {}          # path -> (hparams dict, state_dict)
TRAINING_SIZE = 5         # rows the fake TabPFN was "fitted" on
ENSEMBLE_SEEN = []


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


class LightningModule(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()
        self._hp = _AttrDict()
        self.logged = {}
        self.current_epoch = 0

    def save_hyperparameters(self, *args, ignore=None, **kw):
        for a in args:
            if isinstance(a, dict):
                self._hp.update(a)

    @property
    def hparams(self):
        return self._hp

    def log(self, name, value, **kw):
        self.logged[name] = value

    def log_dict(self, d, **kw):
        self.logged.update(d)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, **kwargs):
        hparams, state = CHECKPOINTS[checkpoint_path]
        model = cls(dict(hparams), **kwargs)
        model.load_state_dict(state, strict=True)
        return model


def register_checkpoint(path, model):
    """What `Trainer.save_checkpoint` keeps of a module: its hparams and its state_dict."""
    CHECKPOINTS[path] = (dict(model.hparams), {k: v.detach().clone() for k, v in model.state_dict().items()})


class _Metric(nn.Module):
    def __init__(self, *a, **kw):
        super().__init__()

    def forward(self, *a, **kw):
        return None

    def compute(self):
        return torch.zeros(())

    def reset(self):
        pass


class _Identity(nn.Identity):
    pass


class _FakeTransformer(nn.Module):
    def __init__(self):
        super().__init__()
        self.decoder = nn.Sequential(_Identity(), nn.GELU(), nn.Linear(1024, 10))


class TabPFNClassifier:
    """Stand-in with the attribute paths the reference touches: `.model[2].decoder[0]`, `.fit`, `.predict_proba`."""

    def __init__(self, device="cpu", N_ensemble_configurations=4, **kw):
        self.n_ens = N_ensemble_configurations
        ENSEMBLE_SEEN.append(N_ensemble_configurations)
        self.model = (None, None, _FakeTransformer())

    def fit(self, x, y, overwrite_warning=False):
        self.x_train = x
        return self

    def predict_proba(self, x, normalize_with_test=False, **kw):
        x = torch.as_tensor(x, dtype=torch.float32)
        g = torch.Generator().manual_seed(1451)
        train_rows = torch.randn((TRAINING_SIZE, x.shape[1]), generator=g)
        seq = torch.cat((train_rows, x), dim=0)                      # (training_size + B, 1024)
        seq = seq.unsqueeze(1).repeat(1, self.n_ens, 1)              # (S, ensemble, 1024) like the TabPFN decoder input
        with torch.no_grad():
            self.model[2].decoder[0](seq)                            # fires the reference's forward hook
        return torch.zeros((x.shape[0], 2)).numpy()


def _get_data(path, binary_classification=True):
    g = torch.Generator().manual_seed(7)
    return torch.randn((TRAINING_SIZE, 9), generator=g).numpy(), torch.zeros(TRAINING_SIZE).numpy()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Install the stand-ins and make `import pkg...` resolve to /root/reference/pkg."""
    if "pytorch_lightning" in sys.modules and getattr(sys.modules["pytorch_lightning"], "_adni_stub", False):
        return
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import medicalnet as omn

    _module("pytorch_lightning", LightningModule=LightningModule, _adni_stub=True)
    tm = _module("torchmetrics")
    tmc = _module("torchmetrics.classification", MulticlassF1Score=_Metric, MulticlassMatthewsCorrCoef=_Metric)
    tm.classification = tmc
    tm.ConfusionMatrix = _Metric
    mpl = _module("matplotlib")
    mpl.pyplot = _module("matplotlib.pyplot")
    mpl.figure = _module("matplotlib.figure", Figure=object)
    mpl.colors = _module("matplotlib.colors", LinearSegmentedColormap=object)
    _module("seaborn")
    _module("nibabel")
    _module("tabpfn", TabPFNClassifier=TabPFNClassifier)

    class _Wrapper:                                                   # generate_model returns nn.DataParallel(model)
        def __init__(self, module):
            self.module = module

    def generate_model(opts):
        return _Wrapper(omn.generate_model(opts.model_depth)), None

    mn = _module("MedicalNet")
    mn.model = _module("MedicalNet.model", generate_model=generate_model)
    mn.setting = _module("MedicalNet.setting", parse_opts=lambda: types.SimpleNamespace())

    if not os.environ.get("CUDA_VISIBLE_DEVICES"):                   # anat_cnn.py:20-24 raises without it
        os.environ["CUDA_VISIBLE_DEVICES"] = "0"

    # the reference package, from where it lies (namespace package: the reference has no __init__.py files)
    for k in [k for k in sys.modules if k == "pkg" or k.startswith("pkg.")]:
        del sys.modules[k]
    sys.path.insert(0, REF)
    import numpy as np
    _module("pkg.utils.dataloader", MultiModalDataset=object)
    # data_preparation.py:16 defines TRAINPATH; tabular_mri_fusion.py:22 reads it through two `import *` hops
    _module("pkg.models.tabular_models.data_preparation", get_data=_get_data, np=np, TRAINPATH="train_path_data_labels.csv")
    importlib.invalidate_caches()


def reference_classes():
    install()
    from pkg.models.fusion_models.all_modalities_fusion import All_Modalities_Fusion
    from pkg.models.fusion_models.anat_pet_featuremapfusion import PET_MRI_FMF
    from pkg.models.fusion_models.anat_pet_fusion import Anat_PET_CNN
    from pkg.models.fusion_models.early_fusion import PET_MRI_EF
    from pkg.models.fusion_models.pet_tabular_fusion import PET_TABULAR_CNN
    from pkg.models.fusion_models.tabular_mri_fusion import Tabular_MRT_Model
    from pkg.models.mri_models.anat_cnn import Anat_CNN
    from pkg.models.pet_models.pet_cnn import Small_PET_CNN
    from pkg.models.pet_models.pet_resnet_cnn import PET_CNN_ResNet
    return dict(Anat_CNN=Anat_CNN, Small_PET_CNN=Small_PET_CNN, PET_CNN_ResNet=PET_CNN_ResNet,
                Anat_PET_CNN=Anat_PET_CNN, Tabular_MRT_Model=Tabular_MRT_Model, PET_TABULAR_CNN=PET_TABULAR_CNN,
                All_Modalities_Fusion=All_Modalities_Fusion, PET_MRI_EF=PET_MRI_EF, PET_MRI_FMF=PET_MRI_FMF)


if __name__ == "__main__":
    for k, v in reference_classes().items():
        print(k, v.__module__, v.__mro__[1].__name__)
