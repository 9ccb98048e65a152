"""PET_CNN_ResNet — the MedicalNet ResNet classifier on PET volumes
(reference pkg/models/pet_models/pet_resnet_cnn.py:12-166): same graph as Anat_CNN, reads batch['pet1451']."""
from ..base_model import adam_or_plateau
from ..mri_models.anat_cnn import Anat_CNN


class PET_CNN_ResNet(Anat_CNN):
    modality = "pet1451"

    def configure_optimizers(self):
        # pet_resnet_cnn.py:149-166: identical grouping, no weight decay argument
        parameters_optim = []
        for name, param in self.model.named_parameters():
            if "conv_seg" in name:
                parameters_optim.append({"params": param, "lr": self.hparams["lr"]})
            elif "lr_pretrained" not in self.hparams or not self.hparams["lr_pretrained"]:
                param.requires_grad = False
                parameters_optim.append({"params": param})
            else:
                param.requires_grad = True
                parameters_optim.append({"params": param, "lr": self.hparams["lr_pretrained"]})
        return adam_or_plateau(self.hparams, parameters_optim)
