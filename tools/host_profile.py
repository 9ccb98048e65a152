"""cProfile of the host-side step issue path (small batch so that the GPU is not the limiter)."""
import cProfile, pstats, sys, time
import torch
sys.path.insert(0, ".")
import bench
from multimodal_alzheimer_b200.pkg.models.fusion_models.anat_pet_fusion import Anat_PET_CNN, ResNet_PET_Trunk
from multimodal_alzheimer_b200.pkg.models.mri_models.anat_cnn import Anat_CNN
from multimodal_alzheimer_b200.pkg.models.pet_models.pet_resnet_cnn import PET_CNN_ResNet
from multimodal_alzheimer_b200.pkg.utils import normalization as norm

dev = torch.device("cuda:0")
enc, fus = bench.hparams_for("pet_mri_fusion_r18", 18)
model = Anat_PET_CNN(dict(fus), model_mri=Anat_CNN(dict(enc)), pet_trunk=ResNet_PET_Trunk(PET_CNN_ResNet(dict(enc))))
model.to(dev).train()
params = [p for p in model.parameters() if p.requires_grad]
opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-4, fused=True)
data = bench.synth_inputs(2, 64, dev, 15)


def step():
    mri = norm.normalize_mri_per_scan_min_max(data["mri_raw"], data["mask"], 0.98, out_dtype=torch.bfloat16)
    pet = norm.normalize_pet(data["pet_raw"], 0.5145, 0.5383, out_dtype=torch.bfloat16)
    out = model.general_step({"mri": mri, "pet1451": pet, "label": data["label"]}, 0, "train")
    out["loss"].backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e3 * (t1 - t0) / 5:.2f} ms/step, incl. drain {1e3 * (t2 - t0) / 5:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
