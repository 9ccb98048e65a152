"""Oracle restatement of the MedicalNet 3D-ResNet encoder (test infrastructure).

MedicalNet (Tencent/MedicalNet, `models/resnet.py`, `model.py`, `setting.py`) is an UN-VENDORED, UN-PINNED
third-party dependency of the reference (README.md:68-88; call sites pkg/models/mri_models/anat_cnn.py:4-5,18-31,
pkg/models/pet_models/pet_resnet_cnn.py:4-5,23-35).  It is absent from /root/reference, so its published
architecture is restated here with the defaults the reference's `parse_opts()` call implies:
shortcut type 'B', conv1 = Conv3d(1, 64, k=7, stride=2, pad=3, bias=False), MaxPool3d(3, 2, 1), layer2 stride 2,
layer3 dilation 2, layer4 dilation 4 (total stride 8), Conv3d init kaiming_normal_(fan_out), BN weight=1 bias=0.

Pinned by the only facts the reference records about it (pkg/utils/outdated/inspect_model.py:100,105,284-285):
ResNet-50 -> 2048 channels; 91x109x91 -> 12x14x12; 159 base parameter tensors (tests/test_oracle.py).
"""
import torch
import torch.nn as nn


def conv3x3x3(in_planes, out_planes, stride=1, dilation=1):
    return nn.Conv3d(in_planes, out_planes, kernel_size=3, dilation=dilation, stride=stride, padding=dilation,
                     bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3x3(inplanes, planes, stride=stride, dilation=dilation)
        self.bn1 = nn.BatchNorm3d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = conv3x3x3(planes, planes, dilation=dilation)
        self.bn2 = nn.BatchNorm3d(planes)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv3d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm3d(planes)
        self.conv2 = nn.Conv3d(planes, planes, kernel_size=3, stride=stride, dilation=dilation, padding=dilation,
                               bias=False)
        self.bn2 = nn.BatchNorm3d(planes)
        self.conv3 = nn.Conv3d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm3d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward(self, x):
        residual = x
        out = self.relu(self.bn1(self.conv1(x)))
        out = self.relu(self.bn2(self.conv2(out)))
        out = self.bn3(self.conv3(out))
        if self.downsample is not None:
            residual = self.downsample(x)
        out = out + residual
        return self.relu(out)


class ResNet(nn.Module):
    def __init__(self, block, layers, shortcut_type="B"):
        super().__init__()
        assert shortcut_type == "B"
        self.inplanes = 64
        self.conv1 = nn.Conv3d(1, 64, kernel_size=7, stride=(2, 2, 2), padding=(3, 3, 3), bias=False)
        self.bn1 = nn.BatchNorm3d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool3d(kernel_size=(3, 3, 3), stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)
        # segmentation head of upstream; every reference model REPLACES it (anat_cnn.py:79)
        self.conv_seg = nn.Identity()
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(
                nn.Conv3d(self.inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False),
                nn.BatchNorm3d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.conv_seg(x)


_DEPTHS = {10: (BasicBlock, [1, 1, 1, 1]), 18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]),
           50: (Bottleneck, [3, 4, 6, 3])}


def generate_model(model_depth):
    """Encoder the reference obtains through `generate_model(opts)[0].module` (anat_cnn.py:30-31)."""
    if model_depth not in _DEPTHS:
        raise ValueError("hparams['resnet_depth'] is not in [10, 18, 34, 50]")
    block, layers = _DEPTHS[model_depth]
    return ResNet(block, layers)


def feature_width(model_depth):
    """n_in of the reference heads (anat_cnn.py:37-46)."""
    if model_depth in (10, 18):
        return 512
    if model_depth == 50:
        return 2048
    raise ValueError("hparams['resnet_depth'] is not in [10, 18, 34, 50]")
