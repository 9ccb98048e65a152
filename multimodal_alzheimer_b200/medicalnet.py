"""B200-native MedicalNet 3D-ResNet encoder (ResNet-10/18/34/50, shortcut type B, dilated layer3/4).

Drop-in for what the reference obtains from `MedicalNet.model.generate_model(opts)[0].module`
(pkg/models/mri_models/anat_cnn.py:18-31, pkg/models/pet_models/pet_resnet_cnn.py:23-35): same module tree and
state_dict keys (`conv1.weight`, `bn1.*`, `layer{1-4}.{i}.{conv,bn}{1-3}.*`, `layer*.0.downsample.{0,1}.*`,
`conv_seg.*`), same random init (Conv3d kaiming_normal_(fan_out), BN weight=1/bias=0), fp32 NCDHW parameters.
The arithmetic is the tcgen05 implicit-GEMM conv + fused BatchNorm/ReLU/residual kernels of libadni_b200.so; each
residual block runs as one autograd Function (multimodal_alzheimer_b200/autograd.py).
"""
import torch.nn as tnn

from . import autograd as A
from . import kernels as K
from . import nn as bnn


def conv3x3x3(in_planes, out_planes, stride=1, dilation=1):
    return bnn.Conv3d(in_planes, out_planes, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)


def _ds_args(downsample):
    if downsample is None:
        return None, None, None, None, None
    conv, bn = downsample[0], downsample[1]
    return conv.weight, bn.weight, bn.bias, bn.state(), conv.cfg


def _with_tail(out, y_last, bnp_last):
    """Tag a block's output with the description of the bn -> (+ residual) -> relu that produced it (raw conv output,
    BatchNorm parameters): the next block's last dgrad folds that BatchNorm's backward sums into its epilogue
    (autograd._block_input_grad).  Training-mode batches only (eval mode has no BatchNorm backward)."""
    if bnp_last.numel():
        out._adni_tail = (y_last, bnp_last)
    return out


def _tail_of(x):
    return getattr(x, "_adni_tail", (None, None))


class BasicBlock(tnn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = conv3x3x3(inplanes, planes, stride=stride, dilation=dilation)
        self.bn1 = bnn.BatchNorm3d(planes)
        self.relu = bnn.ReLU(inplace=True)
        self.conv2 = conv3x3x3(planes, planes, dilation=dilation)
        self.bn2 = bnn.BatchNorm3d(planes)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward(self, x):
        wd, gd, bd, bnd, cd = _ds_args(self.downsample)
        tail_y, tail_p = _tail_of(x)
        out, y2, p2 = A.BasicBlockFn.apply(bnn.as_volume(x), self.conv1.weight, self.bn1.weight, self.bn1.bias,
                                           self.conv2.weight, self.bn2.weight, self.bn2.bias, wd, gd, bd, self.bn1.state(),
                                           self.bn2.state(), bnd, self.conv1.cfg, self.conv2.cfg, cd, tail_y, tail_p)
        return _with_tail(out, y2, p2)


class Bottleneck(tnn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, dilation=1, downsample=None):
        super().__init__()
        self.conv1 = bnn.Conv3d(inplanes, planes, 1, bias=False)
        self.bn1 = bnn.BatchNorm3d(planes)
        self.conv2 = bnn.Conv3d(planes, planes, 3, stride=stride, dilation=dilation, padding=dilation, bias=False)
        self.bn2 = bnn.BatchNorm3d(planes)
        self.conv3 = bnn.Conv3d(planes, planes * 4, 1, bias=False)
        self.bn3 = bnn.BatchNorm3d(planes * 4)
        self.relu = bnn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride
        self.dilation = dilation

    def forward(self, x):
        wd, gd, bd, bnd, cd = _ds_args(self.downsample)
        tail_y, tail_p = _tail_of(x)
        out, y3, p3 = A.BottleneckFn.apply(bnn.as_volume(x), self.conv1.weight, self.bn1.weight, self.bn1.bias,
                                           self.conv2.weight, self.bn2.weight, self.bn2.bias, self.conv3.weight,
                                           self.bn3.weight, self.bn3.bias, wd, gd, bd, self.bn1.state(), self.bn2.state(),
                                           self.bn3.state(), bnd, self.conv1.cfg, self.conv2.cfg, self.conv3.cfg, cd,
                                           tail_y, tail_p)
        return _with_tail(out, y3, p3)


class ResNet(tnn.Module):
    def __init__(self, block, layers, shortcut_type="B"):
        super().__init__()
        if shortcut_type != "B":
            raise NotImplementedError("only MedicalNet shortcut type 'B' (the reference's parse_opts default)")
        self.inplanes = 64
        self.conv1 = bnn.Conv3d(1, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = bnn.BatchNorm3d(64)
        self.relu = bnn.ReLU(inplace=True)
        self.maxpool = bnn.MaxPool3d(3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0])
        self.layer2 = self._make_layer(block, 128, layers[1], stride=2)
        self.layer3 = self._make_layer(block, 256, layers[2], stride=1, dilation=2)
        self.layer4 = self._make_layer(block, 512, layers[3], stride=1, dilation=4)
        # upstream's segmentation head; every reference model replaces it (anat_cnn.py:79)
        self.conv_seg = bnn.Sequential()
        for m in self.modules():
            if isinstance(m, bnn.Conv3d):
                tnn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, bnn.BatchNorm3d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    def _make_layer(self, block, planes, blocks, stride=1, dilation=1):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = bnn.Sequential(
                bnn.Conv3d(self.inplanes, planes * block.expansion, 1, stride=stride, bias=False),
                bnn.BatchNorm3d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride=stride, dilation=dilation, downsample=downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, dilation=dilation))
        return bnn.Sequential(*layers)

    def stem(self, x):
        mp = self.maxpool
        return A.StemFn.apply(bnn.as_volume(x), self.conv1.weight, self.bn1.weight, self.bn1.bias, self.bn1.state(),
                              self.conv1.cfg, (mp.kernel_size, mp.stride, mp.padding))

    def _refresh_kernel_weights(self):
        """bf16 kernel-layout copies of every tensor-core conv weight of the encoder, in one launch."""
        if getattr(self, "_arena", None) is None:
            self._arena = K.WeightArena()
            K.register_arena(self._arena)
        self._arena.refresh([m.weight for m in self.modules() if isinstance(m, bnn.Conv3d) and m is not self.conv1])

    def features(self, x):
        self._refresh_kernel_weights()
        with A.defer_bn_counters():
            x = self.stem(x)
            return self.layer4(self.layer3(self.layer2(self.layer1(x))))

    def forward(self, x):
        return self.conv_seg(self.features(x))


_DEPTHS = {10: (BasicBlock, [1, 1, 1, 1]), 18: (BasicBlock, [2, 2, 2, 2]), 34: (BasicBlock, [3, 4, 6, 3]),
           50: (Bottleneck, [3, 4, 6, 3])}


def generate_model(model_depth):
    if model_depth not in _DEPTHS:
        raise ValueError("hparams['resnet_depth'] is not in [10, 18, 34, 50]")
    block, layers = _DEPTHS[model_depth]
    return ResNet(block, layers)
