"""Bench / test HARNESS model: a CIFAR-style ResNet-18 in plain PyTorch (not product code).

Same state-dict names, shapes, order and counts as the reference's Classification/models/resnet.py ResNet18
(11,173,962 parameters in 62 tensors; checked by tests/test_harness_models.py): 3x3 stem, four stages of two basic blocks
(64-128-256-512), 1x1 projection shortcuts where the shape changes, global average pool, 10-way
classifier.  Written for the end-to-end measurement of BASELINE config 1; the forward/backward of
the reference's models stays in PyTorch and is outside the hot path's scope.
"""
import torch.nn as nn
import torch.nn.functional as F


class _Block(nn.Module):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + self.shortcut(x))


class ResNet18Harness(nn.Module):
    def __init__(self, num_classes=10):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        cin = 64
        for i, (cout, stride) in enumerate(((64, 1), (128, 2), (256, 2), (512, 2)), start=1):
            setattr(self, f"layer{i}", nn.Sequential(_Block(cin, cout, stride), _Block(cout, cout, 1)))
            cin = cout
        self.linear = nn.Linear(512, num_classes)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.linear(F.adaptive_avg_pool2d(x, 1).flatten(1))


if __name__ == "__main__":
    m = ResNet18Harness()
    print(sum(p.numel() for p in m.parameters()), len(list(m.parameters())))
