"""timm.models.regnet.RegStage / Bottleneck restated from the published timm 1.0.3 algorithm (models/regnet.py,
layers/conv_bn_act.py, layers/norm_act.py, layers/squeeze_excite.py); see SURVEY.md §8c for the spec this follows.

Module / parameter names follow timm's (`b{i}.conv{1,2,3}.{conv,bn}`, `se.fc{1,2}`, `downsample.{conv,bn}`) because the
reference's checkpoints are keyed by them.  LN_EPS is a parameter: timm maps `norm_layer=LayerNorm2d` to
`LayerNormAct2d`, whose default eps is 1e-5 (plain LayerNorm2d: 1e-6)."""
import torch
import torch.nn as nn
import torch.nn.functional as F

LN_EPS = 1e-5


class LayerNormAct2d(nn.LayerNorm):
    def __init__(self, num_channels, eps=None, apply_act=True, act_layer=nn.ReLU):
        super().__init__(num_channels, eps=LN_EPS if eps is None else eps)
        self.act = act_layer() if (apply_act and act_layer is not None) else nn.Identity()

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        x = x.permute(0, 3, 1, 2)
        return self.act(x)


class ConvNormAct(nn.Module):
    def __init__(self, in_chs, out_chs, kernel_size=1, stride=1, groups=1, apply_act=True, act_layer=nn.ReLU):
        super().__init__()
        self.conv = nn.Conv2d(in_chs, out_chs, kernel_size, stride=stride, padding=kernel_size // 2, groups=groups,
                              bias=False)
        self.bn = LayerNormAct2d(out_chs, apply_act=apply_act, act_layer=act_layer)

    def forward(self, x):
        return self.bn(self.conv(x))


class SEModule(nn.Module):
    def __init__(self, channels, rd_channels, act_layer=nn.ReLU):
        super().__init__()
        self.fc1 = nn.Conv2d(channels, rd_channels, 1, bias=True)
        self.act = act_layer()
        self.fc2 = nn.Conv2d(rd_channels, channels, 1, bias=True)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.fc2(self.act(self.fc1(s)))
        return x * s.sigmoid()


class Bottleneck(nn.Module):
    def __init__(self, in_chs, out_chs, stride=1, bottle_ratio=1, group_size=1, se_ratio=0.25, act_layer=nn.ReLU,
                 norm_layer=None):
        super().__init__()
        b = int(round(out_chs * bottle_ratio))
        groups = b // group_size
        self.conv1 = ConvNormAct(in_chs, b, 1, act_layer=act_layer)
        self.conv2 = ConvNormAct(b, b, 3, stride=stride, groups=groups, act_layer=act_layer)
        if se_ratio:
            self.se = SEModule(b, rd_channels=int(round(in_chs * se_ratio)), act_layer=act_layer)
        else:
            self.se = nn.Identity()
        self.conv3 = ConvNormAct(b, out_chs, 1, apply_act=False, act_layer=act_layer)
        self.act3 = act_layer()
        if in_chs != out_chs or stride != 1:
            self.downsample = ConvNormAct(in_chs, out_chs, 1, stride=stride, apply_act=False, act_layer=act_layer)
        else:
            self.downsample = nn.Identity()

    def forward(self, x):
        shortcut = x
        x = self.conv1(x)
        x = self.conv2(x)
        x = self.se(x)
        x = self.conv3(x)
        x = x + self.downsample(shortcut)
        return self.act3(x)


class RegStage(nn.Module):
    def __init__(self, depth, in_chs, out_chs, stride, dilation, act_layer=nn.ReLU, norm_layer=None, **_):
        super().__init__()
        for i in range(depth):
            self.add_module(f"b{i + 1}", Bottleneck(in_chs if i == 0 else out_chs, out_chs,
                                                    stride=stride if i == 0 else 1, act_layer=act_layer,
                                                    norm_layer=norm_layer))

    def forward(self, x):
        for blk in self.children():
            x = blk(x)
        return x
