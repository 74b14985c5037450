"""The embedding extractor that FEEDS the matching path: DeepLabv3+ with a ResNet-101 backbone
at output stride 16 (reference: networks/deeplab.py:27-52, networks/backbone/resnet.py, networks/aspp.py,
networks/decoder.py) and the ``IntVOS`` wrapper that owns it (networks/IntVOS.py:527-581).

This is the CALLER of the hot path, written from the published architecture in plain PyTorch (cuDNN convolutions):
out of scope as hand-written kernels, present so that BASELINE config 2 (``IntVOS.forward`` on one 480p frame triple)
runs end to end and so that the matching kernels -- in particular the local-matching numerics guard -- see embeddings
produced by the real topology.  Module and parameter names follow the reference's ``state_dict`` (``backbone.layer3.7.conv2``,
``aspp.aspp2.atrous_conv``, ``decoder.last_conv.4`` ...), so its checkpoints load unchanged; tests/test_deeplab.py checks
that against the reference's own classes when the reference tree is mounted.

Matching, both map memories and both segmentation heads of ``IntVOS`` run on the sm_100a kernels (``engine.py``).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import engine
from ..config import cfg as default_cfg
from .seghead import DynamicSegHead


def _bn(ch, momentum=0.1):
    return nn.BatchNorm2d(ch, momentum=momentum)


def _he_fan_out(module):
    """conv weights ~ N(0, 2/(k*k*out)) and unit batch norms: the initialisation the reference's backbone applies
    (networks/backbone/resnet.py:134-144)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.normal_(m.weight, 0.0, (2.0 / (m.kernel_size[0] * m.kernel_size[1] * m.out_channels)) ** 0.5)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


def _he_fan_in(module):
    """kaiming-normal convs and unit batch norms (ASPP / decoder: networks/aspp.py:78-90, networks/decoder.py:43-53)."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class Bottleneck(nn.Module):
    """1x1 reduce -> 3x3 (stride / dilation) -> 1x1 expand x4, residual add, ReLU."""
    expansion = 4

    def __init__(self, cin, width, stride=1, dilation=1, project=False):
        super().__init__()
        cout = width * self.expansion
        self.conv1 = nn.Conv2d(cin, width, 1, bias=False)
        self.bn1 = _bn(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=dilation, dilation=dilation, bias=False)
        self.bn2 = _bn(width)
        self.conv3 = nn.Conv2d(width, cout, 1, bias=False)
        self.bn3 = _bn(cout)
        self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), _bn(cout)) if project else None

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return F.relu(y + (x if self.downsample is None else self.downsample(x)))


class ResNet101(nn.Module):
    """ResNet-101 trunk [3, 4, 23, 3]; the last stage is the multi-grid unit with dilations (1, 2, 4) x the stage dilation.
    Returns (stride-`output_stride` features [2048 ch], stride-4 low-level features [256 ch])."""

    def __init__(self, output_stride=16):
        super().__init__()
        if output_stride == 16:
            strides, dilations = (1, 2, 2, 1), (1, 1, 1, 2)
        elif output_stride == 8:
            strides, dilations = (1, 2, 1, 1), (1, 1, 2, 4)
        else:
            raise NotImplementedError("output_stride must be 8 or 16")
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = _bn(64)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        cin = 64
        stages = []
        for width, depth, stride, dil in zip((64, 128, 256), (3, 4, 23), strides[:3], dilations[:3]):
            blocks = [Bottleneck(cin, width, stride, dil, project=True)]
            cin = width * Bottleneck.expansion
            blocks += [Bottleneck(cin, width, 1, dil) for _ in range(depth - 1)]
            stages.append(nn.Sequential(*blocks))
        self.layer1, self.layer2, self.layer3 = stages
        grid = (1, 2, 4)
        blocks = [Bottleneck(cin, 512, strides[3], grid[0] * dilations[3], project=True)]
        blocks += [Bottleneck(2048, 512, 1, g * dilations[3]) for g in grid[1:]]
        self.layer4 = nn.Sequential(*blocks)
        _he_fan_out(self)

    def forward(self, x):
        x = self.maxpool(F.relu(self.bn1(self.conv1(x))))
        low = self.layer1(x)
        return self.layer4(self.layer3(self.layer2(low))), low


class _AtrousBranch(nn.Module):
    def __init__(self, cin, cout, kernel, dilation):
        super().__init__()
        self.atrous_conv = nn.Conv2d(cin, cout, kernel, padding=0 if kernel == 1 else dilation, dilation=dilation, bias=False)
        self.bn = _bn(cout)

    def forward(self, x):
        return F.relu(self.bn(self.atrous_conv(x)))


class ASPP(nn.Module):
    """Four atrous branches (1x1 and 3x3 at rates 6/12/18 for stride 16) + image pooling -> concat -> 1x1 -> dropout."""

    def __init__(self, cin=2048, output_stride=16):
        super().__init__()
        rates = {16: (1, 6, 12, 18), 8: (1, 12, 24, 36)}[output_stride]
        self.aspp1 = _AtrousBranch(cin, 256, 1, rates[0])
        self.aspp2 = _AtrousBranch(cin, 256, 3, rates[1])
        self.aspp3 = _AtrousBranch(cin, 256, 3, rates[2])
        self.aspp4 = _AtrousBranch(cin, 256, 3, rates[3])
        self.global_avg_pool = nn.Sequential(nn.AdaptiveAvgPool2d((1, 1)), nn.Conv2d(cin, 256, 1, bias=False), _bn(256), nn.ReLU())
        self.conv1 = nn.Conv2d(1280, 256, 1, bias=False)
        self.bn1 = _bn(256)
        self.dropout = nn.Dropout(0.1)
        _he_fan_in(self)

    def forward(self, x):
        pooled = F.interpolate(self.global_avg_pool(x), size=x.shape[2:], mode="bilinear", align_corners=True)
        y = torch.cat((self.aspp1(x), self.aspp2(x), self.aspp3(x), self.aspp4(x), pooled), 1)
        return self.dropout(F.relu(self.bn1(self.conv1(y))))


class Decoder(nn.Module):
    """Low-level features -> 48 ch, concat with the upsampled ASPP output, two 3x3 convs to 256 ch.  The reference strips the
    classifier (``cls_conv`` / ``upsample4`` replaced by empty Sequentials, IntVOS.py:531-532): the 256-channel map at
    stride 4 is the output.  ``last_conv`` keeps the reference's indices (0,1,4,5) so checkpoints load."""

    def __init__(self, low_level_ch=256):
        super().__init__()
        self.conv1 = nn.Conv2d(low_level_ch, 48, 1, bias=False)
        self.bn1 = _bn(48)
        self.last_conv = nn.Sequential(nn.Conv2d(304, 256, 3, padding=1, bias=False), _bn(256), nn.ReLU(True), nn.Sequential(),
                                       nn.Conv2d(256, 256, 3, padding=1, bias=False), _bn(256), nn.ReLU(True), nn.Sequential())
        _he_fan_in(self)

    def forward(self, x, low):
        low = F.relu(self.bn1(self.conv1(low)))
        x = F.interpolate(x, size=low.shape[2:], mode="bilinear", align_corners=True)
        return self.last_conv(torch.cat((x, low), 1))


class DeepLab(nn.Module):
    """``DeepLab(backbone='resnet', output_stride=16)`` of the reference (networks/deeplab.py:27-52) without its classifier:
    image ``[B,3,H,W]`` -> ``[B,256,H/4,W/4]``.  Only the ResNet-101 backbone the reference's scripts use is provided."""

    def __init__(self, backbone="resnet", output_stride=16, num_classes=21, sync_bn=True, freeze_bn=False):
        super().__init__()
        if backbone != "resnet":
            raise NotImplementedError("only the ResNet-101 backbone (the one MANet's scripts build) is provided")
        del num_classes, sync_bn
        self.backbone = ResNet101(output_stride)
        self.aspp = ASPP(2048, output_stride)
        self.decoder = Decoder(256)
        if freeze_bn:
            self.freeze_bn()

    def forward(self, x):
        feats, low = self.backbone(x)
        return self.decoder(self.aspp(feats), low)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()


class IntVOS(nn.Module):
    """The reference's top-level module (networks/IntVOS.py:527-581, 583-764) with the same constructor, attribute names and
    method signatures: ``extract_feature`` (DeepLab + semantic embedding, torch/cuDNN), ``forward`` / ``prop_seghead`` /
    ``int_seghead`` (matching, map memories and both heads on the sm_100a kernels through ``engine``).

    ``semantic_embedding`` = depthwise 3x3 -> BN -> ReLU -> 1x1 to ``MODEL_SEMANTIC_EMBEDDING_DIM`` -> BN -> ReLU
    (IntVOS.py:534-543); its modules are also registered under their own names as the reference does."""

    def __init__(self, cfg=None, feature_extracter=None):
        super().__init__()
        cfg = default_cfg if cfg is None else cfg
        aspp_dim = getattr(cfg, "MODEL_ASPP_OUTDIM", 256)
        emb_dim = cfg.MODEL_SEMANTIC_EMBEDDING_DIM
        mom = getattr(cfg, "TRAIN_BN_MOM", 0.0003)
        self.feature_extracter = DeepLab() if feature_extracter is None else feature_extracter
        self.seperate_conv = nn.Conv2d(aspp_dim, aspp_dim, 3, padding=1, groups=aspp_dim)
        self.bn1 = _bn(aspp_dim, mom)
        self.relu1 = nn.ReLU(True)
        self.embedding_conv = nn.Conv2d(aspp_dim, emb_dim, 1)
        self.relu2 = nn.ReLU(True)
        self.bn2 = _bn(emb_dim, mom)
        self.semantic_embedding = nn.Sequential(self.seperate_conv, self.bn1, self.relu1, self.embedding_conv, self.bn2, self.relu2)
        for m in self.semantic_embedding:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self.dynamic_seghead = DynamicSegHead(in_dim=emb_dim + 3)
        if getattr(cfg, "MODEL_USEIntSeg", False):
            raise NotImplementedError("MODEL_USEIntSeg=True (the dense IntSegHead, IntVOS.py:463-486) is not provided; the "
                                      "reference's default (config.py:52) is the DynamicSegHead interaction head")
        self.inter_seghead = DynamicSegHead(in_dim=emb_dim + 2)

    def extract_feature(self, x):
        return self.semantic_embedding(self.feature_extracter(x))

    def forward(self, x=None, ref_scribble_label=None, previous_frame_mask=None, normalize_nearest_neighbor_distances=True,
                use_local_map=True, seq_names=None, gt_ids=None, k_nearest_neighbors=1, global_map_tmp_dic=None,
                local_map_dics=None, interaction_num=None, start_annotated_frame=None, frame_num=None):
        """``x``: ``[3*bs,3,H,W]`` = reference, previous and current frames stacked along the batch (IntVOS.py:556-575).
        Returns what ``prop_seghead`` returns for the given memories."""
        emb = self.extract_feature(x)
        ref, prev, cur = torch.split(emb, emb.size(0) // 3, dim=0)
        out = self.prop_seghead(ref, prev, cur, ref_scribble_label, previous_frame_mask, normalize_nearest_neighbor_distances,
                                use_local_map, seq_names, gt_ids, k_nearest_neighbors, global_map_tmp_dic, local_map_dics,
                                interaction_num, start_annotated_frame, frame_num, self.dynamic_seghead)
        if global_map_tmp_dic is None:
            return out
        return out[0], out[1]            # the reference's forward drops the local-map dicts (IntVOS.py:569-575)

    def prop_seghead(self, ref_frame_embedding=None, previous_frame_embedding=None, current_frame_embedding=None,
                     ref_scribble_label=None, previous_frame_mask=None, normalize_nearest_neighbor_distances=True,
                     use_local_map=True, seq_names=None, gt_ids=None, k_nearest_neighbors=1, global_map_tmp_dic=None,
                     local_map_dics=None, interaction_num=None, start_annotated_frame=None, frame_num=None,
                     dynamic_seghead=None):
        return engine.prop_seghead(ref_frame_embedding, previous_frame_embedding, current_frame_embedding, ref_scribble_label,
                                   previous_frame_mask, normalize_nearest_neighbor_distances, use_local_map, seq_names, gt_ids,
                                   k_nearest_neighbors, global_map_tmp_dic, local_map_dics, interaction_num,
                                   start_annotated_frame, frame_num, dynamic_seghead)

    def int_seghead(self, ref_frame_embedding=None, ref_scribble_label=None, prev_round_label=None,
                    normalize_nearest_neighbor_distances=True, global_map_tmp_dic=None, local_map_dics=None,
                    interaction_num=None, seq_names=None, gt_ids=None, k_nearest_neighbors=1, frame_num=None, first_inter=True):
        return engine.int_seghead(ref_frame_embedding, ref_scribble_label, prev_round_label, normalize_nearest_neighbor_distances,
                                  global_map_tmp_dic, local_map_dics, interaction_num, seq_names, gt_ids, k_nearest_neighbors,
                                  frame_num, first_inter, self.inter_seghead)
