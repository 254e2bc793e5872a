"""``load_model`` / ``FineTuneModelPool`` with the surface of the reference's ``src/models/classifiers.py``
(:7-17, :19-111), mnasnet branch only (the resnet branch wraps torchvision models and is outside the hot
path).  ``forward`` runs the fused CUDA program: features -> (last BN+ReLU fused into)
global average pool -> dropout/FC head, all through libmnb200.so."""
import torch.nn as nn

from mnb200.engine import run_module
from models.mnasnet import Mnasnet

_HEADS = {                       # classifiers.py:56-89: (dropout p, hidden widths)
    '256': (0.5, (256,)),
    '512_256': (0.5, (512, 256)),
    '320': (0.2, ()),
    '512': (0.5, (512,)),
}


def load_model(arch='resnet18', pretrained=True):
    if arch.startswith('mnasnet'):
        model = Mnasnet(cut_channels_first=False)     # classifiers.py:13 (pretrained is ignored there too)
        print('Mnasnet initialized')
        return model
    raise ValueError("Finetuning not supported on this architecture yet (mnb200 lowers mnasnet only)")


class FineTuneModelPool(nn.Module):
    _mnb = 'net'

    def __init__(self, original_model, arch, num_classes, classifier_config):
        super().__init__()
        self.num_classes = num_classes
        if not arch.startswith('mnasnet'):
            raise ValueError("Finetuning not supported on this architecture yet")
        self.features = original_model.features
        final_feature_map = 320
        self.pooling = nn.Sequential(nn.AdaptiveAvgPool2d(1))
        self.modelName = 'mnasnet'
        if classifier_config not in _HEADS:
            raise ValueError("Finetuning not supported on this architecture yet")
        p, hidden = _HEADS[classifier_config]
        layers, width = [], final_feature_map
        for h in hidden:
            layers += [nn.Dropout(p), nn.Linear(width, h), nn.ReLU(inplace=True)]
            width = h
        layers += [nn.Dropout(p=p, inplace=(classifier_config == '320')), nn.Linear(width, num_classes)]
        self.classifier = nn.Sequential(*layers)
        self.mean = (0.485, 0.456, 0.406)
        self.std = (0.229, 0.224, 0.225)

    def freeze(self):
        print('Features frozen')
        for p in self.features.parameters():
            p.requires_grad = False

    def unfreeze(self):
        print('Features unfrozen')
        for p in self.features.parameters():
            p.requires_grad = True

    def forward(self, x):
        return run_module(self, x)
