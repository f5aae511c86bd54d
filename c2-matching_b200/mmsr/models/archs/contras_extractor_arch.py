"""Contrastive feature extractor: two independent VGG16[:conv3_1] trunks, no ReLU after
conv3_1 (mmsr/models/archs/contras_extractor_arch.py:8-59).  Keys:
`feature_extraction_image{1,2}.model.conv*_*.{weight,bias}` + `mean`/`std` buffers, so the
reference's `feature_extraction.pth` loads strictly."""
import torch
from torch import nn

from .vgg_arch import build_trunk, run_trunk


class ContrasExtractorLayer(nn.Module):

    def __init__(self):
        super().__init__()
        self.model = build_trunk('vgg16', 'conv3_1')
        self.register_buffer('mean', torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1))

    def forward(self, batch):
        return run_trunk(self.model, (batch - self.mean) / self.std)[0]


class ContrasExtractorSep(nn.Module):

    def __init__(self):
        super().__init__()
        self.feature_extraction_image1 = ContrasExtractorLayer()
        self.feature_extraction_image2 = ContrasExtractorLayer()

    def forward(self, image1, image2):
        return {'dense_features1': self.feature_extraction_image1(image1),
                'dense_features2': self.feature_extraction_image2(image2)}
