"""cti_b200 -- B200-native (sm_100a) drop-ins for the compact-trilinear-interaction hot path of
aioz-ai/ICCV19_VQA-CTI: FCNet, TCNet, TriAttention, BCNet, BiAttention with the reference's
constructors, forward signatures and state_dict keys (reference src/{fc,tc,bc,attention}.py).

The directory is named ``iccv19_vqa-cti_b200`` (not an importable identifier); ``cti_b200.py`` at
the repository root loads it under the module name ``cti_b200``.

All compute goes through ``libcti_sm100.so`` (C ABI in ``include/cti_sm100.h``).  There is no CPU
or eager-PyTorch fallback: without the built library, or on non-CUDA tensors, calls raise.
"""
from . import _lib
from .attention import BiAttention, TriAttention
from .bc import BCNet
from .classifier import SimpleClassifier
from .dropin import install, uninstall
from .fc import FCNet, WNLinear
from .glimpse import glimpse_joint
from .graphs import GraphedStep, reset_caches
from .language_model import QuestionEmbedding
from .loader import FeatureBatch, FeatureStoreBF16, TeacherLogits, prime_features
from .loss_function import Distillation_Loss
from .optim import FusedClipAdamax
from .prepack import bind_grad_buffers, prepack, weight_norm_param_groups
from .tc import TCNet

__all__ = ["FCNet", "WNLinear", "TCNet", "TriAttention", "BCNet", "BiAttention", "SimpleClassifier", "QuestionEmbedding", "Distillation_Loss", "FusedClipAdamax", "prepack", "bind_grad_buffers", "weight_norm_param_groups", "FeatureStoreBF16", "FeatureBatch", "TeacherLogits", "prime_features", "install", "uninstall", "GraphedStep", "reset_caches", "glimpse_joint",
           "library_path", "version"]


def library_path() -> str:
    return _lib.LIB_PATH


def version() -> int:
    return _lib.load().cti_version()
