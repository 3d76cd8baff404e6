"""context_transformer_b200 — B200-native detection hot path of Ze-Yang/Context-Transformer.

Public surface = the reference symbols that ``train.py`` / ``test.py`` import for this path:

  build_net, RFBNet, BasicConv, BasicRFB, BasicRFB_a      models/RFB_Net_vgg.py
  PriorBox, Detect                                        layers/functions/
  MultiBoxLoss_combined                                   layers/modules/multibox_loss_combined.py
  nms, gpu_nms, cpu_nms, cpu_soft_nms                     utils/nms_wrapper.py, utils/nms/
  match, decode, encode, jaccard, point_form              utils/box_utils.py
  init_reweight                                           train.py:252-286 (OBJ(Target) prototype initialisation)
  VOC_300, VOC_512, COCO_300, COCO_512                    data/config.py (prior-box constants)

plus the fused device-side post-processing ``DetectPost`` (test.py:133-161) and the multi-GPU
helper ``shard.ShardedDetector``.  All compute goes through ``libctx_b200.so`` (hand-written
sm_100a CUDA behind the C ABI of ``include/ctx_b200.h``); nothing here falls back to the CPU.
"""
from .config import COCO_300, COCO_512, MBOX, VOC_300, VOC_512, num_priors
from .prior_box import PriorBox
from .detection import BaseTransform, Detect, DetectionCollector, DetectPost, records_to_all_boxes
from .nms_wrapper import cpu_nms, cpu_soft_nms, gpu_nms, nms, nms_device, soft_nms
from .box_utils import decode, encode, hard_negative_rank, jaccard, match, match_batch, point_form
from .multibox_loss import MultiBoxLoss_combined
from .reweight import PrototypeAccumulator, init_reweight
from .rfb_net import BasicConv, BasicRFB, BasicRFB_a, RFBNet, build_net

__all__ = ['build_net', 'RFBNet', 'BasicConv', 'BasicRFB', 'BasicRFB_a', 'PriorBox', 'Detect', 'DetectPost',
           'records_to_all_boxes', 'DetectionCollector', 'BaseTransform', 'MultiBoxLoss_combined', 'nms', 'gpu_nms', 'cpu_nms', 'cpu_soft_nms', 'soft_nms',
           'nms_device', 'match', 'match_batch', 'decode', 'encode', 'jaccard', 'point_form', 'hard_negative_rank',
           'init_reweight', 'PrototypeAccumulator', 'VOC_300', 'VOC_512', 'COCO_300', 'COCO_512', 'MBOX', 'num_priors']
