# -*- coding: utf-8 -*-
from .fullres import fullres_key, fullres_shape, valid_region_and_fullres_shape
from .misc import partial_class
from .targets import InstanceTargetGenerator
from .panoptic_merge import deeplab_merge_batch, naive_merge_semantic_and_instance_batch

__all__ = ['InstanceTargetGenerator', 'fullres_key', 'fullres_shape', 'valid_region_and_fullres_shape', 'partial_class',
           'deeplab_merge_batch', 'naive_merge_semantic_and_instance_batch']
