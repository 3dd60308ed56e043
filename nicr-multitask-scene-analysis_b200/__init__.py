# -*- coding: utf-8 -*-
"""B200-native panoptic post-processing + evaluation (drop-in for that path of
TUI-NICR/nicr-multitask-scene-analysis).  See DESIGN.md."""
__version__ = '0.1.0'
