"""Stand-in for `nicr_scene_analysis_datasets` (absent; test infrastructure only)."""


class ConcatDataset:
    pass
