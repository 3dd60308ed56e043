import importlib.abc


class DependencyImportHook(importlib.abc.MetaPathFinder):
    def __init__(self, *args, **kwargs):
        pass

    def find_spec(self, fullname, path, target=None):
        return None


def is_package_available(*args, **kwargs):
    return True
