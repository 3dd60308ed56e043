class OrientationDict(dict):
    pass


class SemanticLabelList(list):
    pass
