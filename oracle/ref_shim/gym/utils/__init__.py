from . import seeding  # noqa: F401


class EzPickle:
    def __init__(self, *a, **k):
        pass
