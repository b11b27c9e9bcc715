from . import monitoring  # noqa: F401


class TimeLimit:
    pass
