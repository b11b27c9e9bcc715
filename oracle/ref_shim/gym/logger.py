def warn(*a, **k):
    pass


def info(*a, **k):
    pass


def error(*a, **k):
    pass
