class EnvSpec:
    def __init__(self, id, **kw):
        self.id = id


def register(id, **kw):
    pass
