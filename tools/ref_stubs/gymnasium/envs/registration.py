REGISTRY = {}


def register(id, entry_point, **kwargs):
    REGISTRY[id] = entry_point
