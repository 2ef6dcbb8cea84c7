class PartitionSpec(tuple):
    def __new__(cls, *a):
        return super().__new__(cls, a)


class Mesh:
    def __init__(self, *a, **k):
        pass


class NamedSharding:
    def __init__(self, *a, **k):
        pass
