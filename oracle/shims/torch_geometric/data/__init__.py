class Data(dict):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.__dict__.update(kw)


class Batch(Data):
    @staticmethod
    def from_data_list(lst):
        raise NotImplementedError("stub: PyG Batch is out of scope")


class Dataset:
    def __init__(self, *a, **k):
        pass


class DataLoader:
    def __init__(self, *a, **k):
        raise NotImplementedError("stub")
