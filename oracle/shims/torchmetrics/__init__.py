class Metric:
    def __init__(self, *a, **k):
        pass

    def add_state(self, name, default, dist_reduce_fx=None):
        setattr(self, name, default)
