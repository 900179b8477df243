def softmax(*a, **k):
    raise NotImplementedError("stub")


def subgraph(*a, **k):
    raise NotImplementedError("stub")
