import torch.nn as nn


class Linear(nn.Linear):
    pass
