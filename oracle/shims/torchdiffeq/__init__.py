def odeint(*a, **k):
    raise NotImplementedError("stub: torchdiffeq is imported by models/utils/ode_utils.py:7 but never called on the SDE path")
