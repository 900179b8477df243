"""Shim of torchsde.settings (torchsde==0.2.5, reference env.yml:293). TEST INFRASTRUCTURE ONLY."""


class _ContainerMeta(type):
    def all(cls):
        return sorted(getattr(cls, x) for x in dir(cls) if not x.startswith('__'))

    def __str__(cls):
        return str(cls.all())

    def __contains__(cls, item):
        return item in cls.all()


class METHODS(metaclass=_ContainerMeta):
    euler = 'euler'
    milstein = 'milstein'
    srk = 'srk'
    midpoint = 'midpoint'
    reversible_heun = 'reversible_heun'
    adjoint_reversible_heun = 'adjoint_reversible_heun'
    heun = 'heun'
    log_ode_midpoint = 'log_ode'
    euler_heun = 'euler_heun'


class NOISE_TYPES(metaclass=_ContainerMeta):
    general = 'general'
    diagonal = 'diagonal'
    scalar = 'scalar'
    additive = 'additive'


class SDE_TYPES(metaclass=_ContainerMeta):
    ito = 'ito'
    stratonovich = 'stratonovich'


class LEVY_AREA_APPROXIMATIONS(metaclass=_ContainerMeta):
    none = 'none'
    space_time = 'space-time'
    davie = 'davie'
    foster = 'foster'
