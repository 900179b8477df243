"""Import shim standing in for torchsde==0.2.5 (reference env.yml:293; NOT installable here, no network).

TEST INFRASTRUCTURE ONLY: it exists so the reference's own hot-path files run verbatim from /root/reference in the
dev container (golden-vector generation, oracle validation). Nothing under trajsde_b200/ may import it.
"""
from ._brownian import BaseBrownian, BrownianInterval, FixedIncrements  # noqa: F401
from ._core.base_sde import BaseSDE, SDEIto, SDEStratonovich  # noqa: F401
from ._sdeint import sdeint, sdeint_adjoint  # noqa: F401
from . import settings, types  # noqa: F401
