"""trajsde_b200 — B200-native (sm_100a) fused Euler–Maruyama solve behind TrajSDE's ``sdeint`` / ``sdeint_dual`` call sites.

Public API mirrors the reference's operator interface for this path:
    sdeint, sdeint_dual          drop-in solver calls (trajsde_b200/solver.py)
    install / uninstall          rebind the reference's module globals (trajsde_b200/patch.py)
    euler_schedule               the reference's float32 step schedule (trajsde_b200/schedule.py)
    manual_seed, set_default_mode, set_device_seed (Philox key in device memory: CUDA-graph replays with fresh noise)
The CUDA library (trajsde_b200/lib/libtrajsde_b200.so, built by `python -m trajsde_b200.build`) is mandatory: there is
no CPU, eager-PyTorch or Triton fallback.
"""
from .schedule import EulerSchedule, encoder_schedule, encoder_time_pairs, euler_schedule  # noqa: F401
from .solver import get_default_mode, manual_seed, sdeint, sdeint_dual, set_default_mode, set_device_seed  # noqa: F401
from .patch import install, uninstall  # noqa: F401

__version__ = '0.1.0'
