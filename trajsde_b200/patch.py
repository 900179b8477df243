"""Install the fused solve behind the reference's own call sites without editing the reference.

Stages are loaded by ``SourceFileLoader(...).load_module(module_name)`` (models/model_base_mix_sde.py:38-45), so the
decoder's ``sdeint`` and the encoder's ``sdeint_dual`` are module globals of the loaded stage modules; rebinding those
two names is the whole integration."""
from typing import Optional

from .solver import sdeint, sdeint_dual


def _globals_of(obj):
    fwd = getattr(type(obj), 'forward', None)
    g = getattr(fwd, '__globals__', None)
    if g is None:
        raise TypeError(f"cannot locate module globals of {type(obj).__name__}.forward")
    return g


def install(model=None, decoder=None, encoder=None) -> dict:
    """Rebind ``sdeint`` (decoder module) and ``sdeint_dual`` (encoder module).  Pass the LightningModule-style ``model``
    (with ``.decoder`` / ``.encoder``) or the stage modules directly.  Returns the originals for ``uninstall``."""
    decoder = decoder if decoder is not None else getattr(model, 'decoder', None)
    encoder = encoder if encoder is not None else getattr(model, 'encoder', None)
    saved = {}
    if decoder is not None:
        g = _globals_of(decoder)
        if 'sdeint' not in g:
            raise KeyError("decoder module has no global `sdeint` (expected `from torchsde import sdeint`)")
        saved['decoder'] = (g, 'sdeint', g['sdeint'])
        g['sdeint'] = sdeint
    if encoder is not None:
        g = _globals_of(encoder)
        if 'sdeint_dual' not in g:
            raise KeyError("encoder module has no global `sdeint_dual`")
        saved['encoder'] = (g, 'sdeint_dual', g['sdeint_dual'])
        g['sdeint_dual'] = sdeint_dual
    return saved


def uninstall(saved: Optional[dict]) -> None:
    for g, name, orig in (saved or {}).values():
        g[name] = orig
