"""YAML-selectable decoder stage: the reference's ``SDEDecoder`` (models/decoders/dec_hivt_nusargo_sde.py:15-105) with its forward taken
from ``FusedDecoderMixin`` — same constructor, same parameters / state_dict, fused prologue + solve + heads."""
from models.decoders.dec_hivt_nusargo_sde import SDEDecoder          # the reference repository must be on sys.path

from trajsde_b200.stages import FusedDecoderMixin


class SDEDecoderFused(FusedDecoderMixin, SDEDecoder):
    pass
