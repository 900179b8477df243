"""The YAML-selectable stage files (trajsde_b200/plugins) load through the reference's own loader (model_base_mix_sde.py:38-45) and
keep the reference classes' constructor and state_dict (dev container only: needs /root/reference)."""
import os
from importlib.machinery import SourceFileLoader

import pytest

from oracle import ref_runner as rr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(module_name, rel):
    return getattr(SourceFileLoader(module_name, os.path.join(ROOT, rel)).load_module(module_name), module_name)


@pytest.mark.skipif(not rr.reference_available(), reason="reference tree absent (GPU box)")
def test_plugin_stage_classes_keep_the_reference_state_dict():
    from test_boundary_cpu import REF_ENC_KW
    from trajsde_b200.stages import FusedDecoderMixin, FusedEncoderMixin
    m = rr.load_reference()
    Dec = _load('SDEDecoderFused', 'trajsde_b200/plugins/dec_hivt_nusargo_sde_fused.py')
    dec, ref = Dec(**rr.DEC_KW), m['dec'].SDEDecoder(**rr.DEC_KW)
    from models.decoders.dec_hivt_nusargo_sde import SDEDecoder as RefDecoder        # the class the plugin file subclasses
    assert isinstance(dec, RefDecoder) and type(dec).forward is FusedDecoderMixin.forward
    assert list(dec.state_dict()) == list(ref.state_dict())
    dec.load_state_dict(ref.state_dict())
    Enc = _load('LocalEncoderSDESepPara2Fused', 'trajsde_b200/plugins/enc_hivt_nusargo_sde_sep2_fused.py')
    enc, renc = Enc(**REF_ENC_KW), m['enc'].LocalEncoderSDESepPara2(**REF_ENC_KW)
    assert type(enc).forward is FusedEncoderMixin.forward and type(enc).forward_ood is FusedEncoderMixin.forward_ood
    assert list(enc.state_dict()) == list(renc.state_dict())
    enc.load_state_dict(renc.state_dict())
